/* caffe_b200_c.h -- C binding of the C++ Caffe host (libcaffe_b200.so) used by the Python `caffe`
 * shim.  The reference binds the same classes with boost.python (python/caffe/_caffe.cpp:73-347);
 * each function names the member it forwards to.  Return 0 = ok, non-zero = a host CHECK failed
 * (message from caffe_last_error()); pointer-returning functions return NULL on failure. */
#ifndef CAFFE_B200_C_H_
#define CAFFE_B200_C_H_
#ifdef __cplusplus
extern "C" {
#endif

const char* caffe_last_error(void);
int caffe_set_mode(int gpu);                 /* Caffe::set_mode  (_caffe.cpp:38-39) */
int caffe_get_mode(void);
int caffe_set_device(int id);                /* Caffe::SetDevice (_caffe.cpp:221) */
int caffe_device_count(void);
void caffe_set_log_level(int level);         /* 0 = INFO ... 3 = FATAL only */
void* caffe_stream(void);                    /* cudaStream_t the calling thread's forwards run on */
int caffe_sync(void);

void* caffe_net_create(const char* prototxt_path, int phase /*0 TRAIN, 1 TEST*/);   /* Net(file, phase) net.cpp:31-37 */
void* caffe_net_create_from_string(const char* prototxt_text, int phase);
void caffe_net_destroy(void* net);
int caffe_net_copy_trained_from(void* net, const char* caffemodel);   /* Net::CopyTrainedLayersFrom net.cpp:843-858 */
int caffe_net_save(void* net, const char* caffemodel);                /* Net::ToProto + WriteProtoToBinaryFile (_caffe.cpp:120-124) */
int caffe_net_forward(void* net);                                     /* Net::ForwardPrefilled */
int caffe_net_forward_from_to(void* net, int start, int end);         /* Net::ForwardFromTo net.cpp:565-581 */
int caffe_net_reshape(void* net);                                     /* Net::Reshape net.cpp:798-802 */
const char* caffe_net_name(void* net);
int caffe_net_num_blobs(void* net);
const char* caffe_net_blob_name(void* net, int i);
int caffe_net_num_layers(void* net);
const char* caffe_net_layer_name(void* net, int i);
const char* caffe_net_layer_type(void* net, int i);
int caffe_net_layer_num_blobs(void* net, int i);
int caffe_net_layer_num_bottoms(void* net, int i);
int caffe_net_layer_bottom_id(void* net, int i, int j);
int caffe_net_layer_num_tops(void* net, int i);
int caffe_net_layer_top_id(void* net, int i, int j);
int caffe_net_num_inputs(void* net);
int caffe_net_input_index(void* net, int i);
int caffe_net_num_outputs(void* net);
int caffe_net_output_index(void* net, int i);
int caffe_net_layer_weights_changed(void* net, int i);   /* after writing a layer blob through its host pointer */

/* Blob handles hold a shared_ptr: the memory outlives the Net (python/caffe/test/test_net.py:48-60). */
void* caffe_net_blob(void* net, int i);
void* caffe_net_layer_blob(void* net, int layer, int j);
void caffe_blob_release(void* blob);
int caffe_blob_num_axes(void* blob);
int caffe_blob_shape(void* blob, int axis);
int caffe_blob_count(void* blob);
int caffe_blob_reshape(void* blob, int naxes, const int* dims);        /* Blob::Reshape (_caffe.cpp:181-193) */
float* caffe_blob_mutable_cpu_data(void* blob);                        /* zero-copy view base (_caffe.cpp:159-179) */
const float* caffe_blob_cpu_data(void* blob);
float* caffe_blob_mutable_cpu_diff(void* blob);
const float* caffe_blob_gpu_data(void* blob);
float* caffe_blob_mutable_gpu_data(void* blob);
float* caffe_blob_overwrite_gpu_data(void* blob);                      /* device pointer for a writer of every element: no upload first */
/* SyncedMemory::head() of the blob's data (include/caffe/syncedmem.hpp:65-66): 0 UNINITIALIZED, 1 HEAD_AT_CPU, 2 HEAD_AT_GPU,
 * 3 SYNCED; -1 + caffe_last_error() for a blob without memory.  Lets callers check the state machine the reference pins in
 * src/caffe/test/test_syncedmem.cpp:16-125. */
int caffe_blob_data_head(void* blob);

/* B200 extensions */
int caffe_net_set_fusion(void* net, int on);
int caffe_net_materialize_intermediates(void* net, int on);
/* Net outputs the caller will not read ("next_pred": python/pose/estimate_pose.py:231 reads prob and loc_pred only): the fused
 * plan drops their heads from the merged head GEMMs and leaves those blobs unwritten (caffe_net_blob_fresh() == 0).  "" clears. */
int caffe_net_set_skipped_outputs(void* net, const char* comma_separated_blob_names);
int caffe_net_fused_last_forward(void* net);
const char* caffe_net_fusion_diagnostic(void* net);
long long caffe_net_last_forward_launches(void* net);
/* Per-step device timing of the fused plan (the `caffe time` idiom, tools/caffe.cpp:302-388).
 * names receives "Type name\n" per step; ms/flops/bytes the last run's duration and algorithmic work. */
int caffe_net_set_step_timing(void* net, int on);
int caffe_net_num_steps(void* net);
int caffe_net_step_info(void* net, char* names, int names_cap, double* ms, double* flops, double* bytes, int max_steps);
long long caffe_net_arena_bytes(void* net);
/* NetParameter.debug_info (net.cpp:648-735): a forward with it on runs layer by layer and records mean|x| of every top blob
 * (the reference only logs them).  caffe_net_debug_info copies the last forward's records: names = "layer blob\n" per record;
 * returns the record count or -1. */
int caffe_net_set_debug_info(void* net, int on);
int caffe_net_debug_info(void* net, char* names, int names_cap, double* mean_abs, int max_records);
/* 1 when blob i holds the value of the last forward; 0 for an intermediate the fused plan did not write (the reference fills
 * every blob each forward, net.cpp:565-581 -- the shim raises instead of returning stale data). */
int caffe_net_blob_fresh(void* net, int i);
/* Plans the fused execution for the net's CURRENT input shapes without touching a device (works in CPU mode): writes the plan's
 * description (steps, L2-resident segments, launch groups, arena size) into out; returns 0, or 1 + caffe_last_error() when
 * the topology does not fuse or the planner's own schedule check fails. */
int caffe_net_describe_plan(void* net, char* out, int out_cap);
long long caffe_net_weight_bytes(void* net);
int caffe_insert_splits_text(const char* prototxt_text, char* out, int out_cap);   /* InsertSplits, insert_splits.cpp:12 */

#ifdef __cplusplus
}
#endif
#endif
