/* deepcut_b200.h -- C ABI of the B200-native DeeperCut forward kernels.
 *
 * This is the seam between the kept C++ Caffe host (Net/Layer/Blob, caffe_host/) and the
 * hand-written sm_100a CUDA.  The reference has no C ABI; each entry point names the
 * reference interface (file:line under /root/reference) whose GPU work it replaces.
 *
 * Conventions: plain C types only; device pointers are caller-owned; `stream` is a
 * cudaStream_t passed as void*; every function returns 0 on success and a non-zero code on
 * failure with a message available from dc_last_error() (never aborts -- the C++ Forward_gpu
 * wrappers turn non-zero into CHECK failures, matching the reference's glog convention,
 * include/caffe/util/device_alternate.hpp:48-66).  There is no CPU fallback: compute entry
 * points fail with DC_ERR_NO_DEVICE when no sm_100 GPU is present.
 *
 * Activation layout between kernels ("split NHWC"): fp16 [2][N][H][W][C]; plane 0 = hi =
 * fp16(x), plane 1 = lo = fp16(x - hi).  Blob-facing tensors are fp32 NCHW as in
 * include/caffe/blob.hpp:153-164.
 */
#ifndef DEEPCUT_B200_H_
#define DEEPCUT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DC_OK 0
#define DC_ERR_INVALID 1
#define DC_ERR_CUDA 2
#define DC_ERR_NO_DEVICE 3
#define DC_ERR_UNSUPPORTED 4

/* ---- library ---------------------------------------------------------------------- */
int dc_version(void);
const char* dc_last_error(void);
/* Number of usable sm_100 devices (0 on a CPU-only host; never fails). */
int dc_device_count(void);
/* Binds the calling thread to `device`, resolves the driver's tensor-map encoder, raises the
 * kernels' dynamic shared-memory limits.  Replaces Caffe::SetDevice (src/caffe/common.cpp:140-158). */
int dc_init(int device);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
long long dc_launch_count(void);

/* ---- device memory / streams (what SyncedMemory and the activation arena sit on) ------ */
/* Replaces the cudaMalloc / cudaMallocHost / cudaMemcpy calls of SyncedMemory
 * (src/caffe/syncedmem.cpp:7-77, include/caffe/syncedmem.hpp:15-36) and caffe_gpu_memcpy
 * (src/caffe/util/math_functions.cu:77-81).  Copies are asynchronous on `stream`. */
#define DC_H2D 1
#define DC_D2H 2
#define DC_D2D 3
int dc_malloc(void** ptr, size_t bytes);
int dc_free(void* ptr);
int dc_malloc_host(void** ptr, size_t bytes);      /* pinned */
int dc_free_host(void* ptr);
int dc_memcpy_async(void* dst, const void* src, size_t bytes, int kind, void* stream);
int dc_memset_async(void* ptr, int value, size_t bytes, void* stream);
int dc_stream_create(void** stream);
int dc_stream_destroy(void* stream);
int dc_stream_sync(void* stream);
int dc_device_sync(void);
int dc_mem_info(size_t* free_bytes, size_t* total_bytes);
/* Timing with CUDA events on `stream` (Timer, src/caffe/util/benchmark.cpp:8-118). */
int dc_event_create(void** event);
int dc_event_destroy(void* event);
int dc_event_record(void* event, void* stream);
int dc_event_elapsed_ms(void* start, void* stop, float* ms);   /* synchronises on `stop` */
int dc_event_sync(void* event);                                /* host waits until the work recorded before `event` is done */

/* CUDA-graph capture of a launch sequence issued through this ABI on `stream` (per input shape: the fused
 * plan replays ~164 launches, incl. their cluster / programmatic-dependent-launch attributes, from one
 * cudaGraphLaunch instead of re-encoding tensor maps and re-launching every kernel). */
int dc_graph_begin(void* stream);
int dc_graph_end(void* stream, void** graph_exec);      /* ends capture and instantiates */
int dc_graph_launch(void* graph_exec, void* stream);
int dc_graph_destroy(void* graph_exec);

/* ---- load-time weight transforms (host side, no GPU needed) ------------------------- */
/* y = (x - mean*sf) / sqrt(var*sf + eps) * gamma + beta  ==  a*x + b with sf = (factor==0 ? 0 : 1/factor)
 * BatchNormLayer inference branch (src/caffe/layers/batch_norm_layer.cpp:86-93,137-149) folded with
 * ScaleLayer (+bias) (scale_layer.cpp:120-133, bias_layer.cpp:72-88).  gamma/beta may be NULL (=1/0). */
int dc_fold_bn_scale(const float* mean_sum, const float* var_sum, float factor, float eps, const float* gamma,
                     const float* beta, int channels, float* a_out, float* b_out);

/* Rows of the packed weight matrix after padding Cout to the conv kernel's N tile. */
int dc_packed_rows(int cout);
/* N tile (out-channels per CTA tile) the packed weight rows are padded to for `cout` output channels (64 or 128).  A launch
 * whose grid would be under-filled may run 64-channel tiles over 128-padded rows (same results bitwise). */
int dc_tile_n(int cout);
/* Packs a Caffe Convolution weight blob W[cout][cin][kh][kw] (base_conv_layer.cpp:135-140) into the
 * K-major split-fp16 matrix the implicit GEMM reads: packed[2][rows][K], K = (p*kw+q)*cin + ci,
 * rows = dc_packed_rows(cout) (zero padded).  Each row is multiplied by a power of two so its
 * largest |w| lands in [2^9, 2^10) (keeps the lo plane out of fp16 subnormals); rowscale[r]
 * receives the exact inverse (1 for the zero padding rows). */
int dc_pack_conv_weight(const float* w, int cout, int cin, int kh, int kw, uint16_t* packed, float* rowscale);
/* Same for a Deconvolution blob W[cin][cout][kh][kw] (reverse_dimensions, base_conv_layer.cpp:125-131):
 * GEMM row = co*kh*kw + p*kw + q, K = ci; rows = dc_packed_rows(cout*kh*kw). */
int dc_pack_deconv_weight(const float* w, int cin, int cout, int kh, int kw, uint16_t* packed, float* rowscale);
/* conv1 filter bank W[64][3][7][7] -> fp32 [147][64] (k = (ci*7+p)*7+q major, co minor). */
int dc_pack_conv1_weight(const float* w, float* packed);

/* ---- fused convolution (tcgen05 implicit GEMM) ------------------------------------- */
typedef struct dc_conv_args {
  /* input: split NHWC [2][n][h][w][cin], cin % 64 == 0 */
  const void* x;
  int n, h, w, cin;
  /* filter geometry (stride: last field) */
  int cout, kh, kw, pad, dilation;
  const void* w_packed;      /* device copy of dc_pack_*_weight output, rows = dc_packed_rows(cout) */
  const float* scale;        /* device [dc_packed_rows(cout)]: folded a[c] * rowscale[c] */
  const float* shift;        /* device [dc_packed_rows(cout)]: folded b[c] (or bias) */
  const void* residual;      /* split NHWC, output geometry, or NULL (Eltwise SUM shortcut) */
  int relu;
  int out_f32_rows;          /* 0: split NHWC out [2][n][ho][wo][cout]; 1: fp32 rows out[pixel][ldc];
                              * 2: fp32 channel-major out[channel][ldc] (1x1 only; the GEMM runs with A and B
                              *    swapped so its rows are Caffe's col-buffer rows, base_conv_layer.cpp:358-365) */
  int ldc;                   /* row stride in floats: mode 1 >= dc_packed_rows(cout); mode 2 >= n*h*w, multiple of 4 */
  void* out;
  int stride;                /* 0 or 1: unit stride.  > 1 (<= 8): 1x1 pad-0 convolutions with split output only; the A tensor map
                              * traverses W and H with this element stride, which is what im2col does for the reference's
                              * strided 1x1 convs (res3a/res4a branch1 + branch2a; im2col.cu:8-39) */
  void* splitk_workspace;    /* NULL, or device scratch of >= dc_splitk_workspace_bytes() for the split-K clusters of under-filled
                              * launches (see dc_set_split_k); without it such launches simply do not split.  No initialisation
                              * needed; must not be shared by launches that can run concurrently (one per stream). */
  size_t splitk_workspace_bytes;
  /* Sub-batch launches (the L2-resident chunked schedule): x / out / residual may point at image i0 of a
   * larger [2][N][..] tensor, n being the images this launch covers; the distance between the hi and the lo plane is then
   * the FULL tensor's plane, given here in elements.  0 = dense (n*h*w*c of the tensor itself). */
  long long x_plane, out_plane, residual_plane;
  /* 1: load the weight tiles with the L2 evict_last priority (they are re-read by every CTA for each of its pixel tiles while the
   * activations stream through L2).  Ignored for launches with fewer pixel tiles than SMs, which read each weight tile once. */
  int weights_evict_last;
} dc_conv_args;
/* Replaces ConvolutionLayer::Forward_gpu (src/caffe/layers/conv_layer.cu:8-24) =
 * im2col_gpu (util/im2col.cu:8-62) + cublasSgemm (util/math_functions.cu:13-27) per image, and the
 * following BatchNormLayer/ScaleLayer/ReLULayer/EltwiseLayer::Forward_gpu passes
 * (batch_norm_layer.cu:10-90, scale_layer.cu:30-56, relu_layer.cu:17-32, eltwise_layer.cu:47-53).
 * Also serves DeconvolutionLayer's GEMM (base_conv_layer.cpp:351-367) with kh=kw=1 and a
 * dc_pack_deconv_weight matrix (out_f32_rows = 1). */
int dc_conv_forward(const dc_conv_args* args, void* stream);
/* Latency regime (one image; the reference's demo runs batch 1, python/pose/estimate_pose.py:224-243): when a layer has
 * fewer 128-pixel x N-channel work units than SMs, dc_conv_forward first halves the channel tile (bitwise-neutral) and
 * then, given args->splitk_workspace, shares each unit's K loop among a cluster of up to `max_split` CTAs (split-K; the
 * partial tiles meet in the L2-resident workspace and are summed in rank order: deterministic, but the fp32 summation
 * order differs from the unsplit kernel).
 * max_split in {1, 2, 4}; 1 = never split (bitwise batch-independent results at any size).  Default 4 (env DC_SPLIT_K).
 * Process-wide; takes effect for launches (and CUDA-graph captures) made after the call. */
int dc_set_split_k(int max_split);
int dc_get_split_k(void);
/* SMs the persistent convolution grids leave free for kernels running beside a forward (NCCL's, during a multi-GPU batch
 * exchange): a persistent one-CTA-per-SM kernel with a static tile schedule cannot share an SM without stretching to twice its time.
 * Process-wide; takes effect for launches (and CUDA-graph captures) made after the call.  Default 0 (env DC_RESERVED_SMS). */
int dc_set_reserved_sms(int n);
int dc_get_reserved_sms(void);
/* Upper bound of the scratch any split launch needs on this device (SM count x one 128 x 128 fp32 tile). */
size_t dc_splitk_workspace_bytes(void);
/* A K loop is shared only from `min_ksteps` 64-channel K-steps on (taps * cin / 64; env DC_SPLIT_K_MIN_STEPS): the exchange
 * costs about as much as a few K-steps (profiles/r1_microbench_latency.txt). */
int dc_set_split_k_min_steps(int min_ksteps);
int dc_get_split_k_min_steps(void);

/* ---- HBM-bound kernels --------------------------------------------------------------- */
/* conv1 7x7/2 pad 3 (3->64) + folded BN/Scale + ReLU.  x: fp32 NCHW [n][3][h][w] (the `data` blob);
 * w147x64: device dc_pack_conv1_weight output; out: split NHWC [2][n][ho][wo][64].
 * Replaces conv_layer.cu:8-24 (K=147 im2col+SGEMM) + bn/scale/relu for layer conv1. */
int dc_conv1_forward(const float* x, int n, int h, int w, const float* w147x64, const float* scale,
                     const float* shift, void* out, void* stream);
/* Tensor-core stem: the same layer as dc_conv1_forward, computed as a 4x4 stride-1 convolution over the
 * 2x2 space-to-depth image on the tcgen05 kernel.  workspace: device scratch of dc_conv1_tc_workspace_bytes()
 * bytes; w_packed/scale/shift from dc_pack_conv1_tc_weight (host) uploaded by the caller
 * (scale[c] = folded a[c] * rowscale[c]). */
size_t dc_conv1_tc_workspace_bytes(int n, int h, int w);
/* W[64][3][7][7] -> split fp16 [2][64][256], K = p*64 + q*16 + (py*2+px)*3 + ci (zero where the 4x4x16
 * window has no 7x7 tap), rows power-of-two scaled like dc_pack_conv_weight. */
int dc_pack_conv1_tc_weight(const float* w, uint16_t* packed, float* rowscale);
int dc_conv1_tc_forward(const float* x, int n, int h, int w, const void* w_packed, const float* scale,
                        const float* shift, void* workspace, void* out, void* stream);
/* MAX pool, pad 0, ceil-mode (PoolingLayer::Forward_gpu, pooling_layer.cu:10-47,158-180). */
int dc_maxpool_forward(const void* x, int n, int h, int w, int c, int kernel, int stride, void* out, void* stream);
int dc_pool_out_size(int size, int kernel, int stride);
/* out[n,y,x,:] = x[n,s*y,s*x,:] (what im2col does for a strided 1x1 conv, im2col.cu:8-39). */
int dc_subsample_forward(const void* x, int n, int h, int w, int c, int stride, void* out, void* stream);
/* Head finish: col2im of the 3x3/2 deconvolution (im2col.cu:246-305) + Crop to (ho,wo) at offset 0
 * (crop_layer.cu:9-38) + Eltwise SUM with the 1x1 skip head (eltwise_layer.cu:47-53) [+ Sigmoid,
 * sigmoid_layer.cu:8-24].  Both inputs are channel-major fp32 (dc_conv_forward out_f32_rows = 2):
 * col row col_row0 + co*9 + p*3 + q, ldcol >= n*h*w; skip row skip_row0 + co, ldskip >= n*ho*wo;
 * out: fp32 NCHW [n][cout][ho][wo]. */
int dc_head_finish(const float* col, long long ldcol, int col_row0, const float* skip, long long ldskip, int skip_row0,
                   float* out, int n, int cout, int h, int w, int ho, int wo, int sigmoid, void* stream);
/* Pose read-out of the demo on the device (python/pose/estimate_pose.py:131-143, _pose_from_mats):
 * per image n and joint j: arg-max of prob[n][j] (first maximum, row-major), refined by loc[n][2j..2j+1] at
 * that cell.  out: fp32 [n][5][joints] = {x, y, confidence, offset_y, offset_x} exactly as the demo lays them out. */
int dc_pose_from_maps(const float* prob, const float* loc, int n, int joints, int h, int w, float stride,
                      float locref_scale, float scale, float* out, void* stream);
/* Demo pre-processing on the device (python/pose/estimate_pose.py:83-105): uint8 [h][w][3] image (device memory; BGR in
 * the demo) -> 64 px edge-replicated below/right (:90-96) -> scipy.misc.imresize(image, scale, 'bilinear') (:98) == Pillow's
 * 8-bit bilinear resample, bit-exact -> minus mean3 per channel (:99) -> top-left crop / zero fill to the net input
 * (:84-88,101-105) as fp32 [3][out_h][out_w] -- the `data` blob's layout (:225-227).  A plan is immutable after create and
 * owns only its coefficient tables; the caller owns image, output and workspace.  mean3 is read on the host. */
typedef struct dc_preprocess_plan dc_preprocess_plan;
int dc_preprocess_plan_create(int h, int w, double scale, dc_preprocess_plan** plan);
int dc_preprocess_plan_info(const dc_preprocess_plan* plan, int* out_h, int* out_w, size_t* workspace_bytes);
int dc_preprocess_u8_forward(const dc_preprocess_plan* plan, const unsigned char* img, const float* mean3, float* out,
                             void* workspace, void* stream);
int dc_preprocess_plan_destroy(dc_preprocess_plan* plan);
/* A batch of decoded images into the net's input blob: img uint8 [n][h][w][3] (device) -> out fp32 [n][3][h][w] minus
 * mean3[channel] -- the demo's `image.astype('float32') - _MEAN` + HWC->CHW (estimate_pose.py:99,225-227) without its rescale
 * (dc_preprocess_u8_forward covers that).  What the multi-GPU batch exchange ships: 3 B/pixel instead of 12.  mean3 is read on
 * the host. */
int dc_images_u8_to_blob(const unsigned char* img, int n, int h, int w, const float* mean3, float* out, void* stream);
/* Blob materialisation: fp32 NCHW <-> split NHWC. */
int dc_nchw_to_split(const float* x, int n, int c, int h, int w, void* out, void* stream);
int dc_split_to_nchw(const void* x, int n, int c, int h, int w, float* out, void* stream);


/* ---- per-layer fp32 NCHW kernels (Layer::Forward_gpu, one layer at a time) -------------- */
/* (x - mean[c]) / std[c]: BatchNormLayer::Forward_gpu inference branch, batch_norm_layer.cu:22-89 */
int dc_bn_forward_nchw(const float* x, const float* mean, const float* stddev, int n, int c, int hw, float* y, void* stream);
/* x * gamma[c] (+ beta[c]): ScaleLayer::Forward_gpu, scale_layer.cu:19-56 */
int dc_scale_forward_nchw(const float* x, const float* gamma, const float* beta, int n, int c, int hw, float* y, void* stream);
/* ReLULayer::Forward_gpu relu_layer.cu:8-32; SigmoidLayer::Forward_gpu sigmoid_layer.cu:8-24 */
int dc_relu_forward(const float* x, long long count, float negative_slope, float* y, void* stream);
int dc_sigmoid_forward(const float* x, long long count, float* y, void* stream);
/* y = ca*a + cb*b: EltwiseLayer::Forward_gpu SUM, eltwise_layer.cu:47-53 */
int dc_axpby_forward(const float* a, float ca, const float* b, float cb, long long count, float* y, void* stream);
/* CropLayer::Forward_gpu crop_layer.cu:9-38 */
int dc_crop_forward_nchw(const float* x, int n, int c, int h, int w, int off_h, int off_w, int ho, int wo, float* y, void* stream);
/* PoolingLayer::Forward_gpu MAX pooling_layer.cu:10-47 (any kernel/stride/pad; ho/wo from the layer's Reshape) */
int dc_maxpool_forward_nchw(const float* x, int n, int c, int h, int w, int kh, int kw, int sh, int sw, int ph, int pw,
                            int ho, int wo, float* y, void* stream);
/* Generic direct (de)convolution for geometries outside the tcgen05 kernel: conv_layer.cu:8-24, deconv_layer.cu:8-24.
 * w is the Caffe blob ([cout][cin][kh][kw] for conv, [cin][cout][kh][kw] for deconv); bias may be NULL. */
int dc_conv_direct_nchw(const float* x, const float* w, const float* bias, int n, int cin, int h, int wd, int cout,
                        int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, float* y, void* stream);
int dc_deconv_direct_nchw(const float* x, const float* w, const float* bias, int n, int cin, int h, int wd, int cout,
                          int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPCUT_B200_H_ */
