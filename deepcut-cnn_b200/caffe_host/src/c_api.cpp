// C ABI over the C++ Caffe host, for the Python `caffe` shim (python/caffe/_caffe.cpp in the
// reference is a boost.python module over the same classes; boost is not in this image, so the
// binding is ctypes over these functions).  Every call catches the host's CHECK failures
// (caffe::FatalError) and turns them into a non-zero return + caffe_last_error().
#include <cstring>
#include <memory>
#include <string>

#include "caffe/caffe.hpp"
#include "caffe/dc_engine.hpp"
#include "caffe_b200_c.h"

using namespace caffe;  // NOLINT

namespace {
thread_local std::string g_error;
struct BlobHandle { shared_ptr<Blob<float> > blob; };

template <class F>
int Guard(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}
Net<float>* N(void* h) { return static_cast<Net<float>*>(h); }
Blob<float>* B(void* h) { return static_cast<BlobHandle*>(h)->blob.get(); }
}  // namespace

extern "C" {

const char* caffe_last_error(void) { return g_error.c_str(); }

int caffe_set_mode(int gpu) { return Guard([&] { Caffe::set_fatal_throws(true); Caffe::set_mode(gpu ? Caffe::GPU : Caffe::CPU); }); }
int caffe_get_mode(void) { return Caffe::mode() == Caffe::GPU; }
int caffe_set_device(int id) { return Guard([&] { Caffe::set_fatal_throws(true); Caffe::SetDevice(id); }); }
int caffe_device_count(void) { return Caffe::device_count(); }
void caffe_set_log_level(int level) { Caffe::set_log_level(level); }
void* caffe_stream(void) { void* s = nullptr; Guard([&] { s = Caffe::stream(); }); return s; }
int caffe_sync(void) { return Guard([&] { DC_CHECK(dc_stream_sync(Caffe::stream())); }); }

void* caffe_net_create(const char* prototxt, int phase) {
  Net<float>* net = nullptr;
  Caffe::set_fatal_throws(true);
  if (Guard([&] { net = new Net<float>(std::string(prototxt), phase == 0 ? TRAIN : TEST); })) return nullptr;
  return net;
}
void* caffe_net_create_from_string(const char* prototxt_text, int phase) {
  Net<float>* net = nullptr;
  Caffe::set_fatal_throws(true);
  if (Guard([&] {
        NetParameter p;
        std::string err;
        CHECK(p.ParseFromTextString(prototxt_text, &err)) << "prototxt: " << err;
        p.mutable_state()->set_phase(phase == 0 ? TRAIN : TEST);
        net = new Net<float>(p);
      }))
    return nullptr;
  return net;
}
void caffe_net_destroy(void* net) { delete N(net); }
int caffe_net_copy_trained_from(void* net, const char* file) { return Guard([&] { N(net)->CopyTrainedLayersFrom(std::string(file)); }); }
int caffe_net_save(void* net, const char* file) {
  return Guard([&] { NetParameter p; N(net)->ToProto(&p, false); WriteProtoToBinaryFile(p, file); });
}
int caffe_net_forward(void* net) { return Guard([&] { N(net)->ForwardPrefilled(); }); }
int caffe_net_forward_from_to(void* net, int start, int end) { return Guard([&] { N(net)->ForwardFromTo(start, end); }); }
int caffe_net_reshape(void* net) { return Guard([&] { N(net)->Reshape(); }); }
const char* caffe_net_name(void* net) { return N(net)->name().c_str(); }

int caffe_net_num_blobs(void* net) { return static_cast<int>(N(net)->blobs().size()); }
const char* caffe_net_blob_name(void* net, int i) { return N(net)->blob_names()[i].c_str(); }
int caffe_net_num_layers(void* net) { return static_cast<int>(N(net)->layers().size()); }
const char* caffe_net_layer_name(void* net, int i) { return N(net)->layer_names()[i].c_str(); }
const char* caffe_net_layer_type(void* net, int i) { return N(net)->layers()[i]->type(); }
int caffe_net_layer_num_blobs(void* net, int i) { return static_cast<int>(N(net)->layers()[i]->blobs().size()); }
int caffe_net_layer_num_bottoms(void* net, int i) { return static_cast<int>(N(net)->bottom_ids(i).size()); }
int caffe_net_layer_bottom_id(void* net, int i, int j) { return N(net)->bottom_ids(i)[j]; }
int caffe_net_layer_num_tops(void* net, int i) { return static_cast<int>(N(net)->top_ids(i).size()); }
int caffe_net_layer_top_id(void* net, int i, int j) { return N(net)->top_ids(i)[j]; }
int caffe_net_num_inputs(void* net) { return N(net)->num_inputs(); }
int caffe_net_input_index(void* net, int i) { return N(net)->input_blob_indices()[i]; }
int caffe_net_num_outputs(void* net) { return N(net)->num_outputs(); }
int caffe_net_output_index(void* net, int i) { return N(net)->output_blob_indices()[i]; }
int caffe_net_layer_weights_changed(void* net, int i) {
  return Guard([&] { N(net)->layers()[i]->OnWeightsChanged(); N(net)->InvalidatePlan(); });
}

void* caffe_net_blob(void* net, int i) {
  BlobHandle* h = new BlobHandle();
  h->blob = N(net)->blobs()[i];
  return h;
}
void* caffe_net_layer_blob(void* net, int layer, int j) {
  BlobHandle* h = new BlobHandle();
  h->blob = N(net)->layers()[layer]->blobs()[j];
  return h;
}
void caffe_blob_release(void* blob) { delete static_cast<BlobHandle*>(blob); }
int caffe_blob_num_axes(void* blob) { return B(blob)->num_axes(); }
int caffe_blob_shape(void* blob, int axis) { return B(blob)->shape(axis); }
int caffe_blob_count(void* blob) { return B(blob)->count(); }
int caffe_blob_reshape(void* blob, int naxes, const int* dims) {
  return Guard([&] { B(blob)->Reshape(vector<int>(dims, dims + naxes)); });
}
float* caffe_blob_mutable_cpu_data(void* blob) { float* p = nullptr; Guard([&] { p = B(blob)->mutable_cpu_data(); }); return p; }
const float* caffe_blob_cpu_data(void* blob) { const float* p = nullptr; Guard([&] { p = B(blob)->cpu_data(); }); return p; }
float* caffe_blob_mutable_cpu_diff(void* blob) { float* p = nullptr; Guard([&] { p = B(blob)->mutable_cpu_diff(); }); return p; }
const float* caffe_blob_gpu_data(void* blob) { const float* p = nullptr; Guard([&] { p = B(blob)->gpu_data(); }); return p; }
float* caffe_blob_mutable_gpu_data(void* blob) { float* p = nullptr; Guard([&] { p = B(blob)->mutable_gpu_data(); }); return p; }
float* caffe_blob_overwrite_gpu_data(void* blob) { float* p = nullptr; Guard([&] { p = B(blob)->overwrite_gpu_data(); }); return p; }
int caffe_blob_data_head(void* blob) {
  int h = -1;
  Guard([&] { CHECK(B(blob)->data()) << "blob has no memory yet"; h = static_cast<int>(B(blob)->data()->head()); });
  return h;
}

int caffe_net_set_fusion(void* net, int on) { return Guard([&] { N(net)->set_fusion(on != 0); }); }
int caffe_net_materialize_intermediates(void* net, int on) { return Guard([&] { N(net)->materialize_intermediates(on != 0); }); }
int caffe_net_set_skipped_outputs(void* net, const char* comma_separated_blob_names) {
  return Guard([&] {
    std::vector<std::string> names;
    std::string cur;
    for (const char* c = comma_separated_blob_names ? comma_separated_blob_names : ""; ; ++c) {
      if (*c == ',' || *c == 0) { if (!cur.empty()) names.push_back(cur); cur.clear(); if (*c == 0) break; }
      else if (*c != ' ') cur += *c;
    }
    N(net)->set_skipped_outputs(names);
  });
}
int caffe_net_fused_last_forward(void* net) { return N(net)->fused_last_forward(); }
const char* caffe_net_fusion_diagnostic(void* net) { return N(net)->fusion_diagnostic().c_str(); }
long long caffe_net_last_forward_launches(void* net) { return N(net)->last_forward_launches(); }

int caffe_net_set_step_timing(void* net, int on) { return Guard([&] { N(net)->set_step_timing(on != 0); }); }
int caffe_net_num_steps(void* net) { return N(net)->plan() ? N(net)->plan()->num_steps() : 0; }
int caffe_net_step_info(void* net, char* names, int names_cap, double* ms, double* flops, double* bytes, int max_steps) {
  return Guard([&] {
    CHECK(N(net)->plan()) << "no fused plan (run a GPU forward first)";
    std::vector<FusedPlan::StepInfo> info = N(net)->plan()->LastStepInfo();
    CHECK_LE((int)info.size(), max_steps);
    std::string all;
    for (size_t i = 0; i < info.size(); ++i) {
      all += info[i].type + " " + info[i].name + "\n";
      ms[i] = info[i].ms; flops[i] = info[i].flops; bytes[i] = info[i].bytes;
    }
    CHECK_LT((int)all.size(), names_cap);
    memcpy(names, all.c_str(), all.size() + 1);
  });
}
long long caffe_net_arena_bytes(void* net) { return N(net)->plan() ? (long long)N(net)->plan()->arena_bytes() : 0; }
long long caffe_net_weight_bytes(void* net) { return N(net)->plan() ? (long long)N(net)->plan()->weight_bytes() : 0; }
int caffe_net_set_debug_info(void* net, int on) { return Guard([&] { N(net)->set_debug_info(on != 0); }); }
int caffe_net_debug_info(void* net, char* names, int names_cap, double* mean_abs, int max_records) {
  int n = -1;
  Guard([&] {
    const auto& log = N(net)->debug_log();
    CHECK_LE((int)log.size(), max_records);
    std::string all;
    for (size_t i = 0; i < log.size(); ++i) {
      all += log[i].layer + " " + log[i].blob + "\n";
      mean_abs[i] = log[i].mean_abs;
    }
    CHECK_LT((int)all.size(), names_cap);
    memcpy(names, all.c_str(), all.size() + 1);
    n = static_cast<int>(log.size());
  });
  return n;
}
int caffe_net_blob_fresh(void* net, int i) { return N(net)->blob_fresh(i) ? 1 : 0; }
int caffe_net_describe_plan(void* net, char* out, int out_cap) {
  return Guard([&] {
    N(net)->Reshape();
    std::string why;
    std::unique_ptr<FusedPlan> plan(FusedPlan::Build(*N(net), false, &why, nullptr, /*dry_run=*/true));
    CHECK(plan) << "the net does not fuse: " << why;
    const std::string s = plan->Describe();
    CHECK_LT((int)s.size(), out_cap) << "output buffer too small";
    memcpy(out, s.c_str(), s.size() + 1);
  });
}

int caffe_insert_splits_text(const char* prototxt_text, char* out, int out_cap) {
  return Guard([&] {
    NetParameter p, q;
    std::string err;
    CHECK(p.ParseFromTextString(prototxt_text, &err)) << "prototxt: " << err;
    InsertSplits(p, &q);
    const std::string s = q.DebugString();
    CHECK_LT((int)s.size(), out_cap) << "output buffer too small";
    memcpy(out, s.c_str(), s.size() + 1);
  });
}

}  // extern "C"
