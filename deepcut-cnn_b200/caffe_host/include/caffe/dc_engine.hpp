// FusedPlan: the B200 execution plan Net::ForwardPrefilled runs in GPU mode.
//
// Built once per (topology, weights): pattern-matches the layer list into fused steps
//   Conv1        data (fp32 NCHW blob) -> conv 7x7/2 + BN + Scale + ReLU -> split NHWC
//   ConvBN       [Convolution + BatchNorm? + Scale? + ReLU?] (+ Eltwise SUM shortcut + ReLU) on tcgen05
//   Subsample    stride-2 spatial gather feeding the strided 1x1 convolutions
//   MaxPool      3x3/2 ceil-mode
//   HeadGroup    all Deconvolution+Convolution(1x1)+Crop+Eltwise(+Sigmoid) heads that share their two
//                inputs: ONE merged deconv GEMM, ONE merged 1x1 GEMM, one finish kernel per head that
//                writes the fp32 NCHW output blob
// and re-planned (shapes, arena offsets) whenever the input blobs change shape.  Intermediate
// activations live in one device arena as split-fp16 NHWC with liveness-based reuse (the reference
// allocates every top plus BN/Scale scratch separately: blob.cpp:23-43, batch_norm_layer.cpp:50-51).
// A topology the matcher does not recognise makes Build() return NULL with a diagnostic and the Net
// falls back to per-layer Forward_gpu.
#pragma once
#include <string>
#include <vector>

#include "caffe/net.hpp"

namespace caffe {

class FusedPlan {
 public:
  static FusedPlan* Build(Net<float>& net, bool materialize, std::string* why_not);
  ~FusedPlan();
  void Run();
  // true when a parameter blob was written on the host since the weights were packed
  bool WeightsStale() const;
  size_t arena_bytes() const { return arena_bytes_; }
  size_t weight_bytes() const { return weight_bytes_; }
  int num_steps() const { return static_cast<int>(steps_.size()); }
  std::string Describe() const;
  // Per-step device timing: when enabled Run() brackets every step with CUDA events on the forward
  // stream (the reference's `caffe time` idiom, tools/caffe.cpp:302-388, per fused step instead of
  // per layer).  StepInfo gives the last run's duration and the step's algorithmic work.
  struct StepInfo { std::string name, type; double ms, flops, bytes; };
  void set_step_timing(bool on) { step_timing_ = on; }
  std::vector<StepInfo> LastStepInfo();

  struct Tensor;
  struct Step;

 private:
  FusedPlan() {}
  bool Match(Net<float>& net, bool materialize, std::string* why);
  void PlanMemory();
  void UploadWeights(Net<float>& net);

  Net<float>* net_ = nullptr;
  std::vector<Tensor*> tensors_;
  std::vector<Step*> steps_;
  void* arena_ = nullptr;
  size_t arena_bytes_ = 0;
  size_t weight_bytes_ = 0;
  std::vector<void*> weight_allocs_;
  std::vector<int> split_layers_;     // Split layer ids to alias after a materialised run
  bool materialize_ = false;
  bool step_timing_ = false;
  std::vector<void*> events_;
  std::vector<std::pair<SyncedMemory*, unsigned long long> > weight_epochs_;
};

}  // namespace caffe
