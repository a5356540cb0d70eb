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
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "caffe/net.hpp"

namespace caffe {

// Device copies of the transformed weights (packed split-fp16 matrices, folded scale/shift), shared by
// successive plans of one Net: a reshape re-plans shapes and the arena but does not re-pack 263 MB of
// weights.  Dropped when a parameter blob is written on the host (SyncedMemory::host_write_epoch).
class PlanWeightCache {
 public:
  struct Entry { void* w = nullptr; float* scale = nullptr; float* shift = nullptr; };
  ~PlanWeightCache();
  bool Stale() const;
  void Snapshot(Net<float>& net);
  std::map<std::string, Entry> entries;
  std::vector<void*> allocs;
  size_t bytes = 0;

 private:
  // (parameter blob, the SyncedMemory it had when the weights were packed, that memory's write epoch): holding the shared_ptr keeps
  // the comparison valid when a blob is reshaped or re-pointed (ShareData) afterwards
  struct Seen { const Blob<float>* blob; std::shared_ptr<SyncedMemory> mem; unsigned long long epoch; };
  std::vector<Seen> epochs_;
};

class FusedPlan {
 public:
  // dry_run: match, schedule and place the arena WITHOUT touching the device (no allocation, no weight upload): what
  // caffe_net_describe_plan and the CPU tests of the planner use.  A dry-run plan cannot Run().
  static FusedPlan* Build(Net<float>& net, bool materialize, std::string* why_not, std::shared_ptr<PlanWeightCache>* cache, bool dry_run = false);
  ~FusedPlan();
  void Run();
  // true when a parameter blob was written on the host since the weights were packed
  bool WeightsStale() const;
  size_t arena_bytes() const { return arena_bytes_; }
  size_t weight_bytes() const { return weights_ ? weights_->bytes : 0; }
  int num_steps() const { return static_cast<int>(steps_.size()); }
  std::string Describe() const;
  // per Net blob: 1 when Run() leaves the blob's current value in it (outputs; with materialisation the named intermediates
  // that exist as tensors and the Split tops aliasing them)
  std::vector<char> WrittenBlobs() const;
  // Per-step device timing: when enabled Run() brackets every step with CUDA events on the forward
  // stream (the reference's `caffe time` idiom, tools/caffe.cpp:302-388, per fused step instead of
  // per layer).  StepInfo gives the last run's duration and the step's algorithmic work.
  struct StepInfo { std::string name, type; double ms, flops, bytes; };
  void set_step_timing(bool on) { step_timing_ = on; }
  std::vector<StepInfo> LastStepInfo();

  struct Tensor;
  struct Step;
  // One launch group of the schedule: step `step` over images [i0, i0 + cn) of its tensors (cn = 0: the whole batch).
  struct Issue { int step, i0, cn; };
  // A run of consecutive ConvBN steps (the bottleneck blocks of one ResNet stage) executed sub-batch by sub-batch so that
  // the block intermediates stay in L2 (PlanSchedule); chunk = images per pass.
  struct Segment { int first, last, chunk; size_t bytes_per_image; };
  const std::vector<Segment>& segments() const { return segments_; }

 private:
  FusedPlan() {}
  bool Match(Net<float>& net, bool materialize, std::string* why);
  void PlanSchedule();
  void PlanMemory(bool dry_run);
  // Replays the schedule against the arena placement on the host: every (tensor, image) a launch reads must still be owned by
  // that tensor's last writer -- catches liveness / aliasing / chunk-order mistakes of the planner.  Empty string = consistent.
  std::string VerifySchedule() const;
  void UploadWeights(Net<float>& net);

  Net<float>* net_ = nullptr;
  std::vector<Tensor*> tensors_;
  std::vector<Step*> steps_;
  void* arena_ = nullptr;
  size_t arena_bytes_ = 0;
  void* splitk_ws_ = nullptr;           // tail of the arena: scratch of the split-K conv launches
  size_t splitk_ws_bytes_ = 0;
  std::shared_ptr<PlanWeightCache> weights_;
  std::vector<int> split_layers_;     // Split layer ids to alias after a materialised run
  bool materialize_ = false;
  bool step_timing_ = false;
  // CUDA graph of the step sequence, valid while the blob device pointers it baked in are unchanged
  void* graph_ = nullptr;
  std::vector<const void*> graph_ptrs_;
  bool graph_failed_ = false;
  void IssueSteps(const std::vector<const void*>& blob_ptrs, void* stream);
  std::vector<void*> events_;          // step timing: one event before every Issue + one at the end
  std::vector<Issue> schedule_;
  std::vector<Segment> segments_;
};

}  // namespace caffe
