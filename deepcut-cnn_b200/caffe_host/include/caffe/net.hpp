// Net<Dtype>: prototxt -> DAG of layers and named blobs; public API of the reference's
// include/caffe/net.hpp:24-313 for the inference path (Init, Forward*, Reshape,
// CopyTrainedLayersFrom, ToProto, name tables).  In GPU mode ForwardPrefilled runs the fused
// B200 plan (dc_engine.hpp) over the same blobs; Net::set_fusion(false) or a topology the planner
// does not recognise falls back to per-layer Layer::Forward (still CUDA, never CPU).
#pragma once
#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/layer.hpp"
#include "caffe/proto/caffe.pb.h"

namespace caffe {

class FusedPlan;
class PlanWeightCache;

template <typename Dtype>
class Net {
 public:
  explicit Net(const NetParameter& param, const Net* root_net = NULL);
  explicit Net(const string& param_file, Phase phase, const Net* root_net = NULL);
  virtual ~Net();

  void Init(const NetParameter& param);

  const vector<Blob<Dtype>*>& ForwardPrefilled(Dtype* loss = NULL);
  Dtype ForwardFromTo(int start, int end);
  Dtype ForwardFrom(int start) { return ForwardFromTo(start, static_cast<int>(layers_.size()) - 1); }
  Dtype ForwardTo(int end) { return ForwardFromTo(0, end); }
  const vector<Blob<Dtype>*>& Forward(const vector<Blob<Dtype>*>& bottom, Dtype* loss = NULL);
  const vector<Blob<Dtype>*>& Forward(Dtype* loss = NULL) { return ForwardPrefilled(loss); }

  void Reshape();
  void ShareTrainedLayersWith(const Net* other);
  void CopyTrainedLayersFrom(const NetParameter& param);
  void CopyTrainedLayersFrom(const string trained_filename);
  void CopyTrainedLayersFromBinaryProto(const string trained_filename);
  void ToProto(NetParameter* param, bool write_diff = false) const;

  inline const string& name() const { return name_; }
  inline const vector<string>& layer_names() const { return layer_names_; }
  inline const vector<string>& blob_names() const { return blob_names_; }
  inline const vector<shared_ptr<Blob<Dtype> > >& blobs() const { return blobs_; }
  inline const vector<shared_ptr<Layer<Dtype> > >& layers() const { return layers_; }
  inline Phase phase() const { return phase_; }
  inline const vector<vector<Blob<Dtype>*> >& bottom_vecs() const { return bottom_vecs_; }
  inline const vector<vector<Blob<Dtype>*> >& top_vecs() const { return top_vecs_; }
  inline const vector<vector<int> >& bottom_ids() const { return bottom_id_vecs_; }
  inline const vector<vector<int> >& top_ids() const { return top_id_vecs_; }
  inline const vector<int>& top_ids(int i) const { return top_id_vecs_[i]; }
  inline const vector<int>& bottom_ids(int i) const { return bottom_id_vecs_[i]; }
  inline const vector<shared_ptr<Blob<Dtype> > >& params() const { return params_; }
  inline const vector<Blob<Dtype>*>& learnable_params() const { return learnable_params_; }
  inline int num_inputs() const { return static_cast<int>(net_input_blobs_.size()); }
  inline int num_outputs() const { return static_cast<int>(net_output_blobs_.size()); }
  inline const vector<Blob<Dtype>*>& input_blobs() const { return net_input_blobs_; }
  inline const vector<Blob<Dtype>*>& output_blobs() const { return net_output_blobs_; }
  inline const vector<int>& input_blob_indices() const { return net_input_blob_indices_; }
  inline const vector<int>& output_blob_indices() const { return net_output_blob_indices_; }
  bool has_blob(const string& blob_name) const;
  const shared_ptr<Blob<Dtype> > blob_by_name(const string& blob_name) const;
  bool has_layer(const string& layer_name) const;
  const shared_ptr<Layer<Dtype> > layer_by_name(const string& layer_name) const;
  void set_debug_info(const bool value) { debug_info_ = value; }
  // What the last debug_info forward printed (net.cpp:648-735 of the reference logs "[Forward] Layer L, top blob B data: mean|x|"
  // per top): kept so callers and tests can compare the probes blob by blob instead of parsing the log.
  struct DebugRecord { string layer, blob; double mean_abs; };
  const vector<DebugRecord>& debug_log() const { return debug_log_; }

  static void FilterNet(const NetParameter& param, NetParameter* param_filtered);
  static bool StateMeetsRule(const NetState& state, const NetStateRule& rule, const string& layer_name);

  // ---- B200 extensions (not in the reference) ----
  // Fused plan on/off (default on in GPU mode).  With fusion the blobs of fused-away intermediates
  // are not written unless materialize_intermediates(true) asks for every named blob to be filled
  // with its Caffe-final value (the value the in-place BN/Scale/ReLU chain would have left there).
  void set_fusion(bool on);
  bool fusion() const { return fusion_; }
  void materialize_intermediates(bool on);
  // Net outputs the caller will not read (e.g. "next_pred": estimate_pose.py:231 reads only prob and loc_pred).  The fused plan
  // leaves every head whose final blob is named here out of its merged head GEMMs and writes nothing to that blob (it reads as
  // "not written by the last forward", blob_fresh() == false).  Per-layer execution ignores the list and computes everything.
  void set_skipped_outputs(const vector<string>& blob_names);
  const set<string>& skipped_outputs() const { return skipped_outputs_; }
  bool fused_last_forward() const { return fused_last_forward_; }
  // Why the planner declined (empty when the fused plan is active).
  const string& fusion_diagnostic() const { return fusion_diag_; }
  // Device kernels launched by the last ForwardPrefilled.
  long long last_forward_launches() const { return last_launches_; }
  void InvalidatePlan();
  void set_step_timing(bool on) { step_timing_ = on; }
  FusedPlan* plan() const { return plan_; }
  // Does blob i hold the value of the LAST forward?  The fused plan writes only the net's outputs (and, on request, the
  // named intermediates it can reconstruct); every other blob keeps whatever an earlier per-layer forward left there.  The
  // reference fills every blob on every forward, so a caller reading a stale one must be told rather than handed old data.
  bool blob_fresh(int i) const { return blob_fresh_.empty() || blob_fresh_[i] != 0; }

 protected:
  void AppendTop(const NetParameter& param, const int layer_id, const int top_id, set<string>* available_blobs, map<string, int>* blob_name_to_idx);
  int AppendBottom(const NetParameter& param, const int layer_id, const int bottom_id, set<string>* available_blobs, map<string, int>* blob_name_to_idx);
  void AppendParam(const NetParameter& param, const int layer_id, const int param_id);
  void ForwardDebugInfo(const int layer_id);

  string name_;
  Phase phase_;
  vector<shared_ptr<Layer<Dtype> > > layers_;
  vector<string> layer_names_;
  map<string, int> layer_names_index_;
  vector<shared_ptr<Blob<Dtype> > > blobs_;
  vector<string> blob_names_;
  map<string, int> blob_names_index_;
  vector<vector<Blob<Dtype>*> > bottom_vecs_;
  vector<vector<int> > bottom_id_vecs_;
  vector<vector<Blob<Dtype>*> > top_vecs_;
  vector<vector<int> > top_id_vecs_;
  vector<int> net_input_blob_indices_;
  vector<int> net_output_blob_indices_;
  vector<Blob<Dtype>*> net_input_blobs_;
  vector<Blob<Dtype>*> net_output_blobs_;
  vector<shared_ptr<Blob<Dtype> > > params_;
  vector<Blob<Dtype>*> learnable_params_;
  bool debug_info_ = false;
  vector<DebugRecord> debug_log_;
  NetParameter filtered_param_;     // post FilterNet + InsertSplits

  bool fusion_ = true;
  bool materialize_ = false;
  set<string> skipped_outputs_;
  bool step_timing_ = false;
  bool fused_last_forward_ = false;
  string fusion_diag_;
  long long last_launches_ = 0;
  FusedPlan* plan_ = nullptr;
  shared_ptr<PlanWeightCache> plan_weights_;   // survives re-planning on reshape
  vector<vector<int> > plan_input_shapes_;
  vector<char> blob_fresh_;          // empty = every blob is current (no fused forward has run yet)

  DISABLE_COPY_AND_ASSIGN(Net);
};

}  // namespace caffe
