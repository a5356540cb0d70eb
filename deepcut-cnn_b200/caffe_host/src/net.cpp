// Net: graph construction with the reference's semantics (src/caffe/net.cpp:40-284 Init, :287-346
// FilterNet/StateMeetsRule, :385-468 AppendTop/AppendBottom, :565-611 Forward*, :798-858 Reshape /
// CopyTrainedLayersFrom, :984-999 ToProto) + dispatch of ForwardPrefilled to the fused B200 plan.
#include "caffe/net.hpp"

#include <algorithm>

#include "caffe/dc_engine.hpp"
#include "caffe/layer_factory.hpp"
#include "caffe/util/insert_splits.hpp"
#include "caffe/util/io.hpp"
#include "deepcut_b200.h"

namespace caffe {

template <typename Dtype>
Net<Dtype>::Net(const NetParameter& param, const Net* root_net) {
  CHECK(root_net == NULL) << "shared root nets (multi-solver training) are outside the inference path";
  Init(param);
}

template <typename Dtype>
Net<Dtype>::Net(const string& param_file, Phase phase, const Net* root_net) {
  CHECK(root_net == NULL) << "shared root nets (multi-solver training) are outside the inference path";
  NetParameter param;
  ReadNetParamsFromTextFileOrDie(param_file, &param);
  param.mutable_state()->set_phase(phase);
  Init(param);
}

template <typename Dtype>
Net<Dtype>::~Net() { delete plan_; }

template <typename Dtype>
void Net<Dtype>::Init(const NetParameter& in_param) {
  phase_ = in_param.state().phase();
  NetParameter filtered_param;
  FilterNet(in_param, &filtered_param);
  LOG(INFO) << "Initializing net from parameters: \n" << filtered_param.DebugString();
  NetParameter& param = filtered_param_;
  InsertSplits(filtered_param, &param);
  name_ = param.name();
  map<string, int> blob_name_to_idx;
  set<string> available_blobs;
  CHECK(param.input_dim_size() == 0 || param.input_shape_size() == 0) << "Must specify either input_shape OR deprecated input_dim, not both.";
  if (param.input_dim_size() > 0) CHECK_EQ(param.input_size() * 4, param.input_dim_size()) << "Incorrect input blob dimension specifications.";
  else CHECK_EQ(param.input_size(), param.input_shape_size()) << "Exactly one input_shape must be specified per input.";
  for (int input_id = 0; input_id < param.input_size(); ++input_id) AppendTop(param, -1, input_id, &available_blobs, &blob_name_to_idx);

  const int nlayers = param.layer_size();
  bottom_vecs_.resize(nlayers);
  top_vecs_.resize(nlayers);
  bottom_id_vecs_.resize(nlayers);
  top_id_vecs_.resize(nlayers);
  for (int layer_id = 0; layer_id < nlayers; ++layer_id) {
    if (!param.layer(layer_id).has_phase()) param.mutable_layer(layer_id)->set_phase(phase_);   // inherit the net's phase
    const LayerParameter& layer_param = param.layer(layer_id);
    layers_.push_back(LayerRegistry<Dtype>::CreateLayer(layer_param));
    layer_names_.push_back(layer_param.name());
    LOG(INFO) << "Creating Layer " << layer_param.name();
    for (int bottom_id = 0; bottom_id < layer_param.bottom_size(); ++bottom_id) AppendBottom(param, layer_id, bottom_id, &available_blobs, &blob_name_to_idx);
    for (int top_id = 0; top_id < layer_param.top_size(); ++top_id) AppendTop(param, layer_id, top_id, &available_blobs, &blob_name_to_idx);
    Layer<Dtype>* layer = layers_[layer_id].get();
    if (layer->AutoTopBlobs()) {
      const int needed = std::max(layer->MinTopBlobs(), layer->ExactNumTopBlobs());
      for (int n = layer_param.top_size(); n < needed; ++n) AppendTop(param, layer_id, n, NULL, NULL);
    }
    layer->SetUp(bottom_vecs_[layer_id], top_vecs_[layer_id]);
    LOG(INFO) << "Setting up " << layer_names_[layer_id];
    for (size_t top_id = 0; top_id < top_vecs_[layer_id].size(); ++top_id)
      LOG(INFO) << "Top shape: " << top_vecs_[layer_id][top_id]->shape_string();
    const int num_param_blobs = static_cast<int>(layer->blobs().size());
    CHECK_LE(layer_param.param_size(), num_param_blobs) << "Too many params specified for layer " << layer_param.name();
    for (int param_id = 0; param_id < num_param_blobs; ++param_id) AppendParam(param, layer_id, param_id);
  }
  // blobs nobody consumed are the net's outputs, in name order (std::set iteration, net.cpp:268-274)
  for (set<string>::iterator it = available_blobs.begin(); it != available_blobs.end(); ++it) {
    LOG(INFO) << "This network produces output " << *it;
    net_output_blobs_.push_back(blobs_[blob_name_to_idx[*it]].get());
    net_output_blob_indices_.push_back(blob_name_to_idx[*it]);
  }
  for (size_t i = 0; i < blob_names_.size(); ++i) blob_names_index_[blob_names_[i]] = static_cast<int>(i);
  for (size_t i = 0; i < layer_names_.size(); ++i) layer_names_index_[layer_names_[i]] = static_cast<int>(i);
  debug_info_ = param.debug_info();
  LOG(INFO) << "Network initialization done.";
}

template <typename Dtype>
void Net<Dtype>::FilterNet(const NetParameter& param, NetParameter* param_filtered) {
  NetState net_state(param.state());
  param_filtered->CopyFrom(param);
  param_filtered->clear_layer();
  for (int i = 0; i < param.layer_size(); ++i) {
    const LayerParameter& layer_param = param.layer(i);
    const string& layer_name = layer_param.name();
    CHECK(layer_param.include_size() == 0 || layer_param.exclude_size() == 0) << "Specify either include rules or exclude rules; not both.";
    bool layer_included = (layer_param.include_size() == 0);   // no include rules: in by default
    for (int j = 0; layer_included && j < layer_param.exclude_size(); ++j)
      if (StateMeetsRule(net_state, layer_param.exclude(j), layer_name)) layer_included = false;
    for (int j = 0; !layer_included && j < layer_param.include_size(); ++j)
      if (StateMeetsRule(net_state, layer_param.include(j), layer_name)) layer_included = true;
    if (layer_included) param_filtered->add_layer()->CopyFrom(layer_param);
  }
}

template <typename Dtype>
bool Net<Dtype>::StateMeetsRule(const NetState& state, const NetStateRule& rule, const string& layer_name) {
  if (rule.has_phase() && rule.phase() != state.phase()) return false;
  if (rule.has_min_level() && state.level() < rule.min_level()) return false;
  if (rule.has_max_level() && state.level() > rule.max_level()) return false;
  for (int i = 0; i < rule.stage_size(); ++i) {     // every required stage must be present
    bool has_stage = false;
    for (int j = 0; !has_stage && j < state.stage_size(); ++j) has_stage = rule.stage(i) == state.stage(j);
    if (!has_stage) return false;
  }
  for (int i = 0; i < rule.not_stage_size(); ++i)   // none of the forbidden stages may be present
    for (int j = 0; j < state.stage_size(); ++j)
      if (rule.not_stage(i) == state.stage(j)) return false;
  return true;
}

template <typename Dtype>
void Net<Dtype>::AppendTop(const NetParameter& param, const int layer_id, const int top_id, set<string>* available_blobs,
                           map<string, int>* blob_name_to_idx) {
  shared_ptr<LayerParameter> layer_param(layer_id >= 0 ? new LayerParameter(param.layer(layer_id)) : NULL);
  const string blob_name = layer_param ? (layer_param->top_size() > top_id ? layer_param->top(top_id) : "(automatic)") : param.input(top_id);
  if (blob_name_to_idx && layer_param && layer_param->bottom_size() > top_id && blob_name == layer_param->bottom(top_id)) {
    // in-place: top shares the bottom blob (same index)
    LOG(INFO) << layer_param->name() << " -> " << blob_name << " (in-place)";
    top_vecs_[layer_id].push_back(blobs_[(*blob_name_to_idx)[blob_name]].get());
    top_id_vecs_[layer_id].push_back((*blob_name_to_idx)[blob_name]);
  } else if (blob_name_to_idx && blob_name_to_idx->find(blob_name) != blob_name_to_idx->end()) {
    LOG(FATAL) << "Top blob '" << blob_name << "' produced by multiple sources.";
  } else {
    shared_ptr<Blob<Dtype> > blob_pointer(new Blob<Dtype>());
    const int blob_id = static_cast<int>(blobs_.size());
    blobs_.push_back(blob_pointer);
    blob_names_.push_back(blob_name);
    if (blob_name_to_idx) (*blob_name_to_idx)[blob_name] = blob_id;
    if (layer_id == -1) {
      if (param.input_dim_size() > 0)
        blob_pointer->Reshape(param.input_dim(top_id * 4), param.input_dim(top_id * 4 + 1), param.input_dim(top_id * 4 + 2), param.input_dim(top_id * 4 + 3));
      else
        blob_pointer->Reshape(param.input_shape(top_id));
      net_input_blob_indices_.push_back(blob_id);
      net_input_blobs_.push_back(blob_pointer.get());
    } else {
      top_id_vecs_[layer_id].push_back(blob_id);
      top_vecs_[layer_id].push_back(blob_pointer.get());
    }
  }
  if (available_blobs) available_blobs->insert(blob_name);
}

template <typename Dtype>
int Net<Dtype>::AppendBottom(const NetParameter& param, const int layer_id, const int bottom_id, set<string>* available_blobs,
                             map<string, int>* blob_name_to_idx) {
  const LayerParameter& layer_param = param.layer(layer_id);
  const string& blob_name = layer_param.bottom(bottom_id);
  if (available_blobs->find(blob_name) == available_blobs->end())
    LOG(FATAL) << "Unknown bottom blob '" << blob_name << "' (layer '" << layer_param.name() << "', bottom index " << bottom_id << ")";
  const int blob_id = (*blob_name_to_idx)[blob_name];
  LOG(INFO) << layer_names_[layer_id] << " <- " << blob_name;
  bottom_vecs_[layer_id].push_back(blobs_[blob_id].get());
  bottom_id_vecs_[layer_id].push_back(blob_id);
  available_blobs->erase(blob_name);
  return blob_id;
}

template <typename Dtype>
void Net<Dtype>::AppendParam(const NetParameter& param, const int layer_id, const int param_id) {
  // parameter sharing by ParamSpec.name is a training feature; every blob is its own owner here
  params_.push_back(layers_[layer_id]->blobs()[param_id]);
  learnable_params_.push_back(params_.back().get());
}

// --------------------------------------------------------------------------------------- forward
template <typename Dtype>
Dtype Net<Dtype>::ForwardFromTo(int start, int end) {
  CHECK_GE(start, 0);
  CHECK_LT(end, (int)layers_.size());
  Dtype loss = 0;
  fused_last_forward_ = false;
  if (debug_info_) debug_log_.clear();
  for (int i = start; i <= end; ++i) {
    // a per-layer (partial) forward after a fused one: its bottoms must be values the last forward really left in the blobs
    if (!blob_fresh_.empty()) {
      for (int b : bottom_id_vecs_[i])
        CHECK(blob_fresh_[b]) << "layer " << layer_names_[i] << " reads blob " << blob_names_[b] << ", which the last (fused) forward did not "
                              << "materialise: call materialize_intermediates(true) or set_fusion(false) before forwarding from the middle of the net";
      for (int t : top_id_vecs_[i]) blob_fresh_[t] = 1;
    }
    loss += layers_[i]->Forward(bottom_vecs_[i], top_vecs_[i]);
    if (debug_info_) ForwardDebugInfo(i);
  }
  return loss;
}

template <typename Dtype>
const vector<Blob<Dtype>*>& Net<Dtype>::ForwardPrefilled(Dtype* loss) {
  if (loss != NULL) *loss = 0;
  const long long launches_before = dc_launch_count();
  fused_last_forward_ = false;
  if (Caffe::mode() == Caffe::GPU && fusion_ && !debug_info_) {
    vector<vector<int> > shapes;
    for (Blob<Dtype>* b : net_input_blobs_) shapes.push_back(b->shape());
    if (plan_ == nullptr || shapes != plan_input_shapes_ || plan_->WeightsStale()) {
      delete plan_;
      plan_ = nullptr;
      Reshape();                       // propagate input shapes like Layer::Forward's per-call Reshape
      plan_ = FusedPlan::Build(*reinterpret_cast<Net<float>*>(this), materialize_, &fusion_diag_, &plan_weights_);
      plan_input_shapes_ = shapes;
      if (plan_ == nullptr) LOG(WARNING) << "fused plan unavailable (" << fusion_diag_ << "); running layer by layer";
    }
    if (plan_ != nullptr) {
      plan_->set_step_timing(step_timing_);
      plan_->Run();
      fused_last_forward_ = true;
      blob_fresh_ = plan_->WrittenBlobs();
      for (int b : net_input_blob_indices_) blob_fresh_[b] = 1;
      last_launches_ = dc_launch_count() - launches_before;
      return net_output_blobs_;
    }
  }
  ForwardFromTo(0, static_cast<int>(layers_.size()) - 1);
  last_launches_ = dc_launch_count() - launches_before;
  return net_output_blobs_;
}

template <typename Dtype>
const vector<Blob<Dtype>*>& Net<Dtype>::Forward(const vector<Blob<Dtype>*>& bottom, Dtype* loss) {
  CHECK_EQ(bottom.size(), net_input_blobs_.size());
  for (size_t i = 0; i < bottom.size(); ++i) net_input_blobs_[i]->CopyFrom(*bottom[i]);
  return ForwardPrefilled(loss);
}

template <typename Dtype>
void Net<Dtype>::ForwardDebugInfo(const int layer_id) {
  for (size_t top_id = 0; top_id < top_vecs_[layer_id].size(); ++top_id) {
    const Blob<Dtype>& blob = *top_vecs_[layer_id][top_id];
    const Dtype mean_abs = blob.count() ? blob.asum_data() / blob.count() : 0;
    LOG(INFO) << "    [Forward] Layer " << layer_names_[layer_id] << ", top blob " << blob_names_[top_id_vecs_[layer_id][top_id]] << " data: " << mean_abs;
    DebugRecord r = {layer_names_[layer_id], blob_names_[top_id_vecs_[layer_id][top_id]], static_cast<double>(mean_abs)};
    debug_log_.push_back(r);
  }
}

template <typename Dtype>
void Net<Dtype>::Reshape() {
  for (size_t i = 0; i < layers_.size(); ++i) layers_[i]->Reshape(bottom_vecs_[i], top_vecs_[i]);
}

template <typename Dtype> void Net<Dtype>::set_fusion(bool on) { fusion_ = on; }
template <typename Dtype> void Net<Dtype>::materialize_intermediates(bool on) { if (on != materialize_) { materialize_ = on; delete plan_; plan_ = nullptr; } }
template <typename Dtype> void Net<Dtype>::set_skipped_outputs(const vector<string>& blob_names) {
  set<string> want(blob_names.begin(), blob_names.end());
  for (const string& n : want) {
    bool is_output = false;
    for (int b : net_output_blob_indices_) is_output = is_output || blob_names_[b] == n;
    CHECK(is_output) << "set_skipped_outputs: '" << n << "' is not an output blob of this net";
  }
  if (want != skipped_outputs_) { skipped_outputs_.swap(want); delete plan_; plan_ = nullptr; }
}
template <typename Dtype> void Net<Dtype>::InvalidatePlan() { delete plan_; plan_ = nullptr; plan_weights_.reset(); }

// --------------------------------------------------------------------------------------- weights
template <typename Dtype>
void Net<Dtype>::CopyTrainedLayersFrom(const NetParameter& param) {
  const int num_source_layers = param.layer_size();
  for (int i = 0; i < num_source_layers; ++i) {
    const LayerParameter& source_layer = param.layer(i);
    const string& source_layer_name = source_layer.name();
    int target_layer_id = 0;
    while (target_layer_id != (int)layer_names_.size() && layer_names_[target_layer_id] != source_layer_name) ++target_layer_id;
    if (target_layer_id == (int)layer_names_.size()) { LOG(INFO) << "Ignoring source layer " << source_layer_name; continue; }
    vector<shared_ptr<Blob<Dtype> > >& target_blobs = layers_[target_layer_id]->blobs();
    CHECK_EQ((int)target_blobs.size(), source_layer.blobs_size()) << "Incompatible number of blobs for layer " << source_layer_name;
    for (size_t j = 0; j < target_blobs.size(); ++j) {
      if (!target_blobs[j]->ShapeEquals(source_layer.blobs(j))) {
        Blob<Dtype> source_blob;
        source_blob.FromProto(source_layer.blobs(j), true);
        LOG(FATAL) << "Cannot copy param " << j << " weights from layer '" << source_layer_name << "'; shape mismatch.  Source param shape is "
                   << source_blob.shape_string() << "; target param shape is " << target_blobs[j]->shape_string() << ". "
                   << "To learn this layer's parameters from scratch rather than copying from a saved net, rename the layer.";
      }
      target_blobs[j]->FromProto(source_layer.blobs(j), false);
    }
    layers_[target_layer_id]->OnWeightsChanged();
  }
  InvalidatePlan();
}

template <typename Dtype>
void Net<Dtype>::CopyTrainedLayersFrom(const string trained_filename) {
  const size_t n = trained_filename.size();
  CHECK(!(n >= 3 && trained_filename.compare(n - 3, 3, ".h5") == 0)) << "HDF5 weight files are not supported (no HDF5 in this build); use a binaryproto .caffemodel";
  CopyTrainedLayersFromBinaryProto(trained_filename);
}

template <typename Dtype>
void Net<Dtype>::CopyTrainedLayersFromBinaryProto(const string trained_filename) {
  NetParameter param;
  ReadNetParamsFromBinaryFileOrDie(trained_filename, &param);
  CopyTrainedLayersFrom(param);
}

template <typename Dtype>
void Net<Dtype>::ShareTrainedLayersWith(const Net* other) {
  const int num_source_layers = static_cast<int>(other->layers().size());
  for (int i = 0; i < num_source_layers; ++i) {
    Layer<Dtype>* source_layer = other->layers()[i].get();
    const string& source_layer_name = other->layer_names()[i];
    int target_layer_id = 0;
    while (target_layer_id != (int)layer_names_.size() && layer_names_[target_layer_id] != source_layer_name) ++target_layer_id;
    if (target_layer_id == (int)layer_names_.size()) continue;
    vector<shared_ptr<Blob<Dtype> > >& target_blobs = layers_[target_layer_id]->blobs();
    CHECK_EQ(target_blobs.size(), source_layer->blobs().size()) << "Incompatible number of blobs for layer " << source_layer_name;
    for (size_t j = 0; j < target_blobs.size(); ++j) {
      Blob<Dtype>* source_blob = source_layer->blobs()[j].get();
      CHECK(target_blobs[j]->shape() == source_blob->shape());
      target_blobs[j]->ShareData(*source_blob);
    }
    layers_[target_layer_id]->OnWeightsChanged();
  }
  InvalidatePlan();
}

template <typename Dtype>
void Net<Dtype>::ToProto(NetParameter* param, bool write_diff) const {
  param->Clear();
  param->set_name(name_);
  for (size_t i = 0; i < net_input_blob_indices_.size(); ++i) param->add_input(blob_names_[net_input_blob_indices_[i]]);
  for (size_t i = 0; i < layers_.size(); ++i) layers_[i]->ToProto(param->add_layer(), write_diff);
}

template <typename Dtype> bool Net<Dtype>::has_blob(const string& blob_name) const { return blob_names_index_.find(blob_name) != blob_names_index_.end(); }
template <typename Dtype>
const shared_ptr<Blob<Dtype> > Net<Dtype>::blob_by_name(const string& blob_name) const {
  shared_ptr<Blob<Dtype> > blob_ptr;
  if (has_blob(blob_name)) blob_ptr = blobs_[blob_names_index_.find(blob_name)->second];
  else LOG(WARNING) << "Unknown blob name " << blob_name;
  return blob_ptr;
}
template <typename Dtype> bool Net<Dtype>::has_layer(const string& layer_name) const { return layer_names_index_.find(layer_name) != layer_names_index_.end(); }
template <typename Dtype>
const shared_ptr<Layer<Dtype> > Net<Dtype>::layer_by_name(const string& layer_name) const {
  shared_ptr<Layer<Dtype> > layer_ptr;
  if (has_layer(layer_name)) layer_ptr = layers_[layer_names_index_.find(layer_name)->second];
  else LOG(WARNING) << "Unknown layer name " << layer_name;
  return layer_ptr;
}

INSTANTIATE_CLASS(Net);

}  // namespace caffe
