#include "caffe/util/insert_splits.hpp"

#include <map>
#include <sstream>
#include <utility>

#include "caffe/common.hpp"

namespace caffe {

typedef std::pair<int, int> TopRef;   // (producing layer index or -1 for a net input, top index)

void InsertSplits(const NetParameter& param, NetParameter* param_split) {
  param_split->CopyFrom(param);
  param_split->clear_layer();
  map<string, TopRef> latest_producer;             // blob name -> most recent top that wrote it
  map<pair<int, int>, TopRef> bottom_to_producer;  // (layer, bottom idx) -> top it reads
  map<TopRef, int> uses;                           // how many bottoms read each top
  map<TopRef, float> top_loss_weight;
  map<TopRef, int> issued;                         // split outputs handed out so far
  map<int, string> layer_name_of;
  layer_name_of[-1] = "input";

  for (int i = 0; i < param.input_size(); ++i) latest_producer[param.input(i)] = TopRef(-1, i);
  for (int i = 0; i < param.layer_size(); ++i) {
    const LayerParameter& lp = param.layer(i);
    layer_name_of[i] = lp.name();
    for (int j = 0; j < lp.bottom_size(); ++j) {
      const string& blob = lp.bottom(j);
      CHECK(latest_producer.count(blob)) << "Unknown bottom blob '" << blob << "' (layer '" << lp.name() << "', bottom index " << j << ")";
      const TopRef prod = latest_producer[blob];
      bottom_to_producer[std::make_pair(i, j)] = prod;
      ++uses[prod];
    }
    for (int j = 0; j < lp.top_size(); ++j) latest_producer[lp.top(j)] = TopRef(i, j);
    const int nloss = std::min(lp.loss_weight_size(), lp.top_size());
    for (int j = 0; j < nloss; ++j) {
      const TopRef t(i, j);
      top_loss_weight[t] = lp.loss_weight(j);
      if (top_loss_weight[t]) ++uses[t];
    }
  }
  // net inputs read more than once
  for (int i = 0; i < param.input_size(); ++i) {
    const int n = uses[TopRef(-1, i)];
    if (n > 1) ConfigureSplitLayer(layer_name_of[-1], param.input(i), i, n, 0.f, param_split->add_layer());
  }
  for (int i = 0; i < param.layer_size(); ++i) {
    LayerParameter* lp = param_split->add_layer();
    const int lp_index = param_split->layer_size() - 1;
    lp->CopyFrom(param.layer(i));
    for (int j = 0; j < lp->bottom_size(); ++j) {
      const TopRef prod = bottom_to_producer[std::make_pair(i, j)];
      if (uses[prod] > 1) {
        const string& prod_layer = layer_name_of[prod.first];
        lp->set_bottom(j, SplitBlobName(prod_layer, lp->bottom(j), prod.second, issued[prod]++));
      }
    }
    for (int j = 0; j < lp->top_size(); ++j) {
      const TopRef t(i, j);
      const int n = uses[t];
      if (n > 1) {
        const float lw = top_loss_weight.count(t) ? top_loss_weight[t] : 0.f;
        // param_split->add_layer() may reallocate: read what we need from lp first
        const string lname = lp->name(), bname = lp->top(j);
        LayerParameter* split = param_split->add_layer();
        lp = param_split->mutable_layer(lp_index);
        ConfigureSplitLayer(lname, bname, j, n, lw, split);
        if (lw) { lp->clear_loss_weight(); ++issued[t]; }
      }
    }
  }
}

void ConfigureSplitLayer(const string& layer_name, const string& blob_name, const int blob_idx, const int split_count,
                         const float loss_weight, LayerParameter* split_layer_param) {
  split_layer_param->Clear();
  split_layer_param->add_bottom(blob_name);
  split_layer_param->set_name(SplitLayerName(layer_name, blob_name, blob_idx));
  split_layer_param->set_type("Split");
  for (int k = 0; k < split_count; ++k) {
    split_layer_param->add_top(SplitBlobName(layer_name, blob_name, blob_idx, k));
    if (loss_weight) split_layer_param->add_loss_weight(k == 0 ? loss_weight : 0.f);
  }
}

string SplitLayerName(const string& layer_name, const string& blob_name, const int blob_idx) {
  std::ostringstream s;
  s << blob_name << "_" << layer_name << "_" << blob_idx << "_split";
  return s.str();
}

string SplitBlobName(const string& layer_name, const string& blob_name, const int blob_idx, const int split_idx) {
  std::ostringstream s;
  s << blob_name << "_" << layer_name << "_" << blob_idx << "_split_" << split_idx;
  return s.str();
}

}  // namespace caffe
