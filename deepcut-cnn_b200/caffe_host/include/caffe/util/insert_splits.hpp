// InsertSplits (reference src/caffe/util/insert_splits.cpp:12-143): a top consumed by more than
// one bottom gets a Split layer; consumers are renamed <blob>_<layer>_<top idx>_split_<k>.
#pragma once
#include <string>

#include "caffe/proto/caffe.pb.h"

namespace caffe {

void InsertSplits(const NetParameter& param, NetParameter* param_split);
void ConfigureSplitLayer(const std::string& layer_name, const std::string& blob_name, const int blob_idx,
                         const int split_count, const float loss_weight, LayerParameter* split_layer_param);
std::string SplitLayerName(const std::string& layer_name, const std::string& blob_name, const int blob_idx);
std::string SplitBlobName(const std::string& layer_name, const std::string& blob_name, const int blob_idx, const int split_idx);

}  // namespace caffe
