// Layer<Dtype>: the plugin interface the path sits behind (reference include/caffe/layer.hpp:37-445).
// Same virtuals and call protocol: SetUp = CheckBlobCounts + LayerSetUp + Reshape; Forward =
// Reshape + Forward_{cpu,gpu} by Caffe::mode().  Forward_gpu of the layer types in layers/ calls
// the C ABI; there is NO CPU implementation in the product (the CPU oracle lives in oracle/):
// the default Forward_cpu fails loudly.
#pragma once
#include <algorithm>

#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/proto/caffe.pb.h"

namespace caffe {

template <typename Dtype>
class Layer {
 public:
  explicit Layer(const LayerParameter& param) : layer_param_(param) {
    phase_ = param.phase();
    if (layer_param_.blobs_size() > 0) {
      blobs_.resize(layer_param_.blobs_size());
      for (int i = 0; i < layer_param_.blobs_size(); ++i) {
        blobs_[i].reset(new Blob<Dtype>());
        blobs_[i]->FromProto(layer_param_.blobs(i));
      }
    }
  }
  virtual ~Layer() {}

  void SetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    CheckBlobCounts(bottom, top);
    LayerSetUp(bottom, top);
    Reshape(bottom, top);
  }
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) = 0;
  inline Dtype Forward(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);

  vector<shared_ptr<Blob<Dtype> > >& blobs() { return blobs_; }
  const LayerParameter& layer_param() const { return layer_param_; }
  virtual void ToProto(LayerParameter* param, bool write_diff = false);
  virtual inline const char* type() const { return ""; }
  virtual inline int ExactNumBottomBlobs() const { return -1; }
  virtual inline int MinBottomBlobs() const { return -1; }
  virtual inline int MaxBottomBlobs() const { return -1; }
  virtual inline int ExactNumTopBlobs() const { return -1; }
  virtual inline int MinTopBlobs() const { return -1; }
  virtual inline int MaxTopBlobs() const { return -1; }
  virtual inline bool EqualNumBottomTopBlobs() const { return false; }
  virtual inline bool AutoTopBlobs() const { return false; }
  Phase phase() const { return phase_; }
  // Layers whose value is a weight-only function of the blob values (everything on this path) need
  // to know when CopyTrainedLayersFrom / a host write changed blobs_ so cached device-side packs
  // are rebuilt.
  virtual void OnWeightsChanged() {}

 protected:
  LayerParameter layer_param_;
  Phase phase_;
  vector<shared_ptr<Blob<Dtype> > > blobs_;

  virtual void Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    LOG(FATAL) << "Layer " << layer_param_.name() << " (" << type() << "): deepcut-cnn_b200 has no CPU forward path; "
               << "call Caffe::set_mode(Caffe::GPU) (the CPU oracle lives in oracle/, outside the product).";
  }
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    return Forward_cpu(bottom, top);
  }
  virtual void CheckBlobCounts(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    if (ExactNumBottomBlobs() >= 0) CHECK_EQ(ExactNumBottomBlobs(), (int)bottom.size()) << type() << " Layer takes " << ExactNumBottomBlobs() << " bottom blob(s) as input.";
    if (MinBottomBlobs() >= 0) CHECK_LE(MinBottomBlobs(), (int)bottom.size()) << type() << " Layer takes at least " << MinBottomBlobs() << " bottom blob(s) as input.";
    if (MaxBottomBlobs() >= 0) CHECK_GE(MaxBottomBlobs(), (int)bottom.size()) << type() << " Layer takes at most " << MaxBottomBlobs() << " bottom blob(s) as input.";
    if (ExactNumTopBlobs() >= 0) CHECK_EQ(ExactNumTopBlobs(), (int)top.size()) << type() << " Layer produces " << ExactNumTopBlobs() << " top blob(s) as output.";
    if (MinTopBlobs() >= 0) CHECK_LE(MinTopBlobs(), (int)top.size()) << type() << " Layer produces at least " << MinTopBlobs() << " top blob(s) as output.";
    if (MaxTopBlobs() >= 0) CHECK_GE(MaxTopBlobs(), (int)top.size()) << type() << " Layer produces at most " << MaxTopBlobs() << " top blob(s) as output.";
    if (EqualNumBottomTopBlobs()) CHECK_EQ(bottom.size(), top.size()) << type() << " Layer produces one top blob as output for each bottom blob input.";
  }
  DISABLE_COPY_AND_ASSIGN(Layer);
};

template <typename Dtype>
inline Dtype Layer<Dtype>::Forward(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  Reshape(bottom, top);     // the reference re-runs Reshape on every forward (layer.hpp:455)
  switch (Caffe::mode()) {
    case Caffe::CPU: Forward_cpu(bottom, top); break;
    case Caffe::GPU: Forward_gpu(bottom, top); break;
    default: LOG(FATAL) << "Unknown caffe mode.";
  }
  return Dtype(0);          // no loss layers on the inference path
}

template <typename Dtype>
void Layer<Dtype>::ToProto(LayerParameter* param, bool write_diff) {
  param->Clear();
  param->CopyFrom(layer_param_);
  param->clear_blobs();
  for (size_t i = 0; i < blobs_.size(); ++i) blobs_[i]->ToProto(param->add_blobs(), write_diff);
}

}  // namespace caffe
