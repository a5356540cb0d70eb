// The layer types the DeeperCut deploy net instantiates (10 of the reference's 66),
// with the reference's SetUp/Reshape semantics and blob orders, and Forward_gpu through the C ABI.
//   Convolution / Deconvolution  base_conv_layer.cpp:14-254, conv_layer.cpp:8-22, deconv_layer.cpp:8-22
//   BatchNorm                    batch_norm_layer.cpp:10-72
//   Scale (+ owned Bias)         scale_layer.cpp:13-106, bias_layer.cpp:11-70
//   ReLU, Sigmoid                relu_layer.cpp, sigmoid_layer.cpp (neuron_layer.cpp:8-12 reshape)
//   Eltwise                      eltwise_layer.cpp:11-43
//   Pooling                      pooling_layer.cpp:16-123
//   Crop (DeepCut's own)         crop_layer.cpp:14-34
//   Split                        split_layer.cpp:9-24
#pragma once
#include "caffe/blob.hpp"
#include "caffe/layer.hpp"
#include "caffe/proto/caffe.pb.h"

namespace caffe {

// Device-side cache of a layer's transformed weights (packed split-fp16 matrix + per-row scale),
// rebuilt when the weight blobs change.
struct PackedWeights {
  void* w = nullptr;        // device, dc_pack_*_weight layout
  float* scale = nullptr;   // device [rows]
  float* shift = nullptr;   // device [rows]
  int rows = 0;
  bool valid = false;
  unsigned long long epoch = 0;   // host_write_epoch of the weight blob when packed
  ~PackedWeights();
  void Release();
};

template <typename Dtype>
class BaseConvolutionLayer : public Layer<Dtype> {
 public:
  explicit BaseConvolutionLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  inline int MinBottomBlobs() const override { return 1; }
  inline int MinTopBlobs() const override { return 1; }
  inline bool EqualNumBottomTopBlobs() const override { return true; }
  void OnWeightsChanged() override { packed_.valid = false; }
  // geometry accessors used by the fused planner
  int kernel_h() const { return kernel_h_; }
  int kernel_w() const { return kernel_w_; }
  int stride_h() const { return stride_h_; }
  int stride_w() const { return stride_w_; }
  int pad_h() const { return pad_h_; }
  int pad_w() const { return pad_w_; }
  int dilation_h() const { return dilation_h_; }
  int dilation_w() const { return dilation_w_; }
  int num_output() const { return num_output_; }
  int channels() const { return channels_; }
  int group() const { return group_; }
  bool bias_term() const { return bias_term_; }

 protected:
  virtual bool reverse_dimensions() = 0;
  virtual void compute_output_shape() = 0;
  int kernel_h_ = 0, kernel_w_ = 0, stride_h_ = 1, stride_w_ = 1, pad_h_ = 0, pad_w_ = 0, dilation_h_ = 1, dilation_w_ = 1;
  int num_ = 0, channels_ = 0, group_ = 1, num_output_ = 0, height_ = 0, width_ = 0, out_h_ = 0, out_w_ = 0;
  bool bias_term_ = false, is_1x1_ = false;
  PackedWeights packed_;
};

template <typename Dtype>
class ConvolutionLayer : public BaseConvolutionLayer<Dtype> {
 public:
  explicit ConvolutionLayer(const LayerParameter& param) : BaseConvolutionLayer<Dtype>(param) {}
  inline const char* type() const override { return "Convolution"; }

 protected:
  void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  inline bool reverse_dimensions() override { return false; }
  void compute_output_shape() override;
};

template <typename Dtype>
class DeconvolutionLayer : public BaseConvolutionLayer<Dtype> {
 public:
  explicit DeconvolutionLayer(const LayerParameter& param) : BaseConvolutionLayer<Dtype>(param) {}
  inline const char* type() const override { return "Deconvolution"; }

 protected:
  void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  inline bool reverse_dimensions() override { return true; }
  void compute_output_shape() override;
};

template <typename Dtype>
class BatchNormLayer : public Layer<Dtype> {
 public:
  explicit BatchNormLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  inline const char* type() const override { return "BatchNorm"; }
  inline int ExactNumBottomBlobs() const override { return 1; }
  inline int ExactNumTopBlobs() const override { return 1; }
  bool use_global_stats() const { return use_global_stats_; }
  Dtype eps() const { return eps_; }

 protected:
  void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  bool use_global_stats_ = true;
  Dtype eps_ = Dtype(1e-5);
  int channels_ = 0;
  Blob<Dtype> mean_, inv_std_;   // per-channel scratch for the per-layer path
};

template <typename Dtype>
class ScaleLayer : public Layer<Dtype> {
 public:
  explicit ScaleLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  inline const char* type() const override { return "Scale"; }
  inline int MinBottomBlobs() const override { return 1; }
  inline int MaxBottomBlobs() const override { return 2; }
  inline int ExactNumTopBlobs() const override { return 1; }
  bool has_bias() const { return bias_term_; }

 protected:
  void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  bool bias_term_ = false;
  int axis_ = 1, outer_dim_ = 0, scale_dim_ = 0, inner_dim_ = 0;
};

template <typename Dtype>
class NeuronLayer : public Layer<Dtype> {
 public:
  explicit NeuronLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override { top[0]->ReshapeLike(*bottom[0]); }
  inline int ExactNumBottomBlobs() const override { return 1; }
  inline int ExactNumTopBlobs() const override { return 1; }
};

template <typename Dtype>
class ReLULayer : public NeuronLayer<Dtype> {
 public:
  explicit ReLULayer(const LayerParameter& param) : NeuronLayer<Dtype>(param) {}
  inline const char* type() const override { return "ReLU"; }

 protected:
  void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
};

template <typename Dtype>
class SigmoidLayer : public NeuronLayer<Dtype> {
 public:
  explicit SigmoidLayer(const LayerParameter& param) : NeuronLayer<Dtype>(param) {}
  inline const char* type() const override { return "Sigmoid"; }

 protected:
  void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
};

template <typename Dtype>
class EltwiseLayer : public Layer<Dtype> {
 public:
  explicit EltwiseLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  inline const char* type() const override { return "Eltwise"; }
  inline int MinBottomBlobs() const override { return 2; }
  inline int ExactNumTopBlobs() const override { return 1; }
  EltwiseParameter_EltwiseOp op() const { return op_; }
  const vector<Dtype>& coeffs() const { return coeffs_; }

 protected:
  void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  EltwiseParameter_EltwiseOp op_ = EltwiseParameter_EltwiseOp_SUM;
  vector<Dtype> coeffs_;
};

template <typename Dtype>
class PoolingLayer : public Layer<Dtype> {
 public:
  explicit PoolingLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  inline const char* type() const override { return "Pooling"; }
  inline int ExactNumBottomBlobs() const override { return 1; }
  inline int MinTopBlobs() const override { return 1; }
  inline int MaxTopBlobs() const override { return 1; }
  int kernel_h() const { return kernel_h_; }
  int kernel_w() const { return kernel_w_; }
  int stride_h() const { return stride_h_; }
  int stride_w() const { return stride_w_; }
  int pad_h() const { return pad_h_; }
  int pad_w() const { return pad_w_; }
  PoolingParameter_PoolMethod method() const { return this->layer_param_.pooling_param().pool(); }

 protected:
  void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  int kernel_h_ = 0, kernel_w_ = 0, stride_h_ = 1, stride_w_ = 1, pad_h_ = 0, pad_w_ = 0;
  int channels_ = 0, height_ = 0, width_ = 0, pooled_height_ = 0, pooled_width_ = 0;
  bool global_pooling_ = false;
};

template <typename Dtype>
class CropLayer : public Layer<Dtype> {
 public:
  explicit CropLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  inline const char* type() const override { return "Crop"; }
  inline int ExactNumBottomBlobs() const override { return 2; }
  inline int ExactNumTopBlobs() const override { return 1; }
  int crop_h() const { return crop_h_; }
  int crop_w() const { return crop_w_; }

 protected:
  void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  int crop_h_ = 0, crop_w_ = 0;
};

template <typename Dtype>
class SplitLayer : public Layer<Dtype> {
 public:
  explicit SplitLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  inline const char* type() const override { return "Split"; }
  inline int ExactNumBottomBlobs() const override { return 1; }
  inline int MinTopBlobs() const override { return 1; }

 protected:
  void Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override;
  void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) override { Forward_cpu(bottom, top); }
};

}  // namespace caffe
