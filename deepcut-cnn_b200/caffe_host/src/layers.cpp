// Layer implementations: reference SetUp/Reshape semantics + Forward_gpu through the C ABI.
#include "caffe/layers/dc_layers.hpp"

#include <cfloat>
#include <cmath>

#include "caffe/filler.hpp"
#include "caffe/layer_factory.hpp"
#include "deepcut_b200.h"

namespace caffe {

PackedWeights::~PackedWeights() { Release(); }
void PackedWeights::Release() {
  if (w) dc_free(w);
  if (scale) dc_free(scale);
  if (shift) dc_free(shift);
  w = nullptr; scale = shift = nullptr; rows = 0; valid = false;
}

namespace {
// Device pointer of a top blob every element of which the layer writes: skips the reference's
// upload-before-write unless the layer runs in place (then the bottom's data must survive).
float* TopPtr(Blob<float>* top, const Blob<float>* bottom) {
  return top == bottom ? top->mutable_gpu_data() : top->overwrite_gpu_data();
}
// scratch device buffer per thread for the per-layer conv path (split-NHWC staging)
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
  void* Get(size_t bytes) {
    if (bytes > cap) {
      if (p) dc_free(p);
      DC_CHECK(dc_malloc(&p, bytes));
      cap = bytes;
    }
    return p;
  }
  ~Scratch() { if (p) dc_free(p); }
};
Scratch& ScratchA() { static thread_local Scratch s; return s; }
Scratch& ScratchB() { static thread_local Scratch s; return s; }
}  // namespace

// ===================================================================== Convolution / Deconvolution
template <typename Dtype>
void BaseConvolutionLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const ConvolutionParameter& cp = this->layer_param_.convolution_param();
  CHECK_EQ(bottom[0]->num_axes(), 4) << "deepcut-cnn_b200 convolutions are 2-D (4-D blobs)";
  CHECK_EQ(cp.axis(), 1) << "only channel axis 1 is supported";
  // kernel / stride / pad / dilation: either the repeated field (1 or 2 entries) or the _h/_w pair
  if (cp.has_kernel_h() || cp.has_kernel_w()) {
    CHECK_EQ(0, cp.kernel_size_size()) << "Either kernel_size or kernel_h/w should be specified; not both.";
    kernel_h_ = cp.kernel_h(); kernel_w_ = cp.kernel_w();
  } else {
    const int n = cp.kernel_size_size();
    CHECK(n == 1 || n == 2) << "kernel_size must be specified once, or once per spatial dimension (kernel_size specified " << n << " times; 2 spatial dims).";
    kernel_h_ = cp.kernel_size(0); kernel_w_ = cp.kernel_size(n == 1 ? 0 : 1);
  }
  CHECK_GT(kernel_h_, 0) << "Filter dimensions must be nonzero.";
  CHECK_GT(kernel_w_, 0) << "Filter dimensions must be nonzero.";
  if (cp.has_stride_h() || cp.has_stride_w()) {
    CHECK_EQ(0, cp.stride_size()) << "Either stride or stride_h/w should be specified; not both.";
    stride_h_ = cp.stride_h(); stride_w_ = cp.stride_w();
  } else {
    const int n = cp.stride_size();
    CHECK(n == 0 || n == 1 || n == 2) << "stride must be specified once, or once per spatial dimension";
    stride_h_ = n == 0 ? 1 : cp.stride(0); stride_w_ = n == 0 ? 1 : cp.stride(n == 1 ? 0 : 1);
  }
  CHECK_GT(stride_h_, 0); CHECK_GT(stride_w_, 0);
  if (cp.has_pad_h() || cp.has_pad_w()) {
    CHECK_EQ(0, cp.pad_size()) << "Either pad or pad_h/w should be specified; not both.";
    pad_h_ = cp.pad_h(); pad_w_ = cp.pad_w();
  } else {
    const int n = cp.pad_size();
    CHECK(n == 0 || n == 1 || n == 2) << "pad must be specified once, or once per spatial dimension";
    pad_h_ = n == 0 ? 0 : cp.pad(0); pad_w_ = n == 0 ? 0 : cp.pad(n == 1 ? 0 : 1);
  }
  {
    const int n = cp.dilation_size();
    CHECK(n == 0 || n == 1 || n == 2) << "dilation must be specified once, or once per spatial dimension";
    dilation_h_ = n == 0 ? 1 : cp.dilation(0); dilation_w_ = n == 0 ? 1 : cp.dilation(n == 1 ? 0 : 1);
  }
  is_1x1_ = kernel_h_ == 1 && kernel_w_ == 1 && stride_h_ == 1 && stride_w_ == 1 && pad_h_ == 0 && pad_w_ == 0;
  channels_ = bottom[0]->shape(1);
  num_output_ = cp.num_output();
  CHECK_GT(num_output_, 0);
  group_ = cp.group();
  CHECK_EQ(group_, 1) << "grouped convolution is outside the DeeperCut path (SURVEY section 8): group must be 1";
  const int conv_out = reverse_dimensions() ? channels_ : num_output_;
  const int conv_in = reverse_dimensions() ? num_output_ : channels_;
  vector<int> weight_shape = {conv_out, conv_in / group_, kernel_h_, kernel_w_};
  bias_term_ = cp.bias_term();
  vector<int> bias_shape(bias_term_, num_output_);
  if (this->blobs_.size() > 0) {
    CHECK_EQ(1 + bias_term_, (int)this->blobs_.size()) << "Incorrect number of weight blobs.";
    if (weight_shape != this->blobs_[0]->shape()) {
      Blob<Dtype> expected(weight_shape);
      LOG(FATAL) << "Incorrect weight shape: expected shape " << expected.shape_string() << "; instead, shape was " << this->blobs_[0]->shape_string();
    }
    if (bias_term_ && bias_shape != this->blobs_[1]->shape()) {
      Blob<Dtype> expected(bias_shape);
      LOG(FATAL) << "Incorrect bias shape: expected shape " << expected.shape_string() << "; instead, shape was " << this->blobs_[1]->shape_string();
    }
  } else {
    this->blobs_.resize(bias_term_ ? 2 : 1);
    this->blobs_[0].reset(new Blob<Dtype>(weight_shape));
    shared_ptr<Filler<Dtype> > wf(GetFiller<Dtype>(cp.weight_filler()));
    wf->Fill(this->blobs_[0].get());
    if (bias_term_) {
      this->blobs_[1].reset(new Blob<Dtype>(bias_shape));
      shared_ptr<Filler<Dtype> > bf(GetFiller<Dtype>(cp.bias_filler()));
      bf->Fill(this->blobs_[1].get());
    }
  }
}

template <typename Dtype>
void BaseConvolutionLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(bottom[0]->num_axes(), 4) << "bottom[0] must have 4 axes";
  CHECK_EQ(bottom[0]->shape(1), channels_) << "Input size incompatible with convolution kernel.";
  for (size_t i = 1; i < bottom.size(); ++i) CHECK(bottom[0]->shape() == bottom[i]->shape()) << "All inputs must have the same shape.";
  num_ = bottom[0]->shape(0);
  height_ = bottom[0]->shape(2);
  width_ = bottom[0]->shape(3);
  compute_output_shape();
  for (size_t i = 0; i < top.size(); ++i) top[i]->Reshape(num_, num_output_, out_h_, out_w_);
}

template <typename Dtype>
void ConvolutionLayer<Dtype>::compute_output_shape() {
  this->out_h_ = (this->height_ + 2 * this->pad_h_ - (this->dilation_h_ * (this->kernel_h_ - 1) + 1)) / this->stride_h_ + 1;
  this->out_w_ = (this->width_ + 2 * this->pad_w_ - (this->dilation_w_ * (this->kernel_w_ - 1) + 1)) / this->stride_w_ + 1;
}
template <typename Dtype>
void DeconvolutionLayer<Dtype>::compute_output_shape() {
  this->out_h_ = this->stride_h_ * (this->height_ - 1) + this->dilation_h_ * (this->kernel_h_ - 1) + 1 - 2 * this->pad_h_;
  this->out_w_ = this->stride_w_ * (this->width_ - 1) + this->dilation_w_ * (this->kernel_w_ - 1) + 1 - 2 * this->pad_w_;
}

template <typename Dtype>
void ConvolutionLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  void* st = Caffe::stream();
  const bool tensor_ok = this->channels_ % 64 == 0 && this->num_output_ % 32 == 0 && this->stride_h_ == 1 && this->stride_w_ == 1 &&
                         this->pad_h_ == this->pad_w_ && this->dilation_h_ == this->dilation_w_ && this->kernel_h_ * this->kernel_w_ <= 9;
  const float* bias = this->bias_term_ ? reinterpret_cast<const float*>(this->blobs_[1]->gpu_data()) : nullptr;
  if (!tensor_ok) {
    for (size_t i = 0; i < bottom.size(); ++i)
      DC_CHECK(dc_conv_direct_nchw(reinterpret_cast<const float*>(bottom[i]->gpu_data()), reinterpret_cast<const float*>(this->blobs_[0]->gpu_data()), bias,
                                   this->num_, this->channels_, this->height_, this->width_, this->num_output_, this->kernel_h_, this->kernel_w_,
                                   this->stride_h_, this->stride_w_, this->pad_h_, this->pad_w_, this->dilation_h_, this->dilation_w_,
                                   TopPtr(top[i], bottom[i]), st));
    return;
  }
  PackedWeights& pk = this->packed_;
  this->blobs_[0]->cpu_data();
  unsigned long long epoch = this->blobs_[0]->data()->host_write_epoch();
  if (this->bias_term_) { this->blobs_[1]->cpu_data(); epoch += this->blobs_[1]->data()->host_write_epoch(); }
  if (!pk.valid || pk.epoch != epoch) {
    pk.epoch = epoch;
    pk.Release();
    const int rows = dc_packed_rows(this->num_output_);
    const size_t K = static_cast<size_t>(this->kernel_h_) * this->kernel_w_ * this->channels_;
    vector<uint16_t> host(2 * rows * K);
    vector<float> scale(rows), shift(rows, 0.f);
    DC_CHECK(dc_pack_conv_weight(reinterpret_cast<const float*>(this->blobs_[0]->cpu_data()), this->num_output_, this->channels_, this->kernel_h_,
                                 this->kernel_w_, host.data(), scale.data()));
    if (this->bias_term_) for (int c = 0; c < this->num_output_; ++c) shift[c] = this->blobs_[1]->cpu_data()[c];
    DC_CHECK(dc_malloc(&pk.w, host.size() * 2));
    DC_CHECK(dc_malloc(reinterpret_cast<void**>(&pk.scale), rows * 4));
    DC_CHECK(dc_malloc(reinterpret_cast<void**>(&pk.shift), rows * 4));
    DC_CHECK(dc_memcpy_async(pk.w, host.data(), host.size() * 2, DC_H2D, st));
    DC_CHECK(dc_memcpy_async(pk.scale, scale.data(), rows * 4, DC_H2D, st));
    DC_CHECK(dc_memcpy_async(pk.shift, shift.data(), rows * 4, DC_H2D, st));
    DC_CHECK(dc_stream_sync(st));
    pk.rows = rows;
    pk.valid = true;
  }
  for (size_t i = 0; i < bottom.size(); ++i) {
    const size_t in_elems = static_cast<size_t>(bottom[i]->count());
    const size_t out_elems = static_cast<size_t>(top[i]->count());
    void* xs = ScratchA().Get(in_elems * 4);
    void* ys = ScratchB().Get(out_elems * 4);
    DC_CHECK(dc_nchw_to_split(reinterpret_cast<const float*>(bottom[i]->gpu_data()), this->num_, this->channels_, this->height_, this->width_, xs, st));
    dc_conv_args a;
    memset(&a, 0, sizeof(a));
    a.x = xs; a.n = this->num_; a.h = this->height_; a.w = this->width_; a.cin = this->channels_;
    a.cout = this->num_output_; a.kh = this->kernel_h_; a.kw = this->kernel_w_; a.pad = this->pad_h_; a.dilation = this->dilation_h_;
    a.w_packed = pk.w; a.scale = pk.scale; a.shift = pk.shift; a.out = ys;
    DC_CHECK(dc_conv_forward(&a, st));
    DC_CHECK(dc_split_to_nchw(ys, this->num_, this->num_output_, this->out_h_, this->out_w_, TopPtr(top[i], bottom[i]), st));
  }
}

template <typename Dtype>
void DeconvolutionLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const float* bias = this->bias_term_ ? reinterpret_cast<const float*>(this->blobs_[1]->gpu_data()) : nullptr;
  for (size_t i = 0; i < bottom.size(); ++i)
    DC_CHECK(dc_deconv_direct_nchw(reinterpret_cast<const float*>(bottom[i]->gpu_data()), reinterpret_cast<const float*>(this->blobs_[0]->gpu_data()), bias,
                                   this->num_, this->channels_, this->height_, this->width_, this->num_output_, this->kernel_h_, this->kernel_w_,
                                   this->stride_h_, this->stride_w_, this->pad_h_, this->pad_w_, this->dilation_h_, this->dilation_w_,
                                   TopPtr(top[i], bottom[i]), Caffe::stream()));
}

// ===================================================================== BatchNorm
template <typename Dtype>
void BatchNormLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const BatchNormParameter& param = this->layer_param_.batch_norm_param();
  use_global_stats_ = this->phase_ == TEST;
  if (param.has_use_global_stats()) use_global_stats_ = param.use_global_stats();
  CHECK(use_global_stats_) << "BatchNorm layer " << this->layer_param_.name() << ": batch statistics (training mode) are outside the inference path";
  channels_ = bottom[0]->num_axes() == 1 ? 1 : bottom[0]->shape(1);
  eps_ = param.eps();
  if (this->blobs_.size() > 0) {
    CHECK_EQ(this->blobs_.size(), 3u);
  } else {
    this->blobs_.resize(3);
    vector<int> sz = {channels_};
    this->blobs_[0].reset(new Blob<Dtype>(sz));
    this->blobs_[1].reset(new Blob<Dtype>(sz));
    sz[0] = 1;
    this->blobs_[2].reset(new Blob<Dtype>(sz));
    for (int i = 0; i < 3; ++i) {
      Dtype* d = this->blobs_[i]->mutable_cpu_data();
      for (int j = 0; j < this->blobs_[i]->count(); ++j) d[j] = 0;
    }
  }
}

template <typename Dtype>
void BatchNormLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  if (bottom[0]->num_axes() >= 1) CHECK_EQ(bottom[0]->num_axes() == 1 ? 1 : bottom[0]->shape(1), channels_);
  top[0]->ReshapeLike(*bottom[0]);
  vector<int> sz = {channels_};
  mean_.Reshape(sz);
  inv_std_.Reshape(sz);
}

template <typename Dtype>
void BatchNormLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  // stats * (factor == 0 ? 0 : 1/factor); sqrt(var + eps) via pow(.,0.5)  (batch_norm_layer.cpp:86-93,137-140)
  const Dtype f = this->blobs_[2]->cpu_data()[0];
  const Dtype sf = f == 0 ? Dtype(0) : Dtype(1) / f;
  Dtype* m = mean_.mutable_cpu_data();
  Dtype* s = inv_std_.mutable_cpu_data();
  for (int c = 0; c < channels_; ++c) {
    m[c] = this->blobs_[0]->cpu_data()[c] * sf;
    s[c] = std::pow(this->blobs_[1]->cpu_data()[c] * sf + eps_, Dtype(0.5));
  }
  const int n = bottom[0]->shape(0);
  const int hw = bottom[0]->count() / (n * channels_);
  DC_CHECK(dc_bn_forward_nchw(reinterpret_cast<const float*>(bottom[0]->gpu_data()), reinterpret_cast<const float*>(mean_.gpu_data()),
                              reinterpret_cast<const float*>(inv_std_.gpu_data()), n, channels_, hw,
                              TopPtr(top[0], bottom[0]), Caffe::stream()));
}

// ===================================================================== Scale (+bias)
template <typename Dtype>
void ScaleLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const ScaleParameter& param = this->layer_param_.scale_param();
  CHECK_EQ(bottom.size(), 1u) << "Scale with the multiplier as a second bottom is outside the DeeperCut path";
  axis_ = bottom[0]->CanonicalAxisIndex(param.axis());
  const int num_axes = param.num_axes();
  CHECK_GE(num_axes, -1) << "num_axes must be non-negative, or -1 to extend to the end of bottom[0]";
  bias_term_ = param.bias_term();
  if (this->blobs_.size() > 0) {
    CHECK_EQ(this->blobs_.size(), bias_term_ ? 2u : 1u);
  } else {
    if (num_axes >= 0) CHECK_GE(bottom[0]->num_axes(), axis_ + num_axes) << "scale blob's shape extends past bottom[0]'s shape when applied starting with bottom[0] axis = " << axis_;
    const vector<int>& bs = bottom[0]->shape();
    vector<int> scale_shape(bs.begin() + axis_, num_axes == -1 ? bs.end() : bs.begin() + axis_ + num_axes);
    this->blobs_.resize(bias_term_ ? 2 : 1);
    this->blobs_[0].reset(new Blob<Dtype>(scale_shape));
    FillerParameter fp(param.filler());
    if (!param.has_filler()) { fp.set_type("constant"); fp.set_value(1); }   // default: identity
    shared_ptr<Filler<Dtype> > filler(GetFiller<Dtype>(fp));
    filler->Fill(this->blobs_[0].get());
    if (bias_term_) {     // the reference owns a BiasLayer whose blob is shared as blobs_[1] (scale_layer.cpp:44-63)
      this->blobs_[1].reset(new Blob<Dtype>(scale_shape));
      shared_ptr<Filler<Dtype> > bf(GetFiller<Dtype>(param.bias_filler()));
      bf->Fill(this->blobs_[1].get());
    }
  }
}

template <typename Dtype>
void ScaleLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  Blob<Dtype>* scale = this->blobs_[0].get();
  CHECK_GE(bottom[0]->num_axes(), axis_ + scale->num_axes()) << "scale blob's shape extends past bottom[0]'s shape when applied starting with bottom[0] axis = " << axis_;
  for (int i = 0; i < scale->num_axes(); ++i)
    CHECK_EQ(bottom[0]->shape(axis_ + i), scale->shape(i)) << "dimension mismatch between bottom[0]->shape(" << axis_ + i << ") and scale->shape(" << i << ")";
  outer_dim_ = bottom[0]->count(0, axis_);
  scale_dim_ = scale->count();
  inner_dim_ = bottom[0]->count(axis_ + scale->num_axes());
  top[0]->ReshapeLike(*bottom[0]);
}

template <typename Dtype>
void ScaleLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  DC_CHECK(dc_scale_forward_nchw(reinterpret_cast<const float*>(bottom[0]->gpu_data()), reinterpret_cast<const float*>(this->blobs_[0]->gpu_data()),
                                 bias_term_ ? reinterpret_cast<const float*>(this->blobs_[1]->gpu_data()) : nullptr, outer_dim_, scale_dim_,
                                 inner_dim_, TopPtr(top[0], bottom[0]), Caffe::stream()));
}

// ===================================================================== ReLU / Sigmoid
template <typename Dtype>
void ReLULayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  DC_CHECK(dc_relu_forward(reinterpret_cast<const float*>(bottom[0]->gpu_data()), bottom[0]->count(), this->layer_param_.relu_param().negative_slope(),
                           TopPtr(top[0], bottom[0]), Caffe::stream()));
}
template <typename Dtype>
void SigmoidLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  DC_CHECK(dc_sigmoid_forward(reinterpret_cast<const float*>(bottom[0]->gpu_data()), bottom[0]->count(),
                              TopPtr(top[0], bottom[0]), Caffe::stream()));
}

// ===================================================================== Eltwise
template <typename Dtype>
void EltwiseLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const EltwiseParameter& p = this->layer_param_.eltwise_param();
  CHECK(p.coeff_size() == 0 || p.coeff_size() == (int)bottom.size()) << "Eltwise Layer takes one coefficient per bottom blob.";
  CHECK(!(p.operation() == EltwiseParameter_EltwiseOp_PROD && p.coeff_size())) << "Eltwise layer only takes coefficients for summation.";
  op_ = p.operation();
  CHECK(op_ == EltwiseParameter_EltwiseOp_SUM) << "Eltwise PROD/MAX are outside the DeeperCut path (the deploy net only sums)";
  coeffs_ = vector<Dtype>(bottom.size(), 1);
  for (int i = 0; i < p.coeff_size(); ++i) coeffs_[i] = p.coeff(i);
}
template <typename Dtype>
void EltwiseLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  for (size_t i = 1; i < bottom.size(); ++i) CHECK(bottom[i]->shape() == bottom[0]->shape());
  top[0]->ReshapeLike(*bottom[0]);
}
template <typename Dtype>
void EltwiseLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  float* y = (top[0] == bottom[0] || top[0] == bottom[1]) ? top[0]->mutable_gpu_data() : top[0]->overwrite_gpu_data();
  DC_CHECK(dc_axpby_forward(reinterpret_cast<const float*>(bottom[0]->gpu_data()), coeffs_[0], reinterpret_cast<const float*>(bottom[1]->gpu_data()),
                            coeffs_[1], top[0]->count(), y, Caffe::stream()));
  for (size_t i = 2; i < bottom.size(); ++i)
    DC_CHECK(dc_axpby_forward(y, 1.f, reinterpret_cast<const float*>(bottom[i]->gpu_data()), coeffs_[i], top[0]->count(), y, Caffe::stream()));
}

// ===================================================================== Pooling
template <typename Dtype>
void PoolingLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const PoolingParameter& p = this->layer_param_.pooling_param();
  if (p.global_pooling()) {
    CHECK(!(p.has_kernel_size() || p.has_kernel_h() || p.has_kernel_w())) << "With Global_pooling: true Filter size cannot specified";
  } else {
    CHECK(!p.has_kernel_size() != !(p.has_kernel_h() && p.has_kernel_w())) << "Filter size is kernel_size OR kernel_h and kernel_w; not both";
    CHECK(p.has_kernel_size() || (p.has_kernel_h() && p.has_kernel_w())) << "For non-square filters both kernel_h and kernel_w are required.";
  }
  CHECK((!p.has_pad() && p.has_pad_h() && p.has_pad_w()) || (!p.has_pad_h() && !p.has_pad_w())) << "pad is pad OR pad_h and pad_w are required.";
  CHECK((!p.has_stride() && p.has_stride_h() && p.has_stride_w()) || (!p.has_stride_h() && !p.has_stride_w())) << "Stride is stride OR stride_h and stride_w are required.";
  global_pooling_ = p.global_pooling();
  if (global_pooling_) { kernel_h_ = bottom[0]->height(); kernel_w_ = bottom[0]->width(); }
  else if (p.has_kernel_size()) kernel_h_ = kernel_w_ = p.kernel_size();
  else { kernel_h_ = p.kernel_h(); kernel_w_ = p.kernel_w(); }
  CHECK_GT(kernel_h_, 0) << "Filter dimensions cannot be zero.";
  CHECK_GT(kernel_w_, 0) << "Filter dimensions cannot be zero.";
  if (!p.has_pad_h()) pad_h_ = pad_w_ = p.pad(); else { pad_h_ = p.pad_h(); pad_w_ = p.pad_w(); }
  if (!p.has_stride_h()) stride_h_ = stride_w_ = p.stride(); else { stride_h_ = p.stride_h(); stride_w_ = p.stride_w(); }
  if (global_pooling_) CHECK(pad_h_ == 0 && pad_w_ == 0 && stride_h_ == 1 && stride_w_ == 1) << "With Global_pooling: true; only pad = 0 and stride = 1";
  CHECK(p.pool() == PoolingParameter_PoolMethod_MAX) << "AVE/STOCHASTIC pooling are outside the DeeperCut path (pool1 is MAX)";
  if (pad_h_ != 0 || pad_w_ != 0) { CHECK_LT(pad_h_, kernel_h_); CHECK_LT(pad_w_, kernel_w_); }
}
template <typename Dtype>
void PoolingLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(4, bottom[0]->num_axes()) << "Input must have 4 axes, corresponding to (num, channels, height, width)";
  channels_ = bottom[0]->channels();
  height_ = bottom[0]->height();
  width_ = bottom[0]->width();
  if (global_pooling_) { kernel_h_ = height_; kernel_w_ = width_; }
  pooled_height_ = static_cast<int>(std::ceil(static_cast<float>(height_ + 2 * pad_h_ - kernel_h_) / stride_h_)) + 1;
  pooled_width_ = static_cast<int>(std::ceil(static_cast<float>(width_ + 2 * pad_w_ - kernel_w_) / stride_w_)) + 1;
  if (pad_h_ || pad_w_) {
    if ((pooled_height_ - 1) * stride_h_ >= height_ + pad_h_) --pooled_height_;
    if ((pooled_width_ - 1) * stride_w_ >= width_ + pad_w_) --pooled_width_;
    CHECK_LT((pooled_height_ - 1) * stride_h_, height_ + pad_h_);
    CHECK_LT((pooled_width_ - 1) * stride_w_, width_ + pad_w_);
  }
  top[0]->Reshape(bottom[0]->num(), channels_, pooled_height_, pooled_width_);
}
template <typename Dtype>
void PoolingLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  DC_CHECK(dc_maxpool_forward_nchw(reinterpret_cast<const float*>(bottom[0]->gpu_data()), bottom[0]->num(), channels_, height_, width_, kernel_h_, kernel_w_,
                                   stride_h_, stride_w_, pad_h_, pad_w_, pooled_height_, pooled_width_,
                                   TopPtr(top[0], bottom[0]), Caffe::stream()));
}

// ===================================================================== Crop (DeepCut's)
template <typename Dtype>
void CropLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const CropParameter& param = this->layer_param_.crop_param();
  CHECK_EQ(bottom.size(), 2u) << "Wrong number of bottom blobs.";
  CHECK_EQ(bottom[0]->num_axes(), 4) << "Only works with 4D blobs.";
  CHECK_EQ(bottom[1]->num_axes(), 4) << "Only works with 4D blobs.";
  crop_h_ = param.offset_height();
  crop_w_ = param.offset_width();
}
template <typename Dtype>
void CropLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_GT(bottom[0]->height() - crop_h_, bottom[1]->height()) << "invalid offset";
  CHECK_GT(bottom[0]->width() - crop_w_, bottom[1]->width()) << "invalid offset";
  top[0]->Reshape(bottom[0]->num(), bottom[0]->channels(), bottom[1]->height(), bottom[1]->width());
}
template <typename Dtype>
void CropLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  DC_CHECK(dc_crop_forward_nchw(reinterpret_cast<const float*>(bottom[0]->gpu_data()), bottom[0]->num(), bottom[0]->channels(), bottom[0]->height(),
                                bottom[0]->width(), crop_h_, crop_w_, top[0]->height(), top[0]->width(),
                                TopPtr(top[0], bottom[0]), Caffe::stream()));
}

// ===================================================================== Split
template <typename Dtype>
void SplitLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  for (size_t i = 0; i < top.size(); ++i) {
    CHECK_NE(top[i], bottom[0]) << this->type() << " Layer does not allow in-place computation.";
    top[i]->ReshapeLike(*bottom[0]);
    CHECK_EQ(bottom[0]->count(), top[i]->count());
  }
}
template <typename Dtype>
void SplitLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  for (size_t i = 0; i < top.size(); ++i) top[i]->ShareData(*bottom[0]);   // zero-copy alias, any mode
}

INSTANTIATE_CLASS(BaseConvolutionLayer);
INSTANTIATE_CLASS(ConvolutionLayer);
INSTANTIATE_CLASS(DeconvolutionLayer);
INSTANTIATE_CLASS(BatchNormLayer);
INSTANTIATE_CLASS(ScaleLayer);
INSTANTIATE_CLASS(ReLULayer);
INSTANTIATE_CLASS(SigmoidLayer);
INSTANTIATE_CLASS(EltwiseLayer);
INSTANTIATE_CLASS(PoolingLayer);
INSTANTIATE_CLASS(CropLayer);
INSTANTIATE_CLASS(SplitLayer);
REGISTER_LAYER_CLASS(Convolution);
REGISTER_LAYER_CLASS(Deconvolution);
REGISTER_LAYER_CLASS(BatchNorm);
REGISTER_LAYER_CLASS(Scale);
REGISTER_LAYER_CLASS(ReLU);
REGISTER_LAYER_CLASS(Sigmoid);
REGISTER_LAYER_CLASS(Eltwise);
REGISTER_LAYER_CLASS(Pooling);
REGISTER_LAYER_CLASS(Crop);
REGISTER_LAYER_CLASS(Split);

}  // namespace caffe
