#include "caffe/filler.hpp"

namespace caffe {

namespace {
std::mt19937& Rng() {
  static thread_local std::mt19937 gen(Caffe::random_seed());
  return gen;
}

template <typename Dtype>
Dtype FanNorm(const Blob<Dtype>& blob, FillerParameter_VarianceNorm mode) {
  const int fan_in = blob.count() / blob.num();
  const int fan_out = blob.count() / blob.channels();
  if (mode == FillerParameter_VarianceNorm_AVERAGE) return (fan_in + fan_out) / Dtype(2);
  if (mode == FillerParameter_VarianceNorm_FAN_OUT) return static_cast<Dtype>(fan_out);
  return static_cast<Dtype>(fan_in);
}

template <typename Dtype>
class ConstantFiller : public Filler<Dtype> {
 public:
  explicit ConstantFiller(const FillerParameter& p) : Filler<Dtype>(p) {}
  void Fill(Blob<Dtype>* blob) override {
    Dtype* d = blob->mutable_cpu_data();
    const Dtype v = this->filler_param_.value();
    CHECK(blob->count());
    for (int i = 0; i < blob->count(); ++i) d[i] = v;
  }
};
template <typename Dtype>
class UniformFiller : public Filler<Dtype> {
 public:
  explicit UniformFiller(const FillerParameter& p) : Filler<Dtype>(p) {}
  void Fill(Blob<Dtype>* blob) override {
    std::uniform_real_distribution<Dtype> dist(this->filler_param_.min(), this->filler_param_.max());
    Dtype* d = blob->mutable_cpu_data();
    for (int i = 0; i < blob->count(); ++i) d[i] = dist(Rng());
  }
};
template <typename Dtype>
class GaussianFiller : public Filler<Dtype> {
 public:
  explicit GaussianFiller(const FillerParameter& p) : Filler<Dtype>(p) {}
  void Fill(Blob<Dtype>* blob) override {
    std::normal_distribution<Dtype> dist(this->filler_param_.mean(), this->filler_param_.std());
    Dtype* d = blob->mutable_cpu_data();
    for (int i = 0; i < blob->count(); ++i) d[i] = dist(Rng());
    CHECK_EQ(this->filler_param_.sparse(), -1) << "sparse gaussian filling is not supported";
  }
};
template <typename Dtype>
class XavierFiller : public Filler<Dtype> {
 public:
  explicit XavierFiller(const FillerParameter& p) : Filler<Dtype>(p) {}
  void Fill(Blob<Dtype>* blob) override {
    const Dtype scale = std::sqrt(Dtype(3) / FanNorm(*blob, this->filler_param_.variance_norm()));
    std::uniform_real_distribution<Dtype> dist(-scale, scale);
    Dtype* d = blob->mutable_cpu_data();
    for (int i = 0; i < blob->count(); ++i) d[i] = dist(Rng());
  }
};
template <typename Dtype>
class MSRAFiller : public Filler<Dtype> {
 public:
  explicit MSRAFiller(const FillerParameter& p) : Filler<Dtype>(p) {}
  void Fill(Blob<Dtype>* blob) override {
    const Dtype std = std::sqrt(Dtype(2) / FanNorm(*blob, this->filler_param_.variance_norm()));
    std::normal_distribution<Dtype> dist(Dtype(0), std);
    Dtype* d = blob->mutable_cpu_data();
    for (int i = 0; i < blob->count(); ++i) d[i] = dist(Rng());
  }
};
}  // namespace

template <typename Dtype>
Filler<Dtype>* GetFiller(const FillerParameter& param) {
  const string& type = param.type();
  if (type == "constant") return new ConstantFiller<Dtype>(param);
  if (type == "gaussian") return new GaussianFiller<Dtype>(param);
  if (type == "uniform") return new UniformFiller<Dtype>(param);
  if (type == "xavier") return new XavierFiller<Dtype>(param);
  if (type == "msra") return new MSRAFiller<Dtype>(param);
  LOG(FATAL) << "Unknown filler name: " << type;
  return nullptr;
}
template Filler<float>* GetFiller<float>(const FillerParameter& param);

}  // namespace caffe
