#include "caffe/blob.hpp"

#include <climits>
#include <cstring>

#include "deepcut_b200.h"

namespace caffe {

template <typename Dtype>
void Blob<Dtype>::Reshape(const int num, const int channels, const int height, const int width) {
  vector<int> shape(4);
  shape[0] = num; shape[1] = channels; shape[2] = height; shape[3] = width;
  Reshape(shape);
}

template <typename Dtype>
void Blob<Dtype>::Reshape(const vector<int>& shape) {
  CHECK_LE((int)shape.size(), kMaxBlobAxes);
  count_ = 1;
  shape_.resize(shape.size());
  for (size_t i = 0; i < shape.size(); ++i) {
    CHECK_GE(shape[i], 0);
    if (count_ != 0) CHECK_LE(shape[i], INT_MAX / count_) << "blob size exceeds INT_MAX";
    count_ *= shape[i];
    shape_[i] = shape[i];
  }
  if (count_ > capacity_) {     // grows only, never shrinks (reference blob.cpp:38-42)
    capacity_ = count_;
    data_.reset(new SyncedMemory(capacity_ * sizeof(Dtype)));
    diff_.reset(new SyncedMemory(capacity_ * sizeof(Dtype)));
  }
}

template <typename Dtype>
void Blob<Dtype>::Reshape(const BlobShape& shape) {
  CHECK_LE(shape.dim_size(), kMaxBlobAxes);
  vector<int> shape_vec(shape.dim_size());
  for (int i = 0; i < shape.dim_size(); ++i) shape_vec[i] = static_cast<int>(shape.dim(i));
  Reshape(shape_vec);
}

template <typename Dtype>
string Blob<Dtype>::shape_string() const {
  ostringstream stream;
  for (size_t i = 0; i < shape_.size(); ++i) stream << shape_[i] << " ";
  stream << "(" << count_ << ")";
  return stream.str();
}

template <typename Dtype>
int Blob<Dtype>::count(int start_axis, int end_axis) const {
  CHECK_LE(start_axis, end_axis);
  CHECK_GE(start_axis, 0);
  CHECK_LE(end_axis, num_axes());
  int count = 1;
  for (int i = start_axis; i < end_axis; ++i) count *= shape_[i];
  return count;
}

template <typename Dtype>
int Blob<Dtype>::CanonicalAxisIndex(int axis_index) const {
  CHECK_GE(axis_index, -num_axes()) << "axis " << axis_index << " out of range for " << num_axes() << "-D Blob with shape " << shape_string();
  CHECK_LT(axis_index, num_axes()) << "axis " << axis_index << " out of range for " << num_axes() << "-D Blob with shape " << shape_string();
  return axis_index < 0 ? axis_index + num_axes() : axis_index;
}

template <typename Dtype>
int Blob<Dtype>::LegacyShape(int index) const {
  CHECK_LE(num_axes(), 4) << "Cannot use legacy accessors on Blobs with > 4 axes.";
  CHECK_LT(index, 4);
  CHECK_GE(index, -4);
  if (index >= num_axes() || index < -num_axes()) return 1;   // missing trailing axes read as 1
  return shape(index);
}

template <typename Dtype>
int Blob<Dtype>::offset(const int n, const int c, const int h, const int w) const {
  CHECK_GE(n, 0); CHECK_LE(n, num());
  CHECK_GE(channels(), 0); CHECK_LE(c, channels());
  CHECK_GE(height(), 0); CHECK_LE(h, height());
  CHECK_GE(width(), 0); CHECK_LE(w, width());
  return ((n * channels() + c) * height() + h) * width() + w;
}

template <typename Dtype>
int Blob<Dtype>::offset(const vector<int>& indices) const {
  CHECK_LE((int)indices.size(), num_axes());
  int offset = 0;
  for (int i = 0; i < num_axes(); ++i) {
    offset *= shape(i);
    if ((int)indices.size() > i) {
      CHECK_GE(indices[i], 0);
      CHECK_LT(indices[i], shape(i));
      offset += indices[i];
    }
  }
  return offset;
}

template <typename Dtype> const Dtype* Blob<Dtype>::cpu_data() const { CHECK(data_); return (const Dtype*)data_->cpu_data(); }
template <typename Dtype> void Blob<Dtype>::set_cpu_data(Dtype* data) { CHECK(data); data_->set_cpu_data(data); }
template <typename Dtype> const Dtype* Blob<Dtype>::gpu_data() const { CHECK(data_); return (const Dtype*)data_->gpu_data(); }
template <typename Dtype> const Dtype* Blob<Dtype>::cpu_diff() const { CHECK(diff_); return (const Dtype*)diff_->cpu_data(); }
template <typename Dtype> const Dtype* Blob<Dtype>::gpu_diff() const { CHECK(diff_); return (const Dtype*)diff_->gpu_data(); }
template <typename Dtype> Dtype* Blob<Dtype>::mutable_cpu_data() { CHECK(data_); return static_cast<Dtype*>(data_->mutable_cpu_data()); }
template <typename Dtype> Dtype* Blob<Dtype>::mutable_gpu_data() { CHECK(data_); return static_cast<Dtype*>(data_->mutable_gpu_data()); }
template <typename Dtype> Dtype* Blob<Dtype>::mutable_cpu_diff() { CHECK(diff_); return static_cast<Dtype*>(diff_->mutable_cpu_data()); }
template <typename Dtype> Dtype* Blob<Dtype>::mutable_gpu_diff() { CHECK(diff_); return static_cast<Dtype*>(diff_->mutable_gpu_data()); }

template <typename Dtype>
void Blob<Dtype>::ShareData(const Blob& other) { CHECK_EQ(count_, other.count()); data_ = other.data(); }
template <typename Dtype>
void Blob<Dtype>::ShareDiff(const Blob& other) { CHECK_EQ(count_, other.count()); diff_ = other.diff(); }

template <typename Dtype>
Dtype Blob<Dtype>::asum_data() const {
  if (!data_) return 0;
  const Dtype* d = cpu_data();
  double s = 0;
  for (int i = 0; i < count_; ++i) s += std::fabs(d[i]);
  return static_cast<Dtype>(s);
}
template <typename Dtype>
Dtype Blob<Dtype>::sumsq_data() const {
  if (!data_) return 0;
  const Dtype* d = cpu_data();
  double s = 0;
  for (int i = 0; i < count_; ++i) s += static_cast<double>(d[i]) * d[i];
  return static_cast<Dtype>(s);
}
template <typename Dtype>
void Blob<Dtype>::scale_data(Dtype scale_factor) {
  if (!data_) return;
  Dtype* d = mutable_cpu_data();
  for (int i = 0; i < count_; ++i) d[i] *= scale_factor;
}

template <typename Dtype>
bool Blob<Dtype>::ShapeEquals(const BlobProto& other) {
  if (other.has_num() || other.has_channels() || other.has_height() || other.has_width()) {
    // legacy 4-D proto: missing leading axes of this blob count as 1 (reference blob.cpp:395-409)
    return shape_.size() <= 4 && LegacyShape(-4) == other.num() && LegacyShape(-3) == other.channels() &&
           LegacyShape(-2) == other.height() && LegacyShape(-1) == other.width();
  }
  vector<int> other_shape(other.shape().dim_size());
  for (int i = 0; i < other.shape().dim_size(); ++i) other_shape[i] = static_cast<int>(other.shape().dim(i));
  return shape_ == other_shape;
}

template <typename Dtype>
void Blob<Dtype>::CopyFrom(const Blob& source, bool copy_diff, bool reshape) {
  if (source.count() != count_ || source.shape() != shape_) {
    if (reshape) ReshapeLike(source);
    else LOG(FATAL) << "Trying to copy blobs of different sizes.";
  }
  if (Caffe::mode() == Caffe::GPU) {
    const void* src = copy_diff ? (const void*)source.gpu_diff() : (const void*)source.gpu_data();
    void* dst = copy_diff ? (void*)mutable_gpu_diff() : (void*)mutable_gpu_data();
    DC_CHECK(dc_memcpy_async(dst, src, sizeof(Dtype) * count_, DC_D2D, Caffe::stream()));
  } else {
    const void* src = copy_diff ? (const void*)source.cpu_diff() : (const void*)source.cpu_data();
    void* dst = copy_diff ? (void*)mutable_cpu_diff() : (void*)mutable_cpu_data();
    memcpy(dst, src, sizeof(Dtype) * count_);
  }
}

template <typename Dtype>
void Blob<Dtype>::FromProto(const BlobProto& proto, bool reshape) {
  if (reshape) {
    vector<int> shape;
    if (proto.has_num() || proto.has_channels() || proto.has_height() || proto.has_width()) {
      shape.resize(4);
      shape[0] = proto.num(); shape[1] = proto.channels(); shape[2] = proto.height(); shape[3] = proto.width();
    } else {
      shape.resize(proto.shape().dim_size());
      for (int i = 0; i < proto.shape().dim_size(); ++i) shape[i] = static_cast<int>(proto.shape().dim(i));
    }
    Reshape(shape);
  } else {
    CHECK(ShapeEquals(proto)) << "shape mismatch (reshape not set)";
  }
  Dtype* data_vec = mutable_cpu_data();
  if (proto.double_data_size() > 0) {
    CHECK_EQ(count_, proto.double_data_size());
    for (int i = 0; i < count_; ++i) data_vec[i] = static_cast<Dtype>(proto.double_data(i));
  } else {
    CHECK_EQ(count_, proto.data_size());
    for (int i = 0; i < count_; ++i) data_vec[i] = static_cast<Dtype>(proto.data(i));
  }
  if (proto.double_diff_size() > 0) {
    CHECK_EQ(count_, proto.double_diff_size());
    Dtype* diff_vec = mutable_cpu_diff();
    for (int i = 0; i < count_; ++i) diff_vec[i] = static_cast<Dtype>(proto.double_diff(i));
  } else if (proto.diff_size() > 0) {
    CHECK_EQ(count_, proto.diff_size());
    Dtype* diff_vec = mutable_cpu_diff();
    for (int i = 0; i < count_; ++i) diff_vec[i] = static_cast<Dtype>(proto.diff(i));
  }
}

template <typename Dtype>
void Blob<Dtype>::ToProto(BlobProto* proto, bool write_diff) const {
  proto->clear_shape();
  for (size_t i = 0; i < shape_.size(); ++i) proto->mutable_shape()->add_dim(shape_[i]);
  proto->clear_data();
  proto->clear_diff();
  const Dtype* data_vec = cpu_data();
  proto->mutable_data()->assign(data_vec, data_vec + count_);
  if (write_diff) {
    const Dtype* diff_vec = cpu_diff();
    proto->mutable_diff()->assign(diff_vec, diff_vec + count_);
  }
}

INSTANTIATE_CLASS(Blob);
template class Blob<int>;
template class Blob<unsigned int>;

}  // namespace caffe
