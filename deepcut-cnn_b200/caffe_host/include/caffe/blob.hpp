// Blob<Dtype>: N-d tensor over two SyncedMemory (data, diff); public API of
// include/caffe/blob.hpp:24-277 (Reshape, shape/num/channels/height/width, count, offset,
// cpu/gpu accessors, ShareData, FromProto/ToProto, data_at, asum/sumsq).  fp32 NCHW row-major.
#pragma once
#include <algorithm>

#include "caffe/common.hpp"
#include "caffe/proto/caffe.pb.h"
#include "caffe/syncedmem.hpp"

const int kMaxBlobAxes = 32;

namespace caffe {

template <typename Dtype>
class Blob {
 public:
  Blob() {}
  explicit Blob(const int num, const int channels, const int height, const int width) { Reshape(num, channels, height, width); }
  explicit Blob(const vector<int>& shape) { Reshape(shape); }

  void Reshape(const int num, const int channels, const int height, const int width);
  void Reshape(const vector<int>& shape);
  void Reshape(const BlobShape& shape);
  void ReshapeLike(const Blob& other) { Reshape(other.shape()); }
  string shape_string() const;
  const vector<int>& shape() const { return shape_; }
  int shape(int index) const { return shape_[CanonicalAxisIndex(index)]; }
  int num_axes() const { return static_cast<int>(shape_.size()); }
  int count() const { return count_; }
  int count(int start_axis, int end_axis) const;
  int count(int start_axis) const { return count(start_axis, num_axes()); }
  int CanonicalAxisIndex(int axis_index) const;
  int num() const { return LegacyShape(0); }
  int channels() const { return LegacyShape(1); }
  int height() const { return LegacyShape(2); }
  int width() const { return LegacyShape(3); }
  int LegacyShape(int index) const;
  int offset(const int n, const int c = 0, const int h = 0, const int w = 0) const;
  int offset(const vector<int>& indices) const;
  void CopyFrom(const Blob<Dtype>& source, bool copy_diff = false, bool reshape = false);
  Dtype data_at(const int n, const int c, const int h, const int w) const { return cpu_data()[offset(n, c, h, w)]; }
  Dtype data_at(const vector<int>& index) const { return cpu_data()[offset(index)]; }
  const shared_ptr<SyncedMemory>& data() const { return data_; }
  const shared_ptr<SyncedMemory>& diff() const { return diff_; }

  const Dtype* cpu_data() const;
  void set_cpu_data(Dtype* data);
  const Dtype* gpu_data() const;
  const Dtype* cpu_diff() const;
  const Dtype* gpu_diff() const;
  Dtype* mutable_cpu_data();
  Dtype* mutable_gpu_data();
  Dtype* overwrite_gpu_data() { CHECK(data_); return static_cast<Dtype*>(data_->overwrite_gpu_data()); }
  Dtype* mutable_cpu_diff();
  Dtype* mutable_gpu_diff();
  void FromProto(const BlobProto& proto, bool reshape = true);
  void ToProto(BlobProto* proto, bool write_diff = false) const;
  Dtype asum_data() const;
  Dtype sumsq_data() const;
  void scale_data(Dtype scale_factor);
  void ShareData(const Blob& other);
  void ShareDiff(const Blob& other);
  bool ShapeEquals(const BlobProto& other);

 protected:
  shared_ptr<SyncedMemory> data_;
  shared_ptr<SyncedMemory> diff_;
  vector<int> shape_;
  int count_ = 0;
  int capacity_ = 0;
  DISABLE_COPY_AND_ASSIGN(Blob);
};

}  // namespace caffe
