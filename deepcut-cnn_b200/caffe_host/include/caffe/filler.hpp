// Fillers (reference include/caffe/filler.hpp:18-292): constant / uniform / gaussian / xavier / msra.
// The deploy prototxt names none, so blobs default to `constant` 0 (caffe.proto:43-46).
#pragma once
#include <random>

#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/proto/caffe.pb.h"

namespace caffe {

template <typename Dtype>
class Filler {
 public:
  explicit Filler(const FillerParameter& param) : filler_param_(param) {}
  virtual ~Filler() {}
  virtual void Fill(Blob<Dtype>* blob) = 0;

 protected:
  FillerParameter filler_param_;
};

template <typename Dtype>
Filler<Dtype>* GetFiller(const FillerParameter& param);

}  // namespace caffe
