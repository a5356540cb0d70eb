// The protoc-generated header the reference includes everywhere is replaced by the protobuf-free look-alike
// of the product tree (same accessor names over the same caffe.proto subset).
#pragma once
#include "../../../../deepcut-cnn_b200/caffe_host/include/caffe/proto/caffe.pb.h"
