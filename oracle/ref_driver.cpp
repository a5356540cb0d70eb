// oracle/_ref driver (TEST INFRASTRUCTURE): runs the reference's OWN CPU layer code -- compiled verbatim from
// /root/reference by oracle/build_ref.py -- over a deploy prototxt, behind a tiny C API for ctypes.
// What is the reference's: Blob, SyncedMemory, Layer, LayerRegistry, InsertSplits, every layer's
// LayerSetUp / Reshape / Forward_cpu, im2col / col2im, the math wrappers, the Caffe singleton.
// What is restated here (the reference's net.cpp needs HDF5 + the solver headers): the Net::Init wiring loop
// (net.cpp:40-284: blob table by name, in-place tops, AppendBottom/AppendTop) and the four engine-dispatching
// creators of layer_factory.cpp:37-193 (CAFFE engine branch).  The prototxt parser is the protobuf-free
// look-alike (caffe/proto/caffe.pb.h shim).
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/layer.hpp"
#include "caffe/layer_factory.hpp"
#include "caffe/layers/conv_layer.hpp"
#include "caffe/layers/pooling_layer.hpp"
#include "caffe/layers/relu_layer.hpp"
#include "caffe/layers/sigmoid_layer.hpp"
#include "caffe/util/insert_splits.hpp"

namespace caffe {
// layer_factory.cpp:37-73,76-110,150-193 with Engine DEFAULT -> CAFFE (no cuDNN in a CPU_ONLY build)
template <typename Dtype> shared_ptr<Layer<Dtype> > GetConvolutionLayer(const LayerParameter& p) { return shared_ptr<Layer<Dtype> >(new ConvolutionLayer<Dtype>(p)); }
template <typename Dtype> shared_ptr<Layer<Dtype> > GetPoolingLayer(const LayerParameter& p) { return shared_ptr<Layer<Dtype> >(new PoolingLayer<Dtype>(p)); }
template <typename Dtype> shared_ptr<Layer<Dtype> > GetReLULayer(const LayerParameter& p) { return shared_ptr<Layer<Dtype> >(new ReLULayer<Dtype>(p)); }
template <typename Dtype> shared_ptr<Layer<Dtype> > GetSigmoidLayer(const LayerParameter& p) { return shared_ptr<Layer<Dtype> >(new SigmoidLayer<Dtype>(p)); }
REGISTER_LAYER_CREATOR(Convolution, GetConvolutionLayer);
REGISTER_LAYER_CREATOR(Pooling, GetPoolingLayer);
REGISTER_LAYER_CREATOR(ReLU, GetReLULayer);
REGISTER_LAYER_CREATOR(Sigmoid, GetSigmoidLayer);
}  // namespace caffe

using namespace caffe;  // NOLINT

namespace {
thread_local std::string g_err;

struct RefNet {
  NetParameter param;                                    // after InsertSplits
  std::vector<shared_ptr<Layer<float> > > layers;
  std::vector<std::string> layer_names;
  std::vector<shared_ptr<Blob<float> > > blobs;
  std::vector<std::string> blob_names;
  std::map<std::string, int> blob_index;
  std::vector<std::vector<Blob<float>*> > bottoms, tops;

  void Init(const std::string& text) {
    NetParameter in;
    std::string err;
    CHECK(in.ParseFromTextString(text, &err)) << err;
    InsertSplits(in, &param);
    for (int i = 0; i < param.input_size(); ++i) {
      shared_ptr<Blob<float> > b(new Blob<float>());
      b->Reshape(param.input_dim(4 * i), param.input_dim(4 * i + 1), param.input_dim(4 * i + 2), param.input_dim(4 * i + 3));
      blob_index[param.input(i)] = static_cast<int>(blobs.size());
      blobs.push_back(b);
      blob_names.push_back(param.input(i));
    }
    const int n = param.layer_size();
    bottoms.resize(n);
    tops.resize(n);
    for (int i = 0; i < n; ++i) {
      LayerParameter lp(param.layer(i));
      if (!lp.has_phase()) lp.set_phase(TEST);
      layers.push_back(LayerRegistry<float>::CreateLayer(lp));
      layer_names.push_back(lp.name());
      for (int j = 0; j < lp.bottom_size(); ++j) {
        CHECK(blob_index.count(lp.bottom(j))) << "Unknown bottom blob '" << lp.bottom(j) << "'";
        bottoms[i].push_back(blobs[blob_index[lp.bottom(j)]].get());
      }
      for (int j = 0; j < lp.top_size(); ++j) {
        if (j < lp.bottom_size() && lp.top(j) == lp.bottom(j)) {          // in-place (net.cpp:393-400)
          tops[i].push_back(blobs[blob_index[lp.top(j)]].get());
        } else {
          CHECK(!blob_index.count(lp.top(j))) << "Top blob '" << lp.top(j) << "' produced by multiple sources.";
          shared_ptr<Blob<float> > b(new Blob<float>());
          blob_index[lp.top(j)] = static_cast<int>(blobs.size());
          blobs.push_back(b);
          blob_names.push_back(lp.top(j));
          tops[i].push_back(b.get());
        }
      }
      layers[i]->SetUp(bottoms[i], tops[i]);
    }
  }
  void Forward() {
    for (size_t i = 0; i < layers.size(); ++i) layers[i]->Forward(bottoms[i], tops[i]);      // net.cpp:574-579
  }
};

template <class F>
int Guard(F f) {
  try { f(); return 0; } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
}  // namespace

extern "C" {
const char* refcaffe_last_error(void) { return g_err.c_str(); }
void* refcaffe_net_create(const char* prototxt_text) {
  RefNet* net = new RefNet();
  Caffe::set_mode(Caffe::CPU);
  if (Guard([&] { net->Init(prototxt_text); })) { delete net; return nullptr; }
  return net;
}
void refcaffe_net_destroy(void* h) { delete static_cast<RefNet*>(h); }
int refcaffe_net_forward(void* h) { return Guard([&] { static_cast<RefNet*>(h)->Forward(); }); }
int refcaffe_num_layers(void* h) { return static_cast<int>(static_cast<RefNet*>(h)->layers.size()); }
const char* refcaffe_layer_name(void* h, int i) { return static_cast<RefNet*>(h)->layer_names[i].c_str(); }
const char* refcaffe_layer_type(void* h, int i) { return static_cast<RefNet*>(h)->layers[i]->type(); }
int refcaffe_layer_num_blobs(void* h, int i) { return static_cast<int>(static_cast<RefNet*>(h)->layers[i]->blobs().size()); }
int refcaffe_layer_blob_count(void* h, int i, int j) { return static_cast<RefNet*>(h)->layers[i]->blobs()[j]->count(); }
float* refcaffe_layer_blob_data(void* h, int i, int j) { return static_cast<RefNet*>(h)->layers[i]->blobs()[j]->mutable_cpu_data(); }
int refcaffe_num_blobs(void* h) { return static_cast<int>(static_cast<RefNet*>(h)->blobs.size()); }
const char* refcaffe_blob_name(void* h, int i) { return static_cast<RefNet*>(h)->blob_names[i].c_str(); }
int refcaffe_blob_shape(void* h, int i, int* dims) {
  const Blob<float>& b = *static_cast<RefNet*>(h)->blobs[i];
  for (int a = 0; a < b.num_axes(); ++a) dims[a] = b.shape(a);
  return b.num_axes();
}
int refcaffe_blob_reshape(void* h, int i, int n, int c, int hh, int w) {
  return Guard([&] { static_cast<RefNet*>(h)->blobs[i]->Reshape(n, c, hh, w); });
}
float* refcaffe_blob_data(void* h, int i) { return static_cast<RefNet*>(h)->blobs[i]->mutable_cpu_data(); }
}  // extern "C"
