// A user of the reference's C++ plugin API, compiled against THIS repo's headers and linked against libcaffe_b200.so: it
// defines a layer type the library does not have (ScaledSum: top = 2 * bottom0 - 0.5 * bottom1), publishes it with
// REGISTER_LAYER_CLASS exactly as a reference layer file would (include/caffe/layer_factory.hpp:127-137), and runs it in a
// Net next to built-in layers through the public Net / Blob / Caffe API (include/caffe/net.hpp, blob.hpp, common.hpp).
//
//   plugin_demo init  <prototxt>      no GPU needed: registry, Net::Init, shapes, blob-count checks
//   plugin_demo gpu   <prototxt>      Caffe::GPU: Net::Forward, result checked on the host
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "caffe/caffe.hpp"
#include "caffe/layer_factory.hpp"
#include "deepcut_b200.h"

namespace caffe {

template <typename Dtype>
class ScaledSumLayer : public Layer<Dtype> {
 public:
  explicit ScaledSumLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    CHECK(bottom[0]->shape() == bottom[1]->shape()) << "ScaledSum: bottoms must have one shape";
    top[0]->ReshapeLike(*bottom[0]);
  }
  virtual inline const char* type() const { return "ScaledSum"; }
  virtual inline int ExactNumBottomBlobs() const { return 2; }
  virtual inline int ExactNumTopBlobs() const { return 1; }

 protected:
  // the device path goes through the kernels' C ABI, like the built-in layers
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    CHECK_EQ(0, dc_axpby_forward(bottom[0]->gpu_data(), 2.f, bottom[1]->gpu_data(), -0.5f, bottom[0]->count(),
                                 top[0]->mutable_gpu_data(), Caffe::stream()))
        << dc_last_error();
  }
};
REGISTER_LAYER_CLASS(ScaledSum);

}  // namespace caffe

using namespace caffe;  // NOLINT

static int fail(const char* what) {
  std::printf("FAIL: %s\n", what);
  return 1;
}

int main(int argc, char** argv) {
  if (argc != 3) return fail("usage: plugin_demo init|gpu <prototxt>");
  const bool gpu = std::strcmp(argv[1], "gpu") == 0;
  // the registry knows the plugin and the built-in types
  bool have_plugin = false, have_conv = false;
  for (const std::string& t : LayerRegistry<float>::LayerTypeList()) {
    have_plugin |= t == "ScaledSum";
    have_conv |= t == "Convolution";
  }
  if (!have_plugin || !have_conv) return fail("registry is missing ScaledSum or Convolution");
  if (gpu) {
    Caffe::SetDevice(0);
    Caffe::set_mode(Caffe::GPU);
  } else {
    Caffe::set_mode(Caffe::CPU);
  }
  Net<float> net(argv[2], TEST);
  if (!net.has_blob("out") || net.num_outputs() != 1) return fail("net outputs");
  const shared_ptr<Blob<float> > data = net.blob_by_name("data");
  const shared_ptr<Blob<float> > out = net.blob_by_name("out");
  if (out->shape() != data->shape()) return fail("shape propagation through the plugin layer");
  if (std::string(net.layer_by_name("mix")->type()) != "ScaledSum") return fail("layer_by_name");
  if (!gpu) {
    std::printf("OK init: %d layers, out %s\n", static_cast<int>(net.layers().size()), out->shape_string().c_str());
    return 0;
  }
  float* x = data->mutable_cpu_data();
  for (int i = 0; i < data->count(); ++i) x[i] = 0.37f * static_cast<float>((i * 7) % 23 - 11);
  net.Forward();
  const float* y = out->cpu_data();
  double worst = 0;
  for (int i = 0; i < out->count(); ++i) {
    const double r = x[i] > 0 ? x[i] : 0, s = 1.0 / (1.0 + std::exp(-static_cast<double>(x[i])));
    worst = std::max(worst, std::fabs(2 * r - 0.5 * s - y[i]));
  }
  if (worst > 1e-5) return fail("forward result");
  // reshape-and-repeat through the public API (NetTest.TestReshape, src/caffe/test/test_net.cpp:2262-2332)
  data->Reshape(2, 3, 5, 9);
  net.Reshape();
  if (out->shape() != data->shape()) return fail("Net::Reshape");
  x = data->mutable_cpu_data();
  for (int i = 0; i < data->count(); ++i) x[i] = -1.f + 0.01f * i;
  net.Forward();
  y = out->cpu_data();
  for (int i = 0; i < out->count(); ++i) {
    const double r = x[i] > 0 ? x[i] : 0, s = 1.0 / (1.0 + std::exp(-static_cast<double>(x[i])));
    worst = std::max(worst, std::fabs(2 * r - 0.5 * s - y[i]));
  }
  if (worst > 1e-5) return fail("forward after reshape");
  std::printf("OK gpu: max |err| = %.3g\n", worst);
  return 0;
}
