// Layer plugin registry: maps a prototxt `type:` string to the function that builds the layer.
// Same public surface as the reference's registry (include/caffe/layer_factory.hpp:52-137: LayerRegistry<Dtype>::
// {Creator, CreatorRegistry, Registry, AddCreator, CreateLayer, LayerTypeList}, LayerRegisterer, REGISTER_LAYER_CREATOR,
// REGISTER_LAYER_CLASS), so a layer written against the reference registers here unchanged.  Float only, like pycaffe
// (python/caffe/_caffe.cpp:34): the macros instantiate and register the float creator.
#pragma once
#include <map>

#include "caffe/common.hpp"
#include "caffe/layer.hpp"
#include "caffe/proto/caffe.pb.h"

namespace caffe {

template <typename Dtype>
class LayerRegistry {
 public:
  using LayerPtr = shared_ptr<Layer<Dtype> >;
  typedef LayerPtr (*Creator)(const LayerParameter&);
  typedef std::map<string, Creator> CreatorRegistry;

  // one table per Dtype, built on first use (registration runs during static initialisation of the layer files)
  static CreatorRegistry& Registry() {
    static CreatorRegistry table;
    return table;
  }

  static void AddCreator(const string& type, Creator creator) {
    const bool inserted = Registry().emplace(type, creator).second;
    CHECK(inserted) << "Layer type " << type << " already registered.";
  }

  static LayerPtr CreateLayer(const LayerParameter& param) {
    const auto hit = Registry().find(param.type());
    CHECK(hit != Registry().end()) << "Unknown layer type: " << param.type() << " (known types: " << Known() << ")";
    return hit->second(param);
  }

  static vector<string> LayerTypeList() {
    vector<string> names;
    names.reserve(Registry().size());
    for (const auto& entry : Registry()) names.push_back(entry.first);
    return names;
  }

 private:
  LayerRegistry() = delete;
  static string Known() {
    string joined;
    for (const auto& entry : Registry()) {
      if (!joined.empty()) joined += ", ";
      joined += entry.first;
    }
    return joined;
  }
};

// A static object of this type in a layer's translation unit performs the registration.
template <typename Dtype>
struct LayerRegisterer {
  LayerRegisterer(const string& type, typename LayerRegistry<Dtype>::Creator creator) { LayerRegistry<Dtype>::AddCreator(type, creator); }
};

namespace detail {
template <template <typename> class LayerT, typename Dtype>
shared_ptr<Layer<Dtype> > MakeLayer(const LayerParameter& param) {
  return shared_ptr<Layer<Dtype> >(new LayerT<Dtype>(param));
}
}  // namespace detail

#define REGISTER_LAYER_CREATOR(type, creator) static ::caffe::LayerRegisterer<float> g_creator_f_##type(#type, creator<float>)

#define REGISTER_LAYER_CLASS(type)                                                                                            \
  template <typename Dtype>                                                                                                   \
  ::caffe::shared_ptr<::caffe::Layer<Dtype> > Creator_##type##Layer(const ::caffe::LayerParameter& param) {                    \
    return ::caffe::detail::MakeLayer<type##Layer, Dtype>(param);                                                              \
  }                                                                                                                           \
  REGISTER_LAYER_CREATOR(type, Creator_##type##Layer)

}  // namespace caffe
