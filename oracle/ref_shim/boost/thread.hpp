// boost::mutex / boost::thread_specific_ptr stand-ins (src/caffe/layer.cpp:1-24, src/caffe/common.cpp:13-20).
#pragma once
#include <memory>
#include <mutex>
#include <unordered_map>
namespace boost {
class mutex : public std::mutex {};
template <class T>
class thread_specific_ptr {
 public:
  T* get() const {
    auto& m = slots();
    auto it = m.find(this);
    return it == m.end() ? nullptr : static_cast<T*>(it->second.get());
  }
  void reset(T* p) { slots()[this] = std::shared_ptr<void>(p, [](void* q) { delete static_cast<T*>(q); }); }
 private:
  static std::unordered_map<const void*, std::shared_ptr<void> >& slots() {
    static thread_local std::unordered_map<const void*, std::shared_ptr<void> > m;
    return m;
  }
};
}  // namespace boost
