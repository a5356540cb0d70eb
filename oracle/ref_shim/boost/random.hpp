// boost.random stand-ins over <random> for src/caffe/util/math_functions.cpp:229-326 and include/caffe/util/rng.hpp.
#pragma once
#include <random>
namespace boost {
typedef std::mt19937 mt19937;
template <class T = double> using uniform_real = std::uniform_real_distribution<T>;
template <class T = double> using normal_distribution = std::normal_distribution<T>;
template <class T = int> using uniform_int = std::uniform_int_distribution<T>;
template <class T = double>
class bernoulli_distribution {
 public:
  explicit bernoulli_distribution(T p) : d_(static_cast<double>(p)) {}
  template <class E> bool operator()(E& e) { return d_(e); }
 private:
  std::bernoulli_distribution d_;
};
template <class EnginePtr, class Dist>
class variate_generator {
 public:
  variate_generator(EnginePtr e, Dist d) : e_(e), d_(d) {}
  auto operator()() -> decltype(std::declval<Dist&>()(*std::declval<EnginePtr&>())) { return d_(*e_); }
 private:
  EnginePtr e_;
  Dist d_;
};
}  // namespace boost
