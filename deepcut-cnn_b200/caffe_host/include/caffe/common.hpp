// Caffe singleton + logging/CHECK macros on std C++ (the reference builds these on glog, gflags
// and boost: include/caffe/common.hpp:1-182, src/caffe/common.cpp).  Public names are kept:
// Caffe::set_mode / mode / SetDevice / Get, CHECK*, LOG(severity), shared_ptr, NOT_IMPLEMENTED.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace caffe {

using std::map;
using std::ostringstream;
using std::pair;
using std::set;
using std::shared_ptr;
using std::string;
using std::vector;

// ------------------------------------------------------------------------------------ logging
enum LogSeverity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };

// Thrown instead of abort() when Caffe::set_fatal_throws(true) (the Python binding sets it so a
// failed CHECK becomes a Python exception rather than killing the interpreter).
class FatalError : public std::runtime_error {
 public:
  explicit FatalError(const string& m) : std::runtime_error(m) {}
};

class LogMessage {
 public:
  LogMessage(const char* file, int line, LogSeverity sev);
  ~LogMessage() noexcept(false);
  std::ostream& stream() { return ss_; }

 private:
  ostringstream ss_;
  LogSeverity sev_;
};
struct LogVoidify { void operator&(std::ostream&) {} };

#define LOG(sev) ::caffe::LogMessage(__FILE__, __LINE__, ::caffe::sev).stream()
#define LOG_IF(sev, cond) !(cond) ? (void)0 : ::caffe::LogVoidify() & LOG(sev)
#define CHECK(cond) (cond) ? (void)0 : ::caffe::LogVoidify() & LOG(FATAL) << "Check failed: " #cond " "
#define CAFFE_CHECK_OP(a, b, op)                                                               \
  ((a)op(b)) ? (void)0                                                                         \
             : ::caffe::LogVoidify() & LOG(FATAL) << "Check failed: " #a " " #op " " #b " (" << (a) << " vs. " << (b) << ") "
#define CHECK_EQ(a, b) CAFFE_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) CAFFE_CHECK_OP(a, b, !=)
#define CHECK_LE(a, b) CAFFE_CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) CAFFE_CHECK_OP(a, b, <)
#define CHECK_GE(a, b) CAFFE_CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) CAFFE_CHECK_OP(a, b, >)
#define CHECK_NOTNULL(p) ::caffe::CheckNotNull(__FILE__, __LINE__, #p, (p))
#define DCHECK(c) CHECK(c)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define NOT_IMPLEMENTED LOG(FATAL) << "Not Implemented Yet"
// Wraps a call into the C ABI (include/deepcut_b200.h): non-zero -> CHECK failure with its message.
#define DC_CHECK(call)                                                                        \
  do {                                                                                        \
    int dc_rc_ = (call);                                                                      \
    CHECK_EQ(dc_rc_, 0) << #call << ": " << ::caffe::DcLastError();                           \
  } while (0)

template <class T>
T* CheckNotNull(const char* file, int line, const char* expr, T* p) {
  if (p == nullptr) LogMessage(file, line, FATAL).stream() << "'" << expr << "' must be non NULL";
  return p;
}
const char* DcLastError();

#define DISABLE_COPY_AND_ASSIGN(classname) \
 private:                                  \
  classname(const classname&) = delete;    \
  classname& operator=(const classname&) = delete
#define INSTANTIATE_CLASS(classname) template class classname<float>

// ------------------------------------------------------------------------------------ Caffe
// Thread-local like the reference (src/caffe/common.cpp:13-20): mode and device are per thread.
class Caffe {
 public:
  enum Brew { CPU, GPU };
  static Caffe& Get();
  static Brew mode() { return Get().mode_; }
  static void set_mode(Brew mode) { Get().mode_ = mode; }
  // Binds this thread to `device_id` and creates its forward stream (replaces the cuBLAS/cuRAND
  // handle setup of Caffe::SetDevice, src/caffe/common.cpp:140-158).
  static void SetDevice(const int device_id);
  static int device() { return Get().device_; }
  static void* stream();                 // cudaStream_t of this thread's forwards
  static void DeviceQuery();
  static int device_count();
  static void set_random_seed(const unsigned int seed) { Get().seed_ = seed; }
  static unsigned int random_seed() { return Get().seed_; }
  static bool root_solver() { return true; }
  static int solver_count() { return 1; }
  static void set_fatal_throws(bool v);
  static bool fatal_throws();
  static void set_log_level(int min_severity);   // 0 prints INFO (the reference's glog default)

 private:
  Caffe() {}
  Brew mode_ = CPU;
  int device_ = -1;
  void* stream_ = nullptr;
  unsigned int seed_ = 1701;
};

}  // namespace caffe
