// Minimal glog stand-in for building the reference's CPU layer sources as the oracle's "_ref"
// (TEST INFRASTRUCTURE; glog is not installed in this image).  CHECK* / LOG(sev) with stream syntax;
// FATAL throws std::runtime_error so a failed reference CHECK surfaces in the Python test.
#pragma once
#include <unistd.h>
#include <cstring>
#include <cstdlib>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

namespace refshim {
enum Severity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
class LogMessage {
 public:
  LogMessage(const char* file, int line, int sev) : sev_(sev) { ss_ << file << ":" << line << "] "; }
  ~LogMessage() noexcept(false) {
    if (sev_ == FATAL) throw std::runtime_error(ss_.str());
    if (sev_ >= WARNING && std::getenv("REFCAFFE_VERBOSE")) std::cerr << ss_.str() << std::endl;
  }
  std::ostream& stream() { return ss_; }
 private:
  std::ostringstream ss_;
  int sev_;
};
struct Voidify { void operator&(std::ostream&) {} };
template <class T> T* CheckNotNull(const char* f, int l, const char* what, T* p) {
  if (!p) LogMessage(f, l, FATAL).stream() << what << " must be non NULL";
  return p;
}
}  // namespace refshim

#define LOG(sev) ::refshim::LogMessage(__FILE__, __LINE__, ::refshim::sev).stream()
#define LOG_IF(sev, cond) !(cond) ? (void)0 : ::refshim::Voidify() & LOG(sev)
#define LOG_FIRST_N(sev, n) LOG(sev)
#define LOG_EVERY_N(sev, n) LOG(sev)
#define VLOG(n) LOG(INFO)
#define CHECK(c) (c) ? (void)0 : ::refshim::Voidify() & LOG(FATAL) << "Check failed: " #c " "
#define REFSHIM_OP(a, b, op) ((a)op(b)) ? (void)0 : ::refshim::Voidify() & LOG(FATAL) << "Check failed: " #a " " #op " " #b " (" << (a) << " vs. " << (b) << ") "
#define CHECK_EQ(a, b) REFSHIM_OP(a, b, ==)
#define CHECK_NE(a, b) REFSHIM_OP(a, b, !=)
#define CHECK_LE(a, b) REFSHIM_OP(a, b, <=)
#define CHECK_LT(a, b) REFSHIM_OP(a, b, <)
#define CHECK_GE(a, b) REFSHIM_OP(a, b, >=)
#define CHECK_GT(a, b) REFSHIM_OP(a, b, >)
#define CHECK_NOTNULL(p) ::refshim::CheckNotNull(__FILE__, __LINE__, #p, (p))
#define DCHECK(c) CHECK(c)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
namespace google {
inline void InitGoogleLogging(const char*) {}
inline void InstallFailureSignalHandler() {}
}  // namespace google
