#include "caffe/common.hpp"

#include <cstdio>
#include <cstring>

#include "deepcut_b200.h"

namespace caffe {

namespace {
bool g_fatal_throws = false;
int g_min_log_level = 1;   // INFO is chatty (the reference prints the whole net); opt in via set_log_level(0)
}  // namespace

LogMessage::LogMessage(const char* file, int line, LogSeverity sev) : sev_(sev) {
  const char* base = strrchr(file, '/');
  static const char kTag[] = {'I', 'W', 'E', 'F'};
  ss_ << kTag[sev] << " " << (base ? base + 1 : file) << ":" << line << "] ";
}

LogMessage::~LogMessage() noexcept(false) {
  if (sev_ == FATAL) {
    const string msg = ss_.str();
    if (g_fatal_throws) throw FatalError(msg);
    fprintf(stderr, "%s\n*** Check failure stack trace: (none; deepcut-cnn_b200 host) ***\n", msg.c_str());
    fflush(stderr);
    abort();
  }
  if (static_cast<int>(sev_) >= g_min_log_level) fprintf(stderr, "%s\n", ss_.str().c_str());
}

const char* DcLastError() { return dc_last_error(); }

Caffe& Caffe::Get() {
  static thread_local Caffe instance;
  return instance;
}

void Caffe::set_fatal_throws(bool v) { g_fatal_throws = v; }
bool Caffe::fatal_throws() { return g_fatal_throws; }
void Caffe::set_log_level(int min_severity) { g_min_log_level = min_severity; }

int Caffe::device_count() { return dc_device_count(); }

void Caffe::SetDevice(const int device_id) {
  Caffe& c = Get();
  if (c.device_ == device_id && c.stream_ != nullptr) return;
  DC_CHECK(dc_init(device_id));
  if (c.stream_ != nullptr) dc_stream_destroy(c.stream_);
  c.stream_ = nullptr;
  DC_CHECK(dc_stream_create(&c.stream_));
  c.device_ = device_id;
}

void* Caffe::stream() {
  Caffe& c = Get();
  if (c.stream_ == nullptr) SetDevice(c.device_ < 0 ? 0 : c.device_);
  return c.stream_;
}

void Caffe::DeviceQuery() {
  size_t free_b = 0, total_b = 0;
  DC_CHECK(dc_mem_info(&free_b, &total_b));
  LOG(WARNING) << "Device id: " << Get().device_ << "  sm_100 devices: " << dc_device_count() << "  memory free/total MiB: "
               << (free_b >> 20) << "/" << (total_b >> 20);
}

}  // namespace caffe
