#include "caffe/syncedmem.hpp"

#include <cstdlib>
#include <cstring>

#include "deepcut_b200.h"

namespace caffe {

SyncedMemory::~SyncedMemory() {
  if (cpu_ptr_ && own_cpu_data_) {
    if (cpu_malloc_use_cuda_) dc_free_host(cpu_ptr_);
    else free(cpu_ptr_);
  }
  if (gpu_ptr_ && own_gpu_data_) dc_free(gpu_ptr_);
}

void SyncedMemory::alloc_cpu() {
  // pinned iff GPU mode at first touch (reference syncedmem.hpp:15-27), so H2D/D2H can be async
  if (Caffe::mode() == Caffe::GPU) {
    DC_CHECK(dc_malloc_host(&cpu_ptr_, size_ ? size_ : 1));
    cpu_malloc_use_cuda_ = true;
  } else {
    cpu_ptr_ = malloc(size_ ? size_ : 1);
    CHECK(cpu_ptr_) << "host allocation of size " << size_ << " failed";
    cpu_malloc_use_cuda_ = false;
  }
  own_cpu_data_ = true;
}

inline void SyncedMemory::to_cpu() {
  switch (head_) {
    case UNINITIALIZED:
      alloc_cpu();
      memset(cpu_ptr_, 0, size_);
      head_ = HEAD_AT_CPU;
      break;
    case HEAD_AT_GPU:
      if (cpu_ptr_ == nullptr) alloc_cpu();
      DC_CHECK(dc_memcpy_async(cpu_ptr_, gpu_ptr_, size_, DC_D2H, Caffe::stream()));
      DC_CHECK(dc_stream_sync(Caffe::stream()));   // the host is about to read
      head_ = SYNCED;
      break;
    case HEAD_AT_CPU:
    case SYNCED:
      break;
  }
}

inline void SyncedMemory::to_gpu() {
  switch (head_) {
    case UNINITIALIZED:
      DC_CHECK(dc_malloc(&gpu_ptr_, size_ ? size_ : 1));
      DC_CHECK(dc_memset_async(gpu_ptr_, 0, size_, Caffe::stream()));
      gpu_device_ = Caffe::device();
      head_ = HEAD_AT_GPU;
      own_gpu_data_ = true;
      break;
    case HEAD_AT_CPU:
      if (gpu_ptr_ == nullptr) {
        DC_CHECK(dc_malloc(&gpu_ptr_, size_ ? size_ : 1));
        gpu_device_ = Caffe::device();
        own_gpu_data_ = true;
      }
      DC_CHECK(dc_memcpy_async(gpu_ptr_, cpu_ptr_, size_, DC_H2D, Caffe::stream()));
      if (!cpu_malloc_use_cuda_) DC_CHECK(dc_stream_sync(Caffe::stream()));   // pageable source
      head_ = SYNCED;
      break;
    case HEAD_AT_GPU:
    case SYNCED:
      break;
  }
}

const void* SyncedMemory::cpu_data() { to_cpu(); return cpu_ptr_; }

void SyncedMemory::set_cpu_data(void* data) {
  CHECK(data);
  if (own_cpu_data_ && cpu_ptr_) { if (cpu_malloc_use_cuda_) dc_free_host(cpu_ptr_); else free(cpu_ptr_); }
  cpu_ptr_ = data;
  ++host_epoch_;
  head_ = HEAD_AT_CPU;
  own_cpu_data_ = false;
  cpu_malloc_use_cuda_ = false;
}

const void* SyncedMemory::gpu_data() { to_gpu(); return gpu_ptr_; }

void SyncedMemory::set_gpu_data(void* data) {
  CHECK(data);
  if (own_gpu_data_ && gpu_ptr_) dc_free(gpu_ptr_);
  gpu_ptr_ = data;
  head_ = HEAD_AT_GPU;
  own_gpu_data_ = false;
}

void* SyncedMemory::mutable_cpu_data() { to_cpu(); head_ = HEAD_AT_CPU; ++host_epoch_; return cpu_ptr_; }
void* SyncedMemory::mutable_gpu_data() { to_gpu(); head_ = HEAD_AT_GPU; return gpu_ptr_; }

void* SyncedMemory::overwrite_gpu_data() {
  if (gpu_ptr_ == nullptr) {
    DC_CHECK(dc_malloc(&gpu_ptr_, size_ ? size_ : 1));
    gpu_device_ = Caffe::device();
    own_gpu_data_ = true;
  }
  head_ = HEAD_AT_GPU;
  return gpu_ptr_;
}

void SyncedMemory::async_gpu_push(void* stream) {
  CHECK(head_ == HEAD_AT_CPU);
  if (gpu_ptr_ == nullptr) {
    DC_CHECK(dc_malloc(&gpu_ptr_, size_ ? size_ : 1));
    own_gpu_data_ = true;
  }
  DC_CHECK(dc_memcpy_async(gpu_ptr_, cpu_ptr_, size_, DC_H2D, stream));
  head_ = SYNCED;
}

}  // namespace caffe
