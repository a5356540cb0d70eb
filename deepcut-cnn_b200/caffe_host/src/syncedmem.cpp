#include "caffe/syncedmem.hpp"

#include <cstdlib>
#include <cstring>

#include "deepcut_b200.h"

namespace caffe {

// An upload from pinned memory is asynchronous (the reference's is a synchronous cudaMemcpy, syncedmem.cpp:61-69): until it has
// finished the host buffer still belongs to the DMA engine.  Every path that lets the host WRITE the buffer again (or frees
// it) first waits for the event recorded behind the copy -- `blob.data[...] = a; net.forward(); blob.data[...] = b` must not
// change what the first forward reads.
void SyncedMemory::mark_upload(void* stream) {
  if (!cpu_malloc_use_cuda_) return;         // pageable source: the caller synchronises the stream
  if (upload_event_ == nullptr) DC_CHECK(dc_event_create(&upload_event_));
  DC_CHECK(dc_event_record(upload_event_, stream));
  upload_pending_ = true;
}
void SyncedMemory::wait_upload() {
  if (!upload_pending_) return;
  DC_CHECK(dc_event_sync(upload_event_));
  upload_pending_ = false;
}

SyncedMemory::~SyncedMemory() {
  if (upload_pending_) dc_event_sync(upload_event_);
  if (upload_event_) dc_event_destroy(upload_event_);
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
      else mark_upload(Caffe::stream());
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
  wait_upload();
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
  ++host_epoch_;
  head_ = HEAD_AT_GPU;
  own_gpu_data_ = false;
}

void* SyncedMemory::mutable_cpu_data() { to_cpu(); wait_upload(); head_ = HEAD_AT_CPU; ++host_epoch_; return cpu_ptr_; }
// device-side writers count as writes too (a parameter blob updated through mutable_gpu_data must invalidate packed copies of it)
void* SyncedMemory::mutable_gpu_data() { to_gpu(); head_ = HEAD_AT_GPU; ++host_epoch_; return gpu_ptr_; }

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
    gpu_device_ = Caffe::device();
    own_gpu_data_ = true;
  }
  DC_CHECK(dc_memcpy_async(gpu_ptr_, cpu_ptr_, size_, DC_H2D, stream));
  if (!cpu_malloc_use_cuda_) DC_CHECK(dc_stream_sync(stream));
  else mark_upload(stream);
  head_ = SYNCED;
}

}  // namespace caffe
