// SyncedMemory: lazily allocated host/device mirror with the reference's 4-state head
// (include/caffe/syncedmem.hpp:40-84, src/caffe/syncedmem.cpp:7-157).  Differences by design:
// device memory and pinned host memory come from the C ABI (dc_malloc / dc_malloc_host); copies are
// stream-ordered on Caffe::stream() and only synchronised when the host actually reads.
#pragma once
#include <cstddef>

#include "caffe/common.hpp"

namespace caffe {

class SyncedMemory {
 public:
  SyncedMemory() {}
  explicit SyncedMemory(size_t size) : size_(size) {}
  ~SyncedMemory();
  const void* cpu_data();
  void set_cpu_data(void* data);
  const void* gpu_data();
  void set_gpu_data(void* data);
  void* mutable_cpu_data();
  void* mutable_gpu_data();
  // Device pointer for a consumer that overwrites every element: allocates if needed and moves the
  // head to the GPU WITHOUT uploading a host copy first.  (mutable_gpu_data() keeps the reference's
  // semantics, syncedmem.cpp:61-69,124-132: a blob the host has read is re-uploaded before a layer
  // writes its top -- 335 MB of pointless H2D per forward for next_pred at 16x720p.)
  void* overwrite_gpu_data();
  enum SyncedHead { UNINITIALIZED, HEAD_AT_CPU, HEAD_AT_GPU, SYNCED };
  SyncedHead head() { return head_; }
  size_t size() { return size_; }
  void async_gpu_push(void* stream);
  // Bumped by every mutable_cpu_data()/set_cpu_data(): lets device-side caches of transformed
  // weights (packed split-fp16 matrices) notice host writes through Blob::mutable_cpu_data or a
  // pycaffe `.data` view, which the reference needs no notification for.
  unsigned long long host_write_epoch() const { return host_epoch_; }

 private:
  void to_cpu();
  void to_gpu();
  void alloc_cpu();
  void mark_upload(void* stream);     // an async H2D from the pinned host buffer was queued on `stream`
  void wait_upload();                 // ... and must have finished before the host may write that buffer again
  void* upload_event_ = nullptr;
  bool upload_pending_ = false;
  void* cpu_ptr_ = nullptr;
  void* gpu_ptr_ = nullptr;
  size_t size_ = 0;
  SyncedHead head_ = UNINITIALIZED;
  bool own_cpu_data_ = false;
  bool cpu_malloc_use_cuda_ = false;
  bool own_gpu_data_ = false;
  int gpu_device_ = -1;
  unsigned long long host_epoch_ = 0;
  DISABLE_COPY_AND_ASSIGN(SyncedMemory);
};

}  // namespace caffe
