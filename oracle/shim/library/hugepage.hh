// oracle/shim: replaces rdma-library/library/hugepage.hh on the include path.
// TEST INFRASTRUCTURE ONLY.  Same public surface as the reference's HugePage<T>
// (allocate / get_slice[_unaligned] / get_full_buffer / touch_memory / deallocate,
// buffer_size, buffer_length), but backed by a lazily-committed anonymous mapping:
// the reference allocates and touches 35 GB per compute node
// (src/common/constants.hh:6, src/buffer_allocator.hh:14-15), which this box cannot afford.
// Anonymous pages read as zero, so skipping the touch loop does not change behaviour.
#pragma once
#include <sys/mman.h>

#include <cstdlib>
#include <iostream>
#include <memory>

#include <library/types.hh>
#include <library/utils.hh>

template <typename T, bool HUGE_1GB = true>
class HugePage {
public:
  HugePage() = default;
  explicit HugePage(size_t size) { allocate(size); }
  ~HugePage() { deallocate(); }
  HugePage(HugePage&) = delete;
  HugePage& operator=(HugePage&) = delete;

  void allocate(size_t size) {
    lib_assert(buffer_size == 0, "Buffer has been already allocated");
    if (size == 0) size = 64;
    void* p = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    lib_assert(p != MAP_FAILED, "oracle shim: mmap failed");
    base_ = static_cast<T*>(p);
    cursor_ = p;
    buffer_size = size;
    buffer_length = size / sizeof(T);
    left_ = size;
  }

  T* get_slice_unaligned(size_t bytes) {
    lib_assert(left_ >= bytes, "Pre-allocated memory exhausted");
    T* s = static_cast<T*>(cursor_);
    cursor_ = static_cast<byte_t*>(cursor_) + bytes;
    left_ -= bytes;
    return s;
  }

  T* get_slice(size_t bytes) {
    lib_assert(std::align(64, bytes, cursor_, left_) != nullptr, "alignment failed");
    return get_slice_unaligned(bytes);
  }

  T* get_full_buffer() const { return base_; }
  T& operator[](size_t i) { return base_[i]; }
  void touch_memory() {}  // zero pages on demand

  void deallocate() {
    if (base_ != nullptr) munmap(static_cast<void*>(base_), buffer_size);
    base_ = nullptr;
    cursor_ = nullptr;
    buffer_size = buffer_length = left_ = 0;
  }

  size_t get_memory_size() const { return 48UL << 30; }

public:
  size_t buffer_size{0};
  size_t buffer_length{0};

private:
  T* base_{nullptr};
  void* cursor_{nullptr};
  size_t left_{0};
};
