// oracle/shim: stand-in for oneTBB's concurrent_vector (absent in this image).
// TEST INFRASTRUCTURE ONLY.  The reference pulls the type in through
// rdma-library/library/types.hh:4-5,54 and only ever calls push_back on it
// (src/buffer_allocator.hh:79,96), so a mutex-guarded std::vector is enough.
#pragma once
#include <mutex>
#include <vector>

namespace oneapi::tbb {
template <typename T>
class concurrent_vector {
public:
  void push_back(const T& v) {
    std::lock_guard<std::mutex> g(mu_);
    data_.push_back(v);
  }
  size_t size() const { return data_.size(); }
  auto begin() const { return data_.begin(); }
  auto end() const { return data_.end(); }

private:
  std::mutex mu_;
  std::vector<T> data_;
};
}  // namespace oneapi::tbb
