// vmm.h — shares of a partitioned index are allocated with the CUDA virtual-memory-management API (cuMemCreate with
// the device's recommended granularity, i.e. 2 MiB physical pages) and exported as POSIX file descriptors, the way NCCL
// shares buffers between the processes of a node.  Legacy cudaIpc handles of cudaMalloc memory were measured to fall off
// a cliff (about 60 GB/s instead of 770 GB/s peer reads) for shares carved out of a process that had already churned
// through several GB of allocations.  The driver entry points are fetched with cudaGetDriverEntryPoint so that the
// library does not link libcuda (it must load on a machine without a driver for the ABI tests).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstddef>

namespace shn {

struct VmmBlock {
  void* ptr = nullptr;      // mapped address in this process
  size_t size = 0;          // padded to the allocation granularity
  unsigned long long handle = 0;  // CUmemGenericAllocationHandle
  bool imported = false;
  bool plain = false;       // SHN_SHARE_PLAIN=1 diagnostic: cudaMalloc instead of cuMemCreate
};

// Allocate `bytes` of device memory on `device`, read/write mapped for that device, exportable as a POSIX fd.
cudaError_t vmm_alloc(VmmBlock& b, size_t bytes, int device, const char** why);
// File descriptor for the block (the caller owns and closes it).
cudaError_t vmm_export_fd(const VmmBlock& b, int* fd, const char** why);
// Map a block exported by another process (or this one) into this process for `device`.
cudaError_t vmm_import_fd(VmmBlock& b, int fd, size_t size, int device, const char** why);
void vmm_free(VmmBlock& b);

}  // namespace shn
