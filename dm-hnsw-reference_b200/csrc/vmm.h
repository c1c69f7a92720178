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
};

// Allocate `bytes` of device memory on `device`, read/write mapped for that device, exportable as a POSIX fd.
cudaError_t vmm_alloc(VmmBlock& b, size_t bytes, int device, const char** why);
// File descriptor for the block (the caller owns and closes it).
cudaError_t vmm_export_fd(const VmmBlock& b, int* fd, const char** why);
// Map a block exported by another process (or this one) into this process for `device`.
cudaError_t vmm_import_fd(VmmBlock& b, int fd, size_t size, int device, const char** why);
void vmm_free(VmmBlock& b);

// One flat address range per array of a partitioned index (graph.h "flat numbering"): the replicated hot set, this GPU's
// share, every peer's share and the halo are separate physical allocations mapped side by side into ONE reservation, so
// that a row is at base + row * stride wherever it lives and the MMU does what a lookup table would otherwise do.
struct VmmSpace {
  void* base = nullptr;
  size_t size = 0;
};
// physical allocation granularity on `device` (pieces of a space start at multiples of it)
cudaError_t vmm_granularity(int device, size_t* gran, const char** why);
cudaError_t vmm_reserve(VmmSpace& s, size_t bytes, size_t align, const char** why);
void vmm_release(VmmSpace& s);
// A physical allocation that is not mapped anywhere yet (exportable; b.ptr stays null until it is placed).
cudaError_t vmm_create(VmmBlock& b, size_t bytes, int device, const char** why);
// Import without mapping.
cudaError_t vmm_import_handle(VmmBlock& b, int fd, size_t size, const char** why);
// Map allocation `handle` (size bytes) at s.base + offset, read/write for `device` (and, unless peer_access is false, every GPU of this
// process that can reach it).  The same allocation may be placed into several spaces.
cudaError_t vmm_place(const VmmSpace& s, size_t offset, size_t size, unsigned long long handle, int device, const char** why);
void vmm_unplace(const VmmSpace& s, size_t offset, size_t size);
// Drop the allocation handle (mappings made from it stay valid until unmapped).
void vmm_drop(VmmBlock& b);

}  // namespace shn
