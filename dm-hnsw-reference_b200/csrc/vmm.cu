#include "vmm.h"

#include <cstdint>

namespace shn {
namespace {

struct Driver {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  bool ok = false;
};

template <class F>
bool load(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
  fn = reinterpret_cast<F>(p);
  return true;
}

const Driver& driver() {
  static Driver d = [] {
    Driver x;
    x.ok = load("cuMemCreate", x.MemCreate) && load("cuMemRelease", x.MemRelease) &&
           load("cuMemAddressReserve", x.MemAddressReserve) && load("cuMemAddressFree", x.MemAddressFree) &&
           load("cuMemMap", x.MemMap) && load("cuMemUnmap", x.MemUnmap) && load("cuMemSetAccess", x.MemSetAccess) &&
           load("cuMemGetAllocationGranularity", x.MemGetAllocationGranularity) &&
           load("cuMemExportToShareableHandle", x.MemExportToShareableHandle) &&
           load("cuMemImportFromShareableHandle", x.MemImportFromShareableHandle);
    return x;
  }();
  return d;
}

CUmemAllocationProp prop_for(int device) {
  CUmemAllocationProp p = {};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

cudaError_t map_block(const Driver& d, VmmBlock& b, int device, const char** why) {
  CUdeviceptr va = 0;
  if (d.MemAddressReserve(&va, b.size, 0, 0, 0) != CUDA_SUCCESS) { *why = "cuMemAddressReserve"; return cudaErrorMemoryAllocation; }
  if (d.MemMap(va, b.size, 0, b.handle, 0) != CUDA_SUCCESS) { d.MemAddressFree(va, b.size); *why = "cuMemMap"; return cudaErrorMemoryAllocation; }
  // read/write for the mapping device, and for every other GPU of this process that can reach it over NVLink/PCIe peer
  // access: handles that live in the SAME process exchange raw pointers (shn_index_partition_attach raw_ptrs), and a
  // VMM mapping is only visible to the devices named here (cudaDeviceEnablePeerAccess does not cover VMM memory)
  CUmemAccessDesc acc[16] = {};
  size_t n_acc = 0;
  acc[n_acc].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc[n_acc].location.id = device;
  acc[n_acc].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  ++n_acc;
  int count = 0;
  if (!b.imported && cudaGetDeviceCount(&count) == cudaSuccess) {
    for (int other = 0; other < count && n_acc < 16; ++other) {
      int can = 0;
      if (other != device && cudaDeviceCanAccessPeer(&can, other, device) == cudaSuccess && can) {
        acc[n_acc].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc[n_acc].location.id = other;
        acc[n_acc].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        ++n_acc;
      }
    }
  }
  CUresult rs = d.MemSetAccess(va, b.size, acc, n_acc);
  if (rs != CUDA_SUCCESS && n_acc > 1) rs = d.MemSetAccess(va, b.size, acc, 1);  // peers refused: at least the owner
  if (rs != CUDA_SUCCESS) {
    d.MemUnmap(va, b.size); d.MemAddressFree(va, b.size);
    *why = "cuMemSetAccess (no peer access between the two GPUs?)";
    return cudaErrorPeerAccessUnsupported;
  }
  b.ptr = reinterpret_cast<void*>(va);
  return cudaSuccess;
}

}  // namespace

cudaError_t vmm_alloc(VmmBlock& b, size_t bytes, int device, const char** why) {
  const Driver& d = driver();
  if (!d.ok) { *why = "CUDA driver lacks the virtual memory management API"; return cudaErrorNotSupported; }
  cudaFree(nullptr);  // make sure the primary context exists
  const CUmemAllocationProp p = prop_for(device);
  size_t gran = 0;
  if (d.MemGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) {
    *why = "cuMemGetAllocationGranularity"; return cudaErrorUnknown;
  }
  b = VmmBlock{};
  b.size = (bytes + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h = 0;
  if (d.MemCreate(&h, b.size, &p, 0) != CUDA_SUCCESS) { *why = "cuMemCreate"; return cudaErrorMemoryAllocation; }
  b.handle = h;
  const cudaError_t e = map_block(d, b, device, why);
  if (e != cudaSuccess) { d.MemRelease(h); b = VmmBlock{}; }
  return e;
}

cudaError_t vmm_export_fd(const VmmBlock& b, int* fd, const char** why) {
  const Driver& d = driver();
  if (!d.ok || !b.handle) { *why = "nothing to export"; return cudaErrorInvalidValue; }
  int out = -1;
  if (d.MemExportToShareableHandle(&out, b.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) {
    *why = "cuMemExportToShareableHandle"; return cudaErrorUnknown;
  }
  *fd = out;
  return cudaSuccess;
}

cudaError_t vmm_import_fd(VmmBlock& b, int fd, size_t size, int device, const char** why) {
  const Driver& d = driver();
  if (!d.ok) { *why = "CUDA driver lacks the virtual memory management API"; return cudaErrorNotSupported; }
  cudaFree(nullptr);
  b = VmmBlock{};
  b.size = size;
  b.imported = true;
  CUmemGenericAllocationHandle h = 0;
  if (d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) {
    *why = "cuMemImportFromShareableHandle"; b = VmmBlock{}; return cudaErrorUnknown;
  }
  b.handle = h;
  const cudaError_t e = map_block(d, b, device, why);
  if (e != cudaSuccess) { d.MemRelease(h); b = VmmBlock{}; }
  return e;
}

cudaError_t vmm_granularity(int device, size_t* gran, const char** why) {
  const Driver& d = driver();
  if (!d.ok) { *why = "CUDA driver lacks the virtual memory management API"; return cudaErrorNotSupported; }
  cudaFree(nullptr);
  const CUmemAllocationProp p = prop_for(device);
  if (d.MemGetAllocationGranularity(gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || *gran == 0) {
    *why = "cuMemGetAllocationGranularity"; return cudaErrorUnknown;
  }
  return cudaSuccess;
}

cudaError_t vmm_reserve(VmmSpace& s, size_t bytes, size_t align, const char** why) {
  const Driver& d = driver();
  if (!d.ok) { *why = "CUDA driver lacks the virtual memory management API"; return cudaErrorNotSupported; }
  cudaFree(nullptr);
  CUdeviceptr va = 0;
  if (d.MemAddressReserve(&va, bytes, align, 0, 0) != CUDA_SUCCESS) { *why = "cuMemAddressReserve"; return cudaErrorMemoryAllocation; }
  s.base = reinterpret_cast<void*>(va);
  s.size = bytes;
  return cudaSuccess;
}

void vmm_release(VmmSpace& s) {
  const Driver& d = driver();
  if (d.ok && s.base) d.MemAddressFree(reinterpret_cast<CUdeviceptr>(s.base), s.size);
  s = VmmSpace{};
}

cudaError_t vmm_create(VmmBlock& b, size_t bytes, int device, const char** why) {
  const Driver& d = driver();
  size_t gran = 0;
  cudaError_t e = vmm_granularity(device, &gran, why);
  if (e != cudaSuccess) return e;
  const CUmemAllocationProp p = prop_for(device);
  b = VmmBlock{};
  b.size = (bytes + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h = 0;
  if (d.MemCreate(&h, b.size, &p, 0) != CUDA_SUCCESS) { *why = "cuMemCreate"; b = VmmBlock{}; return cudaErrorMemoryAllocation; }
  b.handle = h;
  return cudaSuccess;
}

cudaError_t vmm_import_handle(VmmBlock& b, int fd, size_t size, const char** why) {
  const Driver& d = driver();
  if (!d.ok) { *why = "CUDA driver lacks the virtual memory management API"; return cudaErrorNotSupported; }
  cudaFree(nullptr);
  b = VmmBlock{};
  b.size = size;
  b.imported = true;
  CUmemGenericAllocationHandle h = 0;
  if (d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) {
    *why = "cuMemImportFromShareableHandle"; b = VmmBlock{}; return cudaErrorUnknown;
  }
  b.handle = h;
  return cudaSuccess;
}

cudaError_t vmm_place(const VmmSpace& s, size_t offset, size_t size, unsigned long long handle, int device, const char** why) {
  const Driver& d = driver();
  if (!d.ok || !s.base || offset + size > s.size) { *why = "bad placement"; return cudaErrorInvalidValue; }
  const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(s.base) + offset;
  if (d.MemMap(va, size, 0, handle, 0) != CUDA_SUCCESS) { *why = "cuMemMap"; return cudaErrorMemoryAllocation; }
  CUmemAccessDesc acc = {};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (d.MemSetAccess(va, size, &acc, 1) != CUDA_SUCCESS) {
    d.MemUnmap(va, size);
    *why = "cuMemSetAccess (no peer access between the two GPUs?)";
    return cudaErrorPeerAccessUnsupported;
  }
  return cudaSuccess;
}

void vmm_unplace(const VmmSpace& s, size_t offset, size_t size) {
  const Driver& d = driver();
  if (d.ok && s.base) d.MemUnmap(reinterpret_cast<CUdeviceptr>(s.base) + offset, size);
}

void vmm_drop(VmmBlock& b) {
  const Driver& d = driver();
  if (d.ok && b.handle) d.MemRelease(b.handle);
  b = VmmBlock{};
}

void vmm_free(VmmBlock& b) {
  const Driver& d = driver();
  if (!d.ok || !b.ptr) { b = VmmBlock{}; return; }
  const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(b.ptr);
  d.MemUnmap(va, b.size);
  d.MemRelease(b.handle);
  d.MemAddressFree(va, b.size);
  b = VmmBlock{};
}

}  // namespace shn
