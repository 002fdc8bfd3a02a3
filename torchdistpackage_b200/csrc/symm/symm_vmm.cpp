// Native symmetric memory: CUDA VMM allocations shared between the processes of one node through
// POSIX file descriptors, peer-mapped into every rank's address space, plus one NVLS multicast
// object bound to all of them (multimem.* target).  This is the C++ runtime under
// torchdistpackage_b200.ops.symm (backend "native"); handle *transport* (who gets which fd) is done
// by the Python layer with pidfd_getfd over the process group.
//
// Only driver entry points resolved at run time are used, so the extension still imports on a
// CPU-only host.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../common/tdp_api.h"

namespace tdp {

namespace {

template <typename F>
F drv(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<F>(p);
}

#define TDP_DRV(name) static auto fn_##name = drv<decltype(&name)>(#name)

std::string cu_err(CUresult r) {
  TDP_DRV(cuGetErrorString);
  const char* s = nullptr;
  if (fn_cuGetErrorString) fn_cuGetErrorString(r, &s);
  return std::string(s ? s : "unknown") + " (" + std::to_string(static_cast<int>(r)) + ")";
}

#define TDP_CU(call, what)                                                     \
  do {                                                                         \
    CUresult _r = (call);                                                      \
    if (_r != CUDA_SUCCESS) { err = std::string(what) + ": " + cu_err(_r); return false; } \
  } while (0)

CUmemAllocationProp alloc_prop(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

bool map_rw(CUmemGenericAllocationHandle h, size_t size, int device, uint64_t* ptr_out,
            std::string& err) {
  TDP_DRV(cuMemAddressReserve);
  TDP_DRV(cuMemMap);
  TDP_DRV(cuMemSetAccess);
  CUdeviceptr va = 0;
  TDP_CU(fn_cuMemAddressReserve(&va, size, 0, 0, 0), "cuMemAddressReserve");
  TDP_CU(fn_cuMemMap(va, size, 0, h, 0), "cuMemMap");
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  TDP_CU(fn_cuMemSetAccess(va, size, &acc, 1), "cuMemSetAccess");
  *ptr_out = static_cast<uint64_t>(va);
  return true;
}

}  // namespace

// ---- allocation granularity that satisfies both the VMM allocator and multicast binding
bool vmm_granularity(int device, int num_devices, uint64_t* gran, std::string& err) {
  TDP_DRV(cuMemGetAllocationGranularity);
  TDP_DRV(cuMulticastGetGranularity);
  if (!fn_cuMemGetAllocationGranularity) { err = "CUDA driver without VMM support"; return false; }
  CUmemAllocationProp prop = alloc_prop(device);
  size_t g = 0;
  TDP_CU(fn_cuMemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
         "cuMemGetAllocationGranularity");
  if (fn_cuMulticastGetGranularity && num_devices > 1) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = num_devices;
    mp.size = g;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (fn_cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS &&
        mg > g)
      g = mg;
  }
  *gran = g;
  return true;
}

// ---- local allocation: returns generic handle, mapped pointer and an exportable fd
bool vmm_alloc(uint64_t size, int device, uint64_t* handle, uint64_t* ptr, int* fd,
               std::string& err) {
  TDP_DRV(cuMemCreate);
  TDP_DRV(cuMemExportToShareableHandle);
  if (!fn_cuMemCreate) { err = "CUDA driver without VMM support"; return false; }
  cudaSetDevice(device);
  cudaFree(nullptr);
  CUmemAllocationProp prop = alloc_prop(device);
  CUmemGenericAllocationHandle h;
  TDP_CU(fn_cuMemCreate(&h, size, &prop, 0), "cuMemCreate");
  if (!map_rw(h, size, device, ptr, err)) return false;
  int out_fd = -1;
  TDP_CU(fn_cuMemExportToShareableHandle(&out_fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
         "cuMemExportToShareableHandle");
  *handle = static_cast<uint64_t>(h);
  *fd = out_fd;
  return true;
}

// ---- import a peer's allocation from (a duplicate of) its fd and map it on `device`
bool vmm_import(int fd, uint64_t size, int device, uint64_t* handle, uint64_t* ptr,
                std::string& err) {
  TDP_DRV(cuMemImportFromShareableHandle);
  CUmemGenericAllocationHandle h;
  TDP_CU(fn_cuMemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                           CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
         "cuMemImportFromShareableHandle");
  if (!map_rw(h, size, device, ptr, err)) return false;
  *handle = static_cast<uint64_t>(h);
  return true;
}

// ---- multicast object: created by one rank (exports an fd), imported by the others
bool mc_create(uint64_t size, int num_devices, uint64_t* handle, int* fd, std::string& err) {
  TDP_DRV(cuMulticastCreate);
  TDP_DRV(cuMemExportToShareableHandle);
  if (!fn_cuMulticastCreate) { err = "driver without multicast support"; return false; }
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = num_devices;
  mp.size = size;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h;
  TDP_CU(fn_cuMulticastCreate(&h, &mp), "cuMulticastCreate");
  int out_fd = -1;
  TDP_CU(fn_cuMemExportToShareableHandle(&out_fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
         "cuMemExportToShareableHandle(multicast)");
  *handle = static_cast<uint64_t>(h);
  *fd = out_fd;
  return true;
}

bool mc_import(int fd, uint64_t* handle, std::string& err) {
  TDP_DRV(cuMemImportFromShareableHandle);
  CUmemGenericAllocationHandle h;
  TDP_CU(fn_cuMemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                           CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
         "cuMemImportFromShareableHandle(multicast)");
  *handle = static_cast<uint64_t>(h);
  return true;
}

bool mc_add_device(uint64_t mc_handle, int device, std::string& err) {
  TDP_DRV(cuMulticastAddDevice);
  TDP_DRV(cuDeviceGet);
  CUdevice dev;
  TDP_CU(fn_cuDeviceGet(&dev, device), "cuDeviceGet");
  TDP_CU(fn_cuMulticastAddDevice(static_cast<CUmemGenericAllocationHandle>(mc_handle), dev),
         "cuMulticastAddDevice");
  return true;
}

// bind my physical allocation at offset 0 of the multicast object, then map the multicast VA
bool mc_bind_and_map(uint64_t mc_handle, uint64_t mem_handle, uint64_t size, int device,
                     uint64_t* mc_ptr, std::string& err) {
  TDP_DRV(cuMulticastBindMem);
  TDP_CU(fn_cuMulticastBindMem(static_cast<CUmemGenericAllocationHandle>(mc_handle), 0,
                               static_cast<CUmemGenericAllocationHandle>(mem_handle), 0, size, 0),
         "cuMulticastBindMem");
  return map_rw(static_cast<CUmemGenericAllocationHandle>(mc_handle), size, device, mc_ptr, err);
}

bool vmm_unmap_release(uint64_t handle, uint64_t ptr, uint64_t size, std::string& err) {
  TDP_DRV(cuMemUnmap);
  TDP_DRV(cuMemAddressFree);
  TDP_DRV(cuMemRelease);
  if (ptr) {
    TDP_CU(fn_cuMemUnmap(static_cast<CUdeviceptr>(ptr), size), "cuMemUnmap");
    TDP_CU(fn_cuMemAddressFree(static_cast<CUdeviceptr>(ptr), size), "cuMemAddressFree");
  }
  if (handle) TDP_CU(fn_cuMemRelease(static_cast<CUmemGenericAllocationHandle>(handle)), "cuMemRelease");
  return true;
}

}  // namespace tdp
