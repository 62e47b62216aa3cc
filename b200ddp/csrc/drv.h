// CUDA *driver* API entry points resolved at run time through cudaGetDriverEntryPoint, so the
// extension links only against cudart and still imports on a box without libcuda.so.1 (the CPU
// sandbox).  Used for VMM (cuMem*), NVLS multicast (cuMulticast*) and TMA descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

namespace b200 {

struct Driver {
  decltype(&cuGetErrorString) GetErrorString = nullptr;
  decltype(&cuDeviceGet) DeviceGet = nullptr;
  decltype(&cuDeviceGetAttribute) DeviceGetAttribute = nullptr;
  decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
  decltype(&cuMemCreate) MemCreate = nullptr;
  decltype(&cuMemRelease) MemRelease = nullptr;
  decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
  decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
  decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
  decltype(&cuMemAddressFree) MemAddressFree = nullptr;
  decltype(&cuMemMap) MemMap = nullptr;
  decltype(&cuMemUnmap) MemUnmap = nullptr;
  decltype(&cuMemSetAccess) MemSetAccess = nullptr;
  decltype(&cuMulticastCreate) MulticastCreate = nullptr;
  decltype(&cuMulticastAddDevice) MulticastAddDevice = nullptr;
  decltype(&cuMulticastBindMem) MulticastBindMem = nullptr;
  decltype(&cuMulticastUnbind) MulticastUnbind = nullptr;
  decltype(&cuMulticastGetGranularity) MulticastGetGranularity = nullptr;
  decltype(&cuTensorMapEncodeTiled) TensorMapEncodeTiled = nullptr;

  static Driver& get() {
    static Driver d = load();
    return d;
  }

  std::string err(CUresult r) const {
    const char* s = nullptr;
    if (GetErrorString && GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUresult " + std::to_string((int)r);
  }

 private:
  template <typename F>
  static void resolve(F& fn, const char* name, bool required = true) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
      (void)cudaGetLastError();
      if (required) throw std::runtime_error(std::string("driver entry point not found: ") + name);
      p = nullptr;
    }
    fn = reinterpret_cast<F>(p);
  }
  static Driver load() {
    Driver d;
    resolve(d.GetErrorString, "cuGetErrorString");
    resolve(d.DeviceGet, "cuDeviceGet");
    resolve(d.DeviceGetAttribute, "cuDeviceGetAttribute");
    resolve(d.MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
    resolve(d.MemCreate, "cuMemCreate");
    resolve(d.MemRelease, "cuMemRelease");
    resolve(d.MemExportToShareableHandle, "cuMemExportToShareableHandle");
    resolve(d.MemImportFromShareableHandle, "cuMemImportFromShareableHandle");
    resolve(d.MemAddressReserve, "cuMemAddressReserve");
    resolve(d.MemAddressFree, "cuMemAddressFree");
    resolve(d.MemMap, "cuMemMap");
    resolve(d.MemUnmap, "cuMemUnmap");
    resolve(d.MemSetAccess, "cuMemSetAccess");
    resolve(d.MulticastCreate, "cuMulticastCreate", false);
    resolve(d.MulticastAddDevice, "cuMulticastAddDevice", false);
    resolve(d.MulticastBindMem, "cuMulticastBindMem", false);
    resolve(d.MulticastUnbind, "cuMulticastUnbind", false);
    resolve(d.MulticastGetGranularity, "cuMulticastGetGranularity", false);
    resolve(d.TensorMapEncodeTiled, "cuTensorMapEncodeTiled", false);
    return d;
  }
};

#define B200_DRV_CHECK(expr)                                                                        \
  do {                                                                                              \
    CUresult _r = (expr);                                                                           \
    if (_r != CUDA_SUCCESS)                                                                         \
      throw std::runtime_error(std::string("driver error: ") + ::b200::Driver::get().err(_r) +      \
                               " at " + __FILE__ + ":" + std::to_string(__LINE__) + " in " #expr);  \
  } while (0)

}  // namespace b200
