// Opt-in dynamic shared memory (> 48 KB) is a per-DEVICE attribute of a kernel: a process that drives several devices
// (one engine per device, transcribe.py) has to set it on each of them.  EnsureDynSmem remembers the largest size set
// per (kernel, device) and raises it when a launch needs more.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

namespace rs {

template <typename Kernel>
inline cudaError_t EnsureDynSmem(Kernel kernel, size_t bytes) {
  if (bytes == 0) return cudaSuccess;  // (also below 48 KB: static + dynamic shared memory together may need the opt-in)
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, size_t> configured;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  size_t &have = configured[std::make_pair(reinterpret_cast<const void *>(kernel), dev)];
  if (bytes <= have) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

}  // namespace rs
