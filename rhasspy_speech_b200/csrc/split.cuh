// The split activation format read by the tensor-core layers (nnet_tc.cu): a buffer is two fp16
// planes, value = hi + lo / 2048 with hi = fp16(x), lo = fp16((x - hi) * 2048).  The sum is exactly
// representable in fp32 (two non-overlapping 11-bit significands), so every consumer -- tensor core
// or CUDA core -- sees the same number, 2^-23-relative close to the fp32 value that was stored.
#pragma once
#include <cuda_fp16.h>

namespace rs {

// reads element i of a buffer that is either plain fp32 (lo == nullptr) or split
__device__ __forceinline__ float ld_act(const void *hi, const void *lo, size_t i) {
  if (lo)
    return fmaf(__half2float(reinterpret_cast<const __half *>(lo)[i]), 1.f / 2048.f,
                __half2float(reinterpret_cast<const __half *>(hi)[i]));  // exact
  return reinterpret_cast<const float *>(hi)[i];
}

// returns true if x had to be clamped into the fp16 range
__device__ __forceinline__ bool split_f16(float x, __half &hi, __half &lo) {
  bool sat = false;
  if (fabsf(x) > 65504.f) {
    x = copysignf(65504.f, x);
    sat = true;
  }
  hi = __float2half_rn(x);
  lo = __float2half_rn(__fmul_rn(__fsub_rn(x, __half2float(hi)), 2048.f));
  return sat;
}

// writes element i of a plain (lo == nullptr) or split buffer
__device__ __forceinline__ void st_act(void *hi, void *lo, size_t i, float x, int *range_flag) {
  if (lo) {
    __half h, l;
    if (split_f16(x, h, l) && range_flag) *range_flag = 1;
    reinterpret_cast<__half *>(hi)[i] = h;
    reinterpret_cast<__half *>(lo)[i] = l;
  } else {
    reinterpret_cast<float *>(hi)[i] = x;
  }
}

}  // namespace rs
