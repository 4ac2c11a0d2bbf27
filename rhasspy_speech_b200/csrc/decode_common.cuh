// Small device helpers shared by the decoder kernels (decode.cu, decode_small.cu).
#pragma once
#include <cuda_runtime.h>

namespace rs {

constexpr unsigned long long kEmptyVal = ~0ULL;
constexpr unsigned kArcNone = 0xffffffffu;

// order-preserving map float -> unsigned (and back): a < b  <=>  ord(a) < ord(b)
__device__ __forceinline__ unsigned ord(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unord(unsigned o) {
  unsigned b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}
// (cost, tag) in one word: an atomicMin keeps the cheapest cost and, among equal costs, the smallest tag
__device__ __forceinline__ unsigned long long pack(float cost, unsigned tag) {
  return ((unsigned long long)ord(cost) << 32) | tag;
}

}  // namespace rs
