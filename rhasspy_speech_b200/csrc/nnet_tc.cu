// Stage (ii) on the 5th-generation tensor cores: host side of the affine-layer kernels of the TDNN(-F) forward
// (TdnnComponent::Propagate, kaldi/src/nnet3/nnet-tdnn-component.cc:181-211, and the Affine / Linear components of
// nnet3/nnet-simple-component.cc, which the reference runs as cblas_sgemm, kaldi/src/matrix/kaldi-matrix.cc:171-183):
// tensor maps, weight packing, tile configuration and the dispatch to the device kernels
//   nnet_tc3.cu  gemm_tc3_kernel  (default)  TMA warp, MMA warp, sixteen epilogue warps of 32 columns each
//   nnet_tc2.cu  gemm_tc2_kernel  (RS_B200_TC=v2)  TMA warp, MMA warp, four fold warps, eight tail warps
// One persistent, warp-specialised kernel launch per layer; the activation tensor map of a time-offset slab is a
// row-shifted / row-strided VIEW of the producing layer's buffer (the TDNN splice is never materialised).
//
// Numerics.  The reference computes in fp32; the tolerance on the log-likelihoods is 1e-4.  One
// fp16 (or TF32) product carries an 11-bit significand, ~1e-3.  So every operand is carried as two
// fp16 planes (split.cuh)
//   x ~ hi + lo / 2048,   hi = fp16(x),   lo = fp16((x - hi) * 2048)          (22+ significant bits)
// and a product is three MMAs:  hi*hi  into a "main" accumulator and  hi*lo + lo*hi  into a "cross"
// accumulator that is folded with the exact factor 2^-11.  The dropped lo*lo term and the rounding
// of lo are O(2^-22) relative per product and unbiased.  The scaling keeps lo in fp16's normal range
// whatever the magnitude of x (an unscaled remainder of a weight of 0.03 would be subnormal).
// fp16 rather than TF32 planes: half the operand bytes per product and twice the MMA rate.  Activations are stored
// by the producing epilogue already split, weights are split once at model load; values beyond +-65504 saturate and
// raise a flag that fails the call (the fp32 CUDA-core path, RS_B200_GEMM=simt, has no such limit).
//
// Accumulation.  The tensor core adds each MMA into the fp32 TMEM accumulator with truncation, not round-to-nearest
// (measured with whole-K accumulation in TMEM: the error of a K = 2048 dot product grows linearly with K and is
// biased towards zero, ~1e-5 relative, which breaks the 1e-4 gate after 30 layers).  So a main accumulator only ever
// holds TWO MMAs (K = 32): the issuer starts a fresh one, hands it to the epilogue warps, and they add it into fp32
// registers with round-to-nearest -- the blocked summation a CPU sgemm micro-kernel performs.  Measured on the bench
// model (scripts/debug_ll.py, log-likelihoods against an fp64 forward; the reference's own nnet3-compute is 8e-6 rms /
// 7e-5 max away from it): four MMAs per hand-over 1.5e-5 rms / 1.4e-4 max (the truncation is biased towards zero and
// the bias adds up coherently over the layers), two 9e-6 / 7e-5, one 7e-6 / 5e-5; four on the K = 256 layers only is
// already 1.3e-5 / 1.2e-4.  The cross terms are 2^-11 of the result, so their accumulator stays in TMEM for the whole
// tile and is folded once.  TMEM holds a ring of two main accumulators and two cross accumulators (4 x bn columns).
#include <cuda.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "engine.h"
#include "model.h"
#include "nnet_tc.h"
#include "split.cuh"
#include "tc_ptx.cuh"

namespace rs {

// ------------------------------------------------------------------------------------ host side
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn GetEncodeTiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (!fn) RS_FAIL("cuTensorMapEncodeTiled is not available from the CUDA driver");
  return fn;
}

}  // namespace

// 2-D fp16 tensor [rows x cols], row pitch `pitch_elems`, box = 64 columns x box_rows, 128-byte
// swizzle, out-of-bounds elements read as zero.
void TcEncodeMap(CUtensorMap *map, const __half *base, long long rows, int cols, long long pitch_elems, int box_rows) {
  if (rows < 1) rows = 1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_elems * 2) % 16) RS_FAIL("tensor map operand is not 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = GetEncodeTiled()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS)
    RS_FAIL("cuTensorMapEncodeTiled failed (" << (int)rc << ") rows " << rows << " cols " << cols << " pitch " << pitch_elems);
}

bool TcSplitHost(float x, __half *hi, __half *lo) {
  bool ok = true;
  if (std::fabs(x) > 65504.f) {
    x = std::copysign(65504.f, x);
    ok = false;
  }
  *hi = __float2half_rn(x);
  *lo = __float2half_rn((x - __half2float(*hi)) * kSplitScale);
  return ok;
}

int TcTileN(int n) {
  const int tiles = (n + kTcMaxBN - 1) / kTcMaxBN;  // fewest tiles, then the smallest tile that covers n
  const int per = (n + tiles - 1) / tiles;
  return (per + 31) / 32 * 32;
}

void TcPackWeights(const float *w, int n, int ktot, const std::vector<std::pair<int, int>> &slabs, std::vector<__half> *hi,
                   std::vector<__half> *lo, std::vector<int> *k0, int *kp) {
  int total = 0;
  k0->clear();
  for (const auto &s : slabs) {
    k0->push_back(total);
    total += (s.second + kTcBK - 1) / kTcBK * kTcBK;
  }
  *kp = total;
  hi->assign((size_t)n * total, __float2half_rn(0.f));
  lo->assign((size_t)n * total, __float2half_rn(0.f));
  for (size_t si = 0; si < slabs.size(); si++) {
    const int wcol = slabs[si].first, k = slabs[si].second;
    if (wcol + k > ktot) RS_FAIL("weight slab out of range");
    for (int r = 0; r < n; r++)
      for (int c = 0; c < k; c++)
        if (!TcSplitHost(w[(size_t)r * ktot + wcol + c], &(*hi)[(size_t)r * total + (*k0)[si] + c], &(*lo)[(size_t)r * total + (*k0)[si] + c]))
          RS_FAIL("a weight exceeds the fp16 range of the tensor-core path (set RS_B200_GEMM=simt)");
  }
}

static int g_tc_smem_limit = 0;

void TcConfigure(TcParams *p) {
  if (g_tc_smem_limit == 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_tc_smem_limit = v;
  }
  const int stage_bytes = 2 * kTcBM * 128 + 2 * p->bn * 128;
  int stages = (g_tc_smem_limit - 1024 - 8 * 4096 - 256) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) RS_FAIL("not enough shared memory for the tensor-core GEMM pipeline");
  p->stages = stages;
  int cols = 32;
  while (cols < 4 * p->bn) cols <<= 1;
  p->tmem_cols = cols;
  static const int prof = getenv("RS_B200_TC_PROFILE") ? atoi(getenv("RS_B200_TC_PROFILE")) : 0;
  p->profile = prof;
  p->fold = 2;
  p->tiles_m = (p->m + kTcBM - 1) / kTcBM;
  p->tiles_n = (p->n + p->bn - 1) / p->bn;
  p->total_kb = 0;
  for (int s = 0; s < p->n_slabs; s++) p->total_kb += p->slabs[s].kblocks;
}

void LaunchGemmTc(const TcParams &p, int num_sms, cudaStream_t stream) {
  if (p.m <= 0 || p.n <= 0) return;
  // RS_B200_TC=v2 keeps the arrangement with separate fold and tail warps selectable for A/B measurements
  static const bool v2 = getenv("RS_B200_TC") && !strcmp(getenv("RS_B200_TC"), "v2");
  if (v2)
    LaunchGemmTc2(p, num_sms, g_tc_smem_limit, stream);
  else
    LaunchGemmTc3(p, num_sms, g_tc_smem_limit, stream);
}

}  // namespace rs
