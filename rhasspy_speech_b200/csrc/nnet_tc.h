// Parameter block and host helpers of the tcgen05 / TMA affine-layer kernel (nnet_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <utility>
#include <vector>

#include "engine.h"

namespace rs {

constexpr int kTcBM = 128;      // output rows (time steps) per tile = UMMA M
constexpr int kTcBK = 64;       // fp16 columns per pipeline stage = one 128-byte swizzle atom
constexpr int kTcMaxBN = 128;   // output columns per tile (fp32 running sums live in registers)
constexpr int kTcMaxSlabs = 4;
constexpr float kSplitScale = 2048.f;  // lo plane = (x - hi) * 2^11, see nnet_tc.cu

struct TcSlab {
  int kblocks;  // ceil(k / 64)
  int wk0;      // first column of this slab in the packed weight matrix (multiple of 64)
  int yshift;   // row of the slab's (shifted, strided) source view that output row 0 reads
};

struct TcParams {
  CUtensorMap a_hi[kTcMaxSlabs], a_lo[kTcMaxSlabs];  // activation planes, one view per slab
  CUtensorMap w_hi, w_lo;                            // packed weights [n x kp]
  TcSlab slabs[kTcMaxSlabs];
  int n_slabs;
  int bn;         // output columns per tile = UMMA N (multiple of 32, <= kTcMaxBN)
  int stages;     // smem pipeline depth
  int fold;       // main MMAs (K = 16 each) accumulated inside TMEM per published partial sum: 1, 2 or 4 (nnet_tc.cu)
  int profile;    // RS_B200_TC_PROFILE: block 0 prints where its TMA / MMA / epilogue threads spent their clocks
  int tmem_cols;  // power of two >= 4 * bn (two K-block accumulator pairs)
  int tiles_m, tiles_n;
  int total_kb;   // sum of the slabs' K blocks (TcConfigure): a launch constant the kernels read from the parameter bank
  void *out_hi, *out_lo;  // out_lo == nullptr: plain fp32 output at out_hi, else two fp16 planes
  int out_ld, m, n;
  DevOp ops[kMaxOps];
  int n_ops;
  const int *row_utt;
  int *range_flag;  // set to 1 if a split store had to saturate (|x| > 65504)
};

// tensor map over a row-major fp16 matrix (box 64 columns x box_rows, SWIZZLE_128B, zero fill)
void TcEncodeMap(CUtensorMap *map, const __half *base, long long rows, int cols, long long pitch_elems, int box_rows);
// x ~ hi + lo / 2048 with hi = fp16(x), lo = fp16((x - hi) * 2048); returns false if |x| is out of fp16 range
bool TcSplitHost(float x, __half *hi, __half *lo);
int TcTileN(int n);
// Re-lays W [n x ktot] as [n x kp] with every slab's columns padded to a multiple of 64 (so a
// K block never straddles two slabs) and splits it into the two fp16 planes.
void TcPackWeights(const float *w, int n, int ktot, const std::vector<std::pair<int, int>> &slabs /* (wcol, k) */,
                   std::vector<__half> *hi, std::vector<__half> *lo, std::vector<int> *k0, int *kp);
void TcConfigure(TcParams *p);  // fills stages / tmem_cols / tiles_* from bn, m, n
void LaunchGemmTc(const TcParams &p, int num_sms, cudaStream_t stream);
// second-generation kernel (nnet_tc2.cu): separate fold and tail warps
void LaunchGemmTc2(const TcParams &p, int num_sms, int smem_limit, cudaStream_t stream);
// third arrangement (nnet_tc3.cu): sixteen epilogue warps, 32 columns each, fold and tail in the same warp
void LaunchGemmTc3(const TcParams &p, int num_sms, int smem_limit, cudaStream_t stream);

}  // namespace rs
