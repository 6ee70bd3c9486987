// Definitions shared by the single-CTA and CTA-pair variants of the fused token kernel (head_sm100.cu,
// head_sm100_pair.cu): tile constants, packed-weight layout, kernel parameters, bounded waits, epilogue helpers.
#pragma once
#include "common.cuh"
#include "sm100_prims.cuh"

namespace pasn {
namespace k1 {
using namespace sm100;

constexpr int TILE_M = 128;
constexpr int DD = 256;            // prototype depth D handled by this kernel
constexpr int DH = DD / 2;         // occurrence hidden width
constexpr int PP_MAX = 48;         // padded prototype count limit (multiple of 8)
constexpr int K1_READY_TARGET = 9;  // 8 epilogue warps + the occurrence warp
constexpr uint32_t FE_TILE_BYTES = 131072;  // K2 operand images of one 128-row tile: hi 64 KB | lo 64 KB

// packed weight buffer (bf16 stage images + fp32 biases), see pack_weights_kernel
struct PackedLayout {
  size_t off_l1, off_w4, off_w5, off_bias, off_w2, off_b2, total;
};
__host__ __device__ inline PackedLayout packed_layout(int C) {
  PackedLayout L;
  const int nkc = C / 64;
  L.off_l1 = 0;
  L.off_w4 = (size_t)2 * nkc * 32768;
  L.off_w5 = L.off_w4 + 65536;
  L.off_bias = L.off_w5 + 16384;
  L.off_w2 = (L.off_bias + (DD + DD + DH) * 4 + 1023) / 1024 * 1024;
  L.off_b2 = L.off_w2 + 4 * 32768;
  L.total = L.off_b2 + DD * 4;
  return L;
}

struct K1Params {
  const __nv_bfloat16* feat;   // [N][C][S] (bf16 input)
  const float* feat32;         // same buffer seen as fp32 when f32_in (bf16-compute mode for fp32 feature maps)
  float* occ32;                // occurrence map as fp32 when f32_in
  int f32_in;
  int nsc;                     // feature map is [N][S][C] (channels_last): two-phase kernel file only, bf16 only
  const uint8_t* packed;
  __nv_bfloat16* occ;          // [N][P][S] or null
  uint8_t* feimg;              // K2 operand images, FE_TILE_BYTES per K2 tile
  float* osum;                 // [N][P]
  int N, C, P, S, nkc, clips_per_cta, cpt;  // cpt = clips per K2 tile = 128 / PP
  // small batches: every clip is cut into G voxel groups of S voxels that are processed as G independent "clips"
  // (pooling is linear in the voxels; the partial sums are added after K2).  N and S above are the virtual counts,
  // SR the real voxel count per clip (= row pitch of the feature map and of the occurrence map).  G = 1: no split.
  int G, SR;
  int* err;
  int* fault;                  // host-mapped sticky fault word (or null)
  long long* trace;            // optional [3][16][16] clock64 stamps of CTA 0 (MMA thread, epilogue warp 4)
  int phases;                  // two-phase kernel: 2 = G / A phases overlapped with the previous chain, 1 = serial order
  int dbg_skip;                // timing experiments only: bit0 = do not copy weight stages, bit1 = do not gather X,
                               // bit3 = do not flush finished clips (head_sm100_k1.cu only)
  int flush_kmajor;            // 1: finished clips leave as 2-byte stores into K-major SWIZZLE_128B operand images (one 128-byte
                               // line per warp store); 0: 16-byte chunks of MN-major images (32 lines per warp store)
  int* ready;                  // [N] per-clip hand-off counter to the prototype kernel: +1 per epilogue warp that has stored
                               // its part of the clip's pooled vectors, +1 for the clip's Osum row (K1_READY_TARGET in total)
  int l2_hints;                // 1: pooled-vector images are stored evict_last, the feature map is read evict_first
  int flush_sleep;             // ns slept after every fourth prototype row of a clip flush (0: none)
  int x_drain;                 // 1: feature-gather warps publish everything in flight before they block on a full ring
  int spin;                    // Ctx::spin
};

// 1: build the timing experiments of profiles/README.md into the fused kernels (PASN_DBG_SKIP, PASN_X_DRAIN, PASN_FLUSH_SLEEP,
// PASN_FLUSH_KMAJOR, PASN_L2_HINTS, PASN_K2_EARLY); 0 (the product build): those switches are constants and their code is gone
#ifndef PASN_K1_EXPERIMENTS
#define PASN_K1_EXPERIMENTS 0
#endif

struct Ctx {
  int* err;
  volatile int* abort_s;
  int* fault;   // host-mapped sticky fault word (may be null)
  int spin;     // bit per waiter class (wait code / 100: 1 issuers, 2-3 producers, 4 occurrence warp, 5 epilogue, 6 K2):
                // poll with test_wait (never suspends) instead of try_wait (may suspend the thread for a while)
};

// bounded wait: returns false (and raises the CTA-wide abort flag) instead of hanging on a protocol bug.  The spin is
// out of line so that the fast path at every call site is a single try_wait.
static __device__ __noinline__ bool bwait_slow(uint64_t* bar, uint32_t parity, int* err, volatile int* abort_s, int* fault, int code,
                                               int spin) {
  const long long t0 = clock64();
  const bool poll = (spin >> (code / 100)) & 1;
  while (!(poll ? mbar_test_wait(bar, parity) : mbar_try_wait(bar, parity))) {
    if (*abort_s) return false;
    if (clock64() - t0 > 4000000000ll) {
      *abort_s = 1;
      atomicCAS(err, 0, code);
      if (fault != nullptr) *reinterpret_cast<volatile int*>(fault) = code;
      return false;
    }
  }
  return true;
}
__device__ __forceinline__ bool bwait(uint64_t* bar, uint32_t parity, const Ctx& c, int code) {
  if (mbar_try_wait(bar, parity)) return true;
  return bwait_slow(bar, parity, c.err, c.abort_s, c.fault, code, c.spin);
}

__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 32 fp32 accumulator columns + bias -> ReLU -> 16 packed bf16x2 (bias: 32 floats in smem, 16-byte aligned)
__device__ __forceinline__ void bias_relu_pack(const uint32_t (&r)[32], const float* bias, uint32_t* pk) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * j4);
    pk[2 * j4] = pack_bf16x2_relu(__uint_as_float(r[4 * j4 + 0]) + b.x, __uint_as_float(r[4 * j4 + 1]) + b.y);
    pk[2 * j4 + 1] = pack_bf16x2_relu(__uint_as_float(r[4 * j4 + 2]) + b.z, __uint_as_float(r[4 * j4 + 3]) + b.w);
  }
}

#define K1_TRACE(role, tile, slot)                                                            \
  do {                                                                                        \
    if (p.trace != nullptr && blockIdx.x == 0 && (tile) < 16)                                 \
      p.trace[((role) * 16 + (tile)) * 16 + (slot)] = clock64();                              \
  } while (0)


}  // namespace k1

// two-phase token kernel (head_sm100_k1.cu)
int launch_k1_two_phase(const k1::K1Params& k1p, int ppad, int grid, cudaStream_t st);

}  // namespace pasn
