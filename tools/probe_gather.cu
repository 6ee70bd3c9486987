// Gather-rate probe: how fast can the X producers of K1 move an NCDHW bf16 feature map into the MN-major SWIZZLE_128B
// operand layout when nothing else runs?  Same addressing as head_sm100.cu (tiles of 128 voxels, 64-channel chunks,
// lane = 4 voxels, one warp-wide 8-byte load = a 256-byte run of one channel row), 224 KB of smem carved out like K1.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_gather.bin tools/probe_gather.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>

#include "../protoasnet_b200/csrc/sm100_prims.cuh"

using namespace pasn::sm100;

constexpr int S = 196, C = 512, NKC = C / 64, TILE_M = 128;

// DEPTH units of ROWS channel rows in flight per warp (registers); NW producer warps; each chunk = 64 rows split
// over the warps; REPEAT = how many times each chunk is gathered (2 = the two-phase K1 re-reads it from L2)
template <int NW, int ROWS, int DEPTH, int ALLOC_L1, int REPEAT>
__global__ void __launch_bounds__(NW * 32) gather_kernel(const __nv_bfloat16* __restrict__ feat, int N, int clips_per_cta,
                                                          long long* cycles, unsigned* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c_begin = blockIdx.x * clips_per_cta;
  int ncl = min(N - c_begin, clips_per_cta);
  if (ncl <= 0) return;
  const int ntok = ncl * S, ntiles = (ntok + TILE_M - 1) / TILE_M;
  constexpr int ROWS_PER_WARP = 64 / NW;               // per chunk
  constexpr int UNITS_PER_CHUNK = ROWS_PER_WARP / ROWS;
  static_assert(UNITS_PER_CHUNK >= 1, "unit larger than a warp's share of a chunk");
  const uint32_t x_base = smem_u32(smem);
  const int nunits = ntiles * NKC * REPEAT * UNITS_PER_CHUNK;
  uint2 v[DEPTH][ROWS];
  auto load = [&](int u, uint2* dst) {
    const int job = u / UNITS_PER_CHUNK, part = u - job * UNITS_PER_CHUNK;
    const int tile = job / (NKC * REPEAT), kc = (job % (NKC * REPEAT)) % NKC;
    const int t = tile * TILE_M + 4 * lane;
    const bool valid = t < ntok;
    const int clipl = valid ? t / S : 0, s = valid ? t - clipl * S : 0;
    const __nv_bfloat16* src = feat + ((size_t)(c_begin + clipl) * C + kc * 64 + warp * ROWS_PER_WARP + part * ROWS) * S + s;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      if (!valid) dst[j] = make_uint2(0, 0);
      else if (ALLOC_L1) dst[j] = *reinterpret_cast<const uint2*>(src + (size_t)j * S);
      else dst[j] = ldg_nc_na_v2(src + (size_t)j * S);
    }
  };
  auto store = [&](int u, const uint2* srcv) {
    const int job = u / UNITS_PER_CHUNK, part = u - job * UNITS_PER_CHUNK;
    const uint32_t dst0 = x_base + (job & 3) * 16384;
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
      st_shared_v2(dst0 + off_mnmajor_sw128(4 * lane, warp * ROWS_PER_WARP + part * ROWS + j, 8192), srcv[j]);
    if (part == UNITS_PER_CHUNK - 1) fence_proxy_async();
  };
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll
  for (int d = 0; d < DEPTH - 1; ++d)
    if (d < nunits) load(d, v[d]);
  for (int u0 = 0; u0 < nunits; u0 += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      const int u = u0 + d;
      if (u + DEPTH - 1 < nunits) load(u + DEPTH - 1, v[(d + DEPTH - 1) % DEPTH]);
      if (u < nunits) store(u, v[d]);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (sink != nullptr && smem[threadIdx.x * 16] == 0x5a) atomicAdd(sink, 1u);
}


// cp.async (LDGSTS) variant: no registers, DEPTH commit groups in flight per warp.  MODE 0: 8-byte .ca copies for every
// row; MODE 1: rows whose global address is 16-byte aligned use 16-byte .cg copies (two rows per warp instruction).
template <int NW, int DEPTH, int MODE, int REPEAT = 1, int FENCE_EVERY = 4>
__global__ void __launch_bounds__(NW * 32) gather_cp_kernel(const __nv_bfloat16* __restrict__ feat, int N, int clips_per_cta,
                                                             long long* cycles, unsigned* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c_begin = blockIdx.x * clips_per_cta;
  int ncl = min(N - c_begin, clips_per_cta);
  if (ncl <= 0) return;
  const int ntok = ncl * S, ntiles = (ntok + TILE_M - 1) / TILE_M;
  constexpr int RPW = 64 / NW;
  const uint32_t x_base = smem_u32(smem);
  const int njobs = ntiles * NKC * REPEAT;
  __syncthreads();
  const long long t0 = clock64();
  for (int job = 0; job < njobs; ++job) {
    const int tile = job / (NKC * REPEAT), kc = (job % (NKC * REPEAT)) % NKC;
    const uint32_t dst0 = x_base + (job & 3) * 16384;
    if (MODE == 0) {
      const int t = tile * TILE_M + 4 * lane;
      const bool valid = t < ntok;
      const int clipl = valid ? t / S : 0, s = valid ? t - clipl * S : 0;
      const __nv_bfloat16* src = feat + ((size_t)(c_begin + clipl) * C + kc * 64 + warp * RPW) * S + s;
#pragma unroll
      for (int j = 0; j < RPW; ++j)
        cp_async_8(dst0 + off_mnmajor_sw128(4 * lane, warp * RPW + j, 8192), src + (size_t)j * S, valid ? 8u : 0u);
    } else {
      // rows alternate between 16-byte aligned and 8-byte shifted (392-byte pitch); the parity of the first row of this
      // warp's share decides which rows (even or odd j) are the aligned ones for this lane's voxel offset
#pragma unroll
      for (int jp = 0; jp < RPW / 2; ++jp) {
        // pair of rows (2jp, 2jp+1): one aligned, one shifted
        const int half = lane >> 4, l16 = lane & 15;           // aligned row: 16 lanes x 16 B, two row-halves... one row per half-warp
        (void)half; (void)l16;
        const int t8 = tile * TILE_M + 8 * (lane & 15);
        const bool v8 = t8 < ntok;
        const int c8 = v8 ? t8 / S : 0, s8 = v8 ? t8 - c8 * S : 0;
        const int t4 = tile * TILE_M + 4 * lane;
        const bool v4 = t4 < ntok;
        const int c4 = v4 ? t4 / S : 0, s4 = v4 ? t4 - c4 * S : 0;
        const int row0 = kc * 64 + warp * RPW + 2 * jp;
        // which of the two rows is 16-byte aligned at voxel offset s8 for clip c8 (pitch 392 B = 8 mod 16)
        const size_t e0 = ((size_t)(c_begin + c8) * C + row0) * S + s8;     // element index of row0
        const int r_al = ((e0 * 2) & 15) == 0 ? 0 : 1;
        // aligned row: lanes 0..15 copy it with 16-byte chunks (a clip boundary inside the tile keeps 8-voxel chunks intact only
        // if S % 8 == 0; S = 196 -> fall back to 8-byte copies for chunks that straddle)
        const bool straddle = v8 && (s8 + 8 > S);
        if (lane < 16) {
          const __nv_bfloat16* src = feat + ((size_t)(c_begin + c8) * C + row0 + r_al) * S + s8;
          if (!straddle && ((((size_t)src) & 15) == 0))
            cp_async_16(dst0 + off_mnmajor_sw128(8 * lane, warp * RPW + 2 * jp + r_al, 8192), src, v8 ? 16u : 0u);
          else {
            cp_async_8(dst0 + off_mnmajor_sw128(8 * lane, warp * RPW + 2 * jp + r_al, 8192), src, v8 ? 8u : 0u);
            const int t8b = t8 + 4; const bool vb = t8b < ntok; const int cb = vb ? t8b / S : 0, sb = vb ? t8b - cb * S : 0;
            cp_async_8(dst0 + off_mnmajor_sw128(8 * lane + 4, warp * RPW + 2 * jp + r_al, 8192),
                       feat + ((size_t)(c_begin + cb) * C + row0 + r_al) * S + sb, vb ? 8u : 0u);
          }
        }
        {
          const __nv_bfloat16* src = feat + ((size_t)(c_begin + c4) * C + row0 + (1 - r_al)) * S + s4;
          cp_async_8(dst0 + off_mnmajor_sw128(4 * lane, warp * RPW + 2 * jp + (1 - r_al), 8192), src, v4 ? 8u : 0u);
        }
      }
    }
    cp_async_commit();
    cp_async_wait<DEPTH - 1>();
    if (FENCE_EVERY > 0 && ((job + 1) % FENCE_EVERY) == 0) fence_proxy_async();
  }
  cp_async_wait<0>();
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (sink != nullptr && smem[threadIdx.x * 16] == 0x5a) atomicAdd(sink, 1u);
}

template <int NW, int DEPTH, int MODE, int REPEAT = 1, int FENCE_EVERY = 4>
static void run_cp(const char* name, const __nv_bfloat16* feat, int N, long long* d_cyc, unsigned* d_sink, int smem_kb) {
  auto k = gather_cp_kernel<NW, DEPTH, MODE, REPEAT, FENCE_EVERY>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  const int cpc = (N + 147) / 148, grid = (N + cpc - 1) / cpc;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) k<<<grid, NW * 32, smem_kb * 1024>>>(feat, N, cpc, d_cyc, d_sink);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) k<<<grid, NW * 32, smem_kb * 1024>>>(feat, N, cpc, d_cyc, d_sink);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double us = ms * 1e3 / 5, bytes = (double)N * C * S * 2 * REPEAT;
  const int tiles = (cpc * S + 127) / 128;
  printf("%-58s smem %3d KB: %7.1f us  %6.0f GB/s  CTA0 %.0f cyc/chunk (%.1f B/clk/SM)  %s\n", name, smem_kb, us,
         bytes / us * 1e-3, (double)h[0] / (tiles * NKC * REPEAT), 16384.0 * tiles * NKC * REPEAT / h[0], cudaGetErrorString(e));
  if (e != cudaSuccess) exit(2);
}


// MODE 2: clean mixed gather.  Per channel row and clip segment, rows whose source is 16-byte aligned at even piece
// indices use one 16-byte cp.async.cg per lane pair (issued by the even lane, bypasses L1); everything else 8-byte .ca.
template <int NW, int DEPTH>
__global__ void __launch_bounds__(NW * 32) gather_mixed_kernel(const __nv_bfloat16* __restrict__ feat, int N, int clips_per_cta,
                                                                long long* cycles, unsigned* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c_begin = blockIdx.x * clips_per_cta;
  int ncl = min(N - c_begin, clips_per_cta);
  if (ncl <= 0) return;
  const int ntok = ncl * S, ntiles = (ntok + TILE_M - 1) / TILE_M;
  constexpr int RPW = 64 / NW;
  const uint32_t x_base = smem_u32(smem);
  const int njobs = ntiles * NKC;
  __syncthreads();
  const long long t0 = clock64();
  for (int job = 0; job < njobs; ++job) {
    const int tile = job / NKC, kc = job % NKC;
    const uint32_t dst0 = x_base + (job & 3) * 16384;
    const int t = tile * TILE_M + 4 * lane;
    const bool valid = t < ntok;
    const int clipl = valid ? t / S : 0, s = valid ? t - clipl * S : 0;
    // partner piece (lane ^ 1) in the same clip and valid?  (pieces are 4 voxels; S % 4 == 0)
    const int tp = tile * TILE_M + 4 * (lane | 1);
    const bool pair_ok = valid && tp < ntok && (tp / S) == clipl && ((lane & 1) == 0 ? true : true);
    const __nv_bfloat16* src = feat + ((size_t)(c_begin + clipl) * C + kc * 64 + warp * RPW) * S + s;
    // alignment of this lane's piece for row j: ((rowg + s/4) & 1) == 0 -> 16-byte aligned
    const int par0 = (((c_begin + clipl) * C + kc * 64 + warp * RPW) + (s >> 2)) & 1;
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
      const bool piece_aligned = ((par0 + j) & 1) == 0;   // S odd multiple of 4: alternates with the row index
      const uint32_t dst = dst0 + off_mnmajor_sw128(4 * lane, warp * RPW + j, 8192);
      // even lane with an aligned piece and a good partner: one 16-byte copy for both pieces; its partner issues nothing
      const bool lead16 = (lane & 1) == 0 && piece_aligned && pair_ok;
      const bool covered = (lane & 1) == 1 && !piece_aligned && pair_ok;   // partner (even lane) is aligned and copies for us
      if (lead16) cp_async_16(dst, src + (size_t)j * S, 16u);
      else if (!covered) cp_async_8(dst, src + (size_t)j * S, valid ? 8u : 0u);
    }
    cp_async_commit();
    cp_async_wait<DEPTH - 1>();
    if (((job + 1) & 3) == 0) fence_proxy_async();
  }
  cp_async_wait<0>();
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (sink != nullptr && smem[threadIdx.x * 16] == 0x5a) atomicAdd(sink, 1u);
}

template <int NW, int DEPTH>
static void run_mixed(const char* name, const __nv_bfloat16* feat, int N, long long* d_cyc, unsigned* d_sink, int smem_kb) {
  auto k = gather_mixed_kernel<NW, DEPTH>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  const int cpc = (N + 147) / 148, grid = (N + cpc - 1) / cpc;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) k<<<grid, NW * 32, smem_kb * 1024>>>(feat, N, cpc, d_cyc, d_sink);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) k<<<grid, NW * 32, smem_kb * 1024>>>(feat, N, cpc, d_cyc, d_sink);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double us = ms * 1e3 / 5, bytes = (double)N * C * S * 2;
  const int tiles = (cpc * S + 127) / 128;
  printf("%-58s smem %3d KB: %7.1f us  %6.0f GB/s  CTA0 %.0f cyc/chunk (%.1f B/clk/SM)  %s\n", name, smem_kb, us,
         bytes / us * 1e-3, (double)h[0] / (tiles * NKC), 16384.0 * tiles * NKC / h[0], cudaGetErrorString(e));
  if (e != cudaSuccess) exit(2);
}

template <int NW, int ROWS, int DEPTH, int ALLOC_L1, int REPEAT>
static void run(const char* name, const __nv_bfloat16* feat, int N, long long* d_cyc, unsigned* d_sink, int smem_kb) {
  auto k = gather_kernel<NW, ROWS, DEPTH, ALLOC_L1, REPEAT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  const int cpc = (N + 147) / 148, grid = (N + cpc - 1) / cpc;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) k<<<grid, NW * 32, smem_kb * 1024>>>(feat, N, cpc, d_cyc, d_sink);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) k<<<grid, NW * 32, smem_kb * 1024>>>(feat, N, cpc, d_cyc, d_sink);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double us = ms * 1e3 / 5, bytes = (double)N * C * S * 2 * REPEAT;
  const int tiles = (cpc * S + 127) / 128;
  printf("%-58s smem %3d KB: %7.1f us  %6.0f GB/s  CTA0 %.0f cyc/chunk (%.1f B/clk/SM)  %s\n", name, smem_kb, us,
         bytes / us * 1e-3, (double)h[0] / (tiles * NKC * REPEAT), 16384.0 * tiles * NKC * REPEAT / h[0],
         cudaGetErrorString(e));
  if (e != cudaSuccess) exit(2);
}

int main() {
  const int N = 1024;
  __nv_bfloat16* feat;
  long long* d_cyc; unsigned* d_sink;
  cudaMalloc(&feat, (size_t)N * C * S * 2);
  cudaMemset(feat, 0, (size_t)N * C * S * 2);
  cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_sink, 4);
  cudaMemset(d_sink, 0, 4);
  run_cp<4, 3, 0, 1, 0>("cp.async.ca 8 B, 3 groups, no proxy fence", feat, N, d_cyc, d_sink, 187);
  run_cp<4, 3, 0, 1, 4>("cp.async.ca 8 B, 3 groups, proxy fence every 4 chunks", feat, N, d_cyc, d_sink, 187);
  run_cp<4, 3, 0, 1, 1>("cp.async.ca 8 B, 3 groups, proxy fence every chunk", feat, N, d_cyc, d_sink, 187);
  run_cp<4, 3, 0, 2, 1>("same, every chunk twice", feat, N, d_cyc, d_sink, 187);
  run_cp<4, 3, 0, 2, 0>("no fence, every chunk twice", feat, N, d_cyc, d_sink, 187);
  return 0;
}
