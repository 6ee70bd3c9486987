// Issue-rate probe for the tcgen05 shapes a two-phase K1 would use (run on a B200 via gpurun).
// One CTA per SM on the whole chip (so clocks / power are realistic); thread 0 of each CTA issues ITER MMAs back to
// back on operands that were zero-filled, then commits and waits; cycles per MMA = (t1 - t0) / ITER on CTA 0.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_rate.bin tools/probe_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../protoasnet_b200/csrc/sm100_prims.cuh"

using namespace pasn::sm100;

constexpr int ITER = 1024;

// mode 0: SS M128 N256  A = MN-major SW128 (X tile), B = K-major SW128 (weights)        [current layer-1 pass]
// mode 1: SS M128 N128  A = K-major SW128 (weights), B = MN-major SW128 (X tile)         [transposed add-on pass]
// mode 2: TS M128 N96   A = TMEM bf16, B = MN-major no-swizzle (Os)                      [pooling from TMEM]
// mode 3: modes 0 and 1 issued by two different warps at the same time (N256 into [0,256), N128 into [256,384))
// mode 4: mode 1 while four other warps stream st.shared into a different 64 KB region (X gather stand-in)
// mode 5: TS M128 N128  A = TMEM, B = K-major SW128                                     [G2 pass]
__global__ void __launch_bounds__(256) rate_kernel(int mode, int random, long long* out, int* err, const unsigned char* wsrc,
                                                   const unsigned char* xsrc) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); stop = 0; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int i = tid; i < 160 * 1024 / 4; i += 256) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const float a = ((h & 0xFFFF) / 32768.0f - 1.0f), b = ((h >> 16) / 32768.0f - 1.0f);
    reinterpret_cast<uint32_t*>(base)[i] = random ? pack_bf16x2(a, b) : 0u;
  }
  if (random) {   // A operands of the TS modes live in TMEM columns [256,512)
    uint32_t r[16];
    for (int c = 0; c < 256; c += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint32_t h = (uint32_t)(tid * 512 + c + j) * 2654435761u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        r[j] = pack_bf16x2(((h & 0xFFFF) / 32768.0f - 1.0f), ((h >> 16) / 32768.0f - 1.0f));
      }
      if (warp < 4) tmem_st_x16(tmem_base_s + ((uint32_t)(warp * 32) << 16) + 256 + c, r);
    }
    tmem_st_wait();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + 65536);   // 64 KB of "X", 64 KB of "W"
  auto issue = [&](int m, uint64_t* done) -> long long {
    const long long t0 = clock64();
    if (m == 0) {
      const uint32_t id = make_idesc_bf16(128, 256, 1, 0);
      for (int i = 0; i < ITER; ++i) {
        const int k4 = i & 3, slot = (i >> 2) & 3, ws = (i >> 2) & 1;
        const uint64_t ad = make_smem_desc(a0 + slot * 16384 + k4 * 2048, 8192, 1024, SWZ_128B);
        const uint64_t bd = make_smem_desc(b0 + ws * 32768 + k4 * 32, 16, 1024, SWZ_128B);
        mma_ss(tbase, ad, bd, id, i ? 1u : 0u);
      }
    } else if (m == 1) {
      const uint32_t id = make_idesc_bf16(128, 128, 0, 1);
      for (int i = 0; i < ITER; ++i) {
        const int k4 = i & 3, h = (i >> 2) & 1, slot = (i >> 3) & 3, ws = (i >> 3) & 1;
        const uint64_t ad = make_smem_desc(b0 + ws * 32768 + h * 16384 + k4 * 32, 16, 1024, SWZ_128B);
        const uint64_t bd = make_smem_desc(a0 + slot * 16384 + k4 * 2048, 8192, 1024, SWZ_128B);
        mma_ss(tbase + 256u + 128u * h, ad, bd, id, i > 7 ? 1u : 0u);
      }
    } else if (m == 2) {
      const uint32_t id = make_idesc_bf16(128, 96, 0, 1);
      constexpr uint32_t lbo = (96 / 8) * 128;
      for (int i = 0; i < ITER; ++i) {
        const int ks = i & 7;
        const uint64_t bd = make_smem_desc(a0 + ks * 2 * lbo, lbo, 128, SWZ_NONE);
        mma_ts(tbase, tbase + 256u + 8u * ks, bd, id, i ? 1u : 0u);
      }
    } else {
      const uint32_t id = make_idesc_bf16(128, 128, 0, 0);
      for (int i = 0; i < ITER; ++i) {
        const int ks = i & 15;
        const uint64_t bd = make_smem_desc(b0 + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, SWZ_128B);
        mma_ts(tbase + 64u, tbase + 256u + 8u * ks, bd, id, i ? 1u : 0u);
      }
    }
    const long long t1 = clock64();
    mma_commit(done);
    mbar_wait(done, 0, err, 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[2 * m] = t1 - t0; out[2 * m + 1] = t2 - t0; }
    return t2 - t0;
  };
  if (mode == 3) {
    if (tid == 0) issue(0, &bar[0]);
    if (tid == 32) issue(1, &bar[1]);
  } else if (mode == 4) {
    if (tid == 0) { issue(1, &bar[0]); stop = 1; }
    if (warp >= 4) {   // st.shared stream into the unused upper 32 KB of the scratch area
      const uint32_t dst = smem_u32(base + 131072) + (uint32_t)(tid - 128) * 8;
      int n = 0;
      while (!stop && n < (1 << 20)) {
#pragma unroll
        for (int j = 0; j < 16; ++j) st_shared_v2(dst + j * 1024, make_uint2(n, j));
        fence_proxy_async();
        ++n;
      }
      if (blockIdx.x == 0 && tid == 128) out[14] = n;
    }
  } else if (mode >= 23 && mode <= 34) {
    // issue blocks (23-28: 8 x N128 SS, 29-34: 4 x N256 SS) + one commit, separated by a dependent ALU chain of
    // 0 / 16 / 32 / 64 / 128 / 256 multiply-adds in the issuing thread: how much preparation fits between two blocks?
    __shared__ uint64_t sink[8];
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&sink[i], 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
      const bool n128 = mode <= 28;
      const int gsel = n128 ? mode - 23 : mode - 29;
      const int gap = gsel == 0 ? 0 : (8 << gsel);
      const uint32_t ida = make_idesc_bf16(128, 128, 0, 1), idg = make_idesc_bf16(128, 256, 1, 0);
      uint32_t junk = (uint32_t)clock64();
      const long long t0 = clock64();
      for (int c = 0; c < ITER / 8; ++c) {
        const int slot = c & 3, ws = c & 1;
        if (n128) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k4 = j & 3, h = j >> 2;
            mma_ss_x(tbase + 256u + 128u * h, desc_lo(b0 + ws * 32768 + h * 16384 + k4 * 32, 16), desc_hi(1024, SWZ_128B),
                     desc_lo(a0 + slot * 16384 + k4 * 2048, 8192), desc_hi(1024, SWZ_128B), ida, c ? 1u : (uint32_t)(k4 != 0));
          }
        } else {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            mma_ss_x(tbase, desc_lo(a0 + slot * 16384 + k4 * 2048, 8192), desc_hi(1024, SWZ_128B),
                     desc_lo(b0 + ws * 32768 + k4 * 32, 16), desc_hi(1024, SWZ_128B), idg, c ? 1u : (uint32_t)(k4 != 0));
        }
        mma_commit(&sink[c & 7]);
        for (int g = 0; g < gap; ++g) junk = junk * 1664525u + 1013904223u;
      }
      const long long t1 = clock64();
      mma_commit(&bar[0]);
      mbar_wait(&bar[0], 0, err, 1);
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[2] = t1 - t0; out[3] = (t2 - t0) * (n128 ? 1 : 2); out[15] = junk; }
    }
  } else if (mode >= 15 && mode <= 22) {
    // commit interval sweep: 15-18 shape 0 (N256) with a commit every 2 / 4 / 8 / 16 MMAs; 19-22 shape 1 (N128) every 4 / 8 / 16 / 32
    __shared__ uint64_t sink[8];
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&sink[i], 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
      const bool n256 = mode <= 18;
      const int every = n256 ? (2 << (mode - 15)) : (4 << (mode - 19));
      const uint32_t id = n256 ? make_idesc_bf16(128, 256, 1, 0) : make_idesc_bf16(128, 128, 0, 1);
      const long long t0 = clock64();
      for (int i = 0; i < ITER; ++i) {
        const int k4 = i & 3, h = (i >> 2) & 1, slot = (i >> 3) & 3, ws = (i >> 3) & 1;
        if (n256) {
          const uint64_t ad = make_smem_desc(a0 + slot * 16384 + k4 * 2048, 8192, 1024, SWZ_128B);
          const uint64_t bd = make_smem_desc(b0 + ws * 32768 + k4 * 32, 16, 1024, SWZ_128B);
          mma_ss(tbase, ad, bd, id, i ? 1u : 0u);
        } else {
          const uint64_t ad = make_smem_desc(b0 + ws * 32768 + h * 16384 + k4 * 32, 16, 1024, SWZ_128B);
          const uint64_t bd = make_smem_desc(a0 + slot * 16384 + k4 * 2048, 8192, 1024, SWZ_128B);
          mma_ss(tbase + 256u + 128u * h, ad, bd, id, i > 7 ? 1u : 0u);
        }
        if ((i + 1) % every == 0) mma_commit(&sink[(i / every) & 7]);
      }
      const long long t1 = clock64();
      mma_commit(&bar[0]);
      mbar_wait(&bar[0], 0, err, 1);
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[2] = t1 - t0; out[3] = t2 - t0; }
    }
  } else if (mode >= 11 && mode <= 14) {
    // kernel-like issue loop for shape 1: per chunk of 8 MMAs, 11: one commit; 12: two commits; 13: two commits + two
    // waits on barriers that completed long ago; 14: like 13 plus tcgen05.fence::after_thread_sync after the waits
    __shared__ uint64_t donebar[4];
    __shared__ uint64_t sink[8];
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&donebar[i], 1); for (int i = 0; i < 8; ++i) mbar_init(&sink[i], 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
      for (int i = 0; i < 4; ++i) mbar_arrive(&donebar[i]);   // phase 0 complete: waits on parity 0 succeed at once
      const uint32_t id = make_idesc_bf16(128, 128, 0, 1);
      const long long t0 = clock64();
      for (int c = 0; c < ITER / 8; ++c) {
        if (mode >= 13) {
          mbar_wait(&donebar[c & 1], 0, err, 3);
          mbar_wait(&donebar[2 + (c & 1)], 0, err, 4);
          if (mode == 14) tc_fence_after();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k4 = j & 3, h = j >> 2, slot = c & 3, ws = c & 1;
          const uint64_t ad = make_smem_desc(b0 + ws * 32768 + h * 16384 + k4 * 32, 16, 1024, SWZ_128B);
          const uint64_t bd = make_smem_desc(a0 + slot * 16384 + k4 * 2048, 8192, 1024, SWZ_128B);
          mma_ss(tbase + 256u + 128u * h, ad, bd, id, c ? 1u : (uint32_t)(k4 != 0));
        }
        mma_commit(&sink[c & 3]);
        if (mode >= 12) mma_commit(&sink[4 + (c & 3)]);
      }
      const long long t1 = clock64();
      mma_commit(&bar[0]);
      mbar_wait(&bar[0], 0, err, 1);
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[2] = t1 - t0; out[3] = t2 - t0; }
    }
  } else if (mode == 9 || mode == 10) {
    // shape 1 (9) / shape 0 (10) while the other seven warps spin the way K1's bounded waits do
    __shared__ uint64_t never;
    __shared__ volatile int abort_flag;
    if (tid == 0) { mbar_init(&never, 1); fence_mbar_init(); abort_flag = 0; }
    __syncthreads();
    if (tid == 0) { issue(mode == 9 ? 1 : 0, &bar[0]); stop = 1; }
    if (warp >= 1) {
      long long n = 0;
      const long long t0 = clock64();
      while (!mbar_try_wait(&never, 0)) {
        if (abort_flag || stop) break;
        if (clock64() - t0 > 4000000000ll) break;
        ++n;
      }
      if (blockIdx.x == 0 && tid == 32) out[14] = n;
    }
  } else if (mode >= 6) {
    // modes 6/7/8: shape 0 (6), shape 1 (7), shape 1 + cp.async stream (8) while thread 32 streams 32 KB bulk stages
    // (two 16 KB copies each) from an L2-resident 1 MB buffer into a 3-slot ring at base+64 KB, as fast as they complete
    __shared__ uint64_t wbar[3];
    if (tid == 0) { for (int i = 0; i < 3; ++i) mbar_init(&wbar[i], 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) { issue(mode == 6 ? 0 : 1, &bar[0]); stop = 1; }
    if (tid == 32) {
      int n = 0;
      for (; !stop && n < (1 << 20); ++n) {
        const int sl = n % 3;
        if (n >= 3) mbar_wait(&wbar[sl], ((n / 3) - 1) & 1, err, 2);
        mbar_arrive_expect_tx(&wbar[sl], 32768);
        const unsigned char* src = wsrc + (size_t)((n * 32768 + blockIdx.x * 65536) & (1048576 - 32768));
        bulk_g2s(b0 + sl * 32768, src, 16384, &wbar[sl]);
        bulk_g2s(b0 + sl * 32768 + 16384, src + 16384, 16384, &wbar[sl]);
      }
      if (blockIdx.x == 0) out[13] = n;
    }
    if (mode == 8 && warp >= 4) {
      const int w = warp - 4, lane = tid & 31;
      int n = 0;
      for (; !stop && n < (1 << 20); ++n) {
        const unsigned char* src = xsrc + (size_t)blockIdx.x * 1048576 + (size_t)(n & 63) * 16384 + w * 4096 + lane * 8;
#pragma unroll
        for (int j = 0; j < 16; ++j) cp_async_8(a0 + (n & 3) * 16384 + w * 4096 + j * 256 + lane * 8, src + j * 256, 8);
        cp_async_commit();
        cp_async_wait<2>();
      }
      cp_async_wait<0>();
      if (blockIdx.x == 0 && tid == 128) out[14] = n;
    }
  } else if (tid == 0) {
    issue(mode == 5 ? 5 : mode, &bar[0]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
  long long* d_out; int* d_err;
  cudaMalloc(&d_out, 16 * 8); cudaMalloc(&d_err, 4);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  const char* names[] = {"SS M128 N256 (X as A)", "SS M128 N128 (W as A, X as B)", "TS M128 N96 pool", "two issuers N256 + N128",
                         "SS N128 + st.shared stream", "TS M128 N128 (G2)", "SS N256 + bulk stream", "SS N128 + bulk stream",
                         "SS N128 + bulk + cp.async streams", "SS N128 + 7 spinning warps", "SS N256 + 7 spinning warps", "SS N128 chunks: 1 commit", "SS N128 chunks: 2 commits",
                         "SS N128 chunks: 2 commits + 2 waits", "SS N128 chunks: 2 commits + 2 waits + fence",
                         "N256 commit every 2", "N256 commit every 4", "N256 commit every 8", "N256 commit every 16",
                         "N128 commit every 4", "N128 commit every 8", "N128 commit every 16", "N128 commit every 32",
                         "8xN128 block, gap 0", "8xN128 block, gap 16", "8xN128 block, gap 32", "8xN128 block, gap 64", "8xN128 block, gap 128", "8xN128 block, gap 256",
                         "4xN256 block, gap 0 (x2)", "4xN256 block, gap 16 (x2)", "4xN256 block, gap 32 (x2)", "4xN256 block, gap 64 (x2)", "4xN256 block, gap 128 (x2)", "4xN256 block, gap 256 (x2)"};
  unsigned char *wsrc, *xsrc;
  cudaMalloc(&wsrc, 1 << 20); cudaMemset(wsrc, 0, 1 << 20);
  cudaMalloc(&xsrc, (size_t)148 << 20); cudaMemset(xsrc, 0, (size_t)148 << 20);
  for (int random = 1; random < 2; ++random)
  for (int mode = 23; mode < 35; ++mode) {
    cudaMemset(d_out, 0, 16 * 8); cudaMemset(d_err, 0, 4);
    for (int rep = 0; rep < 3; ++rep) rate_kernel<<<148, 256, 164 * 1024>>>(mode, random, d_out, d_err, wsrc, xsrc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16]; int herr;
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost);
    printf("%s mode %d %-34s cuda=%s err=%d :", random ? "random" : "zeros ", mode, names[mode], cudaGetErrorString(e), herr);
    for (int m = 0; m < 6; ++m)
      if (h[2 * m + 1]) printf("  [shape %d] issue %.1f cyc/MMA, complete %.1f cyc/MMA", m, (double)h[2 * m] / ITER, (double)h[2 * m + 1] / ITER);
    if (mode == 4) printf("  st.shared rounds %lld (x 4 warps x 4 KB)", h[14]);
    if (mode >= 9) printf("  spin iterations per warp %lld", h[14]);
    else if (mode >= 6) printf("  bulk stages %lld x 32 KB = %.1f B/clk", h[13], h[13] * 32768.0 / (h[1] ? h[1] : h[3]));
    if (mode == 8) printf("  cp.async rounds %lld x 16 KB = %.1f B/clk", h[14], h[14] * 16384.0 / h[3]);
    printf("\n");
    if (e != cudaSuccess) return 2;
  }
  return 0;
}
