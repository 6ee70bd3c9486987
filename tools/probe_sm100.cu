// Bring-up probe for the tcgen05 building blocks of the fused head kernel (run on a B200 via gpurun).
// Each mode issues one small GEMM with the exact descriptor / layout helpers of csrc/sm100_prims.cuh and compares
// against a CPU matmul; uncertain conventions are run both ways so one launch disambiguates them.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_sm100.bin tools/probe_sm100.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../protoasnet_b200/csrc/sm100_prims.cuh"

using namespace pasn::sm100;

struct ProbeCfg {
  int mode;      // 1: A Kmaj-SW128 x B Kmaj-SW128; 2: A MNmaj-SW128; 3: A in TMEM; 4: A,B MNmaj no-swizzle;
                 // 5: A Kmaj-SW128 x B MNmaj-SW128 (weights as A, X tile as B)
  int M, N, K;
  int swap;      // try the alternative convention (LBO<->SBO, or bf16 half order for mode 3)
  int b_row0;    // mode 3: B operand starts at this row of a taller image
};

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D, ProbeCfg c, int* err) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* As = base;            // up to 32 KB
  unsigned char* Bs = base + 32768;    // up to 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  // fill operand images
  for (int i = tid; i < 32768 / 4; i += 128) { ((uint32_t*)As)[i] = 0; ((uint32_t*)Bs)[i] = 0; }
  __syncthreads();
  const int M = c.M, N = c.N, K = c.K;
  if (c.mode != 3) {
    for (int i = tid; i < M * K; i += 128) {
      int m = i / K, k = i % K;
      uint32_t off;
      if (c.mode == 1 || c.mode == 5) off = off_kmajor_sw128(m, k);
      else if (c.mode == 2) off = off_mnmajor_sw128(m, k, 8192);
      else off = off_mnmajor_nosw(m, k, M);
      *(__nv_bfloat16*)(As + off) = __float2bfloat16_rn(A[i]);
    }
  }
  const int b_rows_total = c.b_row0 + N;
  for (int i = tid; i < N * K; i += 128) {
    int n = i / K, k = i % K;
    uint32_t off;
    if (c.mode == 4) off = off_mnmajor_nosw(n, k, N);
    else if (c.mode == 5) off = off_mnmajor_sw128(n, k, 8192);
    else off = off_kmajor_sw128(n + c.b_row0, k);
    *(__nv_bfloat16*)(Bs + off) = __float2bfloat16_rn(B[i]);
  }
  (void)b_rows_total;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  if (c.mode == 3) {  // A[m][0..63] -> TMEM columns [256, 288), 2 bf16 per column
    uint32_t r[16];
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float lo = A[tid * K + 32 * h + 2 * j], hi = A[tid * K + 32 * h + 2 * j + 1];
        r[j] = c.swap ? pack_bf16x2(hi, lo) : pack_bf16x2(lo, hi);
      }
      tmem_st_x16(tbase + lane_base + 256 + 16 * h, r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (tid == 0) {
    uint32_t idesc;
    if (c.mode == 1) idesc = make_idesc_bf16(M, N, 0, 0);
    else if (c.mode == 2) idesc = make_idesc_bf16(M, N, 1, 0);
    else if (c.mode == 3) idesc = make_idesc_bf16(M, N, 0, 0);
    else if (c.mode == 5) idesc = make_idesc_bf16(M, N, 0, 1);
    else idesc = make_idesc_bf16(M, N, 1, 1);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t bdesc;
      if (c.mode == 4) {
        uint32_t lbo = (N / 8) * 128, sbo = 128;
        bdesc = c.swap ? make_smem_desc(smem_u32(Bs) + ks * 2 * lbo, sbo, lbo, SWZ_NONE)
                       : make_smem_desc(smem_u32(Bs) + ks * 2 * lbo, lbo, sbo, SWZ_NONE);
      } else if (c.mode == 5) {
        bdesc = make_smem_desc(smem_u32(Bs) + ks * 2048, 8192, 1024, SWZ_128B);
      } else {
        bdesc = make_smem_desc(smem_u32(Bs) + c.b_row0 * 128 + ks * 32, 16, 1024, SWZ_128B);
      }
      if (c.mode == 3) {
        mma_ts(tbase, tbase + 256 + ks * 8, bdesc, idesc, ks > 0);
      } else {
        uint64_t adesc;
        if (c.mode == 1 || c.mode == 5) adesc = make_smem_desc(smem_u32(As) + ks * 32, 16, 1024, SWZ_128B);
        else if (c.mode == 2)
          adesc = c.swap ? make_smem_desc(smem_u32(As) + ks * 2048, 1024, 8192, SWZ_128B)
                         : make_smem_desc(smem_u32(As) + ks * 2048, 8192, 1024, SWZ_128B);
        else {
          uint32_t lbo = (M / 8) * 128, sbo = 128;
          adesc = c.swap ? make_smem_desc(smem_u32(As) + ks * 2 * lbo, sbo, lbo, SWZ_NONE)
                         : make_smem_desc(smem_u32(As) + ks * 2 * lbo, lbo, sbo, SWZ_NONE);
        }
        mma_ss(tbase, adesc, bdesc, idesc, ks > 0);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0, err, 1);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld_x16(tbase + lane_base + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

static bool run(const char* name, ProbeCfg c) {
  std::vector<float> A((size_t)c.M * c.K), B((size_t)c.N * c.K), D((size_t)c.M * c.N), R((size_t)c.M * c.N);
  unsigned s = 12345u + c.mode * 77u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  for (auto& v : A) v = bf16r(rnd());
  for (auto& v : B) v = bf16r(rnd());
  for (int m = 0; m < c.M; ++m)
    for (int n = 0; n < c.N; ++n) {
      double acc = 0;
      for (int k = 0; k < c.K; ++k) acc += (double)A[m * c.K + k] * B[n * c.K + k];
      R[m * c.N + n] = (float)acc;
    }
  float *dA, *dB, *dD; int* dE;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dE, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, D.size() * 4); cudaMemset(dE, 0, 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024 + 1024);
  probe_kernel<<<1, 128, 66 * 1024 + 1024>>>(dA, dB, dD, c, dE);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0;
  cudaMemcpy(&herr, dE, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (size_t i = 0; i < D.size(); ++i) {
    double d = fabs((double)D[i] - R[i]);
    if (!(d <= 1e-3)) ++bad;
    if (d > maxerr || d != d) maxerr = d;
  }
  bool ok = (e == cudaSuccess) && herr == 0 && bad == 0;
  printf("%-44s M=%3d N=%3d K=%3d swap=%d : %s  maxerr=%.3e bad=%d cuda=%s barrier_err=%d\n", name, c.M, c.N, c.K, c.swap,
         ok ? "PASS" : "FAIL", maxerr, bad, cudaGetErrorString(e), herr);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dE);
  if (e != cudaSuccess) { printf("fatal CUDA error, stopping\n"); exit(2); }
  return ok;
}

int main() {
  int pass = 0;
  pass += run("T1 SS  A Kmaj-SW128, B Kmaj-SW128", {1, 128, 256, 64, 0, 0});
  pass += run("T1b SS same, N=64", {1, 128, 64, 64, 0, 0});
  pass += run("T2 SS  A MNmaj-SW128 (LBO=mn-atom, SBO=k-atom)", {2, 128, 256, 64, 0, 0});
  pass += run("T2s SS A MNmaj-SW128 (LBO/SBO swapped)", {2, 128, 256, 64, 1, 0});
  pass += run("T3 TS  A in TMEM (lo half = even k)", {3, 128, 64, 64, 0, 64});
  pass += run("T3s TS A in TMEM (hi half = even k)", {3, 128, 64, 64, 1, 64});
  pass += run("T3c TS A in TMEM, N=128, B row0=0", {3, 128, 128, 64, 0, 0});
  pass += run("T4 SS  A,B MNmaj no-swizzle (LBO=k-grp, SBO=mn-grp)", {4, 128, 80, 128, 0, 0});
  pass += run("T4s SS A,B MNmaj no-swizzle (swapped)", {4, 128, 80, 128, 1, 0});
  pass += run("T5 SS  A Kmaj-SW128, B MNmaj-SW128 N=128", {5, 128, 128, 64, 0, 0});
  printf("passed %d\n", pass);
  return 0;
}
