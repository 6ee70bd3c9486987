// Bring-up probe for cta_group::2 (CTA-pair) MMAs: operand split semantics, TMEM-A variant, multicast commit.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_pair.bin tools/probe_pair.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../protoasnet_b200/csrc/sm100_prims.cuh"
using namespace pasn::sm100;

struct Cfg { int mode; int N; int K; };   // mode 1: SS Kmaj x Kmaj; 2: A MN-major SW128; 3: A in TMEM; 4: MN-major no-swizzle A,B (pool)

// A_full [256][K], B_full [N][K], D_full [256][N].  CTA r owns A rows [128r,+128), B rows [N/2*r, +N/2).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_kernel(const float* A, const float* B, float* D, Cfg c, int* err) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* As = base;
  unsigned char* Bs = base + 32768;
  uint64_t* bar = (uint64_t*)(base + 65536);
  uint32_t* tptr = (uint32_t*)(base + 65536 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int K = c.K, N = c.N, NH = N / 2;
  if (warp == 0) tmem_alloc2(tptr, 512);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  for (int i = tid; i < 65536 / 4; i += 128) ((uint32_t*)As)[i] = 0;
  __syncthreads();
  if (c.mode != 3)
    for (int i = tid; i < 128 * K; i += 128) {
      int m = i / K, k = i % K;
      uint32_t off = c.mode == 1 ? off_kmajor_sw128(m, k) : c.mode == 2 ? off_mnmajor_sw128(m, k, 8192) : off_mnmajor_nosw(m, k, 128);
      *(__nv_bfloat16*)(As + off) = __float2bfloat16_rn(A[(128 * rank + m) * K + k]);
    }
  for (int i = tid; i < NH * K; i += 128) {
    int n = i / K, k = i % K;
    uint32_t off = c.mode == 4 ? off_mnmajor_nosw(n, k, NH) : off_kmajor_sw128(n, k);
    *(__nv_bfloat16*)(Bs + off) = __float2bfloat16_rn(B[(NH * rank + n) * K + k]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (c.mode == 3) {
    uint32_t r[16];
    for (int h = 0; h < K / 32; ++h) {
      for (int j = 0; j < 16; ++j) r[j] = pack_bf16x2(A[(128 * rank + tid) * K + 32 * h + 2 * j], A[(128 * rank + tid) * K + 32 * h + 2 * j + 1]);
      tmem_st_x16(tbase + lane_base + 256 + 16 * h, r);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  cluster_sync_all();
  tc_fence_after();
  if (rank == 0 && tid == 0) {
    uint32_t idesc = make_idesc_bf16(256, N, (c.mode == 2 || c.mode == 4) ? 1 : 0, c.mode == 4 ? 1 : 0);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t bd, ad;
      if (c.mode == 4) {
        uint32_t lbo_b = (NH / 8) * 128, lbo_a = 2048;
        bd = make_smem_desc(smem_u32(Bs) + ks * 2 * lbo_b, lbo_b, 128, SWZ_NONE);
        ad = make_smem_desc(smem_u32(As) + ks * 2 * lbo_a, lbo_a, 128, SWZ_NONE);
      } else {
        bd = make_smem_desc(smem_u32(Bs) + ks * 32, 16, 1024, SWZ_128B);
        ad = c.mode == 1 ? make_smem_desc(smem_u32(As) + ks * 32, 16, 1024, SWZ_128B)
                         : make_smem_desc(smem_u32(As) + ks * 2048, 8192, 1024, SWZ_128B);
      }
      if (c.mode == 3) mma_ts2(tbase, tbase + 256 + ks * 8, bd, idesc, ks > 0);
      else mma_ss2(tbase, ad, bd, idesc, ks > 0);
    }
    mma_commit2(bar, 3);
  }
  mbar_wait(bar, 0, err, 1);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld_x16(tbase + lane_base + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(128 * rank + tid) * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tbase, 512);
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

static bool run(const char* name, Cfg c) {
  const int M = 256;
  std::vector<float> A((size_t)M * c.K), B((size_t)c.N * c.K), D((size_t)M * c.N), R((size_t)M * c.N);
  unsigned s = 777u + c.mode;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  for (auto& v : A) v = bf16r(rnd());
  for (auto& v : B) v = bf16r(rnd());
  for (int m = 0; m < M; ++m) for (int n = 0; n < c.N; ++n) { double a = 0; for (int k = 0; k < c.K; ++k) a += (double)A[m * c.K + k] * B[n * c.K + k]; R[m * c.N + n] = (float)a; }
  float *dA, *dB, *dD; int* dE;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dE, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, D.size() * 4); cudaMemset(dE, 0, 4);
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 68 * 1024);
  pair_kernel<<<2, 128, 68 * 1024>>>(dA, dB, dD, c, dE);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0; cudaMemcpy(&herr, dE, 4, cudaMemcpyDeviceToHost); cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - R[i]); if (!(d <= 1e-3)) ++bad; if (d > maxerr || d != d) maxerr = d; }
  bool ok = e == cudaSuccess && herr == 0 && bad == 0;
  printf("%-52s N=%3d K=%3d : %s maxerr=%.3e bad=%d cuda=%s bar=%d\n", name, c.N, c.K, ok ? "PASS" : "FAIL", maxerr, bad, cudaGetErrorString(e), herr);
  if (!ok && bad) {  // help decode a wrong operand-split assumption: which reference column does output column n match?
    for (int n = 0; n < c.N; n += c.N / 8) { int best = -1; for (int q = 0; q < c.N; ++q) if (fabs(D[5 * c.N + n] - R[5 * c.N + q]) < 1e-3) best = q; printf("   row5 out col %d matches ref col %d\n", n, best); }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dE);
  if (e != cudaSuccess) { printf("fatal\n"); exit(2); }
  return ok;
}

int main() {
  int pass = 0;
  pass += run("P1 pair SS Kmaj x Kmaj M256 N256", {1, 256, 64});
  pass += run("P2 pair SS A MN-major SW128 M256 N256", {2, 256, 64});
  pass += run("P3 pair TS A in TMEM M256 N128", {3, 128, 64});
  pass += run("P3b pair TS A in TMEM M256 N64", {3, 64, 64});
  pass += run("P4 pair SS MN-major no-swizzle A,B M256 N160 (pool)", {4, 160, 128});
  printf("passed %d\n", pass);
  return 0;
}
