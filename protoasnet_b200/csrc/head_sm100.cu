// Fused prototype-head kernels for sm_100a (tcgen05 tensor cores, TMEM accumulators, bulk-async weight streaming).
//
// Reference computation being replaced (per clip, S = T*H*W voxels; src/models/Video_XProtoNet.py:82-98):
//     H1 = relu(W1 x + b1)            add_on_layers[0..1]     :27-39
//     F  = W2 H1 + b2                 add_on_layers[2]
//     G1 = relu(W3 x + b3); G2 = relu(W4 G1 + b4); O = |W5 G2|   occurrence_module + abs   :42-62, :106-109
//     FE[p,:] = sum_s O[p,s] F[:,s]   occurrence-weighted pooling   :87
//     cos / (.+1)/2 / logits          :90-96
//
// Algorithmic restructuring (exact in real arithmetic): pooling commutes with the last add-on conv,
//     FE[p,:] = W2 (sum_s O[p,s] H1[:,s]) + b2 * (sum_s O[p,s]),
// so W2 is applied to P pooled vectors per clip instead of S voxels (D*D*S -> D*D*P MACs per clip).
//
// K1  head_tokens_kernel   one persistent CTA per SM walks a contiguous range of clips in tiles of 128 voxels
//                          ("tokens" = rows of the MMA M dimension).  Per tile, all on tensor cores:
//        acc_G = X W3^T, acc_A = X W1^T     (SS MMA, X tile MN-major from NCDHW, weights streamed by cp.async.bulk)
//        G1 -> TMEM (bf16) -> acc_G2 = G1 W4^T -> G2 -> TMEM -> acc_O = G2 W5^T        (A operand from TMEM)
//        O = |acc_O| -> Os in smem: B operand of the pooling MMA, source of the occurrence-map store and of Osum
//        FEpre^T[d, (slot,p)] += H1^T O      (MN-major SS MMA; "slot" separates the <=2 clips a tile touches)
//      Hidden activations never touch HBM.  Outputs: occurrence_map, Osum [N,P] fp32 and FEpre as bf16 hi/lo
//      operand images (already in the K-major SWIZZLE_128B byte order K2 feeds to its MMAs).
// K2  proto_w2_kernel      FE = FEpre W2^T + b2 Osum on tensor cores (hi/lo split keeps ~fp32 accuracy), then the fp32
//                          cosine -> (.+1)/2 -> logits -> 1-s -> packed argmin keys chain (cf. proto_stage.cu).
#include <cstdlib>

#include "head_sm100_shared.cuh"

namespace pasn {
using namespace sm100;

using namespace k1;

namespace {

constexpr int XSLOTS = 4;
#ifndef PASN_WSLOTS
#define PASN_WSLOTS 3
#endif
constexpr int WSLOTS = PASN_WSLOTS;
constexpr uint32_t XSLOT_BYTES = 16384, WSLOT_BYTES = 32768, HS_BYTES = 32768;
constexpr int K1_WARPS = 16;
constexpr int EPI_WARP0 = 6;   // warps 6..13: epilogue (TMEM quadrant = warp % 4), 14: Osum, 15: occurrence-map store
constexpr int K1_THREADS = K1_WARPS * 32;
// shared-memory map of K1 (offsets from a 1024-byte aligned base)
constexpr uint32_t SM_X = 0;
constexpr uint32_t SM_W = SM_X + XSLOTS * XSLOT_BYTES;            // 65536
constexpr uint32_t SM_HS = SM_W + WSLOTS * WSLOT_BYTES;           // 163840
constexpr uint32_t SM_OS = SM_HS + HS_BYTES;                      // 196608
constexpr uint32_t OS_BYTES_MAX = TILE_M * 2 * PP_MAX * 2;        // 24576
constexpr uint32_t SM_BIAS = SM_OS + OS_BYTES_MAX;                // 221184  b3[256] b1[256] b4[128] fp32
constexpr uint32_t SM_BAR = SM_BIAS + (DD + DD + DH) * 4;         // 223744
constexpr uint32_t SM_MISC = SM_BAR + 32 * 8;                     // 224000
constexpr uint32_t K1_SMEM = SM_MISC + 64;                        // 224064

enum {
  B_XFULL = 0, B_XEMPTY = 4, B_WFULL = 8, B_WEMPTY = 11, B_L1DONE = 14, B_G1READY, B_G2DONE, B_G2READY, B_ODONE,
  B_OSREADY, B_OSEMPTY, B_HSREADY, B_HSEMPTY, B_FEDONE, B_TMEMFREE, B_R1FREE, B_COUNT
};
static_assert(B_COUNT <= 32, "barrier table");

}  // namespace

// =================================================================================================
// K1
// =================================================================================================
template <int PP>  // padded prototype count (multiple of 8, <= PP_MAX)
__global__ void __launch_bounds__(K1_THREADS, 1) head_tokens_kernel(const K1Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + SM_MISC);
  volatile int* abort_s = reinterpret_cast<volatile int*>(smem + SM_MISC + 8);
  float* sb3 = reinterpret_cast<float*>(smem + SM_BIAS);
  float* sb1 = sb3 + DD;
  float* sb4 = sb1 + DD;

  constexpr int NPOOL = 2 * PP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c_begin = blockIdx.x * p.clips_per_cta;
  int ncl = p.N - c_begin;
  if (ncl > p.clips_per_cta) ncl = p.clips_per_cta;
  if (ncl <= 0) return;
  const int S = p.S;
  const int ntok = ncl * S;
  const int ntiles = (ntok + TILE_M - 1) / TILE_M;
  const PackedLayout PL = packed_layout(p.C);
  Ctx ctx{p.err, abort_s};

  if ((smem_u32(smem) & 1023u) != 0) {  // swizzled layouts need the 1024-byte alignment we asked for
    if (tid == 0) atomicCAS(p.err, 0, 900);
    return;
  }

  if (tid == 0) {
    *abort_s = 0;
    for (int i = 0; i < 4; ++i) { mbar_init(&bars[B_XFULL + i], 4); mbar_init(&bars[B_XEMPTY + i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&bars[B_WFULL + i], 1); mbar_init(&bars[B_WEMPTY + i], 1); }
    mbar_init(&bars[B_L1DONE], 1);
    mbar_init(&bars[B_G1READY], 8);
    mbar_init(&bars[B_G2DONE], 1);
    mbar_init(&bars[B_G2READY], 8);
    mbar_init(&bars[B_ODONE], 1);
    mbar_init(&bars[B_OSREADY], 8);
    mbar_init(&bars[B_OSEMPTY], 3);   // pooling MMAs retired + Osum warp + occurrence-map store warp
    mbar_init(&bars[B_HSREADY], 8);
    mbar_init(&bars[B_HSEMPTY], 1);
    mbar_init(&bars[B_FEDONE], 1);
    mbar_init(&bars[B_TMEMFREE], 8);
    mbar_init(&bars[B_R1FREE], 8);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_s, 512);
  {
    const float* gb = reinterpret_cast<const float*>(p.packed + PL.off_bias);
    for (int i = tid; i < DD + DD + DH; i += K1_THREADS) sb3[i] = gb[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch_dependents();   // K2 may start launching (its CTAs only fit on an SM once one of ours has exited)
  const uint32_t tbase = *tmem_ptr_s;
  const uint32_t x_base = smem_u32(smem + SM_X), w_base = smem_u32(smem + SM_W);
  const uint32_t hs_base = smem_u32(smem + SM_HS), os_base = smem_u32(smem + SM_OS);
  const int nkc = p.nkc;
  const int stages_per_tile = 2 * nkc + 3;

  // TMEM column map (512 columns x 128 lanes; lane = voxel row of the tile unless noted)
  //   [  0,256) acc_G fp32  -> G1 bf16 at [0,64) and [192,256) -> acc_O fp32 [0,64) -> FEpartial^T half 0 [0,NPOOL)
  //   [ 64,192) acc_G2 fp32 -> G2 bf16 at [64,96) and [128,160) -> FEpartial^T half 1 [128,128+NPOOL)  (lane = d)
  //   [256,512) acc_A fp32 (H1 pre-activation), drained to smem by the epilogue
  if (warp == 0) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc_l1 = make_idesc_bf16(128, 256, 1, 0);
      const uint32_t idesc_g2 = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_pool = make_idesc_bf16(128, NPOOL, 1, 1);
      constexpr uint32_t lbo_os = (uint32_t)(NPOOL / 8) * 128u;
      uint32_t wst = 0;
      bool ok = true;
      // RA = leading 64-channel chunks whose add-on pass (acc_A) is issued ahead of time, during the previous tile's
      // tail: acc_A's columns are free as soon as the epilogue has pulled H1 into registers, long before acc_G's.
      const int RA = nkc < 4 ? nkc : 4;
      long long xwait = 0, wwait = 0;  // cycles the issue thread spent blocked on X chunks / weight stages (trace only)
      auto wait_x = [&](int tile, int kc) -> bool {
        const uint32_t g = (uint32_t)(tile * nkc + kc);
        const long long t0 = clock64();
        const bool r = bwait(&bars[B_XFULL + (g % XSLOTS)], (g / XSLOTS) & 1, ctx, 102);
        const long long dt = clock64() - t0;
        xwait += dt;
        if (p.trace != nullptr && blockIdx.x == 0 && tile < 16 && kc < 8) p.trace[(2 * 16 + tile) * 16 + kc] = dt;
        return r;
      };
      auto free_x = [&](int tile, int kc) {
        const uint32_t g = (uint32_t)(tile * nkc + kc);
        mma_commit(&bars[B_XEMPTY + (g % XSLOTS)]);
      };
      auto issue_pass = [&](int tile, int kc, int pass) -> bool {  // 4 MMAs: one chunk into acc_G (pass 0) / acc_A (pass 1)
        const uint32_t g = (uint32_t)(tile * nkc + kc), xs = g % XSLOTS;
        const uint32_t ws = wst % WSLOTS, wph = (wst / WSLOTS) & 1;
        ++wst;
        const long long t0 = clock64();
        if (!bwait(&bars[B_WFULL + ws], wph, ctx, 103)) return false;
        {
          const long long dt = clock64() - t0;
          wwait += dt;
          if (p.trace != nullptr && blockIdx.x == 0 && tile < 16 && kc < 8) p.trace[(2 * 16 + tile) * 16 + 8 + kc] += dt;
        }
        tc_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint64_t ad = make_smem_desc(x_base + xs * XSLOT_BYTES + k4 * 2048, 8192, 1024, SWZ_128B);
          const uint64_t bd = make_smem_desc(w_base + ws * WSLOT_BYTES + k4 * 32, 16, 1024, SWZ_128B);
          mma_ss(tbase + (pass ? 256u : 0u), ad, bd, idesc_l1, (kc | k4) ? 1u : 0u);
        }
        mma_commit(&bars[B_WEMPTY + ws]);
        return true;
      };
      for (int kc = 0; kc < RA && ok; ++kc) ok = wait_x(0, kc) && issue_pass(0, kc, 1);  // tile 0 has no predecessor
      for (int tile = 0; tile < ntiles && ok; ++tile) {
        const uint32_t tp = tile & 1;
        K1_TRACE(0, tile, 0);
        if (!(ok = bwait(&bars[B_TMEMFREE], tp ^ 1, ctx, 101))) break;   // acc_G columns drained by the previous tile
        tc_fence_after();
        K1_TRACE(0, tile, 1);
        // ---- layer 1: acc_G (cols 0..255) for every chunk, acc_A (cols 256..511) for the chunks not issued ahead
        for (int kc = 0; kc < RA && ok; ++kc) {
          ok = issue_pass(tile, kc, 0);
          free_x(tile, kc);
        }
        for (int kc = RA; kc < nkc && ok; ++kc) {
          ok = wait_x(tile, kc) && issue_pass(tile, kc, 0) && issue_pass(tile, kc, 1);
          free_x(tile, kc);
        }
        if (!ok) break;
        mma_commit(&bars[B_L1DONE]);
        K1_TRACE(0, tile, 2);
        if (p.trace != nullptr && blockIdx.x == 0 && tile < 16) {
          p.trace[(0 * 16 + tile) * 16 + 11] = xwait;
          p.trace[(0 * 16 + tile) * 16 + 12] = wwait;
        }
        xwait = wwait = 0;
        // ---- G2 = G1 W4^T : A from TMEM (G1 bf16 at cols [0,64) and [192,256)), D = cols [64,192), N = 128
        if (!(ok = bwait(&bars[B_G1READY], tp, ctx, 104))) break;
        tc_fence_after();
        K1_TRACE(0, tile, 3);
        for (int st = 0; st < 2 && ok; ++st, ++wst) {
          const uint32_t ws = wst % WSLOTS, wph = (wst / WSLOTS) & 1;
          if (!(ok = bwait(&bars[B_WFULL + ws], wph, ctx, 105))) break;
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int ks = st * 8 + kk;  // 0..15, K-step of 16 channels
            const uint64_t bd =
                make_smem_desc(w_base + ws * WSLOT_BYTES + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, SWZ_128B);
            const uint32_t a_col = ks < 8 ? 8u * ks : 192u + 8u * (ks - 8);
            mma_ts(tbase + 64u, tbase + a_col, bd, idesc_g2, ks ? 1u : 0u);
          }
          mma_commit(&bars[B_WEMPTY + ws]);
        }
        if (!ok) break;
        mma_commit(&bars[B_G2DONE]);
        K1_TRACE(0, tile, 4);
        // ---- O = G2 W5^T : A from TMEM (G2 bf16 at cols [64,96) and [128,160)), D at cols [0,64)
        if (!(ok = bwait(&bars[B_G2READY], tp, ctx, 106))) break;
        K1_TRACE(0, tile, 5);
        {
          const uint32_t ws = wst % WSLOTS, wph = (wst / WSLOTS) & 1;
          if (!(ok = bwait(&bars[B_WFULL + ws], wph, ctx, 107))) break;
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t a_col = ks < 4 ? 64u + 8u * ks : 128u + 8u * (ks - 4);
            const uint64_t bd =
                make_smem_desc(w_base + ws * WSLOT_BYTES + (ks >> 2) * 8192 + (ks & 3) * 32, 16, 1024, SWZ_128B);
            mma_ts(tbase + 0u, tbase + a_col, bd, idesc_o, ks ? 1u : 0u);
          }
          mma_commit(&bars[B_WEMPTY + ws]);
          ++wst;
        }
        mma_commit(&bars[B_ODONE]);
        K1_TRACE(0, tile, 6);
        // ---- run-ahead: add-on pass of the next tile's first chunks (acc_A is free once H1 sits in registers)
        const bool more = tile + 1 < ntiles;
        const int ra_first = RA / 2;
        if (more) {
          if (!(ok = bwait(&bars[B_R1FREE], tp, ctx, 110))) break;
          tc_fence_after();
          for (int kc = 0; kc < ra_first && ok; ++kc) ok = wait_x(tile + 1, kc) && issue_pass(tile + 1, kc, 1);
          if (!ok) break;
        }
        // ---- pooling: FEpartial^T[d, (slot,p)] = H1^T O ; d halves at cols [0,NPOOL) and [128,128+NPOOL)
        if (!(ok = bwait(&bars[B_OSREADY], tp, ctx, 108))) break;
        K1_TRACE(0, tile, 7);
        for (int half = 0; half < 2 && ok; ++half) {
          if (!(ok = bwait(&bars[B_HSREADY], (uint32_t)half, ctx, 109))) break;  // use #(2*tile+half): parity = half
          tc_fence_after();
          K1_TRACE(0, tile, 8 + half);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ad = make_smem_desc(hs_base + ks * 4096, 2048, 128, SWZ_NONE);
            const uint64_t bd = make_smem_desc(os_base + ks * 2 * lbo_os, lbo_os, 128, SWZ_NONE);
            mma_ss(tbase + (half ? 128u : 0u), ad, bd, idesc_pool, ks ? 1u : 0u);
          }
          mma_commit(&bars[B_HSEMPTY]);
        }
        if (!ok) break;
        mma_commit(&bars[B_OSEMPTY]);
        mma_commit(&bars[B_FEDONE]);
        K1_TRACE(0, tile, 10);
        if (more)
          for (int kc = ra_first; kc < RA && ok; ++kc) ok = wait_x(tile + 1, kc) && issue_pass(tile + 1, kc, 1);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ weight producer (one thread)
    if (lane == 0) {
      uint32_t wst = 0;
      bool ok = true;
      for (int tile = 0; tile < ntiles && ok; ++tile) {
        for (int i = 0; i < stages_per_tile; ++i, ++wst) {
          const uint32_t ws = wst % WSLOTS, wph = (wst / WSLOTS) & 1;
          if (!(ok = bwait(&bars[B_WEMPTY + ws], wph ^ 1, ctx, 201))) break;
          size_t src;
          uint32_t bytes = 32768;
          if (i < 2 * nkc) {  // consumption order: W1[0..RA) (issued ahead), W3[0..RA), then (W3[k], W1[k]) for k >= RA
            const int RA = nkc < 4 ? nkc : 4;
            int kc, pass;
            if (i < RA) { kc = i; pass = 1; }
            else if (i < 2 * RA) { kc = i - RA; pass = 0; }
            else { kc = RA + ((i - 2 * RA) >> 1); pass = (i - 2 * RA) & 1; }
            src = PL.off_l1 + (size_t)(2 * kc + pass) * 32768;
          } else if (i < 2 * nkc + 2) src = PL.off_w4 + (size_t)(i - 2 * nkc) * 32768;
          else { src = PL.off_w5; bytes = 16384; }
          if (p.dbg_skip & 1) { mbar_arrive(&bars[B_WFULL + ws]); continue; }
          mbar_arrive_expect_tx(&bars[B_WFULL + ws], bytes);
          for (uint32_t o = 0; o < bytes; o += 16384)
            bulk_g2s(w_base + ws * WSLOT_BYTES + o, p.packed + src + o, 16384, &bars[B_WFULL + ws]);
        }
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ------------------------------------------------------------------ X producers: NCDHW gather -> MN-major SW128
    // Four warps, each owning 16 of the 64 channels of a chunk; lane l owns voxels 4l..4l+3 of the tile, so one
    // warp-wide 8-byte load is a coalesced 256-byte run of one channel row.  global -> registers with an L1-bypassing
    // load -> st.shared into the swizzled operand layout.  The loads of the next chunk are issued before the stores
    // of the current one and before waiting for its slot, so HBM latency overlaps both (32 KB in flight per SM).
    // cp.async.ca was measured at ~50 cycles per instruction here: with >196 KB of smem carved out there is no L1
    // left for its allocate-on-miss path, and the 16-byte L1-bypassing form needs an alignment NCDHW rows
    // (392 B pitch) only have for every other channel.
    const int xw = warp - 2;
    const uint32_t nchunks = (uint32_t)(ntiles * nkc);
    bool ok = true;
    uint2 va[16], vb[16];
    auto load_unit = [&](uint32_t g, uint2* v) {
      const int tile = (int)(g / (uint32_t)nkc), kc = (int)(g - (uint32_t)tile * nkc);
      const int t = tile * TILE_M + 4 * lane;
      const bool valid = t < ntok && !(p.dbg_skip & 2);
      const int clipl = valid ? t / S : 0;
      const int s = valid ? t - clipl * S : 0;
      const __nv_bfloat16* src = p.feat + ((size_t)(c_begin + clipl) * p.C + kc * 64 + xw * 16) * S + s;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = valid ? ldg_nc_na_v2(src + (size_t)j * S) : make_uint2(0u, 0u);
    };
    auto store_unit = [&](uint32_t g, const uint2* v) -> bool {
      const uint32_t xs = g % XSLOTS, xph = (g / XSLOTS) & 1;
      if (!bwait(&bars[B_XEMPTY + xs], xph ^ 1, ctx, 301)) return false;
      const uint32_t dst0 = x_base + xs * XSLOT_BYTES;
#pragma unroll
      for (int j = 0; j < 16; ++j) st_shared_v2(dst0 + off_mnmajor_sw128(4 * lane, xw * 16 + j, 8192), v[j]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_XFULL + xs]);
      return true;
    };
    if (!p.f32_in) {
      if (nchunks > 0) load_unit(0, va);
      for (uint32_t g = 0; g < nchunks && ok; g += 2) {
        if (g + 1 < nchunks) load_unit(g + 1, vb);
        if (!(ok = store_unit(g, va))) break;
        if (g + 1 < nchunks) {
          if (g + 2 < nchunks) load_unit(g + 2, va);
          ok = store_unit(g + 1, vb);
        }
      }
    } else {
      // fp32 feature maps ("bf16 compute" mode, opt-in): 16-byte loads of 4 voxels, rounded to bf16 on the way into
      // smem.  Units are 8 channel rows (u = 2*chunk + half) so the raw loads of the next unit fit in registers.
      float4* qa = reinterpret_cast<float4*>(va);   // 8 x float4 alias the 16 x uint2 buffers
      float4* qb = reinterpret_cast<float4*>(vb);
      const uint32_t nunits = 2 * nchunks;
      auto load32 = [&](uint32_t u, float4* q) {
        const uint32_t g = u >> 1, h = u & 1;
        const int tile = (int)(g / (uint32_t)nkc), kc = (int)(g - (uint32_t)tile * nkc);
        const int t = tile * TILE_M + 4 * lane;
        const bool valid = t < ntok;
        const int clipl = valid ? t / S : 0;
        const int s = valid ? t - clipl * S : 0;
        const float* src = p.feat32 + ((size_t)(c_begin + clipl) * p.C + kc * 64 + xw * 16 + 8 * h) * S + s;
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = valid ? ldg_nc_na_v4f(src + (size_t)j * S) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      auto store32 = [&](uint32_t u, const float4* q) -> bool {
        const uint32_t g = u >> 1, h = u & 1;
        const uint32_t xs = g % XSLOTS, xph = (g / XSLOTS) & 1;
        if (h == 0 && !bwait(&bars[B_XEMPTY + xs], xph ^ 1, ctx, 301)) return false;
        const uint32_t dst0 = x_base + xs * XSLOT_BYTES;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v2(dst0 + off_mnmajor_sw128(4 * lane, xw * 16 + 8 * h + j, 8192),
                       make_uint2(pack_bf16x2(q[j].x, q[j].y), pack_bf16x2(q[j].z, q[j].w)));
        if (h == 1) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[B_XFULL + xs]);
        }
        return true;
      };
      if (nunits > 0) load32(0, qa);
      for (uint32_t u = 0; u < nunits && ok; u += 2) {
        load32(u + 1, qb);
        if (!(ok = store32(u, qa))) break;
        if (u + 2 < nunits) load32(u + 2, qa);
        ok = store32(u + 1, qb);
      }
    }
  } else if (warp == 14) {
    // ------------------------------------------------------------------ occurrence column sums (bias term of W2)
    float acc0 = 0.f, acc1 = 0.f;  // p = lane, p = lane + 32
    bool ok = true;
    const unsigned char* os = smem + SM_OS;
    for (int tile = 0; tile < ntiles && ok; ++tile) {
      if (!(ok = bwait(&bars[B_OSREADY], tile & 1, ctx, 401))) break;
      float s0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f};  // [slot]
#pragma unroll 1
      for (int slot = 0; slot < 2; ++slot) {
        const int n0 = slot * PP + lane, n1 = n0 + 32;
        if (lane < PP) {
          float a = 0.f;
#pragma unroll 8
          for (int tok = 0; tok < TILE_M; ++tok)
            a += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(os + off_mnmajor_nosw(n0, tok, NPOOL)));
          s0[slot] = a;
        }
        if (lane + 32 < PP) {
          float a = 0.f;
#pragma unroll 8
          for (int tok = 0; tok < TILE_M; ++tok)
            a += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(os + off_mnmajor_nosw(n1, tok, NPOOL)));
          s1[slot] = a;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_OSEMPTY]);
      const int last_tok = min(tile * TILE_M + TILE_M - 1, ntok - 1);
      const int first_clip = (tile * TILE_M) / S, last_clip = last_tok / S;
      acc0 += s0[0]; acc1 += s1[0];
      if (last_clip > first_clip) {
        if (lane < p.P) p.osum[(size_t)(c_begin + first_clip) * p.P + lane] = acc0;
        if (lane + 32 < p.P) p.osum[(size_t)(c_begin + first_clip) * p.P + lane + 32] = acc1;
        acc0 = s0[1]; acc1 = s1[1];
      }
      if ((last_tok + 1) % S == 0) {
        if (lane < p.P) p.osum[(size_t)(c_begin + last_clip) * p.P + lane] = acc0;
        if (lane + 32 < p.P) p.osum[(size_t)(c_begin + last_clip) * p.P + lane + 32] = acc1;
        acc0 = acc1 = 0.f;
      }
    }
  } else if (warp == 15) {
    // ------------------------------------------------------------------ occurrence-map store: Os (smem) -> [N][P][S] bf16
    bool ok = true;
    const unsigned char* os = smem + SM_OS;
    for (int tile = 0; tile < ntiles && ok; ++tile) {
      if (!(ok = bwait(&bars[B_OSREADY], tile & 1, ctx, 402))) break;
      if (p.occ != nullptr || p.occ32 != nullptr) {
        const int first_clip = (tile * TILE_M) / S;
#pragma unroll 1
        for (int grp = 0; grp < 4; ++grp) {
          const int tok = grp * 32 + lane;
          const int t = tile * TILE_M + tok;
          if (t < ntok) {
            const int clipl = t / S, s = t - clipl * S, slot = clipl - first_clip;
            const size_t o0 = ((size_t)(c_begin + clipl) * p.P) * S + s;
            const unsigned char* src = os + off_mnmajor_nosw(slot * PP, tok, NPOOL);
            if (!p.f32_in) {
              __nv_bfloat16* orow = p.occ + o0;
#pragma unroll 8
              for (int pp = 0; pp < p.P; ++pp)
                orow[(size_t)pp * S] = *reinterpret_cast<const __nv_bfloat16*>(src + (pp >> 3) * 128 + (pp & 7) * 2);
            } else {
              float* orow = p.occ32 + o0;
#pragma unroll 8
              for (int pp = 0; pp < p.P; ++pp)
                orow[(size_t)pp * S] = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(src + (pp >> 3) * 128 + (pp & 7) * 2));
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_OSEMPTY]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 6..13
    const int q = warp & 3, hh = (warp - EPI_WARP0) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t tl = tbase + lane_base;
    const int tok = q * 32 + lane;
    float facc[PP];
#pragma unroll
    for (int i = 0; i < PP; ++i) facc[i] = 0.f;
    bool ok = true;

    // acc_A half hf (64 columns of this warp) -> H1 = relu(. + b1) as 32 packed bf16x2 registers
    auto h1_convert = [&](int hf, uint32_t* hp) {
      uint32_t ra[32], rb[32];
      const uint32_t col = 256u + 128u * hf + 64u * hh;
      tmem_ld_x32(tl + col, ra);
      tmem_ld_wait();
      tmem_ld_x32(tl + col + 32, rb);
      bias_relu_pack(ra, sb1 + 128 * hf + 64 * hh, hp);
      tmem_ld_wait();
      bias_relu_pack(rb, sb1 + 128 * hf + 64 * hh + 32, hp + 16);
    };
    // packed H1 -> Hs (MN-major no-swizzle [128 tok x 128 d]) once the pooling MMAs of the previous half retired
    auto h1_store = [&](int hf, const uint32_t* hp) -> bool {
      if (!bwait(&bars[B_HSEMPTY], (uint32_t)(hf ^ 1), ctx, 502)) return false;  // use #(2*tile+hf)
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(smem + SM_HS + off_mnmajor_nosw(64 * hh + 8 * g, tok, 128)) =
            make_uint4(hp[4 * g], hp[4 * g + 1], hp[4 * g + 2], hp[4 * g + 3]);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_HSREADY]);
      return true;
    };

    for (int tile = 0; tile < ntiles && ok; ++tile) {
      const uint32_t tp = tile & 1;
      const int t = tile * TILE_M + tok;
      const bool valid = t < ntok;
      const int first_clip = (tile * TILE_M) / S;
      const int clipl = valid ? t / S : first_clip;
      const int slot = clipl - first_clip;
      uint32_t hp[32];

      // ---- E1: acc_G -> G1 = relu(. + b3) bf16, in place.  Warp hh=0 walks its four 32-column chunks upwards and packs
      //      channels 0..127 into cols [0,64); warp hh=1 walks downwards and packs channels 128..255 into [192,256).
      //      Either order only overwrites columns whose fp32 content was already loaded, and it leaves [64,192)
      //      free as one contiguous N=128 accumulator for G2.
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 0);
      if (!(ok = bwait(&bars[B_L1DONE], tp, ctx, 501))) break;
      tc_fence_after();
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 1);
      {
        // chunk order: hh=0 ascending 0,1,2,3 -> writes [16c,+16); hh=1 descending 3,2,1,0 -> writes [192+16c,+16)
        uint32_t ra[32], rb[32], pk[16];
        const int c0 = hh ? 3 : 0, dc = hh ? -1 : 1;
        const uint32_t src = 128u * hh, dst = hh ? 192u : 0u;
        tmem_ld_x32(tl + src + 32 * c0, ra);
        tmem_ld_wait();
        tmem_ld_x32(tl + src + 32 * (c0 + dc), rb);
        bias_relu_pack(ra, sb3 + src + 32 * c0, pk);
        tmem_st_x16(tl + dst + 16 * c0, pk);
        tmem_ld_wait();
        tmem_ld_x32(tl + src + 32 * (c0 + 2 * dc), ra);
        bias_relu_pack(rb, sb3 + src + 32 * (c0 + dc), pk);
        tmem_st_x16(tl + dst + 16 * (c0 + dc), pk);
        tmem_ld_wait();
        tmem_ld_x32(tl + src + 32 * (c0 + 3 * dc), rb);
        bias_relu_pack(ra, sb3 + src + 32 * (c0 + 2 * dc), pk);
        tmem_st_x16(tl + dst + 16 * (c0 + 2 * dc), pk);
        tmem_ld_wait();
        bias_relu_pack(rb, sb3 + src + 32 * (c0 + 3 * dc), pk);
        tmem_st_x16(tl + dst + 16 * (c0 + 3 * dc), pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_G1READY]);
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 2);

      // ---- E2a: first half of H1 -> Hs (overlaps the G2 MMAs)
      h1_convert(0, hp);
      if (!(ok = h1_store(0, hp))) break;
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 3);

      // ---- E3: acc_G2 (cols [64,192)) -> G2 = relu(. + b4) bf16 in place at [64+64hh, +32)
      if (!(ok = bwait(&bars[B_G2DONE], tp, ctx, 503))) break;
      tc_fence_after();
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 4);
      {
        uint32_t ra[32], rb[32], pk[16];
        const uint32_t col = 64u + 64u * hh;
        tmem_ld_x32(tl + col, ra);
        tmem_ld_wait();
        tmem_ld_x32(tl + col + 32, rb);
        bias_relu_pack(ra, sb4 + 64 * hh, pk);
        tmem_st_x16(tl + col, pk);
        tmem_ld_wait();
        bias_relu_pack(rb, sb4 + 64 * hh + 32, pk);
        tmem_st_x16(tl + col + 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_G2READY]);
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 5);

      // ---- E2b (register part): second half of H1, converted while the O MMAs run
      h1_convert(1, hp);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_R1FREE]);   // acc_A fully consumed: the next tile's add-on pass may start

      // ---- E4: acc_O -> O = |.| bf16 -> Os (pooling B operand, slot-in-N layout; other slot and invalid rows zero)
      if (!(ok = bwait(&bars[B_ODONE], tp, ctx, 504))) break;
      tc_fence_after();
      if (!(ok = bwait(&bars[B_OSEMPTY], tp ^ 1, ctx, 505))) break;
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 6);
      {
        uint32_t r[32];
        tmem_ld_x32(tl + 32 * hh, r);
        tmem_ld_wait();
        const int p0 = 32 * hh;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (p0 + 8 * g < PP) {
            uint32_t w4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              w4[j] = valid ? pack_bf16x2(fabsf(__uint_as_float(r[8 * g + 2 * j])), fabsf(__uint_as_float(r[8 * g + 2 * j + 1])))
                            : 0u;
            const int n_data = slot * PP + p0 + 8 * g, n_zero = (1 - slot) * PP + p0 + 8 * g;
            *reinterpret_cast<uint4*>(smem + SM_OS + off_mnmajor_nosw(n_data, tok, NPOOL)) =
                make_uint4(w4[0], w4[1], w4[2], w4[3]);
            *reinterpret_cast<uint4*>(smem + SM_OS + off_mnmajor_nosw(n_zero, tok, NPOOL)) = make_uint4(0, 0, 0, 0);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_OSREADY]);
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 7);

      // ---- E2b (store part)
      if (!(ok = h1_store(1, hp))) break;
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 8);

      // ---- E5: drain FEpartial^T (lane = d) into per-clip register accumulators; finished clips leave as bf16 hi/lo
      //      rows of the K2 operand images
      if (!(ok = bwait(&bars[B_FEDONE], tp, ctx, 506))) break;
      tc_fence_after();
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 9);
      {
        const int last_tok = min(tile * TILE_M + TILE_M - 1, ntok - 1);
        const int last_clip = last_tok / S;
        const bool boundary = last_clip > first_clip;
        const bool ends = ((last_tok + 1) % S) == 0;
        const int d = 128 * hh + tok;
        const uint32_t fe = tl + 128u * hh;
        uint32_t nb[PP];  // slot-1 partial = start of the next clip (only meaningful when `boundary`)
        {
          uint32_t a[PP];
#pragma unroll
          for (int g = 0; g < PP / 8; ++g) tmem_ld_x8(fe + 8 * g, *reinterpret_cast<uint32_t(*)[8]>(&a[8 * g]));
          if (boundary) {
#pragma unroll
            for (int g = 0; g < PP / 8; ++g) tmem_ld_x8(fe + PP + 8 * g, *reinterpret_cast<uint32_t(*)[8]>(&nb[8 * g]));
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < PP; ++j) facc[j] += __uint_as_float(a[j]);
        }
        // all TMEM reads of this tile are done: hand the accumulators back before the (slow) global flush
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_TMEMFREE]);
        if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 10);
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
          const bool flush = rep == 0 ? boundary : ends;
          if (!flush) continue;
          const int clip = c_begin + (rep == 0 ? first_clip : last_clip);
          // K2's A operand is MN-major (row = (clip,p) contiguous, k = d), so this thread's PP values for its d are
          // PP/8 16-byte chunks per image: rows [rowb, rowb+PP) of k-chunk image d/64, hi at +0 and lo at +64 KB
          const int tile2 = clip / p.cpt;
          const int rowb = (clip - tile2 * p.cpt) * PP;
          uint8_t* img = p.feimg + (size_t)tile2 * FE_TILE_BYTES + (size_t)(d >> 6) * 16384;
#pragma unroll
          for (int c = 0; c < PP / 8; ++c) {
            uint32_t hi4[4], lo4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float v0 = facc[8 * c + 2 * j], v1 = facc[8 * c + 2 * j + 1];
              const float h0 = round_bf16(v0), h1 = round_bf16(v1);
              hi4[j] = pack_bf16x2(h0, h1);
              lo4[j] = pack_bf16x2(v0 - h0, v1 - h1);
            }
            const uint32_t off = off_mnmajor_sw128(rowb + 8 * c, d & 63, 8192);
            *reinterpret_cast<uint4*>(img + off) = make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
            *reinterpret_cast<uint4*>(img + 65536 + off) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
          }
          if (rep == 0) {
#pragma unroll
            for (int j = 0; j < PP; ++j) facc[j] = __uint_as_float(nb[j]);
          } else {
#pragma unroll
            for (int j = 0; j < PP; ++j) facc[j] = 0.f;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// =================================================================================================
// K2: FE = FEpre W2^T + b2 Osum (tensor cores), cosine / similarity / logits / distance / push keys (fp32)
//     warps 0..7 epilogue (row = TMEM lane, two column halves), warp 8 MMA issuer, warp 9 bulk-copy loader
// =================================================================================================
namespace {
constexpr int K2_THREADS = 320;
constexpr uint32_t K2_STAGE = 65536;                 // A_hi chunk 16 KB | A_lo chunk 16 KB | W2 chunk 32 KB
constexpr uint32_t K2_SM_V = 2 * K2_STAGE;           // prototypes fp32 [P][257]
constexpr uint32_t K2_V_BYTES = PP_MAX * 257 * 4;    // 49344
constexpr uint32_t K2_SM_MISC = K2_SM_V + 49408;     // 180480
constexpr uint32_t K2_SMEM = K2_SM_MISC + 8192;

struct K2Params {
  const uint8_t* feimg; const float* osum; const uint8_t* packed; size_t off_w2, off_b2;
  const float* protos; const float* last_layer;
  float* logits; float* sim; float* dist; float* feats;
  const int64_t* labels; const int32_t* proto_class; long long global_offset; unsigned long long* best_key;
  int N, P, PP, K, cpt, ntiles;   // PP = padded P (row stride inside a tile), cpt = clips per tile = 128 / PP
  int* err;
};
}  // namespace

__global__ void __launch_bounds__(K2_THREADS, 1) proto_w2_kernel(const K2Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K2_SM_MISC);   // full[0..1] empty[2..3] accfull[4..5] accempty[6..7]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + K2_SM_MISC + 64);
  volatile int* abort_s = reinterpret_cast<volatile int*>(smem + K2_SM_MISC + 72);
  float* s_vn = reinterpret_cast<float*>(smem + K2_SM_MISC + 128);             // [64] clamped prototype norms
  float* s_b2 = reinterpret_cast<float*>(smem + K2_SM_MISC + 512);             // [256]
  float* s_part = reinterpret_cast<float*>(smem + K2_SM_MISC + 1536);          // [2][128][2] (dot, ff) of column half 1
  float* s_sim = reinterpret_cast<float*>(smem + K2_SM_MISC + 3584);           // [2][128]
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(smem + K2_SM_MISC + 4608);  // [64]
  float* s_v = reinterpret_cast<float*>(smem + K2_SM_V);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Ctx ctx{p.err, abort_s};
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0) atomicCAS(p.err, 0, 901);
    return;
  }
  if (tid == 0) {
    *abort_s = 0;
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 1); mbar_init(&bars[5], 1); mbar_init(&bars[6], 8); mbar_init(&bars[7], 8);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_s, 512);
  for (int i = tid; i < DD; i += K2_THREADS) s_b2[i] = reinterpret_cast<const float*>(p.packed + p.off_b2)[i];
  if (tid < p.P) s_key[tid] = PASN_KEY_NONE;
  {  // prototypes -> smem rows of 257 floats (conflict-free row-per-lane reads); 128-bit loads, several in flight
    const float4* src = reinterpret_cast<const float4*>(p.protos);
    const int n4 = p.P * (DD / 4);
#pragma unroll 4
    for (int i = tid; i < n4; i += K2_THREADS) {
      const float4 v = __ldg(src + i);
      float* dst = s_v + ((i * 4) >> 8) * 257 + ((i * 4) & 255);
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
  }
  __syncthreads();
  for (int pp = warp; pp < p.P; pp += K2_THREADS / 32) {
    float vv = 0.f;
    for (int d = lane; d < DD; d += 32) { const float b = s_v[pp * 257 + d]; vv = fmaf(b, b, vv); }
    vv = warp_sum(vv);
    if (lane == 0) s_vn[pp] = fmaxf(sqrtf(vv), 1e-8f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr_s;
  const uint32_t st_base = smem_u32(smem);
  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 9) {
    // ---------------------------------------------------------------- loader
    if (lane == 0) {
      griddep_wait();   // the pooled-vector images are written by the token kernel
      uint32_t u = 0;
      bool ok = true;
      for (int it = 0; it < my_tiles && ok; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const uint8_t* a_src = p.feimg + (size_t)tile * FE_TILE_BYTES;
        for (int kc = 0; kc < 4; ++kc, ++u) {
          const uint32_t s = u & 1, ph = (u >> 1) & 1;
          if (!(ok = bwait(&bars[2 + s], ph ^ 1, ctx, 611))) break;
          const uint32_t dst = st_base + s * K2_STAGE;
          mbar_arrive_expect_tx(&bars[s], K2_STAGE);
          bulk_g2s(dst, a_src + (size_t)kc * 16384, 16384, &bars[s]);
          bulk_g2s(dst + 16384, a_src + 65536 + (size_t)kc * 16384, 16384, &bars[s]);
          bulk_g2s(dst + 32768, p.packed + p.off_w2 + (size_t)kc * 32768, 16384, &bars[s]);
          bulk_g2s(dst + 49152, p.packed + p.off_w2 + (size_t)kc * 32768 + 16384, 16384, &bars[s]);
        }
      }
    }
  } else if (warp == 8) {
    // ---------------------------------------------------------------- MMA issuer (one elected thread, lean issue path:
    // see head_sm100_k1.cu -- descriptors as 32-bit halves in a single-thread region keep the UTCHMMAs back to back)
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, 256, 1, 0);  // A (pooled vectors) MN-major, B (W2) K-major
      constexpr uint32_t HI = desc_hi(1024, SWZ_128B);
      const uint32_t bar0 = smem_u32(bars);
      uint32_t u = 0;
      bool ok = true;
      for (int it = 0; it < my_tiles && ok; ++it) {
        const uint32_t buf = it & 1;
        if (!(ok = bwait(&bars[6 + buf], ((it >> 1) & 1) ^ 1, ctx, 621))) break;
        tc_fence_after();
        for (int kc = 0; kc < 4 && ok; ++kc, ++u) {
          const uint32_t s = u & 1, ph = (u >> 1) & 1;
          if (!(ok = bwait(&bars[s], ph, ctx, 622))) break;
          const uint32_t sb = st_base + s * K2_STAGE;
          const uint32_t ah = desc_lo(sb, 8192), al = desc_lo(sb + 16384, 8192), bd = desc_lo(sb + 32768, 16);
          const uint32_t d = tbase + 256u * buf;
          mma_ss_x(d, ah, HI, bd, HI, idesc, kc ? 1u : 0u);
          mma_ss_x(d, al, HI, bd, HI, idesc, 1u);
#pragma unroll
          for (int k4 = 1; k4 < 4; ++k4) {
            mma_ss_x(d, ah + k4 * 128, HI, bd + k4 * 2, HI, idesc, 1u);
            mma_ss_x(d, al + k4 * 128, HI, bd + k4 * 2, HI, idesc, 1u);
          }
          mma_commit_a(bar0 + 8u * (2 + s));
        }
        if (ok) mma_commit(&bars[4 + buf]);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: row r = 32q + lane, columns [128ch, +128)
    const int q = warp & 3, ch = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int r = q * 32 + lane;
    griddep_wait();   // Osum is written by the token kernel
    bool ok = true;
    for (int it = 0; it < my_tiles && ok; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const uint32_t buf = it & 1;
      const int clip0 = tile * p.cpt;
      int nclip = p.N - clip0;
      if (nclip > p.cpt) nclip = p.cpt;
      const int cl_r = r / p.PP, pp_r = r - cl_r * p.PP;          // row = clip-in-tile * PP + prototype
      const bool rvalid = cl_r < nclip && pp_r < p.P;
      const int cl = rvalid ? cl_r : 0, pp = rvalid ? pp_r : 0;
      const int n = clip0 + cl;
      const float os = rvalid ? p.osum[(size_t)n * p.P + pp] : 0.f;
      const float* vrow = s_v + pp * 257 + 128 * ch;
      const float* b2 = s_b2 + 128 * ch;
      float* frow = (p.feats && rvalid) ? p.feats + ((size_t)n * p.P + pp) * DD + 128 * ch : nullptr;
      if (!(ok = bwait(&bars[4 + buf], (it >> 1) & 1, ctx, 631))) break;
      tc_fence_after();
      const uint32_t ta = tbase + lane_base + 256u * buf + 128u * ch;
      float ff = 0.f, dot = 0.f;
      uint32_t ra[32], rb[32];
      tmem_ld_x32(ta, ra);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t* cur = (c & 1) ? rb : ra;
        uint32_t* nxt = (c & 1) ? ra : rb;
        if (c < 3) tmem_ld_x32(ta + 32 * (c + 1), *reinterpret_cast<uint32_t(*)[32]>(nxt));
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float f[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            f[j] = fmaf(b2[32 * c + 4 * j4 + j], os, __uint_as_float(cur[4 * j4 + j]));
            ff = fmaf(f[j], f[j], ff);
            dot = fmaf(f[j], vrow[32 * c + 4 * j4 + j], dot);
          }
          if (frow) *reinterpret_cast<float4*>(frow + 32 * c + 4 * j4) = make_float4(f[0], f[1], f[2], f[3]);
        }
        if (c < 3) tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[6 + buf]);   // accumulator buffer may be overwritten by tile it+2
      float* part = s_part + buf * 256;
      float* ssim = s_sim + buf * 128;
      if (ch == 1) { part[2 * r] = dot; part[2 * r + 1] = ff; }
      named_bar_sync(1, 256);
      if (ch == 0) {
        dot += part[2 * r];
        ff += part[2 * r + 1];
        const float nf = fmaxf(sqrtf(ff), 1e-8f);
        const float cosv = dot / (nf * s_vn[pp]);
        const float s = (cosv + 1.0f) / 2.0f;
        const float dd = 1.0f - s;
        ssim[r] = rvalid ? s : 0.f;
        if (rvalid) {
          p.sim[(size_t)n * p.P + pp] = s;
          if (p.dist) p.dist[(size_t)n * p.P + pp] = dd;
          if (p.best_key) {
            const int pc = p.proto_class[pp];
            if (pc < 0 || (long long)pc == p.labels[n]) atomicMin(&s_key[pp], pack_key(dd, (uint32_t)(p.global_offset + n)));
          }
        }
      }
      named_bar_sync(1, 256);
      if (tid < nclip * p.K) {
        const int c2 = tid / p.K, k = tid - c2 * p.K;
        float acc = 0.f;
        for (int qq = 0; qq < p.P; ++qq) acc = fmaf(ssim[c2 * p.PP + qq], p.last_layer[(size_t)k * p.P + qq], acc);
        p.logits[(size_t)(clip0 + c2) * p.K + k] = acc;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.best_key && tid < p.P && s_key[tid] != PASN_KEY_NONE) key_atomic_min_global(&p.best_key[tid], s_key[tid]);
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// =================================================================================================
// weight packing: fp32 state_dict tensors -> bf16 stage images in the exact smem byte order
// =================================================================================================
__global__ void pack_weights_kernel(pasn_weights w, int C, int P, uint8_t* out) {
  const PackedLayout PL = packed_layout(C);
  const int nkc = C / 64;
  const size_t n_l1 = (size_t)2 * nkc * 256 * 64, n_w4 = (size_t)4 * 128 * 64, n_w5 = (size_t)2 * 64 * 64;
  const size_t n_w2 = (size_t)4 * 256 * 64, n_bias = DD + DD + DH, n_b2 = DD;
  const size_t total = n_l1 + n_w4 + n_w5 + n_w2 + n_bias + n_b2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t j = i;
    if (j < n_l1) {  // stage = 2*kc + pass; pass 0: occ_w1 (W3), pass 1: addon_w1 (W1); image [256 rows][64 k]
      const int stage = (int)(j / (256 * 64)), e = (int)(j % (256 * 64));
      const int r = e / 64, k = e % 64, kc = stage >> 1, pass = stage & 1;
      const float v = (pass ? w.addon_w1 : w.occ_w1)[(size_t)r * C + kc * 64 + k];
      *reinterpret_cast<__nv_bfloat16*>(out + PL.off_l1 + (size_t)stage * 32768 + off_kmajor_sw128(r, k)) = __float2bfloat16_rn(v);
      continue;
    }
    j -= n_l1;
    if (j < n_w4) {  // occ_w2 [128][256]: 4 k-chunk images [128 rows][64 k]
      const int kc = (int)(j / (128 * 64)), e = (int)(j % (128 * 64)), r = e / 64, k = e % 64;
      *reinterpret_cast<__nv_bfloat16*>(out + PL.off_w4 + (size_t)kc * 16384 + off_kmajor_sw128(r, k)) =
          __float2bfloat16_rn(w.occ_w2[(size_t)r * DD + kc * 64 + k]);
      continue;
    }
    j -= n_w4;
    if (j < n_w5) {  // occ_w3 [P][128] zero-padded to 64 rows: 2 k-chunk images [64 rows][64 k]
      const int kc = (int)(j / (64 * 64)), e = (int)(j % (64 * 64)), r = e / 64, k = e % 64;
      const float v = r < P ? w.occ_w3[(size_t)r * DH + kc * 64 + k] : 0.f;
      *reinterpret_cast<__nv_bfloat16*>(out + PL.off_w5 + (size_t)kc * 8192 + off_kmajor_sw128(r, k)) = __float2bfloat16_rn(v);
      continue;
    }
    j -= n_w5;
    if (j < n_w2) {  // addon_w2 [256][256]: 4 k-chunk images [256 rows][64 k]
      const int kc = (int)(j / (256 * 64)), e = (int)(j % (256 * 64)), r = e / 64, k = e % 64;
      *reinterpret_cast<__nv_bfloat16*>(out + PL.off_w2 + (size_t)kc * 32768 + off_kmajor_sw128(r, k)) =
          __float2bfloat16_rn(w.addon_w2[(size_t)r * DD + kc * 64 + k]);
      continue;
    }
    j -= n_w2;
    if (j < n_bias) {  // b3 | b1 | b4, bf16-rounded values kept as fp32
      float v;
      if (j < DD) v = w.occ_b1[j];
      else if (j < 2 * DD) v = w.addon_b1[j - DD];
      else v = w.occ_b2[j - 2 * DD];
      reinterpret_cast<float*>(out + PL.off_bias)[j] = round_bf16(v);
      continue;
    }
    j -= n_bias;
    reinterpret_cast<float*>(out + PL.off_b2)[j] = round_bf16(w.addon_b2[j]);
  }
}

// =================================================================================================
// host side
// =================================================================================================
static long long* g_trace = nullptr;
void sm100_set_trace(void* dev_buf) { g_trace = reinterpret_cast<long long*>(dev_buf); }
static int g_k1_variant = -1;   // -1: PASN_K1_PHASES / PASN_K1_PAIR from the environment, else the default
void sm100_set_k1_variant(int variant) { g_k1_variant = variant; }

bool sm100_supported(const pasn_dims& d) {
  // fp32 feature maps take the fused path only on explicit request (bf16 compute: inputs are rounded on the fly)
  // channels_last ([N,S,C]) feature maps: bf16 only
  if (d.layout != PASN_LAYOUT_NCS && !(d.layout == PASN_LAYOUT_NSC && d.dtype == PASN_BF16)) return false;
  if (d.dtype != PASN_BF16 && !(d.dtype == PASN_F32 && d.path == PASN_PATH_TCGEN05)) return false;
  if (d.D != DD) return false;
  if (d.C % 64 != 0 || d.C < 64 || d.C > 1024) return false;
  if (d.P < 1 || d.P > PP_MAX) return false;
  if (d.S % 4 != 0 || d.S < TILE_M) return false;   // a 128-voxel tile may touch at most two clips
  return true;
}

size_t sm100_packed_bytes(const pasn_dims& d) { return packed_layout(d.C).total; }

static inline int k2_ppad(const pasn_dims& d) {
  const int p8 = (d.P + 7) / 8 * 8;
  return p8 <= 16 ? 16 : p8 <= 32 ? 32 : p8 <= 40 ? 40 : 48;   // matches the head_tokens_kernel<PP> instantiations
}
// Small batches: one persistent CTA per clip range leaves most SMs idle when N < 148.  Pooling is linear in the voxels,
// so a clip of S voxels can be processed as G independent "virtual clips" of S/G voxels (each at least one 128-voxel
// tile, multiple of 4) whose pooled features are summed after K2.  G = 1 means no split.
static inline int split_groups(const pasn_dims& d) {
  if (d.N <= 0 || d.N > 74) return 1;
  static const int allow = [] { const char* e = getenv("PASN_NO_SPLIT"); return (e && atoi(e) != 0) ? 0 : 1; }();
  if (!allow) return 1;
  int best = 1;
  for (int g = 2; g <= d.S / TILE_M; ++g) {
    if (d.S % g != 0) continue;
    const int sv = d.S / g;
    if (sv < TILE_M || sv % 4 != 0) continue;
    if ((long long)d.N * g > 148) break;         // one virtual clip range per SM is enough; beyond that the extra
                                                  // kernels of the split cost more than the parallelism gains
    best = g;
  }
  return best;
}
struct WsLayout {
  int G, Nv, Sv, tiles2;
  size_t off_osum, off_err, off_featsv, off_fe, off_simv, off_logv, total;
};
static inline WsLayout ws_layout(const pasn_dims& d) {
  WsLayout L;
  L.G = split_groups(d);
  L.Nv = d.N * L.G;
  L.Sv = d.S / L.G;
  L.tiles2 = ceil_div(L.Nv, TILE_M / k2_ppad(d));
  size_t o = (size_t)L.tiles2 * FE_TILE_BYTES;
  L.off_osum = o; o += align_up((size_t)L.Nv * d.P * 4, 256);
  L.off_err = o; o += 256;
  L.off_featsv = L.off_fe = L.off_simv = L.off_logv = o;
  if (L.G > 1) {
    L.off_featsv = o; o += align_up((size_t)L.Nv * d.P * DD * 4, 256);   // K2's features per virtual clip
    L.off_fe = o; o += align_up((size_t)d.N * d.P * DD * 4, 256);        // their sum per real clip
    L.off_simv = o; o += align_up((size_t)L.Nv * d.P * 4, 256);          // K2's per-virtual-clip similarity / logits: unused
    L.off_logv = o; o += align_up((size_t)L.Nv * d.K * 4, 256);
  }
  L.total = o;
  return L;
}

// workspace: K2 operand images [tiles2][128 KB] | Osum [Nv][P] fp32 | err int | (split only) scratch, see ws_layout
size_t sm100_workspace_bytes(const pasn_dims& d) { return ws_layout(d).total; }

// FE[n][i] = sum_g FEv[n*G + g][i]   (i over P*D)
__global__ void sum_groups_kernel(const float* __restrict__ fev, float* __restrict__ fe, int N, int G, int PD) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * PD) return;
  const long long n = i / PD, r = i - n * PD;
  float a = 0.f;
  for (int g = 0; g < G; ++g) a += fev[((size_t)n * G + g) * PD + r];
  fe[i] = a;
}

int sm100_pack_weights(const pasn_weights& w, const pasn_dims& d, void* packed, cudaStream_t st) {
  if (((uintptr_t)packed & 15) != 0) return PASN_ERR_ALIGN;
  pack_weights_kernel<<<148 * 4, 256, 0, st>>>(w, d.C, d.P, reinterpret_cast<uint8_t*>(packed));
  PASN_LAUNCH_CHECK();
  count_launch();
  return PASN_OK;
}

template <int PP>
static int launch_k1(const K1Params& k1, int grid, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(head_tokens_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM) != cudaSuccess)
      return PASN_ERR_CUDA;
    attr_done = true;
  }
  head_tokens_kernel<PP><<<grid, K1_THREADS, K1_SMEM, st>>>(k1);
  PASN_LAUNCH_CHECK();
  return PASN_OK;
}

int sm100_head_forward(const void* feat, const pasn_weights& w, const void* packed, const pasn_dims& d, float* logits,
                       float* sim, void* occ, float* feats, float* dist, const pasn_push_args* push, void* ws,
                       size_t ws_bytes, cudaStream_t st) {
  if (!sm100_supported(d)) return PASN_ERR_UNSUPPORTED;
  if (ws_bytes < sm100_workspace_bytes(d)) return PASN_ERR_WORKSPACE;
  if (((uintptr_t)feat & 15) != 0 || ((uintptr_t)packed & 15) != 0 || ((uintptr_t)ws & 15) != 0) return PASN_ERR_ALIGN;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(proto_w2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K2_SMEM) != cudaSuccess)
      return PASN_ERR_CUDA;
    attr_done = true;
  }
  const WsLayout L = ws_layout(d);
  char* wsp = reinterpret_cast<char*>(ws);
  uint8_t* feimg = reinterpret_cast<uint8_t*>(wsp);
  float* osum = reinterpret_cast<float*>(wsp + L.off_osum);
  int* err = reinterpret_cast<int*>(wsp + L.off_err);
  if (cudaMemsetAsync(err, 0, 4, st) != cudaSuccess) return PASN_ERR_CUDA;

  const int num_sms = 148;
  K1Params k1{};
  k1.f32_in = d.dtype == PASN_F32;
  k1.nsc = d.layout == PASN_LAYOUT_NSC;
  k1.feat = reinterpret_cast<const __nv_bfloat16*>(feat);
  k1.feat32 = reinterpret_cast<const float*>(feat);
  k1.packed = reinterpret_cast<const uint8_t*>(packed);
  k1.occ = k1.f32_in ? nullptr : reinterpret_cast<__nv_bfloat16*>(occ);
  k1.occ32 = k1.f32_in ? reinterpret_cast<float*>(occ) : nullptr;
  k1.feimg = feimg; k1.osum = osum;
  k1.N = L.Nv; k1.C = d.C; k1.P = d.P; k1.S = L.Sv; k1.nkc = d.C / 64;
  k1.G = L.G; k1.SR = d.S;
  k1.clips_per_cta = ceil_div(L.Nv, num_sms);
  k1.cpt = TILE_M / k2_ppad(d);
  k1.err = err;
  k1.trace = g_trace;
  { const char* e = getenv("PASN_DBG_SKIP"); k1.dbg_skip = e ? atoi(e) : 0; }
  const int grid1 = ceil_div(L.Nv, k1.clips_per_cta);
  const int ppad = (d.P + 7) / 8 * 8;
  // Token-kernel variants (same results; see profiles/README.md for the measurements):
  //   1  head_sm100_k1.cu, serial tile order -- the default
  //   2  head_sm100_k1.cu, two-phase order (G / A phases overlapped with the previous tile's chain)
  //   0  first-generation kernel in this file
  //   3  CTA-pair (cta_group::2) variant, head_sm100_pair.cu
  static const int env_variant = [] {
    const char* e = getenv("PASN_K1_PAIR");
    if (e && atoi(e) != 0) return 3;
    e = getenv("PASN_K1_PHASES");
    return e ? atoi(e) : 1;
  }();
  int variant = g_k1_variant >= 0 ? g_k1_variant : env_variant;
  if ((k1.nsc || L.G > 1) && (variant == 0 || variant == 3)) variant = 1;   // channels_last input, voxel-group split: current kernel only
  const bool use_pair = variant == 3 && !k1.f32_in;
  const int phases = variant == 3 ? 1 : variant;
  k1.phases = phases;
  main_kernel_begin(st);
  int rc;
  if (use_pair) rc = launch_k1_pair(k1, ppad, st);
  else if (phases != 0) rc = launch_k1_two_phase(k1, ppad, grid1, st);
  else if (ppad <= 16) rc = launch_k1<16>(k1, grid1, st);
  else if (ppad <= 32) rc = launch_k1<32>(k1, grid1, st);
  else if (ppad <= 40) rc = launch_k1<40>(k1, grid1, st);
  else rc = launch_k1<48>(k1, grid1, st);
  if (rc) return rc;
  main_kernel_end(st);
  count_launch();

  const PackedLayout PL = packed_layout(d.C);
  K2Params k2{};
  k2.feimg = feimg; k2.osum = osum; k2.packed = k1.packed; k2.off_w2 = PL.off_w2; k2.off_b2 = PL.off_b2;
  k2.protos = w.prototypes; k2.last_layer = w.last_layer;
  const bool split = L.G > 1;
  float* featsv = reinterpret_cast<float*>(wsp + L.off_featsv);
  k2.logits = split ? reinterpret_cast<float*>(wsp + L.off_logv) : logits;
  k2.sim = split ? reinterpret_cast<float*>(wsp + L.off_simv) : sim;
  k2.dist = split ? nullptr : dist;
  k2.feats = split ? featsv : feats;
  k2.labels = (push && !split) ? push->labels : nullptr;
  k2.proto_class = (push && !split) ? push->proto_class : nullptr;
  k2.global_offset = (push && !split) ? (long long)push->global_offset : 0;
  k2.best_key = (push && !split) ? reinterpret_cast<unsigned long long*>(push->best_key) : nullptr;
  k2.N = L.Nv; k2.P = d.P; k2.K = d.K;
  k2.cpt = TILE_M / k2_ppad(d);
  k2.PP = k2_ppad(d);
  k2.ntiles = L.tiles2;
  k2.err = err;
  const int grid2 = k2.ntiles < num_sms ? k2.ntiles : num_sms;
  {  // programmatic dependent launch: K2's prologue (TMEM, barriers, norms, the resident W2 images) overlaps K1's tail
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid2); cfg.blockDim = dim3(K2_THREADS); cfg.dynamicSmemBytes = K2_SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, proto_w2_kernel, k2) != cudaSuccess) return PASN_ERR_CUDA;
  }
  PASN_LAUNCH_CHECK();
  count_launch();
  if (split) {   // sum the pooled features of the voxel groups, then the fp32 cosine / logits / push stage on real clips
    float* fe = feats ? feats : reinterpret_cast<float*>(wsp + L.off_fe);
    const long long tot = (long long)d.N * d.P * DD;
    sum_groups_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(featsv, fe, d.N, L.G, d.P * DD);
    PASN_LAUNCH_CHECK();
    count_launch();
    const int rc2 = launch_proto_stage(fe, w.prototypes, w.last_layer, d.N, d.P, DD, d.K, logits, sim, dist, push, st);
    if (rc2) return rc2;
  }
  return PASN_OK;
}

// surfaced for tests / debugging: non-zero if a kernel hit its bounded-wait limit (protocol bug) on the last call
int sm100_last_error(const void* ws, const pasn_dims& d, cudaStream_t st) {
  const char* wsp = reinterpret_cast<const char*>(ws);
  const int* err = reinterpret_cast<const int*>(wsp + ws_layout(d).off_err);
  int h = 0;
  if (cudaMemcpyAsync(&h, err, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
  return h;
}

}  // namespace pasn
