// CTA-pair (cta_group::2) variant of the fused token kernel.
//
// Why: with one CTA per SM every SM has to pull the full 592 KB weight stream per 128-voxel tile from L2 (72 B/clk at
// MMA speed) on top of its feature tile; measured, the layer-1 MMAs then run at ~60 % of their rate (they run at full
// rate as soon as either stream is switched off).  In a CTA pair one thread of the leader CTA issues M=256 MMAs that
// span both SMs: each CTA keeps its own 128-voxel tile (A operand, TMEM accumulators, epilogue) but supplies only
// HALF of every weight tile (B operand rows [N/2*rank, +N/2)), so the weight bytes entering each SM are halved.
//
// Same math, same outputs, same packed-weight / workspace formats as head_tokens_kernel (head_sm100.cu); differences:
//   * all tcgen05 alloc/mma/commit use cta_group::2; MMA->consumer barriers are signalled in both CTAs by one
//     multicast commit; consumer->MMA barriers live in the leader CTA and are arrived on remotely (mapa)
//   * weight stages are half size (16 KB): 4-slot ring; the peer's warp 0 relays "my half landed" to the leader
//   * pooling: the two CTAs hold different voxels (different K), so the pair MMA is made block-diagonal by widening N:
//     B = [Os(cta0) | Os(cta1)], each CTA reads back only its own NPOOL columns; the other block is ignored
//   * a CTA whose clip range is shorter than its partner's runs "ghost" tiles (all voxels invalid, no stores)
#include <cstdlib>

#include "head_sm100_shared.cuh"

namespace pasn {
using namespace sm100;
using namespace k1;

namespace {

constexpr int XSLOTS = 4, WSLOTS = 4;
constexpr uint32_t XSLOT_BYTES = 16384, WSLOT_BYTES = 16384, HS_BYTES = 32768;
constexpr int KP_WARPS = 16;
constexpr int EPI_WARP0 = 6;   // warps 6..13 epilogue (TMEM quadrant = warp % 4), 14 Osum, 15 occurrence-map store
constexpr int KP_THREADS = KP_WARPS * 32;

constexpr uint32_t SM_X = 0;
constexpr uint32_t SM_W = SM_X + XSLOTS * XSLOT_BYTES;            // 65536
constexpr uint32_t SM_HS = SM_W + WSLOTS * WSLOT_BYTES;           // 131072
constexpr uint32_t SM_OS = SM_HS + HS_BYTES;                      // 163840
constexpr uint32_t OS_BYTES_MAX = TILE_M * 2 * PP_MAX * 2;        // 24576
constexpr uint32_t SM_BIAS = SM_OS + OS_BYTES_MAX;                // 188416
constexpr uint32_t SM_BAR = SM_BIAS + (DD + DD + DH) * 4;         // 190976
constexpr uint32_t SM_MISC = SM_BAR + 32 * 8;                     // 191232
constexpr uint32_t KP_SMEM = SM_MISC + 64;                        // 191296  (< 196 KB carve-out)

enum {
  B_XFULL = 0, B_XEMPTY = 4, B_WFULL = 8, B_WEMPTY = 12, B_L1DONE = 16, B_G1READY, B_G2DONE, B_G2READY, B_ODONE,
  B_OSREADY, B_OSEMPTY, B_HSREADY, B_HSEMPTY, B_FEDONE, B_TMEMFREE, B_COUNT
};
static_assert(B_COUNT <= 32, "barrier table");

}  // namespace

template <int PP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(KP_THREADS, 1) head_tokens_pair_kernel(const K1Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + SM_MISC);
  volatile int* abort_s = reinterpret_cast<volatile int*>(smem + SM_MISC + 8);
  float* sb3 = reinterpret_cast<float*>(smem + SM_BIAS);
  float* sb1 = sb3 + DD;
  float* sb4 = sb1 + DD;

  constexpr int NPOOL = 2 * PP;        // columns of one CTA's pooling block
  constexpr int NPAIR = 2 * NPOOL;     // N of the block-diagonal pair MMA
  static_assert(NPAIR <= 256 && NPAIR % 16 == 0, "pooling MMA shape");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int S = p.S;
  auto clips_of = [&](int b) { int n = p.N - b * p.clips_per_cta; return n < 0 ? 0 : (n > p.clips_per_cta ? p.clips_per_cta : n); };
  const int c_begin = blockIdx.x * p.clips_per_cta;
  const int ncl = clips_of(blockIdx.x);
  const int ntok = ncl * S;
  const int ntok_peer = clips_of(blockIdx.x ^ 1) * S;
  const int ntiles = (max(ntok, ntok_peer) + TILE_M - 1) / TILE_M;   // both CTAs of a pair run the same tile count
  const PackedLayout PL = packed_layout(p.C);
  Ctx ctx{p.err, abort_s};
  const int nkc = p.nkc;
  const int stages_per_tile = 2 * nkc + 3;

  if (tid == 0) {
    *abort_s = 0;
    if ((smem_u32(smem) & 1023u) != 0) atomicCAS(p.err, 0, 900);
    for (int i = 0; i < XSLOTS; ++i) { mbar_init(&bars[B_XFULL + i], 8); mbar_init(&bars[B_XEMPTY + i], 1); }
    for (int i = 0; i < WSLOTS; ++i) { mbar_init(&bars[B_WFULL + i], rank == 0 ? 2 : 1); mbar_init(&bars[B_WEMPTY + i], 1); }
    mbar_init(&bars[B_L1DONE], 1);
    mbar_init(&bars[B_G1READY], 16);
    mbar_init(&bars[B_G2DONE], 1);
    mbar_init(&bars[B_G2READY], 16);
    mbar_init(&bars[B_ODONE], 1);
    mbar_init(&bars[B_OSREADY], 16);   // leader copy: MMA issuer; each CTA's Osum / store warps use B_ODONE-ordered local data
    mbar_init(&bars[B_OSEMPTY], 3);
    mbar_init(&bars[B_HSREADY], 16);
    mbar_init(&bars[B_HSEMPTY], 1);
    mbar_init(&bars[B_FEDONE], 1);
    mbar_init(&bars[B_TMEMFREE], 16);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc2(tmem_ptr_s, 512);
  {
    const float* gb = reinterpret_cast<const float*>(p.packed + PL.off_bias);
    for (int i = tid; i < DD + DD + DH; i += KP_THREADS) sb3[i] = gb[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  griddep_launch_dependents();   // K2 may start launching (its CTAs only fit on an SM once one of ours has exited)
  const uint32_t tbase = *tmem_ptr_s;
  const uint32_t x_base = smem_u32(smem + SM_X), w_base = smem_u32(smem + SM_W);
  const uint32_t hs_base = smem_u32(smem + SM_HS), os_base = smem_u32(smem + SM_OS);
  // cluster address of barrier i in the leader CTA
  auto leader_bar = [&](int i) -> uint32_t { return mapa_u32(smem_u32(&bars[i]), 0); };

  if (warp == 0) {
    if (lane == 0 && rank == 0) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA, one thread)
      const uint32_t idesc_l1 = make_idesc_bf16(256, 256, 1, 0);
      const uint32_t idesc_g2 = make_idesc_bf16(256, 128, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(256, 64, 0, 0);
      const uint32_t idesc_pool = make_idesc_bf16(256, NPAIR, 1, 1);
      constexpr uint32_t lbo_os = (uint32_t)(NPOOL / 8) * 128u;
      uint32_t chunk = 0, wst = 0;
      bool ok = true;
      for (int tile = 0; tile < ntiles && ok; ++tile) {
        const uint32_t tp = tile & 1;
        K1_TRACE(0, tile, 0);
        if (!(ok = bwait(&bars[B_TMEMFREE], tp ^ 1, ctx, 101))) break;
        tc_fence_after();
        K1_TRACE(0, tile, 1);
        for (int kc = 0; kc < nkc && ok; ++kc, ++chunk) {
          const uint32_t xs = chunk % XSLOTS, xph = (chunk / XSLOTS) & 1;
          if (!(ok = bwait(&bars[B_XFULL + xs], xph, ctx, 102))) break;
          for (int pass = 0; pass < 2 && ok; ++pass, ++wst) {
            const uint32_t ws = wst % WSLOTS, wph = (wst / WSLOTS) & 1;
            if (!(ok = bwait(&bars[B_WFULL + ws], wph, ctx, 103))) break;
            tc_fence_after();
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t ad = make_smem_desc(x_base + xs * XSLOT_BYTES + k4 * 2048, 8192, 1024, SWZ_128B);
              const uint64_t bd = make_smem_desc(w_base + ws * WSLOT_BYTES + k4 * 32, 16, 1024, SWZ_128B);
              mma_ss2(tbase + (pass ? 256u : 0u), ad, bd, idesc_l1, (kc | k4) ? 1u : 0u);
            }
            mma_commit2(&bars[B_WEMPTY + ws], 3);
          }
          mma_commit2(&bars[B_XEMPTY + xs], 3);
        }
        if (!ok) break;
        mma_commit2(&bars[B_L1DONE], 3);
        K1_TRACE(0, tile, 2);
        // G2 = G1 W4^T (A from TMEM, N = 128: 64 weight rows per CTA, two k-chunks per 16 KB stage)
        if (!(ok = bwait(&bars[B_G1READY], tp, ctx, 104))) break;
        tc_fence_after();
        K1_TRACE(0, tile, 3);
        for (int st = 0; st < 2 && ok; ++st, ++wst) {
          const uint32_t ws = wst % WSLOTS, wph = (wst / WSLOTS) & 1;
          if (!(ok = bwait(&bars[B_WFULL + ws], wph, ctx, 105))) break;
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int ks = st * 8 + kk;
            const uint64_t bd = make_smem_desc(w_base + ws * WSLOT_BYTES + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024, SWZ_128B);
            const uint32_t a_col = ks < 8 ? 8u * ks : 192u + 8u * (ks - 8);
            mma_ts2(tbase + 64u, tbase + a_col, bd, idesc_g2, ks ? 1u : 0u);
          }
          mma_commit2(&bars[B_WEMPTY + ws], 3);
        }
        if (!ok) break;
        mma_commit2(&bars[B_G2DONE], 3);
        K1_TRACE(0, tile, 4);
        // O = G2 W5^T (A from TMEM, N = 64: 32 weight rows per CTA)
        if (!(ok = bwait(&bars[B_G2READY], tp, ctx, 106))) break;
        K1_TRACE(0, tile, 5);
        {
          const uint32_t ws = wst % WSLOTS, wph = (wst / WSLOTS) & 1;
          if (!(ok = bwait(&bars[B_WFULL + ws], wph, ctx, 107))) break;
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t a_col = ks < 4 ? 64u + 8u * ks : 128u + 8u * (ks - 4);
            const uint64_t bd = make_smem_desc(w_base + ws * WSLOT_BYTES + (ks >> 2) * 4096 + (ks & 3) * 32, 16, 1024, SWZ_128B);
            mma_ts2(tbase + 0u, tbase + a_col, bd, idesc_o, ks ? 1u : 0u);
          }
          mma_commit2(&bars[B_WEMPTY + ws], 3);
          ++wst;
        }
        mma_commit2(&bars[B_ODONE], 3);
        K1_TRACE(0, tile, 6);
        // pooling, block-diagonal over the pair: D[d, (cta, slot, p)] ; halves at cols [0,NPAIR) and [256,256+NPAIR)
        if (!(ok = bwait(&bars[B_OSREADY], tp, ctx, 108))) break;
        K1_TRACE(0, tile, 7);
        for (int half = 0; half < 2 && ok; ++half) {
          if (!(ok = bwait(&bars[B_HSREADY], (uint32_t)half, ctx, 109))) break;
          tc_fence_after();
          K1_TRACE(0, tile, 8 + half);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ad = make_smem_desc(hs_base + ks * 4096, 2048, 128, SWZ_NONE);
            const uint64_t bd = make_smem_desc(os_base + ks * 2 * lbo_os, lbo_os, 128, SWZ_NONE);
            mma_ss2(tbase + (half ? 256u : 0u), ad, bd, idesc_pool, ks ? 1u : 0u);
          }
          mma_commit2(&bars[B_HSEMPTY], 3);
        }
        if (!ok) break;
        mma_commit2(&bars[B_OSEMPTY], 3);
        mma_commit2(&bars[B_FEDONE], 3);
        K1_TRACE(0, tile, 10);
      }
    } else if (lane == 0 && rank == 1) {
      // ---------------------------------------------------------------- peer: relay "my weight half landed" to the leader
      uint32_t wst = 0;
      bool ok = true;
      const int total = ntiles * stages_per_tile;
      for (int i = 0; i < total && ok; ++i, ++wst) {
        const uint32_t ws = wst % WSLOTS, wph = (wst / WSLOTS) & 1;
        if (!(ok = bwait(&bars[B_WFULL + ws], wph, ctx, 151))) break;
        mbar_arrive_cluster(leader_bar(B_WFULL + ws));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ weight producer: this CTA's half of every stage
    if (lane == 0) {
      uint32_t wst = 0;
      bool ok = true;
      for (int tile = 0; tile < ntiles && ok; ++tile) {
        for (int i = 0; i < stages_per_tile; ++i, ++wst) {
          const uint32_t ws = wst % WSLOTS, wph = (wst / WSLOTS) & 1;
          if (!(ok = bwait(&bars[B_WEMPTY + ws], wph ^ 1, ctx, 201))) break;
          const uint32_t dst = w_base + ws * WSLOT_BYTES;
          if (i < 2 * nkc) {            // layer-1 stage [256 rows x 64 k]: rows [128*rank, +128)
            mbar_arrive_expect_tx(&bars[B_WFULL + ws], 16384);
            bulk_g2s(dst, p.packed + PL.off_l1 + (size_t)i * 32768 + rank * 16384, 16384, &bars[B_WFULL + ws]);
          } else if (i < 2 * nkc + 2) {  // W4: two k-chunk images [128 rows x 64 k] per stage: rows [64*rank, +64) of each
            const int j = i - 2 * nkc;
            mbar_arrive_expect_tx(&bars[B_WFULL + ws], 16384);
            for (int kcc = 0; kcc < 2; ++kcc)
              bulk_g2s(dst + kcc * 8192, p.packed + PL.off_w4 + (size_t)(2 * j + kcc) * 16384 + rank * 8192, 8192, &bars[B_WFULL + ws]);
          } else {                       // W5: two k-chunk images [64 rows x 64 k]: rows [32*rank, +32) of each
            mbar_arrive_expect_tx(&bars[B_WFULL + ws], 8192);
            for (int kc5 = 0; kc5 < 2; ++kc5)
              bulk_g2s(dst + kc5 * 4096, p.packed + PL.off_w5 + (size_t)kc5 * 8192 + rank * 4096, 4096, &bars[B_WFULL + ws]);
          }
        }
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ------------------------------------------------------------------ X producers (see head_sm100.cu)
    const int xw = warp - 2;
    const uint32_t nchunks = (uint32_t)(ntiles * nkc);
    bool ok = true;
    uint2 va[16], vb[16];
    auto load_unit = [&](uint32_t g, uint2* v) {
      const int tile = (int)(g / (uint32_t)nkc), kc = (int)(g - (uint32_t)tile * nkc);
      const int t = tile * TILE_M + 4 * lane;
      const bool valid = t < ntok;
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
      if (lane == 0) mbar_arrive_cluster(leader_bar(B_XFULL + xs));
      return true;
    };
    if (nchunks > 0) load_unit(0, va);
    for (uint32_t g = 0; g < nchunks && ok; g += 2) {
      if (g + 1 < nchunks) load_unit(g + 1, vb);
      if (!(ok = store_unit(g, va))) break;
      if (g + 1 < nchunks) {
        if (g + 2 < nchunks) load_unit(g + 2, va);
        ok = store_unit(g + 1, vb);
      }
    }
  } else if (warp == 14) {
    // ------------------------------------------------------------------ occurrence column sums (own Os is complete once
    // this CTA's epilogue warps are done with E4; the leader-side B_OSREADY cannot be waited on from the peer, so the
    // local hand-off uses named barrier 1: 8 epilogue warps arrive, Osum and store warps sync)
    float acc0 = 0.f, acc1 = 0.f;
    bool ok = true;
    const unsigned char* os = smem + SM_OS;
    for (int tile = 0; tile < ntiles && ok; ++tile) {
      named_bar_sync(1, 10 * 32);
      float s0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f};
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
      if (tile * TILE_M < ntok) {
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
    }
  } else if (warp == 15) {
    // ------------------------------------------------------------------ occurrence-map store: Os (smem) -> [N][P][S] bf16
    const unsigned char* os = smem + SM_OS;
    for (int tile = 0; tile < ntiles; ++tile) {
      named_bar_sync(1, 10 * 32);
      if (p.occ != nullptr) {
        const int first_clip = (tile * TILE_M) / S;
#pragma unroll 1
        for (int grp = 0; grp < 4; ++grp) {
          const int tok = grp * 32 + lane;
          const int t = tile * TILE_M + tok;
          if (t < ntok) {
            const int clipl = t / S, s = t - clipl * S, slot = clipl - first_clip;
            __nv_bfloat16* orow = p.occ + ((size_t)(c_begin + clipl) * p.P) * S + s;
            const unsigned char* src = os + off_mnmajor_nosw(slot * PP, tok, NPOOL);
#pragma unroll 8
            for (int pp = 0; pp < p.P; ++pp)
              orow[(size_t)pp * S] = *reinterpret_cast<const __nv_bfloat16*>(src + (pp >> 3) * 128 + (pp & 7) * 2);
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
    auto h1_store = [&](int hf, const uint32_t* hp) -> bool {
      if (!bwait(&bars[B_HSEMPTY], (uint32_t)(hf ^ 1), ctx, 502)) return false;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(smem + SM_HS + off_mnmajor_nosw(64 * hh + 8 * g, tok, 128)) =
            make_uint4(hp[4 * g], hp[4 * g + 1], hp[4 * g + 2], hp[4 * g + 3]);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_bar(B_HSREADY));
      return true;
    };

    for (int tile = 0; tile < ntiles && ok; ++tile) {
      const uint32_t tp = tile & 1;
      const int t = tile * TILE_M + tok;
      const bool valid = t < ntok;
      const bool tile_live = tile * TILE_M < ntok;
      const int first_clip = (tile * TILE_M) / S;
      const int clipl = valid ? t / S : first_clip;
      const int slot = valid ? clipl - first_clip : 0;
      uint32_t hp[32];

      // E1: acc_G -> G1 (bf16, in place; see head_sm100.cu)
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 0);
      if (!(ok = bwait(&bars[B_L1DONE], tp, ctx, 501))) break;
      tc_fence_after();
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 1);
      {
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
      if (lane == 0) mbar_arrive_cluster(leader_bar(B_G1READY));
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 2);

      // E2a
      h1_convert(0, hp);
      if (!(ok = h1_store(0, hp))) break;
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 3);

      // E3: acc_G2 -> G2
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
      if (lane == 0) mbar_arrive_cluster(leader_bar(B_G2READY));
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 5);

      // E2b (register part)
      h1_convert(1, hp);

      // E4: acc_O -> Os
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
      if (lane == 0) mbar_arrive_cluster(leader_bar(B_OSREADY));
      named_bar_arrive(1, 10 * 32);   // local hand-off of Os to the Osum / store warps
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 7);

      // E2b (store part)
      if (!(ok = h1_store(1, hp))) break;
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 8);

      // E5: this CTA's block of FEpartial^T: columns [rank*NPOOL, +NPOOL) of half hh (cols 0.. / 256..)
      if (!(ok = bwait(&bars[B_FEDONE], tp, ctx, 506))) break;
      tc_fence_after();
      if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 9);
      {
        const int last_tok = min(tile * TILE_M + TILE_M - 1, ntok - 1);
        const int last_clip = tile_live ? last_tok / S : 0;
        const bool boundary = tile_live && last_clip > first_clip;
        const bool ends = tile_live && ((last_tok + 1) % S) == 0;
        const int d = 128 * hh + tok;
        const uint32_t fe = tl + 256u * hh + rank * NPOOL;
        uint32_t nb[PP];
        {
          uint32_t a[PP];
#pragma unroll
          for (int g = 0; g < PP / 8; ++g) tmem_ld_x8(fe + 8 * g, *reinterpret_cast<uint32_t(*)[8]>(&a[8 * g]));
          if (boundary) {
#pragma unroll
            for (int g = 0; g < PP / 8; ++g) tmem_ld_x8(fe + PP + 8 * g, *reinterpret_cast<uint32_t(*)[8]>(&nb[8 * g]));
          }
          tmem_ld_wait();
          if (tile_live) {
#pragma unroll
            for (int j = 0; j < PP; ++j) facc[j] += __uint_as_float(a[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_bar(B_TMEMFREE));
        if (warp == EPI_WARP0 && lane == 0) K1_TRACE(1, tile, 10);
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
          const bool flush = rep == 0 ? boundary : ends;
          if (!flush) continue;
          const int clip = c_begin + (rep == 0 ? first_clip : last_clip);
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
  cluster_sync_all();   // no CTA leaves (or frees TMEM) while its partner may still signal it
  if (warp == 0) tmem_dealloc2(tbase, 512);
}

template <int PP>
static int launch_pair(const K1Params& k1p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(head_tokens_pair_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KP_SMEM) != cudaSuccess)
      return PASN_ERR_CUDA;
    attr_done = true;
  }
  int grid = ceil_div(k1p.N, k1p.clips_per_cta);
  grid = (grid + 1) & ~1;   // whole pairs; a trailing CTA without clips only runs ghost tiles
  head_tokens_pair_kernel<PP><<<grid, KP_THREADS, KP_SMEM, st>>>(k1p);
  PASN_LAUNCH_CHECK();
  return PASN_OK;
}

int launch_k1_pair(const K1Params& k1p, int ppad, cudaStream_t st) {
  if (k1p.f32_in) return PASN_ERR_UNSUPPORTED;   // fp32 feature maps: single-CTA kernel only
  if (ppad <= 16) return launch_pair<16>(k1p, st);
  if (ppad <= 32) return launch_pair<32>(k1p, st);
  if (ppad <= 40) return launch_pair<40>(k1p, st);
  return launch_pair<48>(k1p, st);
}

}  // namespace pasn
