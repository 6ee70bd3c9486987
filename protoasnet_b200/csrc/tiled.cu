// "Tiled" tensor-core path of the prototype head: the head as a chain of tcgen05 GEMMs (tc_gemm.cu) over token-major
// activations, for the shapes the fused token kernel (head_sm100*.cu) does not take -- D != 256, feature maps with fewer
// than 128 voxels (image head, BASELINE config 2), thousands of prototypes (config 5: the O layer and the pooling
// contraction are then the dominant GEMMs) -- and for fp32 feature maps, where every product runs as a three-pass bf16
// hi/lo split (x*w ~ xh*wh + xh*wl + xl*wh, fp32 accumulation), i.e. fp32-grade results on the tensor cores.
//
// Reference computation (src/models/Video_XProtoNet.py:82-98, src/models/XProtoNet.py:51-67), per clip with S voxels:
//   [H1 | G1] = relu([W1; W3] x + [b1; b3])          one GEMM over all tokens        (add_on_layers[0], occurrence_module[0])
//   G2 = relu(W4 G1 + b4)                                                              (occurrence_module[2])
//   O  = |W5 G2|            per clip, written straight in the [N,P,S] layout of occurrence_map   (occurrence_module[4], abs)
//   pooled[p,:] = sum_s O[p,s] H1[:,s]     per clip, K = S; H1 (token-major) is the MN-major B operand
//   FE = pooled W2^T + (sum_s O[p,s]) b2     W2 applied after pooling (exact in real arithmetic)
//   ... when that is the cheaper order (P rows per clip, times the hi/lo passes of the pooled vectors, against S voxels);
//   otherwise (image head: S = 49, config 5: P = 4096 > S) the reference's order is kept:
//   F = W2 H1 + b2 over all tokens, FE[p,:] = sum_s O[p,s] F[:,s] straight out of the pooling GEMM in fp32.
//   cosine / (.+1)/2 / logits / 1-s / push keys: proto_stage.cu (fp32)
// Hidden activations are bf16 in HBM between the GEMMs (hi|lo planes in fp32 mode).
#include "tc_gemm.cuh"

namespace pasn {

namespace {

struct Plan {
  int ex;        // bf16 planes per activation / weight: 1 (bf16 mode), 2 (fp32 mode: hi | lo)
  int C, D, D2, P, K, S, Sp;   // Sp: column pitch of one plane of the pooling A operand (occurrence values)
  bool occ_direct;             // bf16 mode and 16-byte aligned rows: the pooling reads the occurrence_map buffer itself
  bool w2_first;               // F = W2 H1 + b2 before the pooling (cheaper when S is small against P)
  bool tok_c;                  // few prototypes: O over all tokens at once (token-major), the map leaves channel-major per clip
  int Pp;                      // column pitch of one plane of the token-major occurrence values
  int bn_c, tiles_n_c;         // tile width / count along S of the O GEMM (psum parts = 2 * tiles_n_c)
  int nb;                      // clips per chunk
  size_t off_xt, off_y, off_g2, off_occ, off_psum, off_pool, off_f, off_fe, total;
};

inline int pick_bn(int n) { return n <= 64 ? 64 : (n <= 128 ? 128 : 256); }

Plan make_plan(const pasn_dims& d) {
  Plan p{};
  p.ex = d.dtype == PASN_F32 ? 2 : 1;
  p.C = d.C; p.D = d.D; p.D2 = d.D / 2; p.P = d.P; p.K = d.K; p.S = d.S;
  p.occ_direct = p.ex == 1 && (d.S * 2) % 16 == 0;
  // MACs of the W2 stage per clip: S*D*D*passes before the pooling, P*D*D*passes after it (bf16 mode: the pooled vectors
  // go in as hi + lo planes, two passes; fp32 mode: three passes either way)
  p.w2_first = p.ex == 1 ? (d.S < 2 * d.P) : (d.S < d.P);
  p.tok_c = d.P <= 64;         // (the full chain uses it only together with w2_first: the other order needs row sums over s)
  p.Pp = (d.P + 7) / 8 * 8;
  p.Sp = p.ex == 2 ? (d.S + 63) / 64 * 64 : (d.S + 7) / 8 * 8;
  p.bn_c = pick_bn(p.ex == 2 ? p.Sp : d.S);
  p.tiles_n_c = ceil_div(p.ex == 2 ? p.Sp : d.S, p.bn_c);
  const size_t S = d.S, ex = p.ex;
  const size_t occ_elems = (size_t)d.P * p.Sp > S * p.Pp ? (size_t)d.P * p.Sp : S * p.Pp;   // either orientation
  const size_t per_clip = S * ex * d.C * 2 + S * ex * 2 * d.D * 2 + S * ex * p.D2 * 2 + occ_elems * ex * 2 +
                          (size_t)d.P * 2 * p.tiles_n_c * 4 + (size_t)d.P * 2 * d.D * 2 + S * ex * d.D * 2 +
                          (size_t)d.P * d.D * 4 + 2048;
  long long nb = (long long)(((size_t)2 << 30) / per_clip);
  if (nb < 1) nb = 1;
  if (nb > d.N) nb = d.N > 0 ? d.N : 1;
  p.nb = (int)nb;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 1024); return r; };
  p.off_xt = take((size_t)p.nb * S * ex * d.C * 2);
  p.off_y = take((size_t)p.nb * S * ex * 2 * d.D * 2);
  p.off_g2 = take((size_t)p.nb * S * ex * p.D2 * 2);
  p.off_occ = take((size_t)p.nb * occ_elems * ex * 2);
  p.off_psum = take((size_t)p.nb * d.P * 2 * p.tiles_n_c * 4);
  p.off_pool = take(p.w2_first ? 0 : (size_t)p.nb * d.P * 2 * d.D * 2);
  p.off_f = take(p.w2_first ? (size_t)p.nb * S * ex * d.D * 2 : 0);
  p.off_fe = take((size_t)p.nb * d.P * d.D * 4);
  p.total = o + 256;
  return p;
}

struct PackLayout {
  size_t off_w13, off_w4, off_w5, off_w2, off_b13, off_b4, off_b2, total;
};
PackLayout pack_layout(const pasn_dims& d) {
  const size_t ex = d.dtype == PASN_F32 ? 2 : 1, D2 = d.D / 2;
  PackLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  L.off_w13 = take((size_t)2 * d.D * ex * d.C * 2);
  L.off_w4 = take(D2 * ex * d.D * 2);
  L.off_w5 = take((size_t)d.P * ex * D2 * 2);
  L.off_w2 = take((size_t)d.D * ex * d.D * 2);
  L.off_b13 = take((size_t)2 * d.D * 4);
  L.off_b4 = take(D2 * 4);
  L.off_b2 = take((size_t)d.D * 4);
  L.total = o;
  return L;
}

// fp32 [rows][cols] -> bf16 planes [rows][ex*cols]: plane 0 = bf16(w), plane 1 = bf16(w - plane 0)
__global__ void pack_planes_kernel(const float* __restrict__ w, int rows, int cols, int ex, __nv_bfloat16* __restrict__ out) {
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float v = w[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[(size_t)r * ex * cols + c] = hi;
    if (ex == 2) *(out + (size_t)r * ex * cols + cols + c) = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}
__global__ void pack_bias_kernel(const float* __restrict__ b, int n, int round, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = round ? round_bf16(b[i]) : b[i];
}

// feature map -> token-major bf16 planes XT[(n*S + s)][ex*C]
//   NCS ([n][C][S]): 64-channel x 32-voxel tiles through shared memory: loads run along s (one channel row segment per warp
//   instruction), stores along c (one full 128-byte line of bf16 channel pairs per token and plane);
//   NSC fp32 ([n][S][C]): plane split only
template <typename T>
__global__ void __launch_bounds__(256) to_tokens_ncs_kernel(const T* __restrict__ x, int C, int S, int ex, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[64][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 64, s0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* xn = x + (size_t)n * C * S;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int cl = warp * 8 + j, c = c0 + cl, s = s0 + lane;
    tile[cl][lane] = (c < C && s < S) ? to_f32<T>(xn[(size_t)c * S + s]) : 0.f;
  }
  __syncthreads();
  const int c = c0 + 2 * lane;           // C is a multiple of 64 on this path
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int sl = warp * 4 + j, s = s0 + sl;
    if (s < S && c < C) {
      const float v0 = tile[2 * lane][sl], v1 = tile[2 * lane + 1][sl];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      __nv_bfloat16* row = out + ((size_t)n * S + s) * ex * C;
      *reinterpret_cast<__nv_bfloat162*>(row + c) = __nv_bfloat162(h0, h1);
      if (ex == 2)
        *reinterpret_cast<__nv_bfloat162*>(row + C + c) =
            __nv_bfloat162(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
    }
  }
}
__global__ void to_tokens_nsc_f32_kernel(const float* __restrict__ x, long long rows, int C, __nv_bfloat16* __restrict__ out) {
  const long long n = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float v = x[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[r * 2 * C + c] = hi;
    *(out + r * 2 * C + C + c) = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// passes of a product of two operands that each come as planes (hi at column 0, lo at column *_lo)
void set_passes(tcg::Gemm& g, int ex, int a_lo, int b_lo) {
  g.pair = 1;   // CTA pairs wherever the shape allows (tc_gemm.cu decides)
  for (int q = 0; q < 4; ++q) g.a_off[q] = g.b_off[q] = 0;
  if (ex == 1) { g.npass = 1; return; }
  // hi*hi, hi*lo, lo*hi.  The lo*lo term is below the representation error of the split itself (each operand keeps 16
  // significant bits: |x - xh - xl| <= 2^-17 |x|, and |xl wl| <= 2^-18 |x w|), so it is not computed.
  g.npass = 3;
  g.b_off[1] = b_lo;
  g.a_off[2] = a_lo;
}

}  // namespace

bool tiled_supported(const pasn_dims& d) {
  if (d.C % 64 != 0 || d.D % 128 != 0) return false;   // every K (C, D, D/2) is a whole number of 64-wide k-blocks
  if (d.P < 1 || d.S < 1) return false;
  return tcg::available();
}
size_t tiled_workspace_bytes(const pasn_dims& d) { return make_plan(d).total; }
size_t tiled_packed_bytes(const pasn_dims& d) { return pack_layout(d).total; }

int tiled_pack_weights(const pasn_weights& w, const pasn_dims& d, void* packed, cudaStream_t st) {
  if (((uintptr_t)packed & 255) != 0) return PASN_ERR_ALIGN;
  const PackLayout L = pack_layout(d);
  const int ex = d.dtype == PASN_F32 ? 2 : 1, D2 = d.D / 2;
  char* pk = reinterpret_cast<char*>(packed);
  auto planes = [&](const float* src, int rows, int cols, size_t off) {
    pack_planes_kernel<<<148 * 2, 256, 0, st>>>(src, rows, cols, ex, reinterpret_cast<__nv_bfloat16*>(pk + off));
    count_launch();
  };
  planes(w.addon_w1, d.D, d.C, L.off_w13);
  planes(w.occ_w1, d.D, d.C, L.off_w13 + (size_t)d.D * ex * d.C * 2);
  planes(w.occ_w2, D2, d.D, L.off_w4);
  planes(w.occ_w3, d.P, D2, L.off_w5);
  planes(w.addon_w2, d.D, d.D, L.off_w2);
  const int rnd = ex == 1;
  auto bias = [&](const float* src, int n, size_t off) {
    pack_bias_kernel<<<ceil_div(n, 256), 256, 0, st>>>(src, n, rnd, reinterpret_cast<float*>(pk + off));
    count_launch();
  };
  bias(w.addon_b1, d.D, L.off_b13);
  bias(w.occ_b1, d.D, L.off_b13 + (size_t)d.D * 4);
  bias(w.occ_b2, D2, L.off_b4);
  bias(w.addon_b2, d.D, L.off_b2);
  PASN_LAUNCH_CHECK();
  return PASN_OK;
}

// the chain for clips [n0, n0+nb); occ_only stops after the occurrence map
static int tiled_chunk(const void* feat, const pasn_weights& w, const void* packed, const pasn_dims& d, const Plan& p, int n0,
                       int nb, float* logits, float* sim, void* occ, float* feats, float* dist, const pasn_push_args* push,
                       char* ws, bool occ_only, cudaStream_t st) {
  const PackLayout L = pack_layout(d);
  const char* pk = reinterpret_cast<const char*>(packed);
  const int ex = p.ex, C = p.C, D = p.D, D2 = p.D2, P = p.P, S = p.S;
  const long long T = (long long)nb * S;
  const size_t elt = d.dtype == PASN_BF16 ? 2 : 4;
  int rc;

  // ---- tokens
  const __nv_bfloat16* xt;
  const char* x = reinterpret_cast<const char*>(feat) + (size_t)n0 * C * S * elt;
  // bf16 NCDHW feature maps with 16-byte aligned channel rows need no transposition: [C][S] per clip IS the MN-major form
  // of the A operand (voxels contiguous), which the TMA unit tiles straight out of the caller's buffer
  const bool x_direct = d.layout == PASN_LAYOUT_NCS && ex == 1 && (S * 2) % 16 == 0 && ((uintptr_t)x & 15) == 0;
  if (d.layout == PASN_LAYOUT_NSC && ex == 1) {
    xt = reinterpret_cast<const __nv_bfloat16*>(x);   // channels_last bf16 feature map: already token-major
    if (((uintptr_t)xt & 15) != 0) return PASN_ERR_ALIGN;
  } else if (x_direct) {
    xt = reinterpret_cast<const __nv_bfloat16*>(x);
  } else {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ws + p.off_xt);
    if (d.layout == PASN_LAYOUT_NSC) {
      to_tokens_nsc_f32_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<const float*>(x), T, C, dst);
    } else {
      dim3 grid(ceil_div(S, 32), ceil_div(C, 64), nb);
      if (ex == 1) to_tokens_ncs_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), C, S, ex, dst);
      else to_tokens_ncs_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), C, S, ex, dst);
    }
    PASN_LAUNCH_CHECK();
    count_launch();
    xt = dst;
  }
  __nv_bfloat16* Y = reinterpret_cast<__nv_bfloat16*>(ws + p.off_y);
  __nv_bfloat16* G2 = reinterpret_cast<__nv_bfloat16*>(ws + p.off_g2);
  __nv_bfloat16* OCC = reinterpret_cast<__nv_bfloat16*>(ws + p.off_occ);
  float* PSUM = reinterpret_cast<float*>(ws + p.off_psum);
  __nv_bfloat16* POOL = reinterpret_cast<__nv_bfloat16*>(ws + p.off_pool);
  float* FE = feats ? feats + (size_t)n0 * P * D : reinterpret_cast<float*>(ws + p.off_fe);
  const __nv_bfloat16* W13 = reinterpret_cast<const __nv_bfloat16*>(pk + L.off_w13);
  const __nv_bfloat16* W4 = reinterpret_cast<const __nv_bfloat16*>(pk + L.off_w4);
  const __nv_bfloat16* W5 = reinterpret_cast<const __nv_bfloat16*>(pk + L.off_w5);
  const __nv_bfloat16* W2 = reinterpret_cast<const __nv_bfloat16*>(pk + L.off_w2);
  const float* b13 = reinterpret_cast<const float*>(pk + L.off_b13);
  const float* b4 = reinterpret_cast<const float*>(pk + L.off_b4);
  const float* b2 = reinterpret_cast<const float*>(pk + L.off_b2);

  // ---- (A) [H1 | G1] = relu(X [W1; W3]^T + [b1; b3])        (occurrence only: G1 alone)
  const int nA = occ_only ? D : 2 * D;          // output columns per plane
  {
    tcg::Gemm g{};
    g.A = xt; g.lda = (long long)ex * C; g.a_bs = 0; g.a_batched = 0; g.ka = ex * C;
    g.B = occ_only ? W13 + (size_t)D * ex * C : W13; g.ldb = (long long)ex * C; g.b_bs = 0; g.b_batched = 0; g.kb = ex * C;
    g.M = (int)T; g.N = nA; g.K = C; g.batch = 1; g.bn = 256;
    set_passes(g, ex, C, C);
    g.bias = occ_only ? b13 + D : b13; g.act = tcg::ACT_RELU;
    g.out[0] = {Y, ex == 1 ? tcg::OUT_BF16 : tcg::OUT_BF16_HILO, (long long)ex * nA, 0, nA};
    if (x_direct) {   // one GEMM per clip: rows = its S voxels, operand [C rows][S] as it lies in the feature map
      g.a_mn_major = 1; g.lda = S; g.a_bs = (long long)C * S; g.a_batched = 1; g.ka = S; g.a_rows = C;
      g.M = S; g.batch = nb;
      g.out[0].bs = (long long)S * ex * nA;
    }
    if ((rc = tcg::launch(g, st))) return rc;
  }
  // ---- (B) G2 = relu(G1 W4^T + b4)
  const int g1_off = occ_only ? 0 : D;          // column of G1 inside a plane of Y
  {
    tcg::Gemm g{};
    g.A = Y + g1_off; g.lda = (long long)ex * nA; g.a_batched = 0; g.ka = ex * nA - g1_off;
    g.B = W4; g.ldb = (long long)ex * D; g.b_batched = 0; g.kb = ex * D;
    g.M = (int)T; g.N = D2; g.K = D; g.batch = 1; g.bn = pick_bn(D2);
    set_passes(g, ex, nA, D);
    g.bias = b4; g.act = tcg::ACT_RELU;
    g.out[0] = {G2, ex == 1 ? tcg::OUT_BF16 : tcg::OUT_BF16_HILO, (long long)ex * D2, 0, D2};
    if ((rc = tcg::launch(g, st))) return rc;
  }
  // ---- (C) O[n] = |W5 G2[n]^T|  -> occurrence_map [n][P][S] (+ the pooling operand copy when the map cannot serve as one)
  void* occ_user = occ ? reinterpret_cast<char*>(occ) + (size_t)n0 * P * S * elt : nullptr;
  const __nv_bfloat16* pool_a = nullptr;     // pooling A operand
  long long pool_lda = 0;
  const bool tok_c = p.tok_c && (occ_only || p.w2_first);
  if (tok_c) {
    // few prototypes: O^T = |G2 W5^T| over all tokens in one GEMM (rows = tokens, N = P).  The occurrence map leaves
    // channel-major per clip through plain stores that are coalesced along the voxels; the token-major copy
    // [T][ex*Pp] is the pooling's MN-major A operand.
    tcg::Gemm g{};
    g.A = G2; g.lda = (long long)ex * D2; g.a_batched = 0; g.ka = ex * D2;
    g.B = W5; g.ldb = (long long)ex * D2; g.b_batched = 0; g.kb = ex * D2;
    g.M = (int)T; g.N = P; g.K = D2; g.batch = 1; g.bn = 64;
    set_passes(g, ex, D2, D2);
    g.act = tcg::ACT_ABS;
    int no = 0;
    if (occ_user != nullptr) {
      g.out[no] = {occ_user, ex == 1 ? tcg::OUT_BF16 : tcg::OUT_F32, (long long)S, (long long)P * S, 0, P, 0, S};
      ++no;
    }
    if (!occ_only) {
      g.out[no] = {OCC, ex == 1 ? tcg::OUT_BF16 : tcg::OUT_BF16_HILO, (long long)ex * p.Pp, 0, p.Pp, P, 0, 0};
      ++no;
      pool_a = OCC; pool_lda = (long long)ex * p.Pp;
    }
    if ((rc = tcg::launch(g, st))) return rc;
  } else {
    tcg::Gemm g{};
    g.A = W5; g.lda = (long long)ex * D2; g.a_batched = 0; g.ka = ex * D2;
    g.B = G2; g.ldb = (long long)ex * D2; g.b_bs = (long long)S * ex * D2; g.b_batched = 1; g.kb = ex * D2;
    g.b_rows = S;   // tokens of this clip; the padded columns of the operand copy come out as exact zeros
    g.M = P; g.K = D2; g.batch = nb; g.bn = p.bn_c;
    g.N = ex == 2 ? p.Sp : S;
    set_passes(g, ex, D2, D2);
    g.act = tcg::ACT_ABS;
    int no = 0;
    const bool need_copy = !occ_only && !(p.occ_direct && occ_user != nullptr);
    if (occ_user != nullptr) {
      g.out[no] = {occ_user, ex == 1 ? tcg::OUT_BF16 : tcg::OUT_F32, (long long)S, (long long)P * S, 0};
      g.out[no].ncols = S;
      ++no;
    }
    if (need_copy) {
      g.out[no] = {OCC, ex == 1 ? tcg::OUT_BF16 : tcg::OUT_BF16_HILO, (long long)ex * p.Sp, (long long)P * ex * p.Sp, p.Sp};
      g.out[no].ncols = ex == 2 ? p.Sp : S;
      ++no;
      pool_a = OCC; pool_lda = (long long)ex * p.Sp;
    } else if (!occ_only) {
      pool_a = reinterpret_cast<const __nv_bfloat16*>(occ_user); pool_lda = S;
    }
    g.psum = (occ_only || p.w2_first) ? nullptr : PSUM;
    g.psum_rounded = ex == 1;
    if ((rc = tcg::launch(g, st))) return rc;
  }
  if (occ_only) return PASN_OK;
  if (p.w2_first) {
    // ---- (D') F = H1 W2^T + b2 over all tokens, then FE[n] = O[n] F[n] (K = S) straight into fp32
    __nv_bfloat16* F = reinterpret_cast<__nv_bfloat16*>(ws + p.off_f);
    {
      tcg::Gemm g{};
      g.A = Y; g.lda = (long long)ex * 2 * D; g.a_batched = 0; g.ka = ex * 2 * D;
      g.B = W2; g.ldb = (long long)ex * D; g.b_batched = 0; g.kb = ex * D;
      g.M = (int)T; g.N = D; g.K = D; g.batch = 1; g.bn = D >= 256 ? 256 : 128;
      set_passes(g, ex, 2 * D, D);
      g.bias = b2; g.act = tcg::ACT_NONE;
      g.out[0] = {F, ex == 1 ? tcg::OUT_BF16 : tcg::OUT_BF16_HILO, (long long)ex * D, 0, D};
      if ((rc = tcg::launch(g, st))) return rc;
    }
    {
      tcg::Gemm g{};
      g.A = pool_a; g.lda = pool_lda; g.a_bs = (long long)P * pool_lda; g.a_batched = 1;
      g.ka = ex == 2 ? 2 * p.Sp : S;
      g.B = F; g.ldb = (long long)ex * D; g.b_bs = (long long)S * ex * D; g.b_batched = 1; g.kb = ex * D;
      g.b_mn_major = 1; g.b_rows = S;
      g.M = P; g.N = D; g.K = ex == 2 ? p.Sp : S; g.batch = nb; g.bn = D >= 256 ? 256 : 128;
      set_passes(g, ex, p.Sp, D);
      if (tok_c) {   // occurrence values token-major: [S rows][P] per clip, prototypes contiguous (MN-major A)
        g.a_mn_major = 1; g.a_bs = (long long)S * pool_lda; g.ka = ex * p.Pp; g.a_rows = S; g.K = S;
        set_passes(g, ex, p.Pp, D);
      }
      g.act = tcg::ACT_NONE;
      g.out[0] = {FE, tcg::OUT_F32, (long long)D, (long long)P * D, 0};
      if ((rc = tcg::launch(g, st))) return rc;
    }
  } else {
  // ---- (D) pooled[n] = O[n] H1[n]   (K = S; H1 token-major = MN-major B operand) -> bf16 hi | lo planes
  {
    tcg::Gemm g{};
    g.A = pool_a; g.lda = pool_lda; g.a_bs = (long long)P * pool_lda; g.a_batched = 1;
    g.ka = ex == 2 ? 2 * p.Sp : S;
    g.B = Y; g.ldb = (long long)ex * 2 * D; g.b_bs = (long long)S * ex * 2 * D; g.b_batched = 1; g.kb = ex * 2 * D;
    g.b_mn_major = 1; g.b_rows = S;
    g.M = P; g.N = D; g.K = ex == 2 ? p.Sp : S; g.batch = nb; g.bn = D >= 256 ? 256 : 128;
    set_passes(g, ex, p.Sp, 2 * D);
    g.act = tcg::ACT_NONE;
    g.out[0] = {POOL, tcg::OUT_BF16_HILO, (long long)2 * D, (long long)P * 2 * D, D};
    if ((rc = tcg::launch(g, st))) return rc;
  }
  // ---- (E) FE = pooled W2^T + (sum_s O) b2^T   (pooled = hi + lo; fp32 mode: W2 = hi + lo as well)
  {
    tcg::Gemm g{};
    g.A = POOL; g.lda = 2 * D; g.a_batched = 0; g.ka = 2 * D;
    g.B = W2; g.ldb = (long long)ex * D; g.b_batched = 0; g.kb = ex * D;
    g.M = nb * P; g.N = D; g.K = D; g.batch = 1; g.bn = D >= 256 ? 256 : 128;
    set_passes(g, ex, D, D);
    if (ex == 1) { g.npass = 2; g.a_off[1] = D; }   // pooled = hi + lo planes, W2 bf16
    g.rowparts = PSUM; g.nparts = 2 * p.tiles_n_c; g.colvec = b2;
    g.act = tcg::ACT_NONE;
    g.out[0] = {FE, tcg::OUT_F32, (long long)D, 0, 0};
    if ((rc = tcg::launch(g, st))) return rc;
  }
  }
  // ---- cosine / similarity / logits / distance / push keys (+ winner capture)
  pasn_push_args pa;
  const pasn_push_args* pp = nullptr;
  if (push) { pa = *push; pa.labels += n0; pa.global_offset += n0; pp = &pa; }
  return launch_proto_stage(FE, w.prototypes, w.last_layer, nb, P, D, d.K, logits + (size_t)n0 * d.K, sim + (size_t)n0 * P,
                            dist ? dist + (size_t)n0 * P : nullptr, pp, st);
}

int tiled_head_forward(const void* feat, const pasn_weights& w, const void* packed, const pasn_dims& d, float* logits,
                       float* sim, void* occ, float* feats, float* dist, const pasn_push_args* push, void* ws,
                       size_t ws_bytes, cudaStream_t st) {
  if (!tiled_supported(d) || packed == nullptr) return PASN_ERR_UNSUPPORTED;
  const Plan p = make_plan(d);
  if (ws_bytes < p.total) return PASN_ERR_WORKSPACE;
  if (((uintptr_t)ws & 255) != 0 || ((uintptr_t)feat & 15) != 0) return PASN_ERR_ALIGN;
  main_kernel_begin(st);   // tiled path: the "dominant kernel" is the GEMM chain
  for (int n0 = 0; n0 < d.N; n0 += p.nb) {
    const int nb = d.N - n0 < p.nb ? d.N - n0 : p.nb;
    const int rc = tiled_chunk(feat, w, packed, d, p, n0, nb, logits, sim, occ, feats, dist, push, reinterpret_cast<char*>(ws),
                               false, st);
    if (rc) return rc;
  }
  main_kernel_end(st);
  return PASN_OK;
}

int tiled_occurrence_only(const void* feat, const pasn_weights& w, const void* packed, const pasn_dims& d, void* occ, void* ws,
                          size_t ws_bytes, cudaStream_t st) {
  if (!tiled_supported(d) || packed == nullptr) return PASN_ERR_UNSUPPORTED;
  const Plan p = make_plan(d);
  if (ws_bytes < p.total) return PASN_ERR_WORKSPACE;
  if (((uintptr_t)ws & 255) != 0 || ((uintptr_t)feat & 15) != 0) return PASN_ERR_ALIGN;
  for (int n0 = 0; n0 < d.N; n0 += p.nb) {
    const int nb = d.N - n0 < p.nb ? d.N - n0 : p.nb;
    const int rc = tiled_chunk(feat, w, packed, d, p, n0, nb, nullptr, nullptr, occ, nullptr, nullptr, nullptr,
                               reinterpret_cast<char*>(ws), true, st);
    if (rc) return rc;
  }
  return PASN_OK;
}

}  // namespace pasn
