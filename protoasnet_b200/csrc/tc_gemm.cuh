// Tiled tcgen05 GEMM with TMA tensor-map loads and stores: the building block of the "tiled" tensor-core path that serves
// the shapes the fused token kernel does not (D != 256, S < 128, P > 48, fp32-exact mode): every 1x1(x1) convolution of
// the head, the occurrence-weighted pooling contraction and the W2 stage are instances of
//     OUT[b][m][n] = act( sum_pass sum_k A[b][m][a_off[pass] + k] * B[b][n][b_off[pass] + k]  + bias[n] + rowvec[m]*colvec[n] )
// with 16-bit operands, fp32 accumulation in TMEM, and up to four operand passes: the hi/lo split that gives fp32-grade
// products, x*w = (xh + xl)(wh + wl) with xh = bf16(x) (8 significant bits, full range) and xl = fp16(x - xh) (11 more bits;
// the residual is 2^-8 of x, well inside fp16's range) -- 19 significant bits per operand, all four partial products kept.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace pasn {
namespace tcg {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_ABS = 2 };
enum { OUT_NONE = 0, OUT_BF16 = 1, OUT_BF16_HILO = 2, OUT_F32 = 3 };   // HILO: bf16 hi plane | fp16 lo plane

struct Output {
  void* ptr;          // base of the [batch][M][ld] array (element (b, m, n) at ptr + (b*bs + m*ld + n) * elt)
  int mode;           // OUT_*
  long long ld, bs;   // row / batch stride in elements
  int lo_off;         // OUT_BF16_HILO: column offset of the lo half (the hi half starts at column 0)
  int ncols;          // columns of this output (0: N); columns beyond it are computed but not stored
};

struct Gemm {
  // A: [a_batched ? batch : 1][M][ka] bf16, row stride lda (elements); K-major (k contiguous)
  const void* A; long long lda, a_bs; int a_batched; int ka;
  // B: K-major [b_batched ? batch : 1][N][kb] (row stride ldb) or, b_mn_major, [batch?][K][nb] with n contiguous
  const void* B; long long ldb, b_bs; int b_batched; int kb; int b_mn_major;
  int b_rows;                    // rows B really has per batch item (0: N, or K when b_mn_major); rows beyond read as zero
  int M, N, K, batch;            // K per pass
  int npass; int a_off[4], b_off[4];   // column offset of each pass' operand plane
  int a_f16[4], b_f16[4];              // that plane holds fp16 (lo planes) instead of bf16
  int bn;                        // tile width: 64, 128 or 256
  const float* bias;             // [N] or null
  const float* rowparts; int nparts; const float* colvec;   // rank-1 term (sum_t rowparts[(b*M+m)*nparts + t]) * colvec[n], or null
  int act;
  Output out[2];
  int lo_f16;                    // OUT_BF16_HILO writes its lo plane as fp16 (else bf16)
  float* psum;                   // optional [batch*M][2*tiles_n]: row sums of the outputs per column half-tile ...
  int psum_rounded;              // ... of the bf16-rounded values (what a bf16 consumer of out[0] will read) or of the fp32 values
};

bool lo_planes_f16();                         // lo planes are fp16 (default) or bf16 (PASN_TILED_LO=bf16)
int launch(const Gemm& g, cudaStream_t st);   // 0 or a negative pasn_status
bool available();                             // driver entry point for tensor-map encoding found

}  // namespace tcg
}  // namespace pasn
