"""The building block of the tiled tensor-core path (csrc/tc_gemm.cu: TMA-fed tcgen05 GEMM) against torch matmul in
float64, through the library's test hook pasn_debug_tc_gemm: operand layouts (K-major / MN-major, batched, zero-filled
edges), operand passes (hi/lo split), split-K, epilogue (bias, rank-1 term, activations, element-wise inputs) and output
modes (bf16 / hi|lo planes / fp32, TMA and plain stores, channel-major per clip)."""
import ctypes as C

import numpy as np
import pytest
import torch

from protoasnet_b200 import _lib

pytestmark = pytest.mark.gpu

ACT_NONE, ACT_RELU, ACT_ABS = 0, 1, 2
OUT_NONE, OUT_BF16, OUT_BF16_HILO, OUT_F32 = 0, 1, 2, 3


class Output(C.Structure):   # pasn::tcg::Output
    _fields_ = [("ptr", C.c_void_p), ("mode", C.c_int), ("ld", C.c_longlong), ("bs", C.c_longlong), ("lo_off", C.c_int),
                ("ncols", C.c_int), ("absval", C.c_int), ("trans_S", C.c_int)]


class Aux(C.Structure):      # pasn::tcg::Aux
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_longlong), ("bs", C.c_longlong), ("cs", C.c_longlong)]


class Gemm(C.Structure):     # pasn::tcg::Gemm
    _fields_ = [("A", C.c_void_p), ("lda", C.c_longlong), ("a_bs", C.c_longlong), ("a_batched", C.c_int), ("ka", C.c_int),
                ("a_mn_major", C.c_int), ("a_rows", C.c_int),
                ("B", C.c_void_p), ("ldb", C.c_longlong), ("b_bs", C.c_longlong), ("b_batched", C.c_int), ("kb", C.c_int),
                ("b_mn_major", C.c_int), ("b_rows", C.c_int),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int), ("k_rows_per_batch", C.c_int),
                ("npass", C.c_int), ("a_off", C.c_int * 4), ("b_off", C.c_int * 4), ("bn", C.c_int),
                ("bias", C.c_void_p), ("rowparts", C.c_void_p), ("nparts", C.c_int), ("colvec", C.c_void_p),
                ("addin", Aux), ("signin", Aux), ("mask", Aux), ("act", C.c_int), ("out", Output * 2),
                ("psum", C.c_void_p), ("psum_rounded", C.c_int), ("pair", C.c_int),
                ("rowstat", C.c_void_p), ("dotvec", C.c_void_p), ("dot_ld", C.c_longlong), ("dot_mod", C.c_int),
                ("group", C.c_int), ("group_rows", C.c_int), ("group_items", C.c_int), ("stat_rows", C.c_longlong),
                ("dot_early", C.c_int)]


def _run(g):
    lib = _lib.load()
    assert lib.pasn_debug_tc_gemm_desc_bytes() == C.sizeof(Gemm), "tests' mirror of pasn::tcg::Gemm is out of date"
    _lib.check(lib.pasn_debug_tc_gemm(C.byref(g), C.sizeof(Gemm), torch.cuda.current_stream().cuda_stream), "tc_gemm")
    torch.cuda.synchronize()
    assert lib.pasn_debug_fault() == 0


def _bf(t):
    return t.to(torch.bfloat16)


def _rand(shape, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, device="cuda", generator=gen)


def _close(got, ref, tol):
    got, ref = got.double(), ref.double()
    scale = float(ref.abs().max()) + 1e-30
    err = float((got - ref).abs().max()) / scale
    assert err <= tol, f"max error {err:.3e} of scale (tolerance {tol:.1e})"


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 256), (300, 200, 192, 256), (1000, 72, 320, 128), (77, 40, 100, 64),
                                      (4096, 1024, 512, 256)])
def test_kmajor_bf16_bias_relu(M, N, K, bn, pair):
    """OUT = relu(A B^T + bias): both operands K-major, ragged edges zero-filled by the TMA unit, bf16 TMA store."""
    ka = (K + 7) // 8 * 8                       # 16-byte row pitch
    A = torch.zeros((M, ka), device="cuda", dtype=torch.bfloat16); A[:, :K] = _bf(_rand((M, K), 1))
    B = torch.zeros((N, ka), device="cuda", dtype=torch.bfloat16); B[:, :K] = _bf(_rand((N, K), 2))
    bias = _rand((N,), 3)
    ldo = (N + 7) // 8 * 8
    out = torch.full((M, ldo), 7.0, device="cuda", dtype=torch.bfloat16)
    g = Gemm()
    g.A, g.lda, g.ka = A.data_ptr(), ka, ka
    g.B, g.ldb, g.kb = B.data_ptr(), ka, ka
    g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair = M, N, K, 1, 1, bn, pair
    g.bias, g.act = bias.data_ptr(), ACT_RELU
    g.out[0] = Output(out.data_ptr(), OUT_BF16, ldo, 0, 0, 0, 0, 0)
    _run(g)
    ref = torch.relu(A[:, :K].double() @ B[:, :K].double().T + bias.double())
    _close(out[:, :N], ref, 6e-3)
    assert bool((out[:, N:] == 7.0).all()), "columns beyond N must not be written"


@pytest.mark.parametrize("pair", [0, 1])
def test_hilo_passes_reach_fp32_grade(pair):
    """fp32 operands as bf16 hi|lo planes, three passes (hi*hi, hi*lo, lo*hi), hi|lo output planes + fp32 output."""
    M, N, K = 520, 384, 256
    A32, B32 = _rand((M, K), 4), _rand((N, K), 5)

    def planes(t):
        hi = _bf(t)
        return torch.cat([hi, _bf(t - hi.float())], dim=1).contiguous()
    A, B = planes(A32), planes(B32)
    o_planes = torch.zeros((M, 2 * N), device="cuda", dtype=torch.bfloat16)
    o32 = torch.zeros((M, N), device="cuda")
    g = Gemm()
    g.A, g.lda, g.ka = A.data_ptr(), 2 * K, 2 * K
    g.B, g.ldb, g.kb = B.data_ptr(), 2 * K, 2 * K
    g.M, g.N, g.K, g.batch, g.bn, g.pair = M, N, K, 1, 128, pair
    g.npass = 3
    g.b_off[1] = K
    g.a_off[2] = K
    g.out[0] = Output(o_planes.data_ptr(), OUT_BF16_HILO, 2 * N, 0, N, 0, 0, 0)
    g.out[1] = Output(o32.data_ptr(), OUT_F32, N, 0, 0, 0, 0, 0)
    _run(g)
    ref = A32.double() @ B32.double().T
    _close(o32, ref, 3e-5)
    _close(o_planes[:, :N].float() + o_planes[:, N:].float(), ref, 5e-5)


@pytest.mark.parametrize("pair", [0, 1])
def test_batched_mn_major_b_and_abs_outputs(pair):
    """Per-clip contraction over tokens: A [b][P][S] K-major, B [b][S][D] MN-major with S not a multiple of the k-block
    (rows beyond S read as zero), fp32 output; second output |.| in bf16 through plain stores (unaligned pitch)."""
    nb, P, S, D = 5, 72, 49, 256
    A = _bf(_rand((nb, P, 56), 6)); A[:, :, S:] = 0
    Bt = _bf(_rand((nb, S, D), 7))
    o32 = torch.zeros((nb, P, D), device="cuda")
    oabs = torch.zeros((nb, P, D + 1), device="cuda", dtype=torch.bfloat16)
    g = Gemm()
    g.A, g.lda, g.a_bs, g.a_batched, g.ka = A.data_ptr(), 56, P * 56, 1, 56
    g.B, g.ldb, g.b_bs, g.b_batched, g.kb, g.b_mn_major, g.b_rows = Bt.data_ptr(), D, S * D, 1, D, 1, S
    g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair = P, D, S, nb, 1, 256, pair
    g.out[0] = Output(o32.data_ptr(), OUT_F32, D, P * D, 0, 0, 0, 0)
    g.out[1] = Output(oabs.data_ptr(), OUT_BF16, D + 1, P * (D + 1), 0, 0, 1, 0)
    _run(g)
    ref = torch.einsum("bps,bsd->bpd", A[:, :, :S].double(), Bt.double())
    _close(o32, ref, 1e-5)
    _close(oabs[:, :, :D], ref.abs(), 6e-3)


def test_split_k_weight_gradient_form():
    """gW[m][n] = sum_t GY[t][m] X[t][n]: both operands token-major (MN-major), K = tokens split over batches, partials."""
    T, M, N, parts = 1000, 136, 320, 4
    kper = 256
    GY, X = _bf(_rand((T, M), 8)), _bf(_rand((T, N), 9))
    out = torch.zeros((parts, M, N), device="cuda")
    g = Gemm()
    g.A, g.lda, g.ka, g.a_mn_major, g.a_rows = GY.data_ptr(), M, M, 1, T
    g.B, g.ldb, g.kb, g.b_mn_major, g.b_rows = X.data_ptr(), N, N, 1, T
    g.M, g.N, g.K, g.batch, g.k_rows_per_batch, g.npass, g.bn = M, N, kper, parts, kper, 1, 128
    g.out[0] = Output(out.data_ptr(), OUT_F32, N, M * N, 0, 0, 0, 0)
    _run(g)
    ref = GY.double().T @ X.double()
    _close(out.sum(0), ref, 1e-5)


def test_epilogue_inputs_and_channel_major_store():
    """value = ((A B^T + addin) * sign(signin)) masked by mask > 0, stored channel-major per clip (trans_S) in fp32 and
    row-major as hi|lo planes; addin is read transposed (cs)."""
    nclip, S, N, K = 6, 49, 40, 128
    T = nclip * S
    A, B = _bf(_rand((T, K), 10)), _bf(_rand((N, K), 11))
    addin = _rand((nclip, N, S), 12)                  # [clip][n][s]: transposed against the [token][n] tile
    sign = _bf(_rand((T, N), 13)); sign[::7] = 0
    mask = _bf(_rand((T, N), 14))
    o_cm = torch.zeros((nclip, N, S), device="cuda")
    o_pl = torch.zeros((T, 2 * N), device="cuda", dtype=torch.bfloat16)
    g = Gemm()
    g.A, g.lda, g.ka = A.data_ptr(), K, K
    g.B, g.ldb, g.kb = B.data_ptr(), K, K
    g.M, g.N, g.K, g.batch, g.npass, g.bn = T, N, K, 1, 1, 64
    # token t = clip*S + s reads addin[clip][n][s]: with ld = 1, cs = S this only lines up inside one clip, so the aux
    # view is built per token row instead: ptr + t*ld with ld = 1 would cross clips -- use a [T][N] transposed copy
    addin_tn = addin.permute(0, 2, 1).reshape(T, N).contiguous()       # [t][n]
    addin_nt = addin_tn.T.contiguous()                                  # [n][t]: element (t, n) at n*T + t
    g.addin = Aux(addin_nt.data_ptr(), 1, 0, T)
    g.signin = Aux(sign.data_ptr(), N, 0, 0)
    g.mask = Aux(mask.data_ptr(), N, 0, 0)
    g.out[0] = Output(o_cm.data_ptr(), OUT_F32, S, N * S, 0, 0, 0, S)
    g.out[1] = Output(o_pl.data_ptr(), OUT_BF16_HILO, 2 * N, 0, N, 0, 0, 0)
    _run(g)
    ref = A.double() @ B.double().T + addin_tn.double()
    sg = torch.sign(sign.double())
    ref = ref * sg
    ref = torch.where(mask.double() > 0, ref, torch.zeros_like(ref))
    _close(o_cm.permute(0, 2, 1).reshape(T, N), ref, 1e-5)
    _close(o_pl[:, :N].float() + o_pl[:, N:].float(), ref, 5e-5)


def test_rank1_term_and_psum():
    """OUT = A B^T + (sum_t rowparts[m][t]) colvec[n]; psum = row sums of the (bf16-rounded) outputs per half tile."""
    M, N, K = 200, 256, 64
    A, B = _bf(_rand((M, K), 15)), _bf(_rand((N, K), 16))
    rp, cv = _rand((M, 3), 17), _rand((N,), 18)
    out = torch.zeros((M, N), device="cuda")
    psum = torch.zeros((M, 2), device="cuda")
    g = Gemm()
    g.A, g.lda, g.ka = A.data_ptr(), K, K
    g.B, g.ldb, g.kb = B.data_ptr(), K, K
    g.M, g.N, g.K, g.batch, g.npass, g.bn = M, N, K, 1, 1, 256
    g.rowparts, g.nparts, g.colvec = rp.data_ptr(), 3, cv.data_ptr()
    g.out[0] = Output(out.data_ptr(), OUT_F32, N, 0, 0, 0, 0, 0)
    g.psum, g.psum_rounded = psum.data_ptr(), 0
    _run(g)
    ref = A.double() @ B.double().T + rp.double().sum(1, keepdim=True) * cv.double()
    _close(out, ref, 1e-5)
    _close(psum.sum(1), ref.sum(1), 1e-5)


@pytest.mark.parametrize("pair", [0, 1])
def test_row_statistics_without_an_output(pair):
    """||row||^2, <row, vector[row % mod]> and ||vector||^2 of the fp32 result per column half-tile, with no output stored at all (the
    prototype stage's reductions folded into the GEMM that makes the pooled features)."""
    M, N, K, mod = 520, 512, 576, 40
    A, B = _bf(_rand((M, K), 21)), _bf(_rand((N, K), 22))
    bias = _rand((N,), 23)
    V = _rand((mod, N), 24)
    tiles_n = 2
    stat = torch.full((M, 2 * tiles_n, 4), -1.0, device="cuda")
    g = Gemm()
    g.A, g.lda, g.ka = A.data_ptr(), K, K
    g.B, g.ldb, g.kb = B.data_ptr(), K, K
    g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair = M, N, K, 1, 1, 256, pair
    g.bias = bias.data_ptr()
    g.rowstat, g.dotvec, g.dot_ld, g.dot_mod = stat.data_ptr(), V.data_ptr(), N, mod
    _run(g)
    ref = A.double() @ B.double().T + bias.double()
    Vr = V.double()[torch.arange(M, device="cuda") % mod]
    _close(stat[:, :, 0].sum(1), (ref * ref).sum(1), 1e-5)
    _close(stat[:, :, 1].sum(1), (ref * Vr).sum(1), 2e-5)
    _close(stat[:, :, 2].sum(1), (Vr * Vr).sum(1), 1e-5)                 # ... and the vector's own squared norm
    for t in range(2 * tiles_n):                      # each entry covers its own 128 columns
        cols = slice(128 * t, 128 * t + 128)
        _close(stat[:, t, 0], (ref[:, cols] ** 2).sum(1), 1e-5)


@pytest.mark.parametrize("bn", [256, 128])
def test_row_statistics_few_rows_per_batch_item(bn):
    """The per-clip pooling shape (M = P = 40 rows per batch item, several tiles per CTA): the vectors' slices are read into
    shared memory once per CTA and reused by every tile with the same columns."""
    M, N, K, batch, mod = 40, 512, 64, 300, 40
    A, B = _bf(_rand((batch, M, K), 31)), _bf(_rand((batch, N, K), 32))
    V = _rand((mod, N), 33)
    tiles_n = N // bn
    stat = torch.full((batch, M, 2 * tiles_n, 4), -1.0, device="cuda")
    g = Gemm()
    g.A, g.lda, g.a_bs, g.a_batched, g.ka = A.data_ptr(), K, M * K, 1, K
    g.B, g.ldb, g.b_bs, g.b_batched, g.kb = B.data_ptr(), K, N * K, 1, K
    g.M, g.N, g.K, g.batch, g.npass, g.bn = M, N, K, batch, 1, bn
    g.rowstat, g.dotvec, g.dot_ld, g.dot_mod = stat.data_ptr(), V.data_ptr(), N, mod
    g.dot_early = 1 if bn == 128 else 0     # the vectors may be read before the kernel in front has completed
    _run(g)
    ref = torch.bmm(A.double(), B.double().transpose(1, 2))
    _close(stat[..., 0].sum(2), (ref * ref).sum(2), 1e-5)
    _close(stat[..., 1].sum(2), (ref * V.double()[None]).sum(2), 2e-5)
    _close(stat[..., 2].sum(2), (V.double() ** 2).sum(1)[None].expand(batch, M), 1e-5)
    half = bn // 2
    for t in range(2 * tiles_n):                      # each entry covers its own column half-tile
        cols = slice(half * t, half * t + half)
        _close(stat[:, :, t, 1], (ref[:, :, cols] * V.double()[None, :, cols]).sum(2), 2e-5)


@pytest.mark.parametrize("items", [11, 148 * 3 + 5])
def test_row_statistics_block_diagonal_groups(items):
    """Three small per-item GEMMs (40 rows, K = 49, both operands MN-major) share one 128-row tile: item j's rows sit at
    [40 j, 40 j + 40), its K rows are the tile's j-th k-block; the last group is ragged."""
    P, K, N, G = 40, 49, 256, 3
    At = _bf(_rand((items, K, P), 41))          # [item][k][m]: m contiguous
    Bt = _bf(_rand((items, K, N), 42))          # [item][k][n]: n contiguous
    V = _rand((P, N), 43)
    groups = (items + G - 1) // G
    tiles_n = N // 128
    stat = torch.full((items * P + 64, 2 * tiles_n, 4), -1.0, device="cuda")   # + guard rows that must stay untouched
    g = Gemm()
    g.A, g.lda, g.a_bs, g.a_batched, g.ka, g.a_mn_major, g.a_rows = At.data_ptr(), P, K * P, 1, P, 1, K
    g.B, g.ldb, g.b_bs, g.b_batched, g.kb, g.b_mn_major, g.b_rows = Bt.data_ptr(), N, K * N, 1, N, 1, K
    g.M, g.N, g.K, g.batch, g.npass, g.bn = G * P, N, K, groups, 1, 128
    g.group, g.group_rows, g.group_items, g.stat_rows = G, P, items, items * P
    g.rowstat, g.dotvec, g.dot_ld, g.dot_mod, g.dot_early = stat.data_ptr(), V.data_ptr(), N, P, 1
    _run(g)
    ref = torch.bmm(At.double().transpose(1, 2), Bt.double())           # [item][P][N]
    got = stat[:items * P].view(items, P, 2 * tiles_n, 4)
    _close(got[..., 0].sum(2), (ref * ref).sum(2), 1e-5)
    _close(got[..., 1].sum(2), (ref * V.double()[None]).sum(2), 2e-5)
    _close(got[..., 2].sum(2), (V.double() ** 2).sum(1)[None].expand(items, P), 1e-5)
    assert bool((stat[items * P:] == -1.0).all()), "rows past the last item were written"
