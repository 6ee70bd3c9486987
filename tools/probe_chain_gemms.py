"""Times the GEMMs of the tiled chain at the image-head shape (config 2, N = 1024) one by one through the library's test
hook, with their outputs switched on and off: which part of a short GEMM's time is the main loop, which the stores.
usage: probe_chain_gemms.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from protoasnet_b200 import _lib
from tests.test_tc_gemm_gpu import Gemm, Output, OUT_NONE, OUT_BF16, OUT_F32, ACT_RELU, ACT_ABS, ACT_NONE

lib = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
NB, S, P, Cc, D = 1024, 49, 40, 512, 512
T = NB * S
bf = torch.bfloat16
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(name, g, flop, nbytes):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        _lib.check(lib.pasn_debug_tc_gemm(C.byref(g), C.sizeof(Gemm), st), name)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()                      # cold L2, as inside the chain (every GEMM reads what another kernel wrote > L2 ago)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.pasn_debug_tc_gemm(C.byref(g), C.sizeof(Gemm), st), name)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    print(f"{name:58s} {us:8.1f} us   {flop / us * 1e-6:8.1f} TFLOP/s   {nbytes / us * 1e-3:8.1f} GB/s", flush=True)
    assert lib.pasn_debug_fault() == 0


X = torch.randn((T, Cc), device=dev).to(bf)
W13 = torch.randn((2 * D, Cc), device=dev).to(bf)
b13 = torch.randn((2 * D,), device=dev)
Y = torch.empty((T, 2 * D), device=dev, dtype=bf)
W4 = torch.randn((D // 2, D), device=dev).to(bf)
G2 = torch.empty((T, D // 2), device=dev, dtype=bf)
W5 = torch.randn((P, D // 2), device=dev).to(bf)
OCCT = torch.empty((T, 64), device=dev, dtype=bf)
OCCM = torch.empty((NB, P, S), device=dev, dtype=bf)
W2 = torch.randn((D, D), device=dev).to(bf)
F = torch.empty((T, D), device=dev, dtype=bf)
FE = torch.empty((NB, P, D), device=dev)

# (A) layer 1
for pair in (1, 0):
    for store in (1, 0):
        g = Gemm()
        g.A, g.lda, g.ka = X.data_ptr(), Cc, Cc
        g.B, g.ldb, g.kb = W13.data_ptr(), Cc, Cc
        g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair = T, 2 * D, Cc, 1, 1, 256, pair
        g.bias, g.act = b13.data_ptr(), ACT_RELU
        if store:
            g.out[0] = Output(Y.data_ptr(), OUT_BF16, 2 * D, 0, 0, 0, 0, 0)
        run(f"A  [H1|G1] {T}x{2 * D}x{Cc} pair={pair} store={store}", g, 2 * T * 2 * D * Cc, T * Cc * 2 + store * T * 2 * D * 2)

# (B) G2
for pair in (1, 0):
    for store in (1, 0):
        g = Gemm()
        g.A, g.lda, g.ka = Y.data_ptr() + D * 2, 2 * D, D
        g.B, g.ldb, g.kb = W4.data_ptr(), D, D
        g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair = T, D // 2, D, 1, 1, 256, pair
        g.act = ACT_RELU
        if store:
            g.out[0] = Output(G2.data_ptr(), OUT_BF16, D // 2, 0, 0, 0, 0, 0)
        run(f"B  G2 {T}x{D // 2}x{D} pair={pair} store={store}", g, 2 * T * (D // 2) * D, T * D * 2 + store * T * D)

# (C) O, token-major
for mode in ("both", "map only", "copy only", "none"):
    g = Gemm()
    g.A, g.lda, g.ka = G2.data_ptr(), D // 2, D // 2
    g.B, g.ldb, g.kb = W5.data_ptr(), D // 2, D // 2
    g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair = T, P, D // 2, 1, 1, 64, 0
    g.act = ACT_ABS
    no = 0
    if mode in ("both", "map only"):
        g.out[no] = Output(OCCM.data_ptr(), OUT_BF16, S, P * S, 0, P, 0, S); no += 1
    if mode in ("both", "copy only"):
        g.out[no] = Output(OCCT.data_ptr(), OUT_BF16, 64, 0, 64, P, 0, 0); no += 1
    run(f"C  O {T}x{P}x{D // 2} outputs: {mode}", g, 2 * T * P * (D // 2), T * D)

# (D') F
for store in (1, 0):
    g = Gemm()
    g.A, g.lda, g.ka = Y.data_ptr(), 2 * D, 2 * D
    g.B, g.ldb, g.kb = W2.data_ptr(), D, D
    g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair = T, D, D, 1, 1, 256, 1
    g.bias = b13.data_ptr()
    if store:
        g.out[0] = Output(F.data_ptr(), OUT_BF16, D, 0, D, 0, 0, 0)
    run(f"D' F {T}x{D}x{D} pair=1 store={store}", g, 2 * T * D * D, T * D * 2 + store * T * D * 2)

# pooling per clip
for store in (1, 0):
    for bn in (256, 128):
        g = Gemm()
        g.A, g.lda, g.a_bs, g.a_batched, g.ka, g.a_mn_major, g.a_rows = OCCT.data_ptr(), 64, S * 64, 1, 64, 1, S
        g.B, g.ldb, g.b_bs, g.b_batched, g.kb, g.b_mn_major, g.b_rows = F.data_ptr(), D, S * D, 1, D, 1, S
        g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair = P, D, S, NB, 1, bn, 0
        if store:
            g.out[0] = Output(FE.data_ptr(), OUT_F32, D, P * D, 0, 0, 0, 0)
        run(f"pool FE[n] {P}x{D}x{S} x{NB} bn={bn} store={store}", g, 2 * NB * P * D * S, T * D * 2 + store * NB * P * D * 4)

# fixed cost of a launch against its per-tile cost: the O GEMM over 1, 148, 296, 444 tiles, no outputs
for tiles in (1, 148, 296, 444, 888):
    g = Gemm()
    g.A, g.lda, g.ka = G2.data_ptr(), D // 2, D // 2
    g.B, g.ldb, g.kb = W5.data_ptr(), D // 2, D // 2
    g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair = min(T, tiles * 128), P, D // 2, 1, 1, 64, 0
    g.act = ACT_ABS
    run(f"C  O, {tiles} tiles, no outputs", g, 2 * g.M * P * (D // 2), g.M * D)
for tiles in (1, 148, 296):
    g = Gemm()
    g.A, g.lda, g.ka = Y.data_ptr() + D * 2, 2 * D, D
    g.B, g.ldb, g.kb = W4.data_ptr(), D, D
    g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair = min(T, tiles * 128), D // 2, D, 1, 1, 256, 0
    g.act = ACT_RELU
    run(f"B  G2, {tiles} tiles, no outputs", g, 2 * g.M * (D // 2) * D, g.M * D * 2)
