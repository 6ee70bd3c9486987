"""Timing of the tiled path's GEMM building block (csrc/tc_gemm.cu) through the debug hook: TFLOP/s per shape / mode."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protoasnet_b200 import _lib
from tests.test_tc_gemm_gpu import Gemm, Output, OUT_BF16, OUT_F32, OUT_BF16_HILO, ACT_RELU, ACT_ABS

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream


def run(g, reps=20):
    for _ in range(3):
        _lib.check(lib.pasn_debug_tc_gemm(C.byref(g), C.sizeof(Gemm), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.pasn_debug_tc_gemm(C.byref(g), C.sizeof(Gemm), st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def plain(M, N, K, bn, pair, out_mode=OUT_BF16, act=0, npass=1, trans_S=0, second=False, batch=1):
    ex = 2 if npass > 1 else 1
    A = torch.randn((batch * M, ex * K), device="cuda").bfloat16()
    B = torch.randn((N, ex * K), device="cuda").bfloat16()
    g = Gemm()
    g.A, g.lda, g.ka = A.data_ptr(), ex * K, ex * K
    g.B, g.ldb, g.kb = B.data_ptr(), ex * K, ex * K
    if batch > 1:
        g.a_batched, g.a_bs = 1, M * ex * K
    g.M, g.N, g.K, g.batch, g.npass, g.bn, g.pair, g.act = M, N, K, batch, npass, bn, pair, act
    if npass == 3:
        g.b_off[1] = K; g.a_off[2] = K
    Np = (N + 7) // 8 * 8
    keep = []
    if trans_S:
        o = torch.empty((batch * M // trans_S, N, trans_S), device="cuda", dtype=torch.bfloat16); keep.append(o)
        g.out[0] = Output(o.data_ptr(), OUT_BF16, trans_S, N * trans_S, 0, N, 0, trans_S)
        if second:
            o2 = torch.empty((batch * M, Np), device="cuda", dtype=torch.bfloat16); keep.append(o2)
            g.out[1] = Output(o2.data_ptr(), OUT_BF16, Np, 0, 0, N, 0, 0)
    else:
        width = {OUT_BF16: Np, OUT_F32: Np, OUT_BF16_HILO: 2 * Np}[out_mode]
        o = torch.empty((batch * M, width), device="cuda", dtype=torch.float32 if out_mode == OUT_F32 else torch.bfloat16); keep.append(o)
        g.out[0] = Output(o.data_ptr(), out_mode, width, M * width, Np, N, 0, 0)
    us = run(g)
    tf = 2.0 * batch * M * N * K * npass / us * 1e-6
    return us, tf, keep


cases = [
    ("layer 1 (cfg3 tokens)  M=200704 N=512 K=512 bf16 relu", dict(M=200704, N=512, K=512, bn=256, act=ACT_RELU)),
    ("same, 3-pass hi/lo -> hi|lo planes", dict(M=200704, N=512, K=512, bn=256, act=ACT_RELU, npass=3, out_mode=OUT_BF16_HILO)),
    ("G2        M=200704 N=128 K=256", dict(M=200704, N=128, K=256, bn=128, act=ACT_RELU)),
    ("O (cfg5)  M=4096 N=3136 K=256 batch 8, |.|", dict(M=4096, N=3136, K=256, bn=256, act=ACT_ABS, batch=8)),
    ("O token-major (cfg2) M=50176 N=40 K=256 -> TMA rows", dict(M=50176, N=40, K=256, bn=64, act=ACT_ABS)),
    ("same -> channel-major per clip (plain stores)", dict(M=50176, N=40, K=256, bn=64, act=ACT_ABS, trans_S=49)),
    ("same -> both outputs", dict(M=50176, N=40, K=256, bn=64, act=ACT_ABS, trans_S=49, second=True)),
    ("square    M=8192 N=8192 K=8192 fp32 out", dict(M=8192, N=8192, K=8192, bn=256, out_mode=OUT_F32)),
]
for name, kw in cases:
    for pair in (0, 1):
        if kw["bn"] < 128 and pair:
            continue
        us, tf, _ = plain(pair=pair, **kw)
        print(f"{name:62s} pair={pair}  {us:9.1f} us  {tf:8.1f} TFLOP/s", flush=True)
