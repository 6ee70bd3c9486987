"""Head forward time for every BASELINE config shape (not the headline: cfg 3 at batch 1024 is bench.py's job)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from protoasnet_b200 import _lib, synth
from tests.util import build_model
lib = _lib.load()
cases = [("cfg1_video_yml", 8, torch.float32), ("cfg1_video_yml", 8, torch.bfloat16), ("cfg1_video_yml", 128, torch.bfloat16),
         ("cfg2_image", 150, torch.float32), ("cfg2_image", 1024, torch.bfloat16),
         ("cfg3_video_b1024", 32, torch.bfloat16), ("cfg3_video_b1024", 1024, torch.bfloat16), ("cfg3_video_b1024", 1024, torch.float32),
         ("cfg5_scaled", 8, torch.bfloat16)]
for name, n, dt in cases:
    dims = synth.CONFIGS[name]
    sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
    m = build_model(dims, sd)
    x = torch.relu(torch.randn((n, dims.C) + dims.spatial, device="cuda")).to(dt)
    d = m._rt.make_dims(x, m.kernel_path)[0]
    fused = bool(lib.pasn_tcgen05_supported(C.byref(d)))
    S = 1
    for v in dims.spatial: S *= v
    flop = 2 * S * (dims.C * dims.D + dims.D * dims.D + dims.C * dims.D + dims.D * (dims.D // 2) + (dims.D // 2) * dims.P + dims.P * dims.D)
    with torch.no_grad():
        for _ in range(2): m(x)
        torch.cuda.synchronize()
        reps = 3 if not fused else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): m(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:18s} N={n:5d} {str(dt)[6:]:9s} {'fused tcgen05' if fused else 'generic fp32 FFMA':18s} {ms:9.3f} ms  {n / ms * 1e3:12.0f} clips/s  {n * flop / ms * 1e-9:8.1f} TFLOP/s")
