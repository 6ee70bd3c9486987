"""Forward + backward of the head through the library at the cfg-3 shape (autograd_mode='kernel')."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protoasnet_b200 import synth
from tests.util import build_model
dims = synth.CONFIGS["cfg3_video_b1024"]
sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
for n in (64, 256):
    for mode in ("kernel", "composite"):
        m = build_model(dims, sd); m.autograd_mode = mode
        x = torch.relu(torch.randn((n, dims.C) + dims.spatial, device="cuda"))
        def step():
            for p in m.parameters(): p.grad = None
            logits, sim, occ = m(x)
            (logits.sum() + sim.sum() + occ.float().abs().sum() * 1e-4).backward()
        for _ in range(2): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): step()
        e1.record(); torch.cuda.synchronize()
        print(f"N={n} {mode}: fwd+bwd {e0.elapsed_time(e1)/5:.2f} ms")
