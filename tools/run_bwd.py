"""A few forward+backward steps of the head through the library (for ncu launch lists): run_bwd.py <N> <bf16|fp32>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protoasnet_b200 import synth
from tests.util import build_model
n = int(sys.argv[1]); dt = torch.bfloat16 if sys.argv[2] == "bf16" else torch.float32
dims = synth.CONFIGS["cfg3_video_b1024"]
sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
m = build_model(dims, sd); m.autograd_mode = "kernel"
x = torch.relu(torch.randn((n, dims.C) + dims.spatial, device="cuda")).to(dt).requires_grad_(True)
for _ in range(2):
    for p in m.parameters(): p.grad = None
    logits, sim, occ = m(x)
    (logits.sum() + sim.sum() + occ.float().abs().sum() * 1e-4).backward()
torch.cuda.synchronize()
