"""One warm forward of a BASELINE config on a chosen path (for ncu captures): run_cfg.py <config> <N> <bf16|fp32> [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protoasnet_b200 import _lib, synth
from tests.util import build_model
name, n, dt = sys.argv[1], int(sys.argv[2]), (torch.bfloat16 if sys.argv[3] == "bf16" else torch.float32)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dims = synth.CONFIGS[name]
sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
m = build_model(dims, sd)
x = torch.relu(torch.randn((n, dims.C) + dims.spatial, device="cuda")).to(dt)
with torch.no_grad():
    for _ in range(reps):
        m(x)
torch.cuda.synchronize()
