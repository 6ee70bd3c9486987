import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protoasnet_b200 import _lib, synth
from tests.util import build_model
dims = synth.CONFIGS["cfg3_video_b1024"]
sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
m = build_model(dims, sd, path=_lib.PASN_PATH_TCGEN05)
x = torch.relu(torch.randn((1024, dims.C) + dims.spatial, device="cuda")).bfloat16()
with torch.no_grad():
    for _ in range(5): m(x)
torch.cuda.synchronize()
