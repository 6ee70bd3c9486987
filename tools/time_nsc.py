"""Fused head at N=1024 (cfg 3): NCDHW vs channels_last_3d feature maps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from protoasnet_b200 import _lib, synth
from tests.util import build_model
dims = synth.CONFIGS["cfg3_video_b1024"]
sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
m = build_model(dims, sd, path=_lib.PASN_PATH_TCGEN05)
x = torch.relu(torch.randn((1024, dims.C) + dims.spatial, device="cuda")).bfloat16()
lib = _lib.load()
for name, xx in (("NCDHW", x), ("channels_last_3d", x.contiguous(memory_format=torch.channels_last_3d))):
    with torch.no_grad():
        for _ in range(3): m(xx)
        torch.cuda.synchronize()
        lib.pasn_debug_time_main_kernel(1)
        ks = []
        for _ in range(10):
            m(xx); ks.append(lib.pasn_debug_last_main_kernel_ms())
        lib.pasn_debug_time_main_kernel(0)
    print(f"{name}: K1 {np.mean(ks)*1e3:.1f} us")
