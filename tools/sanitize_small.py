"""Small fused-path run for compute-sanitizer (memcheck): ragged shapes, both tile orders."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protoasnet_b200 import _lib, synth
from tests.util import build_model
lib = _lib.load()
for (C, P, K, spatial, n) in [(128, 12, 3, (1, 10, 14), 5), (512, 40, 4, (4, 7, 7), 9)]:
    dims = synth.HeadDims(C, 256, P, K, spatial)
    sd = synth.make_head_params(dims, seed=41, bias_scale=0.05, bf16_round=True)
    m = build_model(dims, sd, path=_lib.PASN_PATH_TCGEN05)
    x = torch.from_numpy(synth.make_features(dims, n, seed=23, bf16_round=True)).cuda().bfloat16()
    for variant in (1, 2):
        lib.pasn_debug_set_k1_variant(variant)
        with torch.no_grad():
            out = m.push_forward(x)
        torch.cuda.synchronize()
        print(C, P, n, "variant", variant, "ok", float(out[1].sum()))
