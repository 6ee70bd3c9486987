"""Error of the tiled tensor-core path (and, for reference, the generic CUDA-core path) against the float64 oracle, per output,
for the BASELINE shapes: max |err| / max |ref| (of-scale) and the largest element-wise relative error among entries above 1 %
of the scale."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import head_oracle as ho
from protoasnet_b200 import _lib, synth
from tests.util import build_model

cases = [("cfg3_video_b1024", 6, torch.float32), ("cfg3_video_b1024", 6, torch.bfloat16), ("cfg2_image", 150, torch.bfloat16),
         ("cfg2_image", 20, torch.float32), ("cfg1_video_yml", 2, torch.float32)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0].startswith(sys.argv[1])]
for name, n, dt in cases:
    dims = synth.CONFIGS[name]
    bf = dt == torch.bfloat16
    sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=bf)
    x = synth.make_features(dims, n, seed=3, bf16_round=bf)
    ref = ho.head_forward_f64(x, sd)
    for path, pname in ((_lib.PASN_PATH_TILED, "tiled"), (_lib.PASN_PATH_GENERIC, "generic")):
        m = build_model(dims, sd, path=path)
        with torch.no_grad():
            f, d, occ, lg = m.push_forward(torch.from_numpy(x).cuda().to(dt))
            occ3 = m.compute_occurence_map(torch.from_numpy(x).cuda().to(dt))
        torch.cuda.synchronize()
        out = {"logits": lg, "distance": d, "features_extracted": f, "occurrence_map": occ.reshape(n, dims.P, -1),
               "compute_occurence_map": occ3.reshape(n, dims.P, -1)}
        line = []
        for k, v in out.items():
            r = ref["occurrence_map" if k == "compute_occurence_map" else k]
            g = v.float().cpu().numpy().astype(np.float64)
            e = np.abs(g - r)
            sc = np.abs(r).max()
            big = np.abs(r) > 0.01 * sc
            line.append(f"{k}: {e.max() / sc:.2e} of scale, rel {np.max(e[big] / np.abs(r[big])):.2e}")
        print(f"{name} N={n} {str(dt)[6:]} {pname}: " + "; ".join(line), flush=True)
