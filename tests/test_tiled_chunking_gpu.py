"""The tiled tensor-core chain walks large batches in chunks of clips (workspace bound).  With a forced small chunk
(PASN_TILED_CHUNK, read once per process: hence the subprocess) forward outputs must be bit-identical to the unchunked run
and the backward's gradients, which accumulate across chunks, must agree to fp32 summation order."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os, sys, numpy as np, torch
sys.path.insert(0, %r)
from protoasnet_b200 import _lib, synth
from tests.util import build_model
out = {}
for cfg, n, dt in (("cfg2_image", 23, torch.bfloat16), ("cfg3_video_b1024", 11, torch.float32)):
    dims = synth.CONFIGS[cfg]
    sd = synth.make_head_params(dims, seed=9, bias_scale=0.05, bf16_round=True)
    x = torch.from_numpy(synth.make_features(dims, n, seed=2, bf16_round=True)).cuda().to(dt)
    m = build_model(dims, sd, path=_lib.PASN_PATH_TILED)
    with torch.no_grad():
        f, d, o, l = m.push_forward(x)
        occ = m.compute_occurence_map(x)
    out[cfg + "_fwd"] = [t.float().cpu().numpy() for t in (f, d, o, l, occ)]
    m.autograd_mode = "kernel"; m.train()
    xg = x.clone().requires_grad_(True)
    lg, sim, oc = m(xg)
    g = torch.Generator(device="cuda").manual_seed(1)
    (lg * torch.randn(lg.shape, device="cuda", generator=g)).sum().add((sim * torch.randn(sim.shape, device="cuda", generator=g)).sum()).add(
        (oc.float() * 0.05 * torch.randn(oc.shape, device="cuda", generator=g)).sum()).backward()
    out[cfg + "_bwd"] = [xg.grad.float().cpu().numpy()] + [p.grad.float().cpu().numpy() for k, p in sorted(m.named_parameters()) if p.grad is not None]
np.save(sys.argv[1], np.array([out], dtype=object), allow_pickle=True)
""" % ROOT


def _run(tmp, chunk):
    env = dict(os.environ)
    if chunk:
        env["PASN_TILED_CHUNK"] = str(chunk)
    else:
        env.pop("PASN_TILED_CHUNK", None)
    subprocess.run([sys.executable, "-c", SCRIPT, tmp], check=True, env=env, cwd=ROOT, timeout=600)
    import numpy as np
    return np.load(tmp, allow_pickle=True)[0]


def test_chunked_runs_match_unchunked(tmp_path):
    import numpy as np
    a = _run(str(tmp_path / "a.npy"), 0)
    b = _run(str(tmp_path / "b.npy"), 4)
    for k in a:
        assert len(a[k]) == len(b[k]) and len(a[k]) >= 5
        for u, v in zip(a[k], b[k]):
            if k.endswith("_fwd"):
                assert np.array_equal(u, v), k
            else:
                scale = np.abs(u).max() + 1e-30
                assert np.abs(u - v).max() <= 2e-5 * scale, (k, float(np.abs(u - v).max() / scale))
