"""Soak of the tiled tensor-core chain and the tensor-core backward: repeated calls must be bit-identical (forward) /
identical up to atomic summation order (backward), finite, and never raise the sticky fault word."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protoasnet_b200 import _lib, synth
from tests.util import build_model
lib = _lib.load()
t0 = time.time()
bad = 0
for cfg, n, dt in (("cfg2_image", 300, torch.bfloat16), ("cfg2_image", 37, torch.float32), ("cfg5_scaled", 3, torch.bfloat16),
                   ("cfg3_video_b1024", 97, torch.float32), ("cfg1_video_yml", 5, torch.float32)):
    dims = synth.CONFIGS[cfg]
    sd = synth.make_head_params(dims, seed=3, bias_scale=0.05, bf16_round=True)
    m = build_model(dims, sd, path=_lib.PASN_PATH_TILED)
    x = torch.relu(torch.randn((n, dims.C) + dims.spatial, device="cuda")).to(dt)
    for xx, name in ((x, "ncs"), (x.contiguous(memory_format=torch.channels_last_3d if dims.ndim == 3 else torch.channels_last), "nsc")):
        with torch.no_grad():
            ref = m.push_forward(xx)
            reff = m(xx)      # plain forward: row statistics instead of stored features where the path uses them
            for i in range(40):
                out = m.push_forward(xx)
                outf = m(xx)
                if not all(torch.equal(a, b) for a, b in zip(out, ref)) or not all(torch.equal(a, b) for a, b in zip(outf, reff)):
                    bad += 1; print("FORWARD MISMATCH", cfg, n, name, i); break
            if float((reff[1] - (1 - ref[1])).abs().max()) > 2e-6 * (1 if dt == torch.float32 else 50):
                bad += 1; print("forward / push_forward similarity differ", cfg, n, name, float((reff[1] - (1 - ref[1])).abs().max()))
        print(cfg, n, dt, name, "forward ok", f"{time.time() - t0:.0f}s", flush=True)
    if cfg == "cfg5_scaled":
        continue
    m.autograd_mode = "kernel"; m.train()
    gref = None
    for i in range(15):
        for p in m.parameters(): p.grad = None
        xg = x.clone().requires_grad_(True)
        lg, sim, oc = m(xg)
        (lg.sum() + (sim * sim).sum() + oc.float().sum() * 1e-3).backward()
        g = torch.cat([xg.grad.float().flatten()] + [p.grad.float().flatten() for p in m.parameters() if p.grad is not None])
        if not bool(torch.isfinite(g).all()):
            bad += 1; print("NON-FINITE GRADIENT", cfg, i); break
        if gref is None:
            gref = g
        elif float((g - gref).abs().max()) > 1e-4 * float(gref.abs().max()):
            bad += 1; print("BACKWARD MISMATCH", cfg, i, float((g - gref).abs().max() / gref.abs().max())); break
    print(cfg, n, dt, "backward ok", f"{time.time() - t0:.0f}s", flush=True)
print("fault word:", lib.pasn_debug_fault(), " soak done, problems:", bad)
