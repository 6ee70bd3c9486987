"""Diagnostic (GPU): fused tcgen05 path vs generic CUDA path vs CPU oracle on cfg-3 / cfg-1 shaped inputs."""
import os
import sys

os.environ["PASN_DEBUG_SYNC"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import head_oracle as ho  # noqa: E402
from protoasnet_b200 import _lib, synth  # noqa: E402
from tests.util import build_model  # noqa: E402


def rel(a, b):
    a = a.detach().float().cpu().numpy().astype(np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def run(cfg, n, seed=0):
    dims = synth.CONFIGS[cfg]
    sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
    x = synth.make_features(dims, n, seed=seed, bf16_round=True)
    xg = torch.from_numpy(x).cuda().bfloat16()
    mg = build_model(dims, sd, path=_lib.PASN_PATH_GENERIC)
    mt = build_model(dims, sd, path=_lib.PASN_PATH_TCGEN05)
    with torch.no_grad():
        fg, dg, og, lg = mg.push_forward(xg)
        try:
            ft, dt, ot, lt = mt.push_forward(xg)
        except Exception as e:
            print(f"{cfg} n={n}: tcgen05 FAILED: {e}")
            return
        torch.cuda.synchronize()
    nref = min(n, 6)
    with torch.no_grad():
        rf, rd, ro, rl = ho.push_forward_torch(torch.from_numpy(x[:nref]), ho.to_torch_sd(sd))
    print(f"{cfg} n={n}: tc-vs-generic feats {rel(ft, fg.cpu().numpy()):.2e} dist {rel(dt, dg.cpu().numpy()):.2e} "
          f"occ {rel(ot, og.float().cpu().numpy()):.2e} logits {rel(lt, lg.cpu().numpy()):.2e} | tc-vs-oracle[:{nref}] "
          f"feats {rel(ft[:nref], rf.numpy()):.2e} sim {rel(1 - dt[:nref], (1 - rd).numpy()):.2e} "
          f"logits {rel(lt[:nref], rl.numpy()):.2e} occ {rel(ot[:nref], ro.numpy()):.2e} | nan={bool(torch.isnan(ft).any())}")
    bad = (ft - fg).abs().amax(dim=(1, 2)) / fg.abs().amax()
    worst = torch.topk(bad, min(5, n))
    print("   worst clips:", [(int(i), f"{float(v):.1e}") for v, i in zip(worst.values, worst.indices)])


if __name__ == "__main__":
    cases = (("cfg3_video_b1024", 1), ("cfg3_video_b1024", 2), ("cfg3_video_b1024", 7), ("cfg3_video_b1024", 150),
             ("cfg3_video_b1024", 1024), ("cfg1_video_yml", 3))
    for cfg, n in cases:
        run(cfg, n)
    dims = synth.CONFIGS["cfg3_video_b1024"]
    sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
    m = build_model(dims, sd, path=_lib.PASN_PATH_TCGEN05)
    x = torch.relu(torch.randn((1024, dims.C) + dims.spatial, device="cuda")).bfloat16()
    import protoasnet_b200.head as H

    H._DEBUG_SYNC = False
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"tcgen05 head fwd N=1024: {ms:.3f} ms/step -> {1024 / ms * 1e3:.0f} clips/s")
