"""Wall-clock (globaltimer, ns) stamps of CTA 0 of every GEMM of one tiled-chain forward: where a launch's fixed cost goes
(set-up, wait for the grid in front, first operand stage, first accumulator, last rows out, tear-down) and how large the
gaps between the launches are.  usage: trace_gemm.py [config] [N] [bf16|fp32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protoasnet_b200 import _lib, synth
from tests.util import build_model

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_image"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "fp32") else torch.bfloat16
dims = synth.CONFIGS[name]
sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
m = build_model(dims, sd, path=_lib.PASN_PATH_TILED)
x = torch.relu(torch.randn((n, dims.C) + dims.spatial, device="cuda")).to(dt)
lib = _lib.load()
buf = torch.zeros(64 * 64, dtype=torch.int64, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    lib.pasn_debug_set_gemm_trace(buf.data_ptr())
    m(x)
    torch.cuda.synchronize()
    lib.pasn_debug_set_gemm_trace(None)
t = buf.cpu().view(64, 64).tolist()
t0 = t[0][0]
names = ["entry", "set-up done", "grid in front done", "1st stage landed", "1st tile issued", "1st acc seen", "1st load out", "last acc seen",
         "last rows out", "stores drained", "exit", "block sync", "rows out w4", "rows out cta1", "drained cta1"]
print(f"{name} N={n} {str(dt)[6:]}: CTA 0 of each GEMM launch, ns since the first GEMM's entry")
print("  #  " + "".join(f"{s:>15s}" for s in names if s != "-") + "   span")
prev_exit = None
for i, row in enumerate(t):
    if not row[0]:
        break
    vals = [row[k] - t0 if row[k] else None for k in range(15)]
    line = "".join(f"{(v if v is not None else -1):15d}" for k, v in enumerate(vals) if names[k] != "-")
    gap = f"   gap to previous exit {row[0] - prev_exit:6d}" if prev_exit else ""
    print(f" {i:2d}  {line}   {row[10] - row[0]:6d}{gap}")
    prev_exit = row[10]
print("per tile of CTA 0 (first 16): operands landed / accumulator seen by the epilogue / rows out, ns since the launch's entry")
for i, row in enumerate(t):
    if not row[0]:
        break
    e = row[0]
    print(f" {i:2d}  " + "  ".join(f"{row[48 + k] - e}/{row[16 + k] - e}/{row[32 + k] - e}" for k in range(16) if row[16 + k]))
