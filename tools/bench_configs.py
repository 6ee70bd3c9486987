"""Head forward time for every BASELINE config shape (not the headline: cfg 3 at batch 1024 is bench.py's job).
Prints one line per case and writes gpurun_out/bench_configs.json when that directory exists.  FLOPs are the reference
formulation's; the fraction is against the measured burst bf16 peak."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from protoasnet_b200 import _lib, synth
from tests.util import build_model

lib = _lib.load()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
A, G, F, T = _lib.PASN_PATH_AUTO, _lib.PASN_PATH_GENERIC, _lib.PASN_PATH_TCGEN05, _lib.PASN_PATH_TILED
cases = [("cfg1_video_yml", 8, torch.float32, A), ("cfg1_video_yml", 8, torch.bfloat16, A), ("cfg1_video_yml", 128, torch.bfloat16, A),
         ("cfg2_image", 150, torch.float32, A), ("cfg2_image", 150, torch.bfloat16, A), ("cfg2_image", 1024, torch.bfloat16, A),
         ("cfg2_image", 1024, torch.bfloat16, G),
         ("cfg3_video_b1024", 32, torch.bfloat16, A), ("cfg3_video_b1024", 1024, torch.bfloat16, A),
         ("cfg3_video_b1024", 1024, torch.bfloat16, T), ("cfg3_video_b1024", 1024, torch.float32, A),
         ("cfg3_video_b1024", 1024, torch.float32, G),
         ("cfg5_scaled", 8, torch.bfloat16, A), ("cfg5_scaled", 32, torch.bfloat16, A), ("cfg5_scaled", 8, torch.float32, A),
         ("cfg5_scaled", 8, torch.bfloat16, G)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0].startswith(sys.argv[1])]
rows = []
for name, n, dt, path in cases:
    dims = synth.CONFIGS[name]
    sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
    m = build_model(dims, sd, path=path)
    x = torch.relu(torch.randn((n, dims.C) + dims.spatial, device="cuda")).to(dt)
    d = m._rt.make_dims(x, m.kernel_path)[0]
    fam = "generic FFMA"
    if path != G and lib.pasn_tcgen05_supported(C.byref(d)):
        d2 = m._rt.make_dims(x, F)[0]
        fused = lib.pasn_tcgen05_supported(C.byref(d2)) and (path == F or (path == A and dt == torch.bfloat16))   # AUTO keeps fp32 maps exact: tiled hi/lo chain
        fam = "fused tcgen05" if fused else "tiled tcgen05"
        if fam == "tiled tcgen05" and dt == torch.float32:
            fam += " (hi/lo split)"
    S = dims.S
    flop = 2 * S * (dims.C * dims.D + dims.D * dims.D + dims.C * dims.D + dims.D * (dims.D // 2) + (dims.D // 2) * dims.P + dims.P * dims.D)
    with torch.no_grad():
        for _ in range(2):
            m(x)
        torch.cuda.synchronize()
        reps = 3 if fam == "generic FFMA" else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            m(x)
        e1.record()
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(reps):
            m.compute_occurence_map(x)
        e3.record()
        torch.cuda.synchronize()
    ms, ms_occ = e0.elapsed_time(e1) / reps, e2.elapsed_time(e3) / reps
    tf = n * flop / ms * 1e-9
    rows.append({"config": name, "N": n, "dtype": str(dt)[6:], "path": fam, "ms": ms, "clips_per_s": n / ms * 1e3, "tflops": tf,
                 "frac_of_burst": tf / peaks["bf16_tflops"], "compute_occurence_map_ms": ms_occ})
    print(f"{name:18s} N={n:5d} {str(dt)[6:]:9s} {fam:30s} {ms:9.3f} ms  {n / ms * 1e3:12.0f} clips/s  {tf:8.1f} TFLOP/s "
          f"({100 * tf / peaks['bf16_tflops']:5.1f} % of burst)   compute_occurence_map {ms_occ:8.3f} ms", flush=True)
    del m, x
    torch.cuda.empty_cache()
out = os.path.join(ROOT, "gpurun_out")
if os.path.isdir(out):
    json.dump(rows, open(os.path.join(out, "bench_configs.json"), "w"), indent=1)
