"""GPU parity of push / prototype projection: winner indices bit-exact, pushed prototypes within tolerance."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import push_oracle as po
from protoasnet_b200 import push as pushmod
from protoasnet_b200 import synth
from tests.util import BF16_RTOL, FP32_RTOL, FP32_TC_FEAT_ATOL, assert_close, build_model, feat_atol, load_golden

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PUSH = sorted(glob.glob(os.path.join(GOLDEN, "push_*.npz")))


class _Set(torch.utils.data.Dataset):
    def __init__(self, x, y):
        self.x, self.y = torch.from_numpy(x), torch.from_numpy(y)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return {"cine": self.x[i], "target_AS": self.y[i], "filename": f"clip_{i}"}


def _case(path):
    z, r = load_golden(path)
    dims = synth.CONFIGS[r["config"]]
    sd = synth.make_head_params(dims, **r["params"])
    n_real = dims.K - 1 if r["abstain_class"] else dims.K
    labels = synth.push_labels(r["n_total"], n_real, seed=r["label_seed"])
    x = synth.make_features(dims, r["n_total"], seed=r["feature_seed"], bf16_round=r["params"].get("bf16_round", False))
    return z, r, dims, sd, labels, x


@pytest.mark.parametrize("path", PUSH, ids=[os.path.basename(p)[:-4] for p in PUSH])
def test_push_prototypes_matches_reference_golden(path, tmp_path):
    z, r, dims, sd, labels, x = _case(path)
    m = build_model(dims, sd)
    loader = torch.utils.data.DataLoader(_Set(x, labels), batch_size=r["batch"], shuffle=False)
    res = pushmod.push_prototypes(loader, m, class_specific=True, abstain_class=r["abstain_class"],
                                  root_dir_for_saving_prototypes=str(tmp_path), epoch_number="t", log=lambda *a: None)
    assert np.array_equal(res["index"].cpu().numpy(), z["winner_index"])           # bit-exact indices
    fa = feat_atol(m, torch.from_numpy(x[:1]).cuda())
    assert_close(m.prototype_vectors.data, z["new_prototype_vectors"], FP32_RTOL, "prototype_vectors", atol_frac=fa)
    assert_close(1 - res["distance"].double(), z["winner_similarity"], FP32_RTOL, "winner similarity")
    import pickle
    info = pickle.load(open(os.path.join(str(tmp_path), "epoch-t", "prototypes_info.pickle"), "rb"))
    assert np.array_equal(info["prototypes_gts"], z["winner_gts"])
    assert_close(info["prototypes_preds"], z["winner_logits"], FP32_RTOL, "winner logits")
    assert_close(info["prototypes_occurrence_maps"], z["winner_occurrence_maps"], FP32_RTOL, "winner occ maps", atol_frac=fa)
    assert [str(s) for s in info["prototypes_filenames"]] == [f"clip_{i}" for i in z["winner_index"]]


@pytest.mark.parametrize("path", PUSH, ids=[os.path.basename(p)[:-4] for p in PUSH])
@pytest.mark.parametrize("chunk", [7, 4096])
def test_push_resident_matches_golden_and_no_replace(path, chunk):
    z, r, dims, sd, labels, x = _case(path)
    m = build_model(dims, sd)
    before = m.prototype_vectors.data.clone()
    xg, yg = torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda()
    res = pushmod.push_resident(m, xg, yg, chunk=chunk, abstain_class=r["abstain_class"], replace_prototypes=False)
    assert np.array_equal(res["index"].cpu().numpy(), z["winner_index"])
    assert torch.equal(m.prototype_vectors.data, before)                           # replace_prototypes=False
    pushmod.push_resident(m, xg, yg, chunk=chunk, abstain_class=r["abstain_class"], replace_prototypes=True)
    assert_close(m.prototype_vectors.data, z["new_prototype_vectors"], FP32_RTOL, "prototype_vectors", atol_frac=feat_atol(m, xg))


def test_push_bf16_features_match_oracle_indices():
    z, r, dims, sd, labels, x = _case(os.path.join(GOLDEN, "push_cfg3_bf16in.npz"))
    m = build_model(dims, sd)
    res = pushmod.push_resident(m, torch.from_numpy(x).cuda().bfloat16(), torch.from_numpy(labels).cuda(), chunk=16)
    assert np.array_equal(res["index"].cpu().numpy(), z["winner_index"])
    assert_close(m.prototype_vectors.data, z["new_prototype_vectors"], 4e-3, "prototype_vectors (bf16 mode)")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_push_on_the_tiled_path_matches_oracle(dtype):
    """Push over image-head feature maps (BASELINE config 2 shape: served by the tiled tensor-core chain, whose prototype
    stage folds the running argmin and captures the winners) against the oracle's loop: indices exact, vectors close."""
    dims = synth.CONFIGS["cfg2_image"]
    bf = dtype == torch.bfloat16
    sd = synth.make_head_params(dims, seed=77, bias_scale=0.05, bf16_round=bf)
    n = 120
    x = synth.make_features(dims, n, seed=13, bf16_round=bf)
    labels = synth.push_labels(n, dims.K - 1, seed=3)
    m = build_model(dims, sd)
    res = pushmod.push_resident(m, torch.from_numpy(x).cuda().to(dtype), torch.from_numpy(labels).cuda(), chunk=50)
    new, ref_idx, ref_d = po.push_prototypes_oracle(x, labels, sd, dims.K, batch=40, tie_rule="lowest")
    assert np.array_equal(res["index"].cpu().numpy(), ref_idx)
    tol = 4e-3 if bf else FP32_RTOL
    assert_close(res["distance"], ref_d, BF16_RTOL if bf else FP32_RTOL, "winner distances", atol_frac=BF16_RTOL if bf else FP32_RTOL)
    assert_close(m.prototype_vectors.data, new, tol, "pushed prototype_vectors",
                 atol_frac=1e-3 if bf else feat_atol(m, torch.from_numpy(x).cuda()))


def test_push_ties_break_to_lowest_global_index_and_empty_class():
    dims = synth.CONFIGS["tiny_video"]
    sd = synth.make_head_params(dims, seed=5, bias_scale=0.1)
    base = synth.make_features(dims, 3, seed=2)
    x = np.concatenate([base, base, base[:1]], axis=0)          # clips 3,4,5,6 duplicate clips 0,1,2,0
    labels = np.array([0, 1, 0, 0, 1, 0, 0], dtype=np.int64)    # class 2 never occurs
    m = build_model(dims, sd)
    old = m.prototype_vectors.data.clone()
    res = pushmod.push_resident(m, torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), chunk=2)
    idx = res["index"].cpu().numpy()
    new, ref_idx, _ = po.push_prototypes_oracle(x, labels, sd, dims.K, batch=2, tie_rule="lowest")
    assert np.array_equal(idx, ref_idx)
    assert np.all(idx[idx >= 0] < 3)                            # duplicates never win: lowest global index
    per = dims.P // dims.K
    assert np.all(idx[2 * per:3 * per] == -1)                   # empty class -> -1, prototype kept
    assert torch.equal(m.prototype_vectors.data[2 * per:3 * per], old[2 * per:3 * per])
    assert_close(m.prototype_vectors.data, new, FP32_RTOL, "prototypes")


def test_push_full_size_properties():
    """cfg-4 shape at a size the GPU finishes in well under a second (8192 clips, bf16): re-sharding the set into
    2 / 3 'ranks' and merging keys gives the same winners as one pass; pushing twice is idempotent."""
    dims = synth.CONFIGS["cfg3_video_b1024"]
    sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
    m = build_model(dims, sd)
    n = 8192
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.relu(torch.randn((n, dims.C) + dims.spatial, device="cuda", generator=g)).bfloat16()
    y = torch.randint(0, 3, (n,), device="cuda", generator=g)
    pc = pushmod.proto_class_restriction(m).cuda()
    one = pushmod.new_best_key(dims.P, "cuda")
    m.push_scan(x, y, pc, 0, one, backbone=False)
    for world in (2, 3):
        keys = []
        for rk in range(world):
            lo, hi = synth.shard_range(n, rk, world)
            k = pushmod.new_best_key(dims.P, "cuda")
            for i in range(lo, hi, 1000):
                j = min(i + 1000, hi)
                m.push_scan(x[i:j], y[i:j], pc, i, k, backbone=False)
            keys.append(k)
        merged = torch.stack(keys).min(dim=0).values      # signed int64 order == key order (include/pasn.h)
        # winners are identical for any sharding; the fp32 distance may differ in its last bits because the fused
        # kernel's pooling order depends on where a clip falls in the 128-voxel tiling of its batch
        idx_m, dist_m = pushmod.decode_keys(merged)
        idx_1, dist_1 = pushmod.decode_keys(one)
        assert torch.equal(idx_m, idx_1)
        assert float((dist_m - dist_1).abs().max()) < 1e-5
    idx, dist = pushmod.decode_keys(one)
    assert torch.all(y[idx[:30]] == pc[:30].long())             # class restriction holds
    res1 = pushmod.push_resident(m, x, y, chunk=2048)
    assert torch.equal(res1["index"], idx)
    res2 = pushmod.push_resident(m, x, y, chunk=2048)           # after the overwrite every winner is still its own nearest
    assert torch.equal(res2["index"], idx)
    assert float(res2["distance"].max()) < 1e-3                  # and sits at distance ~0 from its prototype


def _bench_push_set(dims, n_total, dev="cuda"):
    """The data set bench.py pushes over (cfg 4): chunk c = global clips [1000c, 1000c+1000), relu(N(0,1)) from the device
    generator seeded 1000 + c, bf16; labels from synth.push_labels(seed 7)."""
    feats = torch.empty((n_total, dims.C) + dims.spatial, dtype=torch.bfloat16, device=dev)
    for c in range(-(-n_total // synth.PUSH_CHUNK)):
        gg = torch.Generator(device=dev).manual_seed(1000 + c)
        chunk = torch.relu(torch.randn((synth.PUSH_CHUNK, dims.C) + dims.spatial, device=dev, generator=gg)).bfloat16()
        lo, hi = c * synth.PUSH_CHUNK, min(n_total, (c + 1) * synth.PUSH_CHUNK)
        feats[lo:hi] = chunk[: hi - lo]
    labels = synth.push_labels(n_total, dims.K - 1, seed=7)
    return feats, labels


@pytest.mark.parametrize("n_total", [50000])
def test_push_bench_size_matches_oracle_indices(n_total):
    """BASELINE config 4 at its real size: the 50 000-clip set bench.py times, winners against the CPU oracle
    (reference loop of push_abs_revision.py:288-307 over the reference head in fp32 on the same bf16 inputs and
    bf16-rounded weights).  Indices must be bit-exact under both tie rules; the top-2 margin of every prototype is
    reported.  The oracle takes ~40 s on the box's host cores."""
    import json
    import time
    from oracle import head_oracle as ho

    dims = synth.CONFIGS["cfg3_video_b1024"]
    sd = synth.make_head_params(dims, seed=200, bias_scale=0.02, bf16_round=True)
    m = build_model(dims, sd)
    feats, labels = _bench_push_set(dims, n_total)
    yg = torch.from_numpy(labels).cuda()
    res = pushmod.push_resident(m, feats, yg, chunk=8192, replace_prototypes=False)
    idx_gpu = res["index"].cpu().numpy()

    torch.set_num_threads(os.cpu_count() or 1)
    tsd = ho.to_torch_sd(sd)
    dist_all = np.empty((n_total, dims.P), dtype=np.float32)
    t0 = time.time()

    def batches(bs=50):
        with torch.no_grad():
            for i in range(0, n_total, bs):
                xb = feats[i:i + bs].float().cpu()
                f, d, _o, _l = ho.push_forward_torch(xb, tsd)
                dist_all[i:i + bs] = d.numpy()
                yield f.numpy(), d.numpy(), labels[i:i + bs]

    ident = synth.prototype_class_identity(dims.P, dims.K)
    best, idx_ref, _vec = po.push_scan(batches(), ident, dims.K, True, True, "lowest")
    oracle_s = time.time() - t0
    # the reference's `<=` rule over the same distances (cross-batch ties -> later batch)
    _b2, idx_ref2, _v2 = po.push_scan(((np.zeros((min(50, n_total - i), dims.P, 1), np.float32), dist_all[i:i + 50], labels[i:i + 50])
                                       for i in range(0, n_total, 50)), ident, dims.K, True, True, "reference")
    assert np.array_equal(idx_ref, idx_ref2), "tie rules disagree on this data set"
    # top-2 margin per prototype under the class mask
    cls = po.prototype_classes(ident)
    spec = po.class_specific_mask(dims.P, dims.K)
    margins = []
    for j in range(dims.P):
        dj = dist_all[:, j].astype(np.float64)
        if spec[j]:
            dj = np.where(labels == cls[j], dj, np.inf)
        two = np.partition(dj, 1)[:2]
        margins.append(float(two[1] - two[0]))
    dmax = 0.0
    with torch.no_grad():
        for i in range(0, n_total, 2048):
            _f, dg, _o, _l = m.push_forward(feats[i:i + 2048])
            dmax = max(dmax, float(np.abs(dg.cpu().numpy().astype(np.float64) - dist_all[i:i + 2048]).max()))
    report = {"n_total": n_total, "oracle_seconds": oracle_s, "max_abs_distance_error_vs_oracle": dmax, "min_top2_margin": min(margins), "margins": margins,
              "gpu_index": idx_gpu.tolist(), "oracle_index": idx_ref.tolist(),
              "mismatches": [int(j) for j in np.nonzero(idx_gpu != idx_ref)[0]]}
    out = os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out")
    if os.path.isdir(out):
        json.dump(report, open(os.path.join(out, f"push_parity_{n_total}.json"), "w"))
    print(f"push parity at {n_total}: max |d_gpu - d_oracle| {dmax:.3e}, min top-2 margin {min(margins):.3e}, mismatches {report['mismatches']}, oracle {oracle_s:.0f} s")
    assert np.array_equal(idx_gpu, idx_ref), f"winner indices differ for prototypes {report['mismatches']} (margins {[margins[j] for j in report['mismatches']]})"


def test_agent_push_wrapper_matches_reference_golden(tmp_path):
    """XProtoNet_Base.push (src/agents/XProtoNet_Base.py:149-167): loader = data_loaders['train_push'], save dir
    <save_dir>/img/epoch-<current_epoch>_pushed, abstain_class from the config, replace_prototypes passed through."""
    import pickle
    z, r, dims, sd, labels, x = _case(os.path.join(GOLDEN, "push_tiny_video.npz"))
    m = build_model(dims, sd)

    class _Agent:
        current_epoch = 7
        model = m
        config = {"abstain_class": r["abstain_class"], "save_dir": str(tmp_path)}
        data_loaders = {"train_push": torch.utils.data.DataLoader(_Set(x, labels), batch_size=r["batch"], shuffle=False)}

    before = m.prototype_vectors.data.clone()
    res = pushmod.agent_push(_Agent(), replace_prototypes=False)
    assert np.array_equal(res["index"].cpu().numpy(), z["winner_index"])
    assert torch.equal(m.prototype_vectors.data, before)
    d = os.path.join(str(tmp_path), "img", "epoch-7_pushed")
    info = pickle.load(open(os.path.join(d, "prototypes_info.pickle"), "rb"))
    assert np.array_equal(info["prototypes_gts"], z["winner_gts"])
    pushmod.agent_push(_Agent())                                                    # default: prototypes replaced
    assert_close(m.prototype_vectors.data, z["new_prototype_vectors"], FP32_RTOL, "prototype_vectors",
                 atol_frac=feat_atol(m, torch.from_numpy(x[:1]).cuda()))


class _RandomWindowSet(torch.utils.data.Dataset):
    """Like the reference's train_push set (src/data/as_dataloader.py:246-255): every __getitem__ draws a fresh random
    window, so a clip can never be fetched twice.  Everything handed out is logged."""

    def __init__(self, x, y):
        self.x, self.y = torch.from_numpy(x), torch.from_numpy(y)
        self.handed = {}
        self.g = torch.Generator().manual_seed(123)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        assert i not in self.handed, "push must see every clip exactly once"
        clip = self.x[i] * (0.5 + torch.rand(self.x[i].shape, generator=self.g))    # never the same twice
        self.handed[i] = clip.clone()
        return {"cine": clip, "target_AS": self.y[i], "filename": f"clip_{i}"}


@pytest.mark.parametrize("cfg,n,batch", [("tiny_video", 37, 5), ("cfg3_video_b1024", 45, 8)])
def test_push_captures_winners_in_the_pass_that_saw_them(cfg, n, batch, tmp_path):
    """push_abs_revision.py:299-307: the winner's vector / occurrence map / logits / clip come from the forward pass that
    produced the minimum.  With a loader that draws a random window per __getitem__, prototype_vectors must equal the
    pooled features of the clip as it was drawn, and every clip is fetched exactly once."""
    import pickle
    dims = synth.CONFIGS[cfg]
    sd = synth.make_head_params(dims, seed=11, bias_scale=0.05, bf16_round=(cfg != "tiny_video"))
    x = synth.make_features(dims, n, seed=4)
    labels = synth.push_labels(n, dims.K - 1, seed=9)
    m = build_model(dims, sd)
    ds = _RandomWindowSet(x, labels)
    loader = torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=False)
    before = m.prototype_vectors.data.clone()
    res = pushmod.push_prototypes(loader, m, root_dir_for_saving_prototypes=str(tmp_path), epoch_number="r",
                                  log=lambda *a: None, replace_prototypes=False)
    assert len(ds.handed) == n
    assert torch.equal(m.prototype_vectors.data, before)
    idx = res["index"].cpu().numpy()
    assert (idx >= 0).all()
    info = pickle.load(open(os.path.join(str(tmp_path), "epoch-r", "prototypes_info.pickle"), "rb"))
    drawn = torch.stack([ds.handed[int(i)] for i in idx]).cuda()
    with torch.no_grad():
        feats, dist, occ, logits = m.push_forward(drawn)          # row j = the clip that won prototype j, as drawn
    ar = torch.arange(dims.P, device="cuda")
    tol = 1e-5 if cfg == "tiny_video" else 1e-4                   # fused path: pooling order depends on the tile position
    assert_close(res["features"], feats[ar, ar].cpu().numpy(), tol, "captured vector == pooled features of the drawn winner")
    assert_close(res["distance"], dist[ar, ar].cpu().numpy(), tol, "winner distance", atol_frac=tol)
    assert np.array_equal(info["prototypes_src_imgs"], drawn.cpu().numpy())
    assert_close(info["prototypes_occurrence_maps"], occ[ar, ar].float().cpu().numpy(), tol, "winner occurrence maps")
    assert_close(info["prototypes_preds"], logits.cpu().numpy(), tol, "winner logits")
    assert [str(s) for s in info["prototypes_filenames"]] == [f"clip_{i}" for i in idx]
    # and the running argmin really is the minimum over what was drawn
    allx = torch.stack([ds.handed[i] for i in range(n)]).cuda()
    with torch.no_grad():
        _f, d_all, _o, _l = m.push_forward(allx)
    pc = pushmod.proto_class_restriction(m).cuda()
    yg = torch.from_numpy(labels).cuda()
    masked = torch.where((pc[None, :] < 0) | (pc[None, :].long() == yg[:, None]), d_all, torch.full_like(d_all, float("inf")))
    if cfg == "tiny_video":
        assert np.array_equal(masked.argmin(dim=0).cpu().numpy(), idx)
