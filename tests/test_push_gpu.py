"""GPU parity of push / prototype projection: winner indices bit-exact, pushed prototypes within tolerance."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import push_oracle as po
from protoasnet_b200 import push as pushmod
from protoasnet_b200 import synth
from tests.util import BF16_RTOL, FP32_RTOL, assert_close, build_model, load_golden

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
    assert_close(m.prototype_vectors.data, z["new_prototype_vectors"], FP32_RTOL, "prototype_vectors")
    assert_close(1 - res["distance"].double(), z["winner_similarity"], FP32_RTOL, "winner similarity")
    import pickle
    info = pickle.load(open(os.path.join(str(tmp_path), "epoch-t", "prototypes_info.pickle"), "rb"))
    assert np.array_equal(info["prototypes_gts"], z["winner_gts"])
    assert_close(info["prototypes_preds"], z["winner_logits"], FP32_RTOL, "winner logits")
    assert_close(info["prototypes_occurrence_maps"], z["winner_occurrence_maps"], FP32_RTOL, "winner occ maps")
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
    assert_close(m.prototype_vectors.data, z["new_prototype_vectors"], FP32_RTOL, "prototype_vectors")


def test_push_bf16_features_match_oracle_indices():
    z, r, dims, sd, labels, x = _case(os.path.join(GOLDEN, "push_cfg3_bf16in.npz"))
    m = build_model(dims, sd)
    res = pushmod.push_resident(m, torch.from_numpy(x).cuda().bfloat16(), torch.from_numpy(labels).cuda(), chunk=16)
    assert np.array_equal(res["index"].cpu().numpy(), z["winner_index"])
    assert_close(m.prototype_vectors.data, z["new_prototype_vectors"], 4e-3, "prototype_vectors (bf16 mode)")


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
