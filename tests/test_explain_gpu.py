"""Explainability data products (protoasnet_b200.explain) against the CPU oracle of push_forward: same keys, shapes and
values as the reference sweep (src/utils/explainability_utils.py:49-131), pickle round trip included."""
import os

import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from protoasnet_b200 import explain, synth
from tests.util import FP32_RTOL, assert_close, build_model

pytestmark = pytest.mark.gpu


class _Set(torch.utils.data.Dataset):
    def __init__(self, x, y):
        self.x, self.y = torch.from_numpy(x), torch.from_numpy(y)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return {"cine": self.x[i], "target_AS": self.y[i], "filename": f"clip_{i}"}


@pytest.mark.parametrize("abstain", [True, False])
def test_model_products_match_oracle(tmp_path, abstain):
    dims = synth.CONFIGS["tiny_video"]
    sd = synth.make_head_params(dims, seed=11, bias_scale=0.05, last_layer_noise=0.1)
    n = 23
    x = synth.make_features(dims, n, seed=4)
    y = synth.push_labels(n, dims.K - 1, seed=9)
    m = build_model(dims, sd)
    loader = torch.utils.data.DataLoader(_Set(x, y), batch_size=5, shuffle=False)
    cfg = {"view": "all", "frames": 32, "img_size": 112, "interval_quant": 1.0, "interval_unit": "cycle",
           "iterate_intervals": True, "dataset_root": str(tmp_path / "data")}
    logs = []
    data_dict, prod = explain.load_data_and_model_products(m, loader, "val", cfg, abstain, str(tmp_path / "run"), log=logs.append)
    with torch.no_grad():
        rf, rd, ro, rl = ho.push_forward_torch(torch.from_numpy(x), ho.to_torch_sd(sd))
        k = dims.K - 1 if abstain else dims.K
        rp = rl[:, :k].softmax(dim=1)
    assert set(data_dict) == {"inputs", "ys_gt", "filenames"}
    assert set(prod) == {"fc_layer_weights", "protoL_input_", "proto_dist_", "occurrence_map_", "ys_pred"}
    assert data_dict["inputs"].shape == x.shape and np.array_equal(data_dict["inputs"], x)
    assert np.array_equal(data_dict["ys_gt"], y) and data_dict["filenames"] == [f"clip_{i}" for i in range(n)]
    assert prod["occurrence_map_"].shape == (n, dims.P, 1) + dims.spatial
    assert_close(prod["protoL_input_"], rf.numpy(), FP32_RTOL, "protoL_input_")
    assert_close(prod["proto_dist_"], rd.numpy(), FP32_RTOL, "proto_dist_")
    assert_close(prod["occurrence_map_"], ro.numpy(), FP32_RTOL, "occurrence_map_")
    assert_close(prod["ys_pred"], rp.numpy(), FP32_RTOL, "ys_pred")
    assert np.array_equal(prod["fc_layer_weights"], sd["last_layer.weight"])
    # second call loads the pickles instead of running the model
    p1, p2 = explain.product_paths("val", cfg, str(tmp_path / "run"))
    assert os.path.exists(p1) and os.path.exists(p2)
    d2, m2 = explain.load_data_and_model_products(None, None, "val", cfg, abstain, str(tmp_path / "run"), log=logs.append)
    assert np.array_equal(m2["proto_dist_"], prod["proto_dist_"]) and d2["filenames"] == data_dict["filenames"]
    assert any("f1 score" in str(l) for l in logs)
