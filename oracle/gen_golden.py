"""Generate tests/golden/*.npz by running the REFERENCE code itself (build container only).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Run:  python -m oracle.gen_golden

* Head goldens: the reference ``Video_XProtoNet`` / ``XProtoNet`` classes
  (/root/reference/src/models/Video_XProtoNet.py, XProtoNet.py) are instantiated with a stub
  identity backbone (the class names satisfy ``PPNet.get_cnn_backbone_out_channels``,
  src/models/ProtoPNet.py:152-162), loaded with ``protoasnet_b200.synth`` parameters, and
  ``forward`` / ``push_forward`` are run on synthetic feature maps.
* Push goldens: the unmodified ``push_prototypes`` (/root/reference/src/utils/push_abs_revision.py:181-348)
  is executed on CPU.  Its plotting dependencies (matplotlib, moviepy, imageio) are absent in this
  image, so empty stand-in modules are registered for import only; ``prototype_plot`` is replaced by
  a no-op and ``Tensor.cuda`` by identity (the function hard-codes ``.cuda()``, :268/:346).
  None of this touches the arithmetic being pinned.

Inputs are NOT stored: they are re-derived from ``protoasnet_b200.synth`` (numpy PCG64) by the
tests; each fixture stores the recipe (json) plus the reference outputs.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, ROOT)
from protoasnet_b200 import synth  # noqa: E402


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ["matplotlib", "matplotlib.pyplot", "moviepy", "moviepy.video", "moviepy.video.io",
                 "moviepy.video.io.ImageSequenceClip", "moviepy.editor", "imageio"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.ImageSequenceClip = object
                sys.modules[name] = m
    from src.models.Video_XProtoNet import Video_XProtoNet
    from src.models.XProtoNet import XProtoNet
    import src.utils.push_abs_revision as push_mod

    return Video_XProtoNet, XProtoNet, push_mod


class Resnet2p1dStub(nn.Module):
    """Identity backbone whose repr contains RESNET2P1D and whose last Conv3d has C out-channels."""

    def __init__(self, C):
        super().__init__()
        self.marker = nn.Conv3d(1, C, 1)

    def forward(self, x):
        return x


class ResStub(nn.Module):
    """Identity backbone whose repr starts with RES and whose last Conv2d has C out-channels."""

    def __init__(self, C):
        super().__init__()
        self.marker = nn.Conv2d(1, C, 1)

    def forward(self, x):
        return x


def build_reference_model(dims: synth.HeadDims, sd_np):
    Video_XProtoNet, XProtoNet, _ = _import_reference()
    if dims.ndim == 3:
        m = Video_XProtoNet(cnn_backbone=Resnet2p1dStub(dims.C), img_size=112, prototype_shape=dims.prototype_shape,
                            proto_layer_rf_info=None, num_classes=dims.K, init_weights=True)
    else:
        m = XProtoNet(features=ResStub(dims.C), img_size=224, prototype_shape=dims.prototype_shape,
                      proto_layer_rf_info=None, num_classes=dims.K, init_weights=True,
                      prototype_activation_function="log", add_on_layers_type="regular")
    own = m.state_dict()
    for k, v in sd_np.items():
        assert k in own and tuple(own[k].shape) == tuple(v.shape), (k, own.get(k, None) is not None and own[k].shape, v.shape)
        own[k] = torch.from_numpy(v.copy())
    m.load_state_dict(own)
    m.eval()
    return m


HEAD_CASES = [
    # name, config, n, param kwargs, feature seed, bf16_round
    ("head_tiny_video", "tiny_video", 5, dict(seed=11, bias_scale=0.1, last_layer_noise=0.1), 3, False),
    ("head_tiny_image", "tiny_image", 6, dict(seed=12, bias_scale=0.1, incorrect_strength=-0.5), 4, False),
    ("head_odd_video", "odd_video", 3, dict(seed=13, bias_scale=0.05, last_layer_noise=0.2), 5, False),
    ("head_cfg3_fp32", "cfg3_video_b1024", 2, dict(seed=200, bias_scale=0.02), 0, False),
    ("head_cfg3_bf16in", "cfg3_video_b1024", 2, dict(seed=200, bias_scale=0.02, bf16_round=True), 0, True),
    ("head_cfg1_fp32", "cfg1_video_yml", 1, dict(seed=200), 0, False),
    ("head_cfg2_fp32", "cfg2_image", 3, dict(seed=200, bias_scale=0.02), 0, False),
]

PUSH_CASES = [
    # name, config, n_total, loader batch, param kwargs, abstain_class
    ("push_tiny_video", "tiny_video", 37, 5, dict(seed=21, bias_scale=0.1), True),
    ("push_tiny_video_noabstain", "tiny_video", 23, 4, dict(seed=22, bias_scale=0.1), False),
    ("push_tiny_image", "tiny_image", 31, 150, dict(seed=23, bias_scale=0.1), True),
    ("push_cfg3_bf16in", "cfg3_video_b1024", 45, 5, dict(seed=200, bias_scale=0.02, bf16_round=True), True),
]


def gen_head(name, cfg, n, pk, fseed, bf16_round):
    dims = synth.CONFIGS[cfg]
    sd = synth.make_head_params(dims, **pk)
    x = synth.make_features(dims, n, seed=fseed, bf16_round=bf16_round)
    m = build_reference_model(dims, sd)
    with torch.no_grad():
        xt = torch.from_numpy(x)
        logits, sim, occ = m(xt)
        feats, dist, occ2, logits2 = m.push_forward(xt)
        occ3 = m.compute_occurence_map(xt)
    assert torch.equal(occ, occ2) and torch.equal(occ, occ3) and torch.equal(logits, logits2)
    recipe = dict(config=cfg, n=n, params=pk, feature_seed=fseed, bf16_round=bf16_round)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), recipe=json.dumps(recipe),
        logits=logits.numpy(), similarity=sim.numpy(), occurrence_map=occ.numpy(),
        features_extracted=feats.numpy(), distance=dist.numpy(),
        input_checksum=np.float64(x.astype(np.float64).sum()),
    )
    print(f"{name}: logits {tuple(logits.shape)} sim {tuple(sim.shape)} occ {tuple(occ.shape)}")


def gen_push(name, cfg, n_total, batch, pk, abstain):
    _, _, push_mod = _import_reference()
    dims = synth.CONFIGS[cfg]
    sd = synth.make_head_params(dims, **pk)
    n_real = dims.K - 1 if abstain else dims.K
    labels = synth.push_labels(n_total, n_real, seed=7)
    x = synth.make_features(dims, n_total, seed=1000, bf16_round=pk.get("bf16_round", False))
    m = build_reference_model(dims, sd)
    loader = []
    for i in range(0, n_total, batch):
        loader.append({"cine": torch.from_numpy(x[i:i + batch]),
                       "target_AS": torch.from_numpy(labels[i:i + batch]),
                       "filename": [f"clip_{j}" for j in range(i, min(i + batch, n_total))]})
    push_mod.prototype_plot = lambda *a, **k: None
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with tempfile.TemporaryDirectory() as td:
            push_mod.push_prototypes(loader, m, class_specific=True, abstain_class=abstain,
                                     root_dir_for_saving_prototypes=td, epoch_number="golden",
                                     log=lambda *a, **k: None, replace_prototypes=True)
            import pickle
            with open(os.path.join(td, "epoch-golden", "prototypes_info.pickle"), "rb") as f:
                info = pickle.load(f)
    finally:
        torch.Tensor.cuda = orig_cuda
    new_protos = m.prototype_vectors.detach().numpy().copy()
    filenames = [str(s) for s in info["prototypes_filenames"]]
    win_idx = np.array([int(s.split("_")[1]) for s in filenames], dtype=np.int64)
    recipe = dict(config=cfg, n_total=n_total, batch=batch, params=pk, abstain_class=abstain, label_seed=7,
                  feature_seed=1000)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), recipe=json.dumps(recipe),
        new_prototype_vectors=new_protos, winner_index=win_idx,
        winner_similarity=np.asarray(info["prototypes_similarity_to_src_ROIs"], dtype=np.float64),
        winner_gts=np.asarray(info["prototypes_gts"], dtype=np.int64),
        winner_logits=np.asarray(info["prototypes_preds"], dtype=np.float32),
        winner_occurrence_maps=np.asarray(info["prototypes_occurrence_maps"], dtype=np.float32),
    )
    print(f"{name}: winners {win_idx.tolist()}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    for c in HEAD_CASES:
        gen_head(*c)
    for c in PUSH_CASES:
        gen_push(*c)


if __name__ == "__main__":
    main()
