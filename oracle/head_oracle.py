"""CPU restatement of the reference prototype head (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Two independent restatements:

* ``*_torch``  -- op-for-op the reference's eager PyTorch sequence (1x1 conv -> ReLU -> ... ->
  broadcast multiply -> three sums -> nn.CosineSimilarity -> (.+1)/2 -> Linear), fp32, CPU.
  This is the parity oracle and the timed ``cpu_baseline`` (it does exactly the work the
  reference does on the host, including materialising the [N,P,D,*spatial] product).
* ``head_forward_f64`` -- numpy float64 einsum formulation, used as a "ground truth" to
  measure the rounding error of both fp32 implementations.

All citations are relative to /root/reference.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_COS_EPS = 1e-8  # nn.CosineSimilarity default eps (src/models/Video_XProtoNet.py:65)


def _conv1x1(x: torch.Tensor, w: torch.Tensor, b=None) -> torch.Tensor:
    """nn.Conv3d/Conv2d with kernel_size=1 (weights keep their (O,I,1,1[,1]) shape)."""
    if x.dim() == 5:
        return F.conv3d(x, w, b)
    return F.conv2d(x, w, b)


def add_on_torch(x: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """feature_map = add_on_layers(x): Conv -> ReLU -> Conv, no final activation.

    Video: src/models/Video_XProtoNet.py:27-39, :85.  Image: PPNet 'regular' add-on with the
    trailing Sigmoid stripped, src/models/ProtoPNet.py:117-130 + src/models/XProtoNet.py:17.
    """
    h = F.relu(_conv1x1(x, sd["add_on_layers.0.weight"], sd["add_on_layers.0.bias"]))
    return _conv1x1(h, sd["add_on_layers.2.weight"], sd["add_on_layers.2.bias"])


def occurrence_map_torch(x: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """|occurrence_module(x)| with a singleton dim inserted at 2 -> (N, P, 1, *spatial).

    src/models/Video_XProtoNet.py:42-62 (module), :106-109 (abs + unsqueeze); image twin
    src/models/XProtoNet.py:21-41, :82-85.
    """
    g = F.relu(_conv1x1(x, sd["occurrence_module.0.weight"], sd["occurrence_module.0.bias"]))
    g = F.relu(_conv1x1(g, sd["occurrence_module.2.weight"], sd["occurrence_module.2.bias"]))
    o = _conv1x1(g, sd["occurrence_module.4.weight"], None)
    return torch.abs(o).unsqueeze(2)


def pooled_features_torch(occ: torch.Tensor, fmap: torch.Tensor) -> torch.Tensor:
    """(occurrence_map * feature_map).sum(3).sum(3)[.sum(3)] -> (N, P, D).

    src/models/Video_XProtoNet.py:87 / :119; src/models/XProtoNet.py:56 / :95.  Materialises the
    broadcast product exactly like the reference.
    """
    prod = occ * fmap.unsqueeze(1)
    for _ in range(fmap.dim() - 2):
        prod = prod.sum(dim=3)
    return prod


def similarity_torch(feats: torch.Tensor, prototypes: torch.Tensor) -> torch.Tensor:
    """(CosineSimilarity(dim=2)(feats, prototypes.squeeze().unsqueeze(0)) + 1) / 2 -> (N, P).

    src/models/Video_XProtoNet.py:90-93.
    """
    pv = prototypes.reshape(prototypes.shape[0], prototypes.shape[1]).unsqueeze(0)
    cos = F.cosine_similarity(feats, pv, dim=2, eps=_COS_EPS)
    return (cos + 1) / 2.0


def head_forward_torch(x: torch.Tensor, sd: Dict[str, torch.Tensor]):
    """forward() minus the backbone -> (logits, similarity, occurrence_map).  Video_XProtoNet.py:82-98."""
    fmap = add_on_torch(x, sd)
    occ = occurrence_map_torch(x, sd)
    feats = pooled_features_torch(occ, fmap)
    sim = similarity_torch(feats, sd["prototype_vectors"])
    logits = F.linear(sim, sd["last_layer.weight"])
    return logits, sim, occ


def push_forward_torch(x: torch.Tensor, sd: Dict[str, torch.Tensor]):
    """push_forward() minus the backbone -> (features_extracted, 1 - similarity, occurrence_map, logits).

    src/models/Video_XProtoNet.py:111-130.
    """
    fmap = add_on_torch(x, sd)
    occ = occurrence_map_torch(x, sd)
    feats = pooled_features_torch(occ, fmap)
    sim = similarity_torch(feats, sd["prototype_vectors"])
    logits = F.linear(sim, sd["last_layer.weight"])
    return feats, 1 - sim, occ, logits


def push_forward_torch_pchunk(x: torch.Tensor, sd: Dict[str, torch.Tensor], p_chunk: int = 16):
    """push_forward_torch for shapes whose broadcast product does not fit in memory (BASELINE config 5: P=4096, D=512,
    S=3136 -> 26 GB per clip).  Same ops as the reference (Video_XProtoNet.py:111-130); the product
    ``occurrence_map * feature_map`` (:119) is materialised and reduced for ``p_chunk`` prototypes at a time, which does
    not change any individual sum (each (n,p,d) entry still reduces T, then H, then W of its own row)."""
    fmap = add_on_torch(x, sd)
    occ = occurrence_map_torch(x, sd)
    P = occ.shape[1]
    feats = torch.cat([pooled_features_torch(occ[:, i:i + p_chunk], fmap) for i in range(0, P, p_chunk)], dim=1)
    sim = similarity_torch(feats, sd["prototype_vectors"])
    logits = F.linear(sim, sd["last_layer.weight"])
    return feats, 1 - sim, occ, logits


def to_torch_sd(sd_np: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd_np.items()}


def push_forward_chunked(x: np.ndarray, sd_np: Dict[str, np.ndarray], batch: int = 5):
    """Run push_forward_torch over ``x`` in loader-sized batches (video push batch = 5,
    src/configs/Ours_ProtoASNet_Video.yml:25) and concatenate; returns numpy arrays."""
    sd = to_torch_sd(sd_np)
    outs = [[], [], [], []]
    with torch.no_grad():
        for i in range(0, x.shape[0], batch):
            r = push_forward_torch(torch.from_numpy(x[i : i + batch]), sd)
            for o, v in zip(outs, r):
                o.append(v.numpy())
    return tuple(np.concatenate(o, axis=0) for o in outs)


# ---------------------------------------------------------------------------------------------
# float64 ground truth (einsum formulation)
# ---------------------------------------------------------------------------------------------
def head_forward_f64(x: np.ndarray, sd: Dict[str, np.ndarray]):
    """Same mathematics in float64; returns dict of logits, similarity, occurrence_map [N,P,S],
    features_extracted, distance."""
    n, c = x.shape[:2]
    xs = x.reshape(n, c, -1).astype(np.float64)

    def w(name):
        a = sd[name].astype(np.float64)
        return a.reshape(a.shape[0], a.shape[1])

    def b(name):
        return sd[name].astype(np.float64)[None, :, None]

    h = np.maximum(np.einsum("oc,ncs->nos", w("add_on_layers.0.weight"), xs) + b("add_on_layers.0.bias"), 0)
    fmap = np.einsum("oc,ncs->nos", w("add_on_layers.2.weight"), h) + b("add_on_layers.2.bias")
    g = np.maximum(np.einsum("oc,ncs->nos", w("occurrence_module.0.weight"), xs) + b("occurrence_module.0.bias"), 0)
    g = np.maximum(np.einsum("oc,ncs->nos", w("occurrence_module.2.weight"), g) + b("occurrence_module.2.bias"), 0)
    occ = np.abs(np.einsum("oc,ncs->nos", w("occurrence_module.4.weight"), g))
    feats = np.einsum("nps,nds->npd", occ, fmap)
    pv = sd["prototype_vectors"].astype(np.float64)
    pv = pv.reshape(pv.shape[0], pv.shape[1])
    fn = np.maximum(np.linalg.norm(feats, axis=2, keepdims=True), _COS_EPS)
    vn = np.maximum(np.linalg.norm(pv, axis=1, keepdims=True), _COS_EPS)
    cos = np.einsum("npd,pd->np", feats / fn, pv / vn)
    sim = (cos + 1.0) / 2.0
    logits = sim @ sd["last_layer.weight"].astype(np.float64).T
    return {
        "logits": logits,
        "similarity": sim,
        "occurrence_map": occ,
        "features_extracted": feats,
        "distance": 1.0 - sim,
    }
