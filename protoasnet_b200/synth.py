"""Deterministic synthetic inputs for the prototype-head / push hot path.

Everything is generated with numpy's PCG64 (stream-stable across numpy versions by
policy), never with torch's RNG, so that the committed golden fixtures under
``tests/golden/`` (produced in the build container by running the *reference* classes
on these inputs, see ``oracle/gen_golden.py``) can be re-derived on the GPU box where
``/root/reference`` does not exist.

Distributions follow the reference's initialisers:
  * conv weights  ~ kaiming_normal_(fan_out, relu)   -> N(0, sqrt(2 / out_channels)) for 1x1(x1) kernels
    (reference: src/models/ProtoPNet.py:313-320)
  * biases        = 0 in the reference; tests use ``bias_scale > 0`` so biases are exercised
  * prototypes    ~ U[0, 1)                           (reference: src/models/Video_XProtoNet.py:68)
  * last layer    = 1 on the prototype's own class, ``incorrect_strength`` elsewhere
    (reference: src/models/ProtoPNet.py:299-311, called with 0 at Video_XProtoNet.py:80)
  * features      = relu(N(0,1)) (mimics a post-ReLU backbone output)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np


@dataclass(frozen=True)
class HeadDims:
    """Problem dimensions of one prototype head (names follow SURVEY.md §8)."""

    C: int  # backbone output channels
    D: int  # prototype_shape[1]
    P: int  # prototype_shape[0]
    K: int  # num_classes (incl. abstention)
    spatial: Tuple[int, ...]  # (T, H, W) for video, (H, W) for image

    @property
    def S(self) -> int:
        s = 1
        for v in self.spatial:
            s *= v
        return s

    @property
    def ndim(self) -> int:
        return len(self.spatial)

    @property
    def prototype_shape(self) -> Tuple[int, ...]:
        return (self.P, self.D) + (1,) * self.ndim


# The five BASELINE.json configs (SURVEY.md §8 shape table) + small test shapes.
CONFIGS: Dict[str, HeadDims] = {
    "cfg1_video_yml": HeadDims(C=256, D=256, P=40, K=4, spatial=(8, 14, 14)),
    "cfg2_image": HeadDims(C=512, D=512, P=40, K=4, spatial=(7, 7)),
    "cfg3_video_b1024": HeadDims(C=512, D=256, P=40, K=4, spatial=(4, 7, 7)),
    "cfg5_scaled": HeadDims(C=512, D=512, P=4096, K=4, spatial=(16, 14, 14)),
    "tiny_video": HeadDims(C=24, D=16, P=8, K=4, spatial=(2, 3, 3)),
    "tiny_image": HeadDims(C=20, D=12, P=8, K=4, spatial=(5, 3)),
    "odd_video": HeadDims(C=40, D=24, P=12, K=3, spatial=(3, 5, 2)),
}


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(int(seed)))


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32, in numpy (matches torch's .bfloat16())."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounding = ((u >> 16) & 1) + 0x7FFF
    u = ((u + rounding) >> 16) << 16
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


def make_head_params(
    dims: HeadDims,
    seed: int = 200,
    bias_scale: float = 0.0,
    incorrect_strength: float = 0.0,
    last_layer_noise: float = 0.0,
    bf16_round: bool = False,
) -> Dict[str, np.ndarray]:
    """Parameter set of one head, keyed by the reference's ``state_dict`` names.

    Shapes/names: src/models/Video_XProtoNet.py:27-80 (video), src/models/XProtoNet.py:17-49 (image).
    """
    g = _rng(seed)
    C, D, P, K = dims.C, dims.D, dims.P, dims.K
    ones = (1,) * dims.ndim

    def conv(o, i):
        return (g.standard_normal((o, i), dtype=np.float32) * np.float32(np.sqrt(2.0 / o))).reshape((o, i) + ones)

    def bias(o):
        if bias_scale == 0.0:
            return np.zeros((o,), dtype=np.float32)
        return (g.standard_normal((o,), dtype=np.float32) * np.float32(bias_scale)).astype(np.float32)

    sd = {}
    sd["add_on_layers.0.weight"] = conv(D, C)
    sd["add_on_layers.0.bias"] = bias(D)
    sd["add_on_layers.2.weight"] = conv(D, D)
    sd["add_on_layers.2.bias"] = bias(D)
    sd["occurrence_module.0.weight"] = conv(D, C)
    sd["occurrence_module.0.bias"] = bias(D)
    sd["occurrence_module.2.weight"] = conv(D // 2, D)
    sd["occurrence_module.2.bias"] = bias(D // 2)
    sd["occurrence_module.4.weight"] = conv(P, D // 2)
    sd["prototype_vectors"] = g.random((P, D), dtype=np.float32).reshape(dims.prototype_shape)
    sd["ones"] = np.ones(dims.prototype_shape, dtype=np.float32)
    ident = prototype_class_identity(P, K)
    ll = ident.T * 1.0 + (1.0 - ident.T) * incorrect_strength
    if last_layer_noise:
        ll = ll + last_layer_noise * g.standard_normal(ll.shape)
    sd["last_layer.weight"] = ll.astype(np.float32)
    if bf16_round:
        for k in list(sd):
            if k.endswith("weight") or k.endswith("bias"):
                sd[k] = round_to_bf16(sd[k])
    return sd


def prototype_class_identity(P: int, K: int) -> np.ndarray:
    """One-hot (P, K): prototype j belongs to class j // (P/K). Reference: src/models/ProtoPNet.py:326-340."""
    if P % K != 0:
        raise AssertionError("num_prototypes must be divisible by num_classes")
    ident = np.zeros((P, K), dtype=np.float32)
    per = P // K
    for j in range(P):
        ident[j, j // per] = 1.0
    return ident


def make_features(dims: HeadDims, n: int, seed: int = 0, bf16_round: bool = False) -> np.ndarray:
    """Feature map [n, C, *spatial] = relu(N(0,1)), fp32 (optionally bf16-rounded)."""
    g = _rng(seed)
    x = g.standard_normal((n, dims.C) + tuple(dims.spatial), dtype=np.float32)
    np.maximum(x, 0.0, out=x)
    return round_to_bf16(x) if bf16_round else x


# ---------------------------------------------------------------------------------------------
# Push data set: chunked so that any sharding of the global index range sees identical clips.
# ---------------------------------------------------------------------------------------------
PUSH_CHUNK = 1000  # clips per generation chunk (SURVEY.md §8d cfg 4)


def push_chunk_features(dims: HeadDims, chunk_id: int, bf16_round: bool = True, chunk: int = PUSH_CHUNK) -> np.ndarray:
    return make_features(dims, chunk, seed=1000 + int(chunk_id), bf16_round=bf16_round)


def push_labels(n_total: int, num_real_classes: int, seed: int = 7) -> np.ndarray:
    """int64 labels ~ U{0..num_real_classes-1}, every class guaranteed present (if n_total allows)."""
    g = _rng(seed)
    y = g.integers(0, num_real_classes, size=(n_total,), dtype=np.int64)
    for c in range(min(num_real_classes, n_total)):
        y[c] = c
    return y


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ownership [lo, hi) of the unshuffled set."""
    per = -(-n_total // world)
    lo = min(rank * per, n_total)
    hi = min(lo + per, n_total)
    return lo, hi
