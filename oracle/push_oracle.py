"""CPU restatement of the reference push / prototype-projection loop.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows, step by step,
``/root/reference/src/utils/push_abs_revision.py``:

  :226-239  prototype -> class via argmax(prototype_class_identity); every prototype is
            class-specific except the abstention slice [K*P/num_classes : P] when
            ``abstain_class`` (K = num_classes - 1 real classes, asserted >= 2)
  :242      running best distance initialised to +inf (float64 accumulator, fp32 distances)
  :288-307  per loader batch, per prototype j: mask samples whose label differs from the
            prototype's class (class-specific only), skip the batch if everything is masked,
            take the batch minimum, and if it is ``<=`` the running best take the batch argmin
            (lowest index inside the batch) and stash features_extracted[a, j]
  :342-346  stack the stashed vectors, reshape to prototype_shape, overwrite prototype_vectors (fp32)

Tie rules: inside a batch ``np.argmin`` returns the lowest index; ACROSS batches the reference's
``<=`` lets the LATER batch win an exact tie.  The product contract (BASELINE.json north_star) is
"lowest global index wins"; ``tie_rule='lowest'`` restates the loop with ``<``.  The two rules give
identical results unless two clips in different batches have bit-identical fp32 distances for
the same prototype; ``tests`` assert agreement on every data set used for parity.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Optional, Tuple

import numpy as np


def prototype_classes(prototype_class_identity: np.ndarray) -> np.ndarray:
    """push_abs_revision.py:228."""
    return np.argmax(prototype_class_identity, axis=1)


def class_specific_mask(P: int, num_classes: int, class_specific: bool = True, abstain_class: bool = True) -> np.ndarray:
    """push_abs_revision.py:229-239."""
    spec = np.full(P, class_specific)
    if abstain_class:
        K = num_classes - 1
        assert K >= 2, "Abstention-push must have >= 2 classes not including abstain"
        per = P // num_classes
        spec[K * per : P] = False
    return spec


def push_scan(
    batches: Iterable[Tuple[np.ndarray, np.ndarray, np.ndarray]],
    prototype_class_identity: np.ndarray,
    num_classes: int,
    class_specific: bool = True,
    abstain_class: bool = True,
    tie_rule: str = "lowest",
):
    """Running class-restricted argmin over loader batches.

    ``batches`` yields (features_extracted [B,P,D] fp32, distance [B,P] fp32, labels [B] int).
    Returns (best_dist [P] float64, best_index [P] int64 global index or -1, best_vec [P,D] fp32 or NaN rows).
    """
    assert tie_rule in ("lowest", "reference")
    P = prototype_class_identity.shape[0]
    cls = prototype_classes(prototype_class_identity)
    spec = class_specific_mask(P, num_classes, class_specific, abstain_class)
    best = np.full(P, np.inf)
    best_idx = np.full(P, -1, dtype=np.int64)
    best_vec = None
    offset = 0
    for feats, dist, gt in batches:
        if best_vec is None:
            best_vec = np.full((P, feats.shape[2]), np.nan, dtype=np.float32)
        for j in range(P):
            dj = dist[:, j]
            if spec[j]:
                dj = np.ma.masked_array(dj, gt != cls[j])
                if dj.mask.all():
                    continue
            m = np.amin(dj)
            better = (m <= best[j]) if tie_rule == "reference" else (m < best[j])
            if better:
                a = int(np.argmin(dj))
                best[j] = m
                best_idx[j] = offset + a
                best_vec[j] = feats[a, j]
        offset += dist.shape[0]
    return best, best_idx, best_vec


def push_prototypes_oracle(
    features: np.ndarray,
    labels: np.ndarray,
    sd: Dict[str, np.ndarray],
    num_classes: int,
    batch: int = 5,
    class_specific: bool = True,
    abstain_class: bool = True,
    tie_rule: str = "lowest",
    push_forward: Optional[Callable] = None,
):
    """Whole push over an in-memory feature set [N,C,*spatial]; returns
    (new_prototype_vectors shaped like sd['prototype_vectors'], best_index, best_dist).

    Prototypes whose class never appears keep their old vector and report index -1 (the
    reference crashes in that case, push_abs_revision.py:343-346 -- SURVEY.md §7 'Empty class').
    """
    from .head_oracle import push_forward_torch, to_torch_sd
    import torch

    tsd = to_torch_sd(sd)
    pf = push_forward or push_forward_torch
    P = sd["prototype_vectors"].shape[0]
    # one-hot class map of the reference (src/models/ProtoPNet.py:326-340): prototype j belongs to class j // (P / K);
    # restated here so that the oracle does not lean on product code
    assert P % num_classes == 0, "num_prototypes must be divisible by num_classes"
    ident = np.zeros((P, num_classes), dtype=np.float32)
    ident[np.arange(P), np.arange(P) // (P // num_classes)] = 1.0

    def gen():
        with torch.no_grad():
            for i in range(0, features.shape[0], batch):
                f, d, _occ, _lg = pf(torch.from_numpy(features[i : i + batch]), tsd)
                yield f.numpy(), d.numpy(), labels[i : i + batch]

    best, idx, vec = push_scan(gen(), ident, num_classes, class_specific, abstain_class, tie_rule)
    old = sd["prototype_vectors"].reshape(P, -1)
    new = np.where(idx[:, None] >= 0, vec, old).astype(np.float32)
    return new.reshape(sd["prototype_vectors"].shape), idx, best
