"""Push / prototype projection on device, sharded across ranks.

Mirrors ``push_prototypes`` of the reference (src/utils/push_abs_revision.py:181-348) and the agent wrapper
``XProtoNet_Base.push`` (src/agents/XProtoNet_Base.py:149-167), re-designed for B200:

  pass 1  every rank walks its contiguous shard of the unshuffled ``train_push`` set; per batch ONE fused launch
          computes the head and folds the class-restricted running argmin into packed keys
          ``orderable(fp32 dist) << 32 | global index`` kept on device (no per-batch D2H of features, distances,
          occurrence maps or input clips as at push_abs_revision.py:278-285).
  merge   one NCCL all-reduce(MIN) over the P packed keys (ties -> lowest global index).
  pass 2  the owner rank of each winner re-runs ``push_forward`` on just those <= P clips to obtain the winning
          ``features_extracted`` rows and the side data the reference pickles (occurrence map, logits, label,
          filename, input clip; push_abs_revision.py:301-325); one all-reduce(SUM) of the [P,D] fp32 winner rows
          (non-owners contribute exact zeros), then every rank overwrites ``prototype_vectors`` identically
          (push_abs_revision.py:342-346).

Defined behaviour where the reference crashes: a class-specific prototype whose class never occurs keeps its old
vector and reports index -1 (SURVEY.md section 7 'Empty class').  Rendering of prototype images / mp4s
(push_abs_revision.py:13-178, :329-340) is out of scope; the pickle with the same keys is written when a save
directory is given.
"""
from __future__ import annotations

import os
import pickle
import time
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib

_SIGN = -(1 << 63)


def proto_class_restriction(model, class_specific: bool = True, abstain_class: bool = True) -> torch.Tensor:
    """int32 [P]: class each prototype is restricted to, -1 = unrestricted.  push_abs_revision.py:226-239."""
    P = model.num_prototypes
    cls = torch.argmax(model.prototype_class_identity, dim=1).to(torch.int32)
    spec = torch.full((P,), bool(class_specific))
    if abstain_class:
        K = model.num_classes - 1
        assert K >= 2, "Abstention-push must have >= 2 classes not including abstain"
        per = P // model.num_classes
        spec[K * per: P] = False
    return torch.where(spec, cls, torch.full_like(cls, -1))


def new_best_key(P: int, device) -> torch.Tensor:
    key = torch.empty(P, dtype=torch.int64, device=device)  # storage for uint64 keys
    lib = _lib.load()
    with torch.cuda.device(device):
        _lib.check(lib.pasn_push_init(key.data_ptr(), P, torch.cuda.current_stream(device).cuda_stream), "pasn_push_init")
    return key


def merge_keys(best_key: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce(MIN) of the packed keys.  They are stored with the top bit flipped (include/pasn.h), so the signed
    int64 order NCCL / gloo reduce with is the key order: lowest distance first, ties to the lowest global index."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(best_key, op=dist.ReduceOp.MIN, group=group)
    return best_key


_KEY_NONE = (1 << 63) - 1   # INT64_MAX: no candidate


def decode_keys(best_key: torch.Tensor):
    """-> (index int64 [P] (-1 = no candidate), distance fp32 [P])."""
    if not best_key.is_cuda:  # host-side bookkeeping only (gloo tests of the merge logic); no head compute here
        none = best_key == _KEY_NONE
        ku = best_key ^ _SIGN                      # back to the unsigned key bit pattern (held in an int64)
        o = (ku >> 32) & 0xFFFFFFFF
        u = torch.where((o & 0x80000000) != 0, o & 0x7FFFFFFF, (~o) & 0xFFFFFFFF)
        u = torch.where(u >= (1 << 31), u - (1 << 32), u).to(torch.int32)
        d = torch.where(none, torch.full_like(u, 0x7F800000), u).view(torch.float32)
        return torch.where(none, torch.full_like(best_key, -1), ku & 0xFFFFFFFF), d
    lib = _lib.load()
    P = best_key.numel()
    idx = torch.empty(P, dtype=torch.int64, device=best_key.device)
    d = torch.empty(P, dtype=torch.float32, device=best_key.device)
    with torch.cuda.device(best_key.device):
        _lib.check(lib.pasn_push_decode(best_key.data_ptr(), P, idx.data_ptr(), d.data_ptr(),
                                        torch.cuda.current_stream(best_key.device).cuda_stream), "pasn_push_decode")
    return idx, d


def _world(group=None):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def finish_push(model, best_key, fetch_features, lo: int, hi: int, replace_prototypes=True, group=None, dense=False):
    """Merge keys across ranks, gather winner rows from their owners, overwrite prototypes.

    ``fetch_features(global_indices: LongTensor on device) -> backbone feature maps`` for indices in [lo, hi).
    ``dense=True`` (features resident on the GPU): no host synchronisation at all -- every rank re-runs ``push_forward``
    on exactly P clips (its winners, or a clamped in-range clip where it is not the owner) and masks by ownership.
    ``dense=False`` (clips come from a Dataset): only the distinct local winners are fetched.
    Returns dict(index, distance, features (P,D), side) -- ``side`` carries the winners' occurrence maps / logits.
    """
    import torch.distributed as dist

    lib = _lib.load()
    dev = best_key.device
    P, D = model.num_prototypes, model.prototype_shape[1]
    merge_keys(best_key, group)
    side = {}
    if dense:
        # one launch decodes the keys and derives ownership / clamped local indices; one launch collects the rows
        idx = torch.empty(P, dtype=torch.int64, device=dev)
        dmin = torch.empty(P, dtype=torch.float32, device=dev)
        local = torch.empty(P, dtype=torch.int64, device=dev)
        own = torch.empty(P, dtype=torch.int32, device=dev)
        valid = torch.empty(P, dtype=torch.int32, device=dev)
        vec = torch.empty((P, D), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.pasn_push_select(best_key.data_ptr(), P, lo, hi, idx.data_ptr(), dmin.data_ptr(),
                                            local.data_ptr(), own.data_ptr(), valid.data_ptr(), st), "pasn_push_select")
            if hi > lo:
                feats, dist_w, occ, logits = model._rt_push_forward_features(fetch_features(local + lo))
                _lib.check(lib.pasn_push_collect(feats.data_ptr(), own.data_ptr(), vec.data_ptr(), P, D, st),
                           "pasn_push_collect")
                side = {"owned": own, "occurrence_maps": occ, "logits": logits, "distance": dist_w}  # row p <-> prototype p
            else:
                vec.zero_()
    else:
        idx, dmin = decode_keys(best_key)
        mine = (idx >= lo) & (idx < hi)
        valid = (idx >= 0).to(torch.int32)
        vec = torch.zeros((P, D), dtype=torch.float32, device=dev)
        if bool(mine.any()):
            protos = torch.nonzero(mine).flatten()
            uniq, inv = torch.unique(idx[protos], return_inverse=True)
            x = fetch_features(uniq)
            feats, dist_w, occ, logits = model._rt_push_forward_features(x)
            vec[protos] = feats[inv, protos]
            side = {"prototypes": protos, "clips": uniq, "inv": inv, "occurrence_maps": occ[inv, protos],
                    "logits": logits[inv], "distance": dist_w[inv, protos]}
    rank, world = _world(group)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    if replace_prototypes:
        pv = model.prototype_vectors.data
        with torch.cuda.device(dev):
            _lib.check(lib.pasn_push_write_prototypes(pv.data_ptr(), vec.data_ptr(), valid.data_ptr(), P, D,
                                                      torch.cuda.current_stream(dev).cuda_stream),
                       "pasn_push_write_prototypes")
    return {"index": idx, "distance": dmin, "features": vec, "side": side}


def push_resident(model, features: torch.Tensor, labels: torch.Tensor, global_offset: int = 0, n_total: Optional[int] = None,
                  chunk: int = 8192, class_specific=True, abstain_class=True, replace_prototypes=True, group=None):
    """Push over backbone feature maps already resident on this rank's GPU (``features`` [n_local,C,*spatial],
    ``labels`` int64 [n_local]); ``global_offset`` = global index of local clip 0.  This is the path bench.py times."""
    dev = features.device
    P = model.num_prototypes
    pc = proto_class_restriction(model, class_specific, abstain_class).to(dev)
    key = new_best_key(P, dev)
    n_local = features.shape[0]
    for i in range(0, n_local, chunk):
        model.push_scan(features[i:i + chunk], labels[i:i + chunk], pc, global_offset + i, key, backbone=False)
    lo, hi = global_offset, global_offset + n_local

    def fetch(gidx):
        return features.index_select(0, gidx - lo)

    return finish_push(model, key, fetch, lo, hi, replace_prototypes, group, dense=True)


def push_prototypes(
    dataloader,
    model,
    class_specific=True,
    abstain_class=True,
    preprocess_input_function=None,
    root_dir_for_saving_prototypes=None,
    epoch_number=None,
    log=print,
    prototype_img_filename_prefix=None,
    prototype_self_act_filename_prefix=None,
    proto_bound_boxes_filename_prefix=None,
    replace_prototypes=True,
    group=None,
):
    """Same signature and effect as the reference's ``push_prototypes`` (push_abs_revision.py:181-196).

    ``dataloader`` yields dicts with ``"cine"`` [B,3,(T),H,W], ``"target_AS"`` [B] and ``"filename"``; it must be
    unshuffled (src/data/as_dataloader.py:65) and expose ``.dataset`` with ``__getitem__`` for the winner re-fetch.
    Under torch.distributed every rank calls this with the SAME loader; rank r processes the batches whose global
    clip range falls in its shard.
    """
    model.eval()
    log(f"############## push at epoch {epoch_number} #################")
    start = time.time()
    proto_epoch_dir = None
    if root_dir_for_saving_prototypes is not None:
        proto_epoch_dir = root_dir_for_saving_prototypes if epoch_number is None else \
            os.path.join(root_dir_for_saving_prototypes, "epoch-" + str(epoch_number))
        os.makedirs(proto_epoch_dir, exist_ok=True)

    dev = model.prototype_vectors.device
    P = model.num_prototypes
    pc = proto_class_restriction(model, class_specific, abstain_class).to(dev)
    key = new_best_key(P, dev)
    rank, world = _world(group)
    n_batches = len(dataloader)
    per = -(-n_batches // world)
    b_lo, b_hi = min(rank * per, n_batches), min((rank + 1) * per, n_batches)

    offset = 0
    lo = hi = None
    with torch.no_grad():
        for bi, data_sample in enumerate(dataloader):
            x = data_sample["cine"]
            B = x.shape[0]
            if b_lo <= bi < b_hi:
                if lo is None:
                    lo = offset
                if preprocess_input_function is not None:
                    x = preprocess_input_function(x)
                x = x.to(dev, non_blocking=True)
                y = data_sample["target_AS"].to(dev, non_blocking=True).to(torch.int64)
                model.push_scan(x, y, pc, offset, key)
                hi = offset + B
            offset += B
    if lo is None:
        lo = hi = 0

    dataset = getattr(dataloader, "dataset", None)
    fetched = {}

    def fetch(gidx):
        clips = []
        for g in gidx.tolist():
            sample = dataset[g]
            fetched[g] = sample
            xi = sample["cine"]
            if preprocess_input_function is not None:
                xi = preprocess_input_function(xi.unsqueeze(0)).squeeze(0)
            clips.append(xi)
        xb = torch.stack(clips).to(dev)
        with torch.no_grad():
            return model.cnn_backbone(xb)

    with torch.no_grad():
        res = finish_push(model, key, fetch, lo, hi, replace_prototypes, group)

    # side data in the reference's pickle schema (push_abs_revision.py:309-325); local winners only under DDP
    side = res["side"]
    if proto_epoch_dir is not None and side:
        protos = side["prototypes"].tolist()
        clips = side["clips"][side["inv"]].tolist()
        info = {
            "prototype_ids": np.asarray(protos),
            "prototypes_filenames": np.array([fetched[g].get("filename", str(g)) for g in clips]),
            "prototypes_src_imgs": np.array([np.asarray(fetched[g]["cine"]) for g in clips]),
            "prototypes_gts": np.array([int(fetched[g]["target_AS"]) for g in clips]),
            "prototypes_preds": side["logits"].float().cpu().numpy(),
            "prototypes_occurrence_maps": side["occurrence_maps"].float().cpu().numpy(),
            "prototypes_similarity_to_src_ROIs": 1 - res["distance"][side["prototypes"]].double().cpu().numpy(),
        }
        suffix = "" if world == 1 else f".rank{rank}"
        with open(os.path.join(proto_epoch_dir, f"prototypes_info{suffix}.pickle"), "wb") as f:
            pickle.dump(info, f)
    log("\tpush time: \t{0}".format(time.time() - start))
    return res


def agent_push(agent, replace_prototypes=True):
    """``XProtoNet_Base.push`` (src/agents/XProtoNet_Base.py:149-167) for an agent-like object exposing
    ``current_epoch``, ``data_loaders['train_push']``, ``model`` and ``config``."""
    epoch = f"{agent.current_epoch}_pushed"
    return push_prototypes(
        dataloader=agent.data_loaders["train_push"], model=agent.model, abstain_class=agent.config["abstain_class"],
        preprocess_input_function=None,
        root_dir_for_saving_prototypes=os.path.join(agent.config["save_dir"], "img"), epoch_number=epoch,
        prototype_img_filename_prefix="prototype-img", prototype_self_act_filename_prefix="prototype-self-act",
        proto_bound_boxes_filename_prefix="bb", replace_prototypes=replace_prototypes)
