"""Push / prototype projection on device, sharded across ranks.

Mirrors ``push_prototypes`` of the reference (src/utils/push_abs_revision.py:181-348) and the agent wrapper
``XProtoNet_Base.push`` (src/agents/XProtoNet_Base.py:149-167), re-designed for B200:

  scan    every rank walks its contiguous shard of the unshuffled ``train_push`` set; per batch ONE fused launch pair
          computes the head, folds the class-restricted running argmin into packed keys
          ``orderable(fp32 dist) << 32 | global index`` AND keeps the winner's ``features_extracted`` row -- both in one
          device record ``[keys P x u64 | vectors P x D x f32]``.  The winner is captured in the pass that saw it, as the
          reference does (push_abs_revision.py:299-307): its push loader draws a random window per ``__getitem__``
          (src/data/as_dataloader.py:246-255), so a clip cannot be fetched a second time.  No per-batch D2H of features,
          distances, occurrence maps or input clips (push_abs_revision.py:278-285).
  merge   ONE collective: all-gather of the records (41 KB per rank at P=40, D=256); every rank then picks, per
          prototype, the record with the smallest key (lowest distance, ties -> lowest global index) and overwrites
          ``prototype_vectors`` identically (push_abs_revision.py:342-346).
  side    with a Dataset loader the data the reference pickles (occurrence map, logits, label, filename, input clip;
          push_abs_revision.py:301-325) is taken from the batch that improved a prototype, while it is still in memory.

Defined behaviour where the reference crashes: a class-specific prototype whose class never occurs keeps its old
vector and reports index -1 (SURVEY.md section 7 'Empty class').  Rendering of prototype images / mp4s
(push_abs_revision.py:13-178, :329-340) is out of scope; the pickle with the same keys is written when a save
directory is given.
"""
from __future__ import annotations

import os
import pickle
import time
from typing import Dict, Optional, Sequence, Tuple

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
    _init_key(key)
    return key


def _init_key(key: torch.Tensor) -> None:
    lib = _lib.load()
    dev = key.device
    with torch.cuda.device(dev):
        _lib.check(lib.pasn_push_init(key.data_ptr(), key.numel(), torch.cuda.current_stream(dev).cuda_stream), "pasn_push_init")


class PushRecord:
    """One rank's running push state in ONE device buffer ``[keys: P x u64 | vectors: P x D x f32]`` (include/pasn.h,
    ``pasn_push_record_bytes``): ``key`` and ``vec`` are views, ``buf`` is what the all-gather moves."""

    def __init__(self, P: int, D: int, device, buf: Optional[torch.Tensor] = None):
        self.P, self.D = int(P), int(D)
        nbytes = self.P * 8 + self.P * self.D * 4
        # ``buf``: caller-provided storage (a slice of peer-mapped symmetric memory, see PeerRecords)
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device) if buf is None else buf[:nbytes]
        self.peer = None   # (PeerRecords, buffer index, epoch) when the merge runs over peer memory
        self.key = self.buf[: self.P * 8].view(torch.int64)
        self.vec = self.buf[self.P * 8:].view(torch.float32).view(self.P, self.D)
        if self.buf.is_cuda:
            _init_key(self.key)
        else:   # host-side bookkeeping only (gloo tests of the merge logic)
            self.key.fill_(_KEY_NONE)


_KEY_NONE = (1 << 63) - 1   # INT64_MAX: no candidate


class PeerRecords:
    """Push records of all ranks of one node in peer-mapped (symmetric) memory, so that the merge is ONE kernel that reads the
    peers' records over NVLink (``pasn_push_merge_peers``) instead of a collective call.  Per rank: two record buffers
    (alternating from push to push: a rank that starts its next push never overwrites what a slower peer is still reading)
    and one epoch flag.  Built once per (model, device, P, D, group): the rendezvous is collective and not cheap."""

    def __init__(self, P: int, D: int, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.P, self.D = int(P), int(D)
        rec = self.P * 8 + self.P * self.D * 4
        self.stride = (rec + 255) // 256 * 256
        self.mem = symm.empty(2 * self.stride + 256, dtype=torch.uint8, device=device)
        self.mem.zero_()
        pg = group if group is not None else dist.group.WORLD
        self.handle = symm.rendezvous(self.mem, pg)
        self.rank, self.world = int(self.handle.rank), int(self.handle.world_size)
        ptrs = [int(q) for q in self.handle.buffer_ptrs]
        self.rec_ptrs = [torch.tensor([q + b * self.stride for q in ptrs], dtype=torch.int64, device=device) for b in (0, 1)]
        self.flag_ptrs = torch.tensor([q + 2 * self.stride for q in ptrs], dtype=torch.int64, device=device)
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=group)      # every rank's flags are zero before anybody can poll them

    def next_record(self) -> "PushRecord":
        self.epoch += 1
        b = self.epoch & 1
        rec = PushRecord(self.P, self.D, self.mem.device, buf=self.mem[b * self.stride: (b + 1) * self.stride])   # keys reset
        rec.peer = (self, b, self.epoch)
        return rec


def _peer_records(model, P: int, D: int, device, group=None) -> Optional[PeerRecords]:
    """PeerRecords for this model (cached), or None when the merge has to go through the all-gather: single rank, CPU
    bookkeeping, ``PASN_PUSH_PEER=0``, or symmetric memory not available for this group (decided once, by all ranks alike)."""
    import torch.distributed as dist

    _, world = _world(group)
    if world == 1 or torch.device(device).type != "cuda" or os.environ.get("PASN_PUSH_PEER", "1") == "0":
        return None
    cache = model.__dict__.setdefault("_pasn_peer_records", {})
    ck = (P, D, str(device), id(group))
    if ck not in cache:
        ok = torch.ones(1, dtype=torch.int32, device=device)
        pr = None
        try:
            pr = PeerRecords(P, D, device, group)
        except Exception as e:   # noqa: BLE001 -- any failure means "use the collective"; all ranks must agree
            ok.zero_()
            if os.environ.get("PASN_PUSH_PEER_VERBOSE"):
                print(f"[protoasnet_b200] peer-memory push merge unavailable: {e!r}")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        cache[ck] = pr if int(ok.item()) == 1 else None
    return cache[ck]


def merge_keys(best_key: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce(MIN) of packed keys alone (kept for callers that only need indices).  Keys are stored with the top bit
    flipped (include/pasn.h), so the signed int64 order NCCL / gloo reduce with is the key order."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(best_key, op=dist.ReduceOp.MIN, group=group)
    return best_key


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


def gather_records(rec: PushRecord, group=None) -> Tuple[torch.Tensor, int]:
    """The one collective of a push: all-gather of the ranks' records -> (R records back to back, R)."""
    import torch.distributed as dist

    _, world = _world(group)
    if world == 1:
        return rec.buf, 1
    out = torch.empty(world * rec.buf.numel(), dtype=torch.uint8, device=rec.buf.device)
    if rec.buf.is_cuda:
        dist.all_gather_into_tensor(out, rec.buf, group=group)
    else:
        dist.all_gather(list(out.view(world, -1).unbind(0)), rec.buf, group=group)
    return out, world


def reduce_records(gathered: torch.Tensor, R: int, P: int, D: int):
    """Per prototype the record with the smallest key wins -> (index, distance, valid int32, vec [P,D])."""
    dev = gathered.device
    if not gathered.is_cuda:   # host-side bookkeeping only (gloo tests of the merge logic)
        recs = gathered.view(R, -1)
        keys = torch.stack([recs[r, : P * 8].view(torch.int64) for r in range(R)])
        vecs = torch.stack([recs[r, P * 8:].view(torch.float32).view(P, D) for r in range(R)])
        best, who = keys.min(dim=0)
        idx, d = decode_keys(best)
        valid = (idx >= 0).to(torch.int32)
        vec = vecs[who, torch.arange(P)] * valid[:, None]
        return idx, d, valid, vec
    lib = _lib.load()
    idx = torch.empty(P, dtype=torch.int64, device=dev)
    d = torch.empty(P, dtype=torch.float32, device=dev)
    valid = torch.empty(P, dtype=torch.int32, device=dev)
    vec = torch.zeros((P, D), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.pasn_push_reduce(gathered.data_ptr(), R, P, D, idx.data_ptr(), d.data_ptr(), valid.data_ptr(),
                                        vec.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "pasn_push_reduce")
    return idx, d, valid, vec


def finish_push(model, rec: PushRecord, replace_prototypes=True, group=None):
    """One all-gather of the records, per-prototype minimum, prototype overwrite.  No host synchronisation.
    Returns dict(index, distance, features (P,D), valid)."""
    lib = _lib.load()
    if rec.peer is not None:
        # exchange + merge in one kernel over NVLink peer memory
        pr, b, epoch = rec.peer
        dev = rec.buf.device
        idx = torch.empty(rec.P, dtype=torch.int64, device=dev)
        dmin = torch.empty(rec.P, dtype=torch.float32, device=dev)
        valid = torch.empty(rec.P, dtype=torch.int32, device=dev)
        vec = torch.empty((rec.P, rec.D), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.pasn_push_merge_peers(pr.rec_ptrs[b].data_ptr(), pr.flag_ptrs.data_ptr(), pr.world, pr.rank,
                                                 epoch & 0xFFFFFFFF, rec.P, rec.D, idx.data_ptr(), dmin.data_ptr(),
                                                 valid.data_ptr(), vec.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                       "pasn_push_merge_peers")
    else:
        gathered, R = gather_records(rec, group)
        idx, dmin, valid, vec = reduce_records(gathered, R, rec.P, rec.D)
    if replace_prototypes:
        pv = model.prototype_vectors.data
        dev = pv.device
        with torch.cuda.device(dev):
            _lib.check(lib.pasn_push_write_prototypes(pv.data_ptr(), vec.data_ptr(), valid.data_ptr(), rec.P, rec.D,
                                                      torch.cuda.current_stream(dev).cuda_stream),
                       "pasn_push_write_prototypes")
    return {"index": idx, "distance": dmin, "features": vec, "valid": valid, "side": {}}


def push_resident(model, features: torch.Tensor, labels: torch.Tensor, global_offset: int = 0, n_total: Optional[int] = None,
                  chunk: int = 8192, class_specific=True, abstain_class=True, replace_prototypes=True, group=None):
    """Push over backbone feature maps already resident on this rank's GPU (``features`` [n_local,C,*spatial],
    ``labels`` int64 [n_local]); ``global_offset`` = global index of local clip 0.  This is the path bench.py times."""
    dev = features.device
    P, D = model.num_prototypes, model.prototype_shape[1]
    # the class restriction only depends on the (fixed) prototype_class_identity buffer: built once per model / device, so a
    # push issues no host->device copy of its own (at 8 GPUs the fixed cost of a push is what limits its scaling)
    ck = (bool(class_specific), bool(abstain_class), str(dev))
    cache = model.__dict__.setdefault("_pasn_push_pc", {})
    pc = cache.get(ck)
    if pc is None:
        pc = cache[ck] = proto_class_restriction(model, class_specific, abstain_class).to(dev)
    peers = _peer_records(model, P, D, dev, group)
    rec = peers.next_record() if peers is not None else PushRecord(P, D, dev)
    n_local = features.shape[0]
    for i in range(0, n_local, chunk):
        model.push_scan(features[i:i + chunk], labels[i:i + chunk], pc, global_offset + i, rec.key, backbone=False,
                        best_vec=rec.vec)
    return finish_push(model, rec, replace_prototypes, group)


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

    ``dataloader`` yields dicts with ``"cine"`` [B,3,(T),H,W], ``"target_AS"`` [B] and ``"filename"``, unshuffled
    (src/data/as_dataloader.py:65).  Every clip is seen exactly once: the winners' vectors and side data are captured
    from the batch that produced them, so a loader whose ``__getitem__`` is random (as the reference's is) is fine.
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
    P, D = model.num_prototypes, model.prototype_shape[1]
    pc = proto_class_restriction(model, class_specific, abstain_class).to(dev)
    peers = _peer_records(model, P, D, dev, group)
    rec = peers.next_record() if peers is not None else PushRecord(P, D, dev)
    rank, world = _world(group)
    n_batches = len(dataloader)
    per = -(-n_batches // world)
    b_lo, b_hi = min(rank * per, n_batches), min((rank + 1) * per, n_batches)

    # side data of the current winner of each prototype (push_abs_revision.py:301-307), taken from the batch that
    # improved it while that batch is still in memory
    side = {j: None for j in range(P)}
    want_side = proto_epoch_dir is not None
    protos = []
    offset = 0
    with torch.no_grad():
        for bi, data_sample in enumerate(dataloader):
            x = data_sample["cine"]
            B = x.shape[0]
            if b_lo <= bi < b_hi:
                if preprocess_input_function is not None:
                    x = preprocess_input_function(x)
                xd = x.to(dev, non_blocking=True)
                y = data_sample["target_AS"].to(dev, non_blocking=True).to(torch.int64)
                fmap = model.cnn_backbone(xd)
                prev = rec.key.clone() if want_side else None
                model.push_scan(fmap, y, pc, offset, rec.key, backbone=False, best_vec=rec.vec)
                if want_side:
                    changed = torch.nonzero(rec.key != prev).flatten()
                    if changed.numel():                       # O(P log N) batches in total
                        idx_now, _ = decode_keys(rec.key)
                        local = (idx_now[changed] - offset)
                        uniq, inv = torch.unique(local, return_inverse=True)
                        _f, _d, occ, logits = model._rt_push_forward_features(fmap.index_select(0, uniq))
                        occ_rows = occ[inv, changed].float().cpu().numpy()
                        lg = logits[inv].float().cpu().numpy()
                        names = data_sample.get("filename", None)
                        for t, (j, a) in enumerate(zip(changed.tolist(), local.tolist())):
                            side[j] = {"occurrence_map": occ_rows[t], "logits": lg[t], "image": np.asarray(x[a]),
                                       "gt": int(data_sample["target_AS"][a]),
                                       "filename": names[a] if names is not None else str(offset + a)}
            offset += B

    with torch.no_grad():
        res = finish_push(model, rec, replace_prototypes, group)

    # side data in the reference's pickle schema (push_abs_revision.py:309-325); under DDP each rank writes the winners
    # that its own shard produced
    if want_side:
        idx_local, _ = decode_keys(rec.key)
        mine = (idx_local == res["index"]) & (res["index"] >= 0)
        protos = [j for j in torch.nonzero(mine).flatten().tolist() if side[j] is not None]
        if protos:
            info = {
                "prototype_ids": np.asarray(protos),
                "prototypes_filenames": np.array([side[j]["filename"] for j in protos]),
                "prototypes_src_imgs": np.array([side[j]["image"] for j in protos]),
                "prototypes_gts": np.array([side[j]["gt"] for j in protos]),
                "prototypes_preds": np.array([side[j]["logits"] for j in protos]),
                "prototypes_occurrence_maps": np.array([side[j]["occurrence_map"] for j in protos]),
                "prototypes_similarity_to_src_ROIs": 1 - res["distance"][protos].double().cpu().numpy(),
            }
            suffix = "" if world == 1 else f".rank{rank}"
            with open(os.path.join(proto_epoch_dir, f"prototypes_info{suffix}.pickle"), "wb") as f:
                pickle.dump(info, f)
        res["side"] = {j: side[j] for j in protos}
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
