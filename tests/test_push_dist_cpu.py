"""world_size-2 gloo test (CPU) of the host-side push merge: ONE all-gather of the ranks' push records
[keys | winner vectors], per-prototype minimum over the signed-order keys (ties -> lowest global index), and the
keys-only all-reduce(MIN) helper."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from protoasnet_b200 import push as pushmod
from protoasnet_b200 import synth


def _f32_orderable(d):
    u = np.asarray(d, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return np.where(u & 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000).astype(np.uint64)


def _pack(d, idx):
    """global key format of include/pasn.h: (orderable(dist) << 32 | index) with the top bit flipped (signed order)"""
    return (((_f32_orderable(d) << np.uint64(32)) | np.asarray(idx, dtype=np.uint64)) ^ np.uint64(1 << 63)).astype(np.uint64)


_NONE = np.uint64((1 << 63) - 1)


def _worker(rank, world, port, n_total, P, D, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.random.Generator(np.random.PCG64(99))
        dmat = g.random((n_total, P), dtype=np.float32) - np.float32(0.1)      # includes negative distances
        dmat[3, 1] = dmat[40, 1] = dmat[:, 1].min() - np.float32(0.5)          # exact tie across ranks
        feats = g.standard_normal((n_total, P, D), dtype=np.float32)
        lo, hi = synth.shard_range(n_total, rank, world)
        key = np.full(P, _NONE).view(np.int64)
        for n in range(lo, hi):
            key = np.minimum(key, _pack(dmat[n], np.full(P, n)).view(np.int64))   # signed order == key order
        key = key.view(np.uint64).copy()
        key[2] = _NONE                                                         # prototype 2: no candidate anywhere
        kt = torch.from_numpy(key.view(np.int64).copy())
        # this rank's record: running keys + the pooled-feature row of each local winner (captured during the scan)
        rec = pushmod.PushRecord(P, D, "cpu")
        rec.key.copy_(kt)
        idx_l, _ = pushmod.decode_keys(kt)
        for p in range(P):
            if int(idx_l[p]) >= 0:
                rec.vec[p] = torch.from_numpy(feats[int(idx_l[p]), p])
        gathered, R = pushmod.gather_records(rec)                              # the one collective
        assert R == world and gathered.numel() == world * rec.buf.numel()
        idx, dmin, valid, vec = pushmod.reduce_records(gathered, R, P, D)
        k2 = kt.clone()
        pushmod.merge_keys(k2)                                                 # keys-only helper agrees
        idx2, dmin2 = pushmod.decode_keys(k2)
        assert torch.equal(idx, idx2) and torch.equal(dmin, dmin2)
        assert valid.tolist() == [int(i >= 0) for i in idx.tolist()]
        if rank == 0:
            ret["idx"], ret["d"], ret["vec"] = idx.numpy(), dmin.numpy(), vec.numpy()
            exp_idx = dmat.argmin(axis=0)
            exp_idx[2] = -1
            ret["exp_idx"], ret["exp_d"] = exp_idx, dmat.min(axis=0)
            ret["exp_vec"] = np.stack([feats[exp_idx[p], p] if exp_idx[p] >= 0 else np.zeros(D, np.float32) for p in range(P)])
    finally:
        dist.destroy_process_group()


def test_two_rank_merge_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29517, 64, 6, 5, ret), nprocs=2, join=True)
    assert np.array_equal(ret["idx"], ret["exp_idx"])
    assert ret["idx"][1] == 3                                   # cross-rank tie -> lowest global index
    ok = ret["exp_idx"] >= 0
    assert np.array_equal(ret["d"][ok], ret["exp_d"][ok]) and np.isinf(ret["d"][2])
    assert np.array_equal(ret["vec"], ret["exp_vec"])           # the winner's row travels with its key


def test_shard_range_covers_set():
    for n in (0, 1, 7, 50000):
        for w in (1, 2, 3, 8):
            r = [synth.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def test_decode_keys_host_matches_packing():
    d = np.array([0.25, -1e-7, 0.0, 3.5], dtype=np.float32)
    k = _pack(d, [5, 6, 7, 8])
    idx, dd = pushmod.decode_keys(torch.from_numpy(k.view(np.int64).copy()))
    assert idx.tolist() == [5, 6, 7, 8] and np.array_equal(dd.numpy(), d)
    order = np.argsort(k.view(np.int64))
    assert order.tolist() == [1, 2, 0, 3]                       # signed key order == fp32 order
