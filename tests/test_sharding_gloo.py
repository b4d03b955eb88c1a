"""The N > 1 path on CPU: world_size-2 gloo processes run the range-sharded MSM host logic
(zkmember_b200/dist.py) with the oracle standing in for the per-GPU kernel, and must reproduce
the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zkmember_b200.dist import shard_range, record_words


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1000, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q, curve_name="bls12_381"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import capi
    from oracle.py import exact
    from oracle.py.params import CURVES
    from zkmember_b200.dist import sharded_msm
    curve = CURVES[curve_name]
    cid = curve.curve_id
    bases = capi.progression(cid, 1, 17, 3, n)
    scal = capi.random_scalars(cid, n, seed=99)
    W = curve.fq.limbs64

    def local_msm(lo, hi):      # stands in for RegisteredBases.msm_device on this rank's GPU
        xy, inf = capi.msm(cid, 1, bases[lo:hi], scal[lo:hi])
        rec = np.concatenate([xy, np.array([1 if inf else 0], dtype=np.uint64)])
        return torch.from_numpy(rec.view(np.int64).copy())

    def sum_records(allp):      # stands in for zkm_points_sum_device
        G = exact.Group(curve, 1)
        acc = None
        for row in allp.numpy().view(np.uint64):
            acc = G.add(acc, exact.point_from_bytes(curve, 1, row[:2 * W].tobytes(), int(row[2 * W])))
        b, f = exact.point_to_bytes(curve, 1, acc)
        return b, f

    got = sharded_msm(n, rank, world, local_msm, sum_records)
    want_xy, want_inf = capi.msm(cid, 1, bases, scal)
    q.put((rank, got[0] == want_xy.tobytes() and bool(got[1]) == want_inf))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,curve_name", [(1, "bls12_381"), (1001, "bls12_381"), (301, "bw6_761")])
def test_sharded_msm_world2_gloo(n, curve_name):
    assert record_words(6) == 13 and record_words(12) == 25      # BLS12-381 / BW6-761 G1 result records
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q, curve_name)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
