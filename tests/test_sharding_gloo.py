"""CPU tier, world_size 2 over gloo: the only multi-GPU exchange of the path is the scalar loss reduction
(vv_recon.py:383-385,399 reduce over the batch).  The per-cloud operator results here come from the CPU oracle (test
infrastructure); what is under test is the host logic in rfnet_b200.losses: shard bounds, partial sums, all-reduce, and
that the sharded loss equals the single-process loss."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, n, m, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import port as oracle
    from rfnet_b200 import losses
    rng = np.random.default_rng(99)  # same clouds on every rank; each rank evaluates only its slice
    x1 = (rng.random((B, n, 3), dtype=np.float32) - 0.5)
    x2 = (rng.random((B, m, 3), dtype=np.float32) - 0.5)
    lo, hi = losses.shard_bounds(B, rank, world)
    d1, _, d2, _ = oracle.nn_distance(x1[lo:hi], x2[lo:hi])
    sums = losses.chamfer_partial_sums(torch.from_numpy(d1), torch.from_numpy(d2))
    losses.all_reduce_scalars(sums)
    cd = float((sums[0] / sums[1] + sums[2] / sums[3]) / 2)
    match = oracle.approx_match(x1[lo:hi, :64], x2[lo:hi, :64])
    cost = torch.from_numpy(oracle.match_cost(x1[lo:hi, :64], x2[lo:hi, :64], match))
    es = torch.stack([(cost / 64.0).sum(), torch.tensor(float(cost.numel()))])
    losses.all_reduce_scalars(es)
    q.put((rank, cd, float(es[0] / es[1]), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_losses_equal_global_losses_world2():
    from oracle import port as oracle
    B, n, m, world = 5, 120, 90, 2   # odd batch: uneven shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, n, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(99)
    x1 = (rng.random((B, n, 3), dtype=np.float32) - 0.5)
    x2 = (rng.random((B, m, 3), dtype=np.float32) - 0.5)
    d1, _, d2, _ = oracle.nn_distance(x1, x2)
    want_cd = (np.sqrt(d1).mean() + np.sqrt(d2).mean()) / 2          # chamfer_big, vv_recon.py:381-385
    match = oracle.approx_match(x1[:, :64], x2[:, :64])
    want_emd = (oracle.match_cost(x1[:, :64], x2[:, :64], match) / 64.0).mean()   # earth_mover, vv_recon.py:392-399
    spans = sorted((lo, hi) for _, _, _, lo, hi in results)
    assert spans == [(0, 3), (3, 5)]
    for _, cd, emd, _, _ in results:
        assert abs(cd - want_cd) <= 1e-6 * abs(want_cd)
        assert abs(emd - want_emd) <= 1e-6 * abs(want_emd)
