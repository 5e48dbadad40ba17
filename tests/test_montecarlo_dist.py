"""Multi-process (gloo, world_size 2) test of the sharded Monte-Carlo host logic: the N > 1 path of bench.py and
nbldpc_b200.montecarlo.  The per-frame outcome is a deterministic function of the global frame index (as it is on
the GPU, where noise is Philox-keyed by frame index), so the merged counters must not depend on the world size."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nbldpc_b200 import montecarlo as mc


def fake_simulate(lo, hi):
    """stand-in for nbl_simulate_async: frame f is in error iff hash(f) % 37 == 0; iterations 1 + f % 9"""
    f = np.arange(lo, hi, dtype=np.int64)
    bad = ((f * 2654435761) % 4294967296) % 37 == 0
    c = np.zeros(8, np.int64)
    c[0] = len(f); c[1] = bad.sum(); c[2] = 3 * bad.sum(); c[3] = 11 * bad.sum()
    c[4] = (~bad).sum() + (bad & (f % 5 == 0)).sum(); c[5] = (bad & (f % 5 == 0)).sum(); c[6] = (1 + f % 9).sum()
    return c


def _worker(rank, world, port, batch, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def all_reduce(vec):
        t = torch.from_numpy(vec)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    total, rounds = mc.run_point(fake_simulate, all_reduce, rank, world, batch, mc.StopRule(min_frames=2000, min_err_frames=60))
    gathered = [None] * world
    dist.all_gather_object(gathered, (total.tolist(), rounds))
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


def test_two_ranks_match_one_rank_over_the_same_frames():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29531, 64, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (t0, r0), (t1, r1) = res
    assert t0 == t1 and r0 == r1                       # every rank holds the same global counters and stops together
    frames = t0[0]
    assert frames == r0 * 2 * 64
    single = fake_simulate(0, frames)                  # 1-rank run over the same global range
    assert single.tolist() == t0
    # and the 1-rank driver stops no later than the same number of frames rounded up to its own round size
    tot1, rounds1 = mc.run_point(fake_simulate, lambda v: None, 0, 1, 128, mc.StopRule(min_frames=2000, min_err_frames=60))
    assert tot1.tolist() == t0 and rounds1 == r0


def test_shard_ranges_tile_the_frame_axis():
    for world in (1, 2, 4, 8):
        seen = []
        for k in range(3):
            for r in range(world):
                seen.append(mc.shard_range(k, r, world, 10, first_frame=5))
        flat = sorted(seen)
        assert flat[0][0] == 5
        assert all(a[1] == b[0] for a, b in zip(flat, flat[1:]))


def test_stop_rule_matches_reference_semantics():
    rule = mc.StopRule(min_frames=300, min_err_frames=30, min_undetected=-1)
    c = np.zeros(8, np.int64)
    c[0], c[1] = 1000, 30
    assert rule.keep_going(c)                          # errFrame <= minErrFrame keeps going (Simulation.cpp:375)
    c[1] = 31
    assert not rule.keep_going(c)
    c[0] = 300
    assert rule.keep_going(c)                          # frames <= minSimCycle keeps going
    assert mc.StopRule(300, 30, 0).keep_going(np.array([1000, 31, 0, 0, 0, 0, 0, 0]))   # U criterion active when >= 0
