"""Sharded Monte-Carlo driver: one process per GPU, frames split by global frame index, counters all-reduced.

This is the multi-GPU form of the reference's frame loop (main.cpp:43-55): the reference runs `parallel` CPU workers
per round and adds their six error counters serially (Comm.cpp:484-493); here every rank simulates `batch` frames per
round on its GPU (nbl_simulate_async), the 8 x u64 counter vector is summed across ranks with ONE all-reduce per
round (NCCL on GPUs, gloo in the CPU tests), and every rank evaluates the reference's stop rule
(Simulation.cpp:373-376) on the global counters -- so all ranks stop in the same round without further traffic.

Frame f of a run always has the same codeword and the same noise (Philox keyed by (seed, f)), and round k covers the
contiguous global range [k*W*batch, (k+1)*W*batch) split in rank order, so the summed counters after K rounds
with W ranks equal those of a 1-rank run over the same range.
"""
from dataclasses import dataclass

import numpy as np

from .capi import COUNTER_NAMES, NUM_COUNTERS


@dataclass
class StopRule:
    """the reference's rule: keep going while ANY criterion is unmet; a negative limit disables a criterion"""
    min_frames: int = 300
    min_err_frames: int = 30
    min_undetected: int = -1
    max_frames: int = 0          # safety cap (0 = none); not in the reference

    def keep_going(self, c):
        if self.max_frames and c[0] >= self.max_frames:
            return False
        return c[0] <= self.min_frames or c[1] <= self.min_err_frames or c[5] <= self.min_undetected


def shard_range(round_index, rank, world, batch, first_frame=0):
    """global frame range [lo, hi) that `rank` simulates in round `round_index`"""
    lo = first_frame + (round_index * world + rank) * batch
    return lo, lo + batch


def run_point(simulate, all_reduce, rank, world, batch, rule, first_frame=0):
    """One Eb/N0 point.  simulate(lo, hi) -> uint64[8] counters of frames [lo, hi) (device or host array-like);
    all_reduce(vec) sums an int64 numpy vector over ranks in place.  Returns (global counters, rounds)."""
    total = np.zeros(NUM_COUNTERS, np.int64)
    rounds = 0
    while True:
        lo, hi = shard_range(rounds, rank, world, batch, first_frame)
        local = np.asarray(simulate(lo, hi), dtype=np.int64).copy()
        all_reduce(local)
        total += local
        rounds += 1
        if not rule.keep_going(total):
            return total, rounds


def rates(counters, info_syms, bits_per_sym):
    c = dict(zip(COUNTER_NAMES, (int(v) for v in counters)))
    f = max(1, c["frames"])
    return {"frames": c["frames"], "FER": c["err_frames"] / f, "SER": c["err_syms"] / (f * info_syms),
            "BER": c["err_bits"] / (f * info_syms * bits_per_sym), "undetected_FER": c["undetected"] / f,
            "mean_iterations": c["iterations"] / f}
