"""Multi-GPU through the C ABI: nbl_comm_* / nbl_allreduce_counters / nbl_simulate_sharded and the
`--gpus` mode of the C++ driver.  The sharding invariant of BASELINE configs[4]: the merged counters of a G-GPU run equal those
of one GPU simulating the same global frame range, because frame f has the same message, code word and noise wherever it runs.
Tests that need two devices skip on a one-GPU box."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import bundle, load_product_code

pytestmark = pytest.mark.gpu
TEMS = dict(method=4, max_iter=10, p0=2, p1=2, factor=1.0, offset=0.5)
SIM = dict(ebn0_db=4.0, seed=11, random_msg=1, crc_len=8)


def _decoders(code, devices, batch):
    from nbldpc_b200 import capi
    return [capi.Decoder(code, precision=capi.PREC_F32, early_stop=1, max_batch=batch, device=d, **TEMS) for d in devices]


def test_sharded_round_on_one_rank_equals_plain_simulate(code_files):
    """world = 1: no NCCL involved; a round is nbl_simulate of the round's frame range"""
    from nbldpc_b200 import capi
    code = load_product_code(code_files, "H_bit_GF256")
    batch, rounds = 148 * 64, 3
    comm = capi.Comm(1)
    assert (comm.size, comm.local_ranks, comm.rank(0), comm.device(0)) == (1, 1, 0, 0)
    dec = _decoders(code, [0], batch)
    tot = np.zeros(capi.NUM_COUNTERS, np.uint64)
    for r in range(rounds):
        comm.simulate_sharded(dec, batch, r, tot, **SIM)
    ref = dec[0].simulate(frames=rounds * batch, first_frame=0, **SIM)
    assert np.array_equal(tot, ref), (tot, ref)
    assert tot[0] == rounds * batch and tot[1] > 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_cfg5_sharded_counters_do_not_depend_on_the_number_of_gpus(world, code_files):
    """BASELINE configs[4] (H_bit_GF256, T-EMS 2/2/0.5, 4 dB, CRC-8 messages): G GPUs of one process, frames split by global
    index, one NCCL all-reduce of the 8 x u64 per round; merged totals == one GPU over the same range."""
    from nbldpc_b200 import capi
    if capi.device_count() < world:
        pytest.skip("needs %d CUDA devices" % world)
    code = load_product_code(code_files, "H_bit_GF256")
    per_round, rounds = 148 * 256, 4                               # frames per round over all ranks
    comm = capi.Comm(world)
    assert comm.size == world and comm.local_ranks == world
    dec = _decoders(code, [comm.device(i) for i in range(world)], per_round // world)
    tot = np.zeros(capi.NUM_COUNTERS, np.uint64)
    for r in range(rounds):
        comm.simulate_sharded(dec, per_round // world, r, tot, **SIM)
    one = _decoders(code, [0], per_round)[0]
    ref = one.simulate(frames=rounds * per_round, first_frame=0, **SIM)
    assert np.array_equal(tot, ref), (tot, ref)
    assert tot[0] == rounds * per_round and tot[1] > 0


def test_allreduce_counters_sums_device_vectors(code_files):
    """nbl_allreduce_counters alone: in-place sum on the decoders' streams (2 GPUs), identity on one"""
    import torch
    from nbldpc_b200 import capi
    code = load_product_code(code_files, "H_bit_GF256")
    world = 2 if capi.device_count() >= 2 else 1
    comm = capi.Comm(world)
    dec = _decoders(code, list(range(world)), 64)
    bufs = []
    for i in range(world):
        with torch.cuda.device(i):
            bufs.append(torch.arange(8, dtype=torch.int64, device="cuda:%d" % i) * (i + 1) + 100 * i)
            torch.cuda.synchronize()
    want = sum(b.cpu() for b in bufs)
    comm.allreduce_counters(dec, bufs)
    for d in dec:
        d.sync()
    for b in bufs:
        assert torch.equal(b.cpu(), want)


def _profile(tmp_path, code_files):
    """cfg5 profile for the C++ driver: H_bit_GF256 (converted to the reference's code format), T-EMS 2/2/0.5, one point at 4 dB"""
    from oracle import pyoracle as po
    from oracle.refprofile import profile_text
    from conftest import get_code_dict
    (tmp_path / "hbit.txt").write_text(po.code_text(get_code_dict("H_bit_GF256")))
    cons = bundle()["constellations"]["BPSK"]
    (tmp_path / "BPSK.txt").write_text("".join("Point: %d Real: %r Imag: %r\n" % (i, c[0], c[1]) for i, c in enumerate(cons)))
    prof = profile_text(dict(GFq=256, code_file="hbit.txt", method=4, snr=4.0, random_msg=1, parallel=1, seed=11,
                             min_cycle=3 * 148 * 128, min_err=-1, min_uerr=-1))
    (tmp_path / "prof.txt").write_text(prof)
    return "prof.txt"


def test_cpp_driver_gpus_flag_gives_the_same_counters_for_every_g(tmp_path, code_files):
    """nbldpc_sim --gpus G (C++ above the C ABI: sharding, NCCL merge and the reference's stop rule outside Python):
    identical counter line for G = 1 and G = 2 when G * batch is kept constant"""
    from nbldpc_b200 import build, capi
    exe = [p for p in build.build_host() if p.endswith("nbldpc_sim")][0]
    prof = _profile(tmp_path, code_files)
    lines = {}
    for g in (1, 2):
        if capi.device_count() < g:
            continue
        out = subprocess.run([exe, prof, "--fp32", "--gpus", str(g), "--batch", str(148 * 128 // g)], cwd=tmp_path, capture_output=True,
                             text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        m = re.search(r"frames (\d+)\s+mean iterations \S+\s+counters((?: \d+)+)", out.stdout)
        assert m, out.stdout
        lines[g] = [int(x) for x in m.group(2).split()]
        assert lines[g][0] == int(m.group(1)) and lines[g][0] >= 3 * 148 * 128
    assert lines[1][1] > 0
    if 2 in lines:
        assert lines[1] == lines[2], lines
