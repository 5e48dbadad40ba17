#!/usr/bin/env python
"""SURVEY 8(e) on real GPUs: the merged counters of a W-rank Monte-Carlo run equal those of one rank simulating the same
global frame range.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port 29533 tools/shard_check.py

cfg5 of BASELINE.json (H_bit_GF256, T-EMS 2/2/0.5, early stop) at 4 dB with CRC-8 messages; rank 0 prints one JSON line.
With --abi-comm the counters are summed by the library's own communicator (nbl_comm_create_rank + nbl_allreduce_counters: one rank
per process, the NCCL id handed over with torch.distributed) instead of torch.distributed's all_reduce."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    from nbldpc_b200 import capi, montecarlo as mc
    from config_sweep import write_code_files
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = "/tmp/nbldpc_sweep_%d" % rank
    os.makedirs(tmp, exist_ok=True)
    paths, _ = write_code_files(tmp)
    code = capi.Code.from_hbit(paths["H_bit_GF256"], 256)
    batch, rounds, snr, seed = 148 * 256, 6, 4.0, 11
    stream = torch.cuda.Stream()                  # the decoder launches on torch's current stream: zeroing, kernels and the
    torch.cuda.set_stream(stream)                 # read-back of the counters are ordered
    dec = capi.Decoder(code, capi.METHOD_TEMS, 10, 2, 2, 1.0, 0.5, precision=capi.PREC_F32, early_stop=1, max_batch=batch, device=local,
                       stream=stream.cuda_stream)
    dev = torch.zeros(capi.NUM_COUNTERS, dtype=torch.int64, device="cuda")

    def simulate(lo, hi):
        dev.zero_()
        dec.simulate_async(snr, hi - lo, dev, seed=seed, first_frame=lo, random_msg=1, crc_len=8)
        return dev

    abi = "--abi-comm" in sys.argv
    comm = None
    if abi and world > 1:
        ident = torch.from_numpy(capi.Comm.unique_id()).cuda() if rank == 0 else torch.zeros(128, dtype=torch.uint8, device="cuda")
        dist.broadcast(ident, 0)
        comm = capi.Comm.from_rank(world, rank, ident.cpu().numpy(), local)

    def all_reduce(v):
        if world > 1 and comm is not None:        # nbl_allreduce_counters on the decoder's stream, in place on a device vector
            t = torch.from_numpy(v).cuda()
            torch.cuda.current_stream().synchronize()
            comm.allreduce_counters([dec], [t])
            dec.sync()
            v[:] = t.cpu().numpy()
        elif world > 1:
            t = torch.from_numpy(v).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            v[:] = t.cpu().numpy()

    class Fixed(mc.StopRule):                     # a fixed number of rounds, so that W = 1 and W > 1 cover the same frames
        def keep_going(self, c):
            return c[0] < rounds * world * batch

    total, n = mc.run_point(lambda lo, hi: simulate(lo, hi).cpu().numpy(), all_reduce, rank, world, batch, Fixed())
    if rank == 0:
        alone = np.zeros(capi.NUM_COUNTERS, np.int64)
        for k in range(rounds * world):          # the same global range, one rank, batch by batch
            alone += simulate(k * batch, (k + 1) * batch).cpu().numpy()
        again = np.zeros(capi.NUM_COUNTERS, np.int64)
        for k in range(rounds * world):
            again += simulate(k * batch, (k + 1) * batch).cpu().numpy()
        assert (again == alone).all(), (again, alone)           # and the run is deterministic
        print(json.dumps({"world": world, "collective": "nbl_allreduce_counters (C ABI, ncclCommInitRank)" if comm is not None else "torch.distributed all_reduce", "frames": int(total[0]), "merged": [int(x) for x in total], "one_rank": [int(x) for x in alone],
                          "identical": bool((total == alone).all())}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
