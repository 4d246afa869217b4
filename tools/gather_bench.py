#!/usr/bin/env python
"""Time sharding.gather_waveforms alone (the one NCCL exchange of the path) on N ranks: config-3-sized blocks
(256 utterances x 220 672 samples in total, fp32), 10 repetitions after 2 warm-ups, CUDA events, max over ranks.

    torchrun --nproc-per-node N tools/gather_bench.py            # NCCL_* environment variables are passed through
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mb_istft_vits_b200 import sharding  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    total, S = 256, 220672
    per = total // world
    idx = list(range(rank * per, (rank + 1) * per))
    wav = torch.randn(per, 1, S, device="cuda")
    n = torch.full((per,), S, dtype=torch.int64, device="cuda")
    for mode in ("p2p", "allgather"):
        for _ in range(2):
            sharding.gather_waveforms(wav, n, idx, total, mode=mode)
        torch.cuda.synchronize()
        dist.barrier()
        ms = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record()
            out = sharding.gather_waveforms(wav, n, idx, total, mode=mode)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        if rank == 0:
            assert len(out) == total and all(o.numel() == S for o in out) and torch.equal(out[0], wav[0, 0])
        t = torch.tensor([sorted(ms)[len(ms) // 2]], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            inbound = (world - 1) * per * S * 4
            print("world %d  mode %-9s  gather median %.3f ms  needed at rank 0: %.1f MB  = %.1f GB/s  env %s" % (
                world, mode, float(t), inbound / 1e6, inbound / float(t) / 1e6,
                {k: v for k, v in os.environ.items() if k.startswith("NCCL_") and k not in ("NCCL_VERSION", "NCCL_DEBUG")}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
