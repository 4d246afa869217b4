#!/usr/bin/env python
"""Where does the time of the config-5 gather go?  (diagnostic; torchrun --nproc-per-node 2 tools/diag_gather5.py)

Per call of sharding.gather_waveforms on config-5-sized blocks (64 utterances x 256 x ~3700 frames per rank): CUDA-event time
and wall time, (a) alone, (b) after a ~30 ms busy kernel on every rank, (c) after the real flow_decode; then the phases of
one call (table, all_gather, host read, receive buffer, send / recv, assemble) with a synchronise after each.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mb_istft_vits_b200 import sharding  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))))
    per = 64
    total = per * world
    T = 3704 - 33 * rank
    S = 256 * T
    idx = list(range(rank, total, world))
    wav = torch.randn(per, 1, S, device="cuda")
    n = torch.randint(256, S, (per,), dtype=torch.int64, device="cuda")
    busy = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)

    def run(tag, pre, reps=6):
        for i in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            w0 = time.perf_counter()
            e0.record()
            pre()
            e1.record()
            out = sharding.gather_waveforms(wav, n, idx, total)
            e2.record()
            torch.cuda.synchronize()
            w1 = time.perf_counter()
            print(f"[rank {rank}] {tag} call {i}: pre {e0.elapsed_time(e1):8.3f} ms  gather {e1.elapsed_time(e2):8.3f} ms  wall {(w1 - w0) * 1e3:8.3f} ms",
                  flush=True)
            del out

    run("alone", lambda: None)

    def spin():
        for _ in range(24):
            torch.mm(busy, busy)
    run("after 30 ms of GEMMs", spin)

    from mb_istft_vits_b200 import Engine, get_config, synth
    cfg = get_config("uudb_ms_istft_vits_ms")
    sd = synth.make_state_dict(cfg, seed=1234)
    eng = Engine(cfg, sd, precision="bf16", device=torch.cuda.current_device())
    z_p, mask, _ = synth.make_latents(cfg, per, T, seed=4000 + rank)
    z_p, mask = z_p.cuda(), mask.cuda()
    g = sd["emb_g.weight"][torch.arange(per) % cfg["n_speakers"]].unsqueeze(-1).cuda().contiguous()
    eng.flow_decode(z_p, mask, g, want_z=False, out_wav=wav)
    run("after flow_decode", lambda: eng.flow_decode(z_p, mask, g, want_z=False, out_wav=wav))

    # phases of one call, a synchronise after each
    def phase(name, t0):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        print(f"[rank {rank}]   {name:28s} {(t1 - t0) * 1e3:8.3f} ms", flush=True)
        return time.perf_counter()

    for rep in range(2):
        dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        dev = wav.device
        flat = wav.contiguous().reshape(-1)
        table = torch.full((total + 1, 2), -1, dtype=torch.int64, device=dev)
        ii = torch.as_tensor(idx, dtype=torch.int64, device=dev)
        table[ii, 0] = n
        table[ii, 1] = torch.arange(per, dtype=torch.int64, device=dev) * S
        table[total, 0] = flat.numel()
        table[total, 1] = per
        t = phase("table", t)
        tables = torch.empty((world * (total + 1), 2), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(tables, table)
        t = phase("all_gather of the table", t)
        host = tables.cpu().view(world, total + 1, 2)
        t = phase("host read", t)
        sizes = [int(host[r, total, 0]) for r in range(world)]
        ops, bufs = [], [None] * world
        if rank == 0:
            for r in range(1, world):
                bufs[r] = torch.empty(sizes[r], dtype=torch.float32, device=dev)
                ops.append(dist.P2POp(dist.irecv, bufs[r], r))
        else:
            ops.append(dist.P2POp(dist.isend, flat, 0))
        t = phase("receive buffers", t)
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        t = phase("send / recv", t)
        if rank == 0:
            bufs[0] = flat
            sharding._assemble(host, bufs, world, total)
        t = phase("assemble", t)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
