#!/usr/bin/env python
"""Config-5 batch (64 utterances of 1-60 s, seed 1234, uudb_ms_istft_vits_ms) decoded in 1..8 length buckets: ms per batch.

    python tools/bucket_sweep.py [--out file]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mb_istft_vits_b200 import Engine, get_config, synth  # noqa: E402
from mb_istft_vits_b200.sharding import decode_in_buckets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    cfg = get_config("uudb_ms_istft_vits_ms")
    sd = synth.make_state_dict(cfg, seed=1234)
    eng = Engine(cfg, sd, precision="bf16", device=0)
    gen = torch.Generator().manual_seed(1234)
    secs = 1.0 + 59.0 * torch.rand(64, generator=gen)
    lengths = [max(1, int(round(float(s) * cfg["sampling_rate"] / 256))) for s in secs]
    z_p, _, _ = synth.make_latents(cfg, 64, max(lengths), seed=4000, lengths=torch.tensor(lengths))
    z_p = z_p.cuda()
    g = sd["emb_g.weight"][torch.arange(64) % cfg["n_speakers"]].unsqueeze(-1).cuda().contiguous()
    lines = []
    for k, ovh in ((1, 4000), (2, 4000), (3, 4000), (4, 4000), (6, 4000), (8, 4000), (8, 1000), (12, 1000)):
        flat, offs, plan = decode_in_buckets(eng, z_p, lengths, g=g, max_buckets=k, overhead=ovh)
        for _ in range(2):
            decode_in_buckets(eng, z_p, lengths, g=g, out=flat, plan=plan)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            decode_in_buckets(eng, z_p, lengths, g=g, out=flat, plan=plan)
        e1.record()
        torch.cuda.synchronize()
        padded = sum(b * T for _, T, _, _, b in plan[0])
        lines.append("max_buckets %2d overhead %5d -> %2d buckets %s  padded frames %6d (valid %d)  %.2f ms per batch" % (
            k, ovh, len(plan[0]), [(b, T) for _, T, _, _, b in plan[0]], padded, sum(lengths), e0.elapsed_time(e1) / 5))
        print(lines[-1], flush=True)
        del flat
    if a.out:
        open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
