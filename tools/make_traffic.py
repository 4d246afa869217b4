#!/usr/bin/env python
"""Summarise an ncu launch list (CSV of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv ... python tools/profile_step.py`) into the per-launch DRAM
traffic figures bench.py reports as `roofline.traffic`.

    python tools/make_traffic.py profiles/r02_launches.csv profiles/r02_traffic.json
"""
import collections
import csv
import json
import sys


def main(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    d = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < 15:
            continue
        d.setdefault(int(r[0]), {"name": r[4]})[r[12]] = float(r[14].replace(",", ""))
    # one step = from a pack_input launch to the tail launch that follows it; take the LAST complete step
    ids = sorted(d)
    tails = [i for i in ids if "tail_mb" in d[i]["name"] or "tail_kernel" in d[i]["name"] or "tail_fused" in d[i]["name"]]
    packs = [i for i in ids if "pack_input" in d[i]["name"]]
    # candidate steps: a pack_input launch up to the first tail launch after it; the full hot-path step (flow + decoder)
    # is the longest such run (decoder-only and stand-alone tail launches of the bench are shorter)
    best = []
    for p in packs:
        later = [t for t in tails if t > p]
        if not later:
            continue
        run = [i for i in ids if p <= i <= later[0]]
        if all("pack_input" not in d[i]["name"] for i in run[1:]) and len(run) >= len(best):
            best = run
    step = best
    step_tail = step[-1]
    conv = [d[i] for i in step if any(k in d[i]["name"] for k in ("conv_tc", "conv_pair", "pw_tc", "gate_tm", "pair_tm", "conv_tm"))]
    tail = d[step_tail]
    by = lambda x: x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)
    out = {
        "source": f"{src} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum; one step of tools/profile_step.py = the step bench.py times)",
        "launches_in_step": len(step),
        "conv_launches": len(conv),
        "conv_dram_bytes_per_step": sum(by(c) for c in conv),
        "conv_dram_bytes_per_launch": sum(by(c) for c in conv) / max(1, len(conv)),
        "tail_dram_bytes_per_launch": by(tail),
        "conv_us_sum_isolated": sum(c["gpu__time_duration.sum"] for c in conv) / 1e3,
        "tail_us_isolated": tail["gpu__time_duration.sum"] / 1e3,
        "step_us_sum_isolated": sum(d[i]["gpu__time_duration.sum"] for i in step) / 1e3,
    }
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
