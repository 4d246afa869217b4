#!/usr/bin/env python
"""Top stall sites of a kernel from an `ncu --set full --import-source on` report, as text small enough to bring back
from the GPU box (the .ncu-rep itself is ~40 MB):

    python tools/ncu_source_top.py /tmp/n/rs.ncu-rep [N] > gpurun_out/rs_source_top.txt

Reads `ncu -i <rep> --page source --csv` (SASS view with per-instruction warp-stall samples; needs -lineinfo), prints the
column names once, the N instructions with the most stall samples (address, source line, samples, executed count, SASS) and
the per-stall-reason totals when the report carries them.
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True)
    txt = out.stdout
    if not txt.strip():
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True)
        txt = out.stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # the CSV may hold several kernels, each with its own header row: split on rows whose first cell is "Address" / "#"
    hdr_idx = [i for i, r in enumerate(rows) if r and r[0] in ("Address", "#", "Source")]
    if not hdr_idx:
        print("no header found; first lines:")
        print("\n".join(txt.splitlines()[:20]))
        print(out.stderr[-2000:])
        return
    for k, hi in enumerate(hdr_idx):
        hdr = rows[hi]
        body = rows[hi + 1:(hdr_idx[k + 1] if k + 1 < len(hdr_idx) else len(rows))]
        print("=== %s  table %d: %d instructions" % (rep, k, len(body)))
        print("columns:", " | ".join(hdr))
        col = {h: i for i, h in enumerate(hdr)}
        samp = next((h for h in hdr if "Sampling (All" in h), None) or next((h for h in hdr if "Samples" in h), None)
        if samp is None:
            continue
        src = next((h for h in hdr if h in ("Source", "SASS", "Instruction")), hdr[1] if len(hdr) > 1 else hdr[0])
        exe = next((h for h in hdr if "Instructions Executed" in h and "Thread" not in h), None)
        stall_cols = [h for h in hdr if h.startswith("stall_") or "Stall" in h and h != samp]

        def num(r, h):
            try:
                return float(r[col[h]].replace(",", ""))
            except (ValueError, IndexError):
                return 0.0

        total = sum(num(r, samp) for r in body) or 1.0
        print("total samples (%s): %.0f" % (samp, total))
        for h in stall_cols:
            t = sum(num(r, h) for r in body)
            if t > 0:
                print("  %-60s %10.0f  %5.1f %%" % (h, t, 100 * t / total))
        best = sorted(body, key=lambda r: -num(r, samp))[:top_n]
        for r in best:
            extra = "  ".join("%s=%s" % (h.replace("stall_", ""), r[col[h]]) for h in stall_cols if num(r, h) > 0.15 * max(num(r, samp), 1))
            print("%8.0f %5.1f%%  exec %-10s %-70s %s" % (num(r, samp), 100 * num(r, samp) / total, r[col[exe]] if exe else "", r[col[src]][:70], extra[:120]))


if __name__ == "__main__":
    main()
