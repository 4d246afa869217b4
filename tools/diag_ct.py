#!/usr/bin/env python
"""Diagnostic: SNR between the conv_tm path, the generic tensor-core path and the CUDA-core path (same 16-bit operands) on a few cases,
and where along the time axis the conv_tm / generic difference sits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import mbistft_oracle as orc
from mb_istft_vits_b200 import Engine, get_config, synth, lib as L
from helpers import load_case

def run(cfg, sd, flags, z_p, mask, g=None):
    e = Engine(cfg, sd, precision="bf16", flags=flags)
    o = e.flow_decode(z_p.cuda(), mask.cuda(), g)[1].float().cpu()
    torch.cuda.synchronize(); e.close()
    return o

cases = []
for name in ("mb", "mb_long", "mini_mb"):
    cfg, sd, t, meta = load_case(name)
    cases.append((name, cfg, sd, t["z_p"], t["mask"]))
cfg = get_config("ljs_mb_istft_vits"); sd = synth.make_state_dict(cfg, seed=1234)
for lengths in ([48], [40, 33, 7]):
    z_p, mask, _ = synth.make_latents(cfg, len(lengths), max(lengths), seed=5, lengths=lengths)
    cases.append(("ljs_mb %s" % lengths, cfg, sd, z_p, mask))
for name, cfg, sd, z_p, mask in cases:
    a = run(cfg, sd, 0, z_p, mask)
    b = run(cfg, sd, L.FLAG_NO_CONV_TM, z_p, mask)
    c = run(cfg, sd, L.FLAG_FORCE_SIMT, z_p, mask)
    d = run(cfg, sd, L.FLAG_NO_CONV_TM | L.FLAG_NO_PAIR_TM, z_p, mask)
    print("%-22s tm|generic %.1f  tm|simt %.1f  generic|simt %.1f  generic|no_pair_tm %.1f dB" % (name, orc.snr_db(a, b), orc.snr_db(a, c), orc.snr_db(b, c), orc.snr_db(b, d)))
    err = (a - b)[0, 0].double() ** 2
    n = err.numel(); seg = max(1, n // 24)
    sig = (b[0, 0].double() ** 2).mean().item()
    print("   error/signal per 1/24 of utterance 0 (dB):", " ".join("%.0f" % (10 * torch.log10(err[i:i + seg].mean() / sig + 1e-30)).item() for i in range(0, n - seg + 1, seg)))
