#!/usr/bin/env python
"""Mint golden vectors by running the UNMODIFIED reference modules (build container only).

The reference ships no tests or golden vectors for this path, so the parity pin
is "outputs of the reference itself run here".  This script imports models.py from /root/reference
with the three import shims of SURVEY.md section 8c (none of which touches the arithmetic), loads the seeded
state-dict of ``mb_istft_vits_b200.synth`` into the real ``SynthesizerTrn`` with strict key/shape
checks on ``dec.*`` / ``flow.*`` / ``emb_g.*``, runs

    z = net_g.flow(z_p, y_mask, g=g, reverse=True)          (models.py:730)
    o, o_mb, spec, phase = net_g.dec(z * y_mask, g=g)       (models.py:734)

on CPU fp32 and writes ``tests/golden/<case>.npz``.  /root/reference does not exist on the GPU box,
so nothing at test/bench time imports this file.

    python tools/make_golden.py            # regenerate every case
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from mb_istft_vits_b200 import configs as cfgs  # noqa: E402
from mb_istft_vits_b200 import synth  # noqa: E402

CASES = [
    # name, config, overrides, B, T, lengths, g_scale
    ("mini_mb", "ljs_mini_mb_istft_vits", {}, 2, 20, [20, 13], 1.0),
    ("mb", "ljs_mb_istft_vits", {}, 2, 20, [20, 13], 1.0),
    ("mb_gscale", "ljs_mb_istft_vits", {}, 1, 24, [24], 1.7),
    ("ms", "ljs_ms_istft_vits", {}, 2, 20, [20, 11], 1.0),
    ("istft", "ljs_istft_vits", {}, 2, 12, [12, 7], 1.0),
    ("mini_istft", "ljs_mini_istft_vits", {}, 1, 12, [12], 1.0),
    ("uudb_spk8", "uudb_spk8_istft_vits", {}, 1, 20, [20], 1.0),
    ("ms_spk", "uudb_ms_istft_vits_ms", {}, 3, 20, [20, 9, 15], 1.0),
    ("mb_resblock2", "ljs_mb_istft_vits",
     {"resblock": "2", "resblock_dilation_sizes": [[1, 3], [1, 3], [1, 3]]}, 1, 20, [20], 1.0),
    ("mb_long", "ljs_mini_mb_istft_vits", {}, 1, 150, [150], 1.0),
]


def import_reference():
    """The reference's `models` module through baseline/ref_loader.py (the three import shims live there)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_loader
    return ref_loader.import_reference()


def build_reference(models, cfg, sd):
    import ref_loader
    return ref_loader.build_synthesizer(cfg, sd)


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mbistft_oracle as orc
    models = import_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, cname, over, B, T, lengths, g_scale in CASES:
        cfg = cfgs.get_config(cname)
        cfg.update(over)
        sd = synth.make_state_dict(cfg, seed=1234, g_scale=g_scale)
        net = build_reference(models, cfg, sd)
        z_p, mask, lens = synth.make_latents(cfg, B, T, seed=4321, lengths=lengths)
        g = None
        sid = None
        if cfg["gin_channels"]:
            sid = torch.arange(B) % cfg["n_speakers"]
            g = net.emb_g(sid).unsqueeze(-1).detach()
        with torch.no_grad():
            z = net.flow(z_p, mask, g=g, reverse=True)
            o, o_mb, spec, phase = net.dec(z * mask, g=g)
            z_fwd = net.flow(z_p, mask, g=g)  # the forward direction (voice conversion, models.py:796) on the same input
        # cross-check the oracle right here
        zo, (oo, oo_mb, ospec, ophase) = orc.flow_decode(sd, cfg, z_p, mask, g)
        print(f"{name:14s} wav peak {o.abs().max():.4f}  oracle-vs-ref: z {(zo - z).abs().max():.2e} "
              f"wav {orc.max_abs_over_peak(oo, o):.2e} spec {orc.max_abs_over_peak(ospec, spec):.2e} "
              f"phase {(ophase - phase).abs().max():.2e}")
        zfo = orc.flow_forward(sd, cfg, z_p, mask, g)
        print(f"{'':14s} flow forward oracle-vs-ref {(zfo - z_fwd).abs().max():.2e}")
        out = dict(z_p=z_p.numpy(), mask=mask.numpy(), z=z.numpy(), o=o.numpy(),
                   spec=spec.numpy(), phase=phase.numpy(), z_fwd=z_fwd.numpy(),
                   meta=np.array([B, T, 1234, 4321], dtype=np.int64), g_scale=np.float32(g_scale),
                   lengths=np.array(lengths, dtype=np.int64))
        if o_mb is not None:
            out["o_mb"] = o_mb.numpy()
        if g is not None:
            out["g"] = g.numpy()
            out["sid"] = sid.numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)


def mint_infer_cases(models, orc):
    """Golden vectors from the reference's full SynthesizerTrn.infer() (BASELINE configs 1 and 4): synthetic phoneme
    ids through the reference text encoder / duration predictor / path expansion / prior sampling; the inputs the
    seam receives (z_p, y_mask, g) are captured with forward hooks and stored next to infer()'s outputs."""
    for name, cname, B, Tx in (("infer_mini_mb", "ljs_mini_mb_istft_vits", 1, 24), ("infer_istft", "ljs_istft_vits", 2, 12)):
        cfg = cfgs.get_config(cname)
        sd = synth.make_state_dict(cfg, seed=1234)
        torch.manual_seed(77)
        net = build_reference(models, cfg, sd)  # enc_p / dp keep their own (seeded) random init
        cap = {}
        hk = net.flow.register_forward_pre_hook(lambda m, args, kwargs: cap.update(z_p=args[0].detach().clone(), mask=args[1].detach().clone()), with_kwargs=True)
        x = torch.randint(1, 59, (B, Tx))
        x_len = torch.tensor([Tx] + [max(3, Tx - 5)] * (B - 1))
        with torch.no_grad():
            o, o_mb, spec, phase, attn, y_mask, (z, z_p, m_p, logs_p), timings = net.infer(x, x_len, length_scale=1.0)
        hk.remove()
        assert torch.equal(cap["z_p"], z_p) and torch.equal(cap["mask"], y_mask)
        zo, (oo, _, _, _) = orc.flow_decode(sd, cfg, z_p, y_mask, None)
        print(f"{name:14s} T={z_p.shape[-1]} oracle-vs-infer: z {(zo - z).abs().max():.2e} wav {orc.max_abs_over_peak(oo, o):.2e}"
              f"  timings keys {sorted(timings)}")
        out = dict(z_p=z_p.numpy(), mask=y_mask.numpy(), z=z.numpy(), o=o.numpy(), spec=spec.numpy(), phase=phase.numpy(),
                   meta=np.array([B, z_p.shape[-1], 1234, -1], dtype=np.int64), g_scale=np.float32(1.0),
                   lengths=y_mask.sum((1, 2)).long().numpy())
        if o_mb is not None:
            out["o_mb"] = o_mb.numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)


def mint_prior_cases(models, orc):
    """Golden vectors for the alignment expansion + prior sampling (models.py:717-729), taken from inside the reference's
    own infer(): forward hooks capture what enc_p and dp return, and infer()'s own attn,
    y_mask, z_p, expanded m_p / logs_p are the expected outputs.  The noise is recorded by wrapping torch.randn_like for
    the duration of the call."""
    for name, cname, B, Tx, ns, ls in (("prior_mini", "ljs_mini_mb_istft_vits", 3, 20, 0.667, 1.0),
                                       ("prior_long", "ljs_mini_mb_istft_vits", 2, 31, 1.0, 2.3)):
        cfg = cfgs.get_config(cname)
        sd = synth.make_state_dict(cfg, seed=1234)
        torch.manual_seed(91)
        net = build_reference(models, cfg, sd)
        cap = {}
        h1 = net.enc_p.register_forward_hook(lambda m, a, out: cap.update(m_p=out[1].detach().clone(), logs_p=out[2].detach().clone(),
                                                                          x_mask=out[3].detach().clone()))
        h2 = net.dp.register_forward_hook(lambda m, a, out: cap.update(logw=out.detach().clone()))
        x = torch.randint(1, 59, (B, Tx))
        x_len = torch.tensor([Tx] + [max(3, Tx - 4 * (i + 1)) for i in range(B - 1)])
        torch.manual_seed(2024)
        _randn_like = torch.randn_like

        def recording_randn_like(t, *a, **k):  # records the prior noise infer() draws; the draw itself is untouched
            r = _randn_like(t, *a, **k)
            cap.setdefault("noise", []).append(r.detach().clone())
            return r
        torch.randn_like = recording_randn_like
        try:
            with torch.no_grad():
                o, o_mb, spec, phase, attn, y_mask, (z, z_p, m_exp, logs_exp), timings = net.infer(x, x_len, noise_scale=ns, length_scale=ls)
        finally:
            torch.randn_like = _randn_like
        h1.remove(); h2.remove()
        w_ceil = torch.ceil(torch.exp(cap["logw"]) * cap["x_mask"] * ls)          # models.py:717-718
        noise = [n for n in cap["noise"] if n.shape == m_exp.shape][-1].contiguous()  # the randn_like(m_p) of models.py:729
        assert torch.equal(m_exp + noise * torch.exp(logs_exp) * ns, z_p), "noise capture failed"
        got = orc.expand_prior(cap["m_p"], cap["logs_p"], w_ceil, noise, ns, cap["x_mask"])
        print(f"{name:12s} Tx={Tx} Ty={z_p.shape[-1]} y_len={got[5].tolist()} oracle-vs-infer: z_p {(got[0] - z_p).abs().max():.1e} "
              f"attn {(got[2] - attn).abs().max():.0f} mask {(got[1] - y_mask).abs().max():.0f}")
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"),
                            m_p=cap["m_p"].numpy(), logs_p=cap["logs_p"].numpy(), x_mask=cap["x_mask"].numpy(),
                            w_ceil=w_ceil.numpy(), noise=noise.numpy(), noise_scale=np.float32(ns),
                            z_p=z_p.numpy(), y_mask=y_mask.numpy(), attn=attn.numpy(), m_exp=m_exp.numpy(),
                            logs_exp=logs_exp.numpy())


def mint_vc_cases(models, orc):
    """Golden vectors for the posterior encoder and the voice-conversion path (models.py:217-246, 790-798): the reference's
    own enc_q / flow / dec on seeded enc_q.* weights (synth.make_state_dict(enc_q=True)); the noise enc_q draws is
    recorded by wrapping torch.randn_like for the duration of the call."""
    for name, cname, B, T, lens in (("vc_ms_spk", "uudb_ms_istft_vits_ms", 2, 24, [24, 17]),
                                    ("posterior_mini", "ljs_mini_mb_istft_vits", 2, 30, [30, 21])):
        cfg = cfgs.get_config(cname)
        sd = synth.make_state_dict(cfg, seed=1234, enc_q=True)
        net = build_reference(models, cfg, sd)
        y = torch.randn((B, 513, T), generator=torch.Generator().manual_seed(8)).abs()   # a magnitude spectrogram
        y_len = torch.tensor(lens)
        cap = {}
        _randn_like = torch.randn_like

        def recording_randn_like(t, *a, **k):
            r = _randn_like(t, *a, **k)
            cap.setdefault("noise", []).append(r.detach().clone())
            return r
        torch.manual_seed(99)
        torch.randn_like = recording_randn_like
        out = {}
        try:
            with torch.no_grad():
                if cfg["n_speakers"]:
                    sid_s, sid_t = torch.arange(B) % cfg["n_speakers"], (torch.arange(B) + 5) % cfg["n_speakers"]
                    g_s, g_t = net.emb_g(sid_s).unsqueeze(-1), net.emb_g(sid_t).unsqueeze(-1)
                    o_hat, o_hat_mb, y_mask, (z, z_p, z_hat) = net.voice_conversion(y, y_len, sid_s, sid_t)
                    out.update(g_src=g_s.numpy(), g_tgt=g_t.numpy(), sid_src=sid_s.numpy(), sid_tgt=sid_t.numpy(),
                               z_p=z_p.numpy(), z_hat=z_hat.numpy(), o=o_hat.numpy())
                    noise = cap["noise"][0]
                    _, m_q, logs_q, _ = None, None, None, None
                    torch.manual_seed(99)
                    z2, m_q, logs_q, _ = net.enc_q(y, y_len, g=g_s)
                    assert torch.equal(z2, z)
                else:
                    z, m_q, logs_q, y_mask = net.enc_q(y, y_len, g=None)
                    noise = cap["noise"][0]
        finally:
            torch.randn_like = _randn_like
        assert torch.equal((m_q + noise * torch.exp(logs_q)) * y_mask, z), "noise capture failed"
        g_s_t = torch.from_numpy(out["g_src"]) if "g_src" in out else None
        zo, mo, lo, _ = orc.posterior_encoder(sd, cfg, y, y_len, g_s_t, noise)
        print(f"{name:16s} oracle-vs-ref: z {(zo - z).abs().max():.2e} m {(mo - m_q).abs().max():.2e} logs {(lo - logs_q).abs().max():.2e}")
        if "o" in out:
            oo = orc.voice_conversion(sd, cfg, y, y_len, g_s_t, torch.from_numpy(out["g_tgt"]), noise)[0]
            print(f"{'':16s} voice conversion wav oracle-vs-ref {orc.max_abs_over_peak(oo, torch.from_numpy(out['o'])):.2e}")
        out.update(y=y.numpy(), y_lengths=y_len.numpy(), noise=noise.numpy(), z=z.numpy(), m=m_q.numpy(), logs=logs_q.numpy(),
                   y_mask=y_mask.numpy())
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)


def mint_text_cases(models, orc):
    """Golden vectors for the text encoder (models.py:140-181): the reference's own enc_p on seeded enc_p.* weights
    (synth.make_state_dict(enc_p=True)) and synthetic phoneme ids."""
    for name, cname, B, Tx, lens in (("text_mb", "ljs_mb_istft_vits", 3, 37, [37, 20, 5]),
                                     ("text_mini", "ljs_mini_mb_istft_vits", 2, 150, [150, 97]),
                                     ("text_short", "ljs_mb_istft_vits", 2, 3, [3, 1])):
        cfg = cfgs.get_config(cname)
        sd = synth.make_state_dict(cfg, seed=1234, enc_p=True)
        net = build_reference(models, cfg, sd)
        tokens = torch.randint(0, 59, (B, Tx), generator=torch.Generator().manual_seed(12))
        x_len = torch.tensor(lens)
        with torch.no_grad():
            x, m, logs, x_mask = net.enc_p(tokens, x_len)
        xo, mo, lo, mk = orc.text_encoder(sd, tokens, x_len)
        print(f"{name:12s} oracle-vs-ref: x {(xo - x).abs().max():.2e} m {(mo - m).abs().max():.2e} logs {(lo - logs).abs().max():.2e}")
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), tokens=tokens.numpy(), x_lengths=x_len.numpy(),
                            x=x.numpy(), m=m.numpy(), logs=logs.numpy(), x_mask=x_mask.numpy())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "prior":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import mbistft_oracle as _orc
        mint_prior_cases(import_reference(), _orc)
    elif len(sys.argv) > 1 and sys.argv[1] == "text":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import mbistft_oracle as _orc
        mint_text_cases(import_reference(), _orc)
    elif len(sys.argv) > 1 and sys.argv[1] == "vc":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import mbistft_oracle as _orc
        mint_vc_cases(import_reference(), _orc)
    elif len(sys.argv) > 1 and sys.argv[1] == "infer":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import mbistft_oracle as _orc
        mint_infer_cases(import_reference(), _orc)
    else:
        main()
