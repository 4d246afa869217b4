"""Drop-in replacements for ``SynthesizerTrn.flow`` and ``SynthesizerTrn.dec`` (models.py:634-647).

Usage (the reference's models.py stays untouched)::

    net_g = SynthesizerTrn(...); utils.load_checkpoint(path, net_g, None)      # reference code
    from mb_istft_vits_b200 import patch_synthesizer
    patch_synthesizer(net_g, cfg, precision="bf16")                             # swaps flow + dec
    o, o_mb, spec, phase, attn, y_mask, zs, timings = net_g.infer(x, x_lengths, sid=sid)

Signatures honoured:
  flow(x, x_mask, g=None, reverse=False) -> x            (both directions; no gradients: inference / voice conversion)
  dec(x, g=None) -> (o, o_mb, spec, phase)               o_mb is None for the single-band decoder
"""
from __future__ import annotations

import time

import torch
from torch import nn

from .engine import Engine


class NativeFlow(nn.Module):
    """ResidualCouplingBlock.forward(reverse=True) (models.py:207-214) on the B200 library."""

    def __init__(self, engine: Engine):
        super().__init__()
        self.engine = engine

    @torch.no_grad()
    def forward(self, x, x_mask, g=None, reverse=False):
        if not reverse:  # voice conversion (models.py:796); like the reference block it returns x only
            return self.engine.flow_forward(x, x_mask, g)
        return self.engine.flow_reverse(x, x_mask, g)


class NativeDecoder(nn.Module):
    """{iSTFT,Multiband_iSTFT,Multistream_iSTFT}_Generator.forward (models.py:278/344/430)."""

    def __init__(self, engine: Engine, want_mb=True, want_spec=True):
        super().__init__()
        self.engine = engine
        cfg = engine.cfg
        # attributes callers read (inferz_test.ipynb cell 6)
        self.gen_istft_n_fft = cfg["gen_istft_n_fft"]
        self.gen_istft_hop_size = cfg["gen_istft_hop_size"]
        self.subbands = cfg["subbands"]
        self.want_mb = want_mb
        self.want_spec = want_spec

    @torch.no_grad()
    def forward(self, x, g=None):
        return self.engine.decode(x, g, want_mb=self.want_mb, want_spec=self.want_spec)

    def remove_weight_norm(self):
        """Weight-norm is folded at load time; nothing to do (reference: models.py:299/379/469)."""


class NativeTextEncoder(nn.Module):
    """TextEncoder.forward (models.py:172-181): (x, x_lengths) -> (x, m, logs, x_mask)."""

    def __init__(self, engine: Engine):
        super().__init__()
        self.engine = engine

    @torch.no_grad()
    def forward(self, x, x_lengths):
        return self.engine.text_encode(x, x_lengths)


class NativePosteriorEncoder(nn.Module):
    """PosteriorEncoder.forward (models.py:236-246): (x, x_lengths, g=None) -> (z, m, logs, x_mask).  Used by
    ``SynthesizerTrn.voice_conversion`` (models.py:794); with it the whole conversion path runs on the library."""

    def __init__(self, engine: Engine):
        super().__init__()
        self.engine = engine

    @torch.no_grad()
    def forward(self, x, x_lengths, g=None):
        return self.engine.posterior_encode(x, x_lengths, g)


@torch.no_grad()
def infer_native(net_g, engine, x, x_lengths, sid=None, noise_scale=1, length_scale=1, noise_scale_w=1., max_len=None,
                 noise=None, want_attn=True):
    """SynthesizerTrn.infer (models.py:697-737) with everything after the duration predictor on the B200 library: the
    reference's own text encoder / duration predictor modules run as they are (``net_g.enc_p``, ``net_g.dp``,
    ``net_g.emb_g``); alignment expansion + prior sampling (models.py:717-729) is ``Engine.expand_prior`` and flow
    reverse + decoder (models.py:730-734) one ``Engine.flow_decode`` call.  Same return tuple as the reference, including
    the ``timings`` dict with the reference's keys (``time.time()`` deltas without a device synchronise, exactly what
    models.py:698-735 records: on a GPU they are launch times); ``attn`` is None when ``want_attn`` is False.  Because flow
    and decoder are one fused call here, its time is booked under 'flow' and 'waveform_decoder' is the (near-zero) rest."""
    timings = {}
    t0 = time.time()
    x, m_p, logs_p, x_mask = net_g.enc_p(x, x_lengths)
    timings['text_encoder'] = time.time() - t0
    g = net_g.emb_g(sid).unsqueeze(-1) if getattr(net_g, "n_speakers", 0) > 0 else None
    t0 = time.time()
    if getattr(net_g, "use_sdp", False):
        logw = net_g.dp(x, x_mask, g=g, reverse=True, noise_scale=noise_scale_w)
    else:
        logw = net_g.dp(x, x_mask, g=g)
    timings['duration_predictor'] = time.time() - t0
    t0 = time.time()
    w_ceil = torch.ceil(torch.exp(logw) * x_mask * length_scale)
    z_p, y_mask, y_lengths, attn, stats = engine.expand_prior(m_p, logs_p, w_ceil, noise_scale, x_mask=x_mask, noise=noise,
                                                              want_attn=want_attn, want_stats=True)
    timings['alignment_and_projection'] = time.time() - t0
    t0 = time.time()
    if max_len is not None:  # the reference decodes (z * y_mask)[:, :, :max_len]; the flow is causal-free, so run it in full
        z = engine.flow_reverse(z_p, y_mask, g)
        timings['flow'] = time.time() - t0
        t0 = time.time()
        o, o_mb, spec, phase = engine.decode((z * y_mask)[:, :, :max_len].contiguous(), g)
    else:
        z, o, o_mb, spec, phase = engine.flow_decode(z_p, y_mask, g, want_z=True, want_mb=True, want_spec=True)
        timings['flow'] = time.time() - t0
        t0 = time.time()
    timings['waveform_decoder'] = time.time() - t0
    return o, o_mb, spec, phase, attn, y_mask, (z, z_p, stats[0], stats[1]), timings


def patch_synthesizer(net_g, cfg, precision="bf16", device=0, flags=0, residual=None, posterior=False, text=False):
    """Replace net_g.flow and net_g.dec by the native path, using net_g's own weights.  posterior=True also replaces
    net_g.enc_q (PosteriorEncoder), so that ``net_g.voice_conversion`` (models.py:790-798) runs entirely on the library;
    text=True also replaces net_g.enc_p (TextEncoder).  Mind that a 16-bit text encoder can move a predicted duration
    across a ceil() boundary (models.py:717-718) and so change the utterance length; use 'tf32' / 'fp32' when the
    durations must match the reference exactly."""
    keep = ("dec.", "flow.") + (("enc_q.",) if posterior else ()) + (("enc_p.",) if text else ())
    sd = {k: v for k, v in net_g.state_dict().items() if k.startswith(keep)}
    eng = Engine(cfg, sd, precision=precision, device=device, flags=flags, residual=residual)
    net_g.flow = NativeFlow(eng)
    net_g.dec = NativeDecoder(eng)
    if posterior:
        net_g.enc_q = NativePosteriorEncoder(eng)
    if text:
        net_g.enc_p = NativeTextEncoder(eng)
    return eng
