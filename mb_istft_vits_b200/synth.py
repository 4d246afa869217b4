"""Seeded random-init weights in the reference checkpoint layout.

There is no network for checkpoints, so benchmarks and parity tests use random-init weights of the
named architecture.  This generator emits a ``{'dec.*', 'flow.*', 'emb_g.weight'}`` state-dict with
exactly the key names and shapes ``SynthesizerTrn.state_dict()`` has for those sub-modules
(models.py:257-273, 316-336, 394-426, 647, 654; modules.py:126-146, 191-211, 241-249, 328-332):
weight-normed convs as ``weight_g`` / ``weight_v``.  ``tools/make_golden.py`` loads it into the real
reference model with strict key/shape checking, which pins the inventory.

Distributions follow what the reference ends up with at construction time (SURVEY.md section 8c traps):
weight_v ~ U(+-1/sqrt(fan_in)) (PyTorch default; ``init_weights`` does not reach weight_v),
weight_g = ||v|| * g_scale, and the zero-initialised ``flow.*.post`` layers are re-randomised
(N(0, 0.05)) because with zeros the flow reverse is a pure channel permutation.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .configs import FLOW_KERNEL, FLOW_LAYERS, FLOW_N


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def _wn_conv(sd, gen, name, shape, fan_in, g_scale, bias=True, bias_len=None):
    v = _uniform(gen, shape, 1.0 / math.sqrt(fan_in))
    g = v.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], 1, 1) * g_scale
    sd[name + ".weight_g"] = g
    sd[name + ".weight_v"] = v
    if bias:
        sd[name + ".bias"] = _uniform(gen, (bias_len if bias_len is not None else shape[0],), 1.0 / math.sqrt(fan_in))


def _plain_conv(sd, gen, name, shape, fan_in):
    sd[name + ".weight"] = _uniform(gen, shape, 1.0 / math.sqrt(fan_in))
    sd[name + ".bias"] = _uniform(gen, (shape[0],), 1.0 / math.sqrt(fan_in))


def make_state_dict(cfg, seed: int = 1234, g_scale: float = 1.0, enc_q: bool = False, spec_channels: int = 513,
                    enc_q_layers: int = 16, enc_p: bool = False, n_vocab: int = 59) -> Dict[str, torch.Tensor]:
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    inter, hid, gin = cfg["inter_channels"], cfg["hidden_channels"], cfg["gin_channels"]
    c0 = cfg["upsample_initial_channel"]

    # ---- decoder
    _wn_conv(sd, gen, "dec.conv_pre", (c0, inter, 7), inter * 7, g_scale)
    ch = c0
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        cin, cout = c0 // (2 ** i), c0 // (2 ** (i + 1))
        # ConvTranspose1d weight is [C_in, C_out, K]; weight-norm dim 0 = in-channels; bias is [C_out]
        _wn_conv(sd, gen, f"dec.ups.{i}", (cin, cout, k), cout * k, g_scale, bias_len=cout)
        ch = cout
        for j, (rk, dils) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            p = f"dec.resblocks.{i * nk + j}"
            if cfg["resblock"] == "1":
                for q in range(3):
                    _wn_conv(sd, gen, f"{p}.convs1.{q}", (ch, ch, rk), ch * rk, g_scale)
                for q in range(3):
                    _wn_conv(sd, gen, f"{p}.convs2.{q}", (ch, ch, rk), ch * rk, g_scale)
            else:
                for q in range(2):
                    _wn_conv(sd, gen, f"{p}.convs.{q}", (ch, ch, rk), ch * rk, g_scale)
            if gin:
                _plain_conv(sd, gen, f"{p}.cond", (ch, gin, 1), gin)
    n_post = cfg["gen_istft_n_fft"] + 2
    if cfg["variant"] == "istft":
        _wn_conv(sd, gen, "dec.conv_post", (n_post, ch, 7), ch * 7, g_scale)
    else:
        S = cfg["subbands"]
        _wn_conv(sd, gen, "dec.subband_conv_post", (S * n_post, ch, 7), ch * 7, g_scale)
        if cfg["variant"] == "ms":
            filt = torch.zeros(S, S, S)
            for k in range(S):
                filt[k, k, 0] = 1.0
            sd["dec.updown_filter"] = filt
            _wn_conv(sd, gen, "dec.multistream_conv_post", (1, S, 63), S * 63, g_scale, bias=False)

    # ---- flow (4 x ResidualCouplingLayer at even indices; Flip has no parameters)
    half = inter // 2
    for f in range(FLOW_N):
        p = f"flow.flows.{2 * f}"
        _plain_conv(sd, gen, f"{p}.pre", (hid, half, 1), half)
        if gin:
            _wn_conv(sd, gen, f"{p}.enc.cond_layer", (2 * hid * FLOW_LAYERS, gin, 1), gin, g_scale)
        for l in range(FLOW_LAYERS):
            _wn_conv(sd, gen, f"{p}.enc.in_layers.{l}", (2 * hid, hid, FLOW_KERNEL), hid * FLOW_KERNEL, g_scale)
            rs = 2 * hid if l < FLOW_LAYERS - 1 else hid
            _wn_conv(sd, gen, f"{p}.enc.res_skip_layers.{l}", (rs, hid, 1), hid, g_scale)
        sd[f"{p}.post.weight"] = torch.randn((half, hid, 1), generator=gen) * 0.05
        sd[f"{p}.post.bias"] = torch.randn((half,), generator=gen) * 0.05

    if cfg.get("n_speakers", 0) > 1 and gin:
        sd["emb_g.weight"] = torch.randn((cfg["n_speakers"], gin), generator=gen)

    # ---- posterior encoder (models.py:217-246, built at models.py:646 with kernel 5 / dilation_rate 1 / 16 layers); drawn
    # after everything else so that the dec.* / flow.* / emb_g tensors of a seed do not depend on this switch
    if enc_q:
        _plain_conv(sd, gen, "enc_q.pre", (hid, spec_channels, 1), spec_channels)
        if gin:
            _wn_conv(sd, gen, "enc_q.enc.cond_layer", (2 * hid * enc_q_layers, gin, 1), gin, g_scale)
        for l in range(enc_q_layers):
            _wn_conv(sd, gen, f"enc_q.enc.in_layers.{l}", (2 * hid, hid, FLOW_KERNEL), hid * FLOW_KERNEL, g_scale)
            rs = 2 * hid if l < enc_q_layers - 1 else hid
            _wn_conv(sd, gen, f"enc_q.enc.res_skip_layers.{l}", (rs, hid, 1), hid, g_scale)
        _plain_conv(sd, gen, "enc_q.proj", (2 * inter, hid, 1), hid)

    # ---- text encoder (models.py:140-181; attentions.py:13-47): the reference builds it with n_heads 2, kernel 3,
    # window 4 and filter_channels / n_layers from the config (768 / 6; 384 / 3 for the mini configs).  Initial
    # distributions as in the reference constructors; LayerNorm gains / offsets are perturbed so that they matter.
    if enc_p:
        filt, n_layers, n_heads, ks, win = (768, 6, 2, 3, 4) if hid == 192 else (384, 3, 2, 3, 4)
        dk = hid // n_heads
        sd["enc_p.emb.weight"] = torch.randn((n_vocab, hid), generator=gen) * hid ** -0.5
        for i in range(n_layers):
            a = f"enc_p.encoder.attn_layers.{i}"
            sd[a + ".emb_rel_k"] = torch.randn((1, 2 * win + 1, dk), generator=gen) * dk ** -0.5
            sd[a + ".emb_rel_v"] = torch.randn((1, 2 * win + 1, dk), generator=gen) * dk ** -0.5
            for n in ("conv_q", "conv_k", "conv_v", "conv_o"):
                bound = math.sqrt(6.0 / (2 * hid)) if n != "conv_o" else 1.0 / math.sqrt(hid)   # xavier_uniform / default
                sd[f"{a}.{n}.weight"] = _uniform(gen, (hid, hid, 1), bound)
                sd[f"{a}.{n}.bias"] = _uniform(gen, (hid,), 1.0 / math.sqrt(hid))
            for j in (1, 2):
                sd[f"enc_p.encoder.norm_layers_{j}.{i}.gamma"] = 1.0 + 0.1 * torch.randn((hid,), generator=gen)
                sd[f"enc_p.encoder.norm_layers_{j}.{i}.beta"] = 0.1 * torch.randn((hid,), generator=gen)
            _plain_conv(sd, gen, f"enc_p.encoder.ffn_layers.{i}.conv_1", (filt, hid, ks), hid * ks)
            _plain_conv(sd, gen, f"enc_p.encoder.ffn_layers.{i}.conv_2", (hid, filt, ks), filt * ks)
        _plain_conv(sd, gen, "enc_p.proj", (2 * inter, hid, 1), hid)
    return sd


def make_latents(cfg, B: int, T: int, seed: int = 1234, lengths=None):
    """z ~ N(0,1) [B, inter, T] times the length mask, the mask [B,1,T], and g (or None)."""
    gen = torch.Generator().manual_seed(seed)
    z = torch.randn((B, cfg["inter_channels"], T), generator=gen, dtype=torch.float32)
    if lengths is None:
        lengths = torch.full((B,), T, dtype=torch.long)
    lengths = torch.as_tensor(lengths, dtype=torch.long)
    mask = (torch.arange(T)[None, :] < lengths[:, None]).float().unsqueeze(1)
    return z * mask, mask, lengths
