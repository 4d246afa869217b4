"""CPU oracle for the MB-iSTFT-VITS waveform hot path (flow reverse + iSTFT decoders).

TEST INFRASTRUCTURE ONLY.  Nothing under ``mb_istft_vits_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` legs use it, and there only as the checker (or
as the CPU arm being timed), never as the thing shipped.

It is a functional restatement (plain ``torch`` fp32 on the CPU, no nn.Module,
no reference code) of the algorithm in the reference files cited per function.
The arithmetic primitives live in PyTorch (``F.conv1d``, ``F.conv_transpose1d``,
``torch.istft``), exactly as in the reference: the reference pins
no torch version; this image ships torch 2.11.0+cu128.

Parity pin: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the pin is "outputs of the reference itself run here":
``tools/make_golden.py`` imports the unmodified reference modules from
/root/reference in the build container, runs them on seeded weights/inputs and
commits the results under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this oracle against every one of those vectors.

Weights are consumed in the reference checkpoint layout (``utils.py:57-60``
'model' state-dict): ``dec.*`` / ``flow.*`` keys, weight-normed convs stored as
``weight_g`` / ``weight_v`` (plain ``weight`` is accepted too).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # modules.py:17


# ----------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------
def effective_weight(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """torch.nn.utils.weight_norm(dim=0) fold: w = g * v / ||v|| with the norm over
    every dim but 0 (models.py:257,262,336 / modules.py:128-146,191-206).
    For ConvTranspose1d dim 0 is the *input* channel."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"].float()
    v = sd[prefix + ".weight_v"].float()
    g = sd[prefix + ".weight_g"].float()
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / norm)


def _bias(sd, prefix):
    b = sd.get(prefix + ".bias")
    return None if b is None else b.float()


def _conv_same(x, sd, prefix, dilation=1):
    """Conv1d, stride 1, zero 'same' padding d*(K-1)/2 (commons.py:14-15)."""
    w = effective_weight(sd, prefix)
    k = w.shape[-1]
    return F.conv1d(x, w, _bias(sd, prefix), padding=(k * dilation - dilation) // 2, dilation=dilation)


# ----------------------------------------------------------------------------
# flow reverse: models.py:207-214, modules.py:148-176, 280-287, 334-353
# ----------------------------------------------------------------------------
def wn_forward(x, x_mask, sd, prefix, hidden, n_layers=4, kernel_size=5, dilation_rate=1, g=None):
    """modules.WN.forward (modules.py:148-176) with the gate of commons.py:100-107."""
    out = torch.zeros_like(x)
    if g is not None:
        g = F.conv1d(g, effective_weight(sd, prefix + ".cond_layer"), _bias(sd, prefix + ".cond_layer"))
    for i in range(n_layers):
        x_in = _conv_same(x, sd, f"{prefix}.in_layers.{i}", dilation=dilation_rate ** i)
        if g is not None:
            x_in = x_in + g[:, 2 * hidden * i: 2 * hidden * (i + 1), :]
        acts = torch.tanh(x_in[:, :hidden]) * torch.sigmoid(x_in[:, hidden:])
        rs = _conv_same(acts, sd, f"{prefix}.res_skip_layers.{i}")
        if i < n_layers - 1:
            x = (x + rs[:, :hidden]) * x_mask
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out * x_mask


def coupling_reverse(x, x_mask, sd, prefix, hidden, g=None):
    """modules.ResidualCouplingLayer.forward(reverse=True), mean_only (modules.py:334-353)."""
    half = x.shape[1] // 2
    x0, x1 = x[:, :half], x[:, half:]
    h = F.conv1d(x0, sd[prefix + ".pre.weight"].float(), sd[prefix + ".pre.bias"].float()) * x_mask
    h = wn_forward(h, x_mask, sd, prefix + ".enc", hidden, g=g)
    m = F.conv1d(h, sd[prefix + ".post.weight"].float(), sd[prefix + ".post.bias"].float()) * x_mask
    x1 = (x1 - m) * x_mask  # logs == 0 for mean_only
    return torch.cat([x0, x1], 1)


def flow_reverse(sd, cfg, z_p, y_mask, g=None, prefix="flow"):
    """ResidualCouplingBlock.forward(reverse=True) (models.py:207-214): for flow in
    reversed(flows): Flip, RCL3, Flip, RCL2, ... (Flip = channel reversal, modules.py:280-287)."""
    hidden = cfg["hidden_channels"]
    x = z_p.float()
    for i in reversed(range(4)):
        x = torch.flip(x, [1])
        x = coupling_reverse(x, y_mask, sd, f"{prefix}.flows.{2 * i}", hidden, g=g)
    return x


def flow_forward(sd, cfg, x, y_mask, g=None, prefix="flow"):
    """ResidualCouplingBlock.forward(reverse=False) (models.py:207-210), the direction voice conversion uses
    (models.py:790-798): RCL0, Flip, RCL1, Flip, ...; mean-only coupling x1 <- m + x1 * mask (modules.py:345-347).  The
    log-determinant the layers return is discarded by the block, as in the reference."""
    hidden = cfg["hidden_channels"]
    x = x.float()
    for i in range(4):
        pfx = f"{prefix}.flows.{2 * i}"
        half = x.shape[1] // 2
        x0, x1 = x[:, :half], x[:, half:]
        h = F.conv1d(x0, sd[pfx + ".pre.weight"].float(), sd[pfx + ".pre.bias"].float()) * y_mask
        h = wn_forward(h, y_mask, sd, pfx + ".enc", hidden, g=g)
        m = F.conv1d(h, sd[pfx + ".post.weight"].float(), sd[pfx + ".post.bias"].float()) * y_mask
        x = torch.cat([x0, m + x1 * y_mask], 1)
        x = torch.flip(x, [1])
    return x


def posterior_encoder(sd, cfg, y, y_lengths, g=None, noise=None, prefix="enc_q", n_layers=16):
    """PosteriorEncoder.forward (models.py:236-246; built with kernel 5, dilation_rate 1, 16 layers at models.py:646):
    x = pre(y) * mask; x = WN(x, mask, g); stats = proj(x) * mask; m, logs = split(stats);
    z = (m + noise * exp(logs)) * mask.  Returns (z, m, logs, mask)."""
    hidden, inter = cfg["hidden_channels"], cfg["inter_channels"]
    T = y.shape[2]
    mask = (torch.arange(T)[None, :] < torch.as_tensor(y_lengths)[:, None]).to(y.dtype).unsqueeze(1)  # commons.py:121-125
    x = F.conv1d(y.float(), sd[prefix + ".pre.weight"].float(), sd[prefix + ".pre.bias"].float()) * mask
    x = wn_forward(x, mask, sd, prefix + ".enc", hidden, n_layers=n_layers, g=g)
    stats = F.conv1d(x, sd[prefix + ".proj.weight"].float(), sd[prefix + ".proj.bias"].float()) * mask
    m, logs = stats[:, :inter], stats[:, inter:]
    if noise is None:
        noise = torch.randn_like(m)
    z = (m + noise * torch.exp(logs)) * mask
    return z, m, logs, mask


def voice_conversion(sd, cfg, y, y_lengths, g_src, g_tgt, noise):
    """SynthesizerTrn.voice_conversion (models.py:790-798) with the embeddings already looked up."""
    z, m_q, logs_q, y_mask = posterior_encoder(sd, cfg, y, y_lengths, g_src, noise)
    z_p = flow_forward(sd, cfg, z, y_mask, g_src)
    z_hat = flow_reverse(sd, cfg, z_p, y_mask, g_tgt)
    o_hat, o_hat_mb, _, _ = decode(sd, cfg, z_hat * y_mask, g_tgt)
    return o_hat, o_hat_mb, y_mask, (z, z_p, z_hat)


# ----------------------------------------------------------------------------
# text encoder: models.py:140-181, attentions.py:13-47 (Encoder), 101-254 (MultiHeadAttention), 257-303 (FFN),
# modules.py:20-33 (LayerNorm)
# ----------------------------------------------------------------------------
def _layer_norm_channels(x, gamma, beta, eps=1e-5):
    """modules.LayerNorm: normalise over the channel axis of [B, C, T]."""
    mean = x.mean(1, keepdim=True)
    var = ((x - mean) ** 2).mean(1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def relative_attention(q, k, v, mask, emb_rel_k, emb_rel_v, n_heads, window=4):
    """MultiHeadAttention.attention with window_size=4, heads_share=True (attentions.py:142-178): scores = q k^T / sqrt(dk)
    plus the relative-key logits q . E_k[j - i] for |j - i| <= window, masked_fill(mask == 0, -1e4), softmax, p v plus the
    relative-value term sum_r p[i, i + r] E_v[r].  q, k, v: [B, C, T]; mask: [B, 1, T] (1 = valid)."""
    B, C, T = q.shape
    dk = C // n_heads
    qh = q.view(B, n_heads, dk, T).transpose(2, 3) / math.sqrt(dk)
    kh = k.view(B, n_heads, dk, T).transpose(2, 3)
    vh = v.view(B, n_heads, dk, T).transpose(2, 3)
    scores = qh @ kh.transpose(-2, -1)                                   # [B, H, T, T]
    idx = torch.arange(T)
    rel = idx[None, :] - idx[:, None]                                    # rel[i, j] = j - i
    inside = rel.abs() <= window
    e_k = emb_rel_k[0]                                                   # [2w+1, dk]
    rel_logits = qh @ e_k.t()                                            # [B, H, T, 2w+1]
    gathered = torch.gather(rel_logits, 3, (rel.clamp(-window, window) + window).expand(B, n_heads, T, T))
    scores = scores + gathered * inside
    amask = mask.unsqueeze(2) * mask.unsqueeze(-1)                       # [B, 1, T, T]
    scores = scores.masked_fill(amask == 0, -1e4)
    p = F.softmax(scores, dim=-1)
    out = p @ vh
    e_v = emb_rel_v[0]
    for r in range(-window, window + 1):                                  # sum_r p[i, i + r] E_v[r]
        i0, i1 = max(0, -r), min(T, T - r)
        if i1 > i0:
            pr = p[:, :, torch.arange(i0, i1), torch.arange(i0 + r, i1 + r)]   # [B, H, n]
            out[:, :, i0:i1] = out[:, :, i0:i1] + pr.unsqueeze(-1) * e_v[r + window]
    return out.transpose(2, 3).contiguous().view(B, C, T)


def text_encoder(sd, x_tokens, x_lengths, prefix="enc_p", n_heads=2, window=4):
    """TextEncoder.forward (models.py:172-181): embedding * sqrt(H), mask, n_layers x [x = LN(x + attn(x));
    x = LN(x + FFN(x))] (attentions.py:35-47), x * mask, stats = proj(x) * mask.  Returns (x, m, logs, x_mask)."""
    emb = sd[prefix + ".emb.weight"].float()
    H = emb.shape[1]
    T = x_tokens.shape[1]
    mask = (torch.arange(T)[None, :] < torch.as_tensor(x_lengths)[:, None]).float().unsqueeze(1)
    x = (emb[x_tokens] * math.sqrt(H)).transpose(1, 2) * mask
    e = prefix + ".encoder"
    n_layers = 0
    while f"{e}.attn_layers.{n_layers}.conv_q.weight" in sd:
        n_layers += 1
    w = lambda name: sd[name].float()
    for i in range(n_layers):
        a = f"{e}.attn_layers.{i}"
        q = F.conv1d(x, w(a + ".conv_q.weight"), w(a + ".conv_q.bias"))
        k = F.conv1d(x, w(a + ".conv_k.weight"), w(a + ".conv_k.bias"))
        v = F.conv1d(x, w(a + ".conv_v.weight"), w(a + ".conv_v.bias"))
        y = relative_attention(q, k, v, mask, w(a + ".emb_rel_k"), w(a + ".emb_rel_v"), n_heads, window)
        y = F.conv1d(y, w(a + ".conv_o.weight"), w(a + ".conv_o.bias"))
        x = _layer_norm_channels(x + y, w(f"{e}.norm_layers_1.{i}.gamma"), w(f"{e}.norm_layers_1.{i}.beta"))
        f = f"{e}.ffn_layers.{i}"
        ks = sd[f + ".conv_1.weight"].shape[2]
        pad = ((ks - 1) // 2, ks // 2)
        y = F.conv1d(F.pad(x * mask, pad), w(f + ".conv_1.weight"), w(f + ".conv_1.bias"))
        y = torch.relu(y)
        y = F.conv1d(F.pad(y * mask, pad), w(f + ".conv_2.weight"), w(f + ".conv_2.bias")) * mask
        x = _layer_norm_channels(x + y, w(f"{e}.norm_layers_2.{i}.gamma"), w(f"{e}.norm_layers_2.{i}.beta"))
    x = x * mask
    stats = F.conv1d(x, w(prefix + ".proj.weight"), w(prefix + ".proj.bias")) * mask
    C = stats.shape[1] // 2
    return x, stats[:, :C], stats[:, C:], mask


# ----------------------------------------------------------------------------
# decoder body: models.py:278-293 / 344-365 / 430-453, modules.py:213-228, 251-262
# ----------------------------------------------------------------------------
def resblock(x, sd, prefix, kind, k, dils, g=None):
    if g is not None and (prefix + ".cond.weight") in sd:
        x = x + F.conv1d(g, sd[prefix + ".cond.weight"].float(), sd[prefix + ".cond.bias"].float())
    if kind == "1":  # modules.py:216-225
        for p, d in enumerate(dils):
            xt = F.leaky_relu(x, LRELU_SLOPE)
            xt = _conv_same(xt, sd, f"{prefix}.convs1.{p}", dilation=d)
            xt = F.leaky_relu(xt, LRELU_SLOPE)
            xt = _conv_same(xt, sd, f"{prefix}.convs2.{p}")
            x = xt + x
    else:  # modules.py:254-259
        for p, d in enumerate(dils):
            xt = F.leaky_relu(x, LRELU_SLOPE)
            xt = _conv_same(xt, sd, f"{prefix}.convs.{p}", dilation=d)
            x = xt + x
    return x


def decoder_logits(sd, cfg, z, g=None, prefix="dec"):
    """conv_pre .. conv_post: returns the pre-head logits [B, S*(n_fft+2), F] (F = L+1)."""
    x = _conv_same(z.float(), sd, prefix + ".conv_pre")
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, effective_weight(sd, f"{prefix}.ups.{i}"), _bias(sd, f"{prefix}.ups.{i}"),
                               stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            y = resblock(x, sd, f"{prefix}.resblocks.{i * nk + j}", cfg["resblock"],
                         cfg["resblock_kernel_sizes"][j], cfg["resblock_dilation_sizes"][j], g=g)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01 (models.py:291/363/451)
    x = F.pad(x, (1, 0), mode="reflect")  # ReflectionPad1d((1,0))
    post = ".conv_post" if cfg["variant"] == "istft" else ".subband_conv_post"
    return _conv_same(x, sd, prefix + post)


# ----------------------------------------------------------------------------
# head + iSTFT + sub-band synthesis: models.py:294-297 / 366-377 / 454-467
# ----------------------------------------------------------------------------
def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True) (stft.py:187), float32."""
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)


def istft_torch(mag, phase, n_fft, hop):
    """TorchSTFT.inverse (stft.py:197-202): torch.istft(mag * exp(j*phase))."""
    win = torch.from_numpy(hann_periodic(n_fft)).to(mag.device)
    return torch.istft(mag * torch.exp(phase * 1j), n_fft, hop, n_fft, window=win)


def istft_closed_form(mag, phase, n_fft, hop):
    """Same transform written out: real inverse DFT per frame (imaginary part of the
    DC and Nyquist bins ignored), periodic Hann, overlap-add, trim n_fft/2, divide by the
    window-square envelope.  Used to validate the in-register DFT of the CUDA tail kernel."""
    nb, nbins, nfr = mag.shape
    assert nbins == n_fft // 2 + 1
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(nbins, dtype=torch.float64)
    ck = torch.full((nbins,), 2.0, dtype=torch.float64)
    ck[0] = 1.0
    ck[-1] = 1.0
    ang = 2.0 * math.pi * k[:, None] * n[None, :] / n_fft
    cosb = (ck[:, None] * torch.cos(ang) / n_fft)
    sinb = (ck[:, None] * torch.sin(ang) / n_fft)
    sinb[0] = 0.0
    sinb[-1] = 0.0
    re = (mag * torch.cos(phase)).double()
    im = (mag * torch.sin(phase)).double()
    fr = torch.einsum("bkf,kn->bfn", re, cosb) - torch.einsum("bkf,kn->bfn", im, sinb)
    w = torch.from_numpy(hann_periodic(n_fft)).double()
    fr = fr * w
    total = n_fft + hop * (nfr - 1)
    ola = torch.zeros(nb, total, dtype=torch.float64)
    env = torch.zeros(total, dtype=torch.float64)
    for f in range(nfr):
        ola[:, f * hop: f * hop + n_fft] += fr[:, f]
        env[f * hop: f * hop + n_fft] += w * w
    half = n_fft // 2
    y = ola[:, half: total - half] / env[half: total - half]
    return y.float()


def pqmf_synthesis_filter(subbands=4, taps=62, cutoff_ratio=0.15, beta=9.0):
    """pqmf.design_prototype_filter + the h_synthesis rows (pqmf.py:15-43, 64-79), float64."""
    n = np.arange(taps + 1)
    omega_c = np.pi * cutoff_ratio
    with np.errstate(invalid="ignore", divide="ignore"):
        h_i = np.sin(omega_c * (n - 0.5 * taps)) / (np.pi * (n - 0.5 * taps))
    h_i[taps // 2] = cutoff_ratio
    from scipy.signal.windows import kaiser
    h = h_i * kaiser(taps + 1, beta)
    hs = np.zeros((subbands, taps + 1))
    for k in range(subbands):
        hs[k] = 2 * h * np.cos((2 * k + 1) * (np.pi / (2 * subbands)) * (n - (taps - 1) / 2)
                               - (-1) ** k * np.pi / 4)
    return hs


def zero_stuff(y_mb, subbands):
    """F.conv_transpose1d(x, updown_filter * subbands, stride=subbands) (pqmf.py:115, models.py:463)."""
    filt = torch.zeros(subbands, subbands, subbands, device=y_mb.device)
    for k in range(subbands):
        filt[k, k, 0] = 1.0
    return F.conv_transpose1d(y_mb, filt * subbands, stride=subbands)


def pqmf_synthesis(y_mb, subbands=4):
    """PQMF.synthesis (pqmf.py:105-116)."""
    hs = torch.from_numpy(pqmf_synthesis_filter(subbands)).float().unsqueeze(0).to(y_mb.device)  # [1,S,63]
    up = zero_stuff(y_mb, subbands)
    return F.conv1d(F.pad(up, (31, 31)), hs)


def decoder_tail(sd, cfg, logits, prefix="dec", closed_form=False):
    """exp / pi*sin head, iSTFT, sub-band synthesis.  Returns (o, o_mb, spec, phase)."""
    n_fft, hop = cfg["gen_istft_n_fft"], cfg["gen_istft_hop_size"]
    nb = n_fft // 2 + 1
    istft = istft_closed_form if closed_form else istft_torch
    B, _, Fr = logits.shape
    if cfg["variant"] == "istft":
        spec = torch.exp(logits[:, :nb])
        phase = math.pi * torch.sin(logits[:, nb:])
        out = istft(spec, phase, n_fft, hop).unsqueeze(-2)
        return out, None, spec, phase
    S = cfg["subbands"]
    x = logits.reshape(B, S, 2 * nb, Fr)
    spec = torch.exp(x[:, :, :nb])
    phase = math.pi * torch.sin(x[:, :, nb:])
    y_mb = istft(spec.reshape(B * S, nb, Fr), phase.reshape(B * S, nb, Fr), n_fft, hop)
    y_mb = y_mb.reshape(B, S, -1)
    if cfg["variant"] == "mb":
        return pqmf_synthesis(y_mb, S), y_mb, spec, phase
    up = zero_stuff(y_mb, S)  # models.py:463
    w = effective_weight(sd, prefix + ".multistream_conv_post")  # [1,4,63], no bias (models.py:425)
    return F.conv1d(up, w, None, padding=31), up, spec, phase


def decode(sd, cfg, z, g=None, prefix="dec", closed_form=False):
    """{iSTFT,Multiband_iSTFT,Multistream_iSTFT}_Generator.forward (models.py:278-297/344-377/430-467)."""
    with torch.no_grad():
        return decoder_tail(sd, cfg, decoder_logits(sd, cfg, z, g, prefix), prefix, closed_form)


def flow_decode(sd, cfg, z_p, y_mask, g=None):
    """The tail of SynthesizerTrn.infer (models.py:730-734): z = flow(z_p, y_mask, g, reverse=True);
    dec(z * y_mask, g)."""
    with torch.no_grad():
        z = flow_reverse(sd, cfg, z_p, y_mask, g)
        return z, decode(sd, cfg, z * y_mask, g)


# ----------------------------------------------------------------------------
# next-row widening: waveform post-processing of the TTS service (tts_vits.py:204-216), numpy like the reference
# ----------------------------------------------------------------------------
def expand_prior(m_p, logs_p, w_ceil, noise, noise_scale=1.0, x_mask=None):
    """NEXT-row widening: the alignment expansion and prior sampling of SynthesizerTrn.infer,
    models.py:717-729 with commons.generate_path (commons.py:128-143) and commons.sequence_mask (commons.py:120-125):

        y_lengths = clamp_min(sum(w_ceil), 1);  y_mask = (arange(Ty) < y_lengths)
        attn[b, ty, tx] = 1  iff  cum[tx-1] <= ty < cum[tx]   (cum = cumsum(w_ceil)), times x_mask[tx] * y_mask[ty]
        m, logs = attn @ m_p, attn @ logs_p                    -- a row gather, since attn has at most one 1 per row
        z_p = m + noise * exp(logs) * noise_scale              -- noise = the reference's torch.randn_like(m_p)

    m_p, logs_p: [B, C, Tx]; w_ceil: [B, 1, Tx] (ceil'd, already masked durations); noise: [B, C, Ty] with
    Ty = max(y_lengths).  Written with index arithmetic (no attention matrix product) so that it is an independent
    restatement.  Returns (z_p, y_mask [B,1,Ty], attn [B,1,Ty,Tx], m [B,C,Ty], logs [B,C,Ty], y_lengths)."""
    B, C, Tx = m_p.shape
    cum = torch.cumsum(w_ceil[:, 0, :].to(torch.float32), dim=-1)                 # [B, Tx], integer-valued
    y_lengths = torch.clamp_min(cum[:, -1], 1).long()
    Ty = int(y_lengths.max())
    assert noise.shape == (B, C, Ty), (noise.shape, (B, C, Ty))
    ty = torch.arange(Ty, dtype=torch.float32)
    y_mask = (torch.arange(Ty)[None, :] < y_lengths[:, None]).to(m_p.dtype).unsqueeze(1)
    attn = torch.zeros((B, 1, Ty, Tx), dtype=m_p.dtype)
    m = torch.zeros((B, C, Ty), dtype=m_p.dtype)
    logs = torch.zeros((B, C, Ty), dtype=m_p.dtype)
    for b in range(B):
        # first token whose cumulative duration exceeds ty
        tx = torch.searchsorted(cum[b].contiguous(), ty, right=True)              # [Ty], == Tx when none does
        ok = (tx < Tx) & (torch.arange(Ty) < y_lengths[b])
        txc = tx.clamp(max=Tx - 1)
        if x_mask is not None:
            ok = ok & (x_mask[b, 0, txc] != 0)
        rows = torch.nonzero(ok)[:, 0]
        attn[b, 0, rows, txc[rows]] = 1.0
        m[b][:, rows] = m_p[b][:, txc[rows]]
        logs[b][:, rows] = logs_p[b][:, txc[rows]]
    z_p = m + noise * torch.exp(logs) * noise_scale
    return z_p, y_mask, attn, m, logs, y_lengths


def pcm16(audio, auto_normalize=True):
    """One utterance, float32 numpy array -> int16.  Steps 3-5 of tts_vits.py:204-216 verbatim in numpy."""
    audio = np.asarray(audio, dtype=np.float32)
    max_abs_val = np.abs(audio).max()
    if auto_normalize and max_abs_val > 0.01:
        audio = (audio / max_abs_val) * 0.9
    audio = np.clip(audio, -1.0, 1.0)
    return (audio * 32767).astype(np.int16)


# ----------------------------------------------------------------------------
# comparators used by the parity tests (tolerances from BASELINE.json north_star)
# ----------------------------------------------------------------------------
def max_abs_over_peak(test, ref):
    ref = ref.double()
    return float((test.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def snr_db(test, ref):
    ref = ref.double()
    err = (test.double() - ref).pow(2).sum().clamp_min(1e-300)
    return float(10.0 * torch.log10(ref.pow(2).sum() / err))
