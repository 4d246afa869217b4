"""Host-side owner of one libmbistft handle: weight folding / upload, workspace, and the calls.

PyTorch is used for what it is good at here -- device memory, streams, tensors at the boundary.
All arithmetic of the hot path happens in libmbistft.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import lib as _lib
from .configs import receptive_field_frames, samples_per_frame


def fold_weight_norm(sd: Dict[str, torch.Tensor], prefixes=("dec.", "flow.", "enc_q.", "enc_p.")) -> Dict[str, torch.Tensor]:
    """Reference checkpoint layout -> effective fp32 weights: w = g * v / ||v|| with the norm
    over all dims but 0 (torch.nn.utils.weight_norm, dim=0; for ConvTranspose1d dim 0 = in-channels).
    Keys already stored as plain ``weight`` (after remove_weight_norm) pass through."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if not k.startswith(prefixes):
            continue
        if k.endswith(".weight_v"):
            base = k[: -len(".weight_v")]
            g = sd[base + ".weight_g"].detach().float().cpu()
            vv = v.detach().float().cpu()
            norm = vv.reshape(vv.shape[0], -1).norm(dim=1).reshape(g.shape)
            out[base + ".weight"] = (vv * (g / norm)).contiguous()
        elif k.endswith(".weight_g") or k.endswith("updown_filter"):
            continue
        else:
            out[k] = v.detach().float().cpu().contiguous()
    return out


class CapturedStep:
    """A CUDA graph of one Engine call.  The graph bakes in the engine's workspace pointer, so ``replay`` refuses to run
    after the engine has reallocated its workspace (``Engine.ws_generation`` changed)."""

    def __init__(self, engine, graph):
        self.engine, self.graph, self.gen = engine, graph, engine.ws_generation

    def replay(self):
        if self.engine.ws_generation != self.gen:
            raise RuntimeError("the engine reallocated its workspace after this graph was captured; capture it again")
        self.graph.replay()


class Engine:
    """One handle on one GPU.  Not re-entrant across streams that share its workspace."""

    def __init__(self, cfg, state_dict, precision="bf16", device=0, flags=0, residual=None):
        """precision: 'bf16' | 'tf32' | 'fp32' (arithmetic of the dense contractions).
        residual: storage type of the decoder's ResBlock residual stream, 'fp32' or 'fp16' (bf16 precision only;
        default 'fp16' there: one 2^-11 rounding per residual add is invisible next to the 2^-9 operand rounding --
        measured 45-47 dB either way -- and it removes a third of the ResBlock traffic)."""
        self.cfg = dict(cfg)
        self.precision = precision
        if residual is None:
            residual = "fp16" if precision in ("bf16", "fp16") else "fp32"
        if residual not in ("fp16", "fp32"):
            raise ValueError("residual must be 'fp16' or 'fp32'")
        self.residual = residual
        if precision == "fp16" and residual != "fp16":
            raise ValueError("precision 'fp16' keeps its residual streams in the fp16 operand tensors")
        if residual == "fp16" and precision == "bf16":
            flags |= _lib.FLAG_RESIDUAL_FP16
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("mb_istft_vits_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        self.spf = samples_per_frame(cfg)
        self._h = C.c_void_p()
        ccfg = _lib.make_config(cfg, precision, device, flags)
        rc = self.lib.mbv_create(C.byref(ccfg), C.byref(self._h))
        self._check(rc)
        self._load(state_dict)
        self._ws = None
        self.ws_generation = 0   # bumped whenever the workspace is reallocated: captured graphs bake its pointer in

    # ---- plumbing
    def _check(self, rc):
        if rc != 0:
            msg = self.lib.mbv_last_error(self._h)
            raise _lib.MbvError(rc, msg.decode() if msg else "")

    def _load(self, sd):
        eff = fold_weight_norm(sd)
        arr = (_lib.MbvTensor * len(eff))()
        keep = []
        for i, (k, t) in enumerate(eff.items()):
            t = t.contiguous()
            keep.append(t)
            arr[i].name = k.encode()
            arr[i].data = C.cast(t.data_ptr(), C.POINTER(C.c_float))
            arr[i].rank = t.dim()
            for r, s in enumerate(t.shape):
                arr[i].shape[r] = s
        self._check(self.lib.mbv_load_weights(self._h, arr, len(eff)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.mbv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def workspace_bytes(self, B, T):
        n = C.c_size_t()
        self._check(self.lib.mbv_workspace_bytes(self._h, B, T, C.byref(n)))
        return n.value

    def _workspace(self, B, T, need=None):
        need = self.workspace_bytes(B, T) if need is None else need
        if self._ws is None or self._ws.numel() < need + 1024:
            if self._ws is not None:
                # kernels of earlier calls (on any stream) may still be running in the old block, and the caching
                # allocator would hand it out again at once: wait for the device before dropping it
                torch.cuda.synchronize(self.device)
            self._ws = None
            with torch.cuda.device(self.device):
                self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            self.ws_generation += 1
        base = self._ws.data_ptr()
        off = (-base) % 1024
        return base + off, self._ws.numel() - off

    def workspace_capacity(self) -> int:
        """Bytes of the workspace currently held (0 before the first call)."""
        return 0 if self._ws is None else int(self._ws.numel()) - 1024

    def reserve_workspace(self, B, T) -> int:
        """Make sure the workspace fits a [B, T] batch (reallocating, after a device synchronise, if it does not) and
        return ``ws_generation``; holders of captured CUDA graphs compare it to detect a reallocation."""
        self._workspace(B, T)
        return self.ws_generation

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _prep(self, t, shape=None):
        if t is None:
            return None
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _alloc_outputs(self, B, T, want_mb, want_spec, wav=None):
        v = self.cfg["variant"]
        S = self.cfg["subbands"] if v != "istft" else 1
        L = T
        for u in self.cfg["upsample_rates"]:
            L *= u
        if wav is None:
            wav = torch.empty((B, 1, self.spf * T), dtype=torch.float32, device=self.device)
        elif (tuple(wav.shape) != (B, 1, self.spf * T) or wav.dtype != torch.float32 or not wav.is_contiguous()
              or wav.device != self.device):
            raise ValueError("out_wav must be a contiguous fp32 [B, 1, samples] tensor on the engine's device")
        o_mb = None
        if want_mb and v != "istft":
            o_mb = torch.empty((B, S, (4 if v == "mb" else 16) * L), dtype=torch.float32, device=self.device)
        spec = phase = None
        if want_spec:
            shape = (B, S, 9, L + 1) if v != "istft" else (B, 9, L + 1)
            spec = torch.empty(shape, dtype=torch.float32, device=self.device)
            phase = torch.empty(shape, dtype=torch.float32, device=self.device)
        return wav, o_mb, spec, phase

    # ---- the three entry points
    def flow_reverse(self, z_p, y_mask, g=None):
        B, Cz, T = z_p.shape
        z_p = self._prep(z_p)
        y_mask = self._prep(y_mask, (B, 1, T))
        g = self._prep(g)
        out = torch.empty_like(z_p)
        ws, nws = self._workspace(B, T)
        self._check(self.lib.mbv_flow_reverse(self._h, self._ptr(z_p), self._ptr(y_mask), self._ptr(g), self._ptr(out),
                                              B, T, C.c_void_p(ws), nws, self._stream()))
        return out

    def flow_forward(self, x, y_mask, g=None):
        """ResidualCouplingBlock.forward(reverse=False) (models.py:207-210): the direction voice conversion uses."""
        B, Cz, T = x.shape
        x = self._prep(x)
        y_mask = self._prep(y_mask, (B, 1, T))
        g = self._prep(g)
        out = torch.empty_like(x)
        ws, nws = self._workspace(B, T)
        self._check(self.lib.mbv_flow_forward(self._h, self._ptr(x), self._ptr(y_mask), self._ptr(g), self._ptr(out),
                                              B, T, C.c_void_p(ws), nws, self._stream()))
        return out

    def text_encode(self, tokens, x_lengths):
        """TextEncoder.forward (models.py:172-181) on the library: tokens [B, Tx] int64, lengths [B] ->
        (x [B,H,Tx], m_p [B,inter,Tx], logs_p [B,inter,Tx], x_mask [B,1,Tx]).  Needs the enc_p.* weights."""
        tokens = torch.as_tensor(tokens).to(device=self.device, dtype=torch.int64).contiguous()
        B, Tx = tokens.shape
        lens = torch.as_tensor(x_lengths).to(self.device)
        x_mask = (torch.arange(Tx, device=self.device)[None, :] < lens[:, None]).to(torch.float32).unsqueeze(1)  # commons.sequence_mask
        inter, hid = self.cfg["inter_channels"], self.cfg["hidden_channels"]
        x = torch.empty((B, hid, Tx), dtype=torch.float32, device=self.device)
        stats = torch.empty((B, 2 * inter, Tx), dtype=torch.float32, device=self.device)
        n = C.c_size_t()
        self._check(self.lib.mbv_text_workspace_bytes(self._h, B, Tx, C.byref(n)))
        ws, nws = self._workspace(B, Tx, need=max(n.value, self.workspace_capacity()))
        self._check(self.lib.mbv_text_encode(self._h, C.c_void_p(tokens.data_ptr()), self._ptr(x_mask), self._ptr(x), self._ptr(stats),
                                             B, Tx, C.c_void_p(ws), nws, self._stream()))
        return x, stats[:, :inter], stats[:, inter:], x_mask

    def posterior_encode(self, spec, y_lengths=None, g=None, noise=None, y_mask=None):
        """PosteriorEncoder.forward (models.py:236-246) on the library: spec [B, spec_channels, T], lengths [B] (or a ready
        y_mask [B,1,T]) -> (z, m, logs, y_mask).  noise [B,inter,T] or None = drawn here with torch.randn, which is what
        the reference's ``torch.randn_like(m)`` yields for its (contiguous) result.  Needs the enc_q.* weights."""
        spec = self._prep(spec)
        B, _, T = spec.shape
        if y_mask is None:
            lens = torch.as_tensor(y_lengths).to(self.device)
            y_mask = (torch.arange(T, device=self.device)[None, :] < lens[:, None]).to(torch.float32).unsqueeze(1)  # commons.sequence_mask
        y_mask = self._prep(y_mask, (B, 1, T))
        g = self._prep(g)
        inter = self.cfg["inter_channels"]
        if noise is None:
            noise = torch.randn((B, inter, T), dtype=torch.float32, device=self.device)
        noise = self._prep(noise, (B, inter, T))
        z = torch.empty((B, inter, T), dtype=torch.float32, device=self.device)
        stats = torch.empty((B, 2 * inter, T), dtype=torch.float32, device=self.device)
        n = C.c_size_t()
        self._check(self.lib.mbv_posterior_workspace_bytes(self._h, B, T, C.byref(n)))
        ws, nws = self._workspace(B, T, need=max(n.value, self.workspace_capacity()))
        self._check(self.lib.mbv_posterior_encode(self._h, self._ptr(spec), self._ptr(y_mask), self._ptr(g), self._ptr(noise),
                                                  self._ptr(z), self._ptr(stats), B, T, C.c_void_p(ws), nws, self._stream()))
        return z, stats[:, :inter], stats[:, inter:], y_mask

    def decode(self, z, g=None, z_mask=None, want_mb=True, want_spec=True):
        B, Cz, T = z.shape
        z = self._prep(z)
        g = self._prep(g)
        z_mask = self._prep(z_mask, (B, 1, T)) if z_mask is not None else None
        wav, o_mb, spec, phase = self._alloc_outputs(B, T, want_mb, want_spec)
        ws, nws = self._workspace(B, T)
        self._check(self.lib.mbv_decode(self._h, self._ptr(z), self._ptr(z_mask), self._ptr(g), self._ptr(wav),
                                        self._ptr(o_mb), self._ptr(spec), self._ptr(phase), B, T, C.c_void_p(ws), nws,
                                        self._stream()))
        return wav, o_mb, spec, phase

    def flow_decode(self, z_p, y_mask, g=None, want_z=True, want_mb=False, want_spec=False, out_wav=None):
        """out_wav: optional preallocated [B,1,samples] fp32 device tensor to write the waveform into (serving loops that
        must not allocate per step)."""
        B, Cz, T = z_p.shape
        z_p = self._prep(z_p)
        y_mask = self._prep(y_mask, (B, 1, T))
        g = self._prep(g)
        z = torch.empty_like(z_p) if want_z else None
        wav, o_mb, spec, phase = self._alloc_outputs(B, T, want_mb, want_spec, wav=out_wav)
        ws, nws = self._workspace(B, T)
        self._check(self.lib.mbv_flow_decode(self._h, self._ptr(z_p), self._ptr(y_mask), self._ptr(g), self._ptr(z),
                                             self._ptr(wav), self._ptr(o_mb), self._ptr(spec), self._ptr(phase), B, T,
                                             C.c_void_p(ws), nws, self._stream()))
        return z, wav, o_mb, spec, phase

    def capture_flow_decode(self, z_p, y_mask, g=None, want_z=False, want_mb=False, want_spec=False):
        """Capture one flow_decode() on the given (static) device tensors into a CUDA graph.  Returns
        (graph, outputs): ``graph.replay()`` re-runs the whole launch sequence with one submission; the outputs
        tuple (z, wav, o_mb, spec, phase) is overwritten by every replay.  Every library call is allocation-free,
        host-sync-free and enqueues only on the current stream, so it is capturable as is."""
        B, _, T = z_p.shape
        self._workspace(B, T)  # allocate outside the capture
        self.flow_decode(z_p, y_mask, g, want_z=want_z, want_mb=want_mb, want_spec=want_spec)  # warm-up: tensor maps, attributes
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs = self.flow_decode(z_p, y_mask, g, want_z=want_z, want_mb=want_mb, want_spec=want_spec)
        return CapturedStep(self, graph), outs

    def tail(self, logits, T, want_mb=True, want_spec=True):
        """Fused head+iSTFT+synthesis on logits [B, F, n_ch] (channels-last)."""
        logits = self._prep(logits)
        B = logits.shape[0]
        wav, o_mb, spec, phase = self._alloc_outputs(B, T, want_mb, want_spec)
        self._check(self.lib.mbv_tail(self._h, self._ptr(logits), self._ptr(wav), self._ptr(o_mb), self._ptr(spec),
                                      self._ptr(phase), B, T, self._stream()))
        return wav, o_mb, spec, phase

    def tail_fused(self, act, T, want_mb=False, want_spec=False):
        """The fused conv_post + head + iSTFT + synthesis kernel alone, on the operand tensor conv_post consumes:
        act [B, 16T+1, C] in the engine's 16-bit operand type (bf16 / fp16), channels-last."""
        want = torch.bfloat16 if self.precision == "bf16" else torch.float16
        if act.dtype != want or not act.is_contiguous() or act.device != self.device:
            raise ValueError(f"act must be a contiguous {want} tensor on the engine's device")
        B = act.shape[0]
        wav, o_mb, spec, phase = self._alloc_outputs(B, T, want_mb, want_spec)
        self._check(self.lib.mbv_tail_fused(self._h, C.c_void_p(act.data_ptr()), self._ptr(wav), self._ptr(o_mb), self._ptr(spec),
                                            self._ptr(phase), B, T, self._stream()))
        return wav, o_mb, spec, phase

    # ---- widening beyond the seam
    def pcm16(self, wav, n_samples=None, auto_normalize=True):
        """tts_vits.py:204-216 on the GPU: per-utterance peak normalisation (x0.9 when the peak exceeds 0.01), clip,
        int16.  wav: [B, 1, S] or [B, S] fp32; n_samples: [B] valid lengths or None.  Returns int16 [B, S]."""
        w = self._prep(wav)
        if w.dim() == 3:
            w = w[:, 0, :]
        w = w.contiguous()
        B, S = w.shape
        ns = None if n_samples is None else torch.as_tensor(n_samples).to(device=self.device, dtype=torch.int32).contiguous()
        scratch = torch.empty(B, dtype=torch.int32, device=self.device)
        pcm = torch.empty((B, S), dtype=torch.int16, device=self.device)
        self._check(self.lib.mbv_pcm16(self._h, self._ptr(w), self._ptr(ns), B, S, 1 if auto_normalize else 0,
                                       self._ptr(scratch), self._ptr(pcm), self._stream()))
        return pcm

    def expand_prior(self, m_p, logs_p, w_ceil, noise_scale=1.0, x_mask=None, noise=None, want_attn=False,
                     want_stats=False):
        """models.py:717-729 on the GPU: durations -> y_mask, gathered prior statistics, z_p = m + noise * exp(logs) *
        noise_scale.  m_p, logs_p: [B,C,Tx]; w_ceil: [B,1,Tx] (ceil'd, masked durations).  noise: [B,C,Ty] or None = drawn
        here with torch.randn in the memory order of the reference's randn_like (its expanded m_p has the strides of a
        [B,Ty,C] tensor), so a seeded run consumes the generator exactly like the reference.
        Returns (z_p, y_mask, y_lengths, attn | None, (m_exp, logs_exp) | None)."""
        m_p, logs_p, w_ceil = self._prep(m_p), self._prep(logs_p), self._prep(w_ceil)
        B, Cc, Tx = m_p.shape
        if tuple(logs_p.shape) != (B, Cc, Tx) or tuple(w_ceil.shape) != (B, 1, Tx):
            raise ValueError("expand_prior: m_p / logs_p [B,C,Tx] and w_ceil [B,1,Tx] expected")
        x_mask = self._prep(x_mask, (B, 1, Tx)) if x_mask is not None else None
        y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()   # models.py:719
        Ty = int(y_lengths.max())                                          # host sync, as in commons.sequence_mask
        if noise is None:
            noise = torch.randn((B, Ty, Cc), dtype=torch.float32, device=self.device).transpose(1, 2)
        noise = self._prep(noise, (B, Cc, Ty))
        z_p = torch.empty((B, Cc, Ty), dtype=torch.float32, device=self.device)
        y_mask = torch.empty((B, 1, Ty), dtype=torch.float32, device=self.device)
        attn = torch.empty((B, 1, Ty, Tx), dtype=torch.float32, device=self.device) if want_attn else None
        m_exp = torch.empty_like(z_p) if want_stats else None
        logs_exp = torch.empty_like(z_p) if want_stats else None
        self._check(self.lib.mbv_expand_prior(self._h, self._ptr(m_p), self._ptr(logs_p), self._ptr(w_ceil), self._ptr(x_mask),
                                              self._ptr(noise), float(noise_scale), B, Cc, Tx, Ty, self._ptr(z_p),
                                              self._ptr(y_mask), self._ptr(m_exp), self._ptr(logs_exp), self._ptr(attn), None,
                                              self._stream()))
        return z_p, y_mask, y_lengths, attn, ((m_exp, logs_exp) if want_stats else None)

    def decode_chunked(self, z, g=None, chunk_frames=256, halo_frames=None):
        """Exact streaming decode: the decoder is convolutional with a receptive field of +-24 latent frames (MB/MS;
        +-13 single-band; computed from the geometry by configs.receptive_field_frames), so decoding overlapping windows [a-halo, b+halo) and keeping the samples of
        [a, b) reproduces the one-shot result bit for bit -- unlike the approximate overlap-add chunking of the
        reference notebooks (infer.ipynb cells 4-6).  Yields (first_frame, wav_chunk [B,1,256*(b-a)])."""
        B, _, T = z.shape
        need = receptive_field_frames(self.cfg)
        halo = halo_frames if halo_frames is not None else need + 1
        if halo < need:
            raise ValueError(f"halo_frames={halo} is smaller than this decoder's receptive field ({need} latent frames): "
                             "the chunked result would no longer be exact")
        spf = self.spf
        for a0 in range(0, T, chunk_frames):
            b0 = min(T, a0 + chunk_frames)
            lo, hi = max(0, a0 - halo), min(T, b0 + halo)
            wav = self.decode(z[:, :, lo:hi].contiguous(), g, want_mb=False, want_spec=False)[0]
            yield a0, wav[:, :, (a0 - lo) * spf: (b0 - lo) * spf]

    def set_profiling(self, on: bool):
        self._check(self.lib.mbv_set_profiling(self._h, 1 if on else 0))

    def profile_read(self):
        """-> {'conv': (ms, launches), 'tail': (...), 'other': (...)} accumulated since the last read."""
        ms = (C.c_double * 3)()
        cnt = (C.c_int32 * 3)()
        self._check(self.lib.mbv_profile_read(self._h, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(("conv", "tail", "other"))}

    def profile_read_launches(self, cap=4096):
        """-> [(description, ms)] per launch since the last read, in launch order."""
        ms = (C.c_float * cap)()
        stride = 56
        desc = C.create_string_buffer(cap * stride)
        n = C.c_int32()
        self._check(self.lib.mbv_profile_read_launches(self._h, ms, desc, stride, cap, C.byref(n)))
        raw = desc.raw
        return [(raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(n.value)]

    def last_launch_count(self):
        return int(self.lib.mbv_last_launch_count(self._h))

    def decode_flops(self, B, T):
        return float(self.lib.mbv_decode_flops(self._h, B, T))

    def flow_flops(self, B, T):
        return float(self.lib.mbv_flow_flops(self._h, B, T))
