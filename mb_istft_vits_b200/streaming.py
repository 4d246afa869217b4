"""Streaming front of the decoder: exact chunked decoding (mbv_stream_*) plus the PCM chunker of the reference's TTS
service (tts_vits.py:204-226: clip, x32767, int16, 20 ms slices).

The reference synthesises a whole utterance and only then slices the PCM; its notebooks decode 10-frame latent chunks
independently and cross-fade them (infer.ipynb cells 4-6), which is approximate because the decoder's receptive field
(+-24 latent frames) is ignored.  Here the library keeps the halo frames in device state, so the streamed waveform equals
the one-shot ``dec(z)`` bit for bit with a latency of ``halo`` latent frames.

Peak normalisation (tts_vits.py:206-210) needs the whole utterance and therefore cannot be streamed; ``push_pcm`` applies
the reference's ``auto_normalize = False`` arithmetic (or a gain the caller already knows).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch


class StreamingDecoder:
    def __init__(self, engine, batch: int, max_chunk_frames: int, frame_length: float = 0.02, rate: Optional[int] = None):
        """frame_length / rate: the PCM slice length of tts_vits.py:35-38 (chunk_size = round(frame_length * rate))."""
        self.engine = engine
        self.B = batch
        self.max_chunk = max_chunk_frames
        self._s = C.c_void_p()
        engine._check(engine.lib.mbv_stream_open(engine._h, batch, max_chunk_frames, C.byref(self._s)))
        n = C.c_size_t()
        engine._check(engine.lib.mbv_stream_workspace_bytes(self._s, C.byref(n)))
        self._ws_bytes = n.value
        self.halo = int(engine.lib.mbv_stream_halo(self._s))
        self.rate = rate or engine.cfg["sampling_rate"]
        self.chunk_size = round(frame_length * self.rate)
        self._pcm_rest = [torch.empty(0, dtype=torch.int16) for _ in range(batch)]
        self._out = torch.empty((batch, 1, engine.spf * (max_chunk_frames + self.halo)), dtype=torch.float32, device=engine.device)

    def close(self):
        if self._s.value:
            self.engine.lib.mbv_stream_close(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @torch.no_grad()
    def push(self, z_chunk: Optional[torch.Tensor], g=None, last: bool = False) -> Tuple[int, torch.Tensor]:
        """z_chunk: [B, inter, n] (n <= max_chunk_frames; None or n = 0 to flush with last=True), already multiplied by
        the mask like the decoder's input.  Returns (first latent frame, wav [B, 1, 256 * frames]) of what became final."""
        eng = self.engine
        n = 0 if z_chunk is None else int(z_chunk.shape[2])
        z = None if n == 0 else eng._prep(z_chunk, (self.B, eng.cfg["inter_channels"], n))
        g = eng._prep(g)
        ws, nws = eng._workspace(self.B, self.max_chunk, need=max(self._ws_bytes, eng.workspace_capacity()))
        first, cnt = C.c_int64(), C.c_int32()
        cap = self.max_chunk + self.halo
        eng._check(eng.lib.mbv_stream_push(self._s, eng._ptr(z), n, 1 if last else 0, eng._ptr(g), eng._ptr(self._out), cap,
                                           C.byref(first), C.byref(cnt), C.c_void_p(ws), nws, eng._stream()))
        k = cnt.value * eng.spf
        # the library wrote a dense [B, 1, k] block at the start of the buffer
        wav = self._out.view(-1)[: self.B * k].view(self.B, 1, k).clone()
        return int(first.value), wav

    @torch.no_grad()
    def push_pcm(self, z_chunk, g=None, last: bool = False, gain: float = 1.0) -> List[List[torch.Tensor]]:
        """Like ``push`` but returns, per utterance, the list of complete ``chunk_size``-sample int16 slices that are now
        due (tts_vits.py:218-226); the remainder is carried to the next call and flushed (short) when ``last``."""
        _, wav = self.push(z_chunk, g, last)
        out: List[List[torch.Tensor]] = [[] for _ in range(self.B)]
        if wav.shape[-1]:
            if gain != 1.0:
                wav = wav * gain
            pcm = self.engine.pcm16(wav, None, auto_normalize=False).cpu()
        else:
            pcm = torch.empty((self.B, 0), dtype=torch.int16)
        for b in range(self.B):
            buf = torch.cat([self._pcm_rest[b], pcm[b]])
            t = 0
            while t + self.chunk_size <= buf.numel():
                out[b].append(buf[t:t + self.chunk_size])
                t += self.chunk_size
            if last and t < buf.numel():
                out[b].append(buf[t:])
                t = buf.numel()
            self._pcm_rest[b] = buf[t:]
        return out
