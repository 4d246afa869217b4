"""Host-to-host serving loop: overlaps the host<->device copies of neighbouring batches with the compute of the
current one.

The reference's service (tts_vits.py:150-226) synthesises one request at a time: copy in, run, copy out.  On a B200
the hot path takes ~9 ms for 64 x 10 s, while its 42 MB of latents in and 56 MB of waveform out take 2-3 ms over
PCIe when they are serialised with it.  `HostStream` keeps three CUDA streams (copy-in, compute, copy-out) and a ring
of `depth` slots, so batch i+1 is uploaded and batch i-1 downloaded while batch i runs.  Per batch it does exactly
what a caller of the drop-in modules does: `z = flow(z_p, y_mask, g, reverse=True); o = dec(z * y_mask, g)[0]`
(models.py:730-734), with pinned host tensors on both ends.
"""
from __future__ import annotations

import torch

from .modules import NativeDecoder, NativeFlow


class HostStream:
    def __init__(self, engine, depth: int = 2, fused: bool = False, graphs: bool = True):
        """fused=False: the two drop-in module calls with `z * y_mask` between them, exactly as infer() makes them.
        fused=True: one `Engine.flow_decode` call (mbv_flow_decode: the same arithmetic without the round trip of z
        through its fp32 NCT boundary layout)."""
        self.engine = engine
        self.fused = fused
        self.graphs = graphs and fused  # fused mode only: every slot replays one captured CUDA graph (static buffers)
        self.dev = engine.device
        self.flow = NativeFlow(engine)
        self.dec = NativeDecoder(engine, want_mb=False, want_spec=False)
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_cmp = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self.depth = depth
        self.slots = [None] * depth
        self.n = 0

    def _slot(self, i, zp_host, mask_host, g_host):
        s = self.slots[i % self.depth]
        B, T = zp_host.shape[0], zp_host.shape[2]
        key = (tuple(zp_host.shape), tuple(mask_host.shape), None if g_host is None else tuple(g_host.shape))
        # a captured graph bakes in the engine's workspace pointer: the engine reallocates (after synchronising the
        # device) when this batch does not fit, and bumps ws_generation
        gen = self.engine.reserve_workspace(B, T)
        if s is not None and (s["key"] != key or (s["graph"] is not None and s["gen"] != gen)):
            # The slot's buffers are about to be dropped: copies / kernels of batches already submitted may still read
            # or write them, and the caching allocator would hand the blocks out again at once (to the new slot below).
            self.drain()
            s = None
        if s is None:
            with torch.cuda.stream(self.s_cmp):
                s = {"zp": torch.empty(zp_host.shape, dtype=torch.float32, device=self.dev),
                     "m": torch.empty(mask_host.shape, dtype=torch.float32, device=self.dev),
                     "g": None if g_host is None else torch.empty(g_host.shape, dtype=torch.float32, device=self.dev),
                     "wav": torch.empty((B, 1, self.engine.spf * T), dtype=torch.float32, device=self.dev),
                     "ev_in": torch.cuda.Event(), "ev_cmp": torch.cuda.Event(), "ev_out": torch.cuda.Event()}
            for t in (s["zp"], s["m"], s["g"], s["wav"]):
                if t is not None:   # used on all three streams for the slot's whole life
                    t.record_stream(self.s_in)
                    t.record_stream(self.s_out)
            s["graph"] = None
            s["key"] = key
            s["gen"] = gen
            if self.graphs:
                # the slot's buffers are static, so its whole launch sequence is captured once and replayed per batch
                with torch.cuda.stream(self.s_cmp):
                    s["zp"].zero_(); s["m"].fill_(1.0)
                    if s["g"] is not None:
                        s["g"].zero_()
                    self.engine.flow_decode(s["zp"], s["m"], s["g"], want_z=False, out_wav=s["wav"])  # tensor maps, attributes
                self.s_cmp.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=self.s_cmp):
                    self.engine.flow_decode(s["zp"], s["m"], s["g"], want_z=False, out_wav=s["wav"])
                s["graph"] = gr
            s["ev_cmp"].record(self.s_cmp)
            s["ev_out"].record(self.s_out)
            self.slots[i % self.depth] = s
        return s

    @torch.no_grad()
    def submit(self, zp_host, mask_host, wav_host, g_host=None):
        """Enqueue one batch: pinned z_p [B,C,T], y_mask [B,1,T] (and g [B,gin,1]) in, pinned wav [B,1,256T] out.
        Returns the event that fires when wav_host is complete."""
        s = self._slot(self.n, zp_host, mask_host, g_host)
        self.n += 1
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(s["ev_cmp"])  # the compute that last read this slot's inputs is done
            s["zp"].copy_(zp_host, non_blocking=True)
            s["m"].copy_(mask_host, non_blocking=True)
            if g_host is not None:
                s["g"].copy_(g_host, non_blocking=True)
            s["ev_in"].record(self.s_in)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(s["ev_in"])
            self.s_cmp.wait_event(s["ev_out"])  # this slot's waveform buffer has been downloaded
            if s["graph"] is not None:
                s["graph"].replay()
                o = s["wav"]
            elif self.fused:  # writes into the slot's preallocated buffer: no allocation per step
                o = self.engine.flow_decode(s["zp"], s["m"], s["g"], want_z=False, out_wav=s["wav"])[1]
            else:
                z = self.flow(s["zp"], s["m"], g=s["g"], reverse=True)
                o = self.dec(z * s["m"], g=s["g"])[0]
                o.record_stream(self.s_out)
            s["ev_cmp"].record(self.s_cmp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(s["ev_cmp"])
            wav_host.copy_(o, non_blocking=True)
            s["ev_out"].record(self.s_out)
        return s["ev_out"]

    def drain(self):
        for st in (self.s_in, self.s_cmp, self.s_out):
            st.synchronize()
