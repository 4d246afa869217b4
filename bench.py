#!/usr/bin/env python
"""Benchmark of the MB-iSTFT-VITS waveform hot path (flow reverse + iSTFT decoder) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--precision bf16|tf32|fp32]

Workload (BASELINE.json configs[1]): ljs_mb_istft_vits, B = 64 utterances x T = 862 latent frames (10.008 s at
22 050 Hz) per GPU, synthetic latents, seeded random-init weights.  A "step" is one pass of the hot path over
that batch: z = flow(z_p, mask, reverse=True); wav = dec(z * mask)  (models.py:730-734) = 14 123 008 samples.
For N > 1 (launched under torchrun, one rank per GPU) every rank runs its own 64-utterance shard -- utterances are
independent, so there is no collective on the hot path -- and the whole-job value is N * samples / max-over-ranks time.

One JSON line is printed by rank 0:
  value        device-resident throughput (inputs already in HBM), CUDA events around K CUDA-graph replays, max over ranks
  e2e          the same metric through the public API with pinned HOST inputs and a HOST copy of the waveform inside
               the timed region (HostStream serving loop)
  roofline     the conv implicit-GEMM kernels (dominant): algorithmic FLOP/s vs the measured bf16 peak.  Kernel time is
               derived INSIDE the graph-timed region (step time minus the event-timed non-conv launches), so the sum of
               kernel times never exceeds ms_per_step; fractions are given against the sustained AND the burst peak and
               a >= 3 s sustained leg (clocks sampled) is reported next to the short timed region
  roofline_tail  the fused iSTFT/PQMF tail vs the measured HBM copy bandwidth
  cpu_baseline the reference's own modules (baseline/_ref, kind "reference"; the oracle port if the reference is not
               staged) timed on this box's host cores (rank 0, N = 1)
  extras       tf32 path line, PyTorch-eager reference on the same GPU, BASELINE configs 3 and 5 with the NCCL gather

--impl reference times the reference's CPU implementation as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))

CONFIG_NAME = "ljs_mb_istft_vits"
B_PER_GPU = 64
T_FRAMES = 862
CPU_SAMPLE_B = 8   # B = 8 x T = 862 on the host cores (CPU throughput is batch-independent)
METRIC = "decoder audio samples/sec (flow-reverse + MB-iSTFT decoder, batch 64 x 10 s per GPU)"
UNIT = "samples/s"


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML (every ~5 ms) while the timed region runs."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self.th = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
        return self

    def _loop(self):
        nv = self.nv
        names = {}
        for k, label in (("nvmlClocksEventReasonHwSlowdown", "hw_slowdown"),
                         ("nvmlClocksEventReasonHwThermalSlowdown", "hw_thermal_slowdown"),
                         ("nvmlClocksEventReasonSwThermalSlowdown", "sw_thermal_slowdown"),
                         ("nvmlClocksEventReasonSwPowerCap", "sw_power_cap"),
                         ("nvmlClocksThrottleReasonHwSlowdown", "hw_slowdown"),
                         ("nvmlClocksThrottleReasonHwThermalSlowdown", "hw_thermal_slowdown"),
                         ("nvmlClocksThrottleReasonSwThermalSlowdown", "sw_thermal_slowdown"),
                         ("nvmlClocksThrottleReasonSwPowerCap", "sw_power_cap")):
            if hasattr(nv, k):
                names[getattr(nv, k)] = label
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(self.h))
                for bit, label in names.items():
                    if mask & bit:
                        self.reasons.add(label)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(0.005)

    def stop(self):
        self._stop.set()
        if self.th:
            self.th.join(timeout=2)
        sm = sorted(self.sm)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(sm),
               "power_w_max": max(self.power) if self.power else None}
        if self.err:
            out["error"] = self.err
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


def partition_host_cores(rank, world):
    """Give every rank of one box its own slice of the host cores (pinned staging buffers are first-touched and the
    copy / launch threads run there), so that eight ranks do not bounce over the same cores."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cores) >= 2 * world:
            per = len(cores) // world
            os.sched_setaffinity(0, set(cores[rank * per:(rank + 1) * per]))
    except Exception:  # pragma: no cover
        pass


# ------------------------------------------------------------------------------------------------
# reference legs: the unmodified reference modules (baseline/_ref) when staged, else the oracle port
# ------------------------------------------------------------------------------------------------
def make_reference_runner(cfg, sd, device="cpu"):
    """-> (kind, fn(z_p, mask, g) -> waveform) running flow reverse + decoder like models.py:730-734."""
    import torch
    import ref_loader
    if ref_loader.available():
        net = ref_loader.build_synthesizer(cfg, sd, device=device)

        def run(z_p, mask, g=None):
            with torch.no_grad():
                z = net.flow(z_p, mask, g=g, reverse=True)
                return net.dec(z * mask, g=g)[0]
        return "reference", run
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mbistft_oracle as orc
    sd_dev = {k: v.to(device) for k, v in sd.items()}

    def run_port(z_p, mask, g=None):
        with torch.no_grad():
            return orc.flow_decode(sd_dev, cfg, z_p, mask, g)[1][0]
    return "port", run_port


def cpu_reference_run(cfg, sd, B, T, reps, warmup, threads=None):
    """The reference on the host cores: flow reverse + decode, all threads.  Returns (seconds per pass list, cores, kind)."""
    import torch
    from mb_istft_vits_b200 import synth
    n = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(n)
    kind, run = make_reference_runner(cfg, sd, "cpu")
    z_p, mask, _ = synth.make_latents(cfg, B, T, seed=1234)
    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        run(z_p, mask)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, n, kind


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    import contextlib
    from mb_istft_vits_b200 import get_config, synth
    cfg = get_config(CONFIG_NAME)
    sd = synth.make_state_dict(cfg, seed=1234)
    Bs = CPU_SAMPLE_B
    warm = max(1, min(args.warmup, 2))
    with contextlib.redirect_stdout(sys.stderr):  # the reference prints a banner when it builds its decoder
        times, cores, kind = cpu_reference_run(cfg, sd, Bs, T_FRAMES, max(1, args.steps), warm)
    mean = sum(times) / len(times)
    samples = Bs * T_FRAMES * 256
    v = samples / mean
    what = "unmodified reference modules (baseline/_ref: net_g.flow(reverse=True) + net_g.dec)" if kind == "reference" \
        else "oracle port (torch CPU ops of the reference; baseline/_ref not staged)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": warm, "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{CONFIG_NAME} flow-reverse + decoder-from-z, CPU sample B={Bs} x T={T_FRAMES}",
                   "sampling_rate": cfg["sampling_rate"]},
        "rtf": mean / (samples / cfg["sampling_rate"]),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{what}, B={Bs} x T={T_FRAMES} per step, mean of {len(times)} after {warm} warm-up"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _emit(line, fd):
    os.write(fd, (json.dumps(line) + "\n").encode())


# ------------------------------------------------------------------------------------------------
# BASELINE configs 3 and 5 (extras): sharded batches + the final NCCL waveform gather
# ------------------------------------------------------------------------------------------------
def _timed(torch, dist, dev, fn, steps, warmup=1, per_step=None):
    """CUDA-event time of `steps` calls of fn, barrier + synchronize on both sides, max over ranks -> ms per call.
    per_step: a list that receives this rank's event time of every timed call (diagnostic)."""
    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        fn()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    barrier()
    if per_step is not None:
        per_step.extend(round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps))
    t = torch.tensor([ev[0].elapsed_time(ev[steps])], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def run_sharded_config(torch, dist, dev, rank, world, which, steps=5):
    """config 3: ljs_ms_istft_vits, B = 256 utterances x T = 862 in total, strong-scaled (256 / N per GPU).
    config 5: uudb_ms_istft_vits_ms (g-conditioned MS decoder, 16 kHz), 64 utterances per GPU in total with lengths uniform
    in [1, 60] s (seed 1234), length-balanced over the ranks (sharding.balance_utterances), every rank decoding its bin
    as one padded batch.  Both: one flow_decode per rank, then sharding.gather_waveforms to rank 0 over NCCL; timed with
    the gather excluded and included."""
    from mb_istft_vits_b200 import Engine, get_config, synth
    from mb_istft_vits_b200.sharding import balance_utterances, gather_waveforms
    out = {}
    if which == 3:
        cfg = get_config("ljs_ms_istft_vits")
        total = 256
        lengths = [T_FRAMES] * total
        per = total // world
        idx = list(range(rank * per, (rank + 1) * per))
        out["workload"] = f"ljs_ms_istft_vits, B=256 x T={T_FRAMES} in total, {per} utterances per GPU (strong scaling)"
    else:
        cfg = get_config("uudb_ms_istft_vits_ms")
        total = 64 * world
        gen = torch.Generator().manual_seed(1234)
        secs = 1.0 + 59.0 * torch.rand(total, generator=gen)
        lengths = [max(1, int(round(float(s) * cfg["sampling_rate"] / 256))) for s in secs]
        idx = balance_utterances(lengths, world)[rank]
        out["workload"] = (f"uudb_ms_istft_vits_ms (sid-conditioned g), {total} utterances of 1-60 s (seed 1234), "
                           "length-balanced bins, one padded batch per GPU")
    sd = synth.make_state_dict(cfg, seed=1234)
    eng = Engine(cfg, sd, precision="bf16", device=dev.index)
    lens = torch.tensor([lengths[i] for i in idx], dtype=torch.long)
    b, T = len(idx), int(lens.max())
    z_p, mask, _ = synth.make_latents(cfg, b, T, seed=4000 + rank, lengths=lens)
    z_p, mask = z_p.to(dev), mask.to(dev)
    g = None
    if cfg["gin_channels"]:
        sid = torch.tensor([i % cfg["n_speakers"] for i in idx])
        g = sd["emb_g.weight"][sid].unsqueeze(-1).to(dev).contiguous()   # models.py:704-707
    wav = torch.empty((b, 1, 256 * T), dtype=torch.float32, device=dev)
    n_samples = (lens * 256).to(dev)
    state = {}

    def compute():
        eng.flow_decode(z_p, mask, g, want_z=False, out_wav=wav)

    def compute_gather():
        compute()
        state["out"] = gather_waveforms(wav, n_samples, idx, total, dst=0)

    valid = sum(lengths) * 256
    ms_c = _timed(torch, dist, dev, compute, steps)
    out.update({"utterances": total, "valid_samples": valid, "this_rank": {"utterances": b, "T_padded": T},
                "ms_per_step_compute": ms_c, "samples_per_s_compute": valid / (ms_c * 1e-3)})
    if dist is not None:
        # 3 warm-up calls: rank 0 allocates a receive buffer per call while the previous result is still referenced, so the
        # caching allocator reaches its steady state (two alternating blocks, no cudaMalloc) only after two calls
        ps = []
        ms_g = _timed(torch, dist, dev, compute_gather, steps, warmup=3, per_step=ps)
        out["per_step_ms_with_gather_this_rank"] = ps
        out.update({"ms_per_step_with_gather": ms_g, "samples_per_s_with_gather": valid / (ms_g * 1e-3),
                    "gather_ms": ms_g - ms_c, "gather_bytes": valid * 4,
                    "gather": "sharding.gather_waveforms: 1 all_gather of the placement table + grouped NCCL send/recv to rank 0"})
        if rank == 0:
            res = state["out"]
            out["gather_check"] = bool(len(res) == total and all(int(res[i].numel()) == lengths[i] * 256 for i in range(total)))
    else:
        out["gather"] = "n/a at N = 1 (single rank owns every utterance)"
    work = torch.tensor([float(b * T)], dtype=torch.float64, device=dev)
    if dist is not None:
        wmax = work.clone()
        dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
        out["padded_work_balance"] = float(work.item() / world / wmax.item())
    if which == 5:
        # the same utterances, dealt in snake order over the ranks and decoded per rank as <= 4 padded batches of similar
        # length (sharding.decode_in_buckets) written into one flat buffer, which the gather sends as it lies
        from mb_istft_vits_b200.sharding import deal_utterances, decode_in_buckets
        idx_b = deal_utterances(lengths, world)[rank]
        lens_b = [lengths[i] for i in idx_b]
        zb, _, _ = synth.make_latents(cfg, len(idx_b), max(lens_b), seed=4000 + rank, lengths=torch.tensor(lens_b))
        zb = zb.to(dev)
        gb = None
        if cfg["gin_channels"]:
            sid = torch.tensor([i % cfg["n_speakers"] for i in idx_b])
            gb = sd["emb_g.weight"][sid].unsqueeze(-1).to(dev).contiguous()
        flat, offs, plan = decode_in_buckets(eng, zb, lens_b, g=gb)
        ns_b = (torch.tensor(lens_b) * 256).to(dev)

        def compute_b():
            decode_in_buckets(eng, zb, lens_b, g=gb, out=flat, plan=plan)

        def compute_gather_b():
            compute_b()
            state["out_b"] = gather_waveforms(flat, ns_b, idx_b, total, dst=0, offsets=offs)

        ms_cb = _timed(torch, dist, dev, compute_b, steps)
        bk = {"how": "sharding.deal_utterances + sharding.decode_in_buckets (<= 4 length buckets per rank, one flat output buffer)",
              "this_rank": {"utterances": len(idx_b), "buckets": [[bb, TT] for _, TT, _, _, bb in plan[0]],
                            "padded_frames": sum(bb * TT for _, TT, _, _, bb in plan[0]), "valid_frames": sum(lens_b)},
              "ms_per_step_compute": ms_cb, "samples_per_s_compute": valid / (ms_cb * 1e-3)}
        if dist is not None:
            psb = []
            ms_gb = _timed(torch, dist, dev, compute_gather_b, steps, warmup=3, per_step=psb)
            bk["per_step_ms_with_gather_this_rank"] = psb
            bk.update({"ms_per_step_with_gather": ms_gb, "samples_per_s_with_gather": valid / (ms_gb * 1e-3), "gather_ms": ms_gb - ms_cb})
            if rank == 0:
                res = state["out_b"]
                bk["gather_check"] = bool(len(res) == total and all(int(res[i].numel()) == lengths[i] * 256 for i in range(total)))
        out["length_buckets"] = bk
    eng.close()
    del eng
    torch.cuda.empty_cache()
    return out


def run_native(args):
    # Libraries (e.g. NCCL's version banner) print to stdout; the contract is ONE JSON line there.  Everything but that
    # line is routed to stderr at the file-descriptor level.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    rank, local_rank, world = dist_env()
    partition_host_cores(local_rank, world)
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    from mb_istft_vits_b200 import Engine, NativeDecoder, NativeFlow, get_config, synth
    cfg = get_config(CONFIG_NAME)
    sd = synth.make_state_dict(cfg, seed=1234)
    B, T = args.batch, args.frames
    eng = Engine(cfg, sd, precision=args.precision, device=dev.index, residual=args.residual, flags=args.flags)
    z_p_host, mask_host, _ = synth.make_latents(cfg, B, T, seed=1234 + rank)
    z_p, mask = z_p_host.to(dev), mask_host.to(dev)
    samples_per_step = B * T * 256
    sr = cfg["sampling_rate"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return eng.flow_decode(z_p, mask, want_z=False, want_mb=False, want_spec=False)

    # ---------------- device-resident timing: the step's launch sequence is captured once into a CUDA graph
    for _ in range(max(3, args.warmup)):
        step()
    launches_per_step = eng.last_launch_count()
    graph = None
    if not args.no_graph:
        graph, graph_out = eng.capture_flow_decode(z_p, mask)
        for _ in range(2):
            graph.replay()
    run_step = graph.replay if graph is not None else step
    barrier()
    sampler = ClockSampler(dev.index)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t_max = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_step = float(t_max.item()) / args.steps
    value = world * samples_per_step / (ms_step * 1e-3)

    # ---------------- sustained leg: the same replays back to back for >= 3 s (the regime the sustained peak was measured in)
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(3200.0 / ms_step) + 1)
        s_sampler = ClockSampler(dev.index)
        barrier()
        if rank == 0:
            s_sampler.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n_sus):
            run_step()
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sustained = {"steps": n_sus, "seconds": float(ts.item()) * 1e-3, "ms_per_step": float(ts.item()) / n_sus,
                     "clocks": s_sampler.stop() if rank == 0 else None}

    # ---------------- per-launch device times of the NON-conv launches (tail, pack / unpack / gemv): the same step eagerly,
    # every launch bracketed by CUDA events.  (Per-launch events break the programmatic-dependent-launch overlap of the conv
    # launches, so the conv time is NOT taken from here: it is the graph-timed step minus these small launches.)
    eng.set_profiling(True)
    eng.profile_read()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    n_prof = min(args.steps, 5)
    for _ in range(n_prof):
        step()
    p1.record()
    barrier()
    ms_profiled = p0.elapsed_time(p1) / n_prof
    prof = eng.profile_read()
    eng.set_profiling(False)

    # decoder-only (no flow) for the record
    zz = (eng.flow_reverse(z_p, mask) * mask).contiguous()
    for _ in range(2):
        eng.decode(zz, want_mb=False, want_spec=False)
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(args.steps):
        eng.decode(zz, want_mb=False, want_spec=False)
    d1.record()
    barrier()
    dec_ms = d0.elapsed_time(d1) / args.steps
    del zz

    # ---------------- end-to-end through the public API, host buffers in and out.  Every step uploads its own
    # pinned latents + mask, runs the hot path and downloads its waveform into pinned memory.  (a) HostStream: the serving
    # loop, copies of neighbouring steps overlap the compute (three streams, two slots); (b) the drop-in modules (flow,
    # mask, dec: two library calls like infer()) strictly one after the other on one stream.
    from mb_istft_vits_b200 import HostStream
    flow, dec = NativeFlow(eng), NativeDecoder(eng, want_mb=False, want_spec=False)
    n_slots = 2
    zp_pin = [z_p_host.clone().pin_memory() for _ in range(n_slots)]
    mask_pin = [mask_host.clone().pin_memory() for _ in range(n_slots)]
    wav_pin = [torch.empty((B, 1, 256 * T), dtype=torch.float32).pin_memory() for _ in range(n_slots)]

    def e2e_pipelined(n):
        for i in range(n):
            hs.submit(zp_pin[i % n_slots], mask_pin[i % n_slots], wav_pin[i % n_slots])

    def time_pipelined():
        # Same conditions as the device-resident leg above (which starts from an idle GPU): the sustained leg before this one
        # leaves the GPU at its power cap (1.4-1.5 GHz), so idle for 2 s first, then 3 warm-up steps, then the K timed steps.
        hs.drain()
        barrier()
        time.sleep(2.0)
        e2e_pipelined(3)
        hs.drain()
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        x0.record(cur)
        for st in (hs.s_in, hs.s_cmp, hs.s_out):
            st.wait_event(x0)
        w0 = time.perf_counter()
        e2e_pipelined(args.steps)
        for st in (hs.s_in, hs.s_cmp, hs.s_out):
            cur.wait_stream(st)
        x1.record(cur)
        barrier()
        wall = (time.perf_counter() - w0) * 1e3 / args.steps
        t = torch.tensor([x0.elapsed_time(x1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps, wall

    hs = HostStream(eng, depth=n_slots, fused=False)   # module-by-module variant of the serving loop, for the record
    e2e_modules_ms, _ = time_pipelined()
    hs = HostStream(eng, depth=n_slots, fused=True)
    e2e_ms, e2e_wall_ms = time_pipelined()
    e2e_value = world * samples_per_step / (e2e_ms * 1e-3)
    wav_check = float(wav_pin[0].abs().max())  # the downloaded waveform is real data

    def e2e_step():
        zp_d = zp_pin[0].to(dev, non_blocking=True)
        m_d = mask_pin[0].to(dev, non_blocking=True)
        z = flow(zp_d, m_d, g=None, reverse=True)
        o, _, _, _ = dec(z * m_d, g=None)
        wav_pin[0].copy_(o, non_blocking=True)

    for _ in range(3):
        e2e_step()
    barrier()
    y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    y0.record()
    for _ in range(args.steps):
        e2e_step()
    y1.record()
    barrier()
    ty = torch.tensor([y0.elapsed_time(y1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ty, op=dist.ReduceOp.MAX)
    e2e_seq_ms = float(ty.item()) / args.steps
    del hs

    # ---------------- the tail kernels alone, event-timed stand-alone launches: (a) the fused conv_post + head + iSTFT + PQMF
    # kernel the step runs (16-bit paths) on a random operand tensor, (b) the stand-alone tail kernel on fp32 logits
    # (the two-kernel path: conv_post writes 254 MB of logits, this kernel reads them)
    def time_alone(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(n):
            fn()
        t1e.record()
        torch.cuda.synchronize()
        return t0e.elapsed_time(t1e) / n

    L = 16 * T
    logits = torch.randn((B, L + 1, 72), device=dev) * 0.5
    tail_ms = time_alone(lambda: eng.tail(logits, T, want_mb=False, want_spec=False))
    del logits
    fused_tail_ms = None
    if args.precision in ("bf16", "fp16") and not (args.flags & 128):
        act = (torch.randn((B, L + 1, 128), device=dev) * 0.5).to(torch.bfloat16 if args.precision == "bf16" else torch.float16)
        fused_tail_ms = time_alone(lambda: eng.tail_fused(act, T))
        del act

    flops_step = eng.decode_flops(B, T) + eng.flow_flops(B, T)

    # ---------------- the fp32/TF32 precision path of the north star, every run: the same workload through kind::tf32
    # tcgen05 convs (fp32 streams), and the TF32 GEMM peak measured here instead of assumed
    tf32_line = None
    if args.precision == "bf16" and not args.no_tf32:
        try:
            torch.backends.cuda.matmul.allow_tf32 = True
            a = torch.randn((8192, 8192), device=dev)
            b = torch.randn((8192, 8192), device=dev)
            best = None
            for i in range(6):
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                torch.matmul(a, b)
                g1.record()
                torch.cuda.synchronize()
                if i > 0:
                    t = g0.elapsed_time(g1)
                    best = t if best is None else min(best, t)
            tf32_peak = 2 * 8192.0 ** 3 / (best * 1e-3) / 1e12
            torch.backends.cuda.matmul.allow_tf32 = False
            del a, b
            eng32 = Engine(cfg, sd, precision="tf32", device=dev.index)
            for _ in range(2):
                eng32.flow_decode(z_p, mask, want_z=False)
            g32, _ = eng32.capture_flow_decode(z_p, mask)
            g32.replay()
            ms32 = _timed(torch, dist, dev, g32.replay, max(3, args.steps // 4), warmup=1)
            tf32_line = {"ms_per_step": ms32, "value": world * samples_per_step / (ms32 * 1e-3), "unit": UNIT,
                         "dtype": "tf32", "tflops_whole_step": flops_step / (ms32 * 1e-3) / 1e12,
                         "tf32_gemm_peak_tflops": tf32_peak,
                         "tf32_gemm_peak_how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS), best of 5, this run",
                         "frac_of_tf32_peak": flops_step / (ms32 * 1e-3) / 1e12 / tf32_peak,
                         "parity": "<= 1e-3 of peak vs the fp32 reference (tests/test_gpu_parity.py)"}
            del g32
            eng32.close()
            del eng32
            torch.cuda.empty_cache()
        except Exception as e:  # informational leg
            tf32_line = {"error": repr(e)}

    # ---------------- the reference's own modules as PyTorch eager ops ON THIS GPU ("the real bar to beat"):
    # fp32 with TF32 off (the gold setting) and with PyTorch's default TF32 convs.  Bounded sample, rank 0 at N = 1 only.
    torch_eager = None
    if world == 1 and not args.no_cpu:
        try:
            import contextlib
            Bs = min(B, 16)
            with contextlib.redirect_stdout(sys.stderr):
                kind, run_ref = make_reference_runner(cfg, sd, dev)
            zs, ms = z_p[:Bs].contiguous(), mask[:Bs].contiguous()
            what = "unmodified reference modules (baseline/_ref)" if kind == "reference" else "oracle port"
            torch_eager = {"sample": f"{what} on cuda, PyTorch eager, B={Bs} x T={T}, best of 3", "unit": UNIT, "kind": kind}
            for label, tf32 in (("fp32", False), ("tf32", True)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                best = None
                for i in range(4):
                    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    g0.record()
                    with contextlib.redirect_stdout(sys.stderr):
                        run_ref(zs, ms)
                    g1.record()
                    torch.cuda.synchronize()
                    if i > 0:
                        t = g0.elapsed_time(g1)
                        best = t if best is None else min(best, t)
                torch_eager[label] = Bs * T * 256 / (best * 1e-3)
                torch_eager[label + "_ms"] = best
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = False
            torch_eager["native_over_eager_tf32"] = value / torch_eager["tf32"]
            torch_eager["native_over_eager_fp32"] = value / torch_eager["fp32"]
            del run_ref
            torch.cuda.empty_cache()
        except Exception as e:  # the leg is informational: never fail the bench on it
            torch_eager = {"error": repr(e)}

    # ---------------- BASELINE configs 3 and 5 with the final NCCL waveform gather (all ranks take part)
    sharded = {}
    if not args.no_configs:
        for which in (3, 5):
            try:
                sharded["config%d" % which] = run_sharded_config(torch, dist, dev, rank, world, which)
            except Exception as e:
                sharded["config%d" % which] = {"error": repr(e)}
                if dist is not None:
                    raise

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")   # written by tools/make_traffic.py from the latest ncu launch list
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "round2_traffic.json")
    if os.path.exists(tp) and args.precision == "bf16" and B == B_PER_GPU and T == T_FRAMES:
        traffic = json.load(open(tp))  # dram__bytes_read+write per launch from the committed ncu capture of this workload
    conv_ev_ms, conv_n = prof["conv"]
    tail_prof_ms, tail_n = prof["tail"]
    other_ms, other_n = prof["other"]
    n_conv_step = conv_n // max(1, n_prof)
    non_conv_ms = (tail_prof_ms + other_ms) / n_prof
    conv_ms_step = ms_step - non_conv_ms            # conv kernel time INSIDE the graph-timed region
    # conv_post runs inside the tail kernel on the 16-bit paths: its FLOPs and its time belong to roofline_tail
    flops_conv = flops_step - (2.0 * 72 * 896 * B * (16 * T + 1) if fused_tail_ms else 0.0)
    conv_tflops = flops_conv / (conv_ms_step * 1e-3) / 1e12
    tf32 = args.precision == "tf32"
    pk_s = peaks["tflops_sustained"] * (0.5 if tf32 else 1.0)
    pk_b = peaks["tflops_burst"] * (0.5 if tf32 else 1.0)
    roofline = {
        "kernel": "conv kernels (tcgen05 implicit-GEMM convs: conv_tc_kernel, conv_tm_kernel, pair_tm_kernel, pw_tc_kernel, gate_tm_kernel; all %d launches per step)" % n_conv_step,
        "bound": "tensor", "achieved": conv_tflops, "peak": pk_s, "unit": "TFLOP/s", "frac": conv_tflops / pk_s,
        "frac_vs_burst_peak": conv_tflops / pk_b, "peak_burst": pk_b,
        "traffic": traffic.get("conv_dram_bytes_per_launch"), "traffic_source": traffic.get("source"),
        "algorithmic_flops_per_launch": flops_conv / max(1, n_conv_step),
        "peak_source": peaks["source"] + " bf16_tflops_sustained (kernels timed inside a long step); frac_vs_burst_peak uses bf16_tflops"
                       + (" (x 0.5 assumed for tf32; the measured TF32 GEMM peak is in extras.tf32)" if tf32 else ""),
        "flops_per_step": flops_step, "flops_in_conv_launches": flops_conv, "avg_launch_ms": conv_ms_step / max(1, n_conv_step),
        "share_of_step": conv_ms_step / ms_step,
        "timed_in": "the graph-timed region: ms_per_step minus the event-timed non-conv launches (tail %.3f + other %.3f ms)"
                    % (tail_prof_ms / n_prof, other_ms / n_prof),
        "eager_event_sum_ms": conv_ev_ms / n_prof,
    }
    if sustained:
        c_sus = flops_conv / ((sustained["ms_per_step"] - non_conv_ms) * 1e-3) / 1e12
        roofline["sustained_leg"] = {"achieved": c_sus, "frac": c_sus / pk_s, "frac_vs_burst_peak": c_sus / pk_b,
                                     "ms_per_step": sustained["ms_per_step"], "seconds": sustained["seconds"],
                                     "clocks": sustained["clocks"]}
    # fused conv_post + tail kernel (what the step runs): per latent frame it reads 16 frames x 128 ch x 2 B = 4096 B of
    # operands and writes 1024 B of samples ("conv_post fused" variant), and does 2 x 72 x 896 MACs per frame
    # on the tensor cores.  Both ceilings are reported; the binding one is the tensor time.
    roofline_tail = None
    if fused_tail_ms:
        fbytes = B * T * 5120.0 + B * 128 * 2.0        # + the extra reflect-pad frame per utterance
        fflops = 2.0 * 72 * 896 * B * (16 * T + 1)
        t_hbm, t_tc = fbytes / (peaks["hbm_gbs"] * 1e9), fflops / (peaks["tflops_burst"] * 1e12)
        roofline_tail = {
            "kernel": "tail_fused_kernel (conv_post on tcgen05 + head + iSTFT + PQMF, logits stay in TMEM)",
            "variant": "conv_post fused: reads the bf16 stage-1 operand tensor (4096 B per latent frame), writes 1024 B",
            "bound": "tensor" if t_tc > t_hbm else "hbm",
            "achieved": fflops / (fused_tail_ms * 1e-3) / 1e12, "peak": peaks["tflops_burst"], "unit": "TFLOP/s",
            "frac": max(t_tc, t_hbm) / (fused_tail_ms * 1e-3),
            "hbm": {"achieved": fbytes / (fused_tail_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": fbytes / (fused_tail_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
            "traffic": traffic.get("tail_dram_bytes_per_launch"),
            "peak_source": peaks["source"] + " bf16_tflops (burst) and hbm_gbs; kernel timed alone",
            "ms": fused_tail_ms, "ms_inside_step": tail_prof_ms / tail_n if tail_n else None,
            "bytes_per_launch": fbytes, "flops_per_launch": fflops,
            "two_kernel_path": {"conv_post_ms": "see profiles/round2_launch_times_bf16_split_tail.txt", "tail_ms": tail_ms,
                                "tail_hbm_frac": B * T * 5632.0 / (tail_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                "tail_bytes_per_launch": B * T * 5632.0},
        }
    elif tail_ms:
        tail_bytes = B * T * 5632.0  # 4608 B logits in + 1024 B waveform out per latent frame
        roofline_tail = {
            "kernel": "tail_mb3_kernel (head + iSTFT + PQMF on fp32 logits)", "bound": "hbm",
            "achieved": tail_bytes / (tail_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": tail_bytes / (tail_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": traffic.get("tail_dram_bytes_per_launch"),
            "peak_source": peaks["source"] + " hbm_gbs (burst; kernel timed alone)", "ms": tail_ms,
            "ms_inside_step": tail_prof_ms / tail_n if tail_n else None, "bytes_per_launch": tail_bytes,
        }

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        import contextlib
        Bs = CPU_SAMPLE_B
        with contextlib.redirect_stdout(sys.stderr):
            times, cores, kind = cpu_reference_run(cfg, sd, Bs, T, 3, 1)
        best = min(times)
        what = "unmodified reference modules (baseline/_ref)" if kind == "reference" else "oracle port (torch CPU ops of the reference)"
        cpu_baseline = {"value": Bs * T * 256 / best, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{what} on B={Bs} x T={T}, best of 3 after 1 warm-up", "ms": best * 1e3}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"{CONFIG_NAME} flow-reverse + decoder-from-z, B={B} x T={T} per GPU "
                               f"({samples_per_step} samples = {samples_per_step / sr:.1f} s audio per step per GPU)",
                   "sampling_rate": sr, "l2": "working set per step (~4 GB of activations) >> 126 MB L2; no explicit flush",
                   "residual_stream": eng.residual, "accumulate": "fp32",
                   "submission": "eager launches" if graph is None else "one CUDA-graph replay per step",
                   "legs": "value and e2e are each timed from an idle GPU: (2 s idle before the e2e leg,) 3 warm-up steps, K timed steps; "
                           "the >= 3 s sustained leg (roofline.sustained_leg) runs between them"},
        "rtf": ms_step * 1e-3 / (world * samples_per_step / sr),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(zp_pin[0].numel() * 4 + mask_pin[0].numel() * 4),
                "d2h_bytes_per_step": int(wav_pin[0].numel() * 4),
                "how": "HostStream(fused=True) (public API): per step pinned-host latents+mask -> device, one flow_decode "
                       "call, waveform -> pinned host; copies of neighbouring steps overlap the compute (3 streams, 2 slots)",
                "ms_per_step_module_calls": e2e_modules_ms, "ms_per_step_unpipelined": e2e_seq_ms, "wall_ms_per_step": e2e_wall_ms, "wav_abs_max": wav_check},
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "roofline": roofline, "roofline_tail": roofline_tail, "cpu_baseline": cpu_baseline,
        "extras": {"decoder_only_ms": dec_ms, "decoder_only_samples_per_s": samples_per_step / (dec_ms * 1e-3),
                   "kernel_ms_per_step": {"conv_in_graph_region": conv_ms_step, "tail": tail_prof_ms / n_prof, "other": other_ms / n_prof,
                                          "conv_eager_event_sum": conv_ev_ms / n_prof, "eager_step": ms_profiled},
                   "tflops_whole_step": flops_step / (ms_step * 1e-3) / 1e12,
                   "sustained_ms_per_step": sustained["ms_per_step"] if sustained else None,
                   "tf32": tf32_line, "torch_eager_gpu": torch_eager, **sharded},
    }
    if dist is not None:
        dist.destroy_process_group()
    sys.stdout.flush()
    _emit(line, json_fd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["fp16", "bf16", "tf32", "fp32"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--frames", type=int, default=T_FRAMES)
    ap.add_argument("--flags", type=int, default=0, help="extra mbv_config flags (A/B measurements: 16 fused pairs, 32 cluster pairs, 64 branches, 128 split tail)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline and PyTorch-eager legs")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 3 s sustained leg")
    ap.add_argument("--no-tf32", action="store_true", help="skip the tf32 sub-line")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 3 and 5")
    ap.add_argument("--quick", action="store_true", help="only the device-resident and e2e timings (A/B runs)")
    ap.add_argument("--residual", default=None, choices=["fp16", "fp32"], help="ResBlock residual-stream storage (default: fp16 for bf16)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    args = ap.parse_args()
    if args.quick:
        args.no_cpu = args.no_sustained = args.no_tf32 = args.no_configs = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
