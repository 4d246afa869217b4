#!/usr/bin/env python
"""Benchmark of the MB-iSTFT-VITS waveform hot path (flow reverse + iSTFT decoder) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--precision bf16|tf32|fp32]

Workload (BASELINE.json configs[1]): ljs_mb_istft_vits, B = 64 utterances x T = 862 latent frames (10.008 s at
22 050 Hz) per GPU, synthetic latents, seeded random-init weights.  A "step" is one pass of the hot path over
that batch: z = flow(z_p, mask, reverse=True); wav = dec(z * mask)  (models.py:730-734) = 14 123 008 samples.
For N > 1 (launched under torchrun, one rank per GPU) every rank runs its own 64-utterance shard -- utterances are
independent, so there is no collective on the hot path -- and the whole-job value is N * samples / max-over-ranks time.

One JSON line is printed by rank 0:
  value       device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e         the same metric through the public API (NativeFlow / NativeDecoder modules) with pinned HOST inputs
              and a HOST copy of the waveform inside the timed region
  roofline    the conv implicit-GEMM kernel (dominant): algorithmic FLOP/s vs the measured bf16 peak;
              roofline_tail: the fused iSTFT/PQMF tail vs the measured HBM copy bandwidth
  cpu_baseline the CPU oracle port (same torch ops as the reference) timed on this box's host cores (rank 0, N=1)

--impl reference times that CPU oracle port as the reference arm (the reference is Python/PyTorch; its modules
cannot travel to the GPU box, the oracle restates them with the same torch primitives -- see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG_NAME = "ljs_mb_istft_vits"
B_PER_GPU = 64
T_FRAMES = 862
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


def cpu_port_run(cfg, sd, B, T, reps, warmup, threads=None):
    """The oracle port on the host cores: flow reverse + decode, all threads.  Returns best seconds per pass."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mbistft_oracle as orc
    from mb_istft_vits_b200 import synth
    n = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(n)
    z_p, mask, _ = synth.make_latents(cfg, B, T, seed=1234)
    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        orc.flow_decode(sd, cfg, z_p, mask)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, n


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    from mb_istft_vits_b200 import get_config, synth
    cfg = get_config(CONFIG_NAME)
    sd = synth.make_state_dict(cfg, seed=1234)
    Bs = 4  # bounded sample: CPU throughput is batch-independent
    times, cores = cpu_port_run(cfg, sd, Bs, T_FRAMES, max(1, args.steps), max(1, min(args.warmup, 2)))
    mean = sum(times) / len(times)
    samples = Bs * T_FRAMES * 256
    v = samples / mean
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{CONFIG_NAME} flow-reverse + decoder-from-z, CPU sample B={Bs} x T={T_FRAMES}",
                   "sampling_rate": cfg["sampling_rate"]},
        "rtf": mean / (samples / cfg["sampling_rate"]),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle port (torch CPU ops of the reference), B={Bs} x T={T_FRAMES} per step"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _emit(line, fd):
    os.write(fd, (json.dumps(line) + "\n").encode())


def run_native(args):
    # Libraries (e.g. NCCL's version banner) print to stdout; the contract is ONE JSON line there.  Everything but that
    # line is routed to stderr at the file-descriptor level.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    rank, local_rank, world = dist_env()
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
    eng = Engine(cfg, sd, precision=args.precision, device=dev.index, residual=args.residual)
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
    # per-kernel device times for the roofline: the same steps again, eagerly, every launch bracketed by CUDA events
    eng.set_profiling(True)
    eng.profile_read()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for _ in range(args.steps):
        step()
    p1.record()
    barrier()
    ms_profiled = p0.elapsed_time(p1)
    prof = eng.profile_read()
    eng.set_profiling(False)
    t_max = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_step = float(t_max.item()) / args.steps
    value = world * samples_per_step / (ms_step * 1e-3)

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

    # ---------------- end-to-end through the public API, host buffers in and out.  Every step uploads its own
    # pinned latents + mask, runs the drop-in modules (flow, mask, dec: two library calls like infer()) and downloads
    # its waveform into pinned memory.  (a) HostStream: the serving loop, copies of neighbouring steps overlap the
    # compute (three streams, two slots); (b) the same calls strictly one after the other on one stream.
    from mb_istft_vits_b200 import HostStream
    flow, dec = NativeFlow(eng), NativeDecoder(eng, want_mb=False, want_spec=False)
    n_slots = 2
    zp_pin = [z_p_host.clone().pin_memory() for _ in range(n_slots)]
    mask_pin = [mask_host.clone().pin_memory() for _ in range(n_slots)]
    wav_pin = [torch.empty((B, 1, 256 * T), dtype=torch.float32).pin_memory() for _ in range(n_slots)]
    def e2e_pipelined(n):
        for i in range(n):
            hs.submit(zp_pin[i % n_slots], mask_pin[i % n_slots], wav_pin[i % n_slots])

    # (a') module-by-module variant of the serving loop, for the record
    hs = HostStream(eng, depth=n_slots, fused=False)
    e2e_pipelined(3)
    hs.drain()
    barrier()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for st in (hs.s_in, hs.s_cmp, hs.s_out):
        st.wait_event(m0)
    e2e_pipelined(args.steps)
    for st in (hs.s_in, hs.s_cmp, hs.s_out):
        torch.cuda.current_stream().wait_stream(st)
    m1.record()
    barrier()
    e2e_modules_ms = m0.elapsed_time(m1) / args.steps

    hs = HostStream(eng, depth=n_slots, fused=True)

    def e2e_step():
        zp_d = zp_pin[0].to(dev, non_blocking=True)
        m_d = mask_pin[0].to(dev, non_blocking=True)
        z = flow(zp_d, m_d, g=None, reverse=True)
        o, _, _, _ = dec(z * m_d, g=None)
        wav_pin[0].copy_(o, non_blocking=True)

    e2e_pipelined(3)
    hs.drain()
    barrier()
    import time as _time
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    x0.record(cur)
    for st in (hs.s_in, hs.s_cmp, hs.s_out):
        st.wait_event(x0)
    w0 = _time.perf_counter()
    e2e_pipelined(args.steps)
    for st in (hs.s_in, hs.s_cmp, hs.s_out):
        cur.wait_stream(st)
    x1.record(cur)
    barrier()
    e2e_wall_ms = (_time.perf_counter() - w0) * 1e3 / args.steps
    e2e_t = torch.tensor([x0.elapsed_time(x1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t.item()) / args.steps
    e2e_value = world * samples_per_step / (e2e_ms * 1e-3)
    wav_check = float(wav_pin[0].abs().max())  # the downloaded waveform is real data

    for _ in range(3):
        e2e_step()
    barrier()
    y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    y0.record()
    for _ in range(args.steps):
        e2e_step()
    y1.record()
    barrier()
    e2e_seq_ms = y0.elapsed_time(y1) / args.steps

    # ---------------- fused tail alone (HBM roofline of that kernel): event-timed stand-alone launches
    L = 16 * T
    logits = torch.randn((B, L + 1, 72), device=dev) * 0.5
    for _ in range(3):
        eng.tail(logits, T, want_mb=False, want_spec=False)
    torch.cuda.synchronize()
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record()
    for _ in range(10):
        eng.tail(logits, T, want_mb=False, want_spec=False)
    t1e.record()
    torch.cuda.synchronize()
    tail_ms = t0e.elapsed_time(t1e) / 10
    del logits

    # ---------------- the reference's own arithmetic as PyTorch eager ops ON THIS GPU ("the real bar to beat"):
    # the oracle port issues the same F.conv1d / conv_transpose1d / torch.istft calls as the reference modules; fp32 with
    # TF32 off (the gold setting) and with PyTorch's default TF32 convs.  Bounded sample, rank 0 at N = 1 only.
    torch_eager = None
    if world == 1 and not args.no_cpu:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import mbistft_oracle as orc
            Bs = min(B, 16)
            sd_dev = {k: v.to(dev) for k, v in sd.items()}
            zs, ms = z_p[:Bs].contiguous(), mask[:Bs].contiguous()
            torch_eager = {"sample": f"oracle port (the reference's torch ops) on cuda, B={Bs} x T={T}, best of 3", "unit": UNIT}
            for label, tf32 in (("fp32", False), ("tf32", True)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                best = None
                with torch.no_grad():
                    for i in range(4):
                        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        g0.record()
                        orc.flow_decode(sd_dev, cfg, zs, ms)
                        g1.record()
                        torch.cuda.synchronize()
                        if i > 0:
                            t = g0.elapsed_time(g1)
                            best = t if best is None else min(best, t)
                torch_eager[label] = Bs * T * 256 / (best * 1e-3)
                torch_eager[label + "_ms"] = best
            del sd_dev
        except Exception as e:  # the leg is informational: never fail the bench on it
            torch_eager = {"error": repr(e)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tp) and args.precision == "bf16" and B == B_PER_GPU and T == T_FRAMES:
        traffic = json.load(open(tp))  # dram__bytes_read+write per launch from the committed ncu capture of this workload
    flops_step = eng.decode_flops(B, T) + eng.flow_flops(B, T)
    conv_ms, conv_n = prof["conv"]
    tail_prof_ms, tail_n = prof["tail"]
    conv_tflops = flops_step * args.steps / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else None
    tail_bytes = B * T * 5632.0  # 4608 B logits in + 1024 B waveform out per latent frame
    prec_peak = peaks["tflops_sustained"] * (0.5 if args.precision == "tf32" else 1.0)
    roofline = {
        "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all %d launches per step)" % (conv_n // max(1, args.steps)),
        "bound": "tensor", "achieved": conv_tflops, "peak": prec_peak, "unit": "TFLOP/s",
        "frac": (conv_tflops / prec_peak) if conv_tflops else None,
        "traffic": traffic.get("conv_dram_bytes_per_launch"), "traffic_source": traffic.get("source"),
        "algorithmic_flops_per_launch": flops_step / max(1, conv_n // max(1, args.steps)),
        "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)"
                       + (" x 0.5 for tf32" if args.precision == "tf32" else ""),
        "flops_per_step": flops_step, "avg_launch_ms": conv_ms / conv_n if conv_n else None,
        "share_of_step": conv_ms / ms_profiled if ms_profiled else None,
        "timed_in": "eager pass with per-launch CUDA events right after the graph-timed region (%.3f ms/step)" % (ms_profiled / args.steps),
    }
    roofline_tail = {
        "kernel": "tail_kernel (head + iSTFT + PQMF)", "bound": "hbm",
        "achieved": tail_bytes / (tail_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": tail_bytes / (tail_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": traffic.get("tail_dram_bytes_per_launch"),
        "peak_source": peaks["source"] + " hbm_gbs (burst; kernel timed alone)", "ms": tail_ms,
        "ms_inside_step": tail_prof_ms / tail_n if tail_n else None, "bytes_per_launch": tail_bytes,
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        Bs = 4
        times, cores = cpu_port_run(cfg, sd, Bs, T, 3, 1)
        best = min(times)
        cpu_baseline = {"value": Bs * T * 256 / best, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"oracle port (torch CPU ops of the reference) on B={Bs} x T={T}, best of 3 after 1 warm-up",
                        "ms": best * 1e3}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"{CONFIG_NAME} flow-reverse + decoder-from-z, B={B} x T={T} per GPU "
                               f"({samples_per_step} samples = {samples_per_step / sr:.1f} s audio per step per GPU)",
                   "sampling_rate": sr, "l2": "working set per step (~4 GB of activations) >> 126 MB L2; no explicit flush",
                   "residual_stream": eng.residual, "accumulate": "fp32",
                   "submission": "eager launches" if graph is None else "one CUDA-graph replay per step"},
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
                   "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
                   "tflops_whole_step": flops_step / (ms_step * 1e-3) / 1e12,
                   "torch_eager_gpu": torch_eager},
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--residual", default=None, choices=["fp16", "fp32"], help="ResBlock residual-stream storage (default: fp16 for bf16)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
