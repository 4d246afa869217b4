#!/usr/bin/env python
"""Three plain hot-path steps at BASELINE size (config 2: ljs_mb, B = 64, T = 862, bf16), eager launches: the command ncu wraps.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \\
        --log-file gpurun_out/launches.csv python tools/profile_step.py
Launch order inside one step (77 launches): pack_input, 36 flow convs, conv_pre, ups.0, 18 stage-0 ResBlock convs (k3 x6, k7 x6,
k11 x6), ups.1, 18 stage-1 ResBlock convs, tail_fused_kernel.  Among the conv_tc launches of a step: index 50 = stage-0 k=11 c1
(CTA pairs), 58 = stage-1 k=3 c2 (HBM-bound), 1 = flow gate conv.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mb_istft_vits_b200 import Engine, get_config, synth  # noqa: E402

cfg = get_config("ljs_mb_istft_vits")
eng = Engine(cfg, synth.make_state_dict(cfg, seed=1234), precision=os.environ.get("MBV_PREC", "bf16"),
             flags=int(os.environ.get("MBV_FLAGS", "0")))
z_p, mask, _ = synth.make_latents(cfg, 64, 862, seed=1234)
z_p, mask = z_p.cuda(), mask.cuda()
for _ in range(3):
    eng.flow_decode(z_p, mask, want_z=False)
torch.cuda.synchronize()
print("launches per step:", eng.last_launch_count())
