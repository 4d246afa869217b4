#!/usr/bin/env python
"""Debug aid: run one flow-reverse or decode at BASELINE size with MBV_TIMELINE=<epilogue mode> set, so that
libmbistft prints CTA 0's per-tile clock stamps (producer start | MMA begin..end | epilogue begin..end) per launch.
    MBV_TIMELINE=1 python tools/timeline_probe.py decode      # 0 ACT, 1 RES, 2 F32, 3 GATE, 4 RS, 5 POST
"""
import sys
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from mb_istft_vits_b200 import Engine, get_config, synth

cfg = get_config("ljs_mb_istft_vits")
sd = synth.make_state_dict(cfg)
import os
eng = Engine(cfg, sd, precision=os.environ.get("MBV_PREC", "bf16"), flags=int(os.environ.get("MBV_FLAGS", "0")))
z, m, _ = synth.make_latents(cfg, 64, 862)
z, m = z.cuda(), m.cuda()
if len(sys.argv) > 1 and sys.argv[1] == "decode":
    eng.decode(z, want_mb=False, want_spec=False)
else:
    eng.flow_reverse(z, m)
torch.cuda.synchronize()
