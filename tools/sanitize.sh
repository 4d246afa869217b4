#!/bin/bash
# compute-sanitizer racecheck of the three-role mbarrier / TMEM kernels on the mini config: default path, cluster-pair
# multicast variant, fused conv-pair kernel, ResBlock branches.  One tool per gpurun call (profiling guide).
#   gpurun -- 'bash tools/sanitize.sh racecheck > gpurun_out/sanitize_racecheck.log 2>&1'
TOOL=${1:-racecheck}
cat > /tmp/san_probe.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from mb_istft_vits_b200 import Engine, get_config, synth
flags = int(sys.argv[1])
cfg = get_config("ljs_mini_mb_istft_vits" if flags != 16 else "ljs_mb_istft_vits")
sd = synth.make_state_dict(cfg, seed=1234)
eng = Engine(cfg, sd, precision="bf16", flags=flags)
z_p, mask, _ = synth.make_latents(cfg, 2, 40, seed=7, lengths=[40, 29])
z, wav, _, _, _ = eng.flow_decode(z_p.cuda(), mask.cuda())
torch.cuda.synchronize()
print("flags", flags, "wav abs max", float(wav.abs().max()), "launches", eng.last_launch_count())
PY
for FLAGS in 0 32 16 64; do
  echo "=== compute-sanitizer --tool $TOOL, mbv_config.flags = $FLAGS"
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python /tmp/san_probe.py $FLAGS 2>&1 | grep -v "^=========     at \|^=========         in \|Host Frame" | tail -25
done
