import sys, torch
sys.path.insert(0,'/root/repo')
from mb_istft_vits_b200 import Engine, get_config, synth
cfg=get_config('ljs_mb_istft_vits'); sd=synth.make_state_dict(cfg)
eng=Engine(cfg, sd, precision='bf16')
z,m,_=synth.make_latents(cfg,64,862)
z=z.cuda(); m=m.cuda()
eng.flow_reverse(z,m); torch.cuda.synchronize()
