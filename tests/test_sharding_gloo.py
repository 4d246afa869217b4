"""world_size-2 gloo test of the multi-GPU host logic (utterance sharding + final waveform gather) on CPU.
The per-rank 'decoder' here is the CPU oracle: what is under test is the sharding / gather plumbing."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mb_istft_vits_b200.sharding import balance_utterances, gather_waveforms, shard_batch


def test_balance_is_a_partition_and_balances_padded_work():
    lengths = [3750, 63, 125, 312, 625, 1250, 1875, 2812, 900, 901, 77, 3000]
    for w in (1, 2, 4, 8):
        bins = balance_utterances(lengths, w)
        assert sorted(i for b in bins for i in b) == list(range(len(lengths)))
        cost = [len(b) * max((lengths[i] for i in b), default=0) for b in bins]
        assert max(cost) <= 1.6 * (sum(cost) / w) + max(lengths)
    assert balance_utterances([5, 5, 5, 5], 2) == [[0, 2], [1, 3]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import mbistft_oracle as orc
    from mb_istft_vits_b200 import get_config, synth
    torch.set_num_threads(2)
    cfg = get_config("ljs_mini_mb_istft_vits")
    sd = synth.make_state_dict(cfg, seed=1)
    lengths = torch.tensor([30, 7, 18, 25, 11])
    z, mask, _ = synth.make_latents(cfg, 5, 30, seed=2, lengths=lengths)
    z_loc, len_loc, idx = shard_batch(z, lengths, rank, world)
    wav = orc.decode(sd, cfg, z_loc)[0] if len(idx) else torch.zeros((0, 1, 0))
    out = gather_waveforms(wav, len_loc * 256, idx, total=5, dst=0)
    out_ag = gather_waveforms(wav, len_loc * 256, idx, total=5, dst=0, mode="allgather")
    if rank == 0:
        full = orc.decode(sd, cfg, z)[0]
        ok = all(torch.equal(a, b) for a, b in zip(out, out_ag))  # the two transports deliver the same samples
        for i in range(5):
            n = int(lengths[i]) * 256
            # an utterance decoded inside a shorter padded batch equals the full-batch result away from the padded tail
            # (the decoder's receptive field is +-24 latent frames = 6144 samples)
            m = max(0, n - 24 * 256)
            ok &= out[i].shape[0] == n and bool(torch.allclose(out[i][:m], full[i, 0, :m], atol=1e-5))
        torch.save(ok, os.path.join(tmp, "ok.pt"))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_decode_gather_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert torch.load(os.path.join(str(tmp_path), "ok.pt")) is True


# ---- the same plumbing on hardware: two ranks, two GPUs, NCCL, the CUDA decoder (needs `gpurun --gpus 2`)
def _nccl_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from mb_istft_vits_b200 import Engine, get_config, synth
    cfg = get_config("ljs_mini_mb_istft_vits")
    sd = synth.make_state_dict(cfg, seed=1)
    lengths = torch.tensor([30, 7, 18, 25, 11, 29, 3])
    z, mask, _ = synth.make_latents(cfg, 7, 30, seed=2, lengths=lengths)
    eng = Engine(cfg, sd, precision="fp32", device=rank)
    z_loc, len_loc, idx = shard_batch(z, lengths, rank, world)
    wav = eng.decode(z_loc.cuda(rank), want_mb=False, want_spec=False)[0]
    out = gather_waveforms(wav, (len_loc * 256).cuda(rank), idx, total=7, dst=0)
    if rank == 0:
        full = eng.decode(z.cuda(0), want_mb=False, want_spec=False)[0]
        ok = True
        for i in range(7):
            n = int(lengths[i]) * 256
            m = max(0, n - 24 * 256)
            ok &= out[i].shape[0] == n and out[i].device.index == 0
            ok &= bool(torch.allclose(out[i][:m], full[i, 0, :m], atol=1e-5))
        torch.save(bool(ok), os.path.join(tmp, "ok_nccl.pt"))
    else:
        assert out is None
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_shard_decode_gather_two_gpus_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (two NCCL ranks must never share one GPU)")
    port = _free_port()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert torch.load(os.path.join(str(tmp_path), "ok_nccl.pt")) is True
