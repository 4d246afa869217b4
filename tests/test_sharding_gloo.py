"""world_size-2 gloo test of the multi-GPU host logic (utterance sharding + final waveform gather) on CPU.
The per-rank 'decoder' here is the CPU oracle: what is under test is the sharding / gather plumbing."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mb_istft_vits_b200.sharding import (balance_utterances, bucket_by_length, deal_utterances, decode_in_buckets,
                                         gather_waveforms, shard_batch)


def test_balance_is_a_partition_and_balances_padded_work():
    lengths = [3750, 63, 125, 312, 625, 1250, 1875, 2812, 900, 901, 77, 3000]
    for w in (1, 2, 4, 8):
        bins = balance_utterances(lengths, w)
        assert sorted(i for b in bins for i in b) == list(range(len(lengths)))
        cost = [len(b) * max((lengths[i] for i in b), default=0) for b in bins]
        assert max(cost) <= 1.6 * (sum(cost) / w) + max(lengths)
    assert balance_utterances([5, 5, 5, 5], 2) == [[0, 2], [1, 3]]


def test_deal_is_a_partition_with_equal_counts_and_similar_sums():
    gen = torch.Generator().manual_seed(1234)
    lengths = [int(v) for v in torch.randint(60, 3750, (64,), generator=gen)]
    for w in (1, 2, 4, 8):
        bins = deal_utterances(lengths, w)
        assert sorted(i for b in bins for i in b) == list(range(64))
        assert {len(b) for b in bins} == {64 // w}
        sums = [sum(lengths[i] for i in b) for b in bins]
        assert max(sums) - min(sums) <= max(lengths)
        for b in bins:
            assert [lengths[i] for i in b] == sorted((lengths[i] for i in b), reverse=True)
    assert deal_utterances([9, 8, 7, 6, 5], 2) == [[0, 3, 4], [1, 2]]


def test_length_buckets_partition_and_cut_the_padded_work():
    gen = torch.Generator().manual_seed(7)
    lengths = [int(v) for v in torch.randint(60, 3750, (64,), generator=gen)]

    def cost(bk, overhead):
        return sum(len(b) * max(lengths[i] for i in b) + overhead for b in bk)

    one = cost([list(range(64))], 4000)
    prev = one
    for k in (1, 2, 3, 4, 6):
        bk = bucket_by_length(lengths, k, 4000)
        assert 1 <= len(bk) <= k
        assert sorted(i for b in bk for i in b) == list(range(64))
        flat = [lengths[i] for b in bk for i in b]
        assert flat == sorted(flat, reverse=True)          # buckets are runs of the length-sorted list
        assert cost(bk, 4000) <= prev                       # more buckets allowed never costs more
        prev = cost(bk, 4000)
    assert prev < 0.75 * one                                # uniform lengths: a third of the padded work goes away
    assert bucket_by_length([], 4) == []
    assert bucket_by_length([5], 4) == [[0]]
    assert bucket_by_length([10, 10, 10, 10], 4) == [[0, 1, 2, 3]]      # equal lengths: one batch
    assert len(bucket_by_length(lengths, 4, 10 ** 9)) == 1              # a prohibitive call overhead: one batch
    # brute force over every split into two runs
    srt = sorted(lengths, reverse=True)
    best2 = min(min(i * srt[0] + (64 - i) * srt[i] + 8000 for i in range(1, 64)), 64 * srt[0] + 4000)
    assert cost(bucket_by_length(lengths, 2, 4000), 4000) == best2


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
    # the same utterances dealt in snake order and decoded in length buckets into one flat buffer per rank
    class _OracleEngine:   # stands in for Engine: what is under test is the bucketing / placement / gather plumbing
        spf = 256

        def reserve_workspace(self, b, T):
            return 0

        def flow_decode(self, zb, mask, g, want_z=False, out_wav=None):
            out_wav.copy_(orc.decode(sd, cfg, zb)[0])

    idx_b = deal_utterances([int(v) for v in lengths], world)[rank]
    flat, offs, plan = decode_in_buckets(_OracleEngine(), z[idx_b], [int(lengths[i]) for i in idx_b], max_buckets=2, overhead=1)
    flat2, _, _ = decode_in_buckets(_OracleEngine(), z[idx_b], [int(lengths[i]) for i in idx_b], out=torch.empty_like(flat), plan=plan)
    assert torch.equal(flat, flat2) and len(plan[0]) == 2
    out_bk = [gather_waveforms(flat, lengths[idx_b] * 256, idx_b, total=5, dst=0, offsets=offs, mode=m) for m in ("p2p", "allgather")]
    if rank == 0:
        full = orc.decode(sd, cfg, z)[0]
        ok = all(torch.equal(a, b) for a, b in zip(out, out_ag))  # the two transports deliver the same samples
        for res in out_bk:
            for i in range(5):
                n = int(lengths[i]) * 256
                m = max(0, n - 24 * 256)
                ok &= res[i].shape[0] == n and bool(torch.allclose(res[i][:m], full[i, 0, :m], atol=1e-5))
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


def _worker_empty_rank(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # three utterances on two ranks with dst = 1: rank 1 (the consumer) owns one, rank 0 two; then ONE utterance: rank 1 owns nothing
    ok = True
    for lengths in ([5, 9, 2], [7]):
        total = len(lengths)
        idx = deal_utterances(lengths, world)[rank]
        offs, off = [], 0
        for i in idx:
            offs.append(off)
            off += lengths[i] + 3          # utterances lie in the flat buffer with gaps, as bucket padding leaves them
        flat = torch.full((max(off, 1),), -1.0)
        for i, o in zip(idx, offs):
            flat[o: o + lengths[i]] = torch.arange(lengths[i], dtype=torch.float32) + 100 * i
        for mode in ("p2p", "allgather"):
            out = gather_waveforms(flat, torch.tensor([lengths[i] for i in idx], dtype=torch.int64), idx, total, dst=1, mode=mode,
                                   offsets=torch.tensor(offs, dtype=torch.int64))
            if rank == 1:
                ok &= len(out) == total
                for i in range(total):
                    ok &= bool(torch.equal(out[i], torch.arange(lengths[i], dtype=torch.float32) + 100 * i))
            else:
                ok &= out is None
    if rank == 1:
        torch.save(bool(ok), os.path.join(tmp, "ok_empty.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_with_offsets_an_idle_rank_and_a_non_zero_consumer(tmp_path):
    port = _free_port()
    mp.spawn(_worker_empty_rank, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert torch.load(os.path.join(str(tmp_path), "ok_empty.pt")) is True


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


@pytest.mark.gpu
def test_decode_in_length_buckets_matches_one_padded_batch():
    """decode_in_buckets (several padded batches of similar length, written into one flat buffer) against ONE flow_decode of
    the whole batch padded to its longest utterance -- the reference's way (models.py:717-737).  The flow masks every layer,
    so z is identical; the decoder sees `z * mask`, so samples further than its receptive field (24 latent frames) from an
    utterance's end are identical too."""
    from mb_istft_vits_b200 import Engine, get_config, synth
    cfg = get_config("ljs_mini_mb_istft_vits")
    sd = synth.make_state_dict(cfg, seed=1)
    lengths = [200, 61, 137, 190, 88, 33, 140, 75, 199]
    lens = torch.tensor(lengths)
    z_p, mask, _ = synth.make_latents(cfg, len(lengths), 200, seed=5, lengths=lens)
    eng = Engine(cfg, sd, precision="fp32", device=0)
    z_p, mask = z_p.cuda(), mask.cuda()
    full = eng.flow_decode(z_p, mask, want_z=False)[1]
    flat, offs, plan = decode_in_buckets(eng, z_p, lengths, max_buckets=3, overhead=50)
    assert len(plan[0]) == 3 and flat.numel() == sum(b * 256 * T for _, T, _, _, b in plan[0]) < full.numel()
    offs = offs.tolist()
    for i, n in enumerate(lengths):
        m = (n - 24) * 256
        got = flat[offs[i]: offs[i] + n * 256]
        assert torch.allclose(got[:m], full[i, 0, :m], atol=1e-5), i
    flat2, _, _ = decode_in_buckets(eng, z_p, lengths, plan=plan, out=torch.zeros_like(flat))
    assert torch.equal(flat, flat2)
    eng.close()


def test_decode_in_buckets_plan_offsets_and_reuse_without_a_gpu():
    """Host logic of decode_in_buckets against a recording stand-in engine: every utterance gets a slot of its BUCKET's padded
    length, slots tile the flat buffer without gaps, masks follow the lengths, the plan is reusable and the largest bucket is
    reserved first."""
    calls, reserved = [], []

    class _Rec:
        spf = 4

        def reserve_workspace(self, b, T):
            reserved.append((b, T))
            return 0

        def flow_decode(self, zb, mask, g, want_z=False, out_wav=None):
            calls.append((tuple(zb.shape), mask.sum(dim=(1, 2)).tolist(), None if g is None else tuple(g.shape)))
            out_wav.copy_(zb[:, :1, :].repeat_interleave(4, dim=2))   # "waveform" = channel 0 of the masked latent, 4 samples per frame

    lengths = [9, 2, 7, 3, 8, 1]
    z = torch.arange(6 * 2 * 9, dtype=torch.float32).view(6, 2, 9) + 1.0
    g = torch.ones(6, 5, 1)
    flat, offs, plan = decode_in_buckets(_Rec(), z, lengths, g=g, max_buckets=2, overhead=1)
    steps, numel, _ = plan
    assert [(b, T) for _, T, _, _, b in steps] == [(3, 9), (3, 3)] and numel == 4 * (3 * 9 + 3 * 3) == flat.numel()
    assert reserved == [(3, 9), (3, 3)] and [c[0] for c in calls] == [(3, 2, 9), (3, 2, 3)] and calls[0][2] == (3, 5, 1)
    assert calls[0][1] == [9.0, 8.0, 7.0] and calls[1][1] == [3.0, 2.0, 1.0]
    offs = offs.tolist()
    assert sorted(offs) == [0, 36, 72, 108, 120, 132]
    for i, n in enumerate(lengths):
        want = z[i, 0, :n].repeat_interleave(4)
        assert torch.equal(flat[offs[i]: offs[i] + 4 * n], want), i
    n_calls = len(calls)
    flat2, offs2, _ = decode_in_buckets(_Rec(), z, lengths, g=g, out=torch.zeros_like(flat), plan=plan)
    assert torch.equal(flat2, flat) and torch.equal(offs2, torch.tensor(offs)) and len(calls) == n_calls + 2 and len(reserved) == 2
    with pytest.raises(AssertionError):
        decode_in_buckets(_Rec(), z, lengths, g=g, out=torch.zeros(8), plan=plan)
