"""Utterance sharding across the GPUs of one box, and the final variable-length waveform gather.

Every utterance is independent through the flow and the decoder (no batch statistics, no cross-utterance op), so
the path shards by utterance with NO collective on the hot path.  One process per GPU
(as the reference's ``mp.spawn``, train_latest.py:55); weights are replicated.  The only exchange is at the end:
each rank's waveforms go to the consumer rank -- one small placement table, then the samples point-to-point -- over
``torch.distributed`` (NCCL on NVLink 5 / NVSwitch on the box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def balance_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time binning on padded work.  Cost of a bin = (#utterances) x (its longest utterance),
    because a bin is decoded as one padded batch.  Returns utterance indices per rank, each sorted by descending
    length; deterministic for equal inputs."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    bins: List[List[int]] = [[] for _ in range(world_size)]
    tmax = [0] * world_size
    for i in order:
        L = int(lengths[i])
        # the bin whose padded cost grows least; ties -> fewest utterances -> lowest rank
        best = min(range(world_size), key=lambda r: ((len(bins[r]) + 1) * max(tmax[r], L), len(bins[r]), r))
        bins[best].append(i)
        tmax[best] = max(tmax[best], L)
    return bins


def deal_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Snake-order deal of the length-sorted utterances: every rank gets the same count (+-1) and a near-identical length
    distribution -- the right split when each rank then decodes its share in length buckets (``bucket_by_length``), where
    the cost of a rank is close to the SUM of its lengths rather than count x longest.  Deterministic; indices per rank
    sorted by descending length."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    bins: List[List[int]] = [[] for _ in range(world_size)]
    for pos, i in enumerate(order):
        rnd, k = divmod(pos, world_size)
        bins[k if rnd % 2 == 0 else world_size - 1 - k].append(i)
    return bins


def bucket_by_length(lengths: Sequence[int], max_buckets: int = 4, overhead: int = 4000) -> List[List[int]]:
    """Split one rank's utterances into at most ``max_buckets`` padded batches of similar length (the reference pads a
    whole batch to its longest utterance, models.py:717-722 -- with lengths of 1-60 s half of that work is padding).
    Exact dynamic programme over the length-sorted list minimising  sum over buckets (count x longest + overhead);
    ``overhead`` is the fixed cost of one more flow_decode call in utterance-frames (launch sequence + partly filled
    waves; ~0.5 ms of a B200 at 7 frames per microsecond).  Returns positions into ``lengths`` per bucket, longest
    bucket first, each sorted by descending length."""
    n = len(lengths)
    if n == 0:
        return []
    order = sorted(range(n), key=lambda i: (-int(lengths[i]), i))
    ls = [int(lengths[i]) for i in order]
    K = max(1, min(int(max_buckets), n))
    INF = float("inf")
    # best[k][j]: cheapest way to cover the first j sorted utterances with exactly k buckets
    best = [[INF] * (n + 1) for _ in range(K + 1)]
    cut = [[0] * (n + 1) for _ in range(K + 1)]
    best[0][0] = 0
    for k in range(1, K + 1):
        for j in range(1, n + 1):
            for i in range(k - 1, j):
                if best[k - 1][i] == INF:
                    continue
                c = best[k - 1][i] + (j - i) * ls[i] + overhead
                if c < best[k][j]:
                    best[k][j], cut[k][j] = c, i
    k = min(range(1, K + 1), key=lambda kk: (best[kk][n], kk))
    bounds, j = [], n
    while k > 0:
        i = cut[k][j]
        bounds.append((i, j))
        j, k = i, k - 1
    return [[order[p] for p in range(i, j)] for i, j in reversed(bounds)]


def shard_batch(z: torch.Tensor, lengths: torch.Tensor, rank: int, world_size: int
                ) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    """This rank's slice of a padded latent batch [B, C, T]: (z_local trimmed to its own longest utterance,
    lengths_local, global utterance indices)."""
    idx = balance_utterances([int(v) for v in lengths], world_size)[rank]
    if not idx:
        return z[:0, :, :1], lengths[:0], idx
    sel = torch.as_tensor(idx, dtype=torch.long, device=z.device)
    lens = lengths.to(z.device)[sel]
    tmax = int(lens.max())
    return z.index_select(0, sel)[:, :, :tmax].contiguous(), lens, idx


def gather_waveforms(wav: torch.Tensor, n_samples: torch.Tensor, indices: Sequence[int], total: int, dst: int = 0,
                     group: Optional[dist.ProcessGroup] = None, mode: str = "p2p",
                     offsets: Optional[torch.Tensor] = None) -> Optional[List[torch.Tensor]]:
    """Collect every rank's waveforms on rank ``dst`` in the original utterance order.

    wav: [b_local, 1, S_local] (one padded batch) -- or, with ``offsets`` [b_local], ANY contiguous buffer in which
    utterance j starts at element offsets[j] (the flat buffer a rank's length buckets wrote their batches into, see
    ``decode_in_buckets``); n_samples: [b_local] valid sample counts, indices: global utterance ids.
    One small collective, one device->host read and one grouped point-to-point exchange:
      1. every rank writes (valid samples, start offset) of its own utterances into a [total, 2] int64 table (-1 elsewhere)
         and appends the size of its buffer; one ``all_gather_into_tensor`` + ONE ``.cpu()`` gives every rank the whole
         placement -- no per-utterance host sync;
      2. each rank sends its buffer, exactly as it lies in memory (no padding to the longest rank), to ``dst``; ``dst``
         posts one receive per peer.  The operations go out as one ``batch_isend_irecv`` group (NCCL:
         ncclGroupStart/End, all peers in flight at once over NVSwitch; only ``dst`` receives -- the round-1 version
         all-gathered the padded samples to EVERY rank, world x the traffic).
    ``mode="allgather"`` replaces step 2 by ONE ``all_gather_into_tensor`` of the buffers, each padded to the largest:
    world x the bytes, every rank receives everything -- but on an NVSwitch box every GPU has full bandwidth to every peer
    and the collective runs at several times the rate NCCL's grouped send / recv reaches into a single receiver
    (tools/gather_bench.py, profiles/round2_gather_modes.txt), so it is the faster way to get the audio to ``dst`` there.
    Returns the list of trimmed 1-D waveforms (views of the received buffers) on ``dst``, None elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = wav.device
    b_local = len(indices)
    flat = wav.contiguous().reshape(-1) if b_local else None
    numel_local = int(flat.numel()) if b_local else 0
    table = torch.full((total + 1, 2), -1, dtype=torch.int64, device=dev)
    if b_local:
        idx = torch.as_tensor(list(indices), dtype=torch.int64, device=dev)
        if offsets is None:   # one padded batch: row j starts at j x row pitch
            offsets = torch.arange(b_local, dtype=torch.int64, device=dev) * int(wav.shape[-1])
        table[idx, 0] = n_samples.to(dev, torch.int64)
        table[idx, 1] = offsets.to(dev, torch.int64)
    table[total, 0] = numel_local
    table[total, 1] = b_local
    tables = torch.empty((world * (total + 1), 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(tables, table, group=group)
    host = tables.cpu().view(world, total + 1, 2)  # the only device->host read of the gather
    sizes = [int(host[r, total, 0]) for r in range(world)]
    if mode == "allgather":
        numel = max(sizes)
        if flat is None or numel_local < numel:
            padded = torch.zeros(numel, dtype=torch.float32, device=dev)
            if flat is not None:
                padded[:numel_local] = flat
            flat = padded
        everything = torch.empty(world * numel, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(everything, flat, group=group)
        if rank != dst:
            return None
        bufs = [everything[r * numel: r * numel + sizes[r]] if sizes[r] else None for r in range(world)]
        return _assemble(host, bufs, world, total)
    if mode != "p2p":
        raise ValueError("mode must be 'p2p' or 'allgather'")
    ops, bufs = [], [None] * world
    if rank == dst:
        for r in range(world):
            if r == dst or sizes[r] == 0:
                continue
            bufs[r] = torch.empty(sizes[r], dtype=torch.float32, device=dev)
            ops.append(dist.P2POp(dist.irecv, bufs[r], dist.get_global_rank(group, r) if group is not None else r, group))
        bufs[dst] = flat
    elif b_local:
        ops.append(dist.P2POp(dist.isend, flat, dist.get_global_rank(group, dst) if group is not None else dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    return _assemble(host, bufs, world, total)


def _assemble(host, bufs, world, total):
    """Placement table -> the trimmed per-utterance views of the received buffers, in the original utterance order."""
    out: List[Optional[torch.Tensor]] = [None] * total
    ns_all, off_all = host[:, :total, 0].tolist(), host[:, :total, 1].tolist()
    for r in range(world):
        for gi in range(total):
            if ns_all[r][gi] >= 0:
                assert out[gi] is None, f"utterance {gi} was produced by two ranks"
                out[gi] = bufs[r][off_all[r][gi]: off_all[r][gi] + ns_all[r][gi]]
    assert all(o is not None for o in out), "an utterance was not produced by any rank"
    return out  # type: ignore[return-value]


def decode_in_buckets(engine, z_p: torch.Tensor, lengths: Sequence[int], g: Optional[torch.Tensor] = None,
                      max_buckets: int = 4, overhead: int = 4000, out: Optional[torch.Tensor] = None, plan=None):
    """flow reverse + decoder of one rank's variable-length utterances as up to ``max_buckets`` padded batches of similar
    length (``bucket_by_length``) instead of one batch padded to the longest utterance.

    z_p: [b, C, T_max] prior latents (rows in the caller's order, zero beyond an utterance's length is NOT required: the
    mask is rebuilt per bucket from ``lengths``); g: [b, gin, 1] or None.  Every bucket's waveform batch [b_k, 1, 256 T_k]
    is written straight into one flat buffer (``out``, allocated when None), so the gather needs no repacking.
    Returns (flat buffer, offsets [b] int64 on the device: element offset of utterance j, plan).  ``plan`` (the value
    returned by a previous call with the same lengths) skips the bucketing and the index tensors.
    Samples of an utterance further than the decoder's receptive field from its end are those of any other padding."""
    dev = z_p.device
    spf = engine.spf
    if plan is None:
        buckets = bucket_by_length(lengths, max_buckets, overhead)
        plan, off, offs_host = [], 0, [0] * len(lengths)
        for bk in buckets:
            T = max(int(lengths[j]) for j in bk)
            sel = torch.as_tensor(bk, dtype=torch.long, device=dev)
            lens = torch.as_tensor([int(lengths[j]) for j in bk], dtype=torch.long, device=dev)
            mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).to(torch.float32).unsqueeze(1).contiguous()
            for r, j in enumerate(bk):
                offs_host[j] = off + r * spf * T
            plan.append((sel, T, mask, off, len(bk)))
            off += len(bk) * spf * T
        for _, T, _, _, b in sorted(plan, key=lambda e: -e[1] * e[4]):
            engine.reserve_workspace(b, T)   # the largest bucket first: at most one (synchronising) reallocation, here
        plan = (plan, off, torch.as_tensor(offs_host, dtype=torch.int64, device=dev))
    steps, numel, offsets = plan
    if out is None:
        out = torch.empty(numel, dtype=torch.float32, device=dev)
    assert out.numel() >= numel
    for sel, T, mask, off, b in steps:
        zb = (z_p.index_select(0, sel)[:, :, :T] * mask).contiguous()
        gb = None if g is None else g.index_select(0, sel).contiguous()
        engine.flow_decode(zb, mask, gb, want_z=False, out_wav=out[off: off + b * spf * T].view(b, 1, spf * T))
    return out, offsets, plan
