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
                     group: Optional[dist.ProcessGroup] = None, mode: str = "p2p") -> Optional[List[torch.Tensor]]:
    """Collect every rank's waveforms on rank ``dst`` in the original utterance order.

    wav: [b_local, 1, S_local] (padded), n_samples: [b_local] valid sample counts, indices: global utterance ids.
    One small collective, one device->host read and one grouped point-to-point exchange:
      1. every rank writes (valid samples, local row) of its own utterances into a [total, 2] int64 table (-1 elsewhere)
         and appends its (rows, row pitch); one ``all_gather_into_tensor`` + ONE ``.cpu()`` gives every rank the whole
         placement -- no per-utterance host sync;
      2. each rank sends its [b_local, S_local] block, exactly as it lies in memory (no padding to the longest rank), to
         ``dst``; ``dst`` posts one receive per peer.  The operations go out as one ``batch_isend_irecv`` group
         (NCCL: ncclGroupStart/End, all peers in flight at once over NVSwitch; only ``dst`` receives -- the round-1
         version all-gathered the padded samples to EVERY rank, world x the traffic).
    ``mode="allgather"`` replaces step 2 by ONE ``all_gather_into_tensor`` of the blocks, each padded to the largest block:
    world x the bytes, every rank receives everything -- but on an NVSwitch box every GPU has full bandwidth to every peer
    and the collective runs at several times the rate NCCL's grouped send / recv reaches into a single receiver
    (tools/gather_bench.py, profiles/round2_gather_modes.txt), so it is the faster way to get the audio to ``dst`` there.
    Returns the list of trimmed 1-D waveforms (views of the received blocks) on ``dst``, None elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = wav.device
    b_local = int(wav.shape[0])
    s_local = int(wav.shape[-1]) if b_local else 0
    table = torch.full((total + 1, 2), -1, dtype=torch.int64, device=dev)
    if b_local:
        idx = torch.as_tensor(list(indices), dtype=torch.int64, device=dev)
        table[idx, 0] = n_samples.to(dev, torch.int64)
        table[idx, 1] = torch.arange(b_local, dtype=torch.int64, device=dev)
    table[total, 0] = b_local
    table[total, 1] = s_local
    tables = torch.empty((world * (total + 1), 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(tables, table, group=group)
    host = tables.cpu().view(world, total + 1, 2)  # the only device->host read of the gather
    shapes = [(int(host[r, total, 0]), int(host[r, total, 1])) for r in range(world)]
    block = wav.reshape(b_local, s_local).contiguous() if b_local else None
    if mode == "allgather":
        numel = max(r * c for r, c in shapes)
        flat = torch.zeros(numel, dtype=torch.float32, device=dev) if (block is None or block.numel() < numel) else block.reshape(-1)
        if block is not None and block.numel() < numel:
            flat[: block.numel()] = block.reshape(-1)
        everything = torch.empty(world * numel, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(everything, flat, group=group)
        if rank != dst:
            return None
        bufs = [everything[r * numel: r * numel + shapes[r][0] * shapes[r][1]].view(shapes[r]) if shapes[r][0] else None for r in range(world)]
        return _assemble(host, bufs, world, total)
    if mode != "p2p":
        raise ValueError("mode must be 'p2p' or 'allgather'")
    ops, bufs = [], [None] * world
    if rank == dst:
        for r in range(world):
            if r == dst or shapes[r][0] == 0:
                continue
            bufs[r] = torch.empty(shapes[r], dtype=torch.float32, device=dev)
            ops.append(dist.P2POp(dist.irecv, bufs[r], dist.get_global_rank(group, r) if group is not None else r, group))
        bufs[dst] = block
    elif b_local:
        ops.append(dist.P2POp(dist.isend, block, dist.get_global_rank(group, dst) if group is not None else dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    return _assemble(host, bufs, world, total)


def _assemble(host, bufs, world, total):
    """Placement table -> the trimmed per-utterance views of the received blocks, in the original utterance order."""
    out: List[Optional[torch.Tensor]] = [None] * total
    ns_all, row_all = host[:, :total, 0].tolist(), host[:, :total, 1].tolist()
    for r in range(world):
        for gi in range(total):
            if ns_all[r][gi] >= 0:
                assert out[gi] is None, f"utterance {gi} was produced by two ranks"
                out[gi] = bufs[r][row_all[r][gi], : ns_all[r][gi]]
    assert all(o is not None for o in out), "an utterance was not produced by any rank"
    return out  # type: ignore[return-value]
