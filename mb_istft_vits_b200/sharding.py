"""Utterance sharding across the GPUs of one box, and the final variable-length waveform gather.

Every utterance is independent through the flow and the decoder (no batch statistics, no cross-utterance op), so
the path shards by utterance with NO collective on the hot path.  One process per GPU
(as the reference's ``mp.spawn``, train_latest.py:55); weights are replicated.  The only exchange is at the end:
each rank's waveforms go to the consumer rank -- lengths first, then the padded samples -- over
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
                     group: Optional[dist.ProcessGroup] = None) -> Optional[List[torch.Tensor]]:
    """Collect every rank's waveforms on rank ``dst`` in the original utterance order.

    wav: [b_local, 1, S_local] (padded), n_samples: [b_local] valid sample counts, indices: global utterance ids.
    Two collectives: an all_gather of (count, max length) then one padded all_gather of the samples (sub-millisecond
    over NVSwitch for a 256 x 10 s batch = 226 MB).  Returns the list of trimmed 1-D waveforms on ``dst``, None elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = wav.device
    b_local = wav.shape[0]
    meta = torch.tensor([b_local, wav.shape[-1] if b_local else 0], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    b_max = max(int(m[0]) for m in metas)
    s_max = max(int(m[1]) for m in metas)
    pad = torch.zeros((b_max, s_max), dtype=torch.float32, device=dev)
    info = torch.full((b_max, 2), -1, dtype=torch.int64, device=dev)  # (global index, valid samples)
    if b_local:
        pad[:b_local, : wav.shape[-1]] = wav[:, 0, :]
        info[:b_local, 0] = torch.as_tensor(list(indices), dtype=torch.int64, device=dev)
        info[:b_local, 1] = n_samples.to(dev, torch.int64)
    pads = [torch.empty_like(pad) for _ in range(world)]
    infos = [torch.empty_like(info) for _ in range(world)]
    dist.all_gather(pads, pad, group=group)
    dist.all_gather(infos, info, group=group)
    if rank != dst:
        return None
    out: List[Optional[torch.Tensor]] = [None] * total
    for r in range(world):
        for j in range(int(metas[r][0])):
            gi, ns = int(infos[r][j, 0]), int(infos[r][j, 1])
            out[gi] = pads[r][j, :ns].clone()
    assert all(o is not None for o in out), "an utterance was not produced by any rank"
    return out  # type: ignore[return-value]
