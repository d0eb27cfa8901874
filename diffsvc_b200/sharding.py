"""Batch-of-slices sharding across the GPUs of one box.

Utterance slices are independent (conditioning, noise, sampler state and vocoder are per item), so a
job shards with NO data-path collective: rank r runs its own sub-batch through sampler + vocoder
locally with replicated weights, and ONE variable-length all-gather of the final fp32 waveforms
(NCCL over NVLink on GPUs, gloo in the CPU tests) reassembles the job.  The reference itself has no
multi-GPU inference (infer_tools/infer_tool.py:277 batches exactly one item).
"""
from typing import List, Sequence

import torch


def partition_slices(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-first bin packing of slice indices onto ranks (balances total frames)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    bins = [[] for _ in range(world_size)]
    load = [0] * world_size
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(b) for b in bins]


def gather_waveforms(local_wavs: List[torch.Tensor], local_ids: List[int], n_total: int, group=None) -> List[torch.Tensor]:
    """All-gather variable-length waveforms: lengths first, then one padded payload exchange.
    Returns the full job's waveforms in slice order on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = local_wavs[0].device if local_wavs else torch.device("cpu")
    meta = torch.full((n_total, 2), -1, dtype=torch.int64, device=dev)        # per local slot: (slice id, samples)
    for k, (i, w) in enumerate(zip(local_ids, local_wavs)):
        meta[k, 0], meta[k, 1] = i, w.numel()
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    max_total = max(int(m[:, 1].clamp(min=0).sum()) for m in metas)
    payload = torch.zeros(max(max_total, 1), dtype=torch.float32, device=dev)
    if local_wavs:
        flat = torch.cat([w.reshape(-1).to(torch.float32) for w in local_wavs])
        payload[:flat.numel()] = flat
    payloads = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(payloads, payload, group=group)
    out = [None] * n_total
    for m, p in zip(metas, payloads):
        off = 0
        for i, n in m.tolist():
            if i < 0:
                continue
            out[i] = p[off:off + n].clone()
            off += n
    assert all(o is not None for o in out), "a slice was not produced by any rank"
    return out
