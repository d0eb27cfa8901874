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


def gather_waveforms(local_wavs: List[torch.Tensor], local_ids: List[int], n_total: int, group=None,
                     device=None) -> List[torch.Tensor]:
    """All-gather variable-length waveforms: one small metadata exchange (slice id, samples), then ONE padded
    payload exchange.  Returns the full job's waveforms in slice order on every rank.

    `device`: where the collective's buffers live.  Default: the local waveforms' device, else the current CUDA
    device under the NCCL backend (a rank that received no slice must still join the collective with CUDA
    tensors), else the CPU (gloo)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if device is None:
        if local_wavs:
            device = local_wavs[0].device
        elif dist.get_backend(group) == "nccl":
            device = torch.device("cuda", torch.cuda.current_device())
        else:
            device = torch.device("cpu")
    meta = torch.full((n_total, 2), -1, dtype=torch.int64)               # per local slot: (slice id, samples)
    for k, (i, w) in enumerate(zip(local_ids, local_wavs)):
        meta[k, 0], meta[k, 1] = int(i), int(w.numel())
    meta = meta.to(device)
    metas = torch.empty((world * n_total, 2), dtype=torch.int64, device=device)     # concatenated along dim 0
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas = metas.view(world, n_total, 2).cpu()                                                  # the one host sync: payload sizes
    max_total = int(metas[:, :, 1].clamp(min=0).sum(dim=1).max())
    payload = torch.zeros(max(max_total, 1), dtype=torch.float32, device=device)
    if local_wavs:
        flat = torch.cat([w.reshape(-1).to(device=device, dtype=torch.float32) for w in local_wavs])
        payload[:flat.numel()] = flat
    payloads = torch.empty(world * payload.numel(), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(payloads, payload, group=group)
    payloads = payloads.view(world, payload.numel())
    out = [None] * n_total
    for r in range(world):
        off = 0
        for i, n in metas[r].tolist():
            if i < 0:
                continue
            out[i] = payloads[r, off:off + n].clone()
            off += n
    assert all(o is not None for o in out), "a slice was not produced by any rank"
    return out
