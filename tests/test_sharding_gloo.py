"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch of slices and all-gather the
(variable-length) waveforms; every rank must end with the whole job in slice order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffsvc_b200.sharding import gather_waveforms, partition_slices


def test_partition_balances_frames():
    lengths = [689, 120, 900, 450, 300, 700, 50, 610]
    parts = partition_slices(lengths, 3)
    assert sorted(i for p in parts for i in p) == list(range(len(lengths)))
    loads = [sum(lengths[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(lengths)
    assert partition_slices(lengths, 1) == [list(range(len(lengths)))]
    assert partition_slices([5], 4) == [[0], [], [], []]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, lengths, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = partition_slices(lengths, world)[rank]
        # stand-in for sampler + vocoder: a deterministic per-slice "waveform" of hop * frames samples
        wavs = [torch.arange(lengths[i] * 4, dtype=torch.float32) * 0.5 + i for i in mine]
        full = gather_waveforms(wavs, mine, len(lengths))
        ok = all(torch.equal(full[i], torch.arange(lengths[i] * 4, dtype=torch.float32) * 0.5 + i) for i in range(len(lengths)))
        q.put((rank, ok, [int(w.numel()) for w in full]))
    finally:
        dist.destroy_process_group()


def _run(lengths, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lengths, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, sizes in res:
        assert ok, rank
        assert sizes == [n * 4 for n in lengths]


def test_two_rank_gather_gloo():
    _run([7, 3, 11, 5, 2])


def test_gather_with_an_empty_rank():
    """n_slices < world_size: partition_slices leaves a rank without work; it must still join both collectives."""
    _run([9], world=2)
    _run([4, 6], world=3)
