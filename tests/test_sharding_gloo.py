"""N>1 host logic on CPU: world_size-2 gloo run of the ray sharding + frame assembly."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from instant_nvr_b200.sharding import assemble, render_sharded, shard_capacity, shard_indices


def test_shards_partition_the_frame():
    for n, world, tile in ((10000, 3, 256), (262144, 8, 1024), (5, 2, 4), (1, 4, 16)):
        seen = torch.cat([shard_indices(n, r, world, tile) for r in range(world)])
        assert seen.sort().values.tolist() == list(range(n))
        assert max(shard_indices(n, r, world, tile).numel() for r in range(world)) <= shard_capacity(n, world, tile)
    # interleaving balances a centre-heavy image: per-rank share of the central rays is even
    n, world = 512 * 512, 8
    centre = torch.zeros(n, dtype=torch.bool)
    centre[n // 3: 2 * n // 3] = True
    shares = [centre[shard_indices(n, r, world)].float().sum().item() for r in range(world)]
    assert max(shares) / min(shares) < 1.2


def test_peer_frame_index_formula_matches_shard_indices():
    """csrc/nvr_frame.cuh frame_index(): local ray i of a rank's shard -> ((i / tile) * world + rank) * tile + i % tile must be
    exactly the i-th entry of shard_indices (what k_resolve_rays uses to store a ray into the peers' frames)."""
    for n, world, tile in ((10000, 3, 256), (262144, 8, 1024), (236544, 8, 1024), (5, 2, 4), (1, 4, 16), (1000, 1, 64)):
        for rank in range(world):
            idx = shard_indices(n, rank, world, tile)
            i = torch.arange(idx.numel())
            gi = ((i // tile) * world + rank) * tile + i % tile
            assert torch.equal(gi, idx), (n, world, tile, rank)
            # a shard never holds a local index whose frame index is past the frame except in its last (partial) tile
            cap = shard_capacity(n, world, tile)
            j = torch.arange(cap)
            gj = ((j // tile) * world + rank) * tile + j % tile
            assert int((gj < n).sum()) == idx.numel()


def _fake_render(o, d, near, far):
    # any per-ray function: stands in for the CUDA render (no GPU in the CPU suite)
    rgb = torch.stack([o[:, 0] + near, d[:, 1] * far, o[:, 2] - d[:, 0]], dim=1)
    return rgb, near * far


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    o, d = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    near, far = torch.rand(n, generator=g), torch.rand(n, generator=g) + 1
    rgb, acc = render_sharded(_fake_render, o, d, near, far, rank, world, tile=64)
    ref_rgb, ref_acc = _fake_render(o, d, near, far)
    ok = torch.equal(rgb, ref_rgb) and torch.equal(acc, ref_acc)
    gathered = [None] * world
    dist.all_gather_object(gathered, ok)
    if rank == 0:
        q.put(all(gathered))
    dist.destroy_process_group()


def test_world2_gloo_assembles_identical_frame():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1000, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_world1_assemble_is_identity():
    x = torch.arange(30.0).view(10, 3)
    idx = shard_indices(10, 0, 1, 4)
    assert torch.equal(assemble(x[idx], 10, 0, 1, 4), x)
