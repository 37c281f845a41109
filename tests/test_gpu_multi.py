"""N>1 on real GPUs: two ranks (NCCL) render interleaved ray shards of one view through the C-ABI and assemble
the frame with one all-gather; the result must equal the single-GPU render bit for bit (rays are independent).
Skipped on boxes with fewer than two GPUs."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.sharding import render_sharded
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays
    cfg = PathConfig.inb_377(N_samples=32, log2_T_cap=14)
    frame = make_frame(seed=2)
    rays = make_rays(frame, 96, 96)
    net = Network(cfg, device="cpu")
    fill_weights(net.state_dict(), seed=2, table_gain=100.0, bounds=frame["bounds"][0])
    net = net.cuda().eval()
    gframe = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in frame.items()}
    eng = net.engine()
    eng.bind_frame(gframe)
    o, d, nr, fr = (rays[k][0].cuda() for k in ("ray_o", "ray_d", "near", "far"))
    fn = lambda a, b, c, e: eng.render_rays(a, b, c, e, cfg.N_samples)
    rgb, acc = render_sharded(fn, o, d, nr, fr, rank, world, tile=256)
    ref_rgb, ref_acc = fn(o, d, nr, fr)
    ok = bool(torch.equal(rgb, ref_rgb) and torch.equal(acc, ref_acc) and ref_acc.max().item() > 0)
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        q.put(all(flags))
    dist.destroy_process_group()


def test_two_rank_nccl_render_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
