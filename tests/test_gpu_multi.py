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
    rays = make_rays(frame, 208, 208)      # 1.38 M samples: each rank's shard is >= 2^19 samples, i.e. a two-lane render
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
    # the same frame assembled over NVLink peer memory (nvr_render_rays_frame: the compositing kernel stores into every rank's
    # frame buffer + one flag barrier), three frames in a row (the two slots alternate), then the unfused scatter form
    from instant_nvr_b200.sharding import PeerFrame, shard_indices
    n = o.shape[0]
    idx = shard_indices(n, rank, world, tile=256).cuda()
    pf = PeerFrame(eng, n, rank, world, tile=256)
    ok = ok and idx.numel() * cfg.N_samples >= 1 << 19
    for it in range(3):
        frame_t, l_rgb, l_acc = pf.render(o[idx], d[idx], nr[idx], fr[idx], cfg.N_samples, want_local=True)
        torch.cuda.synchronize()
        ok = ok and eng.counters()["n_passes"] == 2     # the compositing kernels of BOTH lanes stored into the peers' frames
        ok = ok and bool(torch.equal(frame_t[:, :3], ref_rgb) and torch.equal(frame_t[:, 3], ref_acc))
        ok = ok and bool(torch.equal(l_rgb, ref_rgb[idx]) and torch.equal(l_acc, ref_acc[idx]))
    # host-buffer form (nvr_render_rays_frame_host): the shard's rays come from pinned memory on the lanes' streams, its own
    # pixels go back the same way
    hr = [t[idx].cpu().contiguous().pin_memory() for t in (o, d, nr, fr)]
    h_rgb, h_acc = torch.empty(idx.numel(), 3).pin_memory(), torch.empty(idx.numel()).pin_memory()
    frame_t = pf.render_host(*hr, cfg.N_samples, h_rgb, h_acc)
    ok = ok and bool(torch.equal(frame_t[:, :3], ref_rgb) and torch.equal(frame_t[:, 3], ref_acc))
    ok = ok and bool(torch.equal(h_rgb, ref_rgb[idx].cpu()) and torch.equal(h_acc, ref_acc[idx].cpu()))
    frame_t = pf.allgather(ref_rgb[idx].contiguous(), ref_acc[idx].contiguous())
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(frame_t[:, :3], ref_rgb) and torch.equal(frame_t[:, 3], ref_acc))
    pf.close()
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


def _train_worker(rank, world, port, q):
    """Each rank back-propagates its half of the rays; the all-reduced gradients must equal the single-GPU gradient of
    the whole batch (the loss is a sum over rays) -- what DDP's gradient all-reduce relies on."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.renderer import Renderer
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays
    cfg = PathConfig.inb_377(N_samples=16, log2_T_cap=12).with_(use_pair_reg=False)
    frame = make_frame(seed=5)
    rays = make_rays(frame, 24, 24)
    net = Network(cfg, device="cpu")
    fill_weights(net.state_dict(), seed=5, table_gain=100.0, bounds=frame["bounds"][0])
    net = net.cuda().train()
    gb = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in {**frame, **rays}.items()}
    R = gb["ray_o"].shape[1]
    tgt = torch.rand(1, R, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    r = Renderer(net)

    def grads_for(sel):
        for p in net.parameters():
            p.grad = None
        b = dict(gb)
        for k in ("ray_o", "ray_d", "near", "far", "occupancy"):
            b[k] = gb[k][:, sel]
        ret = r.render(b)
        (((ret["rgb_map"] - tgt[:, sel]) ** 2).sum() + ret["resd"].pow(2).sum()).backward()
        return {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in net.named_parameters() if p.requires_grad}
    full = grads_for(torch.arange(R, device="cuda"))
    mine = grads_for(torch.arange(rank, R, world, device="cuda"))
    worst = 0.0
    for n, g in mine.items():
        dist.all_reduce(g)
        worst = max(worst, (g - full[n]).abs().max().item() / (full[n].abs().max().item() + 1e-12))
    flags = [None] * world
    dist.all_gather_object(flags, worst)
    if rank == 0:
        q.put(max(flags))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    worst = q.get(timeout=5)
    print("[multi] worst relative difference of all-reduced gradients:", worst)
    assert worst < 1e-3
