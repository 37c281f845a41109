#!/usr/bin/env python
"""Rank skew of the interleaved ray-tile sharding, measured on ONE GPU: every rank's shard of the configs[1] frame is
rendered on its own (per-frame preparation included, CUDA events, 10 steps after 3 warm-ups) for several tile sizes and
world sizes.  An N-GPU step cannot be faster than the slowest shard, so max/mean over the ranks is the efficiency the
sharding itself gives away before any launch or barrier cost.  usage: python tools/shard_skew.py [out.jsonl] [config]"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import __graft_entry__ as ge  # noqa: E402


def main():
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    cfgd = bench.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "c2"]
    ge.build()
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.sharding import shard_indices
    from instant_nvr_b200.synthetic import make_frame, make_rays
    torch.cuda.set_device(0)
    cfg = PathConfig.inb_377(N_samples=cfgd["S"])
    frame = make_frame(seed=0)
    with torch.device("cuda"):
        net = Network(cfg)
    net = net.cuda().eval()
    bench.device_weights(net, frame)
    gframe = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in frame.items()}
    eng = net.engine()
    eng.bind_frame(gframe)
    rays = make_rays(frame, cfgd["H"], cfgd["W"], drop_missing=True)
    rays = {k: rays[k][0].cuda().contiguous() for k in ("ray_o", "ray_d", "near", "far")}
    n = rays["ray_o"].shape[0]

    def time_shard(idx, steps=10, warmup=3):
        d = {k: v[idx].contiguous() for k, v in rays.items()}

        def step():
            eng.bind_frame(gframe, force=True)
            eng.render_rays(d["ray_o"], d["ray_d"], d["near"], d["far"], cfgd["S"])
        for _ in range(warmup):
            step()
        c = eng.counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, c["n_survivors"], sum(c["n_pairs"])      # n_pairs = pairs EVALUATED (far-field pairs are separate)

    full_ms, full_surv, full_pairs = time_shard(torch.arange(n, device="cuda"))
    out.write(json.dumps({"world": 1, "rays": n, "ms": full_ms, "survivors": full_surv, "evaluated_pairs": full_pairs}) + "\n")
    for world in (2, 4, 8):
        for tile in (1024, 256, 64):
            ms, surv, pairs = [], [], []
            for r in range(world):
                t, s, p = time_shard(shard_indices(n, r, world, tile).cuda())
                ms.append(t), surv.append(s), pairs.append(p)
            mean = sum(ms) / world
            out.write(json.dumps({"world": world, "tile": tile, "ms_max": max(ms), "ms_mean": mean, "ms_min": min(ms),
                                  "skew_max_over_mean": max(ms) / mean, "speedup_bound": full_ms / max(ms),
                                  "pairs_max_over_mean": max(pairs) / (sum(pairs) / world),
                                  "survivors_max_over_mean": max(surv) / (sum(surv) / world), "ms": ms}) + "\n")
            out.flush()


if __name__ == "__main__":
    main()
