"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE'S OWN CODE.

Run once in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

It imports the unmodified reference modules (zju3dv/instant-nvr @ a6f4d68) under the stub
harness of SURVEY.md Appendix C -- package stubs for colored_traceback / termcolor /
tensorboardX / matplotlib, a brute-force torch stand-in for pytorch3d.ops.knn.knn_points
(exact K nearest by squared L2 among the first lengths2[b] points), and `.cuda()` neutralised
because the container has no GPU -- builds ``make_network(cfg)`` / ``make_renderer(cfg, net)``
for configs/inb/inb_377.yaml with the hash tables capped at 2**16 rows (CLI overrides the
reference's own yacs accepts), overwrites the weights from the seeded stream in
``instant_nvr_b200.synthetic.fill_weights`` and records the reference's outputs.

Fixtures (float32 / int64 npz, inputs are regenerated from the seeds by the tests):
  e2e_gain1.npz, e2e_gain200.npz   Renderer.render(batch) on 32x32 rays x 32 samples
  state_dict_keys.json              name -> [shape, dtype] of the reference net.state_dict()
  stages.npz                        Embedder / Deformer / grid_sample / LBS / KNN-weights /
                                    PosEnc / volume_rendering called directly
"""
import collections
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
LOG2_T_CAP = 16
N_SAMPLES = 32
IMG = 32
SEED = 7


def install_stubs():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    stub("colored_traceback")
    stub("colored_traceback.auto")
    stub("termcolor", colored=lambda s, *a, **k: s, cprint=lambda *a, **k: None)

    class _Writer:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, n):
            return lambda *a, **k: None

    stub("tensorboardX", SummaryWriter=_Writer)
    stub("matplotlib")
    stub("matplotlib.pyplot")
    KNN = collections.namedtuple("KNN", "dists idx knn")

    def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, return_nn=False, return_sorted=True, **kw):
        outs_d, outs_i = [], []
        for b in range(p1.shape[0]):
            n2 = int(lengths2[b]) if lengths2 is not None else p2.shape[1]
            ds, ix = [], []
            for s in range(0, p1.shape[1], 2048):
                d = ((p1[b, s:s + 2048, None] - p2[b, None, :n2]) ** 2).sum(-1)
                dd, ii = d.topk(K, dim=-1, largest=False)
                ds.append(dd)
                ix.append(ii)
            outs_d.append(torch.cat(ds) if ds else p1.new_zeros(0, K))
            outs_i.append(torch.cat(ix) if ix else torch.zeros(0, K, dtype=torch.long))
        return KNN(torch.stack(outs_d), torch.stack(outs_i), None)

    stub("pytorch3d")
    stub("pytorch3d.ops")
    stub("pytorch3d.ops.knn", knn_points=knn_points)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        _tensor = torch.tensor
        torch.tensor = lambda *a, **k: _tensor(*a, **{kk: v for kk, v in k.items()
                                                      if not (kk == "device" and v == "cuda")})


def main():
    sys.path.insert(0, REPO)
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays

    install_stubs()
    os.chdir(REF)
    sys.path.insert(0, REF)
    opts = ["N_samples", str(N_SAMPLES), "perturb", "0", "silent", "True"]
    for part, log2 in (("body", 20), ("leg", 20), ("head", 18), ("larm", 15), ("rarm", 15)):
        opts += [f"partnet.{part}.embedder.kwargs.log2_hashmap_size", str(min(log2, LOG2_T_CAP))]
    sys.argv = ["x", "--cfg_file", "configs/inb/inb_377.yaml"] + opts
    from lib.config import cfg
    from lib.networks import make_network
    from lib.networks.renderer.make_renderer import make_renderer
    from lib.utils import blend_utils
    from lib.utils.net_utils import volume_rendering

    torch.manual_seed(0)
    net = make_network(cfg)
    net.eval()
    renderer = make_renderer(cfg, net)
    sd = net.state_dict()

    frame = make_frame(seed=SEED)
    rays = make_rays(frame, IMG, IMG)
    batch = dict(frame)
    batch.update(rays)

    for gain in (1.0, 200.0):
        fill_weights(sd, seed=SEED, table_gain=gain, bounds=frame["bounds"][0])
        with torch.no_grad():
            ret = renderer.render(dict(batch))
        out = {k: ret[k].numpy() for k in ("rgb_map", "acc_map", "raw", "occ")}
        surv = int((ret["raw"][0].abs().sum(-1) > 0).sum())
        print(f"gain {gain}: raw {out['raw'].shape}, non-zero samples {surv}, "
              f"acc mean {out['acc_map'].mean():.4f}")
        np.savez_compressed(os.path.join(HERE, f"e2e_gain{int(gain)}.npz"), **out)

    # ---------------- per-stage goldens (weights: gain 200 still loaded) -------------------
    g = torch.Generator().manual_seed(SEED)
    st = {}
    # (1) part Embedder.forward: body + larm, points inside and outside the bbox
    for pid in (0, 3):
        emb = net.tpose_human.part_networks[pid].embedder
        lo, hi = emb.bounds[0], emb.bounds[1]
        x = lo + (hi - lo) * (torch.rand(384, 3, generator=g) * 1.3 - 0.15)
        with torch.no_grad():
            st[f"embed{pid}_x"] = x.numpy()
            st[f"embed{pid}_out"] = emb(x, {}).numpy()
    # (2) deformer grid (concat mode) + Deformer.forward with a flag mask
    with torch.no_grad():
        uvt = torch.rand(256, 3, generator=g) * 1.2 - 0.1
        st["defgrid_x"] = uvt.numpy()
        st["defgrid_out"] = net.tpose_deformer.embedder(uvt, {}).numpy()
        tb = frame["tbounds"][0]
        x0 = tb[0] + (tb[1] - tb[0]) * (torch.rand(1, 300, 3, generator=g) * 1.2 - 0.1)
        flag = torch.rand(1, 300, generator=g) < 0.7
        st["deform_x"] = x0[0].numpy()
        st["deform_flag"] = flag[0].numpy()
        st["deform_out"] = net.tpose_deformer(x0, batch, flag=flag)[0].numpy()
    # (3) trilinear volume lookups (grid_sample border/align_corners)
    with torch.no_grad():
        pb = frame["pbounds"][0]
        pp = pb[0] + (pb[1] - pb[0]) * (torch.rand(1, 500, 3, generator=g) * 1.4 - 0.2)
        st["pnorm_x"] = pp[0].numpy()
        st["pnorm_out"] = blend_utils.pts_sample_blend_weights(pp, frame["pbw"][..., -1:], frame["pbounds"])[0, 0].numpy()
        st["uv_x"] = x0[0].numpy()
        st["uv_out"] = blend_utils.pts_sample_uv(x0, frame["tuv"], frame["tbounds"])[0].T.contiguous().numpy()
    # (4) KNN blend weights (through the stubbed knn_points) and the LBS chain
    with torch.no_grad():
        q = frame["ppts"][0][torch.randperm(6890, generator=g)[:200]] + 0.03 * torch.randn(200, 3, generator=g)
        q = torch.cat([q, pb[0] + (pb[1] - pb[0]) * torch.rand(56, 3, generator=g)])[None]
        mbw = blend_utils.pts_knn_blend_weights_multiassign_batch(q, frame["part_pts"][0], frame["part_pbw"][0],
                                                                  frame["lengths2"][0])       # (1,N,P,25)
        st["knn_x"] = q[0].numpy()
        st["knn_out"] = mbw[0].numpy()
        bw = mbw[0, :, :, :24].reshape(1, -1, 24).permute(0, 2, 1)
        A_bw, R_inv = blend_utils.get_inverse_blend_params(bw, frame["A"])
        big = blend_utils.get_blend_params(bw, frame["big_A"])
        pe = q[:, :, None].expand(1, 256, 5, 3).reshape(1, -1, 3)
        dirs = torch.nn.functional.normalize(torch.randn(1, 256 * 5, 3, generator=g), dim=-1)
        t = blend_utils.pose_points_to_tpose_points(pe, A_bw=A_bw, R_inv=R_inv)
        st["lbs_dirs"] = dirs[0].numpy()
        st["lbs_Rinv"] = R_inv[0].numpy()
        st["lbs_big"] = blend_utils.tpose_points_to_pose_points(t, A_bw=big)[0].numpy()
        td = blend_utils.pose_dirs_to_tpose_dirs(dirs, A_bw=A_bw, R_inv=R_inv)
        st["lbs_bigdirs"] = blend_utils.tpose_dirs_to_pose_dirs(td, A_bw=big)[0].numpy()
    # (5) view-direction encoding and compositing
    with torch.no_grad():
        vd = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1)
        st["posenc_x"] = vd.numpy()
        st["posenc_out"] = net.tpose_human.part_networks[0].embedder_dir(vd, {}).numpy()
        rawc = torch.rand(1, 40, 24, 4, generator=g)
        rawc[..., 3] = rawc[..., 3] * (torch.rand(1, 40, 24, generator=g) < 0.5)
        w, rgb_map, acc = volume_rendering(rawc[..., :3], rawc[..., 3], cfg.random_bg)
        st["comp_raw"] = rawc[0].numpy()
        st["comp_w"], st["comp_rgb"], st["comp_acc"] = w[0].numpy(), rgb_map[0].numpy(), acc[0].numpy()
    np.savez_compressed(os.path.join(HERE, "stages.npz"), **st)
    import json
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump({k: [list(v.shape), str(v.dtype)] for k, v in sd.items()}, f, indent=0)
    meta = dict(seed=SEED, log2_T_cap=LOG2_T_CAP, n_samples=N_SAMPLES, img=IMG, smpl_thresh=float(cfg.smpl_thresh),
                reference_commit="a6f4d68", torch=torch.__version__, numpy=np.__version__)
    with open(os.path.join(HERE, "META.txt"), "w") as f:
        for k, v in meta.items():
            f.write(f"{k}: {v}\n")
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
