"""Runs INSIDE the reference tree (cwd = /root/reference, package stubs as in tests/golden/make_golden.py) and checks the
reference-facing surface of the drop-ins against the reference's own objects -- everything that does not need a GPU:

  1. PathConfig.from_reference_cfg(lib.config.cfg) -- for inb_377.yaml as shipped and with CLI overrides
  2. instant_nvr_b200.network.Network() (no-argument constructor, reads the global cfg, the way
     lib/networks/make_network.py:5-8 calls it): state_dict names / shapes / dtypes == make_network(cfg).state_dict(),
     load_state_dict(strict=True) both ways, the frozen buffers equal bit for bit
  3. instant_nvr_b200.optimizer.make_optimizer(cfg, net): parameter groups (order, lr, weight decay, eps, betas) ==
     lib.train.optimizer.make_optimizer(cfg, reference_net)
  4. the module strings resolve the way the reference resolves them (importlib + attribute names)

Prints one JSON line; tests/test_reference_dropin.py runs it in a subprocess."""
import importlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def main():
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    from make_golden import install_stubs
    install_stubs()
    os.chdir(REF)
    sys.path.insert(0, REF)
    cap = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    opts = ["N_samples", "40", "silent", "True", "smpl_thresh", "0.1"]
    for part, log2 in (("body", 20), ("leg", 20), ("head", 18), ("larm", 15), ("rarm", 15)):
        opts += [f"partnet.{part}.embedder.kwargs.log2_hashmap_size", str(min(log2, cap))]
    sys.argv = ["x", "--cfg_file", "configs/inb/inb_377.yaml"] + opts
    import torch
    from lib.config import cfg
    from lib.networks import make_network
    from lib.train.optimizer import make_optimizer as ref_make_optimizer
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.optimizer import FusedAdam, make_optimizer

    out = {}
    # 1. config mapping
    pc = PathConfig.from_reference_cfg(cfg)
    # (the inb_377() factory keeps eval-oriented defaults -- perturb 0, no distortion term -- for the parity tests; the
    #  yaml trains with perturb 1 and use_reg_distortion, which the live mapping must pick up)
    want = PathConfig.inb_377(N_samples=40, log2_T_cap=cap).with_(smpl_thresh=0.1, perturb=1.0, use_reg_distortion=True)
    diffs = [f for f in pc.__dataclass_fields__ if getattr(pc, f) != getattr(want, f)]
    out["config_diffs"] = diffs
    out["config"] = {"N_samples": pc.N_samples, "smpl_thresh": pc.smpl_thresh, "use_pair_reg": pc.use_pair_reg,
                     "use_reg_distortion": pc.use_reg_distortion, "perturb": pc.perturb}
    # 2. networks
    torch.manual_seed(0)
    ref_net = make_network(cfg)
    mod = importlib.import_module("instant_nvr_b200.network")       # what make_network does with cfg.network_module
    ours = mod.Network()
    rs, os_ = ref_net.state_dict(), ours.state_dict()
    out["keys_equal"] = list(rs.keys()) == list(os_.keys())
    out["shape_mismatch"] = [k for k in rs if k in os_ and (tuple(rs[k].shape) != tuple(os_[k].shape) or rs[k].dtype != os_[k].dtype)]
    trainable = {n for n, p in ref_net.named_parameters() if p.requires_grad}
    frozen = [k for k in rs if k not in trainable]                  # bounds, level tables, offsets, freq_bands ...
    out["frozen_buffers"] = len(frozen)
    out["frozen_mismatch"] = [k for k in frozen if not torch.equal(rs[k], os_[k])]
    ours.load_state_dict(rs, strict=True)                           # a reference checkpoint loads into the drop-in
    ref_net.load_state_dict(ours.state_dict(), strict=True)         # and the other way round
    out["param_names_equal"] = [n for n, _ in ref_net.named_parameters()] == [n for n, _ in ours.named_parameters()]
    out["requires_grad_equal"] = [p.requires_grad for p in ref_net.parameters()] == [p.requires_grad for p in ours.parameters()]
    # 3. optimizers
    ro = ref_make_optimizer(cfg, ref_net)
    oo = make_optimizer(cfg, ours)
    out["optimizer_is_fused"] = isinstance(oo, FusedAdam)
    keys = ("lr", "weight_decay", "eps", "betas")
    out["optimizer_groups"] = len(ro.param_groups)
    out["optimizer_mismatch"] = [i for i, (a, b) in enumerate(zip(ro.param_groups, oo.param_groups))
                                 if any(tuple(a[k]) != tuple(b[k]) if k == "betas" else a[k] != b[k] for k in keys)
                                 or tuple(a["params"][0].shape) != tuple(b["params"][0].shape)]
    out["optimizer_len_equal"] = len(ro.param_groups) == len(oo.param_groups)
    # 4. renderer / trainer module strings
    rmod = importlib.import_module("instant_nvr_b200.renderer")
    r = rmod.Renderer(ours)                                         # make_renderer: imp.load_source(...).Renderer(network)
    out["renderer_has_render"] = callable(getattr(r, "render", None))
    # 5. trainer shim: the reference's NetworkWrapper with the renderer swapped (lib/train/trainers/make_trainer.py:4-12)
    try:
        import types
        for name in ("trimesh", "lpips", "skimage", "skimage.metrics", "imageio", "mcubes", "plyfile", "ipdb"):   # not installed here
            if name not in sys.modules:
                try:
                    importlib.import_module(name)
                except Exception:
                    sys.modules[name] = types.ModuleType(name)
        if "turtle" not in sys.modules:                            # stray `from turtle import forward` in fourier_loss.py:1 (needs tkinter)
            sys.modules["turtle"] = types.ModuleType("turtle")
            sys.modules["turtle"].forward = None
        cfg.use_lpips = False                                      # PerceptualLoss() would download VGG weights
        tmod = importlib.import_module("instant_nvr_b200.trainer")
        wrap = tmod.NetworkWrapper(ours)
        from lib.train.trainers import inb_trainer
        out["trainer"] = {"is_reference_subclass": isinstance(wrap, inb_trainer.NetworkWrapper),
                          "renderer_swapped": type(wrap.renderer).__module__ == "instant_nvr_b200.renderer",
                          "net_shared": wrap.net is ours}
    except Exception as e:                                          # missing third-party packages of the reference's trainer
        out["trainer"] = {"unavailable": f"{type(e).__name__}: {e}"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
