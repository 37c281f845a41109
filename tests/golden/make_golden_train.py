"""Golden fixture for the TRAINING branch: forward outputs and GRADIENTS of the reference's
own code (zju3dv/instant-nvr @ a6f4d68), imported unmodified under the stub harness of make_golden.py.

    python tests/golden/make_golden_train.py      ->  tests/golden/train.npz

What runs: ``make_network(cfg)`` / ``make_renderer(cfg, net)`` for configs/inb/inb_377.yaml (hash tables capped at 2**12
rows, N_samples 24, perturb 0 -- CLI overrides the reference's yacs accepts; use_pair_reg and use_reg_distortion as the
config ships them: True), weights from ``instant_nvr_b200.synthetic.fill_weights``, ``net.train()``, ``iter_step = 2``,
``renderer.render(batch)`` on 20x20 rays, then a loss that touches every differentiable output

    L = sum(rgb_map * W1) + 0.5 sum(acc_map) + 3 sum(resd * W2) + sum(tocc * W3) + 0.1 mean(reg_distortion_loss)
        + 10 sum(oresd ** 2)

(W from a seeded generator) and ``L.backward()`` through the reference's autograd.  Stored: the forward outputs, the
``torch.rand_like`` draw of the pair regulariser (the only random numbers with perturb 0; re-drawn from the same seed),
every non-table gradient in full, and for the dense / hash tables (up to 22 M floats) the L2 norm, the number of non-zero
rows and the dot product with a seeded random vector.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
LOG2_T_CAP = 12
N_SAMPLES = 24
IMG = 20
SEED = 4
RNG_SEED = 123


def loss_weights(n_rays, m5):
    g = torch.Generator().manual_seed(77)
    return torch.randn(n_rays, 3, generator=g), torch.randn(m5, 3, generator=g), torch.randn(m5, generator=g)


def the_loss(ret, n_rays):
    """`ret`: what Renderer.render returns in training mode (resd (1,5N',3): inb_renderer.py:134-136)."""
    m5 = ret["resd"].shape[1]
    W1, W2, W3 = loss_weights(n_rays, m5)
    loss = (ret["rgb_map"][0] * W1).sum() + 0.5 * ret["acc_map"].sum() + 3.0 * (ret["resd"][0] * W2).sum() \
        + (ret["tocc"].reshape(-1) * W3).sum() + 0.1 * ret["reg_distortion_loss"].mean()
    if ret["oresd"].numel():
        loss = loss + 10.0 * (ret["oresd"] ** 2).sum()
    return loss


def table_probe(name, numel):
    g = torch.Generator().manual_seed(sum(name.encode()) + 1000)
    return torch.randn(numel, generator=g)


def main():
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    from make_golden import install_stubs
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

    assert cfg.use_pair_reg and cfg.use_reg_distortion and cfg.perturb == 0
    torch.manual_seed(0)
    net = make_network(cfg)
    renderer = make_renderer(cfg, net)
    frame = make_frame(seed=SEED)
    rays = make_rays(frame, IMG, IMG)
    fill_weights(net.state_dict(), seed=SEED, table_gain=200.0, bounds=frame["bounds"][0])
    batch = dict(frame)
    batch.update(rays)
    batch["iter_step"] = 2                                  # not 1: the bbox overwrite of part_base_embedder.py:107-109 stays off
    net.train()
    torch.manual_seed(RNG_SEED)
    ret = renderer.render(batch)
    n_rays = rays["ray_o"].shape[1]
    loss = the_loss(ret, n_rays)
    loss.backward()

    st = {"loss": np.array(loss.item(), dtype=np.float64)}
    for k in ("rgb_map", "acc_map", "raw", "occ", "resd", "tpts", "tocc", "oresd", "reg_distortion_loss"):
        st["out_" + k] = ret[k].detach().numpy()
    k_pair = ret["oresd"].shape[1] // 2
    torch.manual_seed(RNG_SEED)
    st["pair_noise"] = torch.rand(1, k_pair, 3).numpy()[0]  # the same draw torch.rand_like made inside the render
    n_tab = 0
    for name, p in net.named_parameters():
        if not p.requires_grad:
            continue
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        if name.endswith(".dense") or name.endswith(".hash"):
            g2 = g.reshape(-1, g.shape[-1])
            st["tab_" + name] = np.array([g.double().norm().item(), float((g2.abs().sum(-1) > 0).sum()),
                                          (g.double().reshape(-1) * table_probe(name, g.numel()).double()).sum().item()])
            n_tab += 1
        else:
            st["grad_" + name] = g.numpy()
    np.savez_compressed(os.path.join(HERE, "train.npz"), **st)
    print(f"loss {loss.item():.6f}; (sample, part) rows {ret["resd"].shape[1]}, pair-regulariser points {k_pair}, "
          f"{n_tab} tables, {sum(1 for k in st if k.startswith('grad_'))} small gradients; "
          f"{os.path.getsize(os.path.join(HERE, 'train.npz'))} bytes")


if __name__ == "__main__":
    main()
