"""The TRAINING branch of the oracle (oracle/nvr_oracle.py render_train + torch autograd) against the reference's own
training-mode forward and its own autograd (tests/golden/train.npz, made by tests/golden/make_golden_train.py by running
the unmodified reference): pins SURVEY.md section 8(a) row 16.  The GPU training tests then compare the CUDA backward with
this oracle's autograd.  Runs without a GPU."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REPO

sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, GOLDEN)
import nvr_oracle as O  # noqa: E402
import make_golden_train as T  # noqa: E402  (constants, the loss and the table probes; main() is not run)

GOLD = np.load(os.path.join(GOLDEN, "train.npz"))


@pytest.fixture(scope="module")
def run():
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays
    from instant_nvr_b200.training import trainable
    cfg = PathConfig.inb_377(N_samples=T.N_SAMPLES, log2_T_cap=T.LOG2_T_CAP)
    frame = make_frame(seed=T.SEED)
    rays = make_rays(frame, T.IMG, T.IMG)
    net = Network(cfg, device="cpu")
    fill_weights(net.state_dict(), seed=T.SEED, table_gain=200.0, bounds=frame["bounds"][0])
    names = [n for n, _ in trainable(net)]
    sd = {k: (v.detach().clone().requires_grad_(True) if k in names else v.detach().clone()) for k, v in net.state_dict().items()}
    ret = O.render_train(sd, {**frame, **rays}, cfg.N_samples, cfg.smpl_thresh, use_pair_reg=True, use_reg_distortion=True,
                         pair_noise=torch.from_numpy(GOLD["pair_noise"]))
    ret["resd"] = ret["resd"].reshape(1, -1, 3)                  # Renderer.render's view (inb_renderer.py:134-136)
    loss = T.the_loss(ret, rays["ray_o"].shape[1])
    loss.backward()
    return dict(ret=ret, loss=loss, sd=sd, names=names)


def test_training_forward_outputs_match_reference(run):
    ret = run["ret"]
    for k in ("rgb_map", "acc_map", "raw", "occ", "resd", "tpts", "tocc", "oresd", "reg_distortion_loss"):
        ref = torch.from_numpy(GOLD["out_" + k])
        got = ret[k].detach()
        assert tuple(got.shape) == tuple(ref.shape), (k, tuple(got.shape), tuple(ref.shape))
        err = (got - ref).abs().max().item() if ref.numel() else 0.0
        assert err <= 5e-6 * max(1.0, ref.abs().max().item() if ref.numel() else 1.0), (k, err)
    assert GOLD["out_oresd"].shape[1] > 0                          # the pair regulariser is exercised
    assert abs(run["loss"].item() - float(GOLD["loss"])) <= 2e-5 * abs(float(GOLD["loss"]))


def test_training_gradients_match_reference_autograd(run):
    """Every trainable tensor: MLP weights / biases, latent codes and the deformer's tensors element by element; the dense
    and hashed tables through their L2 norm, non-zero row count and a random projection."""
    sd = run["sd"]
    checked_small = checked_tab = 0
    for name in run["names"]:
        g = sd[name].grad
        g = torch.zeros_like(sd[name]) if g is None else g
        if "grad_" + name in GOLD.files:
            ref = torch.from_numpy(GOLD["grad_" + name])
            scale = max(ref.abs().max().item(), 1e-6)
            assert (g - ref).abs().max().item() <= 2e-4 * scale, (name, (g - ref).abs().max().item(), scale)
            checked_small += 1
        else:
            norm, rows, proj = GOLD["tab_" + name]
            g2 = g.reshape(-1, g.shape[-1])
            assert float((g2.abs().sum(-1) > 0).sum()) == rows, name
            assert abs(g.double().norm().item() - norm) <= 1e-4 * max(norm, 1e-9), name
            p = (g.double().reshape(-1) * T.table_probe(name, g.numel()).double()).sum().item()
            assert abs(p - proj) <= 1e-4 * max(norm, 1e-9) * 3.0, (name, p, proj)
            checked_tab += 1
    assert checked_tab == 12 and checked_small == 55
    nz = sum(1 for n in run["names"] if sd[n].grad is not None and sd[n].grad.abs().sum() > 0)
    assert nz >= 60                                                # every part network and the deformer receive gradient
