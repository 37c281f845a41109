"""Pin the CPU oracle (oracle/nvr_oracle.py) against outputs of the reference's own code.

The reference ships no tests or golden vectors for this path, so the fixtures in
tests/golden/ were produced by running its unmodified modules under the stub harness
(tests/golden/make_golden.py).  Tolerances: the oracle restates the same fp32 torch-CPU math, so
agreement is expected at the 1e-6 level; the only library-order freedom is sgemm / einsum
summation order.
"""
import os
import sys

import numpy as np
import torch

from conftest import REPO, load_golden

sys.path.insert(0, os.path.join(REPO, "oracle"))
import nvr_oracle as O  # noqa: E402


def _close(a, b, atol, rtol=0.0, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    lim = atol + rtol * np.abs(b)
    assert (err <= lim).all(), f"{what}: max err {err.max():.3e} (limit {lim.min():.1e}) at {np.argmax(err - lim)}"


def test_state_dict_layout_matches_reference(golden_setup):
    """Key names / shapes of our Network.state_dict() against the reference's (Appendix D), for
    the golden config.  The reference key list is stored with the fixtures."""
    import json
    from conftest import GOLDEN
    sd = golden_setup["sd"]
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as f:
        ref = json.load(f)
    assert sorted(sd.keys()) == sorted(ref.keys())
    for k, (shape, dtype) in ref.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k
    assert len(sd) == 114
    assert sd["tpose_deformer.embedder.dense"].shape == (12276, 2)
    assert sd["tpose_deformer.embedder.hash"].shape == (2, 16411, 2)
    assert sd["tpose_deformer.mlp.4.weight"].shape == (3, 32)
    assert sd["tpose_human.part_networks.0.rgb.linears.2.weight"].shape == (3, 64)
    assert "tpose_human.part_networks.1.rgb.linears.2.weight" not in sd
    assert sd["tpose_human.part_networks.3.embedder_dir.embedder.freq_bands"].shape == (4, 2, 1)
    assert sd["tpose_human.part_networks.0.embedder.entries_num"].dtype == torch.int64


def test_full_config_table_sizes():
    """(sum-dense-rows, H, T) for the shipped inb_377 config, as probed from the reference
    (SURVEY.md Appendix D)."""
    from instant_nvr_b200.config import PathConfig
    cfg = PathConfig.inb_377()
    want = {"body": (822944, 10, 1048583), "leg": (1385683, 3, 1048583), "head": (199799, 5, 262147),
            "larm": (28143, 7, 32771), "rarm": (28143, 7, 32771)}
    for p in cfg.parts:
        g = p.grid
        assert (g.dense_rows, g.n_hash_levels, g.T) == want[p.name], p.name
    d = cfg.deformer_grid
    assert (d.dense_rows, d.n_hash_levels, d.T) == (12276, 2, 16411)
    assert cfg.parts[0].grid.res == [16, 22, 30, 42, 58, 80, 110, 152, 210, 290, 400, 553, 763, 1053, 1453, 2005]
    assert cfg.parts[1].grid.res == [2, 2, 3, 5, 7, 10, 13, 19, 26, 36, 50, 69, 95, 131, 181, 250]


def test_stage_goldens(golden_setup):
    sd, frame = golden_setup["sd"], golden_setup["frame"]
    st = load_golden("stages.npz")
    t = lambda k: torch.from_numpy(st[k])
    # hash grids (sum mode, body + larm; concat mode, deformer)
    for pid in (0, 3):
        out = O.grid_embed(sd, f"tpose_human.part_networks.{pid}.embedder.", t(f"embed{pid}_x"), True)
        _close(out, st[f"embed{pid}_out"], 2e-6, 1e-6, f"embed{pid}")
    out = O.grid_embed(sd, "tpose_deformer.embedder.", t("defgrid_x"), False)
    _close(out, st["defgrid_out"], 1e-6, 1e-6, "deformer grid")
    # trilinear lookups
    out = O.sample_volume(frame["pbw"][0][..., -1:], t("pnorm_x"), frame["pbounds"][0])[:, 0]
    _close(out, st["pnorm_out"], 1e-6, 1e-6, "pnorm")
    out = O.sample_volume(frame["tuv"][0], t("uv_x"), frame["tbounds"][0])
    _close(out, st["uv_out"], 1e-6, 1e-6, "uv")
    # deformer with flag mask
    x0, flag = t("deform_x"), t("deform_flag")
    res = torch.zeros_like(x0)
    res[flag] = O.deformer(sd, x0[flag], frame["tuv"][0], frame["tbounds"][0], frame["frame_dim"])
    _close(res, st["deform_out"], 1e-6, 1e-5, "deformer")
    # KNN blend weights + LBS
    bw, pd = O.knn_blend_weights(t("knn_x"), frame["part_pts"][0], frame["part_pbw"][0], frame["lengths2"][0])
    _close(bw, st["knn_out"][..., :24], 1e-6, 1e-5, "knn bw")
    _close(pd, st["knn_out"][..., 24], 1e-6, 1e-5, "knn dist")
    gbw = t("knn_out")[..., :24].reshape(-1, 24)
    pe = t("knn_x")[:, None].expand(-1, 5, 3).reshape(-1, 3)
    big, bigd = O.lbs_to_bigpose(pe, t("lbs_dirs"), gbw, frame["A"][0], frame["big_A"][0])
    # far parts have ~1e-11 blend weights, so R_inv reaches 1e10 and fp32 summation order shows;
    # compare relative to the magnitude of the result
    _close(big, st["lbs_big"], 1e-5, 1e-4, "lbs big")
    _close(bigd, st["lbs_bigdirs"], 1e-5, 1e-4, "lbs big dirs")
    # posenc / compositing
    _close(O.posenc(t("posenc_x")), st["posenc_out"], 1e-6, 0, "posenc")
    w, rgb, acc = O.composite(t("comp_raw"))
    _close(w, st["comp_w"], 1e-7, 1e-6, "weights")
    _close(rgb, st["comp_rgb"], 1e-6, 1e-6, "rgb_map")
    _close(acc, st["comp_acc"], 1e-6, 1e-6, "acc_map")


def _e2e(golden_setup, gain):
    from instant_nvr_b200.synthetic import fill_weights
    sd, cfg, frame = golden_setup["sd"], golden_setup["cfg"], golden_setup["frame"]
    fill_weights(sd, seed=golden_setup["seed"], table_gain=gain, bounds=frame["bounds"][0])
    ret = O.render(sd, golden_setup["batch"], cfg.N_samples, cfg.smpl_thresh)
    fill_weights(sd, seed=golden_setup["seed"], table_gain=200.0, bounds=frame["bounds"][0])
    return ret


def test_e2e_golden_gain1(golden_setup):
    ret, gold = _e2e(golden_setup, 1.0), load_golden("e2e_gain1.npz")
    _close(ret["raw"], gold["raw"], 2e-6, 1e-5, "raw")
    _close(ret["occ"], gold["occ"], 2e-6, 1e-5, "occ")
    _close(ret["rgb_map"], gold["rgb_map"], 2e-6, 1e-5, "rgb_map")
    _close(ret["acc_map"], gold["acc_map"], 2e-6, 1e-5, "acc_map")


def test_e2e_golden_gain200(golden_setup):
    ret, gold = _e2e(golden_setup, 200.0), load_golden("e2e_gain200.npz")
    _close(ret["raw"], gold["raw"], 5e-6, 1e-5, "raw")
    _close(ret["rgb_map"], gold["rgb_map"], 5e-6, 1e-5, "rgb_map")
    _close(ret["acc_map"], gold["acc_map"], 5e-6, 1e-5, "acc_map")
    assert abs(O.psnr(ret["rgb_map"], torch.from_numpy(gold["rgb_map"]))) > 90.0
