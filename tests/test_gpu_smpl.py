"""Per-frame SMPL preprocessing on the GPU, through the C-ABI, against the
reference-generated golden vectors (tests/golden/smpl.npz) and the numpy oracle (oracle/smpl_oracle.py), then a render of
a frame prepared on the device against the oracle's render of the reference-prepared frame."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "oracle"))
import smpl_oracle as SO  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(REPO, "tests", "golden", "smpl.npz"))
SEEDS = (3, 11)


def f64(a):
    return np.asarray(a, dtype=np.float64)


@pytest.fixture(scope="module", params=SEEDS)
def case(request):
    from instant_nvr_b200.smpl_frame import SmplSubject, prepare_frame
    from instant_nvr_b200.synthetic import make_subject
    seed = request.param
    sub = make_subject(seed)
    g = {k[len(f"s{seed}_"):]: GOLD[k] for k in GOLD.files if k.startswith(f"s{seed}_")}
    subject = SmplSubject(sub["joints"], sub["parents"], sub["weights"], sub["tpose"])
    out = prepare_frame(subject, sub["wxyz"], sub["Rh"], sub["Th"], sub["poses"])
    torch.cuda.synchronize()
    return seed, sub, g, subject, {k: v[0].cpu().numpy() for k, v in out.items()}


def test_static_tables_bitwise(case):
    seed, sub, g, subject, out = case
    assert np.array_equal(subject.parts.astype(np.int8), g["parts"])
    assert np.array_equal(out["lengths2"], g["lengths2"])
    assert np.array_equal(out["bounds"], g["bounds"]) and np.array_equal(out["tbounds"], g["tbounds"])
    _, part_pbw, _, _ = SO.part_tables(g["ppts"], sub["tpose"], sub["weights"], g["parts"].astype(np.float64))
    assert np.array_equal(out["part_pbw"], part_pbw)


def test_transforms_match_reference(case):
    """A, big_A: float64 chain on the device, identical to the reference's float32 result up to one ulp (np.dot vs fused
    double sums); R = cv2.Rodrigues(Rh): <= 1 ulp."""
    seed, sub, g, subject, out = case
    for name in ("A", "big_A"):
        ref = g[name]
        assert np.abs(out[name] - ref).max() <= 1.2e-7 * max(1.0, np.abs(ref).max()), name
        assert (out[name] != ref).mean() < 0.05, name
    assert np.abs(out["R"] - g["R"]).max() <= 6e-8
    assert np.array_equal(out["Th"].reshape(-1), sub["Th"].reshape(-1))


def test_posed_vertices_and_bounds(case):
    """ppts = (wxyz - Th) . R in float32 (reference: BLAS sgemm; here fused multiply-adds): <= 3e-7 on ~1 m coordinates;
    the part tables and the bboxes are exactly consistent with the device's own ppts."""
    seed, sub, g, subject, out = case
    assert np.abs(out["ppts"] - g["ppts"]).max() <= 3e-7
    parts = subject.parts
    for pid in range(5):
        m = parts == pid
        n = int(m.sum())
        assert np.array_equal(out["part_pts"][pid, :n], out["ppts"][m])
        assert not out["part_pts"][pid, n:].any()
    assert np.array_equal(out["pbounds"], SO.get_bounds(out["ppts"]))
    assert np.array_equal(out["wbounds"], g["wbounds"])
    assert np.abs(out["pbounds"] - g["pbounds"]).max() <= 3e-7


def test_bweights_volume(case):
    """The volume of tools/prepare_zjumocap.get_bweights: same dims as the reference's; the distance channel within one
    float32 ulp of the reference's float64 result; the 24 weight channels are the rows of the same nearest vertex except
    where two vertices are equidistant to 1e-12."""
    seed, sub, g, subject, out = case
    pbw = out["pbw"]
    assert tuple(pbw.shape) == tuple(g["pbw_shape"])
    ref, vid = SO.get_bweights(f64(sub["wxyz"]), f64(sub["Rh"])[0], f64(sub["Th"]), sub["weights"])
    ulp = np.abs(pbw[..., 24].view(np.int32).astype(np.int64) - ref[..., 24].view(np.int32).astype(np.int64))
    assert ulp.max() <= 1 and (ulp > 0).mean() < 1e-3
    same = np.all(pbw[..., :24] == ref[..., :24], axis=-1)
    assert (~same).sum() <= 3
    if "pbw_dist" in g:
        ulp = np.abs(pbw[..., 24].view(np.int32).astype(np.int64) - g["pbw_dist"].view(np.int32).astype(np.int64))
        assert ulp.max() <= 1


def test_get_rigid_transformation_entry(case):
    from instant_nvr_b200.smpl_frame import get_rigid_transformation
    seed, sub, g, subject, out = case
    A = get_rigid_transformation(sub["poses"], sub["joints"], sub["parents"]).cpu().numpy()
    assert np.array_equal(A, out["A"])


def test_render_from_device_prepared_frame():
    """A frame prepared on the device feeds the render: result within the path's parity bar (1e-4 per-sample raw,
    PSNR > 60 dB) of the oracle's render of the same frame prepared by the reference's arithmetic (numpy oracle)."""
    import nvr_oracle as O
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.renderer import Renderer
    from instant_nvr_b200.smpl_frame import SmplSubject, prepare_frame
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays, make_subject

    seed = 1
    sub = make_subject(seed)
    base = make_frame(seed=seed)                                   # tuv volume, frame_dim, latent_index, reference big pose
    subject = SmplSubject(sub["joints"], sub["parents"], sub["weights"], sub["tpose"])
    dev_frame = prepare_frame(subject, sub["wxyz"], sub["Rh"], sub["Th"], sub["poses"], big_poses=sub["big_poses"])
    # the same frame by the oracle (reference arithmetic, numpy)
    wxyz, pxyz, A, big_A, R, Rh, Th = SO.prepare_input(f64(sub["wxyz"]), f64(sub["Rh"]), f64(sub["Th"]), f64(sub["poses"]),
                                                       sub["joints"], sub["parents"], big_poses=f64(sub["big_poses"]))
    parts = SO.smpl_parts(sub["weights"])
    part_pts, part_pbw, lengths2, bounds = SO.part_tables(pxyz, sub["tpose"], sub["weights"], parts)
    pbw, _ = SO.get_bweights(f64(sub["wxyz"]), f64(sub["Rh"])[0], f64(sub["Th"]), sub["weights"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))[None]
    ref_frame = dict(base)
    ref_frame.update({"R": t(R), "Th": t(Th), "A": t(A), "big_A": t(big_A), "ppts": t(pxyz), "pbounds": t(SO.get_bounds(pxyz)),
                      "wbounds": t(SO.get_bounds(wxyz)), "part_pts": t(part_pts), "part_pbw": t(part_pbw),
                      "lengths2": t(lengths2.astype(np.int64)), "bounds": t(bounds), "pbw": t(pbw)})
    rays = make_rays(ref_frame, 20, 20)
    cfg = PathConfig.inb_377(N_samples=16, log2_T_cap=12)
    net = Network(cfg, device="cpu")
    fill_weights(net.state_dict(), seed=seed, table_gain=1.0, bounds=ref_frame["bounds"][0])
    sd_cpu = {k: v.clone() for k, v in net.state_dict().items()}
    ref = O.render(sd_cpu, {**ref_frame, **rays}, cfg.N_samples, cfg.smpl_thresh)
    net = net.cuda().eval()
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in {**base, **rays}.items()}
    batch.update(dev_frame)
    out = Renderer(net, return_raw=True).render(batch)
    raw, rraw = out["raw"].cpu(), ref["raw"]
    n_active = int((rraw[0, :, 3] > 0).sum())
    assert n_active > 20
    # a sample whose cull / part-flag distance sits within float rounding of the threshold may flip: count, do not hide
    bad = ((raw - rraw).abs().max(dim=-1).values > 1e-4).sum().item()
    assert bad <= max(2, n_active // 500), f"{bad} of {n_active} active samples differ"
    assert O.psnr(out["rgb_map"].cpu(), ref["rgb_map"]) > 50.0


def test_error_paths():
    """Bad arguments come back as error codes / exceptions, never a crash: parents out of order, wrong sizes, missing
    workspace."""
    import ctypes as C
    from instant_nvr_b200 import cabi
    from instant_nvr_b200.optimizer import aux_handle
    from instant_nvr_b200.smpl_frame import SmplSubject, _pose_struct, prepare_frame
    from instant_nvr_b200.synthetic import make_subject
    sub = make_subject(3, n_verts=500)
    with pytest.raises(ValueError):
        SmplSubject(sub["joints"], sub["parents"], sub["weights"][:, :23], sub["tpose"])
    subject = SmplSubject(sub["joints"], sub["parents"], sub["weights"], sub["tpose"])
    with pytest.raises(ValueError):
        prepare_frame(subject, sub["wxyz"], sub["Rh"], sub["Th"], sub["poses"][:60])
    out = prepare_frame(subject, sub["wxyz"], sub["Rh"], sub["Th"], sub["poses"], volume=False)      # a 500-vertex body works too
    assert "pbw" not in out and torch.isfinite(out["A"]).all() and out["part_pts"].shape[2] == subject.maxlen
    lib, h = aux_handle(torch.device("cuda"))
    bad = sub["parents"].copy()
    bad[3] = 7                                                    # a child before its parent
    pose = _pose_struct(sub["Rh"], sub["Th"], sub["poses"], np.zeros(72), sub["joints"], bad)
    f = torch.empty(64, device="cuda")
    o = cabi.NvrSmplOut(f.data_ptr(), f.data_ptr(), None, None, f.data_ptr(), None, None, None)
    ws = torch.empty(4096, dtype=torch.uint8, device="cuda")
    rc = lib.nvr_smpl_pose_frame(h, C.byref(pose), f.data_ptr(), 1, None, 0, 0.05, C.byref(o), ws.data_ptr(), ws.numel(), None)
    assert rc != 0 and b"parents" in lib.nvr_last_error(h)
    pose = _pose_struct(sub["Rh"], sub["Th"], sub["poses"], np.zeros(72), sub["joints"], sub["parents"])
    rc = lib.nvr_smpl_pose_frame(h, C.byref(pose), f.data_ptr(), 1, None, 0, 0.05, C.byref(o), None, 0, None)
    assert rc != 0 and b"workspace" in lib.nvr_last_error(h)
    dims, origin = (C.c_int32 * 3)(), (C.c_double * 3)()
    assert lib.nvr_smpl_volume_dims(h, None, dims, origin, None) != 0
