"""The SMPL frame-preprocessing restatement (oracle/smpl_oracle.py) against the reference's own outputs
(tests/golden/smpl.npz, made by tests/golden/make_golden_smpl.py), and the device arithmetic of csrc/nvr_smpl.cuh
(nvr_smpl_chain, nvr_rodrigues_cv, nvr_arange_len: `__host__ __device__`, compiled for the host by tests/host_emul)
against the oracle.  Runs without a GPU."""
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import pytest

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "oracle"))
import smpl_oracle as SO  # noqa: E402
from instant_nvr_b200.synthetic import make_subject  # noqa: E402

GOLD = np.load(os.path.join(REPO, "tests", "golden", "smpl.npz"))
SEEDS = (3, 11)


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def f64(a):
    return np.asarray(a, dtype=np.float64)


@pytest.fixture(scope="module", params=SEEDS)
def case(request):
    seed = request.param
    sub = make_subject(seed)
    g = {k[len(f"s{seed}_"):]: GOLD[k] for k in GOLD.files if k.startswith(f"s{seed}_")}
    return seed, sub, g


def test_rodrigues_restatement_matches_cv2():
    for r, ref32, ref64 in zip(GOLD["rodrigues_in"], GOLD["rodrigues_out"], GOLD["rodrigues_out64"]):
        got32 = SO.rodrigues_cv(r)
        assert got32.dtype == np.float32 and np.array_equal(got32, ref32.astype(np.float32))
        assert np.abs(SO.rodrigues_cv(r.astype(np.float64)) - ref64).max() <= 2.3e-16


def test_parts_and_transforms_bitwise(case):
    seed, sub, g = case
    parts = SO.smpl_parts(sub["weights"])
    assert np.array_equal(parts.astype(np.int8), g["parts"])
    # the files the golden run read hold float64 copies of these float32 values (make_golden_smpl.write_subject)
    wxyz, pxyz, A, big_A, R, Rh, Th = SO.prepare_input(f64(sub["wxyz"]), f64(sub["Rh"]), f64(sub["Th"]), f64(sub["poses"]),
                                                       sub["joints"], sub["parents"])
    for name, got in (("A", A), ("big_A", big_A), ("R", R), ("ppts", pxyz)):
        assert got.dtype == g[name].dtype and np.array_equal(got, g[name]), name
    assert np.array_equal(SO.get_bounds(pxyz), g["pbounds"]) and np.array_equal(SO.get_bounds(wxyz), g["wbounds"])
    assert np.array_equal(SO.get_bounds(sub["tpose"]), g["tbounds"])


def test_part_tables_bitwise(case):
    seed, sub, g = case
    part_pts, part_pbw, lengths2, bounds = SO.part_tables(g["ppts"], sub["tpose"], sub["weights"], g["parts"].astype(np.float64))
    assert np.array_equal(lengths2, g["lengths2"]) and np.array_equal(bounds, g["bounds"])
    assert tuple(part_pts.shape) == tuple(g["part_shape"])
    assert np.array_equal(sha(part_pts), g["part_pts_sha"]) and np.array_equal(sha(part_pbw), g["part_pbw_sha"])


def test_bweights_volume(case):
    """The 25-channel volume of tools/prepare_zjumocap.get_bweights: same dims, same nearest vertices, distances equal
    up to the last float32 bit of a float64 computation (BLAS dgemm vs numpy summation order in pxyz); with identical
    bits the whole array's digest matches the reference's."""
    seed, sub, g = case
    pbw, vid = SO.get_bweights(f64(sub["wxyz"]), f64(sub["Rh"])[0], f64(sub["Th"]), sub["weights"])
    assert tuple(pbw.shape) == tuple(g["pbw_shape"]) and pbw.dtype == np.float32
    if "pbw_dist" in g:
        assert np.array_equal(vid, g["pbw_vid"].astype(np.int64))
        ulp = np.abs(pbw[..., 24].view(np.int32) - g["pbw_dist"].view(np.int32))
        assert ulp.max() <= 1 and (ulp > 0).mean() < 1e-4
    if np.array_equal(sha(pbw), g["pbw_sha"]):
        return
    assert "pbw_dist" in g, "digest differs and this seed stores no per-voxel data to bound the difference"


# ---- device arithmetic compiled for the host ---------------------------------------------------------------------
@pytest.fixture(scope="module")
def emul():
    from test_host_emul import build_emul
    return build_emul()


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_device_chain_matches_oracle(emul, case):
    """nvr_smpl_chain (Rodrigues + kinematic chain in double, float32 out) against get_rigid_transformation on float64
    poses: identical up to one float32 ulp (np.dot vs fused / reordered double sums)."""
    seed, sub, g = case
    joints = np.ascontiguousarray(sub["joints"], dtype=np.float32)
    parents = np.ascontiguousarray(sub["parents"], dtype=np.int32)
    for poses, ref in ((f64(sub["poses"]), g["A"]), (SO.big_poses_default(f64(sub["poses"]).reshape(-1, 3)), g["big_A"])):
        poses = np.ascontiguousarray(poses.reshape(-1))
        out = np.zeros((24, 4, 4), np.float32)
        emul.emul_smpl_chain(ptr(poses), ptr(joints), ptr(parents), ptr(out))
        assert np.abs(out - ref).max() <= 1.2e-7 * max(1.0, np.abs(ref).max())
        assert (out != ref).mean() < 0.05
        assert np.array_equal(out[:, 3], ref[:, 3])


def test_device_rodrigues_matches_cv2(emul):
    for r, ref64 in zip(GOLD["rodrigues_in"], GOLD["rodrigues_out64"]):
        rv = np.ascontiguousarray(r, dtype=np.float64)
        out = np.zeros(9)
        emul.emul_rodrigues_cv(ptr(rv), ptr(out))
        assert np.abs(out.reshape(3, 3) - ref64).max() <= 2.3e-16


def test_device_grid_axis_matches_numpy_arange(emul):
    """np.arange(start, stop, step) in float64: the length is ceil((stop - start) / step) and the values are
    start + i * ((start + step) - start)  (numpy fills from the first two elements)."""
    rng = np.random.default_rng(0)
    emul.emul_arange_len.restype = C.c_int
    for _ in range(300):
        lo = rng.uniform(-2, 2)
        hi = lo + rng.uniform(0.05, 2.5)
        if rng.random() < 0.3:
            hi = lo + 0.025 * rng.integers(2, 90)            # extents that are (nearly) whole voxels: the ceil edge
        start, stop = lo - 0.05, (hi + 0.05) + 0.025
        ref = np.arange(start, stop, 0.025)
        n = emul.emul_arange_len(C.c_double(start), C.c_double(stop), C.c_double(0.025))
        assert n == len(ref)
        vals = np.zeros(n)
        emul.emul_arange_fill(C.c_double(start), C.c_double(0.025), C.c_int(n), ptr(vals))
        assert np.array_equal(vals, ref)


def test_host_side_subject_tables_bitwise(case):
    """SmplSubject (the product's per-subject host bookkeeping) against the reference's load_smpl / use_knn block."""
    from instant_nvr_b200.smpl_frame import SmplSubject, big_poses_default
    seed, sub, g = case
    s = SmplSubject(sub["joints"], sub["parents"], sub["weights"], sub["tpose"], device="cpu")
    assert np.array_equal(s.parts.astype(np.int8), g["parts"])
    assert np.array_equal(s.lengths2.numpy(), g["lengths2"]) and s.maxlen == int(g["part_shape"][1])
    assert np.array_equal(s.bounds.numpy(), g["bounds"]) and np.array_equal(s.tbounds.numpy(), g["tbounds"])
    assert np.array_equal(sha(s.part_pbw.numpy()), g["part_pbw_sha"])
    slot = s.vert_slot.numpy()
    part_pts = np.zeros((5, s.maxlen, 3), np.float32)
    part_pts.reshape(-1, 3)[slot] = g["ppts"]                      # what k_smpl_pose_verts does with the posed vertices
    assert np.array_equal(sha(part_pts), g["part_pts_sha"])
    assert np.array_equal(big_poses_default(), SO.big_poses_default(np.zeros((24, 3))))
