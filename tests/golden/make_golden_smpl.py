"""Golden fixture for the per-frame SMPL preprocessing: outputs of the REFERENCE'S OWN code
(zju3dv/instant-nvr @ a6f4d68), imported unmodified under package stubs and fed with a synthetic subject written to
temporary .npy / .pkl files in the layout the reference reads:

  * lib/utils/if_nerf/if_nerf_data_utils.py   get_rigid_transformation (:545-577), batch_rodrigues (:523-542),
                                               get_bounds (:689-696)
  * lib/datasets/h36m/tpose_dataset.py        Dataset.load_smpl (:96-110), Dataset.prepare_input (:247-293) called on a
                                               bare object carrying the attributes they read, and the `if cfg.use_knn:`
                                               block of __getitem__ (:561-600) -- that block sits inside a 300-line
                                               method that also decodes images, so its source lines are taken with
                                               inspect and executed as they are
  * tools/prepare_zjumocap.py                 get_grid_points (:152-165), get_bweights (:474-508)
  * cv2.Rodrigues                             (OpenCV, the library the reference calls)

Not installed here and stubbed: colored_traceback, termcolor, trimesh, imageio, plyfile, matplotlib, pytorch3d;
`psbody.mesh.Mesh` (third party, absent) is replaced by a stand-in whose `closest_vertices(pts, use_cgal=True)` returns the
exact nearest vertex and its Euclidean distance (scipy cKDTree in float64) -- the published semantics of psbody's
CGALClosestPointTree.nearest.

    python tests/golden/make_golden_smpl.py      ->  tests/golden/smpl.npz

The 25-channel volume is stored as (nearest vertex id, distance) per voxel + a SHA-256 of the full float32 array; the
test rebuilds the array from the ids and compares digests.
"""
import hashlib
import inspect
import os
import pickle
import sys
import tempfile
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
SEEDS = (3, 11)
N_VERTS = 6890                      # Dataset.load_smpl hard-codes np.zeros((6890,))


def install_stubs():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    stub("colored_traceback"); stub("colored_traceback.auto")
    stub("termcolor", colored=lambda s, *a, **k: s, cprint=lambda *a, **k: None)
    stub("trimesh"); stub("imageio"); stub("plyfile", PlyData=object)
    class _Cmap:                                  # img_utils.py:54-61 touches a colormap at import
        _lut = np.zeros((4, 4))

        def _init(self):
            pass
    stub("matplotlib", cm=None); stub("matplotlib.pyplot", get_cmap=lambda *a, **k: _Cmap()); stub("matplotlib.patches")
    if "PIL" not in sys.modules:
        try:
            import PIL  # noqa: F401
        except ImportError:
            stub("PIL", Image=object)
    import collections
    import torch
    _KNN = collections.namedtuple("KNN", "dists idx knn")
    stub("pytorch3d"); stub("pytorch3d.ops"); stub("pytorch3d.ops.knn", knn_points=lambda *a, **k: _KNN(None, None, None))
    if not torch.cuda.is_available():            # blend_utils / embedder build device='cuda' tensors at import
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        _t = torch.tensor
        torch.tensor = lambda *a, **k: _t(*a, **{kk: v for kk, v in k.items() if not (kk == "device" and v == "cuda")})
    from scipy.spatial import cKDTree

    class Mesh:                                   # psbody.mesh.Mesh stand-in: only what get_bweights touches
        def __init__(self, v=None, f=None):
            self.v, self.f = np.asarray(v, dtype=np.float64), f

        def closest_vertices(self, pts, use_cgal=False):
            d, i = cKDTree(self.v).query(np.asarray(pts, dtype=np.float64), k=1)
            return i, d
    stub("psbody"); stub("psbody.mesh", Mesh=Mesh)


def write_subject(root, sub, frame=0):
    """The files the reference reads, float dtypes as EasyMocap / the shipped tools write them."""
    lbs, meta = os.path.join(root, "lbs"), os.path.join(root, "smpl-meta")
    for d in (lbs, meta, os.path.join(root, "new_params"), os.path.join(root, "new_vertices"), os.path.join(lbs, "bweights")):
        os.makedirs(d, exist_ok=True)
    np.save(os.path.join(lbs, "joints.npy"), sub["joints"])
    np.save(os.path.join(lbs, "parents.npy"), sub["parents"])
    np.save(os.path.join(meta, "parents.npy"), sub["parents"])
    np.save(os.path.join(meta, "weights.npy"), sub["weights"])
    np.save(os.path.join(meta, "faces.npy"), np.zeros((1, 3), np.int64))
    # float64 files (EasyMocap writes python-float lists): the dataset casts what it wants in float32 itself, the
    # tool (get_bweights) then works in float64 throughout -- numpy-version-independent promotion
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    params = {"Rh": f64(sub["Rh"]), "Th": f64(sub["Th"]), "poses": f64(sub["poses"])[None], "shapes": np.zeros((1, 10))}
    np.save(os.path.join(root, "new_params", f"{frame}.npy"), params, allow_pickle=True)
    np.save(os.path.join(root, "new_vertices", f"{frame}.npy"), f64(sub["wxyz"]))
    # an SMPL pickle whose shaped template regresses exactly to `joints` is not needed: get_bweights only uses
    # A for a dead computation (can_pts, :495-498); weights / f / kintree_table / v_template are what matter
    V = sub["weights"].shape[0]
    smpl = {"v_template": np.zeros((V, 3)), "shapedirs": np.zeros((V, 3, 10)), "J_regressor": np.zeros((24, V)),
            "kintree_table": np.stack([sub["parents"], np.arange(24)]), "weights": sub["weights"].astype(np.float64),
            "f": np.zeros((1, 3), np.int64)}
    # joints must be distinct for nothing here; J_regressor = 0 gives zero joints, harmless (A unused downstream)
    with open(os.path.join(root, "smpl.pkl"), "wb") as f:
        pickle.dump(smpl, f)


def main():
    sys.path.insert(0, REPO)
    from instant_nvr_b200.synthetic import make_subject
    install_stubs()
    os.chdir(REF)
    sys.path.insert(0, REF)
    sys.argv = ["x", "--cfg_file", "configs/inb/inb_377.yaml"]
    import cv2
    from lib.config import cfg
    from lib.utils.if_nerf import if_nerf_data_utils as U
    from lib.datasets.h36m import tpose_dataset as TD
    sys.path.insert(0, os.path.join(REF, "tools"))
    import prepare_zjumocap as PZ

    # the `if cfg.use_knn:` block of Dataset.__getitem__, verbatim source lines
    src = inspect.getsource(TD.Dataset.__getitem__).split("\n")
    a = next(i for i, l in enumerate(src) if l.strip() == "if cfg.use_knn:")
    b = next(i for i, l in enumerate(src) if i > a and l.strip() == "return ret")
    knn_block = textwrap.dedent("\n".join(src[a:b]))

    st = {}
    for seed in SEEDS:
        sub = make_subject(seed, N_VERTS)
        with tempfile.TemporaryDirectory() as root:
            write_subject(root, sub)
            cfg.smpl_meta = os.path.join(root, "smpl-meta")
            cfg.vertices, cfg.params = "new_vertices", "new_params"
            self = types.SimpleNamespace(data_root=root, lbs_root=os.path.join(root, "lbs"))
            self.joints = np.load(os.path.join(self.lbs_root, "joints.npy")).astype(np.float32)     # :83-84
            self.parents = np.load(os.path.join(self.lbs_root, "parents.npy"))                        # :85
            faces, weights, joints, parents, parts = TD.Dataset.load_smpl(self)
            self.meta_smpl = {"faces": faces, "weights": weights, "joints": joints, "parents": parents, "parts": parts}
            # the pre-baked volume: tools/prepare_zjumocap.py prepare_blend_weights -> get_bweights
            pbw = PZ.get_bweights(os.path.join(root, "new_params", "0.npy"), os.path.join(root, "new_vertices", "0.npy"),
                                  os.path.join(root, "smpl.pkl"))
            np.save(os.path.join(self.lbs_root, "bweights", "0.npy"), pbw)
            wpts, ppts, A, big_A, pbw_loaded, Rh, Th = TD.Dataset.prepare_input(self, 0)
            tpose = sub["tpose"]
            ns = {"cfg": cfg, "np": np, "NUM_PARTS": TD.NUM_PARTS, "self": self, "ret": {}, "ppts": ppts, "wpts": wpts,
                  "tpose": tpose}
            exec(knn_block, ns)
            ret = ns["ret"]
            R = cv2.Rodrigues(Rh)[0].astype(np.float32)                                               # :489
        # nearest-vertex ids from the weight rows (rows of `weights` may repeat: keep the distance-consistent one)
        dist = pbw[..., 24]
        D, H, W = dist.shape
        import scipy.spatial
        pxyz64 = np.dot(sub["wxyz"].astype(np.float64) - sub["Th"].astype(np.float64), cv2.Rodrigues(sub["Rh"][0].astype(np.float64))[0])
        vid = scipy.spatial.cKDTree(pxyz64).query(PZ.get_grid_points(pxyz64).reshape(-1, 3), k=1)[1].reshape(D, H, W)
        assert np.array_equal(pbw[..., :24], sub["weights"][vid]) and pbw.dtype == np.float32
        k = f"s{seed}_"
        st.update({k + "parts": parts.astype(np.int8), k + "A": A, k + "big_A": big_A, k + "R": R, k + "ppts": ppts,
                   k + "pbounds": U.get_bounds(ppts), k + "wbounds": U.get_bounds(wpts), k + "tbounds": U.get_bounds(tpose),
                   k + "lengths2": ret["lengths2"].astype(np.int64), k + "bounds": ret["bounds"],
                   k + "part_pts_sha": np.frombuffer(hashlib.sha256(ret["part_pts"].tobytes()).digest(), np.uint8),
                   k + "part_pbw_sha": np.frombuffer(hashlib.sha256(ret["part_pbw"].tobytes()).digest(), np.uint8),
                   k + "part_shape": np.array(ret["part_pts"].shape), k + "pbw_shape": np.array(pbw.shape),
                   k + "pbw_sha": np.frombuffer(hashlib.sha256(np.ascontiguousarray(pbw).tobytes()).digest(), np.uint8)})
        if seed == SEEDS[0]:                                   # the other volumes are pinned by their digest alone
            st.update({k + "pbw_dist": dist, k + "pbw_vid": vid.astype(np.int16)})
        print(f"seed {seed}: A {A.dtype} big_A {big_A.dtype} ppts {ppts.dtype} pbw {pbw.shape} lengths2 {ret['lengths2']}, "
              f"part_pts {ret['part_pts'].shape} {ret['part_pts'].dtype}")
        # extra rotation vectors for the cv2.Rodrigues restatement (incl. the small-angle branch)
    rng = np.random.default_rng(0)
    rv = np.concatenate([rng.standard_normal((20, 3)), 1e-9 * rng.standard_normal((3, 3)), np.zeros((1, 3)),
                         np.array([[np.pi, 0, 0], [0, 3.0, 0.2]])]).astype(np.float32)
    st["rodrigues_in"] = rv
    st["rodrigues_out"] = np.stack([cv2.Rodrigues(r.reshape(1, 3))[0] for r in rv])
    st["rodrigues_out64"] = np.stack([cv2.Rodrigues(r.astype(np.float64).reshape(1, 3))[0] for r in rv])
    np.savez_compressed(os.path.join(HERE, "smpl.npz"), **st)
    print("wrote", os.path.join(HERE, "smpl.npz"), os.path.getsize(os.path.join(HERE, "smpl.npz")), "bytes")


if __name__ == "__main__":
    main()
