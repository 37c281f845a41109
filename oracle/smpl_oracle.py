"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's per-frame SMPL preprocessing, the step that
produces the frame tensors the per-ray path consumes.  Only tests/ may import this file.

Pinned: tests/golden/smpl.npz holds the outputs of the reference's own functions (imported unmodified by
tests/golden/make_golden_smpl.py: if_nerf_data_utils.get_rigid_transformation / get_bounds, tpose_dataset.Dataset.load_smpl /
prepare_input / the use_knn block of __getitem__, tools/prepare_zjumocap.get_bweights, cv2.Rodrigues);
tests/test_smpl_oracle.py checks this restatement against them (bit for bit except where BLAS summation order enters).

Third-party arithmetic not under /root/reference:
  * cv2.Rodrigues (OpenCV; docs/install.md pins opencv-python via requirements.txt) -- `rodrigues_cv` restates the published
    algorithm (calib3d: theta = |r|; R = cos(theta) I + (1 - cos(theta)) r r^T + sin(theta) [r]_x, identity below DBL_EPSILON).
  * psbody.mesh.Mesh.closest_vertices(use_cgal=True) (MPI-IS/mesh, unpinned in the reference) -- nearest VERTEX and its
    Euclidean distance (CGALClosestPointTree.nearest); restated as an exact float64 nearest-neighbour search.
"""
import numpy as np

NUM_PARTS = 5
# lib/utils/blend_utils.py:9-17
PART_BW_MAP = {
    "body": [14, 13, 9, 6, 3, 0],
    "leg": [1, 2, 4, 5, 7, 8, 10, 11],
    "head": [12, 15],
    "larm": [16, 18, 20, 22],
    "rarm": [17, 19, 21, 23],
}
PARTNAMES = ["body", "leg", "head", "larm", "rarm"]


def batch_rodrigues(poses):
    """if_nerf_data_utils.py:523-542 (same dtype promotion as the reference: float32 poses give a float32 angle)."""
    batch_size = poses.shape[0]
    angle = np.linalg.norm(poses + 1e-8, axis=1, keepdims=True)                        # :527
    rot_dir = poses / angle                                                            # :528
    cos = np.cos(angle)[:, None]
    sin = np.sin(angle)[:, None]
    rx, ry, rz = np.split(rot_dir, 3, axis=1)
    zeros = np.zeros([batch_size, 1])
    K = np.concatenate([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], axis=1).reshape([batch_size, 3, 3])   # :535-537
    return np.eye(3)[None] + sin * K + (1 - cos) * np.matmul(K, K)                     # :539-540


def get_rigid_transformation(poses, joints, parents):
    """if_nerf_data_utils.py:545-577.  poses (24,3), joints (24,3), parents (24) -> (24,4,4) float32."""
    rot_mats = batch_rodrigues(poses)
    rel_joints = joints.copy()
    rel_joints[1:] -= joints[parents[1:]]                                              # :554-555 (in the dtype of `joints`)
    transforms_mat = np.concatenate([rot_mats, rel_joints[..., None]], axis=2)         # :558
    padding = np.zeros([24, 1, 4])
    padding[..., 3] = 1
    transforms_mat = np.concatenate([transforms_mat, padding], axis=1)                 # :561
    chain = [transforms_mat[0]]
    for i in range(1, parents.shape[0]):                                               # :564-567
        chain.append(np.dot(chain[parents[i]], transforms_mat[i]))
    transforms = np.stack(chain, axis=0)
    joints_homogen = np.concatenate([joints, np.zeros([24, 1])], axis=1)               # :571-572
    rel = np.sum(transforms * joints_homogen[:, None], axis=2)                         # :573
    transforms[..., 3] = transforms[..., 3] - rel                                      # :574
    return transforms.astype(np.float32)


def rodrigues_cv(rvec):
    """cv2.Rodrigues(rvec)[0] for a rotation VECTOR: computed in double whatever the input depth; the result has the
    input's depth (float32 in -> float32 out)."""
    r = np.asarray(rvec)
    out_dtype = np.float32 if r.dtype == np.float32 else np.float64
    x, y, z = (float(v) for v in r.reshape(3))
    theta = np.sqrt(x * x + y * y + z * z)
    if theta < np.finfo(np.float64).eps:
        return np.eye(3, dtype=out_dtype)
    c, s = np.cos(theta), np.sin(theta)
    c1 = 1.0 - c
    it = 1.0 / theta
    x, y, z = x * it, y * it, z * it
    rrt = np.array([[x * x, x * y, x * z], [x * y, y * y, y * z], [x * z, y * z, z * z]])
    r_x = np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])
    return (c * np.eye(3) + c1 * rrt + s * r_x).astype(out_dtype)


def big_poses_default(poses, tpose_geometry=True):
    """tpose_dataset.py:276-287: the canonical 'big pose' (cfg.tpose_geometry True in lib/config/config.py:237)."""
    big_poses = np.zeros_like(poses).ravel()
    if tpose_geometry:
        angle = 30
        big_poses[5] = np.deg2rad(angle)
        big_poses[8] = np.deg2rad(-angle)
    else:
        big_poses = big_poses.reshape(-1, 3)
        big_poses[1] = np.array([0, 0, 7. / 180. * np.pi])
        big_poses[2] = np.array([0, 0, -7. / 180. * np.pi])
        big_poses[16] = np.array([0, 0, -55. / 180. * np.pi])
        big_poses[17] = np.array([0, 0, 55. / 180. * np.pi])
    return big_poses.reshape(-1, 3)


def smpl_parts(weights):
    """tpose_dataset.py:96-110 (Dataset.load_smpl): part label of every vertex from its dominant joint."""
    parts = np.zeros((weights.shape[0],))
    weights_max = weights.argmax(axis=-1)
    for pid in range(NUM_PARTS):
        for bwid in PART_BW_MAP[PARTNAMES[pid]]:
            parts[weights_max == bwid] = pid
    return parts


def prepare_input(wxyz, Rh, Th, poses, joints, parents, big_poses=None):
    """tpose_dataset.py:247-293 without the file reads.  -> wxyz f32, pxyz f32, A, big_A f32 (24,4,4), R f32, Rh, Th f32."""
    wxyz = np.asarray(wxyz).astype(np.float32)                                         # :251
    Rh = np.asarray(Rh).astype(np.float32)                                             # :257
    Th = np.asarray(Th).astype(np.float32)                                             # :258
    R = rodrigues_cv(Rh).astype(np.float32)                                            # :259
    pxyz = np.dot(wxyz - Th, R).astype(np.float32)                                     # :269
    poses = np.asarray(poses).reshape(-1, 3)                                           # :272
    A = get_rigid_transformation(poses, joints, parents)                               # :275
    if big_poses is None:
        big_poses = big_poses_default(poses)
    big_A = get_rigid_transformation(big_poses, joints, parents)                       # :288-289
    return wxyz, pxyz, A, big_A, R, Rh, Th


def get_bounds(xyz, box_padding=0.05):
    """if_nerf_data_utils.py:689-696 (cfg.box_padding = 0.05, lib/config/config.py:87)."""
    min_xyz = np.min(xyz, axis=0)
    max_xyz = np.max(xyz, axis=0)
    min_xyz -= box_padding
    max_xyz += box_padding
    return np.stack([min_xyz, max_xyz], axis=0).astype(np.float32)


def part_tables(ppts, tpose, weights, parts, bbox_overlap=0.2):
    """tpose_dataset.py:570-600: per-part posed vertices / skinning rows (ragged, zero padded to the longest part) and
    the per-part big-pose bbox +- cfg.bbox_overlap (lib/config/config.py:27)."""
    N, D = weights.shape
    P = NUM_PARTS
    part_pts = np.zeros((P, N, 3), dtype=np.float32)
    part_pbw = np.zeros((P, N, D), dtype=np.float32)
    lengths2 = np.zeros(P, dtype=int)
    bounds = np.zeros((P, 2, 3), dtype=np.float32)
    for pid in range(P):
        part_flag = (parts == pid)
        lengths2[pid] = np.count_nonzero(part_flag)
        part_pts[pid, :lengths2[pid]] = ppts[part_flag]
        part_pbw[pid, :lengths2[pid]] = weights[part_flag]
        bounds[pid, 0] = tpose[part_flag].min(axis=0) - bbox_overlap
        bounds[pid, 1] = tpose[part_flag].max(axis=0) + bbox_overlap
    max_length = lengths2.max()
    return part_pts[:, :max_length, :], part_pbw[:, :max_length, :], lengths2, bounds


def get_grid_points(xyz, vsize=0.025):
    """tools/prepare_zjumocap.py:152-165: 2.5 cm voxel centres over bbox(xyz) +- 5 cm, 'ij' order."""
    min_xyz = np.min(xyz, axis=0)
    max_xyz = np.max(xyz, axis=0)
    min_xyz -= 0.05
    max_xyz += 0.05
    bounds = np.stack([min_xyz, max_xyz], axis=0)
    x = np.arange(bounds[0, 0], bounds[1, 0] + vsize, vsize)
    y = np.arange(bounds[0, 1], bounds[1, 1] + vsize, vsize)
    z = np.arange(bounds[0, 2], bounds[1, 2] + vsize, vsize)
    return np.stack(np.meshgrid(x, y, z, indexing="ij"), axis=-1)


def closest_vertices(verts, pts, brute=False, chunk=4096):
    """psbody Mesh.closest_vertices(pts, use_cgal=True): nearest vertex id and Euclidean distance, float64, exact
    (lowest index on exact ties).  Candidates come from an exact KD-tree query (scipy cKDTree, eps = 0; 3 candidates so
    that ties are seen) and are re-ranked with the explicit distance expression; brute=True scans every vertex."""
    verts = np.asarray(verts, dtype=np.float64)
    pts = np.asarray(pts, dtype=np.float64)
    ids = np.empty(len(pts), dtype=np.int64)
    dist = np.empty(len(pts), dtype=np.float64)
    if not brute:
        from scipy.spatial import cKDTree
        cand = np.sort(cKDTree(verts).query(pts, k=3)[1], axis=1)                      # ascending ids: argmin keeps the lowest
    for s in range(0, len(pts), chunk):
        v = verts[None] if brute else verts[cand[s:s + chunk]]
        d = pts[s:s + chunk, None, :] - v
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        i = d2.argmin(axis=1)
        ids[s:s + chunk] = i if brute else cand[s:s + chunk][np.arange(len(i)), i]
        dist[s:s + chunk] = np.sqrt(d2[np.arange(len(i)), i])
    return ids, dist


def get_bweights(vertices, Rh, Th, weights):
    """tools/prepare_zjumocap.py:474-508 (get_bweights) without the file reads and without its dead LBS lines (:495-498):
    -> (D,H,W,25) float32 = 24 skinning weights of the nearest posed vertex + the distance to it.
    `vertices` (V,3) world, `Rh` (3,), `Th` (1,3) as stored in the params file (the tool does not cast them)."""
    R = rodrigues_cv(np.asarray(Rh).reshape(3))                                        # :144 (get_transform_params)
    pxyz = np.dot(vertices - Th, R)                                                    # :485
    pts = get_grid_points(pxyz)
    sh = pts.shape
    vert_ids, norm = closest_vertices(pxyz, pts.reshape(-1, 3))                        # :493
    bweights = weights[vert_ids]
    bweights = np.concatenate((bweights, norm[:, None]), axis=1)                       # :505
    return bweights.reshape(*sh[:3], 25).astype(np.float32), vert_ids.reshape(sh[:3])
