"""Golden fixture for the camera-ray step: outputs of the REFERENCE'S OWN
lib/utils/if_nerf/if_nerf_data_utils.get_rays_within_bounds (zju3dv/instant-nvr @ a6f4d68), imported unmodified
under package stubs (colored_traceback / termcolor / trimesh are not installed here).

    python tests/golden/make_golden_rays.py      ->  tests/golden/rays.npz

Cameras: a ZJU-MoCap-like pinhole (float64 K, R, T as the annotation files give them) looking at a 1 x 1.8 x 0.6 m
bbox, at 96x80 (the whole bbox in view, some rays miss) and a close-up at 64x64 (axis-parallel directions present).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def cameras():
    out = []
    rng = np.random.default_rng(5)
    for H, W, f, dist, yaw in ((80, 96, 90.0, 3.2, 0.4), (64, 64, 70.0, 2.5, -1.1), (48, 72, 60.0, 3.0, 0.0)):
        off = 0.0 if yaw == 0.0 else 0.5          # yaw 0: integer principal point -> exactly axis-parallel rays (:96-97 clamps)
        K = np.array([[f, 0.0, W / 2 - off], [0.0, f * 1.01, H / 2 - off], [0.0, 0.0, 1.0]])
        center = np.array([0.05, 0.1, 0.0])
        cam_pos = center + dist * np.array([np.sin(yaw), 0.15 if yaw != 0.0 else 0.0, -np.cos(yaw)])
        z = (center - cam_pos) / np.linalg.norm(center - cam_pos)      # look-at, image y pointing down
        x = np.cross(np.array([0.0, -1.0, 0.0]), z)
        x = x / np.linalg.norm(x)
        R = np.stack([x, np.cross(z, x), z])
        T = (-R @ cam_pos).reshape(3, 1)
        bounds = np.array([[-0.45, -0.8, -0.3], [0.55, 1.0, 0.3]], dtype=np.float32) + rng.normal(0, 0.01, (2, 3)).astype(np.float32)
        out.append((H, W, K, R, T, bounds))
    return out


def main():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    stub("colored_traceback")
    stub("colored_traceback.auto")
    stub("termcolor", colored=lambda s, *a, **k: s, cprint=lambda *a, **k: None)
    stub("trimesh")
    os.chdir(REF)
    sys.path.insert(0, REF)
    sys.argv = ["x", "--cfg_file", "configs/inb/inb_377.yaml"]
    from lib.utils.if_nerf import if_nerf_data_utils as U

    st = {}
    for n, (H, W, K, R, T, bounds) in enumerate(cameras()):
        ray_o, ray_d, near, far, mask = U.get_rays_within_bounds(H, W, K, R, T, bounds)
        st.update({f"c{n}_HW": np.array([H, W]), f"c{n}_K": K, f"c{n}_R": R, f"c{n}_T": T, f"c{n}_bounds": bounds,
                   f"c{n}_ray_o": ray_o, f"c{n}_ray_d": ray_d, f"c{n}_near": near, f"c{n}_far": far, f"c{n}_mask": mask})
        print(f"camera {n}: {H}x{W}, {mask.sum()} of {mask.size} rays hit the bbox; dtypes {ray_o.dtype} {near.dtype}")
    np.savez_compressed(os.path.join(HERE, "rays.npz"), **st)


if __name__ == "__main__":
    main()
