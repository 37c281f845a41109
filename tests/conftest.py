import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_meta():
    meta = {}
    with open(os.path.join(GOLDEN, "META.txt")) as f:
        for line in f:
            k, v = line.strip().split(": ", 1)
            meta[k] = v
    return meta


@pytest.fixture(scope="session")
def golden_setup():
    """Inputs of the golden run, regenerated from the recorded seeds: (cfg, frame, rays, state_dict@gain200)."""
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays

    meta = golden_meta()
    seed = int(meta["seed"])
    cfg = PathConfig.inb_377(N_samples=int(meta["n_samples"]), log2_T_cap=int(meta["log2_T_cap"]))
    cfg = cfg.with_(smpl_thresh=float(meta["smpl_thresh"]))
    frame = make_frame(seed=seed)
    rays = make_rays(frame, int(meta["img"]), int(meta["img"]))
    net = Network(cfg, device="cpu")
    sd = net.state_dict()
    fill_weights(sd, seed=seed, table_gain=200.0, bounds=frame["bounds"][0])
    batch = dict(frame)
    batch.update(rays)
    return dict(cfg=cfg, frame=frame, rays=rays, batch=batch, sd=sd, net=net, seed=seed)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))
