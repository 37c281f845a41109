"""The reference-facing surface of the drop-ins, checked against the reference's OWN objects inside the reference tree
(tests/dropin/check_dropin.py, run in a subprocess because importing the reference changes cwd / sys.argv and installs
package stubs).  Needs /root/reference: runs in the build container, skipped elsewhere (never part of ``-m gpu``)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/lib"), reason="the reference tree is not present")


@pytest.fixture(scope="module")
def report():
    p = subprocess.run([sys.executable, os.path.join(REPO, "tests", "dropin", "check_dropin.py"), "12"], capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def test_config_mapping_from_live_reference_cfg(report):
    assert report["config_diffs"] == [], report["config_diffs"]
    assert report["config"]["N_samples"] == 40 and report["config"]["smpl_thresh"] == 0.1     # CLI overrides are honoured
    assert report["config"]["use_pair_reg"] is True and report["config"]["use_reg_distortion"] is True   # inb_377.yaml


def test_no_arg_network_matches_reference_state_dict(report):
    assert report["keys_equal"] and report["shape_mismatch"] == []
    assert report["frozen_buffers"] > 0 and report["frozen_mismatch"] == []
    assert report["param_names_equal"] and report["requires_grad_equal"]


def test_optimizer_groups_match_reference(report):
    assert report["optimizer_is_fused"] and report["optimizer_len_equal"] and report["optimizer_groups"] == 67
    assert report["optimizer_mismatch"] == []


def test_renderer_module_resolves(report):
    assert report["renderer_has_render"]


def test_trainer_shim_wraps_reference_networkwrapper(report):
    """trainer_module: instant_nvr_b200.trainer -- the reference's own NetworkWrapper (all losses) with only the renderer
    swapped.  Importing the reference's trainer needs a few third-party packages; those absent here are stubbed."""
    t = report["trainer"]
    assert "unavailable" not in t, t
    assert t["is_reference_subclass"] and t["renderer_swapped"] and t["net_shared"]
