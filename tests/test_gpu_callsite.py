"""Executes INTEGRATION.md §1's claim: with gs_localization_b200/dropin first on sys.path the reference's own import
lines and keyword calls (gaussian_renderer/__init__.py:36-95, pipelines/tools/__init__.py:58-141,
scene/gaussian_model.py:20) run on this library and agree with the unmodified reference build."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_call_sites_through_dropin():
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "callsite_replay.py")], cwd="/", env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["ok"]
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "diff_gaussian_rasterization", "_C.so")):
        assert rep["reference_build"] and rep["viewspace_grad_rel"] <= 1e-3
