"""The ctypes binding of `_C` (gsr_alloc_fn callbacks created from Python, raw pointers through
ctypes) is the documented alternative to the compiled glue; the parity suite normally runs on the compiled one.  Here
the reference-build comparisons and the float64 oracle gradients are re-run in a fresh interpreter with
GSR_BINDING=ctypes, and the process is checked to have really used that binding."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parity_suite_under_ctypes_binding():
    env = dict(os.environ, GSR_BINDING="ctypes")
    probe = subprocess.run([sys.executable, "-c",
                            "import gs_localization_b200.diff_gaussian_rasterization._C as C; print(C._B is None)"],
                           cwd=ROOT, env=env, capture_output=True, text=True)
    assert probe.returncode == 0 and probe.stdout.strip() == "True", probe.stdout + probe.stderr
    sel = ("test_forward_vs_reference_bit_exact or test_backward_vs_reference or test_backward_vs_oracle_f64 or "
           "test_empty_and_all_culled or test_speculative_launch_overflow_and_reuse or test_mark_visible or "
           "test_unused_outputs_need_no_zero_gradients or test_precomputed_colors_and_cov")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q", "-m", "gpu",
                        "-k", sel, "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]
