"""Per-warp timeline of the blend backward (measurement build: tests/tools/build_ab_lib.sh WORKTREE stats -DGSR_BWD_STATS)."""
import ctypes, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["GSR_BINDING"] = "ctypes"
from gs_localization_b200 import _lib as _l
_l.LIB_PATH = os.path.join(ROOT, "ab", "libgsr_b200_stats.so")
import torch
import bench
dev = torch.device("cuda:0")
arm = bench.Arm("ours", dev)
name = sys.argv[1] if len(sys.argv) > 1 else "headline"
cfg, gmap, m, cams = bench.build_workload(name, 0, dev)
bg = torch.zeros(3, device=dev)
zD = torch.zeros(1, cfg["H"], cfg["W"], device=dev)
tgt = torch.rand(3, cfg["H"], cfg["W"], device=dev)
lib = _l.load()
for q in range(3):
    view, proj, _, campos = cams[q].matrices(dev)
    for rep in range(3):
        fwd = arm.c_forward(m, bg, view, proj, campos, cams[q])
        arm.c_backward(m, bg, view, proj, campos, cams[q], fwd, bench.l1_grad(fwd[1], tgt), zD, zD)
    torch.cuda.synchronize()
    buf = np.zeros((8192, 4), np.uint64)
    lib.gsr_debug_bwd_stats.argtypes = [ctypes.c_void_p]
    assert lib.gsr_debug_bwd_stats(buf.ctypes.data) == 0
    w = buf[buf[:, 0] > 0]
    t0 = w[:, 0].min()
    start, end = (w[:, 0] - t0) / 1e3, (w[:, 1] - t0) / 1e3
    steps, units = w[:, 3].astype(np.int64), w[:, 2].astype(np.int64)
    print(json.dumps({"pose": q, "workers": int(len(w)), "kernel_us": float(end.max()), "start_us_max": float(start.max()),
                      "end_us_pct": [float(np.percentile(end, p)) for p in (10, 50, 75, 90, 99, 100)],
                      "units_total": int(units.sum()), "steps_total": int(steps.sum()),
                      "steps_per_worker_pct": [int(np.percentile(steps, p)) for p in (10, 50, 90, 100)],
                      "units_per_worker_pct": [int(np.percentile(units, p)) for p in (10, 50, 90, 100)],
                      "us_per_step_median_worker": float(np.median((end - start) / np.maximum(steps, 1)))}))
