"""Per-warp timeline of the blend forward (measurement build: tests/tools/build_ab_lib.sh WORKTREE fstats -DGSR_FWD_STATS [-D...]).
Which warps end the launch, when they started, how many splat steps they ran: is the launch bound by throughput or by
the serial chains of its heaviest quadrants?   usage: fwd_stats.py [workload] [lib tag]"""
import ctypes, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["GSR_BINDING"] = "ctypes"
from gs_localization_b200 import _lib as _l
tag = sys.argv[2] if len(sys.argv) > 2 else "fstats"
_l.LIB_PATH = os.path.join(ROOT, "ab", f"libgsr_b200_{tag}.so")
import torch
import bench
dev = torch.device("cuda:0")
arm = bench.Arm("ours", dev)
name = sys.argv[1] if len(sys.argv) > 1 else "headline"
cfg, gmap, m, cams = bench.build_workload(name, 0, dev)
bg = torch.zeros(3, device=dev)
lib = _l.load()
lib.gsr_debug_fwd_stats.argtypes = [ctypes.c_void_p]
for q in range(3):
    view, proj, _, campos = cams[q].matrices(dev)
    for rep in range(3):
        fwd = arm.c_forward(m, bg, view, proj, campos, cams[q])
    torch.cuda.synchronize()
    buf = np.zeros((16384, 4), np.uint64)
    assert lib.gsr_debug_fwd_stats(buf.ctypes.data) == 0
    T = ((cfg["W"] + 15) // 16) * ((cfg["H"] + 15) // 16)
    w = buf[:4 * T]
    w = w[w[:, 0] > 0]
    t0 = w[:, 0].min()
    start, end = (w[:, 0] - t0) / 1e3, (w[:, 1] - t0) / 1e3
    length, steps = (w[:, 3] & np.uint64(0xffffffff)).astype(np.int64), (w[:, 3] >> np.uint64(32)).astype(np.int64)
    sm = (w[:, 2] >> np.uint64(20)).astype(np.int64)
    dur = end - start
    order = np.argsort(-end)[:8]
    sm_end = np.array([end[sm == i].max() for i in np.unique(sm)])
    sm_steps = np.array([steps[sm == i].sum() for i in np.unique(sm)])
    print(json.dumps({"lib": tag, "workload": name, "pose": q, "warps": int(len(w)), "kernel_us": round(float(end.max()), 1),
                      "start_us_pct": [round(float(np.percentile(start, p)), 1) for p in (50, 75, 90, 100)],
                      "end_us_pct": [round(float(np.percentile(end, p)), 1) for p in (10, 50, 75, 90, 99, 100)],
                      "steps_total": int(steps.sum()), "steps_per_warp_pct": [int(np.percentile(steps, p)) for p in (10, 50, 90, 99, 100)],
                      "sm_end_us_pct": [round(float(np.percentile(sm_end, p)), 1) for p in (0, 10, 50, 90, 100)],
                      "sm_steps_pct": [int(np.percentile(sm_steps, p)) for p in (0, 10, 50, 90, 100)],
                      "last_warps": [{"start_us": round(float(start[i]), 1), "end_us": round(float(end[i]), 1), "list": int(length[i]), "steps": int(steps[i]),
                                      "ns_per_step": round(float(dur[i] * 1e3 / max(steps[i], 1)), 1)} for i in order],
                      "corr_dur_steps": round(float(np.corrcoef(dur, steps)[0, 1]), 3), "corr_dur_list": round(float(np.corrcoef(dur, length)[0, 1]), 3)}))
