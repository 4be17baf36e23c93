"""One fwd+bwd per BASELINE config: stage split (ours) and total device time (ours vs reference)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ablib  # noqa: F401  (GSR_AB_LIB switch)
import torch
import bench
from gs_localization_b200 import synthetic as syn, _lib

names = sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5"]
dev = torch.device("cuda:0")
for name in names:
    cfg, gmap, m, cams = bench.build_workload(name, 0, dev)
    H, W = cfg["H"], cfg["W"]
    bg = torch.zeros(3, device=dev)
    zD = torch.zeros(1, H, W, device=dev)
    gC = torch.full((3, H, W), 1.0 / (3 * H * W), device=dev)
    res = {"config": name}
    for impl in (("ours",) if os.environ.get("ONLY_OURS") else ("ours", "reference")):
        try:
            arm = bench.Arm(impl, dev)
        except Exception as ex:
            res[impl] = f"unavailable: {ex}"
            continue
        mats = [c.matrices(dev) for c in cams]
        def step(i):
            q = i % len(cams)
            view, proj, _, campos = mats[q]
            fwd = arm.c_forward(m, bg, view, proj, campos, cams[q])
            arm.c_backward(m, bg, view, proj, campos, cams[q], fwd, gC, zD, zD)
            return fwd
        for i in range(2 * len(cams)): fwd = step(i)   # every pose twice: the speculative-capacity history is warm
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 2 * len(cams)
        e0.record()
        for i in range(n): fwd = step(i)
        e1.record(); torch.cuda.synchronize()
        res[impl] = {"ms_per_iter": round(e0.elapsed_time(e1) / n, 4), "R": int(fwd[0]), "Pv": int((fwd[4] > 0).sum())}
        if impl == "ours":
            _lib.stage_timing(True)
            for i in range(4): step(i)
            torch.cuda.synchronize()
            res["stage_ms"] = {k: round(v, 4) for k, v in _lib.stage_times().items()}
            _lib.stage_timing(False)
        del arm
    if isinstance(res.get("reference"), dict):
        res["speedup"] = round(res["reference"]["ms_per_iter"] / res["ours"]["ms_per_iter"], 2)
    print(json.dumps(res), flush=True)
    del m, gmap
    torch.cuda.empty_cache()
