"""bench.py's localized-queries block, batch by batch, with each refiner's counters (R, overflow, longest list), capacity
and sort path: shows re-captures, eager fall-backs and long-list switches inside the timed region."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ablib  # noqa: F401
import torch
import bench
from gs_localization_b200 import synthetic as syn, localization as loc
dev = torch.device("cuda:0")
cfg, gmap, m, cams = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "headline", 0, dev)
arm = bench.Arm("ours", dev)
bg = torch.zeros(3, device=dev)
B, iters, nq = 4, 50, 12
qs = []
for q in range(nq + 1):
    gt = syn.make_camera(cfg, 10_000 + q)
    v, p_, _, c = gt.matrices(dev)
    with torch.no_grad():
        target = arm.c_forward(m, bg, v, p_, c, gt)[1].clone()
    qs.append((loc.PoseCamera(gt.perturbed(syn.initial_perturbation(q, trans_m=0.02, rot_deg=1.0)), dev), target, gt))
refiner = loc.BatchedGraphRefiner(m, qs[0][0], batch=B, lr=1e-3)
warm = [loc.PoseCamera(qs[0][2].perturbed(syn.initial_perturbation(0, trans_m=0.02, rot_deg=1.0)), dev) for _ in range(B)]
refiner.refine_batch(warm, [qs[0][1]] * B, iters=iters)
torch.cuda.synchronize()
todo = list(qs[1:])
while todo:
    batch, todo = todo[:B], todo[B:]
    caps0 = [(r.capacity, r.global_sort) for r in refiner.refiners]
    t0 = time.perf_counter()
    refiner.refine_batch([b_[0] for b_ in batch], [b_[1] for b_ in batch], iters=iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"batch_ms": round(dt * 1e3, 2), "ms_per_query_iter": round(dt / len(batch) / iters * 1e3, 4),
                      "caps_before": caps0, "caps_after": [(r.capacity, r.global_sort) for r in refiner.refiners],
                      "counters": [r._counters() for r in refiner.refiners]}), flush=True)
