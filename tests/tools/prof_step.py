"""Short driver for ncu: a few device-resident fwd+bwd steps of a bench workload (ours only).

    ncu --set full --clock-control none --import-source on -o gpurun_out/prof python tools/prof_step.py --steps 2
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gs_localization_b200 import synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="headline")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--impl", default="ours")
a = ap.parse_args()
dev = torch.device("cuda:0")
arm = bench.Arm(a.impl, dev)
cfg, gmap, m, cams = bench.build_workload(a.workload, 0, dev)
bg = torch.zeros(3, device=dev)
mats = [c.matrices(dev) for c in cams]
zD = torch.zeros(1, cfg["H"], cfg["W"], device=dev)
tgt = torch.rand(3, cfg["H"], cfg["W"], device=dev)
for i in range(a.warmup + a.steps):
    if i == a.warmup:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    q = i % len(cams)
    view, proj, _, campos = mats[q]
    fwd = arm.c_forward(m, bg, view, proj, campos, cams[q])
    gC = bench.l1_grad(fwd[1], tgt)
    arm.c_backward(m, bg, view, proj, campos, cams[q], fwd, gC, zD, zD)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", fwd[0])
