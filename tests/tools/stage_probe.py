"""Stage split of the headline workload for A/B switches set through the environment (prints one line)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda:0")
arm = bench.Arm("ours", dev)
cfg, gmap, m, cams = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "headline", 0, dev)
bg = torch.zeros(3, device=dev)
mats = [c.matrices(dev) for c in cams]
zD = torch.zeros(1, cfg["H"], cfg["W"], device=dev)
tgt = torch.rand(3, cfg["H"], cfg["W"], device=dev)
def step(i):
    q = i % len(cams)
    view, proj, _, campos = mats[q]
    fwd = arm.c_forward(m, bg, view, proj, campos, cams[q])
    arm.c_backward(m, bg, view, proj, campos, cams[q], fwd, bench.l1_grad(fwd[1], tgt), zD, zD)
for i in range(16): step(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
for i in range(200): step(i)
b.record(); torch.cuda.synchronize()
total = a.elapsed_time(b) / 200
arm.lib.stage_timing(True)
for i in range(32): step(i)
torch.cuda.synchronize()
split = arm.lib.stage_times()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("GSR_")}, "ms_per_step": round(total, 4),
                  "stage_ms": {k: round(v, 4) for k, v in split.items()}}))
