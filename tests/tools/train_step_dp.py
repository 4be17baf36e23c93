"""Data-parallel map-training step on N GPUs (BASELINE config C4): each rank renders a different view of a
replicated map, runs the full backward, and the per-Gaussian gradients are summed with ONE NCCL all-reduce of a
flat fp32 bucket (gs_localization_b200.parallel.GradientBucket).  Prints per-rank-max step time, the all-reduce
time and bus bandwidth, and checks on rank 0 that the reduced gradient equals the sum of the per-view gradients
computed locally.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_dp.py [C4]
"""
import json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gs_localization_b200 import synthetic as syn, parallel
from gs_localization_b200.diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = syn.CONFIGS[name]
gmap = syn.make_map(cfg["P"], cfg["deg"], cfg["sigma0"], cfg["box"], seed=0).to(dev)
# the reference's parameter split (scene/gaussian_model.py:108-111): xyz, f_dc, f_rest, opacity, scaling, rotation
xyz = gmap.means3D.clone().requires_grad_(True)
f_dc = gmap.shs[:, :1].clone().requires_grad_(True)
f_rest = gmap.shs[:, 1:].clone().requires_grad_(True)
opac, scal, rot = (t.clone().requires_grad_(True) for t in (gmap.opacities, gmap.scales, gmap.rotations))
params = [xyz, f_dc, f_rest, opac, scal, rot]
bucket = parallel.GradientBucket(params)
bg = torch.zeros(3, device=dev)

def view_step(view_id):
    cam = syn.make_camera(cfg, view_id)
    v, p, _, c = cam.matrices(dev)
    rs = GaussianRasterizationSettings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, bg, 1.0, v, p, gmap.sh_degree, c, False, False)
    shs = torch.cat([f_dc, f_rest], 1)
    m2 = torch.zeros_like(xyz, requires_grad=True)
    color, radii, depth, alpha = GaussianRasterizer(rs)(means3D=xyz, means2D=m2, opacities=opac, shs=shs, scales=scal, rotations=rot)
    loss = color.mean() + 0.01 * depth.mean()
    for q in params:
        q.grad = None
    loss.backward()

times, ar_times = [], []
for step in range(6):
    vid = parallel.shard_views(1000, step, rank, world)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    view_step(vid)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    bucket.allreduce()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if step >= 2:
        times.append(t2 - t0); ar_times.append(t2 - t1)
t = torch.tensor([sum(times) / len(times), sum(ar_times) / len(ar_times)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# parity: reduced gradient == sum over the world's views, recomputed on rank 0
ok = None
last_step = 5
reduced = [q.grad.clone() for q in params]
if rank == 0:
    ref = [torch.zeros_like(q) for q in params]
    for r in range(world):
        view_step(parallel.shard_views(1000, last_step, r, world))
        for a, q in zip(ref, params):
            a += q.grad
    errs = [float((a - b).norm() / (a.norm() + 1e-30)) for a, b in zip(ref, reduced)]
    ok = max(errs) < 1e-3
    nbytes = bucket.nbytes()
    busbw = 2 * (world - 1) / world * nbytes / float(t[1]) / 1e9 if world > 1 else None
    print(json.dumps({"config": name, "n_gpus": world, "step_ms": round(float(t[0]) * 1e3, 3), "allreduce_ms": round(float(t[1]) * 1e3, 3),
                      "bucket_MB": round(nbytes / 1e6, 1), "allreduce_busbw_GBs": None if busbw is None else round(busbw, 1),
                      "views_per_s": round(world / float(t[0]), 2), "grad_sum_rel_err": max(errs), "parity_ok": ok}))
if world > 1:
    dist.destroy_process_group()
