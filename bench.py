#!/usr/bin/env python
"""bench.py — fwd+bwd rasterize iterations/s of the LoGS rasterizer hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload headline]

Metric (BASELINE.json): forward+backward rasterize iterations per second at 1M Gaussians,
640x480, SH degree 3 (the "headline" workload; other BASELINE configs are parity-test cases).
One *step* = one pose-refinement style iteration: rasterize forward -> L1 loss gradient ->
rasterize backward (all per-Gaussian gradients, like the reference computes them).

  value   whole-job iterations/s with the map, cameras and target images already resident in
          HBM, called through the `_C` C-ABI layer, timed with CUDA events on the launching stream.
  e2e     the same metric through the public API (GaussianRasterizer + autograd) with HOST inputs:
          every step copies that step's camera matrices and target image from pinned host memory
          and reads the loss back; wall clock, synchronised on both sides.
  N > 1   one process per GPU (torchrun), the map replicated, independent query poses per rank,
          no collective on the data path: weak scaling, value = sum over ranks.

`--impl reference` runs the UNMODIFIED reference CUDA rasterizer built by oracle/build_ref.sh
(oracle/_ref) through its own Python API on the same GPU, same workload.  If that build is
absent it times the CPU oracle port instead.  Both arms attach `cpu_baseline`: the CPU oracle
(oracle/gsr_oracle.cpp, a port of the same math) timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from gs_localization_b200 import synthetic as syn  # noqa: E402

N_POSES = 8  # distinct query cameras cycled through per rank


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline", choices=list(syn.CONFIGS))
    ap.add_argument("--queries", type=int, default=12, help="pose-refinement queries per rank for the queries/s figure (ours only)")
    ap.add_argument("--query-iters", type=int, default=50)
    ap.add_argument("--query-batch", type=int, default=4, help="queries per CUDA-graph launch (localization.BatchedGraphRefiner)")
    ap.add_argument("--c3-queries", type=int, default=512, help="queries of the localization_c3 block (BASELINE config 3), sharded over ranks")
    ap.add_argument("--no-scale-blocks", action="store_true", help="skip the localization_c3 and train_dp blocks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stage-split", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:   # nvidia-smi takes a moment to emit its first sample
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- workload
def build_workload(name, rank, device):
    cfg = syn.CONFIGS[name]
    gmap = syn.make_map(cfg["P"], cfg["deg"], cfg["sigma0"], cfg["box"], seed=0)
    # every rank works through the same N_POSES query poses against its own replica of the map: per-GPU work is
    # identical, which is what "weak scaling" means (`rank` only offsets the pose-refinement queries further down)
    cams = [syn.make_camera(cfg, q).perturbed(syn.initial_perturbation(q)) for q in range(N_POSES)]
    dmap = gmap.to(device) if device is not None else gmap
    return cfg, gmap, dmap, cams


def l1_grad(color, target):
    # d/dcolor of mean |color - target|
    return torch.sign(color - target) / color.numel()


class Arm:
    """One implementation (ours / reference) behind the same two call shapes."""

    def __init__(self, impl, device):
        self.impl = impl
        self.device = device
        if impl == "ours":
            import gs_localization_b200.diff_gaussian_rasterization as pkg
            from gs_localization_b200 import _lib
            _lib.load()
            self.lib = _lib
        else:
            sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
            import diff_gaussian_rasterization as pkg  # the unmodified reference package
            sys.path.pop(0)
            self.lib = None
        self.pkg = pkg
        self.C = pkg._C

    def c_forward(self, m, bg, view, proj, campos, cam):
        e = torch.Tensor([])
        return self.C.rasterize_gaussians(bg, m.means3D, e, m.opacities, m.scales, m.rotations, 1.0, e, view, proj,
                                          cam.tanfovx, cam.tanfovy, cam.H, cam.W, m.shs, m.sh_degree, campos, False, False)

    def c_backward(self, m, bg, view, proj, campos, cam, fwd, gC, gD, gA):
        R, color, depth, alpha, radii, geom, binning, img = fwd
        e = torch.Tensor([])
        return self.C.rasterize_gaussians_backward(bg, m.means3D, radii, e, m.scales, m.rotations, 1.0, e, view, proj,
                                                   cam.tanfovx, cam.tanfovy, gC, gD, gA, m.shs, m.sh_degree, campos, geom, R,
                                                   binning, img, alpha, False)


def run_gpu_arm(args, rank, world, local_rank):
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    arm = Arm(args.impl, device)
    cfg, gmap, m, cams = build_workload(args.workload, rank, device)
    H, W = cfg["H"], cfg["W"]
    bg = torch.zeros(3, device=device)
    mats = [c.matrices(device) for c in cams]
    zerosD = torch.zeros(1, H, W, device=device)

    # targets: render at the unperturbed (ground-truth) poses
    targets = []
    with torch.no_grad():
        for q in range(N_POSES):
            gt = syn.make_camera(cfg, q)
            v, p, _, c = gt.matrices(device)
            targets.append(arm.c_forward(m, bg, v, p, c, gt)[1].clone())

    stats = {}

    def step_device(i):
        q = i % N_POSES
        view, proj, _, campos = mats[q]
        fwd = arm.c_forward(m, bg, view, proj, campos, cams[q])
        gC = l1_grad(fwd[1], targets[q])
        arm.c_backward(m, bg, view, proj, campos, cams[q], fwd, gC, zerosD, zerosD)
        return fwd[0]

    # ---- value: device-resident, CUDA events
    for i in range(args.warmup):
        step_device(i)
    torch.cuda.synchronize()
    launches0 = arm.lib.launch_count() if arm.lib else 0
    if world > 1:
        torch.distributed.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    # one event per step boundary: total = first -> last (the reported mean), plus the per-step median
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    Rs = []
    for i in range(args.steps):
        Rs.append(step_device(i))
        evs[i + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    ms_total = evs[0].elapsed_time(evs[-1])
    per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
    stats["ms_per_step_median"] = per_step[len(per_step) // 2]
    stats["ms_per_step_p10_p90"] = [per_step[len(per_step) // 10], per_step[(len(per_step) * 9) // 10]]
    launches = (arm.lib.launch_count() - launches0) if arm.lib else 0

    # ---- e2e: public API, host inputs, wall clock
    settings_cls, raster_cls = arm.pkg.GaussianRasterizationSettings, arm.pkg.GaussianRasterizer
    host_mats = [[t.cpu().pin_memory() for t in (mt[0], mt[1], mt[3])] for mt in mats]
    host_targets = [t.cpu().pin_memory() for t in targets]
    means = m.means3D.clone().requires_grad_(True)
    shs = m.shs.clone().requires_grad_(True)
    opac = m.opacities.clone().requires_grad_(True)
    scales = m.scales.clone().requires_grad_(True)
    rots = m.rotations.clone().requires_grad_(True)
    params = [means, shs, opac, scales, rots]

    # inputs of step i+1 are uploaded on a copy stream while step i computes (every step still copies its own
    # inputs from pinned host memory inside the timed region; the copy just does not sit on the compute stream).
    # One packed pinned buffer per pose -> one copy per step; view/proj/campos/target are views of the device slot.
    copy_stream = torch.cuda.Stream(device)
    n_in = 16 + 16 + 4 + 3 * H * W
    host_packed = []
    for q in range(N_POSES):
        hp = torch.empty(n_in, dtype=torch.float32).pin_memory()
        hp[0:16] = host_mats[q][0].reshape(-1)
        hp[16:32] = host_mats[q][1].reshape(-1)
        hp[32:35] = host_mats[q][2]
        hp[36:] = host_targets[q].reshape(-1)
        host_packed.append(hp)
    dev_slots = [torch.empty(n_in, dtype=torch.float32, device=device) for _ in range(2)]
    dev_in = [(d[0:16].view(4, 4), d[16:32].view(4, 4), d[32:35], d[36:].view(3, H, W)) for d in dev_slots]
    copy_done = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    # the copy is issued with one cudaMemcpyAsync on the copy stream (entering a torch stream context every step costs
    # more host time than the call it wraps)
    import ctypes
    try:
        _rt = ctypes.CDLL("libcudart.so.12")
        _rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        _rt.cudaMemcpyAsync.restype = ctypes.c_int
    except OSError:
        _rt = None      # runtime library under another name: go through torch's copy_ on the copy stream

    def upload(i):
        q, slot = i % N_POSES, i % 2
        copy_stream.wait_event(consumed[slot])              # the step that last used this slot has finished with it
        if _rt is not None:
            rc = _rt.cudaMemcpyAsync(dev_slots[slot].data_ptr(), host_packed[q].data_ptr(), n_in * 4, 1, copy_stream.cuda_stream)
            if rc != 0:
                raise RuntimeError(f"cudaMemcpyAsync failed ({rc})")
        else:
            with torch.cuda.stream(copy_stream):
                dev_slots[slot].copy_(host_packed[q], non_blocking=True)
        copy_done[slot].record(copy_stream)

    for ev in consumed:
        ev.record(torch.cuda.current_stream(device))

    rasterizer = raster_cls(None)

    def step_e2e(i, first=False):
        q, slot = i % N_POSES, i % 2
        if first:
            upload(i)
        upload(i + 1)
        torch.cuda.current_stream(device).wait_event(copy_done[slot])
        view, proj, campos, target = dev_in[slot]
        cam = cams[q]
        rs = settings_cls(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
                          viewmatrix=view, projmatrix=proj, sh_degree=m.sh_degree, campos=campos, prefiltered=False, debug=False)
        means2D = torch.zeros_like(means, requires_grad=True)
        rasterizer.raster_settings = rs          # one module, new settings per view (the attribute is the reference's own)
        color, radii, depth, alpha = rasterizer(means3D=means, means2D=means2D, opacities=opac, shs=shs, scales=scales,
                                                rotations=rots)
        loss = (color - target).abs().mean()
        loss.backward()
        for p_ in params:
            p_.grad = None
        consumed[slot].record(torch.cuda.current_stream(device))
        # device -> host read of the step's result: asynchronous copy into pinned memory, consumed one step later
        # (the usual logging pattern of a training loop) so that the host can queue step i+1 behind step i
        loss_host[slot].copy_(loss.detach(), non_blocking=True)
        loss_ready[slot].record(torch.cuda.current_stream(device))
        got = None
        if not first:
            loss_ready[1 - slot].synchronize()
            got = float(loss_host[1 - slot])
        return got

    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    losses_read = 0
    nw = max(2 * N_POSES, args.warmup)   # every pose twice: allocator and speculative-capacity history in steady state
    for i in range(nw):
        step_e2e(i, first=(i == 0))
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for i in range(nw, nw + args.steps):
        losses_read += step_e2e(i) is not None
    loss_ready[(nw + args.steps - 1) % 2].synchronize()       # the last step's loss
    last_loss = float(loss_host[(nw + args.steps - 1) % 2])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert losses_read == args.steps and last_loss == last_loss
    clocks = sampler.stop()   # sampled across both timed regions (device-resident and end-to-end)
    if world > 1:
        torch.distributed.barrier()
    h2d = n_in * 4
    d2h = 4

    # ---- stage split + roofline of the dominant kernel (ours, rank 0 only; untimed extra passes)
    roofline, split, workload_stats = None, None, None
    if args.impl == "ours" and rank == 0 and not args.no_stage_split:
        arm.lib.stage_timing(True)
        for i in range(N_POSES * 2):
            step_device(i)
        torch.cuda.synchronize()
        split = arm.lib.stage_times()
        arm.lib.stage_timing(False)
        # measured workload statistics for the bytes model
        q = 0
        view, proj, _, campos = mats[q]
        fwd = arm.c_forward(m, bg, view, proj, campos, cams[q])
        Rq, radii = fwd[0], fwd[4]
        Pv = int((radii > 0).sum().item())
        st = arm.C.export_state(cfg["P"], Rq, W, H, fwd[5], fwd[6], fwd[7])
        n_contrib_sum = int(st["n_contrib"].to(torch.int64).sum().item())
        rng = st["ranges"].to(torch.int64)
        workload_stats = dict(P=cfg["P"], Pv=Pv, R=int(Rq), N=W * H, T=int(rng.shape[0]), n_contrib_sum=n_contrib_sum,
                              mean_n_contrib=n_contrib_sum / (W * H))
        roofline = make_roofline(split, workload_stats, cfg, ms_total / args.steps)
    # ---- localized queries/s (second half of BASELINE's metric): full pose refinements through the pose API,
    # queries sharded over ranks with no collective; fixed iteration count, convergence break disabled
    queries_s = None
    if args.impl == "ours" and args.queries > 0:
        from gs_localization_b200 import localization as loc
        qs = []
        for q in range(args.queries + 1):   # first one is a warm-up
            gq = q          # same queries on every rank (fixed per-GPU work); a deployment would hand each rank its own
            gt = syn.make_camera(cfg, 10_000 + gq)
            v, p_, _, c = gt.matrices(device)
            with torch.no_grad():
                target = arm.c_forward(m, bg, v, p_, c, gt)[1].clone()
            qs.append((loc.PoseCamera(gt.perturbed(syn.initial_perturbation(gq, trans_m=0.02, rot_deg=1.0)), device), target, gt))
        # B queries per CUDA-graph launch (parallel branches of one graph): the latency-bound binning kernels of one query
        # overlap the blend kernels of the others, one graph launch per iteration for the whole batch
        # ... and two such sets take turns on their own streams (PipelinedBatchRefiner): the host loads one batch while the
        # other's replays run
        B = args.query_batch
        refiner = loc.PipelinedBatchRefiner(m, qs[0][0], batch=B, depth=2, lr=1e-3)
        warm = [loc.PoseCamera(qs[0][2].perturbed(syn.initial_perturbation(0, trans_m=0.02, rot_deg=1.0)), device) for _ in range(2 * B)]
        refiner.refine_all(warm, [qs[0][1]] * (2 * B), iters=args.query_iters)      # warm-up: includes both sets' graph captures
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        res_ = refiner.refine_all([b_[0] for b_ in qs[1:]], [b_[1] for b_ in qs[1:]], iters=args.query_iters)
        errs = [r_[0] for r_ in res_]
        torch.cuda.synchronize()
        queries_s = time.perf_counter() - t0
        if world > 1:
            torch.distributed.barrier()
        stats["final_pose_err"] = [syn.pose_error(w.cpu(), q[2].w2c) for w, q in zip(errs, qs[1:])]
    stats.update(ms_total=ms_total, clocks=clocks, launches=launches, e2e_s=e2e_s, h2d=h2d, d2h=d2h, split=split,
                 roofline=roofline, workload_stats=workload_stats, mean_R=sum(Rs) / max(1, len(Rs)), queries_s=queries_s)
    return stats


# --------------------------------------------------------------------------- the two sharded paths of north_star
def run_localization_c3(args, rank, world, device):
    """BASELINE config 3: 2M-Gaussian map, 1024x576, `--c3-queries` DISTINCT queries split over the ranks with
    parallel.shard_queries (no collective on the data path, loop shape of pipelines/7scenes_localize_full_dslam.py:352-365),
    results gathered with parallel.gather_query_results.  Strong scaling: the query set is fixed, value = queries / max-rank time."""
    from gs_localization_b200 import localization as loc
    from gs_localization_b200 import parallel
    cfg = syn.CONFIGS["C3"]
    iters = cfg["iters"]
    m = syn.make_map(cfg["P"], cfg["deg"], cfg["sigma0"], cfg["box"], seed=0).to(device)
    arm_C = __import__("gs_localization_b200.diff_gaussian_rasterization", fromlist=["_C"])._C
    bg = torch.zeros(3, device=device)
    mine = parallel.shard_queries(args.c3_queries, rank, world)
    e = torch.Tensor([])

    def target_of(q):
        gt = syn.make_camera(cfg, 30_000 + q)
        v, p_, _, c = gt.matrices(device)
        with torch.no_grad():
            img = arm_C.rasterize_gaussians(bg, m.means3D, e, m.opacities, m.scales, m.rotations, 1.0, e, v, p_, gt.tanfovx, gt.tanfovy,
                                            gt.H, gt.W, m.shs, m.sh_degree, c, False, False)[1]
        return gt, img

    work = []
    for q in mine:
        gt, img = target_of(q)
        work.append((q, gt, loc.PoseCamera(gt.perturbed(syn.initial_perturbation(q, trans_m=0.05, rot_deg=1.0)), device), img))
    # warm-up batch (graph capture) on a query that is not part of the set
    B = args.query_batch
    gt_w, img_w = target_of(args.c3_queries + 7)
    warm = [loc.PoseCamera(gt_w.perturbed(syn.initial_perturbation(1, trans_m=0.05, rot_deg=1.0)), device) for _ in range(2 * B)]
    refiner = loc.PipelinedBatchRefiner(m, warm[0], batch=B, depth=2, lr=1e-3)
    refiner.refine_all(warm, [img_w] * (2 * B), iters=iters)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    local = {}
    res_ = refiner.refine_all([b_[2] for b_ in work], [b_[3] for b_ in work], iters=iters)
    poses_host = torch.stack([w2c for w2c, _ in res_]).cpu() if res_ else torch.zeros(0, 4, 4)    # the queries' results, on the host
    t_rank = time.perf_counter() - t0       # the error table below is bookkeeping of the benchmark
    for (q, gt, cam_q, img), w2c in zip(work, poses_host):
        et, er = syn.pose_error(w2c, gt.w2c)
        local[q] = torch.tensor([et, er], dtype=torch.float64, device=device)
    table = parallel.gather_query_results(local, args.c3_queries, 2)
    t = torch.tensor([t_rank, -t_rank], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    if rank != 0:
        return None
    tmax, tmin = float(t[0]), -float(t[1])
    errs = table.cpu()
    done = int((~torch.isnan(errs[:, 0])).sum())
    return {"workload": f"C3: {cfg['P']} Gaussians, {cfg['W']}x{cfg['H']}, SH degree {cfg['deg']}, {args.c3_queries} distinct queries, "
                        f"{iters} pose-refinement iterations each from a 5 cm / 1 deg initial error",
            "scaling": "strong", "queries": args.c3_queries, "queries_done": done, "n_gpus": world,
            "queries_per_s": round(args.c3_queries / tmax, 2), "rank_time_s_min_max": [round(tmin, 3), round(tmax, 3)],
            "tail_imbalance": round(tmax / max(tmin, 1e-9), 3),
            "median_final_err_m_deg": [round(float(errs[:, 0].nanmedian()), 5), round(float(errs[:, 1].nanmedian()), 4)],
            "refined_better_than_start": round(float((errs[:, 0] < 0.05).double().mean()), 3)}


def run_train_dp(args, rank, world, device):
    """BASELINE config 4: one data-parallel map-training step per rank-view (loop of gs/7scenes_gs_full_dslam.py:145-241) on the
    3M-Gaussian map at 1297x840: fused render + L1/SSIM loss + backward, gradient exchange over NCCL (dense: one all-reduce
    on the gradient arena; sparse: visible rows only, no host round trip), fused optimiser step.  Weak scaling over views."""
    from gs_localization_b200 import gaussian_model as gm
    from gs_localization_b200 import io as gio
    from gs_localization_b200 import parallel
    cfg = syn.CONFIGS["C4"]
    raw = gio.deactivate(syn.make_map(cfg["P"], cfg["deg"], cfg["sigma0"], cfg["box"], seed=0))
    gen = torch.Generator().manual_seed(0)
    cams = [syn.make_camera(cfg, i) for i in range(16)]
    gts = [torch.rand(3, cams[0].H, cams[0].W, generator=gen).to(device) for _ in range(16)]
    bg = torch.zeros(3, device=device)
    out = {"workload": f"C4: {cfg['P']} Gaussians, {cfg['W']}x{cfg['H']}, SH degree {cfg['deg']}, one view per rank and step, all parameter "
                       "groups trained, densification statistics exchanged", "scaling": "weak", "n_gpus": world}
    for mode in ("dense", "sparse", "auto"):
        model = gm.GaussianModel(cfg["deg"], device=device)
        model.from_raw(raw)
        model.spatial_lr_scale = 1.0
        opt = gm.default_training_args(densify_from_iter=10_000, densify_until_iter=15_000)   # statistics are exchanged, no densification inside the timed steps
        model.training_setup(opt)
        trainer = parallel.DataParallelTrainer(model, opt, mode=mode)
        acc = [[], []]

        def one(it, timed):
            vid = parallel.shard_views(16, it, rank, world)
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            t0 = time.perf_counter()
            loss, g, g2d, o = model.compute_gradients(cams[vid], gts[vid], bg, opt, it, after_forward=trainer.after_forward)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            trainer.reduce(g, g2d, o["radii"], True)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            model.apply_gradients(g, None, None, opt, it, 1.0, stats_done=True)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            if timed:
                acc[0].append(t3 - t0)
                acc[1].append(t2 - t1)

        # warm-up: every view once (allocator, speculative binning capacity and NCCL buffers in steady state), then the
        # median step over another pass through the views
        n_warm, n = 16, 16
        for it in range(1, 1 + n_warm):
            one(it, False)
        for it in range(1 + n_warm, 1 + n_warm + n):
            one(it, True)
        t = torch.tensor([sorted(a)[len(a) // 2] for a in acc], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        # gradient-sum parity: exchanged gradient of one more step against the per-view gradients recomputed here
        it = 1 + n_warm + n
        vid = parallel.shard_views(16, it, rank, world)
        ref = None
        for r in range(world):
            v = parallel.shard_views(16, it, r, world)
            _, gr, _, _ = model.compute_gradients(cams[v], gts[v], bg, opt, it)
            ref = [x.clone() for x in gr] if ref is None else [a + b for a, b in zip(ref, gr)]
        _, g, g2d, o = model.compute_gradients(cams[vid], gts[vid], bg, opt, it, after_forward=trainer.after_forward)
        trainer.reduce(g, g2d, o["radii"], False)
        err = max(float((a - b).norm() / (b.norm() + 1e-30)) for a, b in zip(g, ref))
        step_s, ex_s = float(t[0]), float(t[1])
        nbytes = trainer.exchange_bytes
        used = trainer.last_choice
        bus = (2.0 * (world - 1) / world * nbytes if used == "dense" else (world - 1) / world * nbytes) / max(ex_s, 1e-9) / 1e9 if world > 1 else None
        out[mode] = {"views_per_s": round(world / step_s, 1), "step_ms": round(step_s * 1e3, 3), "exchange_ms": round(ex_s * 1e3, 3),
                     "exchange_MB": round(nbytes / 1e6, 1), "bus_GBs": None if bus is None else round(bus, 1),
                     "grad_sum_rel_err": err, "exchange_used_last_step": used}
        del model, trainer
        torch.cuda.empty_cache()
    return out if rank == 0 else None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)", 1965.0


PROFILED_SOURCES = ("preprocess.cu", "binning.cu", "render.cu", "backward_pre.cu", "api.cu", "gsr_common.cuh", "gsr_kernels.cuh")


def kernel_source_sha256():
    """Hash of the sources of the kernels the ncu profile covers (the rasterizer's forward and backward; not the optimiser,
    loss, kNN or exchange kernels, which are not in it): a committed profile is only used if it was captured on exactly
    these sources."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(ROOT, "gs_localization_b200", "csrc", f) for f in PROFILED_SOURCES) + [os.path.join(ROOT, "include", "gsr_b200.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


PROFILE_JSON = os.path.join(ROOT, "profiles", "r2_roofline_profile.json")   # written by tests/tools/make_roofline_profile.py
STAGE_KERNELS = {   # stage of gsr_stage_timing -> kernels of the ncu capture that run inside it
    "preprocess": ["preprocess_cull_kernel", "preprocess_fwd_kernel", "color_fwd_kernel"],
    "tile_ranges": ["scan_tiles_kernel"],
    "duplicate_with_keys": ["scatter_kernel"],
    "radix_sort": ["tile_sort_kernel", "long_tile"],
    "render": ["render_fwd_kernel"],
    "render_backward": ["render_bwd_kernel"],
    "preprocess_backward": ["preprocess_bwd_kernel"],
}


def load_profile():
    """(per-stage {dram_bytes, warp_instructions, ncu_us} per launch, note).  None if absent or captured on other sources."""
    try:
        d = json.load(open(PROFILE_JSON))
    except Exception:
        return None, "no committed ncu profile (profiles/r2_roofline_profile.json)"
    if d.get("source_sha256") != kernel_source_sha256():
        return None, "committed ncu profile is STALE (captured on other kernel sources) - not used"
    out = {}
    for stage, names in STAGE_KERNELS.items():
        rows = [k for k in d["kernels"] if any(n in k["name"] for n in names)]
        if rows:
            out[stage] = {"dram_bytes": sum(k["dram_bytes_per_launch"] for k in rows),
                          "warp_instructions": sum(k["warp_instructions_per_launch"] for k in rows),
                          "ncu_us": sum(k["us_per_launch"] for k in rows)}
    return out, f"ncu --set full capture of these sources (git {d.get('git', '?')}, {d.get('when', '?')}): {os.path.relpath(PROFILE_JSON, ROOT)}"


def make_roofline(split, ws, cfg, ms_per_step):
    """Roofline record: the dominant kernel against the bound that actually limits it, every stage against HBM on its
    algorithmic bytes (SURVEY.md section 8d terms regrouped for the fused kernels; zero rows counted separately), and the
    whole iteration on SURVEY section 8d's B_fwd + B_bwd.  Everything needed to recompute it is in the record."""
    if not split:
        return None
    M = (cfg["deg"] + 1) ** 2
    P, Pv, R, N, T = ws["P"], ws["Pv"], ws["R"], ws["N"], ws["T"]
    passes = -(-(32 + max(1, (T - 1).bit_length())) // 8)
    fill = (92 + 12 * M) * P        # dense zero rows of the eight returned gradient tensors (B_fill without the internal conic)
    alg = {
        "preprocess": 44 * P + 12 * M * Pv + 4 * P + 97 * Pv,      # map read + SH rows of the visible + radii + per-visible records (89 B) + rect (8 B)
        "tile_ranges": 12 * T,
        "duplicate_with_keys": 12 * Pv + 8 * R + 48 * Pv,          # rect+depth per visible, one 8-byte composite per instance, accumulator rows zeroed
        "radix_sort": 8 * R + 4 * R,                               # tile-local sort: read the composites once, write the sorted slots
        "render": 52 * R + 24 * N + 48 * R + 20 * N,               # gather (4 + 48 B) per instance, images, + the record stream and the checkpoints it leaves for the backward (>= 1 per 64 entries blended)
        "render_backward": 48 * R + 28 * N + 36 * Pv,              # record stream back in, per-pixel reads, accumulator rows
        "preprocess_backward": 189 * Pv + (156 + 12 * M) * Pv,   # slot records in (no map rows are read), accumulator re-zero + gradient rows out
    }
    peak, how, sm_mhz = measured_peaks()
    prof, prof_note = load_profile()
    try:
        sms = torch.cuda.get_device_properties(0).multi_processor_count
    except Exception:
        sms = 148
    issue_peak = sms * 4 * sm_mhz * 1e6            # warp instructions per second: one per SM sub-partition and cycle
    stages = {}
    for k, ms in split.items():
        if ms <= 0:
            continue
        e = {"ms": round(ms, 4), "algorithmic_bytes": int(alg[k]), "hbm_frac": round(alg[k] / (ms * 1e-3) / 1e9 / peak, 4)}
        if k == "render_backward":
            e["algorithmic_bytes_with_zero_rows"] = int(alg[k] + fill)
            e["hbm_frac_with_zero_rows"] = round((alg[k] + fill) / (ms * 1e-3) / 1e9 / peak, 4)
        if prof and k in prof:
            e["dram_bytes_ncu"] = int(prof[k]["dram_bytes"])
            e["warp_instructions_ncu"] = int(prof[k]["warp_instructions"])
            e["issue_frac"] = round(prof[k]["warp_instructions"] / (ms * 1e-3) / issue_peak, 4)
        stages[k] = e
    dom = max(split, key=lambda k: split[k])
    d = stages[dom]
    blend = dom in ("render", "render_backward")
    rec = {"kernel": dom, "ms_per_launch": d["ms"], "peak_source": how, "profile": prof_note}
    if blend and "issue_frac" in d:
        # the blend kernels re-use every staged record for 64 pixels out of shared memory: instruction issue bounds them, not bytes
        rec.update(bound="issue", achieved=round(d["warp_instructions_ncu"] / (d["ms"] * 1e-3) / 1e9, 2), peak=round(issue_peak / 1e9, 2),
                   unit="Gwarp-inst/s", frac=d["issue_frac"],
                   how="warp instructions per launch (smsp__inst_executed.sum of the committed ncu capture) / live CUDA-event time / "
                       f"({sms} SMs x 4 schedulers x {sm_mhz:.0f} MHz)")
    else:
        rec.update(bound="hbm", achieved=round(d["algorithmic_bytes"] / (d["ms"] * 1e-3) / 1e9, 2), peak=peak, unit="GB/s", frac=d["hbm_frac"])
    rec["traffic"] = d.get("dram_bytes_ncu")
    rec["hbm"] = {"algorithmic_bytes_per_launch": d["algorithmic_bytes"], "frac": d["hbm_frac"],
                  "frac_with_zero_rows": d.get("hbm_frac_with_zero_rows"), "peak_GBs": peak}
    # whole iteration, SURVEY.md section 8d
    b_fwd = 64 * P + (83 + 12 * M) * Pv + (72 + 24 * passes) * R + 24 * N + 8 * T
    b_bwd = 44 * R + 28 * N + 4 * P + (207 + 24 * M) * Pv
    b_fill = (108 + 12 * M) * P
    it = {"B_fwd": int(b_fwd), "B_bwd": int(b_bwd), "B_fill": int(b_fill), "passes_reference_sort": passes, "ms_per_step": round(ms_per_step, 4),
          "hbm_frac": round((b_fwd + b_bwd) / (ms_per_step * 1e-3) / 1e9 / peak, 4),
          "hbm_frac_with_fill": round((b_fwd + b_bwd + b_fill) / (ms_per_step * 1e-3) / 1e9 / peak, 4),
          "formula": "SURVEY.md section 8d (the REFERENCE pipeline's bytes for this frame: what a byte-for-byte port would move)"}
    rec["stages"] = stages
    rec["iteration"] = it
    return rec


# --------------------------------------------------------------------------- CPU oracle timing
def cpu_oracle_timing(workload, budget_s=20.0):
    """The CPU oracle (a port of the reference math) on the host cores: bounded sample."""
    from oracle.oracle import Oracle, num_threads
    cfg, gmap, _, cams = build_workload(workload, 0, None)
    H, W = cfg["H"], cfg["W"]
    bg = torch.zeros(3)
    o = Oracle("f32")
    times = []
    t_start = time.perf_counter()
    q = 0
    while True:
        cam = cams[q % N_POSES]
        view, proj, _, campos = cam.matrices()
        t0 = time.perf_counter()
        o.forward(bg, gmap.means3D, None, gmap.opacities, gmap.scales, gmap.rotations, 1.0, None, view, proj, cam.tanfovx,
                  cam.tanfovy, H, W, gmap.shs, gmap.sh_degree, campos)
        color, _, _ = o.images()
        g = (torch.sign(torch.from_numpy(color)) / color.size).numpy()
        o.backward(g, None, None)
        times.append(time.perf_counter() - t0)
        q += 1
        if q >= 2 and (time.perf_counter() - t_start > budget_s or q >= 8):
            break
    times.sort()
    med = times[len(times) // 2]
    return {"value": round(1.0 / med, 4), "unit": "iterations/s", "cores": num_threads(), "kind": "port",
            "sample": f"{len(times)} full fwd+bwd iterations of the {workload} workload (median), OpenMP oracle/gsr_oracle.cpp",
            "pairs_last_pose": o.counters()}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = syn.CONFIGS[args.workload]
    config = {"workload": f"{args.workload}: {cfg['P']} Gaussians, {cfg['W']}x{cfg['H']}, SH degree {cfg['deg']}, "
                          f"{N_POSES} query poses/rank cycled, all per-Gaussian gradients",
              "l2_policy": "inputs larger than L2 (map = 236 B/Gaussian > 126 MB) and a different camera every step",
              "parallelism": f"queries sharded over {world} GPU(s), map replicated, no collective"}
    base = {"metric": "fwd+bwd rasterize iterations/s @1M Gaussians 640x480", "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config}

    ref_built = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "diff_gaussian_rasterization", "_C.so"))
    if args.impl == "reference" and not (ref_built and torch.cuda.is_available()):
        # CPU oracle port as the reference arm (rank 0 only)
        if rank != 0:
            return
        cb = cpu_oracle_timing(args.workload, budget_s=60.0)
        line = dict(base, impl="reference", value=cb["value"], ms_per_step=round(1e3 / cb["value"], 3), n_gpus=1,
                    cpu_baseline=cb, gpu_launches=0,
                    e2e={"value": cb["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gs_localization_b200 has no CPU fallback")
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    st = run_gpu_arm(args, rank, world, local_rank)

    scale_blocks = {}
    if args.impl == "ours" and not args.no_scale_blocks:
        dev_ = torch.device("cuda", local_rank)
        torch.cuda.empty_cache()
        scale_blocks["localization_c3"] = run_localization_c3(args, rank, world, dev_)
        torch.cuda.empty_cache()
        scale_blocks["train_dp"] = run_train_dp(args, rank, world, dev_)
    ms_total, e2e_s, queries_s = st["ms_total"], st["e2e_s"], st["queries_s"] or 0.0
    if world > 1:
        t = torch.tensor([ms_total, e2e_s, queries_s], device=f"cuda:{local_rank}", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total, e2e_s, queries_s = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        value = world * args.steps / (ms_total * 1e-3)
        line = dict(base, value=round(value, 3), ms_per_step=round(ms_total / args.steps, 4),
                    ms_per_step_median=round(st["ms_per_step_median"], 4), ms_per_step_p10_p90=[round(x, 4) for x in st["ms_per_step_p10_p90"]],
                    clocks=st["clocks"],
                    e2e={"value": round(world * args.steps / e2e_s, 3), "unit": "iterations/s",
                         "h2d_bytes_per_step": st["h2d"], "d2h_bytes_per_step": st["d2h"]},
                    gpu_launches=int(st["launches"]))
        if args.impl == "reference":
            line["impl"] = "reference"
            line["config"]["reference"] = "unmodified reference CUDA rasterizer built for sm_100a (oracle/_ref) on this GPU"
        if queries_s > 0:
            errs = st.get("final_pose_err") or []
            line["localization"] = {
                "queries_per_s": round(world * args.queries / queries_s, 3), "iters_per_query": args.query_iters,
                "queries": world * args.queries, "workload": f"{args.workload} map, pose-only refinement from a 2 cm / 1 deg initial error ({args.query_batch} queries per CUDA-graph launch, two such sets taking turns, one launch per iteration: "
                            "sync-free forward, L1 loss+grad kernel, pose-only backward, Adam+SE3 kernel per query)",
                "median_final_err_m_deg": [round(sorted(e[0] for e in errs)[len(errs) // 2], 5),
                                           round(sorted(e[1] for e in errs)[len(errs) // 2], 4)] if errs else None}
        for k, v in scale_blocks.items():
            if v:
                line[k] = v
        if st["roofline"]:
            line["roofline"] = st["roofline"]
        if st["split"]:
            line["stage_ms"] = {k: round(v, 4) for k, v in st["split"].items()}
        if st["workload_stats"]:
            line["workload_stats"] = st["workload_stats"]
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_oracle_timing(args.workload)
            pairs = line["cpu_baseline"].pop("pairs_last_pose", None)
            if pairs and st["split"]:
                # secondary roofline of the two blend kernels: contributing (pixel, splat) pairs against the
                # FP32 non-tensor peak, 148 SMs x 128 lanes x 2 FLOP x max clock; ~34 FLOP/pair forward, ~90 backward
                kc = pairs["pairs_contributing"]
                props = torch.cuda.get_device_properties(0)
                peak_tf = props.multi_processor_count * 128 * 2 * measured_peaks()[2] * 1e6 / 1e12
                fl = {"render": 34.0, "render_backward": 90.0}
                line["blend_compute"] = {k: {"pairs_per_s": round(kc / (st["split"][k] * 1e-3), 0),
                                             "tflops": round(kc * fl[k] / (st["split"][k] * 1e-3) / 1e12, 2),
                                             "frac_of_fp32_peak": round(kc * fl[k] / (st["split"][k] * 1e-3) / 1e12 / peak_tf, 4)}
                                         for k in fl}
                line["blend_compute"]["pairs_contributing"] = kc
                line["blend_compute"]["fp32_peak_tflops"] = round(peak_tf, 1)
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
