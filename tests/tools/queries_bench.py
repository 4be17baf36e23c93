"""queries/s of the graph-replayed refinement loop on several BASELINE configs (1 GPU)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ablib  # noqa: F401  (GSR_AB_LIB switch)
import torch
import bench
from gs_localization_b200 import synthetic as syn, localization as loc
dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["C1", "C2", "headline", "C3"]:
    cfg, gmap, m, cams = bench.build_workload(name, 0, dev)
    iters = cfg["iters"]
    arm = bench.Arm("ours", dev)
    bg = torch.zeros(3, device=dev)
    qs = []
    for q in range(7):
        gt = syn.make_camera(cfg, 20_000 + q)
        v, p_, _, c = gt.matrices(dev)
        target = arm.c_forward(m, bg, v, p_, c, gt)[1].clone()
        qs.append((gt, target))
    res = {"config": name, "iters": iters}
    for mode in (("graph",) if os.environ.get("QUICK") else ("graph", "fused_eager", "framework")):
        cams_q = [loc.PoseCamera(gt.perturbed(syn.initial_perturbation(q, 0.02, 1.0)), dev) for q, (gt, _) in enumerate(qs)]
        if mode == "graph":
            refiner = loc.GraphRefiner(m, cams_q[0])
            run = lambda cam, tgt: refiner.refine(cam, tgt, iters=iters)
        elif mode == "fused_eager":
            run = lambda cam, tgt: loc.refine_pose_fused(m, cam, tgt, iters=iters)
        else:
            run = lambda cam, tgt: loc.refine_pose(m, cam, tgt, iters=iters)
        run(cams_q[0], qs[0][1])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for cam, (gt, tgt) in zip(cams_q[1:], qs[1:]):
            run(cam, tgt)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        errs = [syn.pose_error(cam.w2c.cpu(), gt.w2c) for cam, (gt, _) in zip(cams_q[1:], qs[1:])]
        res[mode] = {"queries_per_s": round(6 / dt, 2), "ms_per_iter": round(dt / 6 / iters * 1e3, 4),
                     "median_err_m": round(sorted(e[0] for e in errs)[3], 5), "median_err_deg": round(sorted(e[1] for e in errs)[3], 4)}
    # two refiners on two streams: independent queries overlap on the GPU
    for nstreams in (2, 3):
        cams_q = [loc.PoseCamera(gt.perturbed(syn.initial_perturbation(q, 0.02, 1.0)), dev) for q, (gt, _) in enumerate(qs)]
        streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
        refs = []
        for s_ in streams:
            with torch.cuda.stream(s_):
                r_ = loc.GraphRefiner(m, cams_q[0])
                r_.refine(loc.PoseCamera(qs[0][0].perturbed(syn.initial_perturbation(0, 0.02, 1.0)), dev), qs[0][1], iters=iters)
                refs.append(r_)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        todo = list(zip(cams_q[1:], qs[1:]))
        while todo:
            batch, todo = todo[:nstreams], todo[nstreams:]
            for (cam, (gt, tgt)), r_, s_ in zip(batch, refs, streams):
                with torch.cuda.stream(s_):
                    r_.submit(cam, tgt, iters)
            for (cam, _), r_, s_ in zip(batch, refs, streams):
                with torch.cuda.stream(s_):
                    r_.collect()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        errs = [syn.pose_error(cam.w2c.cpu(), gt.w2c) for cam, (gt, _) in zip(cams_q[1:], qs[1:])]
        res[f"graph_x{nstreams}"] = {"queries_per_s": round(6 / dt, 2), "ms_per_iter": round(dt / 6 / iters * 1e3, 4),
                                     "median_err_m": round(sorted(e[0] for e in errs)[3], 5)}
        del refs
    # B queries per graph launch (parallel branches of one graph)
    for B in (4, 8):
        cams_q = [loc.PoseCamera(gt.perturbed(syn.initial_perturbation(q, 0.02, 1.0)), dev) for q, (gt, _) in enumerate(qs)]
        br = loc.BatchedGraphRefiner(m, cams_q[0], batch=B)
        warm = [loc.PoseCamera(qs[0][0].perturbed(syn.initial_perturbation(0, 0.02, 1.0)), dev) for _ in range(B)]
        br.refine_batch(warm, [qs[0][1]] * B, iters=iters)
        nq = 2 * B
        batch_cams = [loc.PoseCamera(qs[1 + i % 6][0].perturbed(syn.initial_perturbation(1 + i % 6, 0.02, 1.0)), dev) for i in range(nq)]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for b0 in range(0, nq, B):
            br.refine_batch(batch_cams[b0:b0 + B], [qs[1 + i % 6][1] for i in range(b0, b0 + B)], iters=iters)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        errs = [syn.pose_error(c.w2c.cpu(), qs[1 + i % 6][0].w2c) for i, c in enumerate(batch_cams)]
        res[f"batched_graph_B{B}"] = {"queries_per_s": round(nq / dt, 2), "ms_per_query_iter": round(dt / nq / iters * 1e3, 4),
                                      "median_err_m": round(sorted(e[0] for e in errs)[nq // 2], 5)}
        del br
        torch.cuda.empty_cache()
    print(json.dumps(res), flush=True)
    del m, gmap
    torch.cuda.empty_cache()
