"""Data-parallel map training across a densification step (run under torch.distributed.run with 2+ ranks; with
DP_SAME_GPU=1 all ranks share cuda:0 and the collectives go through gloo — the host logic and the exchange kernels
are the same as over NCCL).  For each mode (sparse, dense) it checks, every iteration:
  * the exchanged gradient equals the sum of the per-view gradients (recomputed on every rank from the same replica);
  * after the step, P and a checksum of every parameter / statistic are identical on all ranks — in particular right
    after densify_and_prune and reset_opacity, which only holds if the densification statistics of ALL views were
    exchanged too.
Prints one JSON line per mode on rank 0; exit status 0 = all checks passed."""
import json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gs_localization_b200 import gaussian_model as gm, io as gio, synthetic as syn, parallel

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
same_gpu = os.environ.get("DP_SAME_GPU") == "1"
dev = torch.device("cuda", 0 if same_gpu else int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("gloo" if same_gpu else "nccl")
cfg = dict(P=20_000, W=160, H=128, deg=2, f=120.0, box=1.0, sigma0=0.06)
cams = [syn.make_camera(cfg, i) for i in range(8)]
gen = torch.Generator().manual_seed(0)
gts = [torch.rand(3, cfg["H"], cfg["W"], generator=gen).to(dev) for _ in range(8)]
bg = torch.zeros(3, device=dev)
ok_all = True
events_by_mode = {}
for mode in ("sparse", "dense", "auto"):
    raw = gio.deactivate(syn.make_map(cfg["P"], cfg["deg"], cfg["sigma0"], cfg["box"], seed=0))
    model = gm.GaussianModel(cfg["deg"], device=dev)
    model.from_raw(raw)
    model.spatial_lr_scale = 1.0
    args = gm.default_training_args(densify_from_iter=2, densification_interval=3, densify_until_iter=14, opacity_reset_interval=8,
                                    densify_grad_threshold=2e-5)
    model.training_setup(args)
    trainer = parallel.DataParallelTrainer(model, args, mode=mode, extent=4.0)
    worst_sum, events, sizes = 0.0, [], []
    for it in range(1, 17):
        vid = parallel.shard_views(8, it, rank, world)
        torch.manual_seed(1234 + it)                   # the split samples must be the same draw on every rank
        # expected sum of the per-view gradients, from this rank's replica (replicas are identical if all is well)
        ref = None
        for r in range(world):
            v = parallel.shard_views(8, it, r, world)
            _, gr, _, _ = model.compute_gradients(cams[v], gts[v], bg, args, it)
            ref = [x.clone() for x in gr] if ref is None else [a + b for a, b in zip(ref, gr)]
        loss, g, g2d, out = model.compute_gradients(cams[vid], gts[vid], bg, args, it, after_forward=trainer.after_forward)
        trainer.reduce(g, g2d, out["radii"], it < args.densify_until_iter)
        errs = [float((a - b).norm() / (b.norm() + 1e-30)) for a, b in zip(g, ref)]
        worst_sum = max(worst_sum, max(errs))
        trainer.before_apply(it)
        info = model.apply_gradients(g, None, None, args, it, 4.0, stats_done=True)
        if info:
            events.append((it, info["cloned"], info["split"], info["pruned"], info["after"]))
        # replicas identical?
        P = int(model._xyz.shape[0])
        sizes.append(P)
        chk = [torch.tensor(float(P), dtype=torch.float64)]
        for t in list(model._params()) + [model.xyz_gradient_accum, model.denom, model.max_radii2D]:
            chk += [t.double().sum().cpu(), t.double().abs().sum().cpu()]
        chk = torch.stack(chk)
        if not same_gpu:
            chk = chk.to(dev)                           # NCCL reduces device tensors
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        # two ranks / dense all-reduce: every rank holds the same bits.  Visible rows at >= 3 ranks: tables are added in a
        # per-rank order, replicas differ in the last ulp between re-synchronisations — P must still be identical
        # (mode "auto" may have exchanged sparsely in an earlier step: the dense all-reduce of a later one does not undo that)
        tol = 1e-9 if (world <= 2 or trainer.mode == "dense") else 1e-5
        same = bool(hi[0] == lo[0]) and bool(((hi - lo).abs() <= tol * hi.abs().clamp_min(1.0)).all())
        if not same:
            ok_all = False
            if rank == 0:
                print(json.dumps({"mode": mode, "iteration": it, "error": "replicas differ", "lo": lo.tolist()[:3], "hi": hi.tolist()[:3]}))
            break
    ok = worst_sum < 1e-4 and len(events) >= 2 and len(set(sizes)) > 1
    events_by_mode[mode] = events
    ok_all &= ok
    if rank == 0:
        print(json.dumps({"mode": mode, "world": world, "grad_sum_rel_err": worst_sum, "densify_events": events, "final_P": sizes[-1], "ok": ok}))
# the exchange mode must not change what is densified: same statistics (sum over views of per-view norms) in every mode.
# The first event is decided by identical inputs; later ones may differ by a few rows (float addition order of the exchange).
same_first = len({tuple(ev[0]) for ev in events_by_mode.values()}) == 1
if rank == 0:
    print(json.dumps({"check": "first densification identical across exchange modes", "ok": same_first,
                      "first_events": {k: v[0] for k, v in events_by_mode.items()}}))
ok_all &= same_first
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
