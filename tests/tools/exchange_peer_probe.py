"""Visible-row gradient exchange on N GPUs, NCCL all_gather + add against the peer-memory pull (symmetric memory: the add
kernel reads the peers' tables over NVLink): same sums required, CUDA-event time of the whole exchange for both, at two
visible fractions of a C4-sized map.   torchrun --nproc-per-node N tests/tools/exchange_peer_probe.py"""
import json, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gs_localization_b200 import parallel
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
P, M = 3_000_000, 16
out = {"world": world}
for frac in (0.035, 0.095, 0.19):
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    radii = (torch.rand(P, generator=g, device=dev) < frac).to(torch.int32) * 5
    base = [torch.randn(P, w, device=dev, generator=g) * (radii > 0).float()[:, None] for w in (3, 3 * M, 1, 3, 4)]
    base[1] = base[1].view(P, M, 3)
    g2d = torch.randn(P, 3, device=dev, generator=g)
    res = {}
    sums = {}
    expect = [b.clone() for b in base]          # sum over ranks of the per-rank gradients, by NCCL
    for e_ in expect:
        dist.all_reduce(e_)
    for name, peer in (("all_gather", False), ("peer_pull", True)):
        ex = parallel.VisibleRowExchange(P, M, dev, peer_memory=peer)
        ms = []
        worst = 0.0
        for it in range(10):
            grads = [b * float(it + 1) for b in base]       # different values in the same exchange buffers every step
            stats = [torch.zeros(P, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev)]
            ex.begin(radii)
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            ex.exchange(grads, radii, g2d, stats)
            e1.record(); torch.cuda.synchronize()
            if it >= 3:
                ms.append(e0.elapsed_time(e1))
            worst = max(worst, max(float((a - e_ * float(it + 1)).abs().max() / (e_.abs().max() * (it + 1))) for a, e_ in zip(grads, expect)))
        t = torch.tensor([sorted(ms)[len(ms) // 2]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = {"exchange_ms": round(float(t[0]), 3), "used_peer_memory": bool(ex.used_peer_memory), "MB_per_table": round((ex.last_rows + 1) * ex.W * 4 / 1e6, 1),
                     "max_rel_err_vs_all_reduce": worst}
        sums[name] = [x.double().sum().item() for x in grads] + [s_.double().sum().item() for s_ in stats], grads
    same = all(torch.equal(a, b) for a, b in zip(sums["all_gather"][1], sums["peer_pull"][1]))
    res["identical_gradients"] = bool(same)
    # dense all-reduce of the same arena for comparison
    arena = torch.cat([b.reshape(-1) for b in base])
    for _ in range(3): dist.all_reduce(arena)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): dist.all_reduce(arena)
    e1.record(); torch.cuda.synchronize()
    res["dense_all_reduce_ms"] = round(e0.elapsed_time(e1) / 5, 3)
    out[f"visible_{frac}"] = res
    del base, arena
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
