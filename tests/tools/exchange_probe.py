"""Pieces of the visible-row gradient exchange at C4 scale on N GPUs: pack kernel, NCCL all_gather, add kernel — each timed
with CUDA events, back to back and with ~4 ms of unrelated GPU work in between (as inside a training step)."""
import ctypes as C, json, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gs_localization_b200 import _lib
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
P, M = 3_000_000, 16
g = torch.Generator(device=dev).manual_seed(rank)
radii = (torch.rand(P, generator=g, device=dev) < 0.035).to(torch.int32) * 5
grads = [torch.randn(P, w, device=dev) for w in (3, 3 * M, 1, 3, 4)]
g2d = torch.randn(P, 3, device=dev)
stats = [torch.zeros(P, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev)]
W = 1 + 11 + 3 * M + 3
cap = 110_592
table = torch.empty((cap + 1) * W, device=dev)
gathered = torch.empty(world * (cap + 1) * W, device=dev)
count = torch.zeros(1, dtype=torch.int32, device=dev)
tab = (C.c_void_p * 5)(*[t.data_ptr() for t in grads])
stream = torch.cuda.current_stream(dev).cuda_stream
filler_a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
def filler():
    for _ in range(3): torch.matmul(filler_a, filler_a)
res = {}
for gap in (False, True):
    acc = [0.0, 0.0, 0.0]
    for it in range(12):
        if gap: filler()
        ev = [torch.cuda.Event(True) for _ in range(4)]
        ev[0].record()
        lib.gsr_pack_visible_rows(radii.data_ptr(), P, M, tab, g2d.data_ptr(), table.data_ptr(), cap, count.data_ptr(), stream)
        ev[1].record()
        dist.all_gather_into_tensor(gathered, table)
        ev[2].record()
        for r in range(world):
            part = gathered[r * (cap + 1) * W:(r + 1) * (cap + 1) * W]
            lib.gsr_add_counted_rows(part.data_ptr(), cap, M, P, tab, int(r != rank), stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), stream)
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 4:
            for k in range(3): acc[k] += ev[k].elapsed_time(ev[k + 1]) / 8
    res["with_gap" if gap else "back_to_back"] = {"pack_ms": round(acc[0], 3), "all_gather_ms": round(acc[1], 3), "add_ms": round(acc[2], 3)}
if rank == 0:
    print(json.dumps({"world": world, "rows": int(count[0]), "table_MB": round(table.numel() * 4 / 1e6, 1), **res}))
dist.destroy_process_group()
