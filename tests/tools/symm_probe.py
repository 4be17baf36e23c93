import os, time, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = 16 * 1024 * 1024   # 64 MB of floats
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    t.fill_(rank + 1.0)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    acc = torch.zeros(n, device=dev)
    for _ in range(3): acc.add_(peer)
    torch.cuda.synchronize(); hdl.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): acc.add_(peer)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(rank, "symm ok: peer value", float(peer[0]), "p2p read+add 64MB:", round(ms, 3), "ms ->", round(n * 4 / ms / 1e6, 1), "GB/s")
except Exception as ex:
    print(rank, "symm failed:", repr(ex)[:300])
# all_gather timing for comparison
src = torch.ones(n, device=dev); out = torch.empty(world * n, device=dev)
for _ in range(3): dist.all_gather_into_tensor(out, src)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10): dist.all_gather_into_tensor(out, src)
e1.record(); torch.cuda.synchronize()
print(rank, "all_gather 64MB/rank:", round(e0.elapsed_time(e1) / 10, 3), "ms")
e0.record()
for _ in range(10): dist.all_reduce(out)
e1.record(); torch.cuda.synchronize()
print(rank, "all_reduce %dMB:" % (world * 64), round(e0.elapsed_time(e1) / 10, 3), "ms")
dist.destroy_process_group()
