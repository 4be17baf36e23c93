"""Host timeline of bench.py's end-to-end step (public API, deferred loss read): where the host spends a step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from gs_localization_b200 import synthetic as syn

dev = torch.device("cuda:0")
arm = bench.Arm(sys.argv[1] if len(sys.argv) > 1 else "ours", dev)
cfg, gmap, m, cams = bench.build_workload("headline", 0, dev)
H, W = cfg["H"], cfg["W"]
bg = torch.zeros(3, device=dev)
mats = [c.matrices(dev) for c in cams]
targets = [torch.rand(3, H, W, device=dev) for _ in range(len(cams))]
params = [t.clone().requires_grad_(True) for t in (m.means3D, m.shs, m.opacities, m.scales, m.rotations)]
S, Rz = arm.pkg.GaussianRasterizationSettings, arm.pkg.GaussianRasterizer
loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
acc = [0.0] * 5
UPLOAD = "--upload" in sys.argv
if UPLOAD:   # bench.py's per-step H2D of the packed inputs on a copy stream
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    n_in = 36 + 3 * H * W
    host = [torch.rand(n_in).pin_memory() for _ in range(len(cams))]
    devs = [torch.empty(n_in, device=dev) for _ in range(2)]
    cs = torch.cuda.Stream(dev)
    done = [torch.cuda.Event(), torch.cuda.Event()]
    used = [torch.cuda.Event(), torch.cuda.Event()]
    for e in used:
        e.record()
def step(i, record):
    q, slot = i % len(cams), i % 2
    t0 = time.perf_counter()
    if UPLOAD:
        nslot = (i + 1) % 2
        cs.wait_event(used[nslot])
        rt.cudaMemcpyAsync(devs[nslot].data_ptr(), host[(i + 1) % len(cams)].data_ptr(), n_in * 4, 1, cs.cuda_stream)
        done[nslot].record(cs)
        torch.cuda.current_stream().wait_event(done[slot])
    view, proj, _, campos = mats[q]
    cam = cams[q]
    rs = S(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0, viewmatrix=view,
           projmatrix=proj, sh_degree=m.sh_degree, campos=campos, prefiltered=False, debug=False)
    m2 = torch.zeros_like(params[0], requires_grad=True)
    color, radii, depth, alpha = Rz(rs)(means3D=params[0], means2D=m2, opacities=params[2], shs=params[1], scales=params[3], rotations=params[4])
    t1 = time.perf_counter()
    loss = (color - targets[q]).abs().mean()
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    for p_ in params:
        p_.grad = None
    if UPLOAD:
        used[slot].record()
    loss_host[slot].copy_(loss.detach(), non_blocking=True)
    loss_ready[slot].record()
    loss_ready[1 - slot].synchronize()
    _ = float(loss_host[1 - slot])
    t4 = time.perf_counter()
    if record:
        for k, (a, b) in enumerate(((t0, t1), (t1, t2), (t2, t3), (t3, t4))):
            acc[k] += b - a
for i in range(10):
    step(i, False)
torch.cuda.synchronize()
n = 200
T0 = time.perf_counter()
for i in range(10, 10 + n):
    step(i, True)
torch.cuda.synchronize()
tot = time.perf_counter() - T0
print(f"step {tot/n*1e6:.0f} us: forward(incl. poll) {acc[0]/n*1e6:.0f}  loss {acc[1]/n*1e6:.0f}  backward {acc[2]/n*1e6:.0f}  tail(read prev loss) {acc[3]/n*1e6:.0f}")
