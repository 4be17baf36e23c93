"""Host-side cost of one forward+backward through the public API on a tiny scene (GPU time negligible)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from gs_localization_b200 import synthetic as syn

impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
dev = torch.device("cuda:0")
arm = bench.Arm(impl, dev)
cfg = dict(P=2000, W=64, H=48, deg=3, f=50.0, box=1.0, sigma0=0.1)
m = syn.make_map(cfg["P"], 3, cfg["sigma0"], 1.0).to(dev)
cam = syn.make_camera(cfg, 0)
view, proj, raw, campos = cam.matrices(dev)
bg = torch.zeros(3, device=dev)
zD = torch.zeros(1, cam.H, cam.W, device=dev)
gC = torch.ones(3, cam.H, cam.W, device=dev)
def step_c():
    fwd = arm.c_forward(m, bg, view, proj, campos, cam)
    arm.c_backward(m, bg, view, proj, campos, cam, fwd, gC, zD, zD)
params = [t.clone().requires_grad_(True) for t in (m.means3D, m.shs, m.opacities, m.scales, m.rotations)]
S, Rz = arm.pkg.GaussianRasterizationSettings, arm.pkg.GaussianRasterizer
def step_api():
    rs = S(image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
           viewmatrix=view, projmatrix=proj, sh_degree=3, campos=campos, prefiltered=False, debug=False)
    m2 = torch.zeros_like(params[0], requires_grad=True)
    out = Rz(rs)(means3D=params[0], means2D=m2, opacities=params[2], shs=params[1], scales=params[3], rotations=params[4])
    out[0].sum().backward()
for name, fn in (("_C fwd+bwd", step_c), ("public API fwd+bwd", step_api)):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 300
    for _ in range(n): fn()
    torch.cuda.synchronize()
    print(f"{impl}: {name}: {(time.perf_counter()-t0)/n*1e6:.1f} us/step")
