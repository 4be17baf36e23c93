"""cProfile of the public-API forward+backward on a tiny scene (host cost only)."""
import cProfile, pstats, os, sys, io
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from gs_localization_b200 import synthetic as syn
dev = torch.device("cuda:0")
arm = bench.Arm("ours", dev)
cfg = dict(P=2000, W=64, H=48, deg=3, f=50.0, box=1.0, sigma0=0.1)
m = syn.make_map(cfg["P"], 3, cfg["sigma0"], 1.0).to(dev)
cam = syn.make_camera(cfg, 0)
view, proj, raw, campos = cam.matrices(dev)
bg = torch.zeros(3, device=dev)
params = [t.clone().requires_grad_(True) for t in (m.means3D, m.shs, m.opacities, m.scales, m.rotations)]
S, Rz = arm.pkg.GaussianRasterizationSettings, arm.pkg.GaussianRasterizer
def step_api():
    rs = S(image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
           viewmatrix=view, projmatrix=proj, sh_degree=3, campos=campos, prefiltered=False, debug=False)
    m2 = torch.zeros_like(params[0], requires_grad=True)
    out = Rz(rs)(means3D=params[0], means2D=m2, opacities=params[2], shs=params[1], scales=params[3], rotations=params[4])
    out[0].sum().backward()
for _ in range(50): step_api()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300): step_api()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
