"""Generates tests/golden/*.npz by running the UNMODIFIED reference CUDA rasterizer
(oracle/_ref, built by oracle/build_ref.sh from /root/reference) on the GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

and copying gpurun_out/golden/*.npz into tests/golden/.  The reference ships no golden
vectors of its own, so these outputs of the reference itself are what pins
the CPU oracle (tests/test_oracle.py) and, through it and directly, the CUDA path.
Inputs are regenerated from seeds (gs_localization_b200/synthetic.py), only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import util  # noqa: E402

GOLDEN_SCENES = {
    "tiny": dict(P=500, W=48, H=32, deg=3, f=40.0, sigma0=0.2),
    "ragged": dict(P=3000, W=75, H=53, deg=2, f=70.0, sigma0=0.1),
    "C1": dict(P=10_000, W=160, H=120, deg=0, f=131.25, sigma0=0.05),
}
BG = [0.1, 0.3, 0.2]


def loss_weights(H, W, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((3, H, W)).astype(np.float32), (0.3 * rng.standard_normal((1, H, W))).astype(np.float32),
            (0.2 * rng.standard_normal((1, H, W))).astype(np.float32))


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = util.load_reference()
    dev = "cuda:0"
    for name, sc in GOLDEN_SCENES.items():
        m, cam = util.scene(**sc)
        bg = torch.tensor(BG)
        args = util.c_args(m, cam, bg, dev)
        R, color, depth, alpha, radii, geom, binning, img = ref._C.rasterize_gaussians(*args)
        torch.cuda.synchronize()
        P = m.means3D.shape[0]
        st = util.ref_unpack_state(P, R, cam.W, cam.H, geom, binning, img)
        wc, wd, wa = loss_weights(cam.H, cam.W)
        (bgt, means3D, colors, opac, scales, rots, smod, cov, view, proj, tfx, tfy, H, W, sh, deg, campos, pf, dbg) = args
        t = lambda a: torch.from_numpy(a).to(dev)
        res = ref._C.rasterize_gaussians_backward(bgt, means3D, radii, colors, scales, rots, smod, cov, view, proj, tfx, tfy,
                                                  t(wc), t(wd), t(wa), sh, deg, campos, geom, R, binning, img, alpha, False)
        torch.cuda.synchronize()
        names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
        vis = (radii > 0)
        out = dict(
            num_rendered=np.int64(R), radii=radii.cpu().numpy(), tiles_touched=st["tiles_touched"].cpu().numpy(),
            keys=st["keys"].cpu().numpy(), point_list=st["list"].cpu().numpy(), ranges=st["ranges"].cpu().numpy(),
            n_contrib=st["n_contrib"].cpu().numpy(),
            vis_depths=st["depths"][vis].cpu().numpy(), vis_means2D=st["means2D"][vis].cpu().numpy(),
            vis_conic_opacity=st["conic_opacity"][vis].cpu().numpy(), vis_cov3D=st["cov3D"][vis].cpu().numpy(),
            vis_rgb=st["rgb"][vis].cpu().numpy(),
            color=color.cpu().numpy().astype(np.float32), depth=depth.cpu().numpy(), alpha=alpha.cpu().numpy())
        for k, r in zip(names, res):
            a = r.cpu().numpy()
            out[k] = a[vis.cpu().numpy()] if a.shape[0] == P else a   # culled rows are zero: store visible rows only
        np.savez_compressed(os.path.join(out_dir, f"ref_{name}.npz"), **out)
        print(name, "R", R, "visible", int(vis.sum()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE))
