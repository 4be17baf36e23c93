"""Generates tests/golden/ref_grad_mask.npz with the reference's image_gradient / image_gradient_mask
(/root/reference/gs_localization/pipelines/tools/descent_utils.py:33-66; they hard-code device="cuda", redirected to
the CPU here) driving the body of Camera.compute_grad_mask (tools/camera_utils.py:164-192)."""
import importlib.util
import os

import numpy as np
import torch

_tensor, _ones = torch.tensor, torch.ones


def _cpu_kw(fn):
    def w(*a, **k):
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return fn(*a, **k)
    return w


torch.tensor, torch.ones = _cpu_kw(_tensor), _cpu_kw(_ones)
spec = importlib.util.spec_from_file_location("ref_descent", "/root/reference/gs_localization/pipelines/tools/descent_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def ref_mask(img, thr, typ):
    gray = img.mean(dim=0, keepdim=True)
    gv, gh = ref.image_gradient(gray)
    mv, mh = ref.image_gradient_mask(gray)
    inten = torch.sqrt((gv * mv) ** 2 + (gh * mh) ** 2)
    if typ == "replica":
        row = col = 32
        _, h, w = img.shape
        for r in range(row):
            for c in range(col):
                block = inten[:, r * int(h / row):(r + 1) * int(h / row), c * int(w / col):(c + 1) * int(w / col)]
                th = block.median()
                block[block > th * thr] = 1
                block[block <= th * thr] = 0
        return inten
    return inten > inten.median() * thr


g = torch.Generator().manual_seed(4)
img = torch.nn.functional.avg_pool2d(torch.rand(3, 64, 96, generator=g)[None], 5, 1, 2)[0]
img[:, :10, :20] = 0.0
gv, gh = ref.image_gradient(img)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_grad_mask.npz"), img=img.numpy(), tum=ref_mask(img, 1.1, "tum").numpy(),
                    replica=ref_mask(img, 4, "replica").numpy(), grad_v=gv.numpy(), grad_h=gh.numpy())
