"""Times distCUDA2 (gsr_dist2_knn3) against the reference build (oracle/_ref/simple_knn) on SfM-like clouds."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from gs_localization_b200.simple_knn._C import distCUDA2  # noqa: E402
from test_knn import _cloud  # noqa: E402


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


ref = None
if os.path.exists("oracle/_ref/simple_knn/_C.so"):
    sys.path.insert(0, "oracle/_ref")
    from simple_knn._C import distCUDA2 as ref  # noqa: E402
for P, kind in [(100_000, "surface"), (1_000_000, "surface"), (1_000_000, "uniform"), (3_000_000, "surface")]:
    pts = torch.from_numpy(_cloud(P, 1, kind)).cuda()
    row = {"tool": "knn_bench", "P": P, "kind": kind, "ours_ms": timed(lambda: distCUDA2(pts))}
    if ref is not None:
        row["reference_ms"] = timed(lambda: ref(pts), n=2)
        row["speedup"] = row["reference_ms"] / row["ours_ms"]
        row["bit_exact"] = bool(torch.equal(distCUDA2(pts), ref(pts)))
    print(json.dumps(row), flush=True)
