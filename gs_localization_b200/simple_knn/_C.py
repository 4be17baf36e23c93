"""`simple_knn._C` on the B200 library: `distCUDA2(points)` as GaussianModel.create_from_pcd calls it
(gaussian_splatting/scene/gaussian_model.py:20,135; binding spatial.cu:15-26)."""
import torch

from .. import _lib


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """points [P,3] float32 CUDA -> [P] mean squared distance to the 3 nearest other points."""
    if not points.is_cuda:
        raise RuntimeError("gs_localization_b200: tensors must be CUDA tensors (no CPU fallback exists)")
    lib = _lib.load()
    pts = points.contiguous().float()
    P = int(pts.shape[0])
    out = torch.zeros(P, dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    with torch.cuda.device(pts.device):
        ws = torch.empty(lib.gsr_knn_workspace_bytes(P), dtype=torch.uint8, device=pts.device)
        _lib.check(lib.gsr_dist2_knn3(pts.data_ptr(), P, out.data_ptr(), ws.data_ptr(),
                                      torch.cuda.current_stream(pts.device).cuda_stream), "gsr_dist2_knn3")
    return out
