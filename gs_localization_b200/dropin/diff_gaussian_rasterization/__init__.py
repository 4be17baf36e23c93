"""Import alias: put `gs_localization_b200/dropin` on PYTHONPATH and the reference's
`from diff_gaussian_rasterization import ...` resolves to the B200 implementation."""
from gs_localization_b200.diff_gaussian_rasterization import *  # noqa: F401,F403
from gs_localization_b200.diff_gaussian_rasterization import (  # noqa: F401
    _C, _RasterizeGaussians, GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians)
