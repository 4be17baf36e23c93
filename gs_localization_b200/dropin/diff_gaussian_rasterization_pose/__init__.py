"""Import alias for the pose variant imported by gs_localization/pipelines/tools/__init__.py:15-18."""
from gs_localization_b200.diff_gaussian_rasterization_pose import (  # noqa: F401
    _RasterizeGaussiansPose, GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians)
