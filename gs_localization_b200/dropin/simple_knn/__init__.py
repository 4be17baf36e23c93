"""Import alias: `from simple_knn._C import distCUDA2` (gaussian_splatting/scene/gaussian_model.py:20) resolves to
the B200 implementation when `gs_localization_b200/dropin` is on PYTHONPATH."""
