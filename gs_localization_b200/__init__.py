"""B200-native (sm_100a) differentiable 3D-Gaussian-splatting rasterizer for LoGS.

A from-scratch replacement for the one hot path of RPL-CS-UCL/gs_localization: the
depth+alpha `diff_gaussian_rasterization` extension (and the `diff_gaussian_rasterization_pose`
surface its localization scripts import).  Python host code -> C ABI (include/gsr_b200.h)
-> hand-written CUDA kernels (csrc/).  There is no CPU or PyTorch fallback: importing the
sub-packages works anywhere, calling them without the built library or a GPU raises.
"""
__version__ = "0.1.0"
