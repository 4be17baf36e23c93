"""Import FIRST in a measurement tool: GSR_AB_LIB=/path/to/other/libgsr_b200.so makes gs_localization_b200 load that build
of the library (through the ctypes binding) instead of the in-tree one — same-box A/B of kernel variants built by
tests/tools/build_ab_lib.sh.  Device times (CUDA events / stage split) do not depend on the binding."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.environ.get("GSR_AB_LIB"):
    os.environ["GSR_BINDING"] = "ctypes"
    from gs_localization_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(os.environ["GSR_AB_LIB"])
    import ctypes as _C
    _probe = _C.CDLL(_l.LIB_PATH)
    for _name in list(_l.SIGNATURES):
        if not hasattr(_probe, _name):
            del _l.SIGNATURES[_name]          # symbols added after that build
