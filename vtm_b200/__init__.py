"""vtm_b200 — B200-native motion search for VTM 9.3 (integer full search + quarter-pel refinement).

The product is the C-ABI CUDA library libvtmme.so (include/vtmme.h); this package is the thin host side
used by the tests and the standalone batched-ME benchmark.  There is no CPU fallback: importing works
anywhere, but every compute entry point needs the built library and a CUDA device.
"""
from .lib import VtmmeError, load_library, library_path, build_library  # noqa: F401
from .me import Amvr, MotionSearch, FrameParams, Job, TzSearch  # noqa: F401
