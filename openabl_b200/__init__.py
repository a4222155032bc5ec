"""openabl_b200 — Python harness around the B200-native OpenABL `cuda` backend.

The product is native: the compiler (`src/`, C++), the device runtime behind a C ABI
(`asset/cuda`, `include/abl_cuda.h`) and the kernels it generates.  This package only
drives them for tests and benchmarks: it builds the pieces in-tree, compiles `.abl`
models through the real CLI, and binds the C ABI with ctypes.  It never imports the
parity oracle (`oracle/`) and has no CPU fallback.
"""
from .paths import REPO_ROOT, ASSET_DIR, COMPILER, RUNTIME_LIB  # noqa: F401
