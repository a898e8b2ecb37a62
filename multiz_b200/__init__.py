"""multiz_b200 -- B200 (sm_100a) implementation of multiz's `yama` hot path.

Only the hot path lives here: `csrc/` (CUDA kernels + the C ABI of include/yama_b200.h) and
`yama.py`, a thin ctypes mirror of the reference's yama() interface (mz_yama.h:4-22) used by the
tests and the benchmark.  The C host of multiz (multiz/tba/roast command lines, MAF I/O,
pre_yama, stitching) stays the reference's own code and links against libyama_b200.so through
integration/ (see INTEGRATION.md).
"""
from .yama import (YamaB200, YamaError, lib_path, load_library, ABI_SYMBOLS, hox70_tables, plan_split, JOB_DTYPE, RESULT_DTYPE)  # noqa: F401
