"""Where does the start-up second of a GPU-backed multiz go?  Times, each in a fresh process: the CUDA runtime's own
initialisation (cudaGetDeviceCount, cudaSetDevice + cudaFree(0)) and yb_create() on top of it.
    python tools/startup_probe.py            (on the GPU box)
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(kind):
    t0 = time.perf_counter()
    out = {}
    if kind == "runtime":
        rt = C.CDLL("libcudart.so.12")
        n = C.c_int()
        rt.cudaGetDeviceCount(C.byref(n)); out["get_device_count_ms"] = (time.perf_counter() - t0) * 1e3
        t1 = time.perf_counter()
        rt.cudaSetDevice(0); rt.cudaFree(None); out["context_ms"] = (time.perf_counter() - t1) * 1e3
    else:
        lib = C.CDLL(os.path.join(ROOT, "multiz_b200", "libyama_b200.so"))
        out["dlopen_ms"] = (time.perf_counter() - t0) * 1e3
        h = C.c_void_p()
        t1 = time.perf_counter()
        dev = (C.c_int * 1)(0)
        rc = lib.yb_create(dev, 1, C.byref(h)); out["yb_create_ms"] = (time.perf_counter() - t1) * 1e3
        out["rc"] = rc
    out["total_ms"] = (time.perf_counter() - t0) * 1e3
    print(json.dumps(out))
    os._exit(0)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    try:
        print(subprocess.run("nvidia-smi -q | grep -i -m2 persistence", shell=True, capture_output=True, text=True).stdout.strip())
    except Exception:
        pass
    for kind in ("runtime", "yb", "runtime", "yb", "yb"):
        t0 = time.perf_counter()
        p = subprocess.run([sys.executable, __file__, kind], capture_output=True, text=True)
        print(kind, p.stdout.strip(), "process_wall_ms=%.0f" % ((time.perf_counter() - t0) * 1e3))
