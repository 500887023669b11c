"""Experiment: set cudaLimitMaxL2FetchGranularity before anything else touches the GPU, then run a bench config."""
import ctypes, glob, os, sys, runpy
gran = int(os.environ.get("L2GRAN", "0"))
if gran:
    import torch  # only to locate the bundled runtime
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    rt = ctypes.CDLL(cands[0])
    rc = rt.cudaDeviceSetLimit(ctypes.c_int(5), ctypes.c_size_t(gran))     # cudaLimitMaxL2FetchGranularity = 0x05
    v = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(v), ctypes.c_int(5))
    print("set L2 fetch granularity", gran, "rc", rc, "now", v.value, file=sys.stderr)
sys.argv = ["bench.py"] + sys.argv[1:]
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench.py"), run_name="__main__")
