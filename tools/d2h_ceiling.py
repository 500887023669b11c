#!/usr/bin/env python
"""Device-to-host copy ceiling of the box, all ranks at once (run under torchrun): what the host-buffer (e2e) path is bounded
by.  Every rank copies a 64 MiB device buffer into pinned host memory (default / write-combined; NUMA-bound or not) in a
loop for ~0.5 s; prints per-rank and aggregate GB/s."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from xenoverse_b200 import HostBuffer, bind_to_gpu_numa_node
    from xenoverse_b200.engine import lib  # noqa: F401
    import ctypes
    cudart = ctypes.CDLL("libcudart.so.12")
    nbytes = 64 << 20
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}
    for label, bind, wc in (("default", False, False), ("wc", False, True), ("numa", True, False), ("numa+wc", True, True)):
        info = bind_to_gpu_numa_node(local) if bind else None
        hb = HostBuffer(nbytes, wc)
        st = torch.cuda.Stream()
        def copy():
            cudart.cudaMemcpyAsync(ctypes.c_void_p(hb.ptr), ctypes.c_void_p(src.data_ptr()), ctypes.c_size_t(nbytes), 2, ctypes.c_void_p(st.cuda_stream))
        for _ in range(3):
            copy()
        st.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 0.5:
            for _ in range(8):
                copy()
            st.synchronize()
            n += 8
        dt = time.perf_counter() - t0
        gbs = torch.tensor([n * nbytes / dt / 1e9], dtype=torch.float64, device=dev)
        tot = gbs.clone()
        if world > 1:
            dist.all_reduce(tot)
        out[label] = (float(gbs), float(tot), info)
        hb.close()
        if world > 1:
            dist.barrier()
    if rank == 0:
        for k, (g, t, info) in out.items():
            print(f"d2h {k:8s}: rank0 {g:6.1f} GB/s, aggregate over {world} ranks {t:7.1f} GB/s  {info or ''}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
