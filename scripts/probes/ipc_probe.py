"""Probe (torchrun, one rank per GPU): does CUDA IPC work between the ranks on this box?  Rank r allocates a buffer,
exports cudaIpcMemHandle, every rank opens every peer's handle and writes its rank into slot [rank] of each peer's buffer
with cudaMemcpy (device-to-device over NVLink); after a barrier every rank checks that all slots arrived."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
    class Handle(C.Structure):          # cudaIpcMemHandle_t: 64 opaque bytes, passed BY VALUE
        _fields_ = [("reserved", C.c_char * 64)]

    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaIpcGetMemHandle.argtypes = [C.POINTER(Handle), C.c_void_p]
    rt.cudaIpcOpenMemHandle.argtypes = [C.POINTER(C.c_void_p), Handle, C.c_uint]
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    own = C.c_void_p()
    assert rt.cudaMalloc(C.byref(own), 4 * world) == 0      # a whole allocation of its own (not a caching-allocator slice)
    zeros = np.zeros(world, np.float32)
    assert rt.cudaMemcpy(own, zeros.ctypes.data_as(C.c_void_p), 4 * world, 1) == 0
    handle = Handle()
    e = rt.cudaIpcGetMemHandle(C.byref(handle), own)
    assert e == 0, "cudaIpcGetMemHandle -> %d" % e
    handles = [None] * world
    dist.all_gather_object(handles, bytes(handle))
    ptrs = []
    for r in range(world):
        if r == rank:
            ptrs.append(own.value)
            continue
        p = C.c_void_p()
        h = Handle.from_buffer_copy(handles[r])
        e = rt.cudaIpcOpenMemHandle(C.byref(p), h, 1)   # cudaIpcMemLazyEnablePeerAccess
        assert e == 0, "cudaIpcOpenMemHandle(rank %d) -> %d" % (r, e)
        ptrs.append(p.value)
    mine = torch.full((1,), float(rank + 1), dtype=torch.float32, device=dev)
    for r in range(world):
        e = rt.cudaMemcpy(C.c_void_p(ptrs[r] + 4 * rank), C.c_void_p(mine.data_ptr()), 4, 3)
        assert e == 0, "cudaMemcpy to rank %d -> %d" % (r, e)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    got = np.zeros(world, np.float32)
    assert rt.cudaMemcpy(got.ctypes.data_as(C.c_void_p), own, 4 * world, 2) == 0
    ok = bool(np.array_equal(got, np.arange(1, world + 1, dtype=np.float32)))
    print("rank %d: ipc %s %s" % (rank, "OK" if ok else "FAILED", got.tolist()), flush=True)
    dist.barrier()
    for r in range(world):
        if r != rank:
            rt.cudaIpcCloseMemHandle(C.c_void_p(ptrs[r]))
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
