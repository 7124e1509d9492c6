#!/usr/bin/env python
"""Does write-combined pinned memory (cudaHostAllocWriteCombined) feed the GPU faster than
ordinary pinned memory when D2H runs at the same time?  H2D 2 x 256 MiB + D2H 288 MiB
concurrently (the solvers' 64:36 byte mix), for both kinds of input buffers."""
import ctypes as C
import time
import torch

rt = C.CDLL("libcudart.so.12")
dev = torch.device("cuda:0")
torch.zeros(1, device=dev)
n_in, n_out = 256 << 20, 288 << 20


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags)) == 0
    return p


d_in = [torch.empty(n_in, dtype=torch.uint8, device=dev) for _ in range(2)]
d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
h_out = host_alloc(n_out, 0)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def run(h_in, with_d2h, reps=6):
    def once():
        for k in range(2):
            assert rt.cudaMemcpyAsync(C.c_void_p(d_in[k].data_ptr()), h_in[k], C.c_size_t(n_in), 1,
                                      C.c_void_p(s_in.cuda_stream)) == 0
        if with_d2h:
            assert rt.cudaMemcpyAsync(h_out, C.c_void_p(d_out.data_ptr()), C.c_size_t(n_out), 2,
                                      C.c_void_p(s_out.cuda_stream)) == 0
    once(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    return 2 * n_in * reps / (time.perf_counter() - t0) / 1e9


for name, flags in (("pinned (default)", 0), ("pinned, write-combined", 4)):
    h_in = [host_alloc(n_in, flags) for _ in range(2)]
    for k in range(2):
        C.memset(h_in[k], 1, n_in)
    print(f"{name:24s}: H2D alone {run(h_in, False):5.1f} GB/s | H2D with concurrent D2H {run(h_in, True):5.1f} GB/s", flush=True)
