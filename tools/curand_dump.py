#!/usr/bin/env python
"""Dumps the first N outputs of cuRAND's HOST API MRG32K3A generator exactly as the reference
draws its sample list (GPU.cu:1443-1446: curandCreateGenerator(CURAND_RNG_PSEUDO_MRG32K3A),
seed 11, curandGenerate) so the ordering can be studied offline.  GPU box only."""
import ctypes as C
import sys

import numpy as np
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 11
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/curand_mrg32k3a.npy"
lib = C.CDLL("libcurand.so.10")
gen = C.c_void_p()
assert lib.curandCreateGenerator(C.byref(gen), 121) == 0          # CURAND_RNG_PSEUDO_MRG32K3A
assert lib.curandSetPseudoRandomGeneratorSeed(gen, C.c_ulonglong(seed)) == 0
buf = torch.empty(n, dtype=torch.int32, device="cuda")
assert lib.curandGenerate(gen, C.c_void_p(buf.data_ptr()), C.c_size_t(n)) == 0
torch.cuda.synchronize()
a = buf.cpu().numpy().view(np.uint32)
np.save(out, a)
print(a[:8], a.shape)
