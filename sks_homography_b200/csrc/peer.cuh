// One-shot max-reduce of the packed RANSAC keys over NVLink peer memory: a
// hand-written replacement for the 8-KiB NCCL all-reduce of the multi-GPU RANSAC
// step.  Every rank owns an exchange block that all peers map
// through CUDA IPC.  A step is two small kernels per rank and no collective call:
//
//   k_peer_push_max  zero the block's NEXT key buffer, push the local keys into
//                    every rank's CURRENT buffer with system-scope atomicMax
//                    (NVLink atomics, performed at the owner's L2), fence, and let
//                    the last CTA add 1 to every rank's arrival counter;
//   k_peer_wait_copy spin (bounded) until the own arrival counter shows that all
//                    `world` ranks have pushed this epoch, then copy the reduced
//                    keys out.
//
// Buffers alternate by epoch parity and the arrival counter is monotonic, so
// nothing is ever reset while a peer may still touch it: a rank enters epoch e+1
// only after every rank's arrival for epoch e, and each rank zeroes its (e+1)
// buffer before it signals arrival e.  No kernel waits on a kernel of the SAME GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sksb {

constexpr int kMaxPeers = 16;

struct PeerBlock {                      // header of an exchange block (device memory)
    unsigned long long arrive;          // += 1 per rank per epoch, by system-scope atomics
    unsigned int done;                  // CTAs of the local push kernel that finished
    unsigned int pad;
    unsigned long long n_keys;
    unsigned long long reserved;
    // followed by unsigned long long keys[2][n_keys]
};

struct PeerTable {
    PeerBlock* blk[kMaxPeers];
};

__device__ __forceinline__ unsigned long long* peer_keys(PeerBlock* b, uint64_t epoch, int64_t n_keys)
{
    return reinterpret_cast<unsigned long long*>(b + 1) + (epoch & 1) * n_keys;
}

__global__ void __launch_bounds__(256)
k_peer_push_max(const unsigned long long* __restrict__ local_keys, int64_t n_keys, PeerTable tab,
                int world, int rank, uint64_t epoch)
{
    PeerBlock* own = tab.blk[rank];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_keys) {
        peer_keys(own, epoch + 1, n_keys)[i] = 0ull;          // next epoch's buffer
        const unsigned long long k = local_keys[i];
        for (int g = 0; g < world; ++g)
            atomicMax_system(peer_keys(tab.blk[g], epoch, n_keys) + i, k);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&own->done, 1u);
        if (prev == gridDim.x - 1) {                          // last CTA of this rank's push
            own->done = 0u;
            __threadfence_system();
            for (int g = 0; g < world; ++g)
                atomicAdd_system(&tab.blk[g]->arrive, 1ull);
        }
    }
}

// status: set to 1 (and never cleared here) when the wait timed out; keys_out is then left
// UNTOUCHED -- a partial max must not be mistaken for the reduced keys.  After a timeout the
// exchange is unusable (the arrival counters no longer line up): the caller has to treat it as
// fatal (dist.PeerReducer raises) and rebuild the blocks or fall back to the NCCL all-reduce.
__global__ void __launch_bounds__(256)
k_peer_wait_copy(PeerBlock* own, int world, uint64_t epoch, unsigned long long* __restrict__ keys_out,
                 int64_t n_keys, int* __restrict__ status, long long timeout_cycles)
{
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const unsigned long long target = (unsigned long long)world * (epoch + 1);
        const long long t0 = clock64();
        int good = 1;
        while (*reinterpret_cast<volatile unsigned long long*>(&own->arrive) < target) {
            if (clock64() - t0 > timeout_cycles) {
                good = 0;
                break;
            }
            __nanosleep(100);
        }
        __threadfence_system();
        ok = good;
        if (status != nullptr && !good) {
            *reinterpret_cast<volatile int*>(status) = 1;
            __threadfence_system();
        }
    }
    __syncthreads();
    if (!ok)
        return;
    const unsigned long long* src = peer_keys(own, epoch, n_keys);
    for (int64_t i = threadIdx.x; i < n_keys; i += blockDim.x)
        keys_out[i] = __ldcg(src + i);
}

}  // namespace sksb
