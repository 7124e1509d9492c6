// In-process multi-GPU fused ACA-RANSAC behind the C ABI (SURVEY.md 8(b) "multi-GPU driver
// entry sks_cuda_*_multi(..., int ngpu)", 8(e) variant A): ONE process, ONE enqueueing host
// thread, no NCCL and no CUDA IPC.  The hypothesis ids of every image pair are cut into ngpu
// contiguous shards; device k scores its shard with the unmodified fused kernel
// (sks_cuda_ransac_aca_shard_f32) and max-combines its per-pair winners into the primary
// device's key array with system-scope atomicMax over NVLink (k_multi_push_max, 8 KiB for 1024
// pairs).  The correspondences reach the peers by a binomial-tree broadcast of peer copies
// over NVLink (device 0 -> 1, then 0 -> 2 and 1 -> 3, then 0..3 -> 4..7: log2(ngpu) stages of
// 64 MiB at ~0.09 ms each, every device starting to score as soon as its copy has landed; the
// host-pointer entry copies the matches to the primary once, over one PCIe link -- eight H2D
// copies of the same 64 MiB would only contend for the host's memory system).  Reading them in
// place over peer access instead (the tile bulk copies and sample gathers of k_ransac_aca accept
// peer addresses, and tests/test_gpu_multi.py ran that way at 2 GPUs) makes every CTA re-fetch
// its pair's 64 KiB tile from the primary: at 8 GPUs with 16 chunks per pair that is 7 GiB of
// NVLink egress from one device per step, 20.2 ms against 12.2 ms (profiles/r02_bench_n8.json),
// so only an explicit sample list -- 16 B per hypothesis, read once -- is still read in place.
// Ordering is all CUDA events between the devices' streams: copies start after the caller's
// stream has reached the call, the primary's finalize starts after every peer's push.  The
// primary then rebuilds the winners from their ids -- no homography travels.  Bit-identical to
// one GPU by construction (integer max of the same keys).
#include <cstdint>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sks_cuda.h"

// pageable-aware H2D copy shared with csrc/host_api.cu
struct SksStageBuf {
    void* p[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
};
extern "C" int sks_stage_h2d_internal(void* dst, const void* src, size_t bytes, cudaStream_t st, SksStageBuf* sb);
extern "C" void sks_stage_free_internal(SksStageBuf* sb);

namespace {

#define CK(x)                                  \
    do {                                       \
        cudaError_t _e = (x);                  \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

__global__ void __launch_bounds__(256)
k_multi_push_max(const unsigned long long* __restrict__ local_keys, unsigned long long* primary_keys,
                 int64_t n_keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_keys) {
        const unsigned long long k = local_keys[i];
        if (k != 0ull)
            atomicMax_system(primary_keys + i, k);   // performed at the owner's L2, over NVLink
    }
}

struct PeerDev {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t copied = nullptr;         // this device's copy of the matches has landed
    unsigned long long* keys = nullptr;   // this device's per-pair winners
    size_t keys_cap = 0;
    float* corr = nullptr;                // this device's copy of the matches
    size_t corr_cap = 0;
};

struct MultiCtx {
    int primary = -1;
    std::mutex mu;                        // one multi-GPU call at a time per primary device
    cudaEvent_t ready = nullptr;          // caller's stream reached the call
    cudaStream_t stream = nullptr;        // primary's own stream (host-pointer entry)
    std::vector<PeerDev> peers;           // devices primary+1 .. (mod visible), grown on demand
    void* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // host entry, primary: H, cnt, key, mask, (unused)
    size_t cap[5] = {0, 0, 0, 0, 0};
    float* corr = nullptr;
    size_t corr_cap = 0;
    uint32_t* samples = nullptr;
    size_t samples_cap = 0;
    SksStageBuf stage;
};

std::mutex g_mu;
std::vector<MultiCtx*> g_ctx;

int grow(void** p, size_t* cap, size_t need)
{
    if (need <= *cap) return SKS_OK;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    CK(cudaMalloc(p, need));
    *cap = need;
    return SKS_OK;
}

// Context of `primary` with at least ngpu - 1 peers: streams, events and peer access
// (peer -> primary: the peers read the primary's memory and push keys into it).
int get_ctx(int primary, int ngpu, int visible, MultiCtx** out)
{
    std::lock_guard<std::mutex> table(g_mu);
    MultiCtx* c = nullptr;
    for (MultiCtx* x : g_ctx)
        if (x->primary == primary) c = x;
    if (c == nullptr) {
        c = new MultiCtx();
        c->primary = primary;
        CK(cudaSetDevice(primary));
        cudaError_t e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            if (c->ready) cudaEventDestroy(c->ready);
            delete c;
            return (int)e;
        }
        g_ctx.push_back(c);
    }
    while ((int)c->peers.size() < ngpu - 1) {
        PeerDev p;
        p.device = (primary + 1 + (int)c->peers.size()) % visible;
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, p.device, primary));
        if (!can) return SKS_ERR_NO_PEER_ACCESS;
        CK(cudaSetDevice(p.device));
        cudaError_t e = cudaDeviceEnablePeerAccess(primary, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) {
            cudaGetLastError();
            e = cudaSuccess;
        }
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.copied, cudaEventDisableTiming);
        // broadcast tree: peer index k = peers.size() + 1 pulls from k - 2^floor(log2 k)
        const int k = (int)c->peers.size() + 1;
        int top = 1;
        while (top * 2 <= k) top *= 2;
        const int parent = k - top;
        if (e == cudaSuccess && parent > 0) {
            e = cudaDeviceEnablePeerAccess(c->peers[parent - 1].device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                e = cudaSuccess;
            }
        }
        if (e != cudaSuccess) {
            if (p.stream) cudaStreamDestroy(p.stream);
            if (p.done) cudaEventDestroy(p.done);
            if (p.copied) cudaEventDestroy(p.copied);
            cudaSetDevice(primary);
            return (int)e;
        }
        c->peers.push_back(p);
    }
    cudaSetDevice(primary);
    *out = c;
    return SKS_OK;
}

// Enqueue the sharded scoring + merge.  corr / samples / best_key live on the primary device; `st`
// is the primary's stream; on return (all asynchronous) `st` is ordered after every peer's push,
// and best_key holds the merged winners once `st` gets there.
int enqueue_sharded(MultiCtx* c, int ngpu, const float* corr, const uint32_t* samples, int64_t n_pairs,
                    int32_t n_pts, uint32_t n_hyp, uint64_t seed, float thr2, unsigned long long* best_key,
                    cudaStream_t st)
{
    const size_t key_bytes = (size_t)n_pairs * sizeof(unsigned long long);
    const size_t corr_bytes = (size_t)n_pairs * n_pts * 4 * sizeof(float);
    CK(cudaSetDevice(c->primary));
    CK(cudaMemsetAsync(best_key, 0, key_bytes, st));
    CK(cudaEventRecord(c->ready, st));       // inputs valid and keys zeroed from here on
    int rc = SKS_OK;
    // binomial-tree broadcast of the matches: peer k pulls from k - 2^floor(log2 k) on its own
    // stream, after the source holds the data AND has finished its previous send (so that a
    // source's NVLink egress serves one child at a time: log2(ngpu) stages in all)
    std::vector<cudaEvent_t> last_send(ngpu, nullptr);
    for (int k = 1; k < ngpu && rc == SKS_OK; ++k) {
        PeerDev& p = c->peers[k - 1];
        int top = 1;
        while (top * 2 <= k) top *= 2;
        const int parent = k - top;
        const float* src = parent == 0 ? corr : c->peers[parent - 1].corr;
        const int src_dev = parent == 0 ? c->primary : c->peers[parent - 1].device;
        if ((rc = (int)cudaSetDevice(p.device)) != SKS_OK) break;
        if ((rc = grow(reinterpret_cast<void**>(&p.corr), &p.corr_cap, corr_bytes)) != SKS_OK) break;
        if ((rc = (int)cudaStreamWaitEvent(p.stream, parent == 0 ? c->ready : c->peers[parent - 1].copied, 0)) != SKS_OK)
            break;
        if (last_send[parent] != nullptr &&
            (rc = (int)cudaStreamWaitEvent(p.stream, last_send[parent], 0)) != SKS_OK)
            break;
        if ((rc = (int)cudaMemcpyPeerAsync(p.corr, p.device, src, src_dev, corr_bytes, p.stream)) != SKS_OK) break;
        if ((rc = (int)cudaEventRecord(p.copied, p.stream)) != SKS_OK) break;
        last_send[parent] = p.copied;
    }
    for (int k = 1; k < ngpu && rc == SKS_OK; ++k) {
        PeerDev& p = c->peers[k - 1];
        int64_t hb = 0, hc = 0;
        sks_cuda_shard_range(n_hyp, k, ngpu, &hb, &hc);
        if ((rc = (int)cudaSetDevice(p.device)) != SKS_OK) break;
        if ((rc = grow(reinterpret_cast<void**>(&p.keys), &p.keys_cap, key_bytes)) != SKS_OK) break;
        if ((rc = (int)cudaStreamWaitEvent(p.stream, c->ready, 0)) != SKS_OK) break;
        if ((rc = (int)cudaMemsetAsync(p.keys, 0, key_bytes, p.stream)) != SKS_OK) break;
        rc = sks_cuda_ransac_aca_shard_f32(p.corr, 0, n_pairs, n_pts, samples, n_hyp, (uint32_t)hb,
                                           (uint32_t)hc, seed, thr2, p.keys, p.stream);
        if (rc != SKS_OK) break;
        k_multi_push_max<<<(unsigned)((n_pairs + 255) / 256), 256, 0, p.stream>>>(p.keys, best_key, n_pairs);
        if ((rc = (int)cudaGetLastError()) != SKS_OK) break;
        rc = (int)cudaEventRecord(p.done, p.stream);
    }
    cudaSetDevice(c->primary);
    if (rc == SKS_OK) {
        int64_t hb = 0, hc = 0;
        sks_cuda_shard_range(n_hyp, 0, ngpu, &hb, &hc);
        rc = sks_cuda_ransac_aca_shard_f32(corr, 0, n_pairs, n_pts, samples, n_hyp, (uint32_t)hb,
                                           (uint32_t)hc, seed, thr2, best_key, st);
    }
    // even after a failure: whatever was enqueued on the peers must be ordered before the
    // caller's stream continues (their atomics target the caller's key array)
    for (int k = 1; k < ngpu; ++k) {
        const cudaError_t e = cudaStreamWaitEvent(st, c->peers[k - 1].done, 0);
        if (rc == SKS_OK && e != cudaSuccess) rc = (int)e;
    }
    return rc;
}

int clamp_gpus(int ngpu, int* visible)
{
    if (int rc = sks_cuda_device_count(visible)) return rc;
    if (*visible <= 0) return SKS_ERR_NO_DEVICE;
    return SKS_OK;
}

}  // namespace

extern "C" {

int sks_cuda_ransac_aca_multi_f32(const float* corr, int64_t n_pairs, int32_t n_pts, const uint32_t* samples,
                                  uint32_t n_hyp, uint64_t seed, float thr2, int ngpu,
                                  unsigned long long* best_key, float* H_best, uint32_t* inlier_count,
                                  uint8_t* inlier_mask, void* stream)
{
    if (corr == nullptr || best_key == nullptr || n_pairs < 0 || n_pts <= 0 || n_hyp == 0 || ngpu < 0)
        return SKS_ERR_INVALID_ARG;
    int visible = 0;
    if (int rc = clamp_gpus(ngpu, &visible)) return rc;
    if (ngpu == 0 || ngpu > visible) ngpu = visible;
    if ((uint32_t)ngpu > n_hyp) ngpu = (int)n_hyp;
    if (n_pairs == 0) return SKS_OK;
    int primary = 0;
    CK(cudaGetDevice(&primary));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = SKS_OK;
    if (ngpu == 1) {
        CK(cudaMemsetAsync(best_key, 0, (size_t)n_pairs * sizeof(unsigned long long), st));
        rc = sks_cuda_ransac_aca_f32(corr, n_pairs, n_pts, samples, n_hyp, 0, n_hyp, seed, thr2, best_key, st);
    } else {
        MultiCtx* c = nullptr;
        if ((rc = get_ctx(primary, ngpu, visible, &c)) != SKS_OK) return rc;
        std::lock_guard<std::mutex> lk(c->mu);
        rc = enqueue_sharded(c, ngpu, corr, samples, n_pairs, n_pts, n_hyp, seed, thr2, best_key, st);
    }
    if (rc != SKS_OK) return rc;
    if (H_best != nullptr)
        rc = sks_cuda_ransac_finalize_f32(corr, n_pairs, n_pts, samples, n_hyp, seed, thr2, best_key, H_best,
                                          inlier_count, inlier_mask, st);
    return rc;
}

// Host-pointer form of the above, used by sks_host_ransac_aca_f32 when sks_host_set_device_count(g != 1):
// one H2D copy to the current device, then exactly the device-pointer flow.
int sks_host_ransac_aca_multi_f32(const float* corr, int64_t n_pairs, int32_t n_pts, const uint32_t* samples,
                                  uint32_t n_hyp, uint64_t seed, float thr2, int ngpu, float* H_best,
                                  uint32_t* inlier_count, uint8_t* inlier_mask, unsigned long long* best_key)
{
    if (corr == nullptr || H_best == nullptr || n_pairs < 0 || n_pts <= 0 || n_hyp == 0 || ngpu < 0)
        return SKS_ERR_INVALID_ARG;
    int visible = 0;
    if (int rc = clamp_gpus(ngpu, &visible)) return rc;
    if (ngpu == 0 || ngpu > visible) ngpu = visible;
    if ((uint32_t)ngpu > n_hyp) ngpu = (int)n_hyp;
    if (n_pairs == 0) return SKS_OK;
    int primary = 0;
    CK(cudaGetDevice(&primary));
    MultiCtx* c = nullptr;
    if (int rc = get_ctx(primary, ngpu, visible, &c)) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t corr_bytes = (size_t)n_pairs * n_pts * 4 * sizeof(float);
    const size_t samp_bytes = samples ? (size_t)n_pairs * n_hyp * 4 * sizeof(uint32_t) : 0;
    const size_t need[4] = {(size_t)n_pairs * 9 * sizeof(float), (size_t)n_pairs * sizeof(uint32_t),
                            (size_t)n_pairs * sizeof(unsigned long long),
                            inlier_mask ? (size_t)n_pairs * n_pts : 0};
    for (int k = 0; k < 4; ++k)
        if (int rc = grow(&c->buf[k], &c->cap[k], need[k])) return rc;
    if (int rc = grow(reinterpret_cast<void**>(&c->corr), &c->corr_cap, corr_bytes)) return rc;
    if (int rc = grow(reinterpret_cast<void**>(&c->samples), &c->samples_cap, samp_bytes)) return rc;
    cudaStream_t st = c->stream;
    float* d_H = static_cast<float*>(c->buf[0]);
    uint32_t* d_cnt = static_cast<uint32_t*>(c->buf[1]);
    unsigned long long* d_key = static_cast<unsigned long long*>(c->buf[2]);
    uint8_t* d_mask = inlier_mask ? static_cast<uint8_t*>(c->buf[3]) : nullptr;

    int rc = SKS_OK;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == SKS_OK) rc = (int)e; return e == cudaSuccess; };
    rc = sks_stage_h2d_internal(c->corr, corr, corr_bytes, st, &c->stage);
    if (rc == SKS_OK && samples) rc = sks_stage_h2d_internal(c->samples, samples, samp_bytes, st, &c->stage);
    if (rc == SKS_OK)
        rc = enqueue_sharded(c, ngpu, c->corr, samples ? c->samples : nullptr, n_pairs, n_pts, n_hyp, seed, thr2,
                             d_key, st);
    if (rc == SKS_OK)
        rc = sks_cuda_ransac_finalize_f32(c->corr, n_pairs, n_pts, samples ? c->samples : nullptr, n_hyp, seed, thr2,
                                          d_key, d_H, d_cnt, d_mask, st);
    if (rc == SKS_OK) cu(cudaMemcpyAsync(H_best, d_H, need[0], cudaMemcpyDeviceToHost, st));
    if (rc == SKS_OK && inlier_count) cu(cudaMemcpyAsync(inlier_count, d_cnt, need[1], cudaMemcpyDeviceToHost, st));
    if (rc == SKS_OK && inlier_mask) cu(cudaMemcpyAsync(inlier_mask, d_mask, need[3], cudaMemcpyDeviceToHost, st));
    if (rc == SKS_OK && best_key) cu(cudaMemcpyAsync(best_key, d_key, need[2], cudaMemcpyDeviceToHost, st));
    // success or not: nothing of this call may still be in flight when it returns
    for (int k = 1; k < ngpu; ++k) {
        cudaSetDevice(c->peers[k - 1].device);
        cu(cudaStreamSynchronize(c->peers[k - 1].stream));
    }
    cudaSetDevice(primary);
    cu(cudaStreamSynchronize(st));
    if (rc != SKS_OK) cudaGetLastError();
    return rc;
}

// called by sks_cuda_shutdown (host_api.cu)
void sks_multi_shutdown_internal(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int prev = 0;
    cudaGetDevice(&prev);
    for (MultiCtx* c : g_ctx) {
        { std::lock_guard<std::mutex> busy(c->mu); }
        for (PeerDev& p : c->peers) {
            cudaSetDevice(p.device);
            if (p.stream) { cudaStreamSynchronize(p.stream); cudaStreamDestroy(p.stream); }
            if (p.done) cudaEventDestroy(p.done);
            if (p.copied) cudaEventDestroy(p.copied);
            if (p.keys) cudaFree(p.keys);
            if (p.corr) cudaFree(p.corr);
        }
        cudaSetDevice(c->primary);
        sks_stage_free_internal(&c->stage);
        if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
        if (c->ready) cudaEventDestroy(c->ready);
        for (void* b : c->buf)
            if (b) cudaFree(b);
        if (c->corr) cudaFree(c->corr);
        if (c->samples) cudaFree(c->samples);
        delete c;
    }
    g_ctx.clear();
    cudaSetDevice(prev);
}

}  // extern "C"
