// Host-pointer entry points: the drop-in for the reference's C++ call sites,
// which hold plain host arrays (CPU/main.cpp:47-58,87-114).  A batch is cut
// into chunks (up to 64 MiB per input array) that flow through a 4-slot ring of
// device buffers on independent streams, so the H2D copy of chunk c+1, the kernel of chunk c and the D2H copy
// of chunk c-1 overlap (the path is PCIe-bound: 100 B cross the bus per
// homography against ~100 flops of work).  Pinned caller buffers are copied
// directly; pageable ones are staged through an internal pinned ring with a
// multi-threaded memcpy.  The layer sits strictly above the device-pointer
// C ABI: it calls sks_cuda_* like any other client.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sks_cuda.h"

extern "C" void sks_multi_shutdown_internal(void);   // csrc/multi.cu
extern "C" void sks_tuning_get_internal(int* out9);     // csrc/capi.cu: per-thread tuning knobs
extern "C" void sks_tuning_set_internal(const int* in9);

namespace {

constexpr int kRing = 4;
}  // namespace
#include <atomic>
namespace {
// per input array per chunk: small enough that filling and draining the pipeline costs
// ~1 % of a 2^25-quadruple batch, large enough for full-rate PCIe DMA
// Upper bound of one input array per chunk.  Measured on B200 / PCIe Gen5
// (tools/host_chunk_sweep.py): 1 MiB 36.7, 8 MiB 44.5, 32 MiB 49.2, 64 MiB 49.8 GB/s
// H2D -- large DMAs win, so chunks are as large as still leaves ~8 of them to overlap.
std::atomic<int64_t> g_chunk_bytes{64ll << 20};

struct Slot {
    void* d_in[3] = {nullptr, nullptr, nullptr};   // src, tar, M
    void* d_out = nullptr;
    void* p_in[3] = {nullptr, nullptr, nullptr};   // pinned staging (pageable callers)
    void* p_out = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    int64_t pending_off = -1, pending_cnt = 0;     // chunk whose D2H is in flight
};

struct HostCtx {
    int device = -1;
    std::mutex mu;                     // one batch at a time per device
    int64_t cap_in[3] = {0, 0, 0}, cap_out = 0;   // bytes per device buffer
    bool staged_in[3] = {false, false, false}, staged_out = false;
    Slot slot[kRing];
};

std::mutex g_mu;                       // guards g_ctx and context (re)allocation
std::vector<HostCtx*> g_ctx;
std::atomic<int> g_host_devices{1};    // GPUs a host-pointer batch is sharded over (1 = current only)

#define CK(x)                                  \
    do {                                       \
        cudaError_t _e = (x);                  \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

void free_ctx(HostCtx* c)
{
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    for (Slot& s : c->slot) {
        for (int k = 0; k < 3; ++k) {
            if (s.d_in[k]) cudaFree(s.d_in[k]);
            if (s.p_in[k]) cudaFreeHost(s.p_in[k]);
        }
        if (s.d_out) cudaFree(s.d_out);
        if (s.p_out) cudaFreeHost(s.p_out);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.done) cudaEventDestroy(s.done);
    }
    cudaSetDevice(prev);
    delete c;
}

// Find or create the per-device context (table lock only; no device memory is
// touched here so a batch running on the device is not disturbed).
int get_ctx(int dev, HostCtx** out)
{
    std::lock_guard<std::mutex> table(g_mu);
    HostCtx* c = nullptr;
    for (HostCtx* x : g_ctx)
        if (x->device == dev) c = x;
    if (c == nullptr) {
        c = new HostCtx();
        c->device = dev;
        for (Slot& s : c->slot) {
            cudaError_t e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
            if (e != cudaSuccess) {      // never publish a half-initialised context
                free_ctx(c);
                return (int)e;
            }
        }
        g_ctx.push_back(c);
    }
    *out = c;
    return SKS_OK;
}

// Grow the ring's device / pinned buffers.  Caller holds c->mu, so no batch is in
// flight on these buffers.
int ensure_capacity(HostCtx* c, const int64_t* need_in, int n_in, int64_t need_out, bool stage_in,
                    bool stage_out)
{
    // input array k gets exactly what it needs (the M array of ACA-rect is a quarter of the
    // corner array, and most calls have no third input at all)
    for (int k = 0; k < n_in; ++k) {
        if (need_in[k] <= c->cap_in[k]) continue;
        for (Slot& s : c->slot) {
            if (s.d_in[k]) CK(cudaFree(s.d_in[k]));
            s.d_in[k] = nullptr;
            CK(cudaMalloc(&s.d_in[k], (size_t)need_in[k]));
            if (s.p_in[k]) CK(cudaFreeHost(s.p_in[k]));
            s.p_in[k] = nullptr;
        }
        c->cap_in[k] = need_in[k];
        c->staged_in[k] = false;
    }
    if (need_out > c->cap_out) {
        for (Slot& s : c->slot) {
            if (s.d_out) CK(cudaFree(s.d_out));
            s.d_out = nullptr;
            CK(cudaMalloc(&s.d_out, (size_t)need_out));
            if (s.p_out) CK(cudaFreeHost(s.p_out));
            s.p_out = nullptr;
        }
        c->cap_out = need_out;
        c->staged_out = false;
    }
    for (int k = 0; k < n_in; ++k)
        if (stage_in && !c->staged_in[k]) {
            for (Slot& s : c->slot)
                CK(cudaHostAlloc(&s.p_in[k], (size_t)c->cap_in[k], cudaHostAllocDefault));
            c->staged_in[k] = true;
        }
    if (stage_out && !c->staged_out) {
        for (Slot& s : c->slot)
            CK(cudaHostAlloc(&s.p_out, (size_t)c->cap_out, cudaHostAllocDefault));
        c->staged_out = true;
    }
    return SKS_OK;
}

bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// Staging copy with non-temporal stores (SSE2, baseline x86-64).  The pageable path is bound by the
// host's memory system, not by PCIe: per homography 64 B are copied into the pinned ring and 36 B out
// of it, and an ordinary store to a line that is not in cache first READS the line (read for
// ownership), so a plain memcpy moves 3 bytes over the memory bus per byte copied.  Streaming stores
// skip that read and keep the (never re-read) staging data out of the caches.
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
static void stream_copy(void* dst, const void* src, size_t n)
{
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    if (head > n) head = n;
    std::memcpy(d, s, head);
    d += head; s += head; n -= head;
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; ++i) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
        s += 64; d += 64;
    }
    std::memcpy(d, s, n - blocks * 64);
    _mm_sfence();
}
#else
static void stream_copy(void* dst, const void* src, size_t n) { std::memcpy(dst, src, n); }
#endif
std::atomic<int> g_stream_copy{3};     // sks_host_set_staging_copy: bit 0 = non-temporal stores into the pinned ring,
                                       // bit 1 = non-temporal stores out of it into the caller's result buffer

// Persistent memcpy workers for the pageable staging path: spawning threads per copy costs
// about as much as copying a few MiB, and the in- and out-stagers copy at the same time.
class CopyPool {
public:
    static CopyPool& get()
    {
        static CopyPool* p = new CopyPool();   // leaked on purpose: workers may outlive static destructors
        return *p;
    }
    // copy [src, src+bytes) to dst in slices of >= 1 MiB on up to `max_parts` workers; blocks
    void copy(void* dst, const void* src, size_t bytes, unsigned max_parts, bool nt)
    {
        const unsigned parts = (unsigned)std::min<size_t>(std::min<size_t>(max_parts, workers_.size()),
                                                          bytes / (1u << 20) + 1);
        if (parts <= 1 || workers_.empty()) {
            if (nt) stream_copy(dst, src, bytes); else std::memcpy(dst, src, bytes);
            return;
        }
        Job job;
        job.left = (int)parts;
        const size_t per = ((bytes / parts) + 4095) & ~size_t(4095);
        {
            std::lock_guard<std::mutex> g(m_);
            for (unsigned t = 0; t < parts; ++t) {
                const size_t lo = std::min(bytes, per * t), hi = (t + 1 == parts) ? bytes : std::min(bytes, lo + per);
                q_.push_back(Task{(char*)dst + lo, (const char*)src + lo, hi - lo, &job, nt});
            }
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> g(job.m);
        job.cv.wait(g, [&] { return job.left == 0; });
    }

private:
    struct Job {
        std::mutex m;
        std::condition_variable cv;
        int left = 0;
    };
    struct Task {
        char* dst;
        const char* src;
        size_t bytes;
        Job* job;
        bool nt;
    };
    CopyPool()
    {
        const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
        const unsigned n = std::min(16u, hw);
        for (unsigned i = 0; i < n; ++i)
            workers_.emplace_back([this] { loop(); }).detach();
    }
    void loop()
    {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return !q_.empty(); });
                t = q_.front();
                q_.pop_front();
            }
            if (t.bytes) {
                if (t.nt) stream_copy(t.dst, t.src, t.bytes); else std::memcpy(t.dst, t.src, t.bytes);
            }
            std::lock_guard<std::mutex> g(t.job->m);
            if (--t.job->left == 0) t.job->cv.notify_one();
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Task> q_;
    std::vector<std::thread> workers_;
};

// worker threads one staging copy is spread over (sks_host_set_staging_threads); 0 = three quarters of the
// pool (12 of 16: measured best on a 16-vCPU B200 host, 0.65 vs 0.58 G H/s at 8 and 0.59 at 16 -- the in- and
// out-stager copy at the same time and share the pool)
std::atomic<int> g_copy_parts{0};
void parallel_copy(void* dst, const void* src, size_t bytes, bool nt)
{
    int parts = g_copy_parts.load();
    if (parts <= 0) {
        const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
        parts = (int)(std::min(16u, hw) * 3 / 4);
        if (parts < 1) parts = 1;
    }
    CopyPool::get().copy(dst, src, bytes, (unsigned)parts, nt);
}

}  // namespace

// H2D copy of a host array that may be pageable, for the RANSAC host entries (also csrc/multi.cu):
// pinned sources go out as one cudaMemcpyAsync; pageable ones are cut into 8 MiB pieces that a
// worker pool copies (non-temporal stores) into two alternating pinned buffers while the previous
// piece is on the bus -- cudaMemcpyAsync straight from pageable memory managed ~13 GB/s here.
// `sb` holds the two pinned buffers and their events (owned by the caller's per-device context).
struct SksStageBuf {
    void* p[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
};
extern "C" int sks_stage_h2d_internal(void* dst, const void* src, size_t bytes, cudaStream_t st, SksStageBuf* sb)
{
    constexpr size_t PIECE = 8u << 20;
    if (bytes == 0) return SKS_OK;
    if (is_pinned(src) || bytes < (1u << 20)) {
        const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
        return e == cudaSuccess ? SKS_OK : (int)e;
    }
    for (int i = 0; i < 2; ++i)
        if (sb->p[i] == nullptr) {
            CK(cudaHostAlloc(&sb->p[i], PIECE, cudaHostAllocDefault));
            CK(cudaEventCreateWithFlags(&sb->ev[i], cudaEventDisableTiming));
            sb->used[i] = false;
        }
    const bool nt = (g_stream_copy.load() & 1) != 0;
    int i = 0;
    for (size_t off = 0; off < bytes; off += PIECE, i ^= 1) {
        const size_t n = bytes - off < PIECE ? bytes - off : PIECE;
        if (sb->used[i]) CK(cudaEventSynchronize(sb->ev[i]));      // the DMA that last read this buffer
        parallel_copy(sb->p[i], static_cast<const char*>(src) + off, n, nt);
        CK(cudaMemcpyAsync(static_cast<char*>(dst) + off, sb->p[i], n, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(sb->ev[i], st));
        sb->used[i] = true;
    }
    return SKS_OK;
}
extern "C" void sks_stage_free_internal(SksStageBuf* sb)
{
    for (int i = 0; i < 2; ++i) {
        if (sb->ev[i]) { cudaEventSynchronize(sb->ev[i]); cudaEventDestroy(sb->ev[i]); }
        if (sb->p[i]) cudaFreeHost(sb->p[i]);
        sb->p[i] = nullptr; sb->ev[i] = nullptr; sb->used[i] = false;
    }
}

namespace {

// Generic pipeline.  in[k] (k < n_in) are host arrays of in_elems[k] elements
// per quadruple; out is 9 elements per quadruple; `launch` enqueues the solver
// for `cnt` quadruples on device buffers.
template <typename T, typename Launch>
int run_on_device(int dev, const T* const* in, const int* in_elems, int n_in, T* out, int64_t n,
                  Launch launch)
{
    CK(cudaSetDevice(dev));
    bool stage_in = false;
    for (int k = 0; k < n_in; ++k) stage_in = stage_in || !is_pinned(in[k]);
    const bool stage_out = !is_pinned(out);
    // non-temporal staging copies (measured on B200 hosts, tools/host_pageable_probe.py): into the ring
    // always (0.40 -> 0.55 G H/s with pageable in/out); out of it when the input is staged as well and
    // the host's memory bus is the bottleneck (0.55 -> 0.59), not when only the output is (0.72 -> 0.68)
    const bool nt_in = (g_stream_copy.load() & 1) != 0;
    const bool nt_out = (g_stream_copy.load() & 2) != 0 && stage_in;
    // staged (pageable) batches use smaller chunks: the two host memcpys overlap the DMA at a
    // finer grain and the pipeline fills sooner
    const int64_t cap_bytes = (stage_in || stage_out) ? std::min<int64_t>(g_chunk_bytes.load(), 16ll << 20)
                                                      : g_chunk_bytes.load();
    const int64_t cap = cap_bytes / (8 * (int64_t)sizeof(T));
    const int64_t floor_q = std::min<int64_t>(cap, (8ll << 20) / (8 * (int64_t)sizeof(T)));
    const int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>(n, std::min(cap, std::max(floor_q, (n + 7) / 8))));
    // Chunk schedule: full-size chunks in the middle, a ramp of chunk/8, /4, /2 at both ends.
    // Nothing overlaps the first H2D and the last D2H, so the pipeline's fill and drain should
    // move as few bytes as possible (with 64 MiB chunks they were ~4 ms of a 43 ms batch).
    std::vector<int64_t> offs, cnts;
    {
        std::vector<int64_t> head, tail;
        int64_t lo = 0, hi = n;
        if (n >= 6 * chunk)
            for (int64_t part = std::max<int64_t>(1024, chunk / 8); part < chunk && hi - lo > 2 * part; part *= 2) {
                head.push_back(part);
                tail.push_back(part);
                lo += part;
                hi -= part;
            }
        int64_t off = 0;
        for (int64_t c0 : head) { offs.push_back(off); cnts.push_back(c0); off += c0; }
        while (off < hi) {
            const int64_t c0 = std::min(chunk, hi - off);
            offs.push_back(off); cnts.push_back(c0); off += c0;
        }
        for (auto it = tail.rbegin(); it != tail.rend(); ++it) { offs.push_back(off); cnts.push_back(*it); off += *it; }
    }
    HostCtx* c = nullptr;
    if (int rc = get_ctx(dev, &c)) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    int64_t need_in[3] = {0, 0, 0};
    for (int k = 0; k < n_in; ++k) need_in[k] = chunk * in_elems[k] * (int64_t)sizeof(T);
    if (int rc = ensure_capacity(c, need_in, n_in, chunk * 9 * (int64_t)sizeof(T), stage_in, stage_out))
        return rc;

    if (stage_in || stage_out) {
        // Pageable caller buffers: three actors so that the memcpy into the pinned ring, the
        // DMA + kernel, and the memcpy out of the ring all run at the same time instead of
        // taking turns on one host thread --
        //   in-stager  : chunk ci -> slot's pinned inputs (waits until the slot was drained)
        //   this thread: H2D, kernel, D2H, event on the slot's stream (waits for the in-stager)
        //   out-stager : waits for the event, copies the slot's pinned output to the caller
        // Chunks complete in order, so three monotonic counters are the whole protocol.
        const int64_t n_chunks = (int64_t)offs.size();
        std::mutex m;
        std::condition_variable cv;
        int64_t staged = 0, enqueued = 0, freed = 0;     // chunks that passed each stage
        std::atomic<int> err{SKS_OK};
        auto fail = [&](int rc) {
            int ok = SKS_OK;
            err.compare_exchange_strong(ok, rc);
            std::lock_guard<std::mutex> g(m);
            cv.notify_all();
        };
        std::thread in_stager([&] {
            cudaSetDevice(dev);
            for (int64_t ci = 0; ci < n_chunks && err.load() == SKS_OK; ++ci) {
                {
                    std::unique_lock<std::mutex> g(m);     // slot reuse: chunk ci - kRing must be drained
                    cv.wait(g, [&] { return freed >= ci - kRing + 1 || err.load() != SKS_OK; });
                }
                if (err.load() != SKS_OK) break;
                Slot& sl = c->slot[ci % kRing];
                const int64_t off = offs[ci], cnt = cnts[ci];
                if (stage_in)
                    for (int k = 0; k < n_in; ++k)
                        parallel_copy(sl.p_in[k], in[k] + off * in_elems[k],
                                      (size_t)cnt * in_elems[k] * sizeof(T), nt_in);
                std::lock_guard<std::mutex> g(m);
                staged = ci + 1;
                cv.notify_all();
            }
        });
        std::thread out_stager([&] {
            cudaSetDevice(dev);
            for (int64_t ci = 0; ci < n_chunks && err.load() == SKS_OK; ++ci) {
                {
                    std::unique_lock<std::mutex> g(m);
                    cv.wait(g, [&] { return enqueued > ci || err.load() != SKS_OK; });
                }
                if (err.load() != SKS_OK) break;
                Slot& sl = c->slot[ci % kRing];
                const cudaError_t e = cudaEventSynchronize(sl.done);
                if (e != cudaSuccess) { fail((int)e); break; }
                const int64_t off = offs[ci], cnt = cnts[ci];
                if (stage_out)
                    parallel_copy(out + off * 9, sl.p_out, (size_t)cnt * 9 * sizeof(T), nt_out);
                std::lock_guard<std::mutex> g(m);
                freed = ci + 1;
                cv.notify_all();
            }
        });
        auto enqueue = [&](int64_t ci) -> int {
            Slot& sl = c->slot[ci % kRing];
            const int64_t off = offs[ci], cnt = cnts[ci];
            for (int k = 0; k < n_in; ++k) {
                const size_t bytes = (size_t)cnt * in_elems[k] * sizeof(T);
                const T* hsrc = stage_in ? static_cast<const T*>(sl.p_in[k]) : in[k] + off * in_elems[k];
                CK(cudaMemcpyAsync(sl.d_in[k], hsrc, bytes, cudaMemcpyHostToDevice, sl.stream));
            }
            if (int rc = launch(sl, cnt)) return rc;
            T* hdst = stage_out ? static_cast<T*>(sl.p_out) : out + off * 9;
            CK(cudaMemcpyAsync(hdst, sl.d_out, (size_t)cnt * 9 * sizeof(T), cudaMemcpyDeviceToHost, sl.stream));
            CK(cudaEventRecord(sl.done, sl.stream));
            return SKS_OK;
        };
        for (int64_t ci = 0; ci < n_chunks && err.load() == SKS_OK; ++ci) {
            {
                std::unique_lock<std::mutex> g(m);
                cv.wait(g, [&] { return staged > ci || err.load() != SKS_OK; });
            }
            if (err.load() != SKS_OK) break;
            if (int rc = enqueue(ci)) { fail(rc); break; }
            std::lock_guard<std::mutex> g(m);
            enqueued = ci + 1;
            cv.notify_all();
        }
        in_stager.join();
        out_stager.join();
        if (err.load() != SKS_OK) cudaDeviceSynchronize();   // nothing of this batch left in flight
        return err.load();
    }

    auto drain = [&](Slot& s) -> int {   // finish the chunk this slot last produced
        if (s.pending_off < 0) return SKS_OK;
        CK(cudaEventSynchronize(s.done));
        if (stage_out)
            parallel_copy(out + s.pending_off * 9, s.p_out, (size_t)s.pending_cnt * 9 * sizeof(T), nt_out);
        s.pending_off = -1;
        return SKS_OK;
    };

    const int64_t n_chunks = (int64_t)offs.size();
    int rc = SKS_OK;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == SKS_OK) rc = (int)e; return e == cudaSuccess; };
    for (int64_t ci = 0; ci < n_chunks && rc == SKS_OK; ++ci) {
        Slot& s = c->slot[ci % kRing];
        if ((rc = drain(s)) != SKS_OK) break;
        const int64_t off = offs[ci], cnt = cnts[ci];
        for (int k = 0; k < n_in && rc == SKS_OK; ++k) {
            const size_t bytes = (size_t)cnt * in_elems[k] * sizeof(T);
            const T* hsrc = in[k] + off * in_elems[k];
            if (stage_in) {
                parallel_copy(s.p_in[k], hsrc, bytes, nt_in);
                hsrc = static_cast<const T*>(s.p_in[k]);
            }
            cu(cudaMemcpyAsync(s.d_in[k], hsrc, bytes, cudaMemcpyHostToDevice, s.stream));
        }
        if (rc != SKS_OK) break;
        rc = launch(s, cnt);
        if (rc != SKS_OK) break;
        T* hdst = stage_out ? static_cast<T*>(s.p_out) : out + off * 9;
        if (!cu(cudaMemcpyAsync(hdst, s.d_out, (size_t)cnt * 9 * sizeof(T), cudaMemcpyDeviceToHost, s.stream)))
            break;
        if (!cu(cudaEventRecord(s.done, s.stream))) break;
        s.pending_off = off;
        s.pending_cnt = cnt;
    }
    if (rc != SKS_OK) {
        // an enqueue failed mid-batch: nothing of this batch may stay in flight towards the
        // caller's buffers, and the persistent context must not remember a pending chunk
        for (Slot& s : c->slot) {
            cudaStreamSynchronize(s.stream);
            s.pending_off = -1;
        }
        cudaGetLastError();
        return rc;
    }
    for (Slot& s : c->slot) {
        const int r2 = drain(s);
        if (rc == SKS_OK) rc = r2;
        if (r2 != SKS_OK) {
            cudaStreamSynchronize(s.stream);
            s.pending_off = -1;
        }
    }
    return rc;
}

// Generic pipeline.  in[k] (k < n_in) are host arrays of in_elems[k] elements
// per quadruple; out is 9 elements per quadruple; `launch` enqueues the solver
// for `cnt` quadruples on device buffers.  With sks_host_set_device_count(g > 1)
// the batch is cut into g contiguous shards, one host thread
// and one PCIe link per GPU, no inter-GPU traffic.
template <typename T, typename Launch>
int run_pipeline(const T* const* in, const int* in_elems, int n_in, T* out, int64_t n, Launch launch)
{
    if (n < 0 || (n > 0 && out == nullptr)) return SKS_ERR_INVALID_ARG;
    for (int k = 0; k < n_in; ++k)
        if (n > 0 && in[k] == nullptr) return SKS_ERR_INVALID_ARG;
    int visible = 0;
    if (int rc = sks_cuda_device_count(&visible)) return rc;
    if (visible <= 0) return SKS_ERR_NO_DEVICE;
    if (n == 0) return SKS_OK;
    int cur = 0;
    CK(cudaGetDevice(&cur));
    int want = g_host_devices.load();
    if (want <= 0 || want > visible) want = visible;
    const int g = (int)std::min<int64_t>(want, std::max<int64_t>(1, n / (1 << 16)));
    if (g <= 1) return run_on_device<T>(cur, in, in_elems, n_in, out, n, launch);

    std::vector<int> rcs(g, SKS_OK);
    std::vector<std::thread> workers;
    int knobs[9];
    sks_tuning_get_internal(knobs);          // the workers launch with the caller's tuning
    for (int d = 0; d < g; ++d) {
        int64_t begin = 0, count = 0;
        sks_cuda_shard_range(n, d, g, &begin, &count);
        workers.emplace_back([=, &rcs] {
            sks_tuning_set_internal(knobs);
            const T* sub[3] = {nullptr, nullptr, nullptr};
            for (int k = 0; k < n_in; ++k) sub[k] = in[k] + begin * in_elems[k];
            rcs[d] = run_on_device<T>(d, sub, in_elems, n_in, out + begin * 9, count, launch);
        });
    }
    for (auto& w : workers) w.join();
    cudaSetDevice(cur);
    for (int rc : rcs)
        if (rc != SKS_OK) return rc;
    return SKS_OK;
}

template <typename T, typename Fn>
int host_general(Fn fn, const T* src, const T* tar, T* H, int64_t n, int flags)
{
    const T* in[2] = {src, tar};
    const int elems[2] = {8, 8};
    return run_pipeline<T>(in, elems, 2, H, n, [&](Slot& s, int64_t cnt) {
        return fn(static_cast<const T*>(s.d_in[0]), static_cast<const T*>(s.d_in[1]),
                  static_cast<T*>(s.d_out), cnt, SKS_LAYOUT_AOS, 0, flags, nullptr, s.stream);
    });
}

template <typename T, typename Fn>
int host_rect(Fn fn, const T* tar, const T* M, T mx, T my, T width, T ratio, T* H, int64_t n, int flags)
{
    const T* in[2] = {tar, M};
    const int elems[2] = {8, 2};
    return run_pipeline<T>(in, elems, M ? 2 : 1, H, n, [&](Slot& s, int64_t cnt) {
        return fn(static_cast<const T*>(s.d_in[0]), M ? static_cast<const T*>(s.d_in[1]) : nullptr,
                  mx, my, width, ratio, static_cast<T*>(s.d_out), cnt, SKS_LAYOUT_AOS, 0, flags,
                  nullptr, s.stream);
    });
}

}  // namespace

extern "C" {

int sks_host_aca_f32(const float* src, const float* tar, float* H, int64_t n, int flags)
{
    return host_general<float>(sks_cuda_aca_f32, src, tar, H, n, flags);
}
int sks_host_aca_f64(const double* src, const double* tar, double* H, int64_t n, int flags)
{
    return host_general<double>(sks_cuda_aca_f64, src, tar, H, n, flags);
}
int sks_host_sks_f32(const float* src, const float* tar, float* H, int64_t n, int flags)
{
    return host_general<float>(sks_cuda_sks_f32, src, tar, H, n, flags);
}
int sks_host_sks_f64(const double* src, const double* tar, double* H, int64_t n, int flags)
{
    return host_general<double>(sks_cuda_sks_f64, src, tar, H, n, flags);
}
int sks_host_ge_f32(const float* src, const float* tar, float* H, int64_t n, int flags)
{
    return host_general<float>(sks_cuda_ge_f32, src, tar, H, n, flags);
}
int sks_host_ge_f64(const double* src, const double* tar, double* H, int64_t n, int flags)
{
    return host_general<double>(sks_cuda_ge_f64, src, tar, H, n, flags);
}
int sks_host_aca_rect_f32(const float* tar, const float* M, float mx, float my, float width,
                          float ratio, float* H, int64_t n, int flags)
{
    return host_rect<float>(sks_cuda_aca_rect_f32, tar, M, mx, my, width, ratio, H, n, flags);
}
int sks_host_aca_rect_f64(const double* tar, const double* M, double mx, double my, double width,
                          double ratio, double* H, int64_t n, int flags)
{
    return host_rect<double>(sks_cuda_aca_rect_f64, tar, M, mx, my, width, ratio, H, n, flags);
}

// Whole robust estimate for callers that hold their matches in host memory: H2D of the
// correspondences (and of an explicit sample list, if any), fused scoring of all hypothesis
// ids, finalize, D2H of models / counts / masks -- on the current device.  Device buffers
// and the stream are kept per device and only ever grow (an allocation per call costs more
// than the copies: 19 ms of a 122 ms call at 1024 pairs x 4096 matches).
struct RansacHostCtx {
    int device = -1;
    std::mutex mu;
    cudaStream_t stream = nullptr;
    void* buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // corr, H, cnt, key, samples, mask
    size_t cap[6] = {0, 0, 0, 0, 0, 0};
    SksStageBuf stage;
};
static std::vector<RansacHostCtx*> g_ransac_ctx;      // guarded by g_mu

int sks_host_ransac_aca_f32(const float* corr, int64_t n_pairs, int32_t n_pts, const uint32_t* samples,
                            uint32_t n_hyp, uint64_t seed, float thr2, float* H_best,
                            uint32_t* inlier_count, uint8_t* inlier_mask, unsigned long long* best_key)
{
    if (corr == nullptr || H_best == nullptr || n_pairs < 0 || n_pts <= 0 || n_hyp == 0)
        return SKS_ERR_INVALID_ARG;          // no hypotheses: there would be no model to return
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
        cudaGetLastError();
        return SKS_ERR_NO_DEVICE;
    }
    if (n_pairs == 0) return SKS_OK;
    if (g_host_devices.load() != 1 && dev_count > 1)     // in-library multi-GPU driver (csrc/multi.cu)
        return sks_host_ransac_aca_multi_f32(corr, n_pairs, n_pts, samples, n_hyp, seed, thr2, g_host_devices.load(),
                                             H_best, inlier_count, inlier_mask, best_key);
    int dev = 0;
    CK(cudaGetDevice(&dev));
    RansacHostCtx* c = nullptr;
    {
        std::lock_guard<std::mutex> table(g_mu);
        for (RansacHostCtx* x : g_ransac_ctx)
            if (x->device == dev) c = x;
        if (c == nullptr) {
            c = new RansacHostCtx();
            c->device = dev;
            CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
            g_ransac_ctx.push_back(c);
        }
    }
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t need[6] = {(size_t)n_pairs * n_pts * 4 * sizeof(float), (size_t)n_pairs * 9 * sizeof(float),
                            (size_t)n_pairs * sizeof(uint32_t), (size_t)n_pairs * sizeof(unsigned long long),
                            samples ? (size_t)n_pairs * n_hyp * 4 * sizeof(uint32_t) : 0,
                            inlier_mask ? (size_t)n_pairs * n_pts : 0};
    for (int k = 0; k < 6; ++k)
        if (need[k] > c->cap[k]) {
            if (c->buf[k]) CK(cudaFree(c->buf[k]));
            c->buf[k] = nullptr;
            c->cap[k] = 0;
            CK(cudaMalloc(&c->buf[k], need[k]));
            c->cap[k] = need[k];
        }
    cudaStream_t st = c->stream;
    float* d_corr = static_cast<float*>(c->buf[0]);
    float* d_H = static_cast<float*>(c->buf[1]);
    uint32_t* d_cnt = static_cast<uint32_t*>(c->buf[2]);
    unsigned long long* d_key = static_cast<unsigned long long*>(c->buf[3]);
    uint32_t* d_samp = samples ? static_cast<uint32_t*>(c->buf[4]) : nullptr;
    uint8_t* d_mask = inlier_mask ? static_cast<uint8_t*>(c->buf[5]) : nullptr;
    int rc = SKS_OK;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == SKS_OK) rc = (int)e; return rc == SKS_OK; };
    rc = sks_stage_h2d_internal(d_corr, corr, need[0], st, &c->stage);
    if (rc == SKS_OK && samples) rc = sks_stage_h2d_internal(d_samp, samples, need[4], st, &c->stage);
    if (rc == SKS_OK) cu(cudaMemsetAsync(d_key, 0, need[3], st));
    if (rc == SKS_OK)
        rc = sks_cuda_ransac_aca_f32(d_corr, n_pairs, n_pts, d_samp, n_hyp, 0, n_hyp, seed, thr2, d_key, st);
    if (rc == SKS_OK)
        rc = sks_cuda_ransac_finalize_f32(d_corr, n_pairs, n_pts, d_samp, n_hyp, seed, thr2, d_key, d_H, d_cnt,
                                          d_mask, st);
    if (rc == SKS_OK) cu(cudaMemcpyAsync(H_best, d_H, need[1], cudaMemcpyDeviceToHost, st));
    if (rc == SKS_OK && inlier_count) cu(cudaMemcpyAsync(inlier_count, d_cnt, need[2], cudaMemcpyDeviceToHost, st));
    if (rc == SKS_OK && inlier_mask) cu(cudaMemcpyAsync(inlier_mask, d_mask, need[5], cudaMemcpyDeviceToHost, st));
    if (rc == SKS_OK && best_key) cu(cudaMemcpyAsync(best_key, d_key, need[3], cudaMemcpyDeviceToHost, st));
    // success or not: no copy towards the caller's buffers may still be in flight on return
    const cudaError_t es = cudaStreamSynchronize(st);
    if (rc == SKS_OK && es != cudaSuccess) rc = (int)es;
    if (rc != SKS_OK) cudaGetLastError();
    return rc;
}

int sks_host_alloc_pinned(void** ptr, int64_t bytes)
{
    if (ptr == nullptr || bytes < 0) return SKS_ERR_INVALID_ARG;
    cudaError_t e = cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? SKS_ERR_NO_DEVICE : (int)e;
    }
    return SKS_OK;
}
int sks_host_free_pinned(void* ptr)
{
    cudaError_t e = cudaFreeHost(ptr);
    return e == cudaSuccess ? SKS_OK : (int)e;
}

// Pin a buffer the caller already owns (a std::vector's storage, a numpy array): afterwards the
// host-pointer entry points DMA from / into it directly instead of staging through the pinned ring.
int sks_host_register(void* ptr, int64_t bytes)
{
    if (ptr == nullptr || bytes <= 0) return SKS_ERR_INVALID_ARG;
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return SKS_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? SKS_ERR_NO_DEVICE : (int)e;
    }
    return SKS_OK;
}
int sks_host_unregister(void* ptr)
{
    if (ptr == nullptr) return SKS_ERR_INVALID_ARG;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) cudaGetLastError();
    return e == cudaSuccess ? SKS_OK : (int)e;
}

int sks_host_set_chunk_bytes(int64_t bytes_per_input_array)
{
    if (bytes_per_input_array < (64 << 10) || bytes_per_input_array > (1ll << 30)) return SKS_ERR_INVALID_ARG;
    g_chunk_bytes.store(bytes_per_input_array);
    return SKS_OK;
}

int sks_host_set_staging_copy(int non_temporal)
{
    if (non_temporal < 0 || non_temporal > 3) return SKS_ERR_INVALID_ARG;
    g_stream_copy.store(non_temporal);
    return SKS_OK;
}

int sks_host_set_staging_threads(int threads_per_copy)
{
    if (threads_per_copy < 0 || threads_per_copy > 64) return SKS_ERR_INVALID_ARG;   // 0 = automatic
    g_copy_parts.store(threads_per_copy);
    return SKS_OK;
}

int sks_host_set_device_count(int count)
{
    if (count < 0) return SKS_ERR_INVALID_ARG;
    g_host_devices.store(count);       // 0 = every visible GPU
    return SKS_OK;
}

int sks_cuda_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (HostCtx* c : g_ctx) {
        { std::lock_guard<std::mutex> busy(c->mu); }   // wait for a batch that is still running on it
        free_ctx(c);
    }
    g_ctx.clear();
    int prev = 0;
    cudaGetDevice(&prev);
    for (RansacHostCtx* c : g_ransac_ctx) {
        { std::lock_guard<std::mutex> busy(c->mu); }
        cudaSetDevice(c->device);
        sks_stage_free_internal(&c->stage);
        for (void* b : c->buf)
            if (b) cudaFree(b);
        if (c->stream) cudaStreamDestroy(c->stream);
        delete c;
    }
    g_ransac_ctx.clear();
    cudaSetDevice(prev);
    sks_multi_shutdown_internal();
    return SKS_OK;
}

}  // extern "C"
