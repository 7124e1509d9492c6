// TEST INFRASTRUCTURE ONLY -- batch / thread wrapper around the reference's own
// C++ solvers, linked with "C++ Codes/modules/ACA_SKS.cpp" compiled IN PLACE
// from /root/reference by oracle/Makefile into oracle/_ref/libsks_ref.so.
// No reference source is copied into this repository; this file only declares
// the four entry points (MOD/ACA_SKS.hpp:17-20) plus the competitor cv::runKernel_GE
// (MOD/GE.hpp:9, from MOD/GE.cpp compiled the same way) and loops over them the way
// the reference's CPU harness does (CPU/main.cpp:87-114), but over DISTINCT
// quadruples streamed from memory and optionally on several host threads.
//
// Used by: tests (to pin oracle/sks_oracle.c), tests/golden/make_golden.py,
// bench.py cpu_baseline (kind "reference") and bench.py --impl reference.
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace sks {
int runKernel_ACA(float* src, float* tar, float* result);
int runKernel_ACA_double(double* src, double* tar, double* result);
int runKernel_SKS(float* src, float* tar, float* result);
int runKernel_SKS_double(double* src, double* tar, double* result);
}  // namespace sks
namespace cv {
void runKernel_GE(float* src, float* tar, float* result);   // MOD/GE.hpp:9 (RHO-GE competitor, fp32 only)
}

namespace {

int ge_as_int(float* s, float* t, float* r)
{
    cv::runKernel_GE(s, t, r);
    return 0;
}
}  // namespace

namespace {

template <typename T, int (*Solve)(T*, T*, T*)>
void run_batch(const T* src, const T* tar, T* H, int64_t n, int threads)
{
    auto body = [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i)
            Solve(const_cast<T*>(src) + 8 * i, const_cast<T*>(tar) + 8 * i, H + 9 * i);
    };
    if (threads <= 1 || n < 2 * (int64_t)threads) {
        body(0, n);
        return;
    }
    std::vector<std::thread> pool;
    const int64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const int64_t lo = std::min<int64_t>(n, per * t), hi = std::min<int64_t>(n, lo + per);
        if (lo < hi)
            pool.emplace_back(body, lo, hi);
    }
    for (auto& th : pool)
        th.join();
}

}  // namespace

extern "C" {

int ref_hardware_threads() { return (int)std::max(1u, std::thread::hardware_concurrency()); }

void ref_aca_f32(const float* s, const float* t, float* H, int64_t n, int threads)
{
    run_batch<float, sks::runKernel_ACA>(s, t, H, n, threads);
}
void ref_aca_f64(const double* s, const double* t, double* H, int64_t n, int threads)
{
    run_batch<double, sks::runKernel_ACA_double>(s, t, H, n, threads);
}
void ref_sks_f32(const float* s, const float* t, float* H, int64_t n, int threads)
{
    run_batch<float, sks::runKernel_SKS>(s, t, H, n, threads);
}
void ref_sks_f64(const double* s, const double* t, double* H, int64_t n, int threads)
{
    run_batch<double, sks::runKernel_SKS_double>(s, t, H, n, threads);
}
void ref_ge_f32(const float* s, const float* t, float* H, int64_t n, int threads)
{
    run_batch<float, ge_as_int>(s, t, H, n, threads);
}

}  // extern "C"
