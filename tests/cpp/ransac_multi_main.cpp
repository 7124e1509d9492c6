// A single-process C++ caller (the shape of the reference's CPU/main.cpp: host arrays, no
// Python, no torch.distributed) that runs the fused ACA-RANSAC on one GPU and on all GPUs of
// the box through the drop-in header, and checks that the multi-GPU result is bit-identical.
//   ransac_multi [pairs] [points] [hypotheses]
// Prints "gpus G  single_ms .. multi_ms ..  identical 0|1"; exit code 0 iff identical.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sks_homography.hpp"

int main(int argc, char** argv)
{
    const std::int64_t P = argc > 1 ? std::atoll(argv[1]) : 64;
    const std::int32_t n_pts = argc > 2 ? std::atoi(argv[2]) : 1024;
    const std::uint32_t n_hyp = argc > 3 ? (std::uint32_t)std::atoll(argv[3]) : 4096;
    int gpus = 0;
    if (int rc = sks_cuda_device_count(&gpus)) { std::printf("status %d\n", rc); return 3; }
    // a planted homography + noise + outliers, generated on the host like a caller's matches
    std::vector<float> corr((size_t)P * n_pts * 4);
    std::uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((s >> 40) * (1.0 / 16777216.0)); };
    for (std::int64_t p = 0; p < P; ++p) {
        const float a = 1.0f + 0.1f * rnd(), b = 0.05f * rnd(), tx = 5 * rnd(), ty = 5 * rnd();
        const float g = 1e-4f * rnd(), h = 1e-4f * rnd();
        for (std::int32_t i = 0; i < n_pts; ++i) {
            float* c = &corr[((size_t)p * n_pts + i) * 4];
            c[0] = 160 * rnd(); c[1] = 160 * rnd();
            if (i & 1) { c[2] = 160 * rnd(); c[3] = 160 * rnd(); continue; }   // outlier
            const float w = g * c[0] + h * c[1] + 1;
            c[2] = (a * c[0] - b * c[1] + tx) / w + (rnd() - 0.5f);
            c[3] = (b * c[0] + a * c[1] + ty) / w + (rnd() - 0.5f);
        }
    }
    std::vector<float> H1((size_t)P * 9), Hn((size_t)P * 9);
    std::vector<std::uint32_t> c1(P), cn(P);
    std::vector<std::uint8_t> m1((size_t)P * n_pts), mn((size_t)P * n_pts);
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    int rc = sks::runRansac_ACA(corr.data(), P, n_pts, n_hyp, 11, 2.25f, H1.data(), c1.data(), m1.data());   // warm-up
    if (rc) { std::printf("status %d\n", rc); return 4; }
    auto t0 = now();
    rc = sks::runRansac_ACA(corr.data(), P, n_pts, n_hyp, 11, 2.25f, H1.data(), c1.data(), m1.data());
    auto t1 = now();
    if (rc) { std::printf("status %d\n", rc); return 4; }
    rc = sks::runRansac_ACA_multi(corr.data(), P, n_pts, n_hyp, 11, 2.25f, 0, Hn.data(), cn.data(), mn.data());   // warm-up
    if (rc) { std::printf("multi status %d (%s)\n", rc, sks_cuda_error_string(rc)); return 5; }
    auto t2 = now();
    rc = sks::runRansac_ACA_multi(corr.data(), P, n_pts, n_hyp, 11, 2.25f, 0, Hn.data(), cn.data(), mn.data());
    auto t3 = now();
    if (rc) { std::printf("multi status %d (%s)\n", rc, sks_cuda_error_string(rc)); return 5; }
    // and through the unchanged single-GPU entry point with the library-wide device count
    std::vector<float> Hs((size_t)P * 9);
    sks_host_set_device_count(0);
    rc = sks::runRansac_ACA(corr.data(), P, n_pts, n_hyp, 11, 2.25f, Hs.data());
    sks_host_set_device_count(1);
    if (rc) { std::printf("set_device_count status %d\n", rc); return 6; }
    const bool same = std::memcmp(H1.data(), Hn.data(), H1.size() * 4) == 0 &&
                      std::memcmp(c1.data(), cn.data(), c1.size() * 4) == 0 &&
                      std::memcmp(m1.data(), mn.data(), m1.size()) == 0 &&
                      std::memcmp(H1.data(), Hs.data(), H1.size() * 4) == 0;
    double mean = 0;
    for (auto c : c1) mean += c;
    std::printf("gpus %d  pairs %lld  single_ms %.3f  multi_ms %.3f  mean_inliers %.1f  identical %d\n", gpus,
                (long long)P, ms(t0, t1), ms(t2, t3), mean / (double)P, same ? 1 : 0);
    sks_cuda_shutdown();
    return same ? 0 : 1;
}
