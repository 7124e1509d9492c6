// A reference-style caller (cf. CPU/main.cpp:47-58,87-114) compiled against the
// drop-in header.  Prints the status and, on success, the veri_4Pts.m known
// answer computed by all four solvers.  Exit code = status of the first call.
#include <cstdio>
#include <vector>

#include "sks_homography.hpp"

int main()
{
    // ML/veri_4Pts.m:9-12 source quad and its projection
    double src_d[8] = {0, 0, 200, 0, 50, 139, 181, 93};
    double tar_d[8] = {482, 378.571428571429, 639.240222867399, 453.531347049346,
                       601.89683773457, 610.680390715948, 673.458433551996, 563.547292039412};
    float src_f[8], tar_f[8];
    for (int i = 0; i < 8; ++i) { src_f[i] = (float)src_d[i]; tar_f[i] = (float)tar_d[i]; }
    float Hf[9];
    double Hd[9];
    int rc = sks::runKernel_ACA(src_f, tar_f, Hf);          // reference signature
    std::printf("status %d\n", rc);
    if (rc != 0) return rc < 0 ? -rc : rc;
    std::printf("ACA   "); for (float v : Hf) std::printf("%.9g ", v); std::printf("\n");
    sks::runKernel_SKS(src_f, tar_f, Hf);
    std::printf("SKS   "); for (float v : Hf) std::printf("%.9g ", v); std::printf("\n");
    sks::runKernel_ACA_double(src_d, tar_d, Hd);
    std::printf("ACA64 "); for (double v : Hd) std::printf("%.17g ", v); std::printf("\n");
    sks::runKernel_SKS_double(src_d, tar_d, Hd);
    std::printf("SKS64 "); for (double v : Hd) std::printf("%.17g ", v); std::printf("\n");
    // batched form: the caller's loop collapsed into one call
    const int n = 1000;
    std::vector<float> s(8 * n), t(8 * n), H(9 * n);
    for (int k = 0; k < n; ++k)
        for (int i = 0; i < 8; ++i) { s[8 * k + i] = src_f[i] + k * 0.01f; t[8 * k + i] = tar_f[i]; }
    rc = sks::runKernel_ACA(s.data(), t.data(), H.data(), n);
    std::printf("batch status %d  H[999][0]=%.9g\n", rc, H[9 * 999]);
    return rc;
}
