// dhfk_bwd.cu -- instantiates the fused backward kernels for one (trig policy, bone-grad, input mode)
// triple (-DDHFK_TRIG=0|1 -DDHFK_GBONE=0|1 -DDHFK_GEN=0|1).
// Backward is issue-bound: the table sincos (15 instead of 23 instructions) is worth its 31 L1-hit loads
// per pose (0.151 -> 0.143 ms -- profiles/r1_ab_staging.md).
#ifndef DHFK_ACCURATE_TABLE
#define DHFK_ACCURATE_TABLE 1
#endif
#include "dhfk_launch.h"
#if !defined(DHFK_TRIG) || !defined(DHFK_GBONE) || !defined(DHFK_GEN)
#error "compile with -DDHFK_TRIG=0|1 -DDHFK_GBONE=0|1 -DDHFK_GEN=0|1|2   (2 = raw mode, wide rows)"
#endif
#define DHFK_CAT_(a, b, c, d, e, f) a##b##c##d##e##f
#define DHFK_CAT(a, b, c, d, e, f) DHFK_CAT_(a, b, c, d, e, f)
namespace dhfk {
int DHFK_CAT(launch_bwd_t, DHFK_TRIG, _b, DHFK_GBONE, _g, DHFK_GEN)(const BwdParams& p, bool guv, cudaStream_t st,
                                                                    const char** where) {
    constexpr bool G = DHFK_GEN == 1, W = DHFK_GEN == 2, B = DHFK_GBONE != 0;
    const bool gw = p.g_world != nullptr, gc = p.g_cam != nullptr;
    const size_t smem = bwd_smem_bytes(gw, gc, guv, G, p.w);
    // which upstream gradients exist is a compile-time property of the kernel (7 combinations)
#define DHFK_BWD(GW_, C, U) return launch_tiles(dhfk_bwd_kernel<GW_, C, U, B, DHFK_TRIG, G, W>, smem, p, st, where)
    if (gw && !gc && guv) DHFK_BWD(true, false, true);       // the GAN step: world + 2D critics
    if (gw && !gc && !guv) DHFK_BWD(true, false, false);     // FK only
    if (gw && gc && guv) DHFK_BWD(true, true, true);
    if (gw && gc && !guv) DHFK_BWD(true, true, false);
    if (!gw && gc && guv) DHFK_BWD(false, true, true);
    if (!gw && gc && !guv) DHFK_BWD(false, true, false);
    if (!gw && !gc && guv) DHFK_BWD(false, false, true);
#undef DHFK_BWD
    *where = "no upstream gradient";
    return (int)cudaErrorInvalidValue;
}
}  // namespace dhfk
