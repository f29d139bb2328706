// dhfk_bwd.cu -- instantiates the fused backward kernels for one (trig policy, bone-grad, input mode)
// triple (-DDHFK_TRIG=0|1 -DDHFK_GBONE=0|1 -DDHFK_GEN=0|1).
// Backward is issue-bound: the table sincos (15 instead of 23 instructions) is worth its 31 L1-hit loads
// per pose (0.151 -> 0.143 ms -- profiles/r1_ab_staging.md).
#ifndef DHFK_ACCURATE_TABLE
#define DHFK_ACCURATE_TABLE 1
#endif
#include "dhfk_launch.h"
#if !defined(DHFK_TRIG) || !defined(DHFK_GBONE) || !defined(DHFK_GEN)
#error "compile with -DDHFK_TRIG=0|1 -DDHFK_GBONE=0|1 -DDHFK_GEN=0|1"
#endif
#define DHFK_CAT_(a, b, c, d, e, f) a##b##c##d##e##f
#define DHFK_CAT(a, b, c, d, e, f) DHFK_CAT_(a, b, c, d, e, f)
namespace dhfk {
int DHFK_CAT(launch_bwd_t, DHFK_TRIG, _b, DHFK_GBONE, _g, DHFK_GEN)(const BwdParams& p, bool guv, cudaStream_t st,
                                                                    const char** where) {
    constexpr bool G = DHFK_GEN != 0, B = DHFK_GBONE != 0;
    const size_t smem = bwd_smem_bytes(p.g_world != nullptr, p.g_cam != nullptr, guv, G);
    if (guv) return launch_tiles(dhfk_bwd_kernel<true, B, DHFK_TRIG, G>, smem, p, st, where);
    return launch_tiles(dhfk_bwd_kernel<false, B, DHFK_TRIG, G>, smem, p, st, where);
}
}  // namespace dhfk
