// dhfk_fwd.cu -- instantiates the fused forward kernels for one (trig policy, input mode) pair
// (-DDHFK_TRIG=0|1 -DDHFK_GEN=0|1).
// Forward is HBM-bound (96 % of the copy roofline): the polynomial sincos keeps the LSU free for the
// tile traffic.  (The table variant measured 0.099 ms vs 0.090 ms here -- profiles/r1_ab_staging.md.)
#ifndef DHFK_ACCURATE_TABLE
#define DHFK_ACCURATE_TABLE 0
#endif
#include "dhfk_launch.h"
#if !defined(DHFK_TRIG) || !defined(DHFK_GEN)
#error "compile with -DDHFK_TRIG=0|1 -DDHFK_GEN=0|1|2   (2 = raw mode, wide rows)"
#endif
#define DHFK_CAT_(a, b, c, d) a##b##c##d
#define DHFK_CAT(a, b, c, d) DHFK_CAT_(a, b, c, d)
namespace dhfk {
int DHFK_CAT(launch_fwd_t, DHFK_TRIG, _g, DHFK_GEN)(const FwdParams& p, bool cam, bool uv, cudaStream_t st,
                                                    const char** where) {
    constexpr bool G = DHFK_GEN == 1;
    constexpr bool W = DHFK_GEN == 2;
    const size_t smem = fwd_smem_bytes(cam, uv, G, p.w);
    if (cam && uv) return launch_tiles(dhfk_fwd_kernel<true, true, DHFK_TRIG, G, W>, smem, p, st, where);
    if (uv) return launch_tiles(dhfk_fwd_kernel<false, true, DHFK_TRIG, G, W>, smem, p, st, where);
    if (cam) return launch_tiles(dhfk_fwd_kernel<true, false, DHFK_TRIG, G, W>, smem, p, st, where);
    return launch_tiles(dhfk_fwd_kernel<false, false, DHFK_TRIG, G, W>, smem, p, st, where);
}
}  // namespace dhfk
