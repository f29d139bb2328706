// dhfk_fwd.cu -- instantiates the fused forward kernels for one trig policy (-DDHFK_TRIG=0|1).
// Forward is HBM-bound (96 % of the copy roofline): the polynomial sincos keeps the LSU free for the
// tile traffic.  (The table variant measured 0.099 ms vs 0.090 ms here -- profiles/r1_ab_staging.md.)
#ifndef DHFK_ACCURATE_TABLE
#define DHFK_ACCURATE_TABLE 0
#endif
#include "dhfk_launch.h"
#ifndef DHFK_TRIG
#error "compile with -DDHFK_TRIG=0 (polynomial) or 1 (MUFU)"
#endif
namespace dhfk {
#if DHFK_TRIG == 0
int launch_fwd_trig0
#else
int launch_fwd_trig1
#endif
(const FwdParams& p, bool cam, bool uv, cudaStream_t st, const char** where) {
    const size_t smem = fwd_smem_bytes(cam, uv);
    if (cam && uv) return launch_tiles(dhfk_fwd_kernel<true, true, DHFK_TRIG>, smem, p, st, where);
    if (uv) return launch_tiles(dhfk_fwd_kernel<false, true, DHFK_TRIG>, smem, p, st, where);
    if (cam) return launch_tiles(dhfk_fwd_kernel<true, false, DHFK_TRIG>, smem, p, st, where);
    return launch_tiles(dhfk_fwd_kernel<false, false, DHFK_TRIG>, smem, p, st, where);
}
}  // namespace dhfk
