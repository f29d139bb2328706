// dhfk_kernels.cuh -- fused DH-FK + global rotation + world->camera + pinhole projection,
// forward and analytic backward, for sm_100a.  One thread per pose, one warp (32 poses) per CTA.
//
// Data movement (see DESIGN.md "HBM layout"):
//   inputs  [N,33] [N,3] [N,15] [N,3] are AoS rows with odd lengths.  A tile of 32 rows of each
//           is one contiguous, 16-byte aligned slab; lane 0 fetches the four slabs with TMA bulk
//           copies (exact image in shared memory, odd row stride => conflict-free per-thread
//           scalar reads).
//   outputs [N,16,3] / [N,16,2] (and upstream gradients in the backward) have 48 / 32 float
//           rows; they are staged in shared memory with the row stride padded to 13 / 9
//           16-byte chunks so that both the per-thread 128-bit accesses and the cooperative
//           coalesced 128-bit global accesses (LDGSTS in, STG out) are bank-conflict free.
//   backward results d(ang), d(grot), d(root), d(bone) overwrite the input slabs in place and
//           leave through the same coalesced path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dhfk_device.cuh"

namespace dhfk {

constexpr int kTile = 32;        // poses per CTA: exactly one warp, so every barrier is a __syncwarp
constexpr int kWorldChunks = 12; // 48 floats
constexpr int kUvChunks = 8;     // 32 floats
constexpr int kWorldRow4 = kWorldChunks + 1;  // padded row stride in float4 (odd => conflict-free 128-bit access)
constexpr int kUvRow4 = kUvChunks + 1;

static_assert(kTile == 32, "the kernels assume one warp per tile");
static_assert(nth_child(-1, 0) == 0 && nth_child(-1, 1) == 5 && nth_child(-1, 2) == 10 &&
              nth_child(-1, 3) == -1, "chain roots are joints 0, 5, 10");

// Tunables for A/B measurement (profiles/r1_ab_staging.md).
// L2 prefetch distance in tiles (148 SMs x 12 resident warps = one resident wave); 0 = off.
#ifndef DHFK_PREFETCH_TILES_FWD
#define DHFK_PREFETCH_TILES_FWD 1776
#endif
// backward: a third of a wave ahead measured +3 % (0.1353 -> 0.1313 ms); one wave ahead (1776) measured nothing
#ifndef DHFK_PREFETCH_TILES_BWD
#define DHFK_PREFETCH_TILES_BWD 592
#endif

struct RowSrc {
    const float* p;
    long long stride;  // floats
    int vec;           // 1: packed + 16-byte aligned => slab path
};
struct RowDst {
    float* p;
    long long stride;
    int vec;
};

// "Wide" rows (raw mode only): the generator hands the angles and the global rotation over as column slices of ONE
// [N,S] tensor (S = 37: angles in columns 0..32, global rotation in 34..36, Fk_generator.py:136-184).  When that tensor
// is 16-byte aligned a tile of it is one contiguous slab like any other input: `wide` = S makes the kernels stage the
// [32,S] slab as it is (row stride S in shared memory too, odd => conflict-free) and read the global rotation at
// column `goff` of the same rows; 0 = separate packed arrays.  `grot_slab` = 0 drops the separate [32,3] slab from
// the shared-memory layout (possible when no ragged last tile needs it), which keeps 12 CTAs per SM.
struct WideRows {
    int wide, goff, grot_slab;
};
// WIDE is a template parameter of the kernels: carrying the run-time row stride in the packed-row instantiations cost
// the headline backward 0.8 % (A/B on one box, profiles/r2_summary.md), so the packed kernels are compiled without it.
struct FwdParams {
    RowSrc ang, grot, bone, root;   // GEN mode: `ang` is the raw network output [N,35]; grot/root unused
    WideRows w;
    float* out_world;
    float* out_cam;
    float* out_uv;
    long long n;
    CamConst cam;
    GenScale gs;                    // GEN mode only
};
struct BwdParams {
    RowSrc ang, grot, bone, root;
    WideRows w;
    int g_wide;                     // the gradient of the wide tensor is wanted as ONE [N,S] tensor (same S, goff)
    const float* g_world;
    const float* g_cam;
    const float* g_uv;
    RowDst g_ang, g_grot, g_root, g_bone;   // GEN mode: g_ang is d(raw network output) [N,35]
    long long n;
    CamConst cam;
    GenScale gs;                            // GEN mode only
};

// ---- asynchronous staging: LDGSTS in, TMA bulk (UBLKCP) out, TMA L2 prefetch ---------------------
// A full tile's inputs are fetched with ONE round trip: every lane issues its share of coalesced
// 16-byte LDGSTS (cp.async) for the four input slabs and the padded gradient rows, then blocks on
// cp.async.wait_all -- a scoreboard wait that costs no issue slots.  Measured alternatives
// (profiles/r1_ab_staging.md): TMA bulk loads completing on an mbarrier need a try_wait loop, which
// with one waiter per warp doubled the executed instructions; one cp.async.bulk per row was 25 %
// slower (the TMA unit serialises small requests).  TMA is kept where no wait loop is needed:
// bulk stores of the result slabs and L2 prefetch of the next wave's inputs.
DHFK_DI uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DHFK_DI void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DHFK_DI void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Wait without burning issue slots: try_wait with a suspend-time hint lets the hardware park the
// thread; if it still returns early, back off with nanosleep instead of polling (a bare try_wait
// loop was measured to DOUBLE the kernel's executed instructions -- profiles/r1_prof_v4).
DHFK_DI void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(4000u)
            : "memory");
        if (done) break;
        __nanosleep(128);
    }
}
DHFK_DI void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
DHFK_DI void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
                 "r"(bytes) : "memory");
}
DHFK_DI void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
DHFK_DI void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DHFK_DI void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make this thread's generic-proxy shared-memory writes visible to the async proxy (TMA)
DHFK_DI void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DHFK_DI void ldgsts16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
DHFK_DI void ldgsts4(void* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
DHFK_DI void ldgsts_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- shared <-> global staging (all warp-wide; lane = threadIdx.x) -------------------------------
// Full-tile padded rows: the warp walks the tile's 32*CH 16-byte chunks in global order (512 B per
// instruction, fully coalesced); chunk i lives in shared at row i/CH, column i%CH of a (CH+1)-chunk
// padded row.  (row, col) advance incrementally: no division in the loop.
template <int CH, class F>
DHFK_DI void for_each_tile_chunk(F&& f) {
    // chunk gi = 32 m + lane of the tile lives at shared chunk gi + gi / CH (row gi / CH of a (CH+1)-chunk row).
    // gi / CH in closed form, so every address is one per-lane base plus a compile-time immediate (the
    // incremental (row, col) update it replaces cost ~5 integer instructions per chunk):
    //   CH = 8 :  gi / 8  = 4 m + lane / 8
    //   CH = 12:  32 m = 12 (2 m + 2 (m / 3)) + {0, 8, 16}[m % 3]  =>  gi / 12 = 2 m + 2 (m / 3) + {l/12, (l+8)/12, 1 + (l+4)/12}
    const int lane = threadIdx.x;
    static_assert(CH == 8 || CH == 12 || CH == 24, "closed forms are written for 32-, 48- and 96-float rows");
    if constexpr (CH == 8) {
        const int b = lane + (lane >> 3);
#pragma unroll
        for (int m = 0; m < CH; ++m) f(m * 32 + lane, m * 36 + b);
    } else if constexpr (CH == 12) {
        const int b0 = lane + lane / 12, b1 = lane + (lane + 8) / 12, b2 = lane + (lane + 4) / 12;
#pragma unroll
        for (int m = 0; m < CH; ++m) {
            const int k = m / 3, r = m % 3;
            f(m * 32 + lane, 34 * m + 2 * k + (r == 2 ? 1 : 0) + (r == 0 ? b0 : (r == 1 ? b1 : b2)));
        }
    } else {   // CH = 24: 32 m = 24 (m + m / 3) + {0, 8, 16}[m % 3]  =>  gi / 24 = m + m / 3 + (8 (m % 3) + lane) / 24
        const int b0 = lane + lane / 24, b1 = lane + (lane + 8) / 24, b2 = lane + (lane + 16) / 24;
#pragma unroll
        for (int m = 0; m < CH; ++m) {
            const int k = m / 3, r = m % 3;
            f(m * 32 + lane, 33 * m + k + (r == 0 ? b0 : (r == 1 ? b1 : b2)));
        }
    }
}
template <int CH>
DHFK_DI void ldgsts_padded_tile(float4* s4, const float* gbase, long long row0) {
    const float4* g4 = reinterpret_cast<const float4*>(gbase) + row0 * CH;
    for_each_tile_chunk<CH>([&](int gi, int si) { ldgsts16(s4 + si, g4 + gi); });
}
// exact-image slab of kTile rows x NCOLS floats (contiguous, 16-byte aligned): coalesced LDGSTS
template <int NCOLS>
DHFK_DI void ldgsts_slab(float* s, const float* g) {
    constexpr int NV = kTile * NCOLS / 4;
    static_assert(kTile * NCOLS % 4 == 0, "slab must be a whole number of 16-byte chunks");
    const int lane = threadIdx.x;
#pragma unroll
    for (int k = 0; k < (NV + 31) / 32; ++k) {
        const int i = lane + 32 * k;
        if (NV % 32 == 0 || i < NV)
            ldgsts16(reinterpret_cast<float4*>(s) + i, reinterpret_cast<const float4*>(g) + i);
    }
}
// same for a slab whose size is only known at run time (wide rows)
DHFK_DI void ldgsts_slab_rt(float* s, const float* g, int nv /* 16-byte chunks */) {
    for (int i = threadIdx.x; i < nv; i += kTile)
        ldgsts16(reinterpret_cast<float4*>(s) + i, reinterpret_cast<const float4*>(g) + i);
}
// block until every cp.async (LDGSTS) this thread issued has landed: a scoreboard wait, no polling
DHFK_DI void ldgsts_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int CH>
DHFK_DI void store_padded_tile(const float4* s4, float* gbase, long long row0) {
    float4* g4 = reinterpret_cast<float4*>(gbase) + row0 * CH;
    for_each_tile_chunk<CH>([&](int gi, int si) { __stcs(g4 + gi, s4[si]); });
}

// Generic (ragged last tile, strided / unaligned views) paths: plain loads and stores.
// Asynchronous like the slab path: every element is requested with cp.async (16-byte chunks when the rows are packed,
// 4-byte otherwise) and the CALLER waits once with ldgsts_wait_all() after all arrays of the tile have been queued --
// one DRAM round trip per tile also for ragged tiles and strided views (synchronous LDG -> STS staging serialised one
// round trip per array: the [N,37]-view forward measured 40 % of the copy roofline with it).
template <int NCOLS>
DHFK_DI void stage_rows_in(float* s, const RowSrc& src, long long row0, int rows) {
    const int lane = threadIdx.x;
    const float* g = src.p + row0 * src.stride;
    const int nfl = rows * NCOLS;
    if (src.vec) {
        const int nv = nfl >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(g);
        float4* s4 = reinterpret_cast<float4*>(s);
        for (int i = lane; i < nv; i += kTile) ldgsts16(s4 + i, g4 + i);
        for (int i = (nv << 2) + lane; i < nfl; i += kTile) ldgsts4(s + i, g + i);
    } else {
        for (int i = lane; i < nfl; i += kTile) {
            int r = i / NCOLS, c = i - r * NCOLS;
            ldgsts4(s + i, g + (long long)r * src.stride + c);
        }
    }
}
// `ss` = row stride of the shared image (NCOLS unless the tile was staged as a wide slab)
template <int NCOLS>
DHFK_DI void stage_rows_out(const float* s, const RowDst& dst, long long row0, int rows, int ss = NCOLS) {
    const int lane = threadIdx.x;
    float* g = dst.p + row0 * dst.stride;
    const int nfl = rows * NCOLS;
    if (ss != NCOLS) {
        for (int i = lane; i < nfl; i += kTile) {
            int r = i / NCOLS, c = i - r * NCOLS;
            g[(long long)r * dst.stride + c] = s[r * ss + c];
        }
    } else if (dst.vec) {
        const int nv = nfl >> 2;
        float4* g4 = reinterpret_cast<float4*>(g);
        const float4* s4 = reinterpret_cast<const float4*>(s);
        for (int i = lane; i < nv; i += kTile) __stcs(g4 + i, s4[i]);
        for (int i = (nv << 2) + lane; i < nfl; i += kTile) __stcs(g + i, s[i]);
    } else {
        for (int i = lane; i < nfl; i += kTile) {
            int r = i / NCOLS, c = i - r * NCOLS;
            g[(long long)r * dst.stride + c] = s[i];
        }
    }
}
template <int CH>
DHFK_DI void stage_padded_in(float4* s4, const float* gbase, long long row0, int rows) {
    const float4* g4 = reinterpret_cast<const float4*>(gbase) + row0 * CH;
    for (int i = threadIdx.x; i < rows * CH; i += kTile) {   // cp.async: the caller waits with ldgsts_wait_all()
        int r = i / CH, c = i - r * CH;
        ldgsts16(s4 + r * (CH + 1) + c, g4 + i);
    }
}
template <int CH>
DHFK_DI void stage_padded_out(const float4* s4, float* gbase, long long row0, int rows) {
    float4* g4 = reinterpret_cast<float4*>(gbase) + row0 * CH;
    for (int i = threadIdx.x; i < rows * CH; i += kTile) {
        int r = i / CH, c = i - r * CH;
        __stcs(g4 + i, s4[r * (CH + 1) + c]);
    }
}

// GEN mode: the raw network output slab becomes tanh(slab) in place, ONCE per column, by the whole warp: 128-bit
// shared loads / stores over the contiguous slab (conflict-free), packed arithmetic, 2 MUFU per element.  The tree
// walk then reads t and forms angle = t * half + mid with one FFMA; the backward gets sech^2 = 1 - t^2 from the same
// cell.  (r1 decoded each column where the walk consumed it with scalar code: ~11 instructions per column and pass;
// generator mode measured 0.2262 ms per fwd+bwd step against 0.2027 ms for the plain step.)
DHFK_DI void tanh_slab_inplace(float* s) {
    constexpr int NV = kTile * GEN_NCOL / 4;
    static_assert(kTile * GEN_NCOL % 4 == 0, "slab must be a whole number of 16-byte chunks");
    float4* s4 = reinterpret_cast<float4*>(s);
    const int lane = threadIdx.x;
#pragma unroll
    for (int k = 0; k < (NV + 31) / 32; ++k) {
        const int i = lane + 32 * k;
        if (NV % 32 == 0 || i < NV) {
            const float4 v = s4[i];
            const float2 lo = tanh2(make_float2(v.x, v.y)), hi = tanh2(make_float2(v.z, v.w));
            s4[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
    }
}

// global rotation R = Rx(gx) Ry(gy) Rz(gz), forward_kinematics_DH_model.py:141-191 (row-major)
template <int TRIG>
DHFK_DI void global_rotation(const float* g, float* R, float& sx, float& cx, float& sy, float& cy) {
    float sz, cz;
    sincos_deg<TRIG, 0>(g[0], sx, cx);
    sincos_deg<TRIG, 0>(g[1], sy, cy);
    sincos_deg<TRIG, 0>(g[2], sz, cz);
    R[0] = cy * cz;                      R[1] = -cy * sz;                     R[2] = sy;
    R[3] = fmaf(sx * sy, cz, cx * sz);   R[4] = fmaf(-sx * sy, sz, cx * cz);  R[5] = -sx * cy;
    R[6] = fmaf(-cx * sy, cz, sx * sz);  R[7] = fmaf(cx * sy, sz, sx * cz);   R[8] = cx * cy;
}

// ---- forward -------------------------------------------------------------------------------------
template <bool CAM, bool UV, bool GEN>
struct FwdCtx {
    const float* ang;     // this pose's 33 angles, or (GEN) tanh of its 35 raw network outputs (tanh_slab_inplace)
    const float* bone;
    const CamConst* cc;
    const GenScale* gs;

    template <int J>
    DHFK_DI float angle() const {
        if constexpr (!GEN) return ang[J];
        else {
            constexpr int SRC = gen_src_col(J);
            if constexpr (SRC < 0) return gs->hm[J].y;
            else return fmaf(ang[SRC], gs->hm[J].x, gs->hm[J].y);
        }
    }
    float R[9];
    V3 root;
    float w[48];
    float cm[CAM ? 48 : 1];
    float uv[UV ? 32 : 1];

    template <int K>
    DHFK_DI void emit(V3 o) {
        constexpr bool kZero = origin_is_zero(OUT16[K]);
        V3 W;                       // the chain runs in world axes (base frame = columns of R): no R*o per joint
        if constexpr (kZero) W = root;
        else W = o + root;
        w[3 * K] = W.x; w[3 * K + 1] = W.y; w[3 * K + 2] = W.z;
        if (CAM || UV) {
            V3 X = cam_space(*cc, W);
            if (CAM) { cm[3 * K] = X.x; cm[3 * K + 1] = X.y; cm[3 * K + 2] = X.z; }
            if (UV) {
                ProjAux a;
                project_point(*cc, X, uv[2 * K], uv[2 * K + 1], a);
            }
        }
    }
};

template <int LO, int HI>
DHFK_DI void flush_chunks(float4* row4, const float* v) {
#pragma unroll
    for (int c = LO; c < HI; ++c) row4[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

template <bool CAM, bool UV, int TRIG, bool GEN, bool WIDE = false>
__global__ void __launch_bounds__(kTile) dhfk_fwd_kernel(const __grid_constant__ FwdParams p) {
    static_assert(!(GEN && WIDE), "wide rows are a raw-mode layout");
    constexpr int NANG = GEN ? GEN_NCOL : 33;      // floats per pose in the first slab
    const int wide = WIDE ? p.w.wide : 0;          // launch-uniform
    extern __shared__ __align__(16) float smem[];
    float* s_ang = smem;
    float* s_grot = s_ang + kTile * (wide ? wide : NANG);
    float* s_bone = s_grot + ((GEN || (WIDE && !p.w.grot_slab)) ? 0 : kTile * 3);
    float* s_root = s_bone + kTile * 15;
    float4* s_world = reinterpret_cast<float4*>(s_root + (GEN ? 0 : kTile * 3));
    float4* s_cam = s_world + kTile * kWorldRow4;
    float4* s_uv = s_cam + (CAM ? kTile * kWorldRow4 : 0);

    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    // full tile of packed, aligned rows: async slab path; ragged last tile / strided views: gather path
    const bool wtile = wide != 0 && rows == kTile && p.bone.vec && p.root.vec;     // this tile is staged as a wide slab
    const bool bulk = wtile || (rows == kTile && p.ang.vec && p.bone.vec && (GEN || (p.grot.vec && p.root.vec)));
    const int sa = wtile ? wide : NANG;            // row stride of the first slab, in global and in shared memory

    if (bulk) {
        if (wtile) ldgsts_slab_rt(s_ang, p.ang.p + row0 * sa, kTile / 4 * sa);
        else ldgsts_slab<NANG>(s_ang, p.ang.p + row0 * NANG);
        ldgsts_slab<15>(s_bone, p.bone.p + row0 * 15);
        if (!GEN) {
            if (!wtile) ldgsts_slab<3>(s_grot, p.grot.p + row0 * 3);
            ldgsts_slab<3>(s_root, p.root.p + row0 * 3);
        }
#if DHFK_PREFETCH_TILES_FWD > 0
        if (lane == 0) {
            const long long rowp = row0 + (long long)DHFK_PREFETCH_TILES_FWD * kTile;
            if (rowp + kTile <= p.n) {   // the warp one resident wave later finds its slabs in L2
                bulk_prefetch_l2(p.ang.p + rowp * sa, kTile * sa * 4);
                bulk_prefetch_l2(p.bone.p + rowp * 15, kTile * 15 * 4);
                if (!GEN) {
                    if (!wtile) bulk_prefetch_l2(p.grot.p + rowp * 3, kTile * 3 * 4);
                    bulk_prefetch_l2(p.root.p + rowp * 3, kTile * 3 * 4);
                }
            }
        }
#endif
        ldgsts_wait_all();
    } else {
        stage_rows_in<NANG>(s_ang, p.ang, row0, rows);
        stage_rows_in<15>(s_bone, p.bone, row0, rows);
        if (!GEN) {
            stage_rows_in<3>(s_grot, p.grot, row0, rows);
            stage_rows_in<3>(s_root, p.root, row0, rows);
        }
        ldgsts_wait_all();
    }
    __syncwarp();
    if (GEN) {
        tanh_slab_inplace(s_ang);
        __syncwarp();
    }

    if (lane < rows) {
        FwdCtx<CAM, UV, GEN> ctx;
        ctx.gs = &p.gs;
        ctx.ang = s_ang + lane * sa;
        ctx.bone = s_bone + lane * 15;
        ctx.cc = &p.cam;
        float sx, cx, sy, cy;
        if (GEN) {
            // global rotation = slots 34..36 (columns 28..30), root = tanh(columns 32..34) * 10
            float g[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                g[i] = fmaf(ctx.ang[gen_src_col(GEN_GROT_SLOT) + i], p.gs.hm[GEN_GROT_SLOT + i].x,
                            p.gs.hm[GEN_GROT_SLOT + i].y);
            global_rotation<TRIG>(g, ctx.R, sx, cx, sy, cy);
            ctx.root = v3(ctx.ang[GEN_ROOT_COL] * p.gs.root_scale, ctx.ang[GEN_ROOT_COL + 1] * p.gs.root_scale,
                          ctx.ang[GEN_ROOT_COL + 2] * p.gs.root_scale);
        } else {
            global_rotation<TRIG>(wtile ? ctx.ang + p.w.goff : s_grot + lane * 3, ctx.R, sx, cx, sy, cy);
            ctx.root = v3(s_root[lane * 3], s_root[lane * 3 + 1], s_root[lane * 3 + 2]);
        }
        float4* wrow = s_world + lane * kWorldRow4;
        float4* crow = s_cam + lane * kWorldRow4;
        float4* urow = s_uv + lane * kUvRow4;
        Frame I;                    // chain root frame: world axes rotated by the global rotation, origin = root point
        I.X = v3(ctx.R[0], ctx.R[3], ctx.R[6]); I.Y = v3(ctx.R[1], ctx.R[4], ctx.R[7]);
        I.Z = v3(ctx.R[2], ctx.R[5], ctx.R[8]); I.O = v3(0.f, 0.f, 0.f);
        // body, head, arms: outputs 0,7,8,9,13,14,15,10,11,12 -> joints 8..15 complete
        fwd_walk<TRIG, 10>(I, ctx);
        flush_chunks<6, 12>(wrow, ctx.w);
        if (CAM) flush_chunks<6, 12>(crow, ctx.cm);
        if (UV) flush_chunks<4, 8>(urow, ctx.uv);
        // right leg: outputs 1,2,3 -> joints 0..3 complete
        fwd_walk<TRIG, 0>(I, ctx);
        flush_chunks<0, 3>(wrow, ctx.w);
        if (CAM) flush_chunks<0, 3>(crow, ctx.cm);
        if (UV) flush_chunks<0, 2>(urow, ctx.uv);
        // left leg: outputs 4,5,6 -> joints 4..7 complete
        fwd_walk<TRIG, 5>(I, ctx);
        flush_chunks<3, 6>(wrow, ctx.w);
        if (CAM) flush_chunks<3, 6>(crow, ctx.cm);
        if (UV) flush_chunks<2, 4>(urow, ctx.uv);
    }
    __syncwarp();

    if (rows == kTile) {
        store_padded_tile<kWorldChunks>(s_world, p.out_world, row0);
        if (CAM) store_padded_tile<kWorldChunks>(s_cam, p.out_cam, row0);
        if (UV) store_padded_tile<kUvChunks>(s_uv, p.out_uv, row0);
    } else {
        stage_padded_out<kWorldChunks>(s_world, p.out_world, row0, rows);
        if (CAM) stage_padded_out<kWorldChunks>(s_cam, p.out_cam, row0, rows);
        if (UV) stage_padded_out<kUvChunks>(s_uv, p.out_uv, row0, rows);
    }
}

// ---- backward ------------------------------------------------------------------------------------
// GW / GCAM / GUV: which upstream gradients exist -- compile-time, so absent ones cost no zero-fills, no
// predicated loads and no null tests in the limb loop (measured: -110 instructions per warp vs runtime tests)
template <bool GW, bool GCAM, bool GUV, bool GBONE, bool GEN>
struct BwdCtx {
    static constexpr bool kBoneGrad = GBONE;
    const float* ang;   // this pose's 33 angles, or (GEN) tanh of its 35 raw network outputs
    const float* bone;
    float* g_ang;   // same shared row as ang (in place)
    float* g_bone;  // same shared row as bone (in place)
    const CamConst* cc;
    const GenScale* gs;

    // Joint angle in degrees.  GEN: the cell holds t = tanh(network output) (tanh_slab_inplace); the angle is the affine
    // slot map of t, and grad_angle turns d/d(angle) into d/d(network output) = g * half * (1 - t^2) in place.
    template <int J>
    DHFK_DI float angle() {
        if constexpr (!GEN) return ang[J];
        else {
            constexpr int SRC = gen_src_col(J);
            if constexpr (SRC < 0) return gs->hm[J].y;
            else return fmaf(ang[SRC], gs->hm[J].x, gs->hm[J].y);
        }
    }
    // joint K (0..3) of limb L: the limb is a run-time value (warp-uniform), the position inside it is not.  GEN: the
    // column comes from the limb descriptor (uniform datapath) and (half, mid) from one 64-bit constant load.
    template <int K>
    DHFK_DI float limb_angle(const LimbDesc& L) {
        if (!GEN) return ang[L.ang0 + K];
        const float2 hm = gs->hm[L.ang0 + K];
        if (K == 0) return L.gcol0 < 0 ? hm.y : fmaf(ang[L.gcol0], hm.x, hm.y);
        return fmaf(ang[L.gcol1 + (K - 1)], hm.x, hm.y);
    }
    template <int K>
    DHFK_DI void limb_grad(const LimbDesc& L, float g) {
        if (!GEN) { g_ang[L.ang0 + K] = g; return; }
        const int col = K == 0 ? L.gcol0 : L.gcol1 + (K - 1);
        if (K == 0 && col < 0) return;
        g_ang[col] = g * gs->hm[L.ang0 + K].x * sech2_from_tanh(g_ang[col]);
    }
    DHFK_DI void limb_zero_leaf(const LimbDesc& L) {
        if (!GEN) { g_ang[L.ang0 + 4] = 0.f; return; }
        if (L.gcol4 >= 0) g_ang[L.gcol4] = 0.f;
    }
    template <int J>
    DHFK_DI void grad_angle(float g) {
        if constexpr (!GEN) g_ang[J] = g;
        else {
            constexpr int SRC = gen_src_col(J);
            if constexpr (SRC >= 0) g_ang[SRC] = g * gs->hm[J].x * sech2_from_tanh(g_ang[SRC]);
        }
    }
    template <int J>
    DHFK_DI void zero_grad_angle() {
        if constexpr (!GEN) g_ang[J] = 0.f;
        else {
            constexpr int SRC = gen_src_col(J);
            if constexpr (SRC >= 0) g_ang[SRC] = 0.f;
        }
    }
    const float4* gw4;   // padded shared rows of the upstream gradients (may be null)
    const float4* gc4;
    const float4* gu4;
    float R[9];
    V3 root;
    // The chain is walked directly in the frame the upstream gradients live in, with the root point as origin:
    // camera axes (base = columns of M*R) when any camera-space gradient exists, world axes (columns of R)
    // otherwise.  Positions then need only "+ v0" to become camera coordinates and gradients need no
    // per-joint rotation back into a chain frame.
    Frame base;
    static constexpr bool cam_frame = GUV || GCAM;
    V3 v0;         // root point in camera coordinates, M * (root - t)
    Wrench legs;   // filled by bwd_all_limbs
    // d/d root = sum over the 16 outputs of the world-space gradient: see root_grad()
    V3 sum_gw;

    DHFK_DI void setup_camera() {
        sum_gw = v3(0.f, 0.f, 0.f);
        base.O = v3(0.f, 0.f, 0.f);
        if constexpr (cam_frame) {
            float MR[9];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    MR[3 * i + j] = fmaf(cc->M[3 * i], R[j], fmaf(cc->M[3 * i + 1], R[3 + j], cc->M[3 * i + 2] * R[6 + j]));
            base.X = v3(MR[0], MR[3], MR[6]); base.Y = v3(MR[1], MR[4], MR[7]); base.Z = v3(MR[2], MR[5], MR[8]);
            v0 = cam_space(*cc, root);
        } else {
            base.X = v3(R[0], R[3], R[6]); base.Y = v3(R[1], R[4], R[7]); base.Z = v3(R[2], R[5], R[8]);
            v0 = v3(0.f, 0.f, 0.f);
        }
    }
    // total dL/d(origin) in the working frame from the world-space gradient g and the camera-space gradient gc.
    // Only the world-space part is summed on the side (for d/d root); the camera-space sum falls out of the
    // total force at the end:  sum_gc = F_total - M sum_gw.
    DHFK_DI V3 to_frame(V3 g, V3 gc) {
        if constexpr (!cam_frame) return g;
        else if constexpr (GW) {
            sum_gw = sum_gw + g;
            return cam_mat_vec_add(*cc, g, gc);    // M g_w + g_c
        } else return gc;
    }
    // d/d root (world axes) from the total force over the 16 outputs (working frame)
    DHFK_DI V3 root_grad(V3 Ft) const {
        if constexpr (!cam_frame) return Ft;                       // working frame = world axes
        else if constexpr (!GW) return matT_vec(cc->M, Ft);        // M^T sum_gc
        else {
            const V3 mg = mat_vec(cc->M, sum_gw);
            return matT_vec_add(cc->M, v3(Ft.x - mg.x, Ft.y - mg.y, Ft.z - mg.z), sum_gw);
        }
    }

    // 3 consecutive floats starting at float index 3K of a padded row, via 128-bit loads only
    template <int K>
    DHFK_DI V3 load3(const float4* row) const {
        constexpr int F0 = 3 * K, C0 = F0 / 4, OFF = F0 % 4;
        float4 a = row[C0];
        if (OFF == 0) return v3(a.x, a.y, a.z);
        if (OFF == 1) return v3(a.y, a.z, a.w);
        float4 b = row[C0 + 1];
        if (OFF == 2) return v3(a.z, a.w, b.x);
        return v3(a.w, b.x, b.y);
    }

    // total dL/d(origin of output K) in the working frame; o = origin relative to the root, working axes
    template <int K>
    DHFK_DI V3 upstream(V3 o) {
        V3 g = v3(0.f, 0.f, 0.f);
        if constexpr (GW) g = load3<K>(gw4);
        V3 gc = v3(0.f, 0.f, 0.f);
        if constexpr (GCAM) gc = load3<K>(gc4);
        if (GUV) {
            constexpr bool kZero = origin_is_zero(OUT16[K]);
            V3 X;
            if constexpr (kZero) X = v0;
            else X = o + v0;
            float u, v;
            ProjAux a;
            project_point(*cc, X, u, v, a);
            float4 q = gu4[K / 2];
            gc = gc + project_point_bwd(*cc, a, (K & 1) ? q.z : q.x, (K & 1) ? q.w : q.y);
        }
        return to_frame(g, gc);
    }
    // The shared limb routine (runtime limb index, warp-uniform): the limb's three outputs are 9 consecutive floats of
    // the [16,3] rows and 6 of the [16,2] row.  Rows are padded for 128-bit access (odd stride in 16-byte chunks), so a
    // 32- or 64-bit load at a runtime offset is a 4- / 2-way bank conflict whatever the padding (every lane's address
    // is congruent mod 16 bytes): r1 measured 46 % of the shared wavefronts as conflicts.  Instead the 3 (2) chunks
    // that hold them are read with LDS.128 and the floats picked with a warp-uniform offset.
    float lw[GW ? 9 : 1], lc[GCAM ? 9 : 1], lu[GUV ? 6 : 1];
    // ld.shared.v4 spelled out: left to itself the compiler narrows the first chunk to a 64-bit + conditional 32-bit
    // loads (only part of it is used for two of the three offsets), which brings the conflicts back
    DHFK_DI static float4 lds128(const float4* p) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
        return v;
    }
    DHFK_DI static void pick9(const float4* row, int c0, int off, float* o) {
        const float4 a = lds128(row + c0), b = lds128(row + c0 + 1), c = lds128(row + c0 + 2);
        if (off == 0) {
            o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w; o[8] = c.x;
        } else if (off == 2) {
            o[0] = a.z; o[1] = a.w; o[2] = b.x; o[3] = b.y; o[4] = b.z; o[5] = b.w; o[6] = c.x; o[7] = c.y; o[8] = c.z;
        } else {   // off == 3
            o[0] = a.w; o[1] = b.x; o[2] = b.y; o[3] = b.z; o[4] = b.w; o[5] = c.x; o[6] = c.y; o[7] = c.z; o[8] = c.w;
        }
    }
    DHFK_DI void load_limb(const LimbDesc& L) {
        if constexpr (GW) pick9(gw4, L.wc0, L.woff, lw);
        if constexpr (GCAM) pick9(gc4, L.wc0, L.woff, lc);
        if constexpr (GUV) {
            const float4 a = lds128(gu4 + L.uc0), b = lds128(gu4 + L.uc0 + 1);
            if (L.uoff == 0) { lu[0] = a.x; lu[1] = a.y; lu[2] = a.z; lu[3] = a.w; lu[4] = b.x; lu[5] = b.y; }
            else { lu[0] = a.z; lu[1] = a.w; lu[2] = b.x; lu[3] = b.y; lu[4] = b.z; lu[5] = b.w; }
        }
    }
    // total dL/d(origin of the limb's I-th output) in the working frame
    template <int I>
    DHFK_DI V3 upstream_limb(V3 o) {
        V3 g = v3(0.f, 0.f, 0.f);
        if constexpr (GW) g = v3(lw[3 * I], lw[3 * I + 1], lw[3 * I + 2]);
        V3 gc = v3(0.f, 0.f, 0.f);
        if constexpr (GCAM) gc = v3(lc[3 * I], lc[3 * I + 1], lc[3 * I + 2]);
        if (GUV) {
            V3 X = o + v0;
            float u, v;
            ProjAux a;
            project_point(*cc, X, u, v, a);
            gc = gc + project_point_bwd(*cc, a, lu[2 * I], lu[2 * I + 1]);
        }
        return to_frame(g, gc);
    }
    DHFK_DI void grad_bone(int b, float g) { g_bone[b] = g; }
};

template <bool GW, bool GCAM, bool GUV, bool GBONE, int TRIG, bool GEN, bool WIDE = false>
__global__ void __launch_bounds__(kTile) dhfk_bwd_kernel(const __grid_constant__ BwdParams p) {
    static_assert(!(GEN && WIDE), "wide rows are a raw-mode layout");
    static_assert(GW || GCAM || GUV, "at least one upstream gradient");
    static_assert(!(GEN && GBONE), "bone-length gradients are not produced in generator mode");
    constexpr int NANG = GEN ? GEN_NCOL : 33;
    const int wide = WIDE ? p.w.wide : 0;          // launch-uniform (see WideRows)
    extern __shared__ __align__(16) float smem[];
    float* s_ang = smem;
    float* s_grot = s_ang + kTile * (wide ? wide : NANG);
    float* s_bone = s_grot + ((GEN || (WIDE && !p.w.grot_slab)) ? 0 : kTile * 3);
    float* s_root = s_bone + kTile * 15;
    float4* s_gw = reinterpret_cast<float4*>(s_root + (GEN ? 0 : kTile * 3));
    float4* s_gc = s_gw + (GW ? kTile * kWorldRow4 : 0);
    float4* s_gu = s_gc + (GCAM ? kTile * kWorldRow4 : 0);

    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    const bool wtile = wide != 0 && rows == kTile && p.bone.vec && p.root.vec;
    const bool bulk = wtile || (rows == kTile && p.ang.vec && p.bone.vec && (GEN || (p.grot.vec && p.root.vec)));
    const int sa = wtile ? wide : NANG;

    if (bulk) {
        // one round trip: every byte of the tile is requested before anything is waited for
        if (wtile) ldgsts_slab_rt(s_ang, p.ang.p + row0 * sa, kTile / 4 * sa);
        else ldgsts_slab<NANG>(s_ang, p.ang.p + row0 * NANG);
        ldgsts_slab<15>(s_bone, p.bone.p + row0 * 15);
        if (!GEN) {
            if (!wtile) ldgsts_slab<3>(s_grot, p.grot.p + row0 * 3);
            ldgsts_slab<3>(s_root, p.root.p + row0 * 3);
        }
        if (GW) ldgsts_padded_tile<kWorldChunks>(s_gw, p.g_world, row0);
        if (GCAM) ldgsts_padded_tile<kWorldChunks>(s_gc, p.g_cam, row0);
        if (GUV) ldgsts_padded_tile<kUvChunks>(s_gu, p.g_uv, row0);
#if DHFK_PREFETCH_TILES_BWD > 0
        if (lane == 0) {
            const long long rowp = row0 + (long long)DHFK_PREFETCH_TILES_BWD * kTile;
            if (rowp + kTile <= p.n) {
                bulk_prefetch_l2(p.ang.p + rowp * sa, kTile * sa * 4);
                bulk_prefetch_l2(p.bone.p + rowp * 15, kTile * 15 * 4);
                if (!GEN) {
                    if (!wtile) bulk_prefetch_l2(p.grot.p + rowp * 3, kTile * 3 * 4);
                    bulk_prefetch_l2(p.root.p + rowp * 3, kTile * 3 * 4);
                }
                if (GW) bulk_prefetch_l2(p.g_world + rowp * 48, kTile * 48 * 4);
                if (GCAM) bulk_prefetch_l2(p.g_cam + rowp * 48, kTile * 48 * 4);
                if (GUV) bulk_prefetch_l2(p.g_uv + rowp * 32, kTile * 32 * 4);
            }
        }
#endif
        ldgsts_wait_all();
    } else {
        stage_rows_in<NANG>(s_ang, p.ang, row0, rows);
        stage_rows_in<15>(s_bone, p.bone, row0, rows);
        if (!GEN) {
            stage_rows_in<3>(s_grot, p.grot, row0, rows);
            stage_rows_in<3>(s_root, p.root, row0, rows);
        }
        if (GW) stage_padded_in<kWorldChunks>(s_gw, p.g_world, row0, rows);
        if (GCAM) stage_padded_in<kWorldChunks>(s_gc, p.g_cam, row0, rows);
        if (GUV) stage_padded_in<kUvChunks>(s_gu, p.g_uv, row0, rows);
        ldgsts_wait_all();
    }
    __syncwarp();
    if (GEN) {
        tanh_slab_inplace(s_ang);
        __syncwarp();
    }

    if (lane < rows) {
        BwdCtx<GW, GCAM, GUV, GBONE, GEN> ctx;
        ctx.gs = &p.gs;
        ctx.ang = s_ang + lane * sa;
        ctx.g_ang = s_ang + lane * sa;
        ctx.bone = s_bone + lane * 15;
        ctx.g_bone = s_bone + lane * 15;
        ctx.cc = &p.cam;
        ctx.gw4 = GW ? s_gw + lane * kWorldRow4 : nullptr;
        ctx.gc4 = GCAM ? s_gc + lane * kWorldRow4 : nullptr;
        ctx.gu4 = GUV ? s_gu + lane * kUvRow4 : nullptr;
        float sx, cx, sy, cy;
        float chain_g[3], chain_r[3];   // GEN: d(grot_i)/d(col), d(root_i)/d(col)
        if (GEN) {
            float g[3], r[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float tg = ctx.ang[gen_src_col(GEN_GROT_SLOT) + i], tr = ctx.ang[GEN_ROOT_COL + i];
                g[i] = fmaf(tg, p.gs.hm[GEN_GROT_SLOT + i].x, p.gs.hm[GEN_GROT_SLOT + i].y);
                chain_g[i] = sech2_from_tanh(tg) * p.gs.hm[GEN_GROT_SLOT + i].x;
                r[i] = tr * p.gs.root_scale;
                chain_r[i] = sech2_from_tanh(tr) * p.gs.root_scale;
            }
            global_rotation<TRIG>(g, ctx.R, sx, cx, sy, cy);
            ctx.root = v3(r[0], r[1], r[2]);
        } else {
            global_rotation<TRIG>(wtile ? ctx.ang + p.w.goff : s_grot + lane * 3, ctx.R, sx, cx, sy, cy);
            ctx.root = v3(s_root[lane * 3], s_root[lane * 3 + 1], s_root[lane * 3 + 2]);
        }
        ctx.setup_camera();
        // body + head chain unrolled; at joint 18 the walker runs the shared limb loop (arms AND legs)
        Wrench wb = bwd_walk<TRIG, 10>(ctx.base, ctx);
        V3 Ft = wb.F + ctx.legs.F;
        V3 Mt = wb.M + ctx.legs.M;
        const V3 gr = ctx.root_grad(Ft);
        // d/d global angles: torque about the world axes e_x, Rx e_y, Rx Ry e_z
        V3 tw = Mt;                                      // moment about the root, working axes -> world axes
        if (ctx.cam_frame) tw = matT_vec(p.cam.M, Mt);
        const float gg0 = kDegToRad * tw.x;
        const float gg1 = kDegToRad * fmaf(cx, tw.y, sx * tw.z);
        const float gg2 = kDegToRad * fmaf(sy, tw.x, fmaf(-sx * cy, tw.y, cx * cy * tw.z));
        if (GEN) {
            float* o = ctx.g_ang;
            constexpr int C = gen_src_col(GEN_GROT_SLOT);
            o[C] = gg0 * chain_g[0]; o[C + 1] = gg1 * chain_g[1]; o[C + 2] = gg2 * chain_g[2];
            o[GEN_UNUSED_COL] = 0.f;      // column 31 never reaches a slot
            o[GEN_ROOT_COL] = gr.x * chain_r[0]; o[GEN_ROOT_COL + 1] = gr.y * chain_r[1];
            o[GEN_ROOT_COL + 2] = gr.z * chain_r[2];
        } else {
            s_root[lane * 3] = gr.x; s_root[lane * 3 + 1] = gr.y; s_root[lane * 3 + 2] = gr.z;
            if (wtile) {
                // the row image becomes the gradient of the whole wide row: angles, global rotation at `goff`, and
                // zeros in the columns that feed nothing
                float* row = s_ang + lane * sa;
                for (int c = 33; c < sa; ++c) row[c] = 0.f;
                row[p.w.goff] = gg0; row[p.w.goff + 1] = gg1; row[p.w.goff + 2] = gg2;
            } else {
                s_grot[lane * 3] = gg0; s_grot[lane * 3 + 1] = gg1; s_grot[lane * 3 + 2] = gg2;
            }
        }
    }
    const bool wide_out = wtile && p.g_wide;       // one slab store carries d(angles) and d(global rotation)
    const bool bulk_out = wide_out || (!wtile && rows == kTile && p.g_ang.vec &&
                                       (GEN || (p.g_grot.vec && p.g_root.vec && (!GBONE || p.g_bone.vec))));
    if (bulk_out && (!wide_out || (p.g_root.vec && (!GBONE || p.g_bone.vec)))) {
        // results overwrote the input slabs in place; lane 0 ships the slabs
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(p.g_ang.p + row0 * sa, s_ang, kTile * sa * 4);
            if (!GEN) {
                if (!wide_out) bulk_s2g(p.g_grot.p + row0 * 3, s_grot, kTile * 3 * 4);
                bulk_s2g(p.g_root.p + row0 * 3, s_root, kTile * 3 * 4);
                if (GBONE) bulk_s2g(p.g_bone.p + row0 * 15, s_bone, kTile * 15 * 4);
            }
            bulk_commit();
            bulk_wait_read_all();
        }
        return;
    }
    __syncwarp();
    stage_rows_out<NANG>(s_ang, p.g_ang, row0, rows, sa);
    if (!GEN) {
        if (wtile) stage_rows_out<3>(s_ang + p.w.goff, p.g_grot, row0, rows, sa);
        else stage_rows_out<3>(s_grot, p.g_grot, row0, rows);
        stage_rows_out<3>(s_root, p.g_root, row0, rows);
        if (GBONE) stage_rows_out<15>(s_bone, p.g_bone, row0, rows);
    }
}

}  // namespace dhfk
