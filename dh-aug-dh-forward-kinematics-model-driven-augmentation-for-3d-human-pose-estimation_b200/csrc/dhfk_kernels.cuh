// dhfk_kernels.cuh -- fused DH-FK + global rotation + world->camera + pinhole projection,
// forward and analytic backward, for sm_100a.  One thread per pose, 96 poses per CTA.
//
// Data movement (see DESIGN.md "HBM layout"):
//   inputs  [N,33] [N,3] [N,15] [N,3] are AoS rows with odd lengths.  A tile of 96 rows of each
//           is one contiguous, 16-byte aligned slab; the CTA copies it to shared memory with
//           coalesced 128-bit loads (exact image, odd row stride => conflict-free per-thread
//           scalar reads).
//   outputs [N,16,3] / [N,16,2] (and upstream gradients in the backward) have 48 / 32 float
//           rows; they are staged in shared memory with the row stride padded to 13 / 9
//           16-byte chunks so that both the per-thread 128-bit accesses and the cooperative
//           coalesced 128-bit global accesses are bank-conflict free.  With 96 threads the
//           cooperative copy needs no index arithmetic (96 = 8*12 = 12*8).
//   backward results d(ang), d(grot), d(root), d(bone) overwrite the input slabs in place and
//           leave through the same coalesced path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dhfk_device.cuh"

namespace dhfk {

constexpr int kTile = 96;        // poses (threads) per CTA
constexpr int kWorldChunks = 12; // 48 floats
constexpr int kUvChunks = 8;     // 32 floats
constexpr int kWorldRow4 = kWorldChunks + 1;  // padded row stride in float4
constexpr int kUvRow4 = kUvChunks + 1;

static_assert(kTile % kWorldChunks == 0 && kTile % kUvChunks == 0, "tile must tile both row shapes");
static_assert(kTile % 4 == 0, "tile slabs must stay 16-byte aligned");
static_assert(nth_child(-1, 0) == 0 && nth_child(-1, 1) == 5 && nth_child(-1, 2) == 10 &&
              nth_child(-1, 3) == -1, "chain roots are joints 0, 5, 10");

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

struct FwdParams {
    RowSrc ang, grot, bone, root;
    float* out_world;
    float* out_cam;
    float* out_uv;
    long long n;
    CamConst cam;
};
struct BwdParams {
    RowSrc ang, grot, bone, root;
    const float* g_world;
    const float* g_cam;
    const float* g_uv;
    RowDst g_ang, g_grot, g_root, g_bone;
    long long n;
    CamConst cam;
};

// Tunables for A/B measurement (see profiles/): how padded rows enter / leave shared memory on the
// bulk path and who waits on the mbarrier.
#ifndef DHFK_ROWS_IN
#define DHFK_ROWS_IN 1    // 0: one cp.async.bulk per row and thread, 1: 16-byte LDGSTS by all threads
#endif
#ifndef DHFK_ROWS_OUT
#define DHFK_ROWS_OUT 1   // 0: one cp.async.bulk per row and thread, 1: cooperative LDS.128 -> STG.128
#endif
// L2 prefetch distance in tiles (148 SMs x 4 resident CTAs = one resident wave); 0 = off.
// Measured (profiles/r1_ab_staging.md): forward 0.110 -> 0.098 ms; backward unchanged/slightly worse
// (it is issue / i-cache bound, not load-latency bound), so it stays off there.
#ifndef DHFK_PREFETCH_TILES_FWD
#define DHFK_PREFETCH_TILES_FWD 592
#endif
#ifndef DHFK_PREFETCH_TILES_BWD
#define DHFK_PREFETCH_TILES_BWD 0
#endif
#ifndef DHFK_WAIT_ONE
#define DHFK_WAIT_ONE 1   // 1: thread 0 waits on the mbarrier, the CTA waits on bar.sync (no spinning)
#endif

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier -------------------------------------
// A full tile's inputs are fetched with ONE round trip: four slab copies issued by thread 0 and
// (backward) one row copy per thread and gradient tensor, all completing on a single mbarrier.
// Outputs leave the same way: each thread bulk-stores its own padded row(s); slabs by thread 0.
DHFK_DI uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DHFK_DI void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DHFK_DI void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DHFK_DI void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DHFK_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DHFK_DONE;\n"
        "bra DHFK_WAIT;\n"
        "DHFK_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
DHFK_DI void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
DHFK_DI void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
                 "r"(bytes) : "memory");
}
// Ampere-style 16-byte async copy (SASS LDGSTS) whose completion is tracked by an mbarrier
DHFK_DI void ldgsts16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
DHFK_DI void ldgsts_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int CH>
DHFK_DI void ldgsts_padded_rows(float4* s4, const float* gbase, long long row0) {
    constexpr int RPI = kTile / CH;
    const int tid = threadIdx.x;
    const int r0 = tid / CH, c0 = tid - r0 * CH;
    const float4* g4 = reinterpret_cast<const float4*>(gbase) + row0 * CH + r0 * CH + c0;
    float4* d4 = s4 + r0 * (CH + 1) + c0;
#pragma unroll
    for (int m = 0; m < CH; ++m) ldgsts16(d4 + m * RPI * (CH + 1), g4 + m * RPI * CH);
}
// L2 prefetch of a contiguous global range (multiple of 16 bytes, 16-byte aligned)
DHFK_DI void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
DHFK_DI void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DHFK_DI void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make this thread's generic-proxy shared-memory writes visible to the async proxy (TMA)
DHFK_DI void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- shared <-> global staging -----------------------------------------------------------------
// exact-image rows (row stride NCOLS in shared)
template <int NCOLS>
DHFK_DI void stage_rows_in(float* s, const RowSrc& src, long long row0, int rows) {
    const int tid = threadIdx.x;
    const float* g = src.p + row0 * src.stride;
    if (src.vec) {
        const float4* g4 = reinterpret_cast<const float4*>(g);
        float4* s4 = reinterpret_cast<float4*>(s);
        if (rows == kTile) {
            constexpr int NV = kTile * NCOLS / 4;
#pragma unroll
            for (int k = 0; k < (NV + kTile - 1) / kTile; ++k) {
                int i = tid + k * kTile;
                if (i < NV) s4[i] = __ldcs(g4 + i);
            }
        } else {
            const int nfl = rows * NCOLS, nv = nfl >> 2;
            for (int i = tid; i < nv; i += kTile) s4[i] = __ldcs(g4 + i);
            for (int i = (nv << 2) + tid; i < nfl; i += kTile) s[i] = __ldcs(g + i);
        }
    } else {
        const int nfl = rows * NCOLS;
        for (int i = tid; i < nfl; i += kTile) {
            int r = i / NCOLS, c = i - r * NCOLS;
            s[i] = __ldg(g + (long long)r * src.stride + c);
        }
    }
}
template <int NCOLS>
DHFK_DI void stage_rows_out(const float* s, const RowDst& dst, long long row0, int rows) {
    const int tid = threadIdx.x;
    float* g = dst.p + row0 * dst.stride;
    if (dst.vec) {
        float4* g4 = reinterpret_cast<float4*>(g);
        const float4* s4 = reinterpret_cast<const float4*>(s);
        if (rows == kTile) {
            constexpr int NV = kTile * NCOLS / 4;
#pragma unroll
            for (int k = 0; k < (NV + kTile - 1) / kTile; ++k) {
                int i = tid + k * kTile;
                if (i < NV) __stcs(g4 + i, s4[i]);
            }
        } else {
            const int nfl = rows * NCOLS, nv = nfl >> 2;
            for (int i = tid; i < nv; i += kTile) __stcs(g4 + i, s4[i]);
            for (int i = (nv << 2) + tid; i < nfl; i += kTile) __stcs(g + i, s[i]);
        }
    } else {
        const int nfl = rows * NCOLS;
        for (int i = tid; i < nfl; i += kTile) {
            int r = i / NCOLS, c = i - r * NCOLS;
            g[(long long)r * dst.stride + c] = s[i];
        }
    }
}
// padded rows: CH 16-byte chunks per row in global (packed, aligned), CH+1 in shared
template <int CH>
DHFK_DI void stage_padded_in(float4* s4, const float* gbase, long long row0, int rows) {
    constexpr int RPI = kTile / CH;
    const int tid = threadIdx.x;
    const int r0 = tid / CH, c0 = tid - r0 * CH;
    const float4* g4 = reinterpret_cast<const float4*>(gbase) + row0 * CH;
#pragma unroll
    for (int m = 0; m < CH; ++m) {
        int r = r0 + m * RPI;
        if (r < rows) s4[r * (CH + 1) + c0] = __ldcs(g4 + r * CH + c0);
    }
}
template <int CH>
DHFK_DI void stage_padded_out(const float4* s4, float* gbase, long long row0, int rows) {
    constexpr int RPI = kTile / CH;
    const int tid = threadIdx.x;
    const int r0 = tid / CH, c0 = tid - r0 * CH;
    float4* g4 = reinterpret_cast<float4*>(gbase) + row0 * CH;
#pragma unroll
    for (int m = 0; m < CH; ++m) {
        int r = r0 + m * RPI;
        if (r < rows) __stcs(g4 + r * CH + c0, s4[r * (CH + 1) + c0]);
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
template <bool CAM, bool UV>
struct FwdCtx {
    const float* ang;
    const float* bone;
    const CamConst* cc;
    float R[9];
    V3 root;
    float w[48];
    float cm[CAM ? 48 : 1];
    float uv[UV ? 32 : 1];

    template <int K>
    DHFK_DI void emit(V3 o) {
        constexpr bool kZero = origin_is_zero(OUT16[K]);
        V3 W;
        if constexpr (kZero) W = root;
        else W = mat_vec_add(R, o, root);
        w[3 * K] = W.x; w[3 * K + 1] = W.y; w[3 * K + 2] = W.z;
        if (CAM || UV) {
            V3 X = mat_vec(cc->M, v3(W.x - cc->t[0], W.y - cc->t[1], W.z - cc->t[2]));
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

template <bool CAM, bool UV, int TRIG>
__global__ void __launch_bounds__(kTile) dhfk_fwd_kernel(const __grid_constant__ FwdParams p) {
    extern __shared__ __align__(16) float smem[];
    float* s_ang = smem;
    float* s_grot = s_ang + kTile * 33;
    float* s_bone = s_grot + kTile * 3;
    float* s_root = s_bone + kTile * 15;
    float4* s_world = reinterpret_cast<float4*>(s_root + kTile * 3);
    float4* s_cam = s_world + kTile * kWorldRow4;
    float4* s_uv = s_cam + (CAM ? kTile * kWorldRow4 : 0);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_uv + (UV ? kTile * kUvRow4 : 0));

    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    // full tile of packed, aligned rows: TMA bulk path; ragged last tile / strided views: gather path
    const bool bulk = rows == kTile && p.ang.vec && p.grot.vec && p.bone.vec && p.root.vec;

    if (bulk) {
        if (tid == 0) mbar_init(s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(s_bar, kTile * 54 * 4);
            bulk_g2s(s_ang, p.ang.p + row0 * 33, kTile * 33 * 4, s_bar);
            bulk_g2s(s_bone, p.bone.p + row0 * 15, kTile * 15 * 4, s_bar);
            bulk_g2s(s_grot, p.grot.p + row0 * 3, kTile * 3 * 4, s_bar);
            bulk_g2s(s_root, p.root.p + row0 * 3, kTile * 3 * 4, s_bar);
#if DHFK_PREFETCH_TILES_FWD > 0
            const long long rowp = row0 + (long long)DHFK_PREFETCH_TILES_FWD * kTile;
            if (rowp + kTile <= p.n) {   // the CTA one resident wave later finds its slabs in L2
                bulk_prefetch_l2(p.ang.p + rowp * 33, kTile * 33 * 4);
                bulk_prefetch_l2(p.bone.p + rowp * 15, kTile * 15 * 4);
                bulk_prefetch_l2(p.grot.p + rowp * 3, kTile * 3 * 4);
                bulk_prefetch_l2(p.root.p + rowp * 3, kTile * 3 * 4);
            }
#endif
        }
#if DHFK_WAIT_ONE
        if (tid == 0) mbar_wait(s_bar, 0);
        __syncthreads();
#else
        mbar_wait(s_bar, 0);
#endif
    } else {
        stage_rows_in<33>(s_ang, p.ang, row0, rows);
        stage_rows_in<3>(s_grot, p.grot, row0, rows);
        stage_rows_in<15>(s_bone, p.bone, row0, rows);
        stage_rows_in<3>(s_root, p.root, row0, rows);
        __syncthreads();
    }

    if (tid < rows) {
        FwdCtx<CAM, UV> ctx;
        ctx.ang = s_ang + tid * 33;
        ctx.bone = s_bone + tid * 15;
        ctx.cc = &p.cam;
        float sx, cx, sy, cy;
        global_rotation<TRIG>(s_grot + tid * 3, ctx.R, sx, cx, sy, cy);
        ctx.root = v3(s_root[tid * 3], s_root[tid * 3 + 1], s_root[tid * 3 + 2]);
        float4* wrow = s_world + tid * kWorldRow4;
        float4* crow = s_cam + tid * kWorldRow4;
        float4* urow = s_uv + tid * kUvRow4;
        const Frame I = identity_frame();
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
        if (bulk && DHFK_ROWS_OUT == 0) {
            // every thread ships its own rows: no CTA barrier, no cooperative copy loop
            fence_proxy_async();
            bulk_s2g(p.out_world + (row0 + tid) * 48, wrow, 48 * 4);
            if (CAM) bulk_s2g(p.out_cam + (row0 + tid) * 48, crow, 48 * 4);
            if (UV) bulk_s2g(p.out_uv + (row0 + tid) * 32, urow, 32 * 4);
            bulk_commit();
            bulk_wait_read_all();
        }
    }
    if (bulk && DHFK_ROWS_OUT == 0) return;
    __syncthreads();

    stage_padded_out<kWorldChunks>(s_world, p.out_world, row0, rows);
    if (CAM) stage_padded_out<kWorldChunks>(s_cam, p.out_cam, row0, rows);
    if (UV) stage_padded_out<kUvChunks>(s_uv, p.out_uv, row0, rows);
}

// ---- backward ------------------------------------------------------------------------------------
template <bool GUV, bool GBONE>
struct BwdCtx {
    static constexpr bool kBoneGrad = GBONE;
    const float* ang;
    const float* bone;
    float* g_ang;   // same shared row as ang (in place)
    float* g_bone;  // same shared row as bone (in place)
    const CamConst* cc;
    const float4* gw4;   // padded shared rows of the upstream gradients (may be null)
    const float4* gc4;
    const float4* gu4;
    float R[9];
    V3 root;

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

    // total dL/d(origin of output K) rotated back into the chain frame
    template <int K>
    DHFK_DI V3 upstream(V3 o) const {
        V3 g = v3(0.f, 0.f, 0.f);
        if (gw4) g = load3<K>(gw4);            // block-uniform branch
        if (GUV || gc4) {
            V3 gc = v3(0.f, 0.f, 0.f);
            if (gc4) gc = load3<K>(gc4);       // block-uniform branch
            if (GUV) {
                constexpr bool kZero = origin_is_zero(OUT16[K]);
                V3 W;
                if constexpr (kZero) W = root;
                else W = mat_vec_add(R, o, root);
                V3 X = mat_vec(cc->M, v3(W.x - cc->t[0], W.y - cc->t[1], W.z - cc->t[2]));
                float u, v;
                ProjAux a;
                project_point(*cc, X, u, v, a);
                float4 q = gu4[K / 2];
                V3 gp = project_point_bwd(*cc, a, (K & 1) ? q.z : q.x, (K & 1) ? q.w : q.y);
                gc = gc + gp;
            }
            g = matT_vec_add(cc->M, gc, g);
        }
        return matT_vec(R, g);
    }
    DHFK_DI void grad_angle(int j, float g) { g_ang[j] = g; }
    DHFK_DI void grad_bone(int b, float g) { g_bone[b] = g; }
};

template <bool GUV, bool GBONE, int TRIG>
__global__ void __launch_bounds__(kTile) dhfk_bwd_kernel(const __grid_constant__ BwdParams p) {
    extern __shared__ __align__(16) float smem[];
    float* s_ang = smem;
    float* s_grot = s_ang + kTile * 33;
    float* s_bone = s_grot + kTile * 3;
    float* s_root = s_bone + kTile * 15;
    float4* s_gw = reinterpret_cast<float4*>(s_root + kTile * 3);
    const bool GW = p.g_world != nullptr, GCAM = p.g_cam != nullptr;
    float4* s_gc = s_gw + (GW ? kTile * kWorldRow4 : 0);
    float4* s_gu = s_gc + (GCAM ? kTile * kWorldRow4 : 0);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_gu + (GUV ? kTile * kUvRow4 : 0));

    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    const bool bulk = rows == kTile && p.ang.vec && p.grot.vec && p.bone.vec && p.root.vec;

    if (bulk) {
#if DHFK_ROWS_IN == 1
        if (tid == 0) mbar_init(s_bar, 1 + kTile);   // thread 0's expect_tx arrive + one LDGSTS arrive per thread
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(s_bar, kTile * 4 * 54);
#else
        if (tid == 0) mbar_init(s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(s_bar, kTile * 4 * (54 + (GW ? 48 : 0) + (GCAM ? 48 : 0) + (GUV ? 32 : 0)));
#endif
            bulk_g2s(s_ang, p.ang.p + row0 * 33, kTile * 33 * 4, s_bar);
            bulk_g2s(s_bone, p.bone.p + row0 * 15, kTile * 15 * 4, s_bar);
            bulk_g2s(s_grot, p.grot.p + row0 * 3, kTile * 3 * 4, s_bar);
            bulk_g2s(s_root, p.root.p + row0 * 3, kTile * 3 * 4, s_bar);
#if DHFK_PREFETCH_TILES_BWD > 0
            const long long rowp = row0 + (long long)DHFK_PREFETCH_TILES_BWD * kTile;
            if (rowp + kTile <= p.n) {
                bulk_prefetch_l2(p.ang.p + rowp * 33, kTile * 33 * 4);
                bulk_prefetch_l2(p.bone.p + rowp * 15, kTile * 15 * 4);
                bulk_prefetch_l2(p.grot.p + rowp * 3, kTile * 3 * 4);
                bulk_prefetch_l2(p.root.p + rowp * 3, kTile * 3 * 4);
                if (GW) bulk_prefetch_l2(p.g_world + rowp * 48, kTile * 48 * 4);
                if (GCAM) bulk_prefetch_l2(p.g_cam + rowp * 48, kTile * 48 * 4);
                if (GUV) bulk_prefetch_l2(p.g_uv + rowp * 32, kTile * 32 * 4);
            }
#endif
        }
#if DHFK_ROWS_IN == 1
        // padded gradient rows: coalesced 16-byte LDGSTS, completion counted on the same mbarrier
        if (GW) ldgsts_padded_rows<kWorldChunks>(s_gw, p.g_world, row0);
        if (GCAM) ldgsts_padded_rows<kWorldChunks>(s_gc, p.g_cam, row0);
        if (GUV) ldgsts_padded_rows<kUvChunks>(s_gu, p.g_uv, row0);
        ldgsts_arrive_noinc(s_bar);
#else
        // one row copy per thread into the padded rows (192 B / 128 B, 16-byte aligned both sides)
        if (GW) bulk_g2s(s_gw + tid * kWorldRow4, p.g_world + (row0 + tid) * 48, 48 * 4, s_bar);
        if (GCAM) bulk_g2s(s_gc + tid * kWorldRow4, p.g_cam + (row0 + tid) * 48, 48 * 4, s_bar);
        if (GUV) bulk_g2s(s_gu + tid * kUvRow4, p.g_uv + (row0 + tid) * 32, 32 * 4, s_bar);
#endif
#if DHFK_WAIT_ONE
        if (tid == 0) mbar_wait(s_bar, 0);
        __syncthreads();
#else
        mbar_wait(s_bar, 0);
#endif
    } else {
        stage_rows_in<33>(s_ang, p.ang, row0, rows);
        stage_rows_in<3>(s_grot, p.grot, row0, rows);
        stage_rows_in<15>(s_bone, p.bone, row0, rows);
        stage_rows_in<3>(s_root, p.root, row0, rows);
        if (GW) stage_padded_in<kWorldChunks>(s_gw, p.g_world, row0, rows);
        if (GCAM) stage_padded_in<kWorldChunks>(s_gc, p.g_cam, row0, rows);
        if (GUV) stage_padded_in<kUvChunks>(s_gu, p.g_uv, row0, rows);
        __syncthreads();
    }

    if (tid < rows) {
        BwdCtx<GUV, GBONE> ctx;
        ctx.ang = s_ang + tid * 33;
        ctx.g_ang = s_ang + tid * 33;
        ctx.bone = s_bone + tid * 15;
        ctx.g_bone = s_bone + tid * 15;
        ctx.cc = &p.cam;
        ctx.gw4 = GW ? s_gw + tid * kWorldRow4 : nullptr;
        ctx.gc4 = GCAM ? s_gc + tid * kWorldRow4 : nullptr;
        ctx.gu4 = GUV ? s_gu + tid * kUvRow4 : nullptr;
        float sx, cx, sy, cy;
        global_rotation<TRIG>(s_grot + tid * 3, ctx.R, sx, cx, sy, cy);
        ctx.root = v3(s_root[tid * 3], s_root[tid * 3 + 1], s_root[tid * 3 + 2]);
        const Frame I = identity_frame();
        Wrench wb = bwd_walk<TRIG, 10>(I, ctx);
        Wrench wr = bwd_walk<TRIG, 0>(I, ctx);
        Wrench wl = bwd_walk<TRIG, 5>(I, ctx);
        V3 Ft = wb.F + wr.F + wl.F;
        V3 Mt = wb.M + wr.M + wl.M;
        // d/d root = sum_k g_k = R * sum_k (R^T g_k)
        V3 gr = mat_vec(ctx.R, Ft);
        s_root[tid * 3] = gr.x; s_root[tid * 3 + 1] = gr.y; s_root[tid * 3 + 2] = gr.z;
        // d/d global angles: torque about the world axes e_x, Rx e_y, Rx Ry e_z
        V3 tw = mat_vec(ctx.R, Mt);
        s_grot[tid * 3] = kDegToRad * tw.x;
        s_grot[tid * 3 + 1] = kDegToRad * fmaf(cx, tw.y, sx * tw.z);
        s_grot[tid * 3 + 2] = kDegToRad * fmaf(sy, tw.x, fmaf(-sx * cy, tw.y, cx * cy * tw.z));
    }
    const bool bulk_out = rows == kTile && p.g_ang.vec && p.g_grot.vec && p.g_root.vec && (!GBONE || p.g_bone.vec);
    if (bulk_out) {
        // results overwrote the input slabs in place; ship the four slabs with one thread
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(p.g_ang.p + row0 * 33, s_ang, kTile * 33 * 4);
            bulk_s2g(p.g_grot.p + row0 * 3, s_grot, kTile * 3 * 4);
            bulk_s2g(p.g_root.p + row0 * 3, s_root, kTile * 3 * 4);
            if (GBONE) bulk_s2g(p.g_bone.p + row0 * 15, s_bone, kTile * 15 * 4);
            bulk_commit();
            bulk_wait_read_all();
        }
        return;
    }
    __syncthreads();

    stage_rows_out<33>(s_ang, p.g_ang, row0, rows);
    stage_rows_out<3>(s_grot, p.g_grot, row0, rows);
    stage_rows_out<3>(s_root, p.g_root, row0, rows);
    if (GBONE) stage_rows_out<15>(s_bone, p.g_bone, row0, rows);
}

}  // namespace dhfk

