// dhfk_video.cu -- SURVEY 8 (f2), video part: the inputs of the motion critics, fused.
//   Video_motion_Fk_3D_Discriminator.forward     models_Fk_GAN/Fk_discriminator.py:436-512
//       per-frame KCS-15 (video_mode_special_KCS_Input_transform, :269-377)           -> kcs  [B,F,15]
//       adjacent-frame KCS differences, a Python loop of F-1 slice writes + clones (:450-461) -> dkcs [B,F-1,15]
//       adjacent-frame 3-D differences, the same loop on the poses (:478-492)         -> dpos [B,F-1,48]
//   Video_motion_Fk_2D_Discriminator.forward     :554-587   root-joint 2-D differences  -> [B,F-1,2]
//   temporal playback reverse                    models_Fk_GAN/video_GAN_fun.py:222-223,269-270 (torch.flip(dims=[1]))
// One pass over the poses: one thread per frame computes that frame's KCS once, takes the next frame's from its
// neighbour lane (warp shuffle), and the differences leave through the same shared-memory staging as everything else.
// The reverse is index math on the way out: features of the reversed clip are the reversed features, and
//     d_rev[f] = x[F-2-f] - x[F-1-f] = -d[F-2-f],
// so the kernels always difference in storage order and write row (F-2-f) negated when DHFK_VIDEO_REVERSE is set.
//   forward : pose -> kcs, dkcs [, dpos] [, pos (the clip in playback order)]
//   backward: vector-Jacobian product of all of the above -> g_pose                (recomputes the bones)
//   jvp     : Jacobian-vector product (same shapes as forward) -- the derivative of `backward` w.r.t. its upstream
//             gradients, which WGAN-GP's create_graph=True pass differentiates through (Fk_discriminator.py:208-233)
#include "dhfk_launch.h"

namespace dhfk {

__device__ constexpr int VB0[15] = {5, 2, 4, 1, 0, 0, 0, 7, 8, 8, 10, 13, 11, 14, 8};     // bone = x[VB1] - x[VB0]
__device__ constexpr int VB1[15] = {6, 3, 5, 2, 4, 1, 7, 8, 10, 13, 11, 14, 12, 15, 9};   // (special_operate.py:515-531)
__device__ constexpr int VP0[15] = {0, 1, 2, 3, 4, 4, 5, 6, 7, 7, 7, 8, 9, 10, 11};       // feature = cos(bone VP0, bone VP1)
__device__ constexpr int VP1[15] = {2, 3, 4, 5, 5, 6, 6, 7, 14, 8, 9, 10, 11, 12, 13};    // (Fk_discriminator.py:324-372)

constexpr unsigned kVideoReverse = 1u;
constexpr int kOwn = kTile - 1;       // forward / jvp tiles advance 31 rows: lane 31 is the halo frame of lane 30

struct VideoParams {
    const float* pose;    // [B*F,16,3]
    const float* v;       // jvp: tangent [B*F,16,3]
    const float* g_kcs;   // backward: [B,F,15] or null
    const float* g_dkcs;  // backward: [B,F-1,15] or null
    const float* g_dpos;  // backward: [B,F-1,48] or null
    const float* g_pos;   // backward: [B,F,48] or null
    float* out_kcs;       // forward / jvp: [B,F,15]
    float* out_dkcs;      // forward / jvp: [B,F-1,15]
    float* out_dpos;      // forward / jvp: [B,F-1,48] or null
    float* out_pos;       // forward / jvp: [B,F,48] (playback order) or null;  backward: g_pose [B*F,16,3]
    long long n;          // B*F
    int frames;           // F
    unsigned flags;
};

DHFK_DI void vload48(const float4* row, float* x) {
#pragma unroll
    for (int c = 0; c < 12; ++c) {
        const float4 v = row[c];
        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
    }
}
DHFK_DI void vstore48(float4* row, const float* y) {
#pragma unroll
    for (int c = 0; c < 12; ++c) row[c] = make_float4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
}
DHFK_DI float vdot(const float* a, const float* b) { return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])); }

struct VBones {
    float v[45];
    float inv[15];   // 1 / length
};
DHFK_DI void vbones(const float* x, VBones& B) {
#pragma unroll
    for (int b = 0; b < 15; ++b) {
        const float dx = x[3 * VB1[b]] - x[3 * VB0[b]], dy = x[3 * VB1[b] + 1] - x[3 * VB0[b] + 1],
                    dz = x[3 * VB1[b] + 2] - x[3 * VB0[b] + 2];
        B.v[3 * b] = dx; B.v[3 * b + 1] = dy; B.v[3 * b + 2] = dz;
        B.inv[b] = rsqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    }
}
DHFK_DI void vkcs(const VBones& B, float* k) {
#pragma unroll
    for (int q = 0; q < 15; ++q) k[q] = vdot(B.v + 3 * VP0[q], B.v + 3 * VP1[q]) * (B.inv[VP0[q]] * B.inv[VP1[q]]);
}
// d kcs along the tangent tx of the pose
DHFK_DI void vkcs_jvp(const VBones& B, const float* tx, float* tk) {
    float tb[45], tl[15];
#pragma unroll
    for (int b = 0; b < 15; ++b) {
#pragma unroll
        for (int a = 0; a < 3; ++a) tb[3 * b + a] = tx[3 * VB1[b] + a] - tx[3 * VB0[b] + a];
        tl[b] = vdot(B.v + 3 * b, tb + 3 * b) * (B.inv[b] * B.inv[b]);
    }
#pragma unroll
    for (int q = 0; q < 15; ++q) {
        const int i = VP0[q], j = VP1[q];
        const float ij = B.inv[i] * B.inv[j];
        const float c = vdot(B.v + 3 * i, B.v + 3 * j) * ij;
        const float d = (vdot(tb + 3 * i, B.v + 3 * j) + vdot(B.v + 3 * i, tb + 3 * j)) * ij;
        tk[q] = fmaf(-c, tl[i] + tl[j], d);
    }
}
// g_x += J_kcs(x)^T gk
DHFK_DI void vkcs_vjp(const VBones& B, const float* gk, float* gx) {
    float gb[45];
#pragma unroll
    for (int i = 0; i < 45; ++i) gb[i] = 0.f;
#pragma unroll
    for (int q = 0; q < 15; ++q) {
        const int i = VP0[q], j = VP1[q];
        const float ii = B.inv[i], ij = B.inv[j];
        const float c = vdot(B.v + 3 * i, B.v + 3 * j) * (ii * ij);
        const float w = gk[q] * (ii * ij), ci = gk[q] * c * (ii * ii), cj = gk[q] * c * (ij * ij);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            gb[3 * i + a] += fmaf(w, B.v[3 * j + a], -ci * B.v[3 * i + a]);
            gb[3 * j + a] += fmaf(w, B.v[3 * i + a], -cj * B.v[3 * j + a]);
        }
    }
#pragma unroll
    for (int b = 0; b < 15; ++b)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            gx[3 * VB1[b] + a] += gb[3 * b + a];
            gx[3 * VB0[b] + a] -= gb[3 * b + a];
        }
}

// Output rows of storage row r = (clip b, frame f): per-frame tensors [B,F,*] and difference tensors [B,F-1,*]
// (difference f = frame f+1 minus frame f, defined for f < F-1; -1 otherwise).
// (n_rows < 2^31 is checked at the C ABI: 32-bit arithmetic; a 64-bit division is ~60 emulated instructions)
struct VideoRows { int per_frame, diff; };
DHFK_DI VideoRows video_rows(unsigned r, unsigned F, bool rev) {
    const unsigned b = r / F;
    const unsigned f = r - b * F;
    VideoRows o;
    o.per_frame = (int)(b * F + (rev ? F - 1 - f : f));
    o.diff = f < F - 1 ? (int)(b * (F - 1) + (rev ? F - 2 - f : f)) : -1;
    return o;
}

// rows of 15 floats from a compact shared image (slot * 15) to mapped global rows (-1 = skip): each half-warp ships
// one row per pass (lanes 0..14 of the half: 60 contiguous bytes), so a pass costs one LDS, one broadcast LDS of the
// row index and one STG -- no per-element division, no 64-bit multiply
DHFK_DI void store_rows15_mapped(const float* s, float* g, const int* map, int slots) {
    const int lane = threadIdx.x, half = lane >> 4, c = lane & 15;
#pragma unroll 4
    for (int slot = half; slot < slots; slot += 2) {
        const int row = map[slot];
        if (c < 15 && row >= 0) __stcs(g + (size_t)row * 15 + c, s[slot * 15 + c]);
    }
}
// 48-float rows from padded shared rows to mapped global rows: 16-byte stores; 8 rows of 12 chunks per 3 passes
DHFK_DI void store_rows48_mapped(const float4* s4, float* g, const int* map, int slots) {
    float4* g4 = reinterpret_cast<float4*>(g);
    for (int i = threadIdx.x; i < slots * kWorldChunks; i += kTile) {
        const int slot = i / kWorldChunks, c = i - slot * kWorldChunks;
        const int row = map[slot];
        if (row >= 0) __stcs(g4 + (size_t)row * kWorldChunks + c, s4[slot * kWorldRow4 + c]);
    }
}

// adjacent-frame differences of 48-float rows, formed in the store loop straight from the staged rows (slot+1 minus slot,
// times sgn): no difference buffer in shared memory, so more CTAs fit an SM
DHFK_DI void store_rowdiff48_mapped(const float4* s4, float* g, const int* map, int slots, float sgn) {
    float4* g4 = reinterpret_cast<float4*>(g);
    for (int i = threadIdx.x; i < slots * kWorldChunks; i += kTile) {
        const int slot = i / kWorldChunks, c = i - slot * kWorldChunks;
        const int row = map[slot];
        if (row >= 0) {
            const float4 a = s4[slot * kWorldRow4 + c], b = s4[(slot + 1) * kWorldRow4 + c];
            __stcs(g4 + (size_t)row * kWorldChunks + c,
                   make_float4(sgn * (b.x - a.x), sgn * (b.y - a.y), sgn * (b.z - a.z), sgn * (b.w - a.w)));
        }
    }
}

// MODE 0: forward.  MODE 2: jvp (tangent p.v through the same maps).
template <int MODE, bool DPOS, bool POS>
__global__ void __launch_bounds__(kTile) dhfk_video_critic_kernel(const __grid_constant__ VideoParams p) {
    constexpr bool JVP = MODE == 2;
    extern __shared__ __align__(16) float smem[];
    float4* s_pose = reinterpret_cast<float4*>(smem);                       // 32 padded rows
    float4* s_v = s_pose + kTile * kWorldRow4;                              // jvp: tangents
    float* s_k = reinterpret_cast<float*>(s_v + (JVP ? kTile * kWorldRow4 : 0));
    float* s_dk = s_k + kTile * 15;
    int* s_map_f = reinterpret_cast<int*>(s_dk + kTile * 15);              // per-frame output row of each slot
    int* s_map_d = s_map_f + kTile;                                         // difference output row (or -1)

    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kOwn;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;        // staged rows (own + halo)
    const int own = rows < kOwn ? rows : kOwn;                // rows this tile emits
    if (rows == kTile) {
        ldgsts_padded_tile<kWorldChunks>(s_pose, p.pose, row0);
        if (JVP) ldgsts_padded_tile<kWorldChunks>(s_v, p.v, row0);
    } else {
        stage_padded_in<kWorldChunks>(s_pose, p.pose, row0, rows);
        if (JVP) stage_padded_in<kWorldChunks>(s_v, p.v, row0, rows);
    }
    const bool rev = (p.flags & kVideoReverse) != 0;
    VideoRows o;
    o.per_frame = o.diff = -1;
    if (lane < own) o = video_rows((unsigned)(row0 + lane), (unsigned)p.frames, rev);
    s_map_f[lane] = o.per_frame;
    s_map_d[lane] = o.diff;
    ldgsts_wait_all();
    __syncwarp();

    float k[15];
#pragma unroll
    for (int q = 0; q < 15; ++q) k[q] = 0.f;
    float x[48];
    if (lane < rows) {
        vload48(s_pose + lane * kWorldRow4, x);
        VBones B;
        vbones(x, B);
        if (JVP) {
            vload48(s_v + lane * kWorldRow4, x);      // from here on x is the tangent: positions enter linearly
            vkcs_jvp(B, x, k);
        } else {
            vkcs(B, k);
        }
    }
    // next frame's features / positions: neighbour lane (the halo lane 31 feeds lane 30)
    const float sgn = rev ? -1.f : 1.f;
#pragma unroll
    for (int q = 0; q < 15; ++q) {
        const float kn = __shfl_down_sync(0xffffffffu, k[q], 1);
        s_k[lane * 15 + q] = k[q];
        s_dk[lane * 15 + q] = sgn * (kn - k[q]);
    }
    __syncwarp();
    store_rows15_mapped(s_k, p.out_kcs, s_map_f, own);
    store_rows15_mapped(s_dk, p.out_dkcs, s_map_d, own);
    if (DPOS) store_rowdiff48_mapped(JVP ? s_v : s_pose, p.out_dpos, s_map_d, own, sgn);
    if (POS) store_rows48_mapped(JVP ? s_v : s_pose, p.out_pos, s_map_f, own);
}

// backward: g_pose[r] = J_kcs(x_r)^T ( g_kcs[pf(r)] + s g_dkcs[d(r-1)] - s g_dkcs[d(r)] )
//                       + g_pos[pf(r)] + s g_dpos[d(r-1)] - s g_dpos[d(r)]
// with pf / d the output rows of video_rows(), s = -1 in reverse mode, terms outside the clip dropped.
// ONE 33-row tile of shared memory is used twice: first for the pose rows (each lane copies its row into registers),
// then -- while the lanes do the KCS part on those registers -- cp.async refills it with the 33 g_dpos rows the tile
// needs (difference rows r-1 .. r+31, gathered through the slot -> row table), and once more with the g_pos rows when
// that gradient exists.  10.9 kB per CTA; the loads stay coalesced 16-byte cp.async and overlap the arithmetic.
// (Staging everything at once: 17.7 kB, 12 CTAs per SM, 0.75 of the copy peak; adding the positional part from global
// memory in the store loop: dependent loads, 0.67; per-lane 128-bit global loads: 32 sectors per instruction, 0.67.)
template <bool GK, bool GDK, bool GDP, bool GP>
__global__ void __launch_bounds__(kTile) dhfk_video_critic_bwd_kernel(const __grid_constant__ VideoParams p) {
    extern __shared__ __align__(16) float smem[];
    float4* s_tile = reinterpret_cast<float4*>(smem);                        // 33 padded rows, reused (see above)
    float* s_gk = reinterpret_cast<float*>(s_tile + (kTile + 1) * kWorldRow4);   // g_kcs, 32 x 15 (odd stride)
    float* s_gdk = s_gk + (GK ? kTile * 15 : 0);                             // g_dkcs, 33 x 15
    int* s_map_f = reinterpret_cast<int*>(s_gdk + (GDK ? (kTile + 1) * 15 : 0));
    int* s_map_d = s_map_f + kTile;                                          // 33 entries

    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    const bool rev = (p.flags & kVideoReverse) != 0;
    constexpr bool KCS = GK || GDK;
    VideoRows o;
    o.per_frame = o.diff = -1;
    if (lane < rows) o = video_rows((unsigned)(row0 + lane), (unsigned)p.frames, rev);
    int prev_diff = __shfl_up_sync(0xffffffffu, o.diff, 1);
    if (lane == 0) prev_diff = row0 > 0 ? video_rows((unsigned)(row0 - 1), (unsigned)p.frames, rev).diff : -1;
    s_map_f[lane] = o.per_frame;
    s_map_d[lane + 1] = o.diff;
    if (lane == 0) s_map_d[0] = prev_diff;
    if (KCS) {
        if (rows == kTile) ldgsts_padded_tile<kWorldChunks>(s_tile, p.pose, row0);
        else stage_padded_in<kWorldChunks>(s_tile, p.pose, row0, rows);
    }
    __syncwarp();
    // mapped gathers of the 15-float upstream gradients (cp.async, one row per half-warp and pass)
    if (GK) {
        const int half = lane >> 4, c = lane & 15;
        for (int slot = half; slot < rows; slot += 2)
            if (c < 15) ldgsts4(s_gk + slot * 15 + c, p.g_kcs + (size_t)s_map_f[slot] * 15 + c);
    }
    if (GDK) {
        const int half = lane >> 4, c = lane & 15;
        for (int slot = half; slot < rows + 1; slot += 2) {
            const int row = s_map_d[slot];
            if (c < 15 && row >= 0) ldgsts4(s_gdk + slot * 15 + c, p.g_dkcs + (size_t)row * 15 + c);
        }
    }
    ldgsts_wait_all();
    __syncwarp();
    const float sgn = rev ? -1.f : 1.f;
    const bool has_prev = prev_diff >= 0, has_next = o.diff >= 0;
    float x[48], gk[15];
    if (KCS && lane < rows) {
        vload48(s_tile + lane * kWorldRow4, x);
#pragma unroll
        for (int q = 0; q < 15; ++q) {
            float a = GK ? s_gk[lane * 15 + q] : 0.f;
            if (GDK) {
                if (has_prev) a = fmaf(sgn, s_gdk[lane * 15 + q], a);
                if (has_next) a = fmaf(-sgn, s_gdk[(lane + 1) * 15 + q], a);
            }
            gk[q] = a;
        }
    }
    __syncwarp();        // every lane holds its pose row: the tile can be refilled
    if (GDP) {           // difference rows r-1 .. r+31 of g_dpos -> slots 0 .. 32
        const float4* g4 = reinterpret_cast<const float4*>(p.g_dpos);
        for (int i = lane; i < (rows + 1) * kWorldChunks; i += kTile) {
            const int slot = i / kWorldChunks, c = i - slot * kWorldChunks;
            const int row = s_map_d[slot];
            if (row >= 0) ldgsts16(s_tile + slot * kWorldRow4 + c, g4 + (size_t)row * kWorldChunks + c);
        }
    }
    float g[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) g[i] = 0.f;
    if (KCS && lane < rows) {        // the KCS part, on registers, while the refill is in flight
        VBones B;
        vbones(x, B);
        vkcs_vjp(B, gk, g);
    }
    if (GDP) {
        ldgsts_wait_all();
        __syncwarp();
        if (lane < rows) {
            float t[48];
            if (has_prev) {
                vload48(s_tile + lane * kWorldRow4, t);
#pragma unroll
                for (int i = 0; i < 48; ++i) g[i] = fmaf(sgn, t[i], g[i]);
            }
            if (has_next) {
                vload48(s_tile + (lane + 1) * kWorldRow4, t);
#pragma unroll
                for (int i = 0; i < 48; ++i) g[i] = fmaf(-sgn, t[i], g[i]);
            }
        }
        __syncwarp();
    }
    if (GP) {            // g_pos rows (per-frame output rows of this tile) -> slots 0 .. 31
        const float4* g4 = reinterpret_cast<const float4*>(p.g_pos);
        for (int i = lane; i < rows * kWorldChunks; i += kTile) {
            const int slot = i / kWorldChunks, c = i - slot * kWorldChunks;
            ldgsts16(s_tile + slot * kWorldRow4 + c, g4 + (size_t)s_map_f[slot] * kWorldChunks + c);
        }
        ldgsts_wait_all();
        __syncwarp();
        if (lane < rows) {
            float t[48];
            vload48(s_tile + lane * kWorldRow4, t);
#pragma unroll
            for (int i = 0; i < 48; ++i) g[i] += t[i];
        }
        __syncwarp();
    }
    if (lane < rows) vstore48(s_tile + lane * kWorldRow4, g);
    __syncwarp();
    if (rows == kTile) store_padded_tile<kWorldChunks>(s_tile, p.out_pos, row0);
    else stage_padded_out<kWorldChunks>(s_tile, p.out_pos, row0, rows);
}

// ---- 2-D motion critic: root-joint differences (Fk_discriminator.py:566-579) ------------------------------------
// forward : uv [B*F,16,2] -> diff [B,F-1,2] = uv[b,f+1,0,:] - uv[b,f,0,:]  [, uv in playback order]
// backward: the transpose: g_uv [B*F,16,2] = scatter of +-g_diff into joint 0 [+ g_uv_pb gathered back]
// One thread per frame; the forward touches 8 bytes of every 128-byte row (one sector), the backward writes full rows.
struct RootDiffParams {
    const float* uv;       // forward: [B*F,16,2]
    const float* g_diff;   // backward: [B,F-1,2] or null
    const float* g_pb;     // backward: gradient of the playback-order copy [B,F,32] or null
    float* out_diff;       // forward
    float* out_pb;         // forward: playback-order copy or null
    float* g_uv;           // backward: [B*F,16,2]
    long long n;
    int frames;
    unsigned flags;
};
__global__ void __launch_bounds__(256) dhfk_video_root_diff_fwd_kernel(const __grid_constant__ RootDiffParams p) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n) return;
    const bool rev = (p.flags & kVideoReverse) != 0;
    const VideoRows o = video_rows((unsigned)r, (unsigned)p.frames, rev);
    const float2* u2 = reinterpret_cast<const float2*>(p.uv);
    if (o.diff >= 0) {
        const float2 a = __ldg(u2 + r * 16), b = __ldg(u2 + (r + 1) * 16);
        const float s = rev ? -1.f : 1.f;
        reinterpret_cast<float2*>(p.out_diff)[o.diff] = make_float2(s * (b.x - a.x), s * (b.y - a.y));
    }
    if (p.out_pb) {
        const float4* s4 = reinterpret_cast<const float4*>(p.uv) + r * kUvChunks;
        float4* d4 = reinterpret_cast<float4*>(p.out_pb) + o.per_frame * kUvChunks;
#pragma unroll
        for (int c = 0; c < kUvChunks; ++c) d4[c] = __ldg(s4 + c);
    }
}
__global__ void __launch_bounds__(256) dhfk_video_root_diff_bwd_kernel(const __grid_constant__ RootDiffParams p) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n) return;
    const bool rev = (p.flags & kVideoReverse) != 0;
    const VideoRows o = video_rows((unsigned)r, (unsigned)p.frames, rev);
    const unsigned b = (unsigned)r / (unsigned)p.frames;
    const int f = (int)((unsigned)r - b * (unsigned)p.frames);
    float gx = 0.f, gy = 0.f;
    if (p.g_diff) {
        const float2* g2 = reinterpret_cast<const float2*>(p.g_diff);
        const float s = rev ? -1.f : 1.f;
        if (f > 0) {                      // difference f-1 = frame f - frame f-1
            const float2 g = __ldg(g2 + video_rows((unsigned)(r - 1), (unsigned)p.frames, rev).diff);
            gx += s * g.x; gy += s * g.y;
        }
        if (o.diff >= 0) {
            const float2 g = __ldg(g2 + o.diff);
            gx -= s * g.x; gy -= s * g.y;
        }
    }
    float4* d4 = reinterpret_cast<float4*>(p.g_uv) + r * kUvChunks;
    const float4* s4 = p.g_pb ? reinterpret_cast<const float4*>(p.g_pb) + o.per_frame * kUvChunks : nullptr;
#pragma unroll
    for (int c = 0; c < kUvChunks; ++c) {
        float4 v = s4 ? __ldg(s4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (c == 0) { v.x += gx; v.y += gy; }
        d4[c] = v;
    }
}

static size_t video_fwd_smem(bool jvp, bool dpos) {
    (void)dpos;
    return sizeof(float4) * kTile * kWorldRow4 * (1 + (jvp ? 1 : 0)) + sizeof(float) * kTile * 30 + sizeof(int) * kTile * 2;
}

// launch_tiles sizes the grid as ceil(p.n / 32); forward / jvp tiles advance 31 rows, so they get their own launcher
template <typename K>
static int launch_video(K kernel, size_t smem, const VideoParams& p, long long blocks, cudaStream_t st, const char** where) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { *where = "cudaGetDevice"; return (int)e; }
    if (!func_attrs_done((const void*)kernel, dev)) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) {
            func_attrs_forget((const void*)kernel, dev);
            *where = "cudaFuncSetAttribute(video critic kernel)";
            return (int)e;
        }
    }
    void* args[] = {const_cast<VideoParams*>(&p)};
    e = cudaLaunchKernel((const void*)kernel, dim3((unsigned)blocks), dim3(kTile), args, smem, st);
    if (e != cudaSuccess) { *where = "cudaLaunchKernel(video critic kernel)"; return (int)e; }
    return 0;
}

// mode 0 forward, 2 jvp (v = tangent)
int launch_video_critic(int mode, const float* pose, const float* v, int frames, unsigned flags, float* out_kcs,
                        float* out_dkcs, float* out_dpos, float* out_pos, long long n, cudaStream_t st, const char** where) {
    VideoParams p = {};
    p.pose = pose; p.v = v; p.out_kcs = out_kcs; p.out_dkcs = out_dkcs; p.out_dpos = out_dpos; p.out_pos = out_pos;
    p.n = n; p.frames = frames; p.flags = flags;
    const long long blocks = (n + kOwn - 1) / kOwn;
    const bool dp = out_dpos != nullptr, ps = out_pos != nullptr;
#define DHFK_VIDEO(M)                                                                                                \
    if (dp && ps) return launch_video(dhfk_video_critic_kernel<M, true, true>, video_fwd_smem(M == 2, true), p, blocks, st, where);   \
    if (dp) return launch_video(dhfk_video_critic_kernel<M, true, false>, video_fwd_smem(M == 2, true), p, blocks, st, where);        \
    if (ps) return launch_video(dhfk_video_critic_kernel<M, false, true>, video_fwd_smem(M == 2, false), p, blocks, st, where);       \
    return launch_video(dhfk_video_critic_kernel<M, false, false>, video_fwd_smem(M == 2, false), p, blocks, st, where);
    if (mode == 0) { DHFK_VIDEO(0) }
    DHFK_VIDEO(2)
#undef DHFK_VIDEO
}

int launch_video_critic_bwd(const float* pose, int frames, unsigned flags, const float* g_kcs, const float* g_dkcs,
                            const float* g_dpos, const float* g_pos, float* g_pose, long long n, cudaStream_t st,
                            const char** where) {
    VideoParams p = {};
    p.pose = pose; p.g_kcs = g_kcs; p.g_dkcs = g_dkcs; p.g_dpos = g_dpos; p.g_pos = g_pos; p.out_pos = g_pose;
    p.n = n; p.frames = frames; p.flags = flags;
    const bool gk = g_kcs != nullptr, gdk = g_dkcs != nullptr, gdp = g_dpos != nullptr, gp = g_pos != nullptr;
    const size_t smem = sizeof(float4) * (kTile + 1) * kWorldRow4 +
                        sizeof(float) * ((gk ? kTile * 15 : 0) + (gdk ? (kTile + 1) * 15 : 0)) + sizeof(int) * (2 * kTile + 2);
    const long long blocks = (n + kTile - 1) / kTile;
    // the combinations the critics produce: everything (3-D motion critic with both extra branches), features only,
    // and the two single-branch configurations (motion_Dis_whether_use_3dPos_branch / _3dDiff_branch)
#define DHFK_VB(A, B, C, D) return launch_video(dhfk_video_critic_bwd_kernel<A, B, C, D>, smem, p, blocks, st, where)
    const int key = (gk ? 8 : 0) | (gdk ? 4 : 0) | (gdp ? 2 : 0) | (gp ? 1 : 0);
    switch (key) {
        case 15: DHFK_VB(true, true, true, true);
        case 14: DHFK_VB(true, true, true, false);
        case 13: DHFK_VB(true, true, false, true);
        case 12: DHFK_VB(true, true, false, false);
        case 8: DHFK_VB(true, false, false, false);
        case 4: DHFK_VB(false, true, false, false);
        case 2: DHFK_VB(false, false, true, false);
        case 1: DHFK_VB(false, false, false, true);
        case 3: DHFK_VB(false, false, true, true);
        case 10: DHFK_VB(true, false, true, false);
        case 9: DHFK_VB(true, false, false, true);
        case 11: DHFK_VB(true, false, true, true);
        case 6: DHFK_VB(false, true, true, false);
        case 5: DHFK_VB(false, true, false, true);
        case 7: DHFK_VB(false, true, true, true);
        default: break;
    }
#undef DHFK_VB
    *where = "video critic backward: no upstream gradient";
    return (int)cudaErrorInvalidValue;
}

int launch_video_root_diff(bool bwd, const float* uv, const float* g_diff, const float* g_pb, int frames, unsigned flags,
                           float* out_diff, float* out_pb, float* g_uv, long long n, cudaStream_t st, const char** where) {
    RootDiffParams p = {};
    p.uv = uv; p.g_diff = g_diff; p.g_pb = g_pb; p.out_diff = out_diff; p.out_pb = out_pb; p.g_uv = g_uv;
    p.n = n; p.frames = frames; p.flags = flags;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (bwd) dhfk_video_root_diff_bwd_kernel<<<blocks, 256, 0, st>>>(p);
    else dhfk_video_root_diff_fwd_kernel<<<blocks, 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *where = "dhfk_video_root_diff kernel"; return (int)e; }
    return 0;
}

}  // namespace dhfk
