// dhfk_critic.cu -- SURVEY 8 (f2): critic input transforms, fused.
//   special_KCS_Input_transform / video_mode_special_KCS_Input_transform
//       models_Fk_GAN/Fk_discriminator.py:36-146, :269-377   (15 bone-pair cosines [+ 15 bone lengths])
//   Fk_get_boneVecByPose3d            models_Fk_GAN/special_operate.py:513-539  (bone = child - parent)
//   root-centring                     models_Fk_GAN/model_fk_gan_train.py:295,312,437
//   left/right flip                   models_Fk_GAN/model_fk_gan_train.py:320-331,393-405,449-461
// The reference runs two [N,3,16]x[N,16,15] matmuls, 30 row writes into a [30,N] buffer and a transpose
// (~120 autograd nodes) per critic evaluation, 4-8 evaluations per iteration.  Here: one thread per pose,
// one warp per 32-pose tile, the same LDGSTS-in / coalesced-out staging as the FK kernels.
//   forward : pose -> (flip) -> (centre) -> pos', KCS features
//   backward: vector-Jacobian product  d<g_pos,pos'> + <g_kcs,kcs> / d pose        (recomputes the bones)
//   jvp     : Jacobian-vector product, i.e. the derivative of `backward` w.r.t. its upstream gradients --
//             what WGAN-GP's create_graph=True pass needs (Fk_discriminator.py:224-231)
#include "dhfk_launch.h"

namespace dhfk {

// bone b = x[KB1[b]] - x[KB0[b]], used_16key_15bone_len_table order (special_operate.py:515-531)
__device__ constexpr int KB0[15] = {5, 2, 4, 1, 0, 0, 0, 7, 8, 8, 10, 13, 11, 14, 8};
__device__ constexpr int KB1[15] = {6, 3, 5, 2, 4, 1, 7, 8, 10, 13, 11, 14, 12, 15, 9};
// feature p = cos(bone KP0[p], bone KP1[p])  (Fk_discriminator.py:81-139)
__device__ constexpr int KP0[15] = {0, 1, 2, 3, 4, 4, 5, 6, 7, 7, 7, 8, 9, 10, 11};
__device__ constexpr int KP1[15] = {2, 3, 4, 5, 5, 6, 6, 7, 14, 8, 9, 10, 11, 12, 13};
// joint j of the flipped pose = mirrored joint FLIP16[j] (out_left/out_right swap, model_fk_gan_train.py:321-327)
__device__ constexpr int FLIP16[16] = {0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 13, 14, 15, 10, 11, 12};

constexpr unsigned kCentre = 1u, kFlip = 2u;

struct CriticParams {
    const float* pose;   // [N,16,3]
    const float* a;      // backward: g_pos [N,16,3] or null;  jvp: v_pose [N,16,3]
    const float* b;      // backward: g_kcs [N,KC] or null
    float* out_pos;      // forward: pos' ; backward: g_pose ; jvp: t_pos   (null = not wanted, fwd/jvp only)
    float* out_kcs;      // forward: kcs [N,KC] ; jvp: t_kcs                (null = not wanted)
    long long n;
    unsigned flags;
};

DHFK_DI void load_row48(const float4* row, float* x) {
#pragma unroll
    for (int c = 0; c < 12; ++c) {
        float4 v = row[c];
        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
    }
}
DHFK_DI void store_row48(float4* row, const float* y) {
#pragma unroll
    for (int c = 0; c < 12; ++c) row[c] = make_float4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
}
// y = centre(flip(x)); linear, so the same routine maps tangents
DHFK_DI void flip_centre(const float* x, float* y, unsigned flags) {
    const bool f = flags & kFlip;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int s = FLIP16[j];
        y[3 * j] = f ? -x[3 * s] : x[3 * j];
        y[3 * j + 1] = f ? x[3 * s + 1] : x[3 * j + 1];
        y[3 * j + 2] = f ? x[3 * s + 2] : x[3 * j + 2];
    }
    if (flags & kCentre) {
        const float rx = y[0], ry = y[1], rz = y[2];
#pragma unroll
        for (int j = 0; j < 16; ++j) { y[3 * j] -= rx; y[3 * j + 1] -= ry; y[3 * j + 2] -= rz; }
    }
}
// transpose of flip_centre: g_x from g_y (in place safe: uses a copy of row 0 sums first)
DHFK_DI void flip_centre_T(float* g, float* gx, unsigned flags) {
    if (flags & kCentre) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { sx += g[3 * j]; sy += g[3 * j + 1]; sz += g[3 * j + 2]; }
        g[0] -= sx; g[1] -= sy; g[2] -= sz;
    }
    const bool f = flags & kFlip;
#pragma unroll
    for (int j = 0; j < 16; ++j) {   // FLIP16 is an involution: gx[j] = mirror(g[FLIP16[j]])
        const int s = FLIP16[j];
        gx[3 * j] = f ? -g[3 * s] : g[3 * j];
        gx[3 * j + 1] = f ? g[3 * s + 1] : g[3 * j + 1];
        gx[3 * j + 2] = f ? g[3 * s + 2] : g[3 * j + 2];
    }
}
struct Bones {
    float v[45];    // bone vectors
    float s[15];    // squared lengths
    float inv[15];  // 1 / length
};
DHFK_DI void bones_of(const float* y, Bones& B) {
#pragma unroll
    for (int b = 0; b < 15; ++b) {
        const float dx = y[3 * KB1[b]] - y[3 * KB0[b]], dy = y[3 * KB1[b] + 1] - y[3 * KB0[b] + 1],
                    dz = y[3 * KB1[b] + 2] - y[3 * KB0[b] + 2];
        B.v[3 * b] = dx; B.v[3 * b + 1] = dy; B.v[3 * b + 2] = dz;
        B.s[b] = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        B.inv[b] = rsqrtf(B.s[b]);
    }
}
DHFK_DI float bdot(const float* a, const float* b) { return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])); }

template <int KC, bool POS>
__global__ void __launch_bounds__(kTile) dhfk_critic_fwd_kernel(const __grid_constant__ CriticParams p) {
    extern __shared__ __align__(16) float smem[];
    float4* s_pose = reinterpret_cast<float4*>(smem);
    float* s_kcs = reinterpret_cast<float*>(s_pose + kTile * kWorldRow4);
    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    if (rows == kTile) {
        ldgsts_padded_tile<kWorldChunks>(s_pose, p.pose, row0);
    } else {
        stage_padded_in<kWorldChunks>(s_pose, p.pose, row0, rows);
    }
    ldgsts_wait_all();     // full tiles and ragged tiles alike: everything above was queued with cp.async
    __syncwarp();
    if (lane < rows) {
        float x[48], y[48];
        load_row48(s_pose + lane * kWorldRow4, x);
        flip_centre(x, y, p.flags);
        if (POS) store_row48(s_pose + lane * kWorldRow4, y);
        if (KC > 0) {
            Bones B;
            bones_of(y, B);
            float* k = s_kcs + lane * KC;
#pragma unroll
            for (int q = 0; q < 15; ++q)
                k[q] = bdot(B.v + 3 * KP0[q], B.v + 3 * KP1[q]) * (B.inv[KP0[q]] * B.inv[KP1[q]]);
            if (KC == 30) {
#pragma unroll
                for (int b = 0; b < 15; ++b) k[15 + b] = sqrtf(B.s[b]);
            }
        }
    }
    __syncwarp();
    if (POS) {
        if (rows == kTile) store_padded_tile<kWorldChunks>(s_pose, p.out_pos, row0);
        else stage_padded_out<kWorldChunks>(s_pose, p.out_pos, row0, rows);
    }
    if (KC > 0) {
        RowDst d; d.p = p.out_kcs; d.stride = KC; d.vec = 1;
        stage_rows_out<(KC > 0 ? KC : 1)>(s_kcs, d, row0, rows);
    }
}

// gradient of the KCS features w.r.t. the bone vectors, contracted with gk (VJP) -> gb[45]
template <int KC>
DHFK_DI void kcs_vjp(const Bones& B, const float* gk, float* gb) {
#pragma unroll
    for (int i = 0; i < 45; ++i) gb[i] = 0.f;
#pragma unroll
    for (int q = 0; q < 15; ++q) {
        const int i = KP0[q], j = KP1[q];
        const float ii = B.inv[i], ij = B.inv[j];
        const float c = bdot(B.v + 3 * i, B.v + 3 * j) * (ii * ij);
        const float w = gk[q] * (ii * ij), ci = gk[q] * c * (ii * ii), cj = gk[q] * c * (ij * ij);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            gb[3 * i + a] += fmaf(w, B.v[3 * j + a], -ci * B.v[3 * i + a]);
            gb[3 * j + a] += fmaf(w, B.v[3 * i + a], -cj * B.v[3 * j + a]);
        }
    }
    if (KC == 30) {
#pragma unroll
        for (int b = 0; b < 15; ++b) {
            const float w = gk[15 + b] * B.inv[b];
#pragma unroll
            for (int a = 0; a < 3; ++a) gb[3 * b + a] = fmaf(w, B.v[3 * b + a], gb[3 * b + a]);
        }
    }
}

template <int KC, bool GPOS>
__global__ void __launch_bounds__(kTile) dhfk_critic_bwd_kernel(const __grid_constant__ CriticParams p) {
    extern __shared__ __align__(16) float smem[];
    float4* s_pose = reinterpret_cast<float4*>(smem);
    float4* s_gp = s_pose + kTile * kWorldRow4;
    float* s_gk = reinterpret_cast<float*>(s_gp + (GPOS ? kTile * kWorldRow4 : 0));
    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    if (rows == kTile) {
        ldgsts_padded_tile<kWorldChunks>(s_pose, p.pose, row0);
        if (GPOS) ldgsts_padded_tile<kWorldChunks>(s_gp, p.a, row0);
        if (KC > 0) ldgsts_slab<(KC > 0 ? KC : 4)>(s_gk, p.b + row0 * KC);
    } else {
        stage_padded_in<kWorldChunks>(s_pose, p.pose, row0, rows);
        if (GPOS) stage_padded_in<kWorldChunks>(s_gp, p.a, row0, rows);
        if (KC > 0) {
            RowSrc s; s.p = p.b; s.stride = KC; s.vec = 1;
            stage_rows_in<(KC > 0 ? KC : 1)>(s_gk, s, row0, rows);
        }
    }
    ldgsts_wait_all();     // full tiles and ragged tiles alike: everything above was queued with cp.async
    __syncwarp();
    if (lane < rows) {
        float x[48], y[48], g[48];
        if (GPOS) load_row48(s_gp + lane * kWorldRow4, g);
        else {
#pragma unroll
            for (int i = 0; i < 48; ++i) g[i] = 0.f;
        }
        if (KC > 0) {
            load_row48(s_pose + lane * kWorldRow4, x);
            flip_centre(x, y, p.flags);
            Bones B;
            bones_of(y, B);
            float gk[KC > 0 ? KC : 1], gb[45];
#pragma unroll
            for (int q = 0; q < KC; ++q) gk[q] = s_gk[lane * KC + q];
            kcs_vjp<KC>(B, gk, gb);
#pragma unroll
            for (int b = 0; b < 15; ++b)
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    g[3 * KB1[b] + a] += gb[3 * b + a];
                    g[3 * KB0[b] + a] -= gb[3 * b + a];
                }
        }
        flip_centre_T(g, x, p.flags);
        store_row48(s_pose + lane * kWorldRow4, x);
    }
    __syncwarp();
    if (rows == kTile) store_padded_tile<kWorldChunks>(s_pose, p.out_pos, row0);
    else stage_padded_out<kWorldChunks>(s_pose, p.out_pos, row0, rows);
}

template <int KC, bool POS>
__global__ void __launch_bounds__(kTile) dhfk_critic_jvp_kernel(const __grid_constant__ CriticParams p) {
    extern __shared__ __align__(16) float smem[];
    float4* s_pose = reinterpret_cast<float4*>(smem);
    float4* s_v = s_pose + kTile * kWorldRow4;
    float* s_kcs = reinterpret_cast<float*>(s_v + kTile * kWorldRow4);
    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    if (rows == kTile) {
        if (KC > 0) ldgsts_padded_tile<kWorldChunks>(s_pose, p.pose, row0);
        ldgsts_padded_tile<kWorldChunks>(s_v, p.a, row0);
    } else {
        if (KC > 0) stage_padded_in<kWorldChunks>(s_pose, p.pose, row0, rows);
        stage_padded_in<kWorldChunks>(s_v, p.a, row0, rows);
    }
    ldgsts_wait_all();     // full tiles and ragged tiles alike: everything above was queued with cp.async
    __syncwarp();
    if (lane < rows) {
        float v[48], ty[48];
        load_row48(s_v + lane * kWorldRow4, v);
        flip_centre(v, ty, p.flags);
        if (POS) store_row48(s_v + lane * kWorldRow4, ty);
        if (KC > 0) {
            float x[48], y[48];
            load_row48(s_pose + lane * kWorldRow4, x);
            flip_centre(x, y, p.flags);
            Bones B, Tb;
            bones_of(y, B);
            float tl[15];   // (b . tb) / s  = relative change of the length
#pragma unroll
            for (int b = 0; b < 15; ++b) {
#pragma unroll
                for (int a = 0; a < 3; ++a) Tb.v[3 * b + a] = ty[3 * KB1[b] + a] - ty[3 * KB0[b] + a];
                tl[b] = bdot(B.v + 3 * b, Tb.v + 3 * b) * (B.inv[b] * B.inv[b]);
            }
            float* k = s_kcs + lane * KC;
#pragma unroll
            for (int q = 0; q < 15; ++q) {
                const int i = KP0[q], j = KP1[q];
                const float ij = B.inv[i] * B.inv[j];
                const float c = bdot(B.v + 3 * i, B.v + 3 * j) * ij;
                const float d = (bdot(Tb.v + 3 * i, B.v + 3 * j) + bdot(B.v + 3 * i, Tb.v + 3 * j)) * ij;
                k[q] = fmaf(-c, tl[i] + tl[j], d);
            }
            if (KC == 30) {
#pragma unroll
                for (int b = 0; b < 15; ++b) k[15 + b] = tl[b] * sqrtf(B.s[b]);
            }
        }
    }
    __syncwarp();
    if (POS) {
        if (rows == kTile) store_padded_tile<kWorldChunks>(s_v, p.out_pos, row0);
        else stage_padded_out<kWorldChunks>(s_v, p.out_pos, row0, rows);
    }
    if (KC > 0) {
        RowDst d; d.p = p.out_kcs; d.stride = KC; d.vec = 1;
        stage_rows_out<(KC > 0 ? KC : 1)>(s_kcs, d, row0, rows);
    }
}

// left/right flip of [N,16,2] keypoints: same tile staging as everything else (one thread per pose permutes its
// 32 floats in registers; 128-bit coalesced both ways).  A one-thread-per-chunk gather measured 87 %.
struct Flip2dParams {
    const float* x;
    float* out;
    long long n;
};
__global__ void __launch_bounds__(kTile) dhfk_flip2d_tile_kernel(const __grid_constant__ Flip2dParams p) {
    extern __shared__ __align__(16) float smem[];
    float4* s_uv = reinterpret_cast<float4*>(smem);
    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    if (rows == kTile) {
        ldgsts_padded_tile<kUvChunks>(s_uv, p.x, row0);
    } else {
        stage_padded_in<kUvChunks>(s_uv, p.x, row0, rows);
    }
    ldgsts_wait_all();     // full tiles and ragged tiles alike: everything above was queued with cp.async
    __syncwarp();
    if (lane < rows) {
        float4* row = s_uv + lane * kUvRow4;
        float x[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = row[c];
            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
            row[c] = make_float4(-x[2 * FLIP16[2 * c]], x[2 * FLIP16[2 * c] + 1], -x[2 * FLIP16[2 * c + 1]],
                                 x[2 * FLIP16[2 * c + 1] + 1]);
    }
    __syncwarp();
    if (rows == kTile) store_padded_tile<kUvChunks>(s_uv, p.out, row0);
    else stage_padded_out<kUvChunks>(s_uv, p.out, row0, rows);
}

template <int KC, bool POS>
static int launch_fwd(const CriticParams& p, cudaStream_t st, const char** where) {
    const size_t smem = sizeof(float4) * kTile * kWorldRow4 + sizeof(float) * kTile * KC;
    return launch_tiles(dhfk_critic_fwd_kernel<KC, POS>, smem, p, st, where);
}
template <int KC, bool GPOS>
static int launch_bwd(const CriticParams& p, cudaStream_t st, const char** where) {
    const size_t smem = sizeof(float4) * kTile * kWorldRow4 * (GPOS ? 2 : 1) + sizeof(float) * kTile * KC;
    return launch_tiles(dhfk_critic_bwd_kernel<KC, GPOS>, smem, p, st, where);
}
template <int KC, bool POS>
static int launch_jvp(const CriticParams& p, cudaStream_t st, const char** where) {
    const size_t smem = sizeof(float4) * kTile * kWorldRow4 * 2 + sizeof(float) * kTile * KC;
    return launch_tiles(dhfk_critic_jvp_kernel<KC, POS>, smem, p, st, where);
}

// mode 0 forward, 1 backward (vjp), 2 jvp.  kc in {0, 15, 30}; `pos` = positional output / g_pos present.
int launch_critic(int mode, int kc, bool pos, const float* pose, const float* a, const float* b, float* out_pos,
                  float* out_kcs, long long n, unsigned flags, cudaStream_t st, const char** where) {
    CriticParams p;
    p.pose = pose; p.a = a; p.b = b; p.out_pos = out_pos; p.out_kcs = out_kcs; p.n = n; p.flags = flags;
#define DHFK_CRITIC_DISPATCH(FN)                                                   \
    if (kc == 30) return pos ? FN<30, true>(p, st, where) : FN<30, false>(p, st, where); \
    if (kc == 15) return pos ? FN<15, true>(p, st, where) : FN<15, false>(p, st, where); \
    return FN<0, true>(p, st, where);
    if (mode == 0) { DHFK_CRITIC_DISPATCH(launch_fwd) }
    if (mode == 1) { DHFK_CRITIC_DISPATCH(launch_bwd) }
    DHFK_CRITIC_DISPATCH(launch_jvp)
#undef DHFK_CRITIC_DISPATCH
}

int launch_flip(const float* x, float* out, long long n, int dims, cudaStream_t st, const char** where) {
    // 3-D: the tiled critic kernel with only the flip flag (smem-staged, 128-bit both ways: ~98 % of copy peak;
    // a per-chunk gather of 12-byte joints measured 41 %).  2-D: its own tiled kernel.
    if (dims == 3) return launch_critic(0, 0, true, x, nullptr, nullptr, out, nullptr, n, kFlip, st, where);
    Flip2dParams p;
    p.x = x; p.out = out; p.n = n;
    return launch_tiles(dhfk_flip2d_tile_kernel, sizeof(float4) * kTile * kUvRow4, p, st, where);
}

}  // namespace dhfk
