// dhfk_camera.cu -- tiled versions of the standalone camera ops for 16-joint poses (the shape every reference call
// site has): GAN_torch_world_to_camera (common/camera.py:36-38) and project_to_2d with per-row intrinsics
// (common/camera.py:62-94), forward and backward.  One thread per pose, one warp per 32-pose tile, the same
// LDGSTS-in / coalesced-128-bit-out staging as the fused kernels (the one-thread-per-point kernels in
// dhfk_cabi.cu, kept for other joint counts and unaligned buffers, measured 70-87 % of the copy roofline).
#include "dhfk_launch.h"

namespace dhfk {

struct CamTileParams {
    const float* x;          // [N,16,3]: world / camera-space points, or (w2c backward) the upstream gradient
    const float* g_uv;       // [N,16,2] upstream gradient (projection backward)
    const float* cam_rows;   // [N, cam_stride] f2 c2 k3 p2 (projection)
    long long cam_stride;
    const float* q_dev;      // camera quaternion / translation in device memory, or nullptr: use M / t below
    const float* t_dev;
    float M[9];
    float t[3];
    float* out;              // [N,16,3] or [N,16,2]
    long long n;
};

DHFK_DI void quat_matrix(const float* __restrict__ q, float* M) {   // linear map of qrot(conj q, .), quaternion.py:6-35
    const float w = q[0], ux = -q[1], uy = -q[2], uz = -q[3];
    const float xx = ux * ux, yy = uy * uy, zz = uz * uz, xy = ux * uy, xz = ux * uz, yz = uy * uz;
    M[0] = 1.f - 2.f * (yy + zz); M[1] = 2.f * (xy - w * uz);   M[2] = 2.f * (xz + w * uy);
    M[3] = 2.f * (xy + w * uz);   M[4] = 1.f - 2.f * (xx + zz); M[5] = 2.f * (yz - w * ux);
    M[6] = 2.f * (xz - w * uy);   M[7] = 2.f * (yz + w * ux);   M[8] = 1.f - 2.f * (xx + yy);
}

// MODE 0: out = M (x - t)   1: out = M^T x   2: uv = project(x; row)   3: g_x = d project / d x ^T g_uv
template <int MODE>
__global__ void __launch_bounds__(kTile) dhfk_camera_tile_kernel(const __grid_constant__ CamTileParams p) {
    extern __shared__ __align__(16) float smem[];
    float4* s_x = reinterpret_cast<float4*>(smem);                   // padded rows, 13 chunks
    float4* s_uv = s_x + kTile * kWorldRow4;                         // padded rows, 9 chunks (MODE 2 out, MODE 3 in)
    float* s_cam = reinterpret_cast<float*>(s_uv + (MODE >= 2 ? kTile * kUvRow4 : 0));
    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    const bool cam_slab = MODE >= 2 && rows == kTile && p.cam_stride == 9 &&
                          (reinterpret_cast<unsigned long long>(p.cam_rows) & 15ull) == 0;
    if (rows == kTile) {
        ldgsts_padded_tile<kWorldChunks>(s_x, p.x, row0);
        if (MODE == 3) ldgsts_padded_tile<kUvChunks>(s_uv, p.g_uv, row0);
        if (cam_slab) ldgsts_slab<9>(s_cam, p.cam_rows + row0 * 9);
    } else {
        stage_padded_in<kWorldChunks>(s_x, p.x, row0, rows);
        if (MODE == 3) stage_padded_in<kUvChunks>(s_uv, p.g_uv, row0, rows);
    }
    ldgsts_wait_all();     // full tiles and ragged tiles alike: everything above was queued with cp.async
    __syncwarp();
    if (lane < rows) {
        float4* xrow = s_x + lane * kWorldRow4;
        float x[48];
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            const float4 v = xrow[c];
            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
        }
        if (MODE <= 1) {
            float M[9], t[3] = {0.f, 0.f, 0.f};
            if (p.q_dev) {
                quat_matrix(p.q_dev, M);
                if (MODE == 0) { t[0] = __ldg(p.t_dev); t[1] = __ldg(p.t_dev + 1); t[2] = __ldg(p.t_dev + 2); }
            } else {
#pragma unroll
                for (int i = 0; i < 9; ++i) M[i] = p.M[i];
                if (MODE == 0) { t[0] = p.t[0]; t[1] = p.t[1]; t[2] = p.t[2]; }
            }
            float y[48];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const V3 v = v3(x[3 * k] - t[0], x[3 * k + 1] - t[1], x[3 * k + 2] - t[2]);
                const V3 o = MODE == 0 ? mat_vec(M, v) : matT_vec(M, v);
                y[3 * k] = o.x; y[3 * k + 1] = o.y; y[3 * k + 2] = o.z;
            }
#pragma unroll
            for (int c = 0; c < 12; ++c) xrow[c] = make_float4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
        } else {
            float cr[9];
            if (cam_slab) {
#pragma unroll
                for (int i = 0; i < 9; ++i) cr[i] = s_cam[lane * 9 + i];
            } else {
                const float* g = p.cam_rows + (row0 + lane) * p.cam_stride;
#pragma unroll
                for (int i = 0; i < 9; ++i) cr[i] = __ldg(g + i);
            }
            CamConst cc;
            cc.f = make_float2(cr[0], cr[1]); cc.c = make_float2(cr[2], cr[3]);
            cc.k[0] = cr[4]; cc.k[1] = cr[5]; cc.k[2] = cr[6];
            cc.p = make_float2(cr[7], cr[8]);
            cc.k1x2 = 2.f * cc.k[1]; cc.k2x3 = 3.f * cc.k[2];
            float4* urow = s_uv + lane * kUvRow4;
            if (MODE == 2) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    ProjAux a;
                    float u0, v0, u1, v1;
                    project_point(cc, v3(x[6 * c], x[6 * c + 1], x[6 * c + 2]), u0, v0, a);
                    project_point(cc, v3(x[6 * c + 3], x[6 * c + 4], x[6 * c + 5]), u1, v1, a);
                    urow[c] = make_float4(u0, v0, u1, v1);
                }
            } else {
                float y[48];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 g = urow[c];
                    ProjAux a;
                    float u, v;
                    project_point(cc, v3(x[6 * c], x[6 * c + 1], x[6 * c + 2]), u, v, a);
                    V3 o = project_point_bwd(cc, a, g.x, g.y);
                    y[6 * c] = o.x; y[6 * c + 1] = o.y; y[6 * c + 2] = o.z;
                    project_point(cc, v3(x[6 * c + 3], x[6 * c + 4], x[6 * c + 5]), u, v, a);
                    o = project_point_bwd(cc, a, g.z, g.w);
                    y[6 * c + 3] = o.x; y[6 * c + 4] = o.y; y[6 * c + 5] = o.z;
                }
#pragma unroll
                for (int c = 0; c < 12; ++c) xrow[c] = make_float4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
            }
        }
    }
    __syncwarp();
    if (MODE == 2) {
        if (rows == kTile) store_padded_tile<kUvChunks>(s_uv, p.out, row0);
        else stage_padded_out<kUvChunks>(s_uv, p.out, row0, rows);
    } else {
        if (rows == kTile) store_padded_tile<kWorldChunks>(s_x, p.out, row0);
        else stage_padded_out<kWorldChunks>(s_x, p.out, row0, rows);
    }
}

// mode as above.  Every pointer that is given must be 16-byte aligned (the caller checks and falls back otherwise).
int launch_camera_tiles(int mode, const float* x, const float* g_uv, const float* cam_rows, long long cam_stride,
                        const float* q_dev, const float* t_dev, const float* M, const float* t, float* out,
                        long long n, cudaStream_t st, const char** where) {
    CamTileParams p;
    p.x = x; p.g_uv = g_uv; p.cam_rows = cam_rows; p.cam_stride = cam_stride; p.q_dev = q_dev; p.t_dev = t_dev;
    for (int i = 0; i < 9; ++i) p.M[i] = M ? M[i] : 0.f;
    for (int i = 0; i < 3; ++i) p.t[i] = t ? t[i] : 0.f;
    p.out = out; p.n = n;
    const size_t base = sizeof(float4) * kTile * kWorldRow4;
    const size_t proj = base + sizeof(float4) * kTile * kUvRow4 + sizeof(float) * kTile * 9;
    switch (mode) {
        case 0: return launch_tiles(dhfk_camera_tile_kernel<0>, base, p, st, where);
        case 1: return launch_tiles(dhfk_camera_tile_kernel<1>, base, p, st, where);
        case 2: return launch_tiles(dhfk_camera_tile_kernel<2>, proj, p, st, where);
        default: return launch_tiles(dhfk_camera_tile_kernel<3>, proj, p, st, where);
    }
}

}  // namespace dhfk
