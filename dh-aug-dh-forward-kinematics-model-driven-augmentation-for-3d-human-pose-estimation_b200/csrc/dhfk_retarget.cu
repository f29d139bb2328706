// dhfk_retarget.cu -- SURVEY 8 (f3): per-epoch dataset re-augmentation, fused.
//   function_aug/dataloader_update.py:18-41  random_bl_aug: root-centre, unit bone vectors
//       (utils/gan_utils.py:90-110,130-134), multiply by a randomly chosen S1/5/6/7/8 bone-length template
//       row, rebuild the pose along the 16-joint tree (utils/gan_utils.py:56-86), add the root back
//   function_aug/dataloader_update.py:69      project_to_2d(targets_3d, cam_param) with per-row intrinsics
// One thread per pose, one warp per 32-pose tile, same staging as the FK kernels: poses travel through
// padded shared rows (LDGSTS in, coalesced STG.128 out).  Forward only: the reference detaches the result.
#include "dhfk_launch.h"

namespace dhfk {

// 16-joint H36M tree in utils/gan_utils.py bone order: bone b joins PARENT16[b+1] -> joint b+1
__device__ constexpr int PARENT16[16] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 8, 10, 11, 8, 13, 14};

struct RetargetParams {
    const float* pose;        // [N,16,3] packed, 16-byte aligned
    const int* tmpl_idx;      // [N] row of the template table per pose, or nullptr (row 0 for all)
    const float* templates;   // [T,15] device table, utils/gan_utils.py bone order
    int num_templates;
    const float* cam_rows;    // [N, cam_stride] f2 c2 k3 p2 (...); stride 0 = one shared row
    long long cam_stride;
    float* out_pose;          // [N,16,3]
    float* out_uv;            // [N,16,2] or nullptr
    long long n;
};

template <bool PROJ>
__global__ void __launch_bounds__(kTile) dhfk_retarget_kernel(const __grid_constant__ RetargetParams p) {
    extern __shared__ __align__(16) float smem[];
    float4* s_pose = reinterpret_cast<float4*>(smem);                 // padded rows, 13 chunks
    float4* s_uv = s_pose + kTile * kWorldRow4;                        // padded rows, 9 chunks
    float* s_cam = reinterpret_cast<float*>(s_uv + (PROJ ? kTile * kUvRow4 : 0));   // exact image of 32 packed 9-float rows
    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;

    // packed [N,9] intrinsics rows are staged as one 1152-byte slab (coalesced 128-bit requests instead of nine
    // 36-byte-strided scalar loads per lane); wider / shared / unaligned rows are read in place
    const bool cam_slab = PROJ && rows == kTile && p.cam_stride == 9 &&
                          (reinterpret_cast<unsigned long long>(p.cam_rows) & 15ull) == 0;
    if (rows == kTile) {
        ldgsts_padded_tile<kWorldChunks>(s_pose, p.pose, row0);
        if (cam_slab) ldgsts_slab<9>(s_cam, p.cam_rows + row0 * 9);
    } else {
        stage_padded_in<kWorldChunks>(s_pose, p.pose, row0, rows);
    }
    ldgsts_wait_all();     // full tiles and ragged tiles alike: everything above was queued with cp.async
    __syncwarp();

    if (lane < rows) {
        float4* prow = s_pose + lane * kWorldRow4;
        float x[48];
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            float4 v = prow[c];
            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
        }
        int t = p.tmpl_idx ? p.tmpl_idx[row0 + lane] : 0;   // nullptr: one template row for the whole sequence
        t = t < 0 ? 0 : (t >= p.num_templates ? p.num_templates - 1 : t);
        const float* L = p.templates + t * 15;
        float y[48];
        y[0] = x[0]; y[1] = x[1]; y[2] = x[2];                        // the root keeps its position
#pragma unroll
        for (int j = 1; j < 16; ++j) {
            const int q = PARENT16[j];
            float dx = x[3 * j] - x[3 * q], dy = x[3 * j + 1] - x[3 * q + 1], dz = x[3 * j + 2] - x[3 * q + 2];
            float s = __ldg(L + (j - 1)) * rsqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
            y[3 * j] = fmaf(dx, s, y[3 * q]);
            y[3 * j + 1] = fmaf(dy, s, y[3 * q + 1]);
            y[3 * j + 2] = fmaf(dz, s, y[3 * q + 2]);
        }
#pragma unroll
        for (int c = 0; c < 12; ++c) prow[c] = make_float4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
        if (PROJ) {
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
            float4* urow = s_uv + lane * kUvRow4;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                ProjAux a;
                float u0, v0, u1, v1;
                project_point(cc, v3(y[6 * c], y[6 * c + 1], y[6 * c + 2]), u0, v0, a);
                project_point(cc, v3(y[6 * c + 3], y[6 * c + 4], y[6 * c + 5]), u1, v1, a);
                urow[c] = make_float4(u0, v0, u1, v1);
            }
        }
    }
    __syncwarp();
    if (rows == kTile) {
        store_padded_tile<kWorldChunks>(s_pose, p.out_pose, row0);
        if (PROJ) store_padded_tile<kUvChunks>(s_uv, p.out_uv, row0);
    } else {
        stage_padded_out<kWorldChunks>(s_pose, p.out_pose, row0, rows);
        if (PROJ) stage_padded_out<kUvChunks>(s_uv, p.out_uv, row0, rows);
    }
}

int launch_retarget(const float* pose, const int* tmpl_idx, const float* templates, int num_templates,
                    const float* cam_rows, long long cam_stride, float* out_pose, float* out_uv, long long n,
                    cudaStream_t st, const char** where) {
    RetargetParams p;
    p.pose = pose; p.tmpl_idx = tmpl_idx; p.templates = templates; p.num_templates = num_templates;
    p.cam_rows = cam_rows; p.cam_stride = cam_stride; p.out_pose = out_pose; p.out_uv = out_uv; p.n = n;
    const bool proj = out_uv != nullptr;
    const size_t smem = sizeof(float4) * kTile * (kWorldRow4 + (proj ? kUvRow4 : 0)) + (proj ? sizeof(float) * kTile * 9 : 0);
    if (proj) return launch_tiles(dhfk_retarget_kernel<true>, smem, p, st, where);
    return launch_tiles(dhfk_retarget_kernel<false>, smem, p, st, where);
}

}  // namespace dhfk
