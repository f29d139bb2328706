// dhfk_scatter32.cu -- the reference's 32-slot output layout, in one launch each way.
// Forward_Kinematics_DH_Model.change_3d_joint_angle returns [N,32,3] (forward_kinematics_DH_model.py:745-820): the 16
// joints sit in the H36M 32-joint slots (common/h36m_dataset.py:37-38), slot 14 repeats the head joint of slot 15, every
// other slot holds 0 + root.  The plain drop-in path (dropin.install() without the generator swap) has to hand that
// tensor back; built from torch ops it is an expand + clone + index_copy + slice write (and their autograd mirror).
//   forward : world16 [N,16,3], root [N,3]  ->  world32 [N,32,3]
//   backward: g32 [N,32,3]  ->  g16 [N,16,3] (slot 14 folded into the head joint), g_root [N,3] = sum of the 15 free slots
// One thread per pose, tile staging as everywhere else (96-float rows: 24 + 1 chunks).
#include "dhfk_launch.h"

namespace dhfk {

constexpr int kW32Chunks = 24;
constexpr int kW32Row4 = kW32Chunks + 1;
constexpr int kHeadOut = 9;        // output joint stored in slot 15 and repeated in slot 14 (:797-803)

struct Scatter32Params {
    const float* in;      // forward: world16 [N,16,3]; backward: g32 [N,32,3]
    const float* root;    // forward: [N, root_stride]
    long long root_stride;
    float* out;           // forward: world32 [N,32,3]; backward: g16 [N,16,3]
    float* g_root;        // backward: [N,3] packed
    long long n;
};

// slot -> output joint (-1: free slot); checked against H36M_32_TO_16 below
__device__ constexpr int SLOT_JOINT[32] = {0, 1, 2, 3, -1, -1, 4, 5, 6, -1, -1, -1, 7, 8, 9, 9,
                                           -1, 10, 11, 12, -1, -1, -1, -1, -1, 13, 14, 15, -1, -1, -1, -1};
constexpr bool slot_table_ok() {
    constexpr int T[32] = {0, 1, 2, 3, -1, -1, 4, 5, 6, -1, -1, -1, 7, 8, 9, 9,
                           -1, 10, 11, 12, -1, -1, -1, -1, -1, 13, 14, 15, -1, -1, -1, -1};
    for (int k = 0; k < NOUT; ++k)
        if (T[H36M_32_TO_16[k]] != k) return false;
    int used = 0;
    for (int s = 0; s < 32; ++s) used += T[s] >= 0;
    return used == 17 && T[14] == kHeadOut && T[15] == kHeadOut;
}
static_assert(slot_table_ok(), "SLOT_JOINT must invert H36M_32_TO_16 (plus the repeated head slot 14)");

template <bool BWD>
__global__ void __launch_bounds__(kTile) dhfk_scatter32_kernel(const __grid_constant__ Scatter32Params p) {
    extern __shared__ __align__(16) float smem[];
    float4* s16 = reinterpret_cast<float4*>(smem);                     // 13-chunk padded rows
    float4* s32 = s16 + kTile * kWorldRow4;                            // 25-chunk padded rows
    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kTile;
    const long long left = p.n - row0;
    const int rows = left < kTile ? (int)left : kTile;
    if (rows == kTile) {
        if (BWD) ldgsts_padded_tile<kW32Chunks>(s32, p.in, row0);
        else ldgsts_padded_tile<kWorldChunks>(s16, p.in, row0);
    } else {
        if (BWD) stage_padded_in<kW32Chunks>(s32, p.in, row0, rows);
        else stage_padded_in<kWorldChunks>(s16, p.in, row0, rows);
    }
    ldgsts_wait_all();     // full tiles and ragged tiles alike: everything above was queued with cp.async
    __syncwarp();
    if (lane < rows) {
        float4* r16 = s16 + lane * kWorldRow4;
        float4* r32 = s32 + lane * kW32Row4;
        float a[48], b[96];
        if (!BWD) {
#pragma unroll
            for (int c = 0; c < 12; ++c) {
                const float4 v = r16[c];
                a[4 * c] = v.x; a[4 * c + 1] = v.y; a[4 * c + 2] = v.z; a[4 * c + 3] = v.w;
            }
            const float* rt = p.root + (row0 + lane) * p.root_stride;
            const float rx = __ldg(rt), ry = __ldg(rt + 1), rz = __ldg(rt + 2);
#pragma unroll
            for (int s = 0; s < 32; ++s) {
                const int k = SLOT_JOINT[s];
                b[3 * s] = k >= 0 ? a[3 * k] : rx;
                b[3 * s + 1] = k >= 0 ? a[3 * k + 1] : ry;
                b[3 * s + 2] = k >= 0 ? a[3 * k + 2] : rz;
            }
#pragma unroll
            for (int c = 0; c < 24; ++c) r32[c] = make_float4(b[4 * c], b[4 * c + 1], b[4 * c + 2], b[4 * c + 3]);
        } else {
#pragma unroll
            for (int c = 0; c < 24; ++c) {
                const float4 v = r32[c];
                b[4 * c] = v.x; b[4 * c + 1] = v.y; b[4 * c + 2] = v.z; b[4 * c + 3] = v.w;
            }
            float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
            for (int k = 0; k < 48; ++k) a[k] = 0.f;
#pragma unroll
            for (int s = 0; s < 32; ++s) {
                const int k = SLOT_JOINT[s];
                if (k >= 0) { a[3 * k] += b[3 * s]; a[3 * k + 1] += b[3 * s + 1]; a[3 * k + 2] += b[3 * s + 2]; }
                else { gx += b[3 * s]; gy += b[3 * s + 1]; gz += b[3 * s + 2]; }
            }
#pragma unroll
            for (int c = 0; c < 12; ++c) r16[c] = make_float4(a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
            float* g = p.g_root + (row0 + lane) * 3;
            g[0] = gx; g[1] = gy; g[2] = gz;
        }
    }
    __syncwarp();
    if (BWD) {
        if (rows == kTile) store_padded_tile<kWorldChunks>(s16, p.out, row0);
        else stage_padded_out<kWorldChunks>(s16, p.out, row0, rows);
    } else {
        if (rows == kTile) store_padded_tile<kW32Chunks>(s32, p.out, row0);
        else stage_padded_out<kW32Chunks>(s32, p.out, row0, rows);
    }
}

int launch_scatter32(bool bwd, const float* in, const float* root, long long root_stride, float* out, float* g_root,
                     long long n, cudaStream_t st, const char** where) {
    Scatter32Params p;
    p.in = in; p.root = root; p.root_stride = root_stride; p.out = out; p.g_root = g_root; p.n = n;
    const size_t smem = sizeof(float4) * kTile * (kWorldRow4 + kW32Row4);
    if (bwd) return launch_tiles(dhfk_scatter32_kernel<true>, smem, p, st, where);
    return launch_tiles(dhfk_scatter32_kernel<false>, smem, p, st, where);
}

}  // namespace dhfk
