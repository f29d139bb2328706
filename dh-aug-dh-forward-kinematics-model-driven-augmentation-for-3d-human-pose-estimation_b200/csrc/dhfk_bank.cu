// dhfk_bank.cu -- SURVEY 8 (f4): device-resident fake-pair bank.
// The reference copies pos_3d_cam / uv / cam to host numpy every iteration
// (models_Fk_GAN/model_fk_gan_train.py:486-488), concatenates them and re-serves them through a CPU
// DataLoader with pin_memory (:504-510; common/data_loader.py:9-36), i.e. one D2H and one H2D per batch.
// Here the pairs stay in HBM; a shuffled mini-batch is one gather launch.  Wire format per pose, identical
// to what PoseDataSet.__getitem__ yields: pose3d [16,3], pose2d [16,2], cam [cam_cols] fp32.
#include "dhfk_launch.h"

namespace dhfk {

constexpr int kBankChunks = kWorldChunks + kUvChunks;   // 16-byte chunks per pose: 12 (3-D) + 8 (2-D)

// One thread per 16-byte chunk of the output batch; the camera row (cam_cols <= 20 floats, not a multiple of
// 16 bytes) is copied one float per thread by the first cam_cols threads of the pose.
__global__ void dhfk_bank_gather_kernel(const float4* __restrict__ bank3d, const float4* __restrict__ bank2d,
                                        const float* __restrict__ bank_cam, int cam_cols,
                                        const long long* __restrict__ idx, long long nb, long long bank_rows,
                                        float4* __restrict__ out3d, float4* __restrict__ out2d,
                                        float* __restrict__ out_cam) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb * kBankChunks) return;
    const long long b = i / kBankChunks;
    const int c = (int)(i - b * kBankChunks);
    const long long r = idx[b];
    const bool ok = r >= 0 && r < bank_rows;          // an out-of-range index yields a NaN row, never a wild read
    const float nan = __int_as_float(0x7fc00000);
    const float4 bad = make_float4(nan, nan, nan, nan);
    if (c < kWorldChunks) out3d[b * kWorldChunks + c] = ok ? __ldg(bank3d + r * kWorldChunks + c) : bad;
    else out2d[b * kUvChunks + (c - kWorldChunks)] = ok ? __ldg(bank2d + r * kUvChunks + (c - kWorldChunks)) : bad;
    if (out_cam && c < cam_cols) out_cam[b * cam_cols + c] = ok ? __ldg(bank_cam + r * cam_cols + c) : nan;
}

int launch_bank_gather(const float* bank3d, const float* bank2d, const float* bank_cam, int cam_cols,
                       const long long* idx, long long nb, long long bank_rows, float* out3d, float* out2d,
                       float* out_cam, cudaStream_t st, const char** where) {
    const long long nthreads = nb * kBankChunks;
    const unsigned blocks = (unsigned)((nthreads + 255) / 256);
    dhfk_bank_gather_kernel<<<blocks, 256, 0, st>>>(
        reinterpret_cast<const float4*>(bank3d), reinterpret_cast<const float4*>(bank2d), bank_cam, cam_cols, idx, nb,
        bank_rows, reinterpret_cast<float4*>(out3d), reinterpret_cast<float4*>(out2d), out_cam);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *where = "dhfk_bank_gather_kernel"; return (int)e; }
    return 0;
}

}  // namespace dhfk
