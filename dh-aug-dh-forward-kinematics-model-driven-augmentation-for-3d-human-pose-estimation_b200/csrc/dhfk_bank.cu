// dhfk_bank.cu -- SURVEY 8 (f4): device-resident fake-pair bank.
// The reference copies pos_3d_cam / uv / cam to host numpy every iteration
// (models_Fk_GAN/model_fk_gan_train.py:486-488), concatenates them and re-serves them through a CPU
// DataLoader with pin_memory (:504-510; common/data_loader.py:9-36), i.e. one D2H and one H2D per batch.
// Here the pairs stay in HBM as one record per pose -- [pose3d 48 | pose2d 32 | cam, padded] floats, 384 bytes
// by default -- so that a shuffled mini-batch reads ONE contiguous 384-byte span per pose (three separate
// arrays with 192 / 128 / 36-byte rows measured 50 % of the copy roofline, the 36-byte camera rows straddling
// sectors).  Output wire format per pose, identical to what PoseDataSet.__getitem__ yields:
// pose3d [16,3], pose2d [16,2], cam [cam_cols] fp32, three packed tensors.
#include "dhfk_launch.h"

namespace dhfk {

constexpr int kRec3d = kWorldChunks;              // chunks 0..11  : pose3d
constexpr int kRec2d = kWorldChunks + kUvChunks;  // chunks 12..19 : pose2d;  chunks 20.. : camera row

// One thread per 16-byte chunk of a record, kBankUnroll chunks per thread a whole block apart: the kBankUnroll index
// loads are issued back to back, then the kBankUnroll record loads, then the stores -- a random gather is latency-bound,
// so throughput is bytes in flight.  (r1: one chunk per thread = 32 KB in flight per SM = 64 % of the streaming peak.)
// rec_chunks = record stride in 16-byte chunks (>= 20 + ceil(cam_cols/4)).
constexpr int kBankThreads = 256;
#ifndef DHFK_BANK_UNROLL
#define DHFK_BANK_UNROLL 4
#endif
constexpr int kBankUnroll = DHFK_BANK_UNROLL;

// IDX32: nb * used < 2^32, so (record, chunk) come from 32-bit arithmetic -- a 64-bit division is ~60 emulated
// instructions, four of them per thread made this copy kernel issue-bound (ncu r2e: issue active 58 %, DRAM 65 %).
template <bool IDX32>
__global__ void __launch_bounds__(kBankThreads)
dhfk_bank_gather_kernel(const float4* __restrict__ bank, int rec_chunks, int cam_cols, const long long* __restrict__ idx,
                        long long nb, long long bank_rows, float4* __restrict__ out3d, float4* __restrict__ out2d,
                        float* __restrict__ out_cam) {
    const int used = kRec2d + (out_cam ? (cam_cols + 3) / 4 : 0);   // chunks of a record that are consumed
    const long long total = nb * used;
    const long long i0 = (long long)blockIdx.x * (kBankThreads * kBankUnroll) + threadIdx.x;
    long long b[kBankUnroll], r[kBankUnroll];
    int c[kBankUnroll];
#pragma unroll
    for (int u = 0; u < kBankUnroll; ++u) {
        const long long i = i0 + (long long)u * kBankThreads;
        if (IDX32) {
            const unsigned bi = (unsigned)i / (unsigned)used;
            b[u] = bi;
            c[u] = (int)((unsigned)i - bi * (unsigned)used);
        } else {
            b[u] = i / used;
            c[u] = (int)(i - b[u] * used);
        }
        r[u] = i < total ? __ldg(idx + b[u]) : -1;
    }
    const float nan = __int_as_float(0x7fc00000);
    float4 v[kBankUnroll];
#pragma unroll
    for (int u = 0; u < kBankUnroll; ++u) {
        const bool ok = r[u] >= 0 && r[u] < bank_rows;   // an out-of-range index yields a NaN row, never a wild read
        v[u] = ok ? __ldg(bank + r[u] * rec_chunks + c[u]) : make_float4(nan, nan, nan, nan);
    }
#pragma unroll
    for (int u = 0; u < kBankUnroll; ++u) {
        if (i0 + (long long)u * kBankThreads >= total) continue;
        const int cu = c[u];
        if (cu < kRec3d) __stcs(out3d + b[u] * kWorldChunks + cu, v[u]);
        else if (cu < kRec2d) __stcs(out2d + b[u] * kUvChunks + (cu - kRec3d), v[u]);
        else {
            const int k0 = 4 * (cu - kRec2d);
            float* o = out_cam + b[u] * cam_cols + k0;
            if (k0 < cam_cols) o[0] = v[u].x;
            if (k0 + 1 < cam_cols) o[1] = v[u].y;
            if (k0 + 2 < cam_cols) o[2] = v[u].z;
            if (k0 + 3 < cam_cols) o[3] = v[u].w;
        }
    }
}

int launch_bank_gather(const float* bank, long long rec_floats, int cam_cols, const long long* idx, long long nb,
                       long long bank_rows, float* out3d, float* out2d, float* out_cam, cudaStream_t st,
                       const char** where) {
    const int used = kRec2d + (out_cam ? (cam_cols + 3) / 4 : 0);
    const long long nthreads = nb * used;
    const long long per_block = kBankThreads * kBankUnroll;
    const unsigned blocks = (unsigned)((nthreads + per_block - 1) / per_block);
    if (nthreads + per_block < (1LL << 32))
        dhfk_bank_gather_kernel<true><<<blocks, kBankThreads, 0, st>>>(
            reinterpret_cast<const float4*>(bank), (int)(rec_floats / 4), cam_cols, idx, nb, bank_rows,
            reinterpret_cast<float4*>(out3d), reinterpret_cast<float4*>(out2d), out_cam);
    else
        dhfk_bank_gather_kernel<false><<<blocks, kBankThreads, 0, st>>>(
            reinterpret_cast<const float4*>(bank), (int)(rec_floats / 4), cam_cols, idx, nb, bank_rows,
            reinterpret_cast<float4*>(out3d), reinterpret_cast<float4*>(out2d), out_cam);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *where = "dhfk_bank_gather_kernel"; return (int)e; }
    return 0;
}

}  // namespace dhfk
