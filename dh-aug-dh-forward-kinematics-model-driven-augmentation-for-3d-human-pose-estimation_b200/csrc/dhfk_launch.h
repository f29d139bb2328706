// dhfk_launch.h -- internal launch entry points, one translation unit per kernel family so the
// (long, fully unrolled) kernels compile in parallel.
#pragma once
#include "dhfk_kernels.cuh"

namespace dhfk {
// return 0 or a cudaError_t; `where` receives a static string naming the failing call.
// Naming: launch_fwd_t<TRIG>_g<GEN>, launch_bwd_t<TRIG>_b<GBONE>_g<GEN>
int launch_fwd_t0_g0(const FwdParams& p, bool cam, bool uv, cudaStream_t st, const char** where);
int launch_fwd_t1_g0(const FwdParams& p, bool cam, bool uv, cudaStream_t st, const char** where);
int launch_fwd_t0_g1(const FwdParams& p, bool cam, bool uv, cudaStream_t st, const char** where);
int launch_fwd_t1_g1(const FwdParams& p, bool cam, bool uv, cudaStream_t st, const char** where);
// g2 = raw mode with wide rows (the generator's [N,37] slot tensor as one slab per tile); no bone-gradient variant
int launch_fwd_t0_g2(const FwdParams& p, bool cam, bool uv, cudaStream_t st, const char** where);
int launch_fwd_t1_g2(const FwdParams& p, bool cam, bool uv, cudaStream_t st, const char** where);
int launch_bwd_t0_b0_g2(const BwdParams& p, bool guv, cudaStream_t st, const char** where);
int launch_bwd_t1_b0_g2(const BwdParams& p, bool guv, cudaStream_t st, const char** where);
int launch_bwd_t0_b0_g0(const BwdParams& p, bool guv, cudaStream_t st, const char** where);
int launch_bwd_t0_b1_g0(const BwdParams& p, bool guv, cudaStream_t st, const char** where);
int launch_bwd_t1_b0_g0(const BwdParams& p, bool guv, cudaStream_t st, const char** where);
int launch_bwd_t1_b1_g0(const BwdParams& p, bool guv, cudaStream_t st, const char** where);
int launch_bwd_t0_b0_g1(const BwdParams& p, bool guv, cudaStream_t st, const char** where);
int launch_bwd_t1_b0_g1(const BwdParams& p, bool guv, cudaStream_t st, const char** where);

// bone-length retarget (+ per-row projection), forward only
int launch_retarget(const float* pose, const int* tmpl_idx, const float* templates, int num_templates,
                    const float* cam_rows, long long cam_stride, float* out_pose, float* out_uv, long long n,
                    cudaStream_t st, const char** where);

// critic input transforms (flip / root-centre / KCS features): mode 0 forward, 1 vjp, 2 jvp
int launch_critic(int mode, int kc, bool pos, const float* pose, const float* a, const float* b, float* out_pos,
                  float* out_kcs, long long n, unsigned flags, cudaStream_t st, const char** where);
int launch_flip(const float* x, float* out, long long n, int dims, cudaStream_t st, const char** where);

// SURVEY 8 f2, video part: inputs of the motion critics (per-frame KCS, adjacent-frame differences, playback reverse).
// mode 0 forward, 2 jvp (v = tangent)
int launch_video_critic(int mode, const float* pose, const float* v, int frames, unsigned flags, float* out_kcs,
                        float* out_dkcs, float* out_dpos, float* out_pos, long long n, cudaStream_t st, const char** where);
int launch_video_critic_bwd(const float* pose, int frames, unsigned flags, const float* g_kcs, const float* g_dkcs,
                            const float* g_dpos, const float* g_pos, float* g_pose, long long n, cudaStream_t st,
                            const char** where);
int launch_video_root_diff(bool bwd, const float* uv, const float* g_diff, const float* g_pb, int frames, unsigned flags,
                           float* out_diff, float* out_pb, float* g_uv, long long n, cudaStream_t st, const char** where);

// shuffled mini-batch gather out of the device-resident fake-pair bank
int launch_bank_gather(const float* bank, long long rec_floats, int cam_cols, const long long* idx, long long nb,
                       long long bank_rows, float* out3d, float* out2d, float* out_cam, cudaStream_t st,
                       const char** where);

// SURVEY 8 e: in-place gradient all-reduce over NVLink peer memory / NVLS multicast (dhfk_allreduce.cu)
int launch_grad_allreduce(float* const* peer_bufs, float* mc_buf, unsigned* const* peer_flags, unsigned* status, int rank,
                          int world, long long n_floats, float scale, int max_ctas, int threads,
                          unsigned long long timeout_ns, cudaStream_t st, const char** where);

// tiled standalone camera ops for 16-joint poses: mode 0 w2c fwd, 1 w2c bwd, 2 project fwd, 3 project bwd
int launch_camera_tiles(int mode, const float* x, const float* g_uv, const float* cam_rows, long long cam_stride,
                        const float* q_dev, const float* t_dev, const float* M, const float* t, float* out,
                        long long n, cudaStream_t st, const char** where);

// the reference's 32-slot output layout: forward world16 + root -> world32, backward g32 -> g16 + g_root
int launch_scatter32(bool bwd, const float* in, const float* root, long long root_stride, float* out, float* g_root,
                     long long n, cudaStream_t st, const char** where);

// floats per pose in the input slabs: raw mode ang33+grot3+bone15+root3 (wide rows: S + [grot3] + bone15 + root3),
// generator mode out35+bone15
inline size_t in_floats(bool gen, const WideRows& w) {
    if (gen) return GEN_NCOL + 15;
    return (w.wide ? (size_t)w.wide : 33) + (w.grot_slab ? 3 : 0) + 15 + 3;
}
inline size_t fwd_smem_bytes(bool cam, bool uv, bool gen, const WideRows& w) {
    return sizeof(float) * kTile * in_floats(gen, w) +
           sizeof(float4) * kTile * (kWorldRow4 * (1 + (cam ? 1 : 0)) + (uv ? kUvRow4 : 0)) + 16;
}
inline size_t bwd_smem_bytes(bool gw, bool gcam, bool guv, bool gen, const WideRows& w) {
    return sizeof(float) * kTile * in_floats(gen, w) +
           sizeof(float4) * kTile * (kWorldRow4 * ((gw ? 1 : 0) + (gcam ? 1 : 0)) + (guv ? kUvRow4 : 0)) + 16;
}

// cudaFuncSetAttribute is needed once per (kernel, device), not once per launch: at the reference's real batch sizes
// (1 024 / 4 608 poses) the kernels take ~7 us and two attribute calls per launch were a measurable part of the
// 13 us a C-ABI forward+backward pair cost in round 1.  Lock-free set keyed by (kernel address, device).
// MaxDynamicSharedMemorySize is set ONCE to a bound that covers every launch of the library (it is a limit, not a
// reservation: launches asking for more than the value last set fail with cudaErrorInvalidValue, and the same kernel
// is launched with different sizes -- packed rows, wide rows, with / without the separate global-rotation slab).
constexpr int kMaxDynSmem = 64 * 1024;
bool func_attrs_done(const void* kernel, int device);     // true when already recorded; records it otherwise
void func_attrs_forget(const void* kernel, int device);   // undo after a failed attribute call

template <typename K, typename P>
int launch_tiles(K kernel, size_t smem, const P& p, cudaStream_t st, const char** where) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { *where = "cudaGetDevice"; return (int)e; }
    if ((int)smem > kMaxDynSmem) { *where = "dynamic shared memory request exceeds kMaxDynSmem"; return (int)cudaErrorInvalidValue; }
    if (!func_attrs_done((const void*)kernel, dev)) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) {
            func_attrs_forget((const void*)kernel, dev);
            *where = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize / PreferredSharedMemoryCarveout)";
            return (int)e;
        }
    }
    long long blocks = (p.n + kTile - 1) / kTile;
    void* args[] = {const_cast<P*>(&p)};
    e = cudaLaunchKernel((const void*)kernel, dim3((unsigned)blocks), dim3(kTile), args, smem, st);
    if (e != cudaSuccess) { *where = "cudaLaunchKernel"; return (int)e; }
    return 0;
}
}  // namespace dhfk
