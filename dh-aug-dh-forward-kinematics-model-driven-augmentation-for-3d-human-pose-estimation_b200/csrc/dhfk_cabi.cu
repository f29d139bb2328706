// dhfk_cabi.cu -- extern "C" boundary of libdhfk.so (include/dhfk.h): argument validation,
// camera constants, kernel dispatch, the standalone camera kernels and the host-buffer pipeline.
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/dhfk.h"
#include "dhfk_launch.h"

namespace dhfk {

// ---- per-(kernel, device) "attributes already set" set: open addressing over atomics, never shrinks -----------
namespace {
constexpr int kAttrDevices = 16, kAttrSlots = 256;     // ~90 kernel instantiations in the library
std::atomic<uintptr_t> g_attr_keys[kAttrDevices][kAttrSlots];
constexpr uintptr_t kTombstone = ~uintptr_t(0);
int attr_slot(uintptr_t key) { return (int)(((unsigned long long)key * 0x9E3779B97F4A7C15ull) >> 56) & (kAttrSlots - 1); }
}  // namespace

bool func_attrs_done(const void* kernel, int device) {
    if (device < 0 || device >= kAttrDevices) return false;          // unknown ordinal: set the attributes every time
    std::atomic<uintptr_t>* keys = g_attr_keys[device];
    const uintptr_t key = reinterpret_cast<uintptr_t>(kernel);
    int s = attr_slot(key);
    for (int probe = 0; probe < kAttrSlots; ++probe, s = (s + 1) & (kAttrSlots - 1)) {
        uintptr_t cur = keys[s].load(std::memory_order_acquire);
        if (cur == key) return true;
        if (cur == 0) {
            if (keys[s].compare_exchange_strong(cur, key, std::memory_order_acq_rel)) return false;
            if (cur == key) return true;
        }
    }
    return false;   // table full: fall back to setting the attributes on every launch
}
void func_attrs_forget(const void* kernel, int device) {
    if (device < 0 || device >= kAttrDevices) return;
    std::atomic<uintptr_t>* keys = g_attr_keys[device];
    const uintptr_t key = reinterpret_cast<uintptr_t>(kernel);
    int s = attr_slot(key);
    for (int probe = 0; probe < kAttrSlots; ++probe, s = (s + 1) & (kAttrSlots - 1)) {
        uintptr_t cur = keys[s].load(std::memory_order_acquire);
        if (cur == key) { keys[s].store(kTombstone, std::memory_order_release); return; }
        if (cur == 0) return;
    }
}

// ---- standalone camera ops (common/camera.py used on its own) ---------------------------------
struct RotConst { float M[9]; float t[3]; };

// out = M (x - t): one thread per point, 3 floats; consecutive threads touch consecutive 12-byte
// records so every 128-byte line is fully consumed across the three loads.
__global__ void __launch_bounds__(256) w2c_fwd_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                     long long npts, const __grid_constant__ RotConst rc) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    V3 v = v3(x[3 * i] - rc.t[0], x[3 * i + 1] - rc.t[1], x[3 * i + 2] - rc.t[2]);
    V3 o = mat_vec(rc.M, v);
    out[3 * i] = o.x; out[3 * i + 1] = o.y; out[3 * i + 2] = o.z;
}
__global__ void __launch_bounds__(256) w2c_bwd_kernel(const float* __restrict__ g, float* __restrict__ gx,
                                                     long long npts, const __grid_constant__ RotConst rc) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    V3 o = matT_vec(rc.M, v3(g[3 * i], g[3 * i + 1], g[3 * i + 2]));
    gx[3 * i] = o.x; gx[3 * i + 1] = o.y; gx[3 * i + 2] = o.z;
}

// same map with q / t read from device memory (fp32 matrix build per thread; 7 broadcast loads)
DHFK_DI void rot_from_quat_dev(const float* __restrict__ q, float* M) {
    const float w = q[0], ux = -q[1], uy = -q[2], uz = -q[3];
    // M = I + 2 w [u]x + 2 [u]x^2
    const float xx = ux * ux, yy = uy * uy, zz = uz * uz, xy = ux * uy, xz = ux * uz, yz = uy * uz;
    M[0] = 1.f - 2.f * (yy + zz); M[1] = 2.f * (xy - w * uz);   M[2] = 2.f * (xz + w * uy);
    M[3] = 2.f * (xy + w * uz);   M[4] = 1.f - 2.f * (xx + zz); M[5] = 2.f * (yz - w * ux);
    M[6] = 2.f * (xz - w * uy);   M[7] = 2.f * (yz + w * ux);   M[8] = 1.f - 2.f * (xx + yy);
}
__global__ void __launch_bounds__(256) w2c_fwd_dev_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                         long long npts, const float* __restrict__ q,
                                                         const float* __restrict__ t) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    float M[9];
    rot_from_quat_dev(q, M);
    V3 v = v3(x[3 * i] - t[0], x[3 * i + 1] - t[1], x[3 * i + 2] - t[2]);
    V3 o = mat_vec(M, v);
    out[3 * i] = o.x; out[3 * i + 1] = o.y; out[3 * i + 2] = o.z;
}
__global__ void __launch_bounds__(256) w2c_bwd_dev_kernel(const float* __restrict__ g, float* __restrict__ gx,
                                                         long long npts, const float* __restrict__ q) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    float M[9];
    rot_from_quat_dev(q, M);
    V3 o = matT_vec(M, v3(g[3 * i], g[3 * i + 1], g[3 * i + 2]));
    gx[3 * i] = o.x; gx[3 * i + 1] = o.y; gx[3 * i + 2] = o.z;
}

DHFK_DI CamConst load_cam_row(const float* row) {
    CamConst cc;
    cc.f = make_float2(row[0], row[1]); cc.c = make_float2(row[2], row[3]);
    cc.k[0] = row[4]; cc.k[1] = row[5]; cc.k[2] = row[6]; cc.p = make_float2(row[7], row[8]);
    cc.k1x2 = 2.f * cc.k[1]; cc.k2x3 = 3.f * cc.k[2];
    return cc;
}
// project_to_2d with per-row intrinsics; one thread per (row, joint) point
__global__ void __launch_bounds__(256) project_fwd_kernel(const float* __restrict__ x, const float* __restrict__ cam,
                                                         long long cam_stride, float* __restrict__ uv,
                                                         long long npts, int joints) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    CamConst cc = load_cam_row(cam + (i / joints) * cam_stride);
    ProjAux a;
    float u, v;
    project_point(cc, v3(x[3 * i], x[3 * i + 1], x[3 * i + 2]), u, v, a);
    reinterpret_cast<float2*>(uv)[i] = make_float2(u, v);
}
__global__ void __launch_bounds__(256) project_bwd_kernel(const float* __restrict__ x, const float* __restrict__ cam,
                                                         long long cam_stride, const float* __restrict__ g_uv,
                                                         float* __restrict__ gx, long long npts, int joints) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    CamConst cc = load_cam_row(cam + (i / joints) * cam_stride);
    ProjAux a;
    float u, v;
    project_point(cc, v3(x[3 * i], x[3 * i + 1], x[3 * i + 2]), u, v, a);
    float2 g = reinterpret_cast<const float2*>(g_uv)[i];
    V3 o = project_point_bwd(cc, a, g.x, g.y);
    gx[3 * i] = o.x; gx[3 * i + 1] = o.y; gx[3 * i + 2] = o.z;
}

}  // namespace dhfk

// =================================================================================================
// Host side: C ABI
// =================================================================================================

namespace {

using namespace dhfk;

thread_local char g_err[512] = "";

int fail(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof g_err, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

bool bad_trig_flags(uint32_t flags) {
    return (flags & DHFK_FLAG_FAST_TRIG) && (flags & DHFK_FLAG_ACCURATE_TRIG);
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

RowSrc row_src(const float* p, int64_t stride, int ncols) {
    RowSrc r;
    r.p = p;
    r.stride = stride;
    r.vec = (stride == ncols && aligned16(p)) ? 1 : 0;
    return r;
}
RowDst row_dst(float* p, int64_t stride, int ncols) {
    RowDst r;
    r.p = p;
    r.stride = stride;
    r.vec = (p != nullptr && stride == ncols && aligned16(p)) ? 1 : 0;
    return r;
}

// v -> qrot(conj(q), v) as a 3x3 matrix (double arithmetic on the host), common/quaternion.py:6-35
void camera_matrix(const float* q, float* M) {
    double w = q[0], u[3] = {-(double)q[1], -(double)q[2], -(double)q[3]};
    for (int col = 0; col < 3; ++col) {
        double v[3] = {0, 0, 0};
        v[col] = 1;
        double uv[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        double uuv[3] = {u[1] * uv[2] - u[2] * uv[1], u[2] * uv[0] - u[0] * uv[2], u[0] * uv[1] - u[1] * uv[0]};
        for (int r = 0; r < 3; ++r) M[r * 3 + col] = (float)(v[r] + 2 * (w * uv[r] + uuv[r]));
    }
}
CamConst make_cam(const float* cam) {
    CamConst cc;
    memset(&cc, 0, sizeof cc);
    if (cam) {
        camera_matrix(cam, cc.M);
        for (int i = 0; i < 3; ++i) cc.t[i] = cam[4 + i];
        for (int i = 0; i < 3; ++i) cc.Mc[i] = make_float2(cc.M[i], cc.M[3 + i]);
        cc.txy = make_float2(cc.t[0], cc.t[1]);
        cc.f = make_float2(cam[7], cam[8]); cc.c = make_float2(cam[9], cam[10]);
        cc.k[0] = cam[11]; cc.k[1] = cam[12]; cc.k[2] = cam[13]; cc.p = make_float2(cam[14], cam[15]);
        cc.k1x2 = 2.f * cc.k[1];
        cc.k2x3 = 3.f * cc.k[2];
    }
    return cc;
}

// Angles and global rotation handed over as column slices of ONE 16-byte aligned [N,S] tensor (the generator's [N,37]
// layout): stage it as a slab.  Max S keeps 12 CTAs per SM only for S <= 37; larger rows still work, at lower occupancy.
WideRows detect_wide(const float* ang, int64_t as, const float* grot, int64_t gs, const float* bone, int64_t bs,
                     const float* root, int64_t rs, int64_t n) {
    WideRows w = {0, 0, 1};
    if (as == gs && as > 33 && as <= 64 && aligned16(ang) && grot >= ang + 33 && grot + 3 <= ang + as) {
        w.wide = (int)as;
        w.goff = (int)(grot - ang);
        // no ragged last tile and every tile takes the slab path => the separate [32,3] slab is never touched
        w.grot_slab = (n % dhfk::kTile == 0 && bs == 15 && aligned16(bone) && rs == 3 && aligned16(root)) ? 0 : 1;
    }
    return w;
}

int check_inputs(const float* ang, int64_t as, const float* grot, int64_t gs, const float* bone, int64_t bs,
                 const float* root, int64_t rs, int64_t n) {
    if (n < 0) return fail(DHFK_E_INVAL, "n must be >= 0");
    if (n == 0) return DHFK_OK;
    if (!ang || !grot || !bone || !root) return fail(DHFK_E_INVAL, "ang/grot/bone/root must be non-null");
    if (as < 33 || gs < 3 || bs < 15 || rs < 3)
        return fail(DHFK_E_INVAL, "row strides must be >= 33 (ang), 3 (grot), 15 (bone), 3 (root)");
    return DHFK_OK;
}

}  // namespace

extern "C" {

int dhfk_abi_version(void) { return DHFK_ABI_VERSION; }
const char* dhfk_last_error(void) { return g_err; }
int dhfk_tile_rows(void) { return dhfk::kTile; }

int dhfk_topology(int32_t* parent33, int32_t* out16, float* alpha33, float* theta0_33, int32_t* len_kind33,
                  int32_t* len_bone33, int32_t* len_sign33, int32_t* h36m_32_to_16) {
    for (int j = 0; j < dhfk::NJ; ++j) {
        if (parent33) parent33[j] = dhfk::PARENT[j];
        if (alpha33) alpha33[j] = 90.0f * (float)dhfk::ALPHA_Q[j];
        if (theta0_33) theta0_33[j] = 90.0f * (float)dhfk::THETA0_Q[j];
        if (len_kind33) len_kind33[j] = dhfk::LEN_KIND[j];
        if (len_bone33) len_bone33[j] = dhfk::LEN_BONE[j];
        if (len_sign33) len_sign33[j] = dhfk::LEN_SIGN[j];
    }
    for (int k = 0; k < dhfk::NOUT; ++k) {
        if (out16) out16[k] = dhfk::OUT16[k];
        if (h36m_32_to_16) h36m_32_to_16[k] = dhfk::H36M_32_TO_16[k];
    }
    return DHFK_OK;
}

int dhfk_forward(const float* ang, int64_t ang_stride, const float* grot, int64_t grot_stride, const float* bone,
                 int64_t bone_stride, const float* root, int64_t root_stride, const float* cam, float* out_world,
                 float* out_cam, float* out_uv, int64_t n, uint32_t flags, void* stream) {
    int rc = check_inputs(ang, ang_stride, grot, grot_stride, bone, bone_stride, root, root_stride, n);
    if (rc != DHFK_OK) return rc;
    if (n == 0) return DHFK_OK;
    if (bad_trig_flags(flags)) return fail(DHFK_E_INVAL, "DHFK_FLAG_FAST_TRIG and DHFK_FLAG_ACCURATE_TRIG are mutually exclusive");
    if (!out_world) return fail(DHFK_E_INVAL, "out_world is required");
    if ((out_cam || out_uv) && !cam) return fail(DHFK_E_INVAL, "cam block required for out_cam / out_uv");
    if (!aligned16(out_world) || !aligned16(out_cam) || !aligned16(out_uv))
        return fail(DHFK_E_ALIGN, "outputs must be 16-byte aligned");
    FwdParams p;
    memset(&p, 0, sizeof p);
    p.ang = row_src(ang, ang_stride, 33);
    p.grot = row_src(grot, grot_stride, 3);
    p.bone = row_src(bone, bone_stride, 15);
    p.root = row_src(root, root_stride, 3);
    p.w = detect_wide(ang, ang_stride, grot, grot_stride, bone, bone_stride, root, root_stride, n);
    p.out_world = out_world;
    p.out_cam = out_cam;
    p.out_uv = out_uv;
    p.n = n;
    p.cam = make_cam(cam);
    if ((n + dhfk::kTile - 1) / dhfk::kTile > 2147483647LL) return fail(DHFK_E_INVAL, "n too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    const char* where = "";
    const bool oc = out_cam != nullptr, ou = out_uv != nullptr;
    int e;
    if (p.w.wide) e = (flags & DHFK_FLAG_FAST_TRIG) ? dhfk::launch_fwd_t1_g2(p, oc, ou, st, &where)
                                                    : dhfk::launch_fwd_t0_g2(p, oc, ou, st, &where);
    else e = (flags & DHFK_FLAG_FAST_TRIG) ? dhfk::launch_fwd_t1_g0(p, oc, ou, st, &where)
                                           : dhfk::launch_fwd_t0_g0(p, oc, ou, st, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}

int dhfk_backward(const float* ang, int64_t ang_stride, const float* grot, int64_t grot_stride, const float* bone,
                  int64_t bone_stride, const float* root, int64_t root_stride, const float* cam, const float* g_world,
                  const float* g_cam, const float* g_uv, float* g_ang, int64_t g_ang_stride, float* g_grot,
                  int64_t g_grot_stride, float* g_root, int64_t g_root_stride, float* g_bone, int64_t g_bone_stride,
                  int64_t n, uint32_t flags, void* stream) {
    int rc = check_inputs(ang, ang_stride, grot, grot_stride, bone, bone_stride, root, root_stride, n);
    if (rc != DHFK_OK) return rc;
    if (n == 0) return DHFK_OK;
    if (bad_trig_flags(flags)) return fail(DHFK_E_INVAL, "DHFK_FLAG_FAST_TRIG and DHFK_FLAG_ACCURATE_TRIG are mutually exclusive");
    if (!g_world && !g_cam && !g_uv) return fail(DHFK_E_INVAL, "at least one upstream gradient is required");
    if ((g_cam || g_uv) && !cam) return fail(DHFK_E_INVAL, "cam block required for g_cam / g_uv");
    if (!g_ang || !g_grot || !g_root) return fail(DHFK_E_INVAL, "g_ang, g_grot and g_root are required");
    if (g_ang_stride < 33 || g_grot_stride < 3 || g_root_stride < 3 || (g_bone && g_bone_stride < 15))
        return fail(DHFK_E_INVAL, "gradient row strides too small");
    if (!aligned16(g_world) || !aligned16(g_cam) || !aligned16(g_uv))
        return fail(DHFK_E_ALIGN, "upstream gradients must be 16-byte aligned");
    BwdParams p;
    memset(&p, 0, sizeof p);
    p.ang = row_src(ang, ang_stride, 33);
    p.grot = row_src(grot, grot_stride, 3);
    p.bone = row_src(bone, bone_stride, 15);
    p.root = row_src(root, root_stride, 3);
    p.g_world = g_world;
    p.g_cam = g_cam;
    p.g_uv = g_uv;
    p.g_ang = row_dst(g_ang, g_ang_stride, 33);
    p.g_grot = row_dst(g_grot, g_grot_stride, 3);
    p.g_root = row_dst(g_root, g_root_stride, 3);
    p.g_bone = row_dst(g_bone, g_bone_stride, 15);
    p.w = detect_wide(ang, ang_stride, grot, grot_stride, bone, bone_stride, root, root_stride, n);
    // the gradient of the wide tensor as one [N,S] tensor: same stride, d(global rotation) at the same column
    p.g_wide = (p.w.wide && g_ang_stride == p.w.wide && g_grot_stride == p.w.wide && g_grot == g_ang + p.w.goff &&
                aligned16(g_ang)) ? 1 : 0;
    p.n = n;
    p.cam = make_cam(cam);
    if ((n + dhfk::kTile - 1) / dhfk::kTile > 2147483647LL) return fail(DHFK_E_INVAL, "n too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    const bool gu = g_uv != nullptr;
    const bool fast = (flags & DHFK_FLAG_ACCURATE_TRIG) == 0;   // backward default: MUFU trig (see dhfk.h)
    const char* where = "";
    int e;
    if (g_bone) {     // no wide-row instantiation with bone gradients: the strided views take the gather path
        p.w = WideRows{0, 0, 1};
        p.g_wide = 0;
        e = fast ? dhfk::launch_bwd_t1_b1_g0(p, gu, st, &where) : dhfk::launch_bwd_t0_b1_g0(p, gu, st, &where);
    } else if (p.w.wide) {
        e = fast ? dhfk::launch_bwd_t1_b0_g2(p, gu, st, &where) : dhfk::launch_bwd_t0_b0_g2(p, gu, st, &where);
    } else {
        e = fast ? dhfk::launch_bwd_t1_b0_g0(p, gu, st, &where) : dhfk::launch_bwd_t0_b0_g0(p, gu, st, &where);
    }
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}

// ---- generator-epilogue mode -----------------------------------------------------
static int fill_gen_scale(dhfk::GenScale& gs, const float* half37, const float* mid37, float root_scale) {
    if (!half37 || !mid37) return fail(DHFK_E_INVAL, "gen_half37 / gen_mid37 (host, 37 floats each) are required");
    for (int i = 0; i < dhfk::GEN_NSLOT; ++i) gs.hm[i] = make_float2(half37[i], mid37[i]);
    gs.root_scale = root_scale;
    return DHFK_OK;
}

int dhfk_generator_forward(const float* net_out, int64_t net_out_stride, const float* bone, int64_t bone_stride,
                           const float* gen_half37, const float* gen_mid37, float root_scale, const float* cam,
                           float* out_world, float* out_cam, float* out_uv, int64_t n, uint32_t flags, void* stream) {
    if (n < 0) return fail(DHFK_E_INVAL, "n must be >= 0");
    if (n == 0) return DHFK_OK;
    if (bad_trig_flags(flags)) return fail(DHFK_E_INVAL, "DHFK_FLAG_FAST_TRIG and DHFK_FLAG_ACCURATE_TRIG are mutually exclusive");
    if (!net_out || !bone) return fail(DHFK_E_INVAL, "net_out / bone must be non-null");
    if (net_out_stride < dhfk::GEN_NCOL || bone_stride < 15)
        return fail(DHFK_E_INVAL, "row strides must be >= 35 (net_out), 15 (bone)");
    if (!out_world) return fail(DHFK_E_INVAL, "out_world is required");
    if ((out_cam || out_uv) && !cam) return fail(DHFK_E_INVAL, "cam block required for out_cam / out_uv");
    if (!aligned16(out_world) || !aligned16(out_cam) || !aligned16(out_uv))
        return fail(DHFK_E_ALIGN, "outputs must be 16-byte aligned");
    if ((n + dhfk::kTile - 1) / dhfk::kTile > 2147483647LL) return fail(DHFK_E_INVAL, "n too large for one launch");
    FwdParams p;
    memset(&p, 0, sizeof p);
    int rc = fill_gen_scale(p.gs, gen_half37, gen_mid37, root_scale);
    if (rc != DHFK_OK) return rc;
    p.ang = row_src(net_out, net_out_stride, dhfk::GEN_NCOL);
    p.bone = row_src(bone, bone_stride, 15);
    p.out_world = out_world; p.out_cam = out_cam; p.out_uv = out_uv;
    p.n = n;
    p.cam = make_cam(cam);
    const char* where = "";
    const bool oc = out_cam != nullptr, ou = out_uv != nullptr;
    int e = (flags & DHFK_FLAG_FAST_TRIG) ? dhfk::launch_fwd_t1_g1(p, oc, ou, (cudaStream_t)stream, &where)
                                          : dhfk::launch_fwd_t0_g1(p, oc, ou, (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}

int dhfk_generator_backward(const float* net_out, int64_t net_out_stride, const float* bone, int64_t bone_stride,
                            const float* gen_half37, const float* gen_mid37, float root_scale, const float* cam,
                            const float* g_world, const float* g_cam, const float* g_uv, float* g_net_out,
                            int64_t g_net_out_stride, int64_t n, uint32_t flags, void* stream) {
    if (n < 0) return fail(DHFK_E_INVAL, "n must be >= 0");
    if (n == 0) return DHFK_OK;
    if (bad_trig_flags(flags)) return fail(DHFK_E_INVAL, "DHFK_FLAG_FAST_TRIG and DHFK_FLAG_ACCURATE_TRIG are mutually exclusive");
    if (!net_out || !bone || !g_net_out) return fail(DHFK_E_INVAL, "net_out / bone / g_net_out must be non-null");
    if (net_out_stride < dhfk::GEN_NCOL || bone_stride < 15 || g_net_out_stride < dhfk::GEN_NCOL)
        return fail(DHFK_E_INVAL, "row strides must be >= 35 (net_out, g_net_out), 15 (bone)");
    if (!g_world && !g_cam && !g_uv) return fail(DHFK_E_INVAL, "at least one upstream gradient is required");
    if ((g_cam || g_uv) && !cam) return fail(DHFK_E_INVAL, "cam block required for g_cam / g_uv");
    if (!aligned16(g_world) || !aligned16(g_cam) || !aligned16(g_uv))
        return fail(DHFK_E_ALIGN, "upstream gradients must be 16-byte aligned");
    if ((n + dhfk::kTile - 1) / dhfk::kTile > 2147483647LL) return fail(DHFK_E_INVAL, "n too large for one launch");
    BwdParams p;
    memset(&p, 0, sizeof p);
    int rc = fill_gen_scale(p.gs, gen_half37, gen_mid37, root_scale);
    if (rc != DHFK_OK) return rc;
    p.ang = row_src(net_out, net_out_stride, dhfk::GEN_NCOL);
    p.bone = row_src(bone, bone_stride, 15);
    p.g_world = g_world; p.g_cam = g_cam; p.g_uv = g_uv;
    p.g_ang = row_dst(g_net_out, g_net_out_stride, dhfk::GEN_NCOL);
    p.n = n;
    p.cam = make_cam(cam);
    const char* where = "";
    const bool gu = g_uv != nullptr;
    int e = !(flags & DHFK_FLAG_ACCURATE_TRIG) ? dhfk::launch_bwd_t1_b0_g1(p, gu, (cudaStream_t)stream, &where)
                                          : dhfk::launch_bwd_t0_b0_g1(p, gu, (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}

int dhfk_world_to_camera_forward(const float* x, const float* cam_q, const float* cam_t, int32_t cam_on_device,
                                 float* out, int64_t num_points, void* stream) {
    if (num_points < 0) return fail(DHFK_E_INVAL, "num_points must be >= 0");
    if (num_points == 0) return DHFK_OK;
    if (!x || !cam_q || !cam_t || !out) return fail(DHFK_E_INVAL, "null argument");
    const bool tiles = num_points % 16 == 0 && aligned16(x) && aligned16(out);   // 16-joint poses: tiled kernel
    if (tiles) {
        dhfk::RotConst rc = {};
        if (!cam_on_device) {
            camera_matrix(cam_q, rc.M);
            for (int i = 0; i < 3; ++i) rc.t[i] = cam_t[i];
        }
        const char* where = "";
        int e = dhfk::launch_camera_tiles(0, x, nullptr, nullptr, 0, cam_on_device ? cam_q : nullptr,
                                          cam_on_device ? cam_t : nullptr, rc.M, rc.t, out, num_points / 16,
                                          (cudaStream_t)stream, &where);
        return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
    }
    if (cam_on_device) {
        long long nb = (num_points + 255) / 256;
        dhfk::w2c_fwd_dev_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(x, out, num_points, cam_q, cam_t);
        cudaError_t e2 = cudaGetLastError();
        return e2 == cudaSuccess ? DHFK_OK : cuda_fail(e2, "w2c_fwd_dev_kernel");
    }
    dhfk::RotConst rc;
    camera_matrix(cam_q, rc.M);
    for (int i = 0; i < 3; ++i) rc.t[i] = cam_t[i];
    long long blocks = (num_points + 255) / 256;
    dhfk::w2c_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, out, num_points, rc);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DHFK_OK : cuda_fail(e, "w2c_fwd_kernel");
}
int dhfk_world_to_camera_backward(const float* g_out, const float* cam_q, int32_t cam_on_device, float* g_x,
                                  int64_t num_points, void* stream) {
    if (num_points < 0) return fail(DHFK_E_INVAL, "num_points must be >= 0");
    if (num_points == 0) return DHFK_OK;
    if (!g_out || !cam_q || !g_x) return fail(DHFK_E_INVAL, "null argument");
    if (num_points % 16 == 0 && aligned16(g_out) && aligned16(g_x)) {
        dhfk::RotConst rc = {};
        if (!cam_on_device) camera_matrix(cam_q, rc.M);
        const char* where = "";
        int e = dhfk::launch_camera_tiles(1, g_out, nullptr, nullptr, 0, cam_on_device ? cam_q : nullptr, nullptr, rc.M,
                                          nullptr, g_x, num_points / 16, (cudaStream_t)stream, &where);
        return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
    }
    if (cam_on_device) {
        long long nb = (num_points + 255) / 256;
        dhfk::w2c_bwd_dev_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(g_out, g_x, num_points, cam_q);
        cudaError_t e2 = cudaGetLastError();
        return e2 == cudaSuccess ? DHFK_OK : cuda_fail(e2, "w2c_bwd_dev_kernel");
    }
    dhfk::RotConst rc;
    camera_matrix(cam_q, rc.M);
    rc.t[0] = rc.t[1] = rc.t[2] = 0.f;
    long long blocks = (num_points + 255) / 256;
    dhfk::w2c_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g_out, g_x, num_points, rc);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DHFK_OK : cuda_fail(e, "w2c_bwd_kernel");
}

int dhfk_project_forward(const float* x, const float* cam_rows, int64_t cam_rows_stride, float* uv, int64_t n,
                         int64_t joints, void* stream) {
    if (n < 0 || joints < 0) return fail(DHFK_E_INVAL, "n and joints must be >= 0");
    if (n == 0 || joints == 0) return DHFK_OK;
    if (!x || !cam_rows || !uv) return fail(DHFK_E_INVAL, "null argument");
    if (cam_rows_stride < 9) return fail(DHFK_E_INVAL, "camera rows need >= 9 columns");
    if ((reinterpret_cast<uintptr_t>(uv) & 7u) != 0) return fail(DHFK_E_ALIGN, "uv must be 8-byte aligned");
    if (joints == 16 && aligned16(x) && aligned16(uv) && n <= 2147483647LL * 32) {
        const char* where = "";
        int e = dhfk::launch_camera_tiles(2, x, nullptr, cam_rows, cam_rows_stride, nullptr, nullptr, nullptr, nullptr, uv,
                                          n, (cudaStream_t)stream, &where);
        return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
    }
    long long npts = n * joints, blocks = (npts + 255) / 256;
    dhfk::project_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, cam_rows, cam_rows_stride, uv,
                                                                                npts, (int)joints);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DHFK_OK : cuda_fail(e, "project_fwd_kernel");
}
int dhfk_project_backward(const float* x, const float* cam_rows, int64_t cam_rows_stride, const float* g_uv,
                          float* g_x, int64_t n, int64_t joints, void* stream) {
    if (n < 0 || joints < 0) return fail(DHFK_E_INVAL, "n and joints must be >= 0");
    if (n == 0 || joints == 0) return DHFK_OK;
    if (!x || !cam_rows || !g_uv || !g_x) return fail(DHFK_E_INVAL, "null argument");
    if (cam_rows_stride < 9) return fail(DHFK_E_INVAL, "camera rows need >= 9 columns");
    if ((reinterpret_cast<uintptr_t>(g_uv) & 7u) != 0) return fail(DHFK_E_ALIGN, "g_uv must be 8-byte aligned");
    if (joints == 16 && aligned16(x) && aligned16(g_uv) && aligned16(g_x) && n <= 2147483647LL * 32) {
        const char* where = "";
        int e = dhfk::launch_camera_tiles(3, x, g_uv, cam_rows, cam_rows_stride, nullptr, nullptr, nullptr, nullptr, g_x,
                                          n, (cudaStream_t)stream, &where);
        return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
    }
    long long npts = n * joints, blocks = (npts + 255) / 256;
    dhfk::project_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, cam_rows, cam_rows_stride, g_uv,
                                                                                g_x, npts, (int)joints);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DHFK_OK : cuda_fail(e, "project_bwd_kernel");
}

// ---- the reference's [N,32,3] output layout -----------------------------------------------------
int dhfk_scatter32_forward(const float* world16, const float* root, int64_t root_stride, float* world32, int64_t n,
                           void* stream) {
    if (n < 0) return fail(DHFK_E_INVAL, "n must be >= 0");
    if (n == 0) return DHFK_OK;
    if (!world16 || !root || !world32) return fail(DHFK_E_INVAL, "world16 / root / world32 must be non-null");
    if (root_stride < 3) return fail(DHFK_E_INVAL, "root row stride must be >= 3");
    if (!aligned16(world16) || !aligned16(world32)) return fail(DHFK_E_ALIGN, "world16 / world32 must be 16-byte aligned");
    if ((n + dhfk::kTile - 1) / dhfk::kTile > 2147483647LL) return fail(DHFK_E_INVAL, "n too large for one launch");
    const char* where = "";
    int e = dhfk::launch_scatter32(false, world16, root, root_stride, world32, nullptr, n, (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}
int dhfk_scatter32_backward(const float* g_world32, float* g_world16, float* g_root, int64_t n, void* stream) {
    if (n < 0) return fail(DHFK_E_INVAL, "n must be >= 0");
    if (n == 0) return DHFK_OK;
    if (!g_world32 || !g_world16 || !g_root) return fail(DHFK_E_INVAL, "g_world32 / g_world16 / g_root must be non-null");
    if (!aligned16(g_world32) || !aligned16(g_world16))
        return fail(DHFK_E_ALIGN, "g_world32 / g_world16 must be 16-byte aligned");
    if ((n + dhfk::kTile - 1) / dhfk::kTile > 2147483647LL) return fail(DHFK_E_INVAL, "n too large for one launch");
    const char* where = "";
    int e = dhfk::launch_scatter32(true, g_world32, nullptr, 0, g_world16, g_root, n, (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}

// ---- bone-length retarget + per-row projection ---------------------------------------
int dhfk_retarget_project(const float* pose, const int32_t* tmpl_idx, const float* templates, int32_t num_templates,
                          const float* cam_rows, int64_t cam_rows_stride, float* out_pose, float* out_uv, int64_t n,
                          void* stream) {
    if (n < 0) return fail(DHFK_E_INVAL, "n must be >= 0");
    if (n == 0) return DHFK_OK;
    if (!pose || !templates || !out_pose) return fail(DHFK_E_INVAL, "pose / templates / out_pose must be non-null");
    if (num_templates <= 0) return fail(DHFK_E_INVAL, "num_templates must be > 0");
    if (out_uv && (!cam_rows || (cam_rows_stride != 0 && cam_rows_stride < 9)))
        return fail(DHFK_E_INVAL, "out_uv needs cam_rows with >= 9 columns (row stride 0 = one shared row)");
    if (!aligned16(pose) || !aligned16(out_pose) || !aligned16(out_uv))
        return fail(DHFK_E_ALIGN, "pose / out_pose / out_uv must be 16-byte aligned");
    if ((n + dhfk::kTile - 1) / dhfk::kTile > 2147483647LL) return fail(DHFK_E_INVAL, "n too large for one launch");
    const char* where = "";
    int e = dhfk::launch_retarget(pose, tmpl_idx, templates, num_templates, cam_rows, cam_rows_stride, out_pose, out_uv,
                                  n, (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}

// ---- critic input transforms -----------------------------------------------------------
static int critic_common(int32_t kcs_cols, int64_t n, uint32_t flags) {
    if (n < 0) return fail(DHFK_E_INVAL, "n must be >= 0");
    if (kcs_cols != 0 && kcs_cols != 15 && kcs_cols != 30) return fail(DHFK_E_INVAL, "kcs_cols must be 0, 15 or 30");
    if (flags & ~(DHFK_CRITIC_CENTRE | DHFK_CRITIC_FLIP)) return fail(DHFK_E_INVAL, "unknown critic flag");
    if ((n + dhfk::kTile - 1) / dhfk::kTile > 2147483647LL) return fail(DHFK_E_INVAL, "n too large for one launch");
    return DHFK_OK;
}
int dhfk_critic_input_forward(const float* pose, float* out_pos, float* out_kcs, int32_t kcs_cols, int64_t n,
                              uint32_t flags, void* stream) {
    if (int rc = critic_common(kcs_cols, n, flags)) return rc;
    if (n == 0) return DHFK_OK;
    if (!pose) return fail(DHFK_E_INVAL, "pose must be non-null");
    if ((kcs_cols > 0) != (out_kcs != nullptr)) return fail(DHFK_E_INVAL, "out_kcs must be given iff kcs_cols > 0");
    if (!out_pos && !out_kcs) return fail(DHFK_E_INVAL, "nothing to compute: out_pos and out_kcs are both null");
    if (!aligned16(pose) || !aligned16(out_pos) || !aligned16(out_kcs))
        return fail(DHFK_E_ALIGN, "pose / out_pos / out_kcs must be 16-byte aligned");
    const char* where = "";
    int e = dhfk::launch_critic(0, kcs_cols, out_pos != nullptr, pose, nullptr, nullptr, out_pos, out_kcs, n, flags,
                                (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}
int dhfk_critic_input_backward(const float* pose, const float* g_pos, const float* g_kcs, int32_t kcs_cols,
                               float* g_pose, int64_t n, uint32_t flags, void* stream) {
    if (int rc = critic_common(kcs_cols, n, flags)) return rc;
    if (n == 0) return DHFK_OK;
    if (!pose || !g_pose) return fail(DHFK_E_INVAL, "pose / g_pose must be non-null");
    if ((kcs_cols > 0) != (g_kcs != nullptr)) return fail(DHFK_E_INVAL, "g_kcs must be given iff kcs_cols > 0");
    if (!g_pos && !g_kcs) return fail(DHFK_E_INVAL, "no upstream gradient: g_pos and g_kcs are both null");
    if (!aligned16(pose) || !aligned16(g_pos) || !aligned16(g_kcs) || !aligned16(g_pose))
        return fail(DHFK_E_ALIGN, "pose / g_pos / g_kcs / g_pose must be 16-byte aligned");
    const char* where = "";
    int e = dhfk::launch_critic(1, kcs_cols, g_pos != nullptr, pose, g_pos, g_kcs, g_pose, nullptr, n, flags,
                                (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}
int dhfk_critic_input_jvp(const float* pose, const float* v_pose, float* t_pos, float* t_kcs, int32_t kcs_cols,
                          int64_t n, uint32_t flags, void* stream) {
    if (int rc = critic_common(kcs_cols, n, flags)) return rc;
    if (n == 0) return DHFK_OK;
    if (!pose || !v_pose) return fail(DHFK_E_INVAL, "pose / v_pose must be non-null");
    if ((kcs_cols > 0) != (t_kcs != nullptr)) return fail(DHFK_E_INVAL, "t_kcs must be given iff kcs_cols > 0");
    if (!t_pos && !t_kcs) return fail(DHFK_E_INVAL, "nothing to compute: t_pos and t_kcs are both null");
    if (!aligned16(pose) || !aligned16(v_pose) || !aligned16(t_pos) || !aligned16(t_kcs))
        return fail(DHFK_E_ALIGN, "pose / v_pose / t_pos / t_kcs must be 16-byte aligned");
    const char* where = "";
    int e = dhfk::launch_critic(2, kcs_cols, t_pos != nullptr, pose, v_pose, nullptr, t_pos, t_kcs, n, flags,
                                (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}
int dhfk_flip_pose(const float* x, float* out, int64_t n, int32_t dims, void* stream) {
    if (n < 0) return fail(DHFK_E_INVAL, "n must be >= 0");
    if (dims != 2 && dims != 3) return fail(DHFK_E_INVAL, "dims must be 2 or 3");
    if (n == 0) return DHFK_OK;
    if (!x || !out) return fail(DHFK_E_INVAL, "x / out must be non-null");
    if ((n + dhfk::kTile - 1) / dhfk::kTile > 2147483647LL) return fail(DHFK_E_INVAL, "n too large for one launch");
    if (!aligned16(x) || !aligned16(out)) return fail(DHFK_E_ALIGN, "x / out must be 16-byte aligned");
    const char* where = "";
    int e = dhfk::launch_flip(x, out, n, dims, (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}

// ---- SURVEY 8 f2, video part: motion-critic inputs ---------------------------------------------------
static int video_common(int32_t frames, uint32_t flags, int64_t n) {
    if (n < 0) return fail(DHFK_E_INVAL, "n_rows must be >= 0");
    if (frames < 1) return fail(DHFK_E_INVAL, "frames must be >= 1");
    if (n % frames != 0) return fail(DHFK_E_INVAL, "n_rows must be a multiple of frames");
    if (flags & ~DHFK_VIDEO_REVERSE) return fail(DHFK_E_INVAL, "unknown video flag");
    if (n > 2147483647LL) return fail(DHFK_E_INVAL, "n_rows too large for one launch (< 2^31)");
    return DHFK_OK;
}
int dhfk_video_critic_forward(const float* pose, int32_t frames, uint32_t flags, float* out_kcs, float* out_dkcs,
                              float* out_dpos, float* out_pos, int64_t n, void* stream) {
    if (int rc = video_common(frames, flags, n)) return rc;
    if (n == 0) return DHFK_OK;
    if (!pose || !out_kcs) return fail(DHFK_E_INVAL, "pose / out_kcs must be non-null");
    if (frames > 1 && !out_dkcs) return fail(DHFK_E_INVAL, "out_dkcs must be non-null when frames > 1");
    if (!aligned16(pose) || !aligned16(out_dpos) || !aligned16(out_pos))
        return fail(DHFK_E_ALIGN, "pose / out_dpos / out_pos must be 16-byte aligned");
    const char* where = "";
    int e = dhfk::launch_video_critic(0, pose, nullptr, frames, flags, out_kcs, out_dkcs, frames > 1 ? out_dpos : nullptr,
                                      out_pos, n, (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}
int dhfk_video_critic_jvp(const float* pose, const float* v_pose, int32_t frames, uint32_t flags, float* t_kcs,
                          float* t_dkcs, float* t_dpos, float* t_pos, int64_t n, void* stream) {
    if (int rc = video_common(frames, flags, n)) return rc;
    if (n == 0) return DHFK_OK;
    if (!pose || !v_pose || !t_kcs) return fail(DHFK_E_INVAL, "pose / v_pose / t_kcs must be non-null");
    if (frames > 1 && !t_dkcs) return fail(DHFK_E_INVAL, "t_dkcs must be non-null when frames > 1");
    if (!aligned16(pose) || !aligned16(v_pose) || !aligned16(t_dpos) || !aligned16(t_pos))
        return fail(DHFK_E_ALIGN, "pose / v_pose / t_dpos / t_pos must be 16-byte aligned");
    const char* where = "";
    int e = dhfk::launch_video_critic(2, pose, v_pose, frames, flags, t_kcs, t_dkcs, frames > 1 ? t_dpos : nullptr, t_pos,
                                      n, (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}
int dhfk_video_critic_backward(const float* pose, int32_t frames, uint32_t flags, const float* g_kcs, const float* g_dkcs,
                               const float* g_dpos, const float* g_pos, float* g_pose, int64_t n, void* stream) {
    if (int rc = video_common(frames, flags, n)) return rc;
    if (n == 0) return DHFK_OK;
    if (!pose || !g_pose) return fail(DHFK_E_INVAL, "pose / g_pose must be non-null");
    if (frames == 1) { g_dkcs = nullptr; g_dpos = nullptr; }
    if (!g_kcs && !g_dkcs && !g_dpos && !g_pos) return fail(DHFK_E_INVAL, "at least one upstream gradient is required");
    if (!aligned16(pose) || !aligned16(g_dpos) || !aligned16(g_pos) || !aligned16(g_pose))
        return fail(DHFK_E_ALIGN, "pose / g_dpos / g_pos / g_pose must be 16-byte aligned");
    const char* where = "";
    int e = dhfk::launch_video_critic_bwd(pose, frames, flags, g_kcs, g_dkcs, g_dpos, g_pos, g_pose, n,
                                          (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}
int dhfk_video_root_diff_forward(const float* uv, int32_t frames, uint32_t flags, float* out_diff, float* out_uv,
                                 int64_t n, void* stream) {
    if (int rc = video_common(frames, flags, n)) return rc;
    if (n == 0) return DHFK_OK;
    if (!uv) return fail(DHFK_E_INVAL, "uv must be non-null");
    if (frames > 1 && !out_diff) return fail(DHFK_E_INVAL, "out_diff must be non-null when frames > 1");
    if (!aligned16(uv) || !aligned16(out_uv) || (reinterpret_cast<uintptr_t>(out_diff) & 7u))
        return fail(DHFK_E_ALIGN, "uv / out_uv must be 16-byte aligned, out_diff 8-byte aligned");
    if (frames == 1 && !out_uv) return DHFK_OK;
    const char* where = "";
    int e = dhfk::launch_video_root_diff(false, uv, nullptr, nullptr, frames, flags, out_diff, out_uv, nullptr, n,
                                         (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}
int dhfk_video_root_diff_backward(const float* g_diff, const float* g_uv_playback, int32_t frames, uint32_t flags,
                                  float* g_uv, int64_t n, void* stream) {
    if (int rc = video_common(frames, flags, n)) return rc;
    if (n == 0) return DHFK_OK;
    if (!g_uv) return fail(DHFK_E_INVAL, "g_uv must be non-null");
    if (frames == 1) g_diff = nullptr;
    if (!g_diff && !g_uv_playback && frames > 1) return fail(DHFK_E_INVAL, "at least one upstream gradient is required");
    if (!aligned16(g_uv) || !aligned16(g_uv_playback) || (reinterpret_cast<uintptr_t>(g_diff) & 7u))
        return fail(DHFK_E_ALIGN, "g_uv / g_uv_playback must be 16-byte aligned, g_diff 8-byte aligned");
    const char* where = "";
    int e = dhfk::launch_video_root_diff(true, nullptr, g_diff, g_uv_playback, frames, flags, nullptr, nullptr, g_uv, n,
                                         (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}

// ---- device-resident fake-pair bank ----------------------------------------------------
int dhfk_bank_gather(const float* bank, int64_t rec_floats, int32_t cam_cols, const int64_t* idx, int64_t nb,
                     int64_t bank_rows, float* out3d, float* out2d, float* out_cam, void* stream) {
    if (nb < 0 || bank_rows < 0) return fail(DHFK_E_INVAL, "nb / bank_rows must be >= 0");
    if (nb == 0) return DHFK_OK;
    if (!bank || !idx || !out3d || !out2d) return fail(DHFK_E_INVAL, "bank / idx / out3d / out2d must be non-null");
    if (out_cam && (cam_cols < 1 || cam_cols > 32)) return fail(DHFK_E_INVAL, "cam_cols must be in 1..32");
    const int64_t need = 80 + (out_cam ? (cam_cols + 3) / 4 * 4 : 0);
    if (rec_floats < need || rec_floats % 4 != 0 || rec_floats > (1 << 20))
        return fail(DHFK_E_INVAL, "rec_floats must be a multiple of 4 and >= 80 + cam_cols rounded up to 4");
    if (!aligned16(bank) || !aligned16(out3d) || !aligned16(out2d))
        return fail(DHFK_E_ALIGN, "bank / out3d / out2d must be 16-byte aligned");
    if (nb > (1LL << 33)) return fail(DHFK_E_INVAL, "nb too large for one launch");
    const char* where = "";
    int e = dhfk::launch_bank_gather(bank, rec_floats, cam_cols, reinterpret_cast<const long long*>(idx), nb, bank_rows,
                                     out3d, out2d, out_cam, (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}

// ---- SURVEY 8 e: gradient exchange over peer memory ---------------------------------------------------
int dhfk_grad_allreduce(float* const* peer_bufs, float* multicast_buf, uint32_t* const* peer_flags, uint32_t* status_dev,
                        int32_t rank, int32_t world, int64_t n_floats, float scale, int32_t max_ctas, int32_t cta_threads,
                        int64_t timeout_ms, void* stream) {
    if (world < 1 || world > DHFK_AR_MAX_WORLD || rank < 0 || rank >= world)
        return fail(DHFK_E_INVAL, "need 0 <= rank < world <= DHFK_AR_MAX_WORLD");
    if (n_floats < 0 || n_floats % 4 != 0) return fail(DHFK_E_INVAL, "n_floats must be a non-negative multiple of 4");
    if (n_floats == 0) return DHFK_OK;
    if (!peer_bufs || !peer_flags || !status_dev) return fail(DHFK_E_INVAL, "peer_bufs / peer_flags / status_dev must be non-null");
    if (max_ctas < 1 || max_ctas > DHFK_AR_MAX_CTAS) return fail(DHFK_E_INVAL, "max_ctas must be in 1..DHFK_AR_MAX_CTAS");
    if (cta_threads < 32 || cta_threads > 512 || cta_threads % 32 != 0)
        return fail(DHFK_E_INVAL, "cta_threads must be a multiple of 32 in 32..512");
    if (timeout_ms <= 0) return fail(DHFK_E_INVAL, "timeout_ms must be positive");
    for (int r = 0; r < world; ++r) {
        if (!peer_bufs[r] || !peer_flags[r]) return fail(DHFK_E_INVAL, "null peer buffer / flag block");
        if (!aligned16(peer_bufs[r])) return fail(DHFK_E_ALIGN, "peer buffers must be 16-byte aligned");
    }
    if (multicast_buf && !aligned16(multicast_buf)) return fail(DHFK_E_ALIGN, "multicast_buf must be 16-byte aligned");
    const char* where = "";
    int e = dhfk::launch_grad_allreduce(peer_bufs, multicast_buf, reinterpret_cast<unsigned* const*>(peer_flags),
                                        reinterpret_cast<unsigned*>(status_dev), rank, world, n_floats, scale,
                                        max_ctas, cta_threads, (unsigned long long)timeout_ms * 1000000ull, (cudaStream_t)stream, &where);
    return e == 0 ? DHFK_OK : cuda_fail((cudaError_t)e, where);
}

// ---- host-buffer end-to-end entry --------------------------------------------------------------
// per-row device scratch: inputs 54, world 48, uv 32, g_world 48, g_uv 32, g_ang 33, g_grot 3, g_root 3
static const int64_t kHostRowFloats = 54 + 48 + 32 + 48 + 32 + 33 + 3 + 3;

int64_t dhfk_host_workspace_bytes(int64_t chunk_rows, int32_t num_streams) {
    if (chunk_rows <= 0 || num_streams <= 0) return 0;
    int64_t rows = (chunk_rows + 3) / 4 * 4;  // keep every sub-buffer 16-byte aligned
    return rows * kHostRowFloats * (int64_t)sizeof(float) * num_streams;
}

int dhfk_forward_backward_host(const float* ang_h, const float* grot_h, const float* bone_h, const float* root_h,
                               const float* cam, const float* g_world_h, const float* g_uv_h, float* out_world_h,
                               float* out_uv_h, float* g_ang_h, float* g_grot_h, float* g_root_h, int64_t n,
                               int64_t chunk_rows, int32_t num_streams, void* workspace, int64_t workspace_bytes,
                               uint32_t flags) {
    if (n < 0) return fail(DHFK_E_INVAL, "n must be >= 0");
    if (n == 0) return DHFK_OK;
    if (bad_trig_flags(flags)) return fail(DHFK_E_INVAL, "DHFK_FLAG_FAST_TRIG and DHFK_FLAG_ACCURATE_TRIG are mutually exclusive");
    if (!ang_h || !grot_h || !bone_h || !root_h || !cam || !out_world_h || !out_uv_h)
        return fail(DHFK_E_INVAL, "null host argument");
    const bool do_bwd = g_world_h != nullptr || g_uv_h != nullptr;
    if (do_bwd && (!g_world_h || !g_uv_h || !g_ang_h || !g_grot_h || !g_root_h))
        return fail(DHFK_E_INVAL, "backward needs g_world, g_uv and the three gradient outputs");
    if (chunk_rows <= 0 || chunk_rows % 4 != 0) return fail(DHFK_E_INVAL, "chunk_rows must be a positive multiple of 4");
    if (num_streams <= 0 || num_streams > 8) return fail(DHFK_E_INVAL, "num_streams must be in 1..8");
    if (!workspace || workspace_bytes < dhfk_host_workspace_bytes(chunk_rows, num_streams))
        return fail(DHFK_E_INVAL, "workspace too small (see dhfk_host_workspace_bytes)");
    if (!aligned16(workspace)) return fail(DHFK_E_ALIGN, "workspace must be 16-byte aligned");

    // Three-stage pipeline over a ring of `num_streams` device slots: one stream only uploads, one only computes,
    // one only downloads, chained by events.  Each chunk moves in two half-steps -- inputs up, forward, world/uv
    // down; upstream gradients up, backward, gradients down -- so the first download starts after 216 B/row have
    // arrived and the last one is only 156 B/row (short pipeline fill and drain), and each DMA engine sees one
    // in-order queue of copies that are ready when they reach its head.
    const int slots = num_streams;
    cudaStream_t s_up = nullptr, s_run = nullptr, s_down = nullptr;
    enum { kUpIn, kUpGrad, kFwd, kBwd, kDown, kNumEv };
    cudaEvent_t ev[8][kNumEv];
    int n_ev = 0;
    int rc = DHFK_OK;
    cudaError_t e = cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking);
    for (; e == cudaSuccess && n_ev < slots * kNumEv; ++n_ev)
        e = cudaEventCreateWithFlags(&ev[n_ev / kNumEv][n_ev % kNumEv], cudaEventDisableTiming);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamCreate / cudaEventCreate");

    const size_t F = sizeof(float);
    const int64_t slot_floats = chunk_rows * kHostRowFloats;
    int64_t chunk_idx = 0;
#define DHFK_CP(dst, src, cnt, kind, st)                                                  \
    if (rc == DHFK_OK) {                                                                  \
        e = cudaMemcpyAsync(dst, src, (size_t)(cnt)*F, kind, st);                         \
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync");                       \
    }
#define DHFK_EV(call)                                                                     \
    if (rc == DHFK_OK) {                                                                  \
        e = call;                                                                         \
        if (e != cudaSuccess) rc = cuda_fail(e, #call);                                   \
    }
    for (int64_t r0 = 0; r0 < n && rc == DHFK_OK; r0 += chunk_rows, ++chunk_idx) {
        const int64_t rows = (n - r0 < chunk_rows) ? (n - r0) : chunk_rows;
        const int slot = (int)(chunk_idx % slots);
        cudaEvent_t* E = ev[slot];
        float* base = reinterpret_cast<float*>(workspace) + slot * slot_floats;
        float* d_ang = base;
        float* d_grot = d_ang + chunk_rows * 33;
        float* d_bone = d_grot + chunk_rows * 3;
        float* d_root = d_bone + chunk_rows * 15;
        float* d_world = d_root + chunk_rows * 3;
        float* d_uv = d_world + chunk_rows * 48;
        float* d_gw = d_uv + chunk_rows * 32;
        float* d_gu = d_gw + chunk_rows * 48;
        float* d_gang = d_gu + chunk_rows * 32;
        float* d_ggrot = d_gang + chunk_rows * 33;
        float* d_groot = d_ggrot + chunk_rows * 3;
        // the slot is free once the last download of the chunk that used it before has finished
        if (chunk_idx >= slots) DHFK_EV(cudaStreamWaitEvent(s_up, E[kDown], 0))
        DHFK_CP(d_ang, ang_h + r0 * 33, rows * 33, cudaMemcpyHostToDevice, s_up)
        DHFK_CP(d_grot, grot_h + r0 * 3, rows * 3, cudaMemcpyHostToDevice, s_up)
        DHFK_CP(d_bone, bone_h + r0 * 15, rows * 15, cudaMemcpyHostToDevice, s_up)
        DHFK_CP(d_root, root_h + r0 * 3, rows * 3, cudaMemcpyHostToDevice, s_up)
        DHFK_EV(cudaEventRecord(E[kUpIn], s_up))
        if (do_bwd) {
            DHFK_CP(d_gw, g_world_h + r0 * 48, rows * 48, cudaMemcpyHostToDevice, s_up)
            DHFK_CP(d_gu, g_uv_h + r0 * 32, rows * 32, cudaMemcpyHostToDevice, s_up)
            DHFK_EV(cudaEventRecord(E[kUpGrad], s_up))
        }
        DHFK_EV(cudaStreamWaitEvent(s_run, E[kUpIn], 0))
        if (rc == DHFK_OK)
            rc = dhfk_forward(d_ang, 33, d_grot, 3, d_bone, 15, d_root, 3, cam, d_world, nullptr, d_uv,
                              rows, flags, s_run);
        DHFK_EV(cudaEventRecord(E[kFwd], s_run))
        if (do_bwd) {
            DHFK_EV(cudaStreamWaitEvent(s_run, E[kUpGrad], 0))
            if (rc == DHFK_OK)
                rc = dhfk_backward(d_ang, 33, d_grot, 3, d_bone, 15, d_root, 3, cam, d_gw, nullptr, d_gu,
                                   d_gang, 33, d_ggrot, 3, d_groot, 3, nullptr, 0, rows, flags, s_run);
            DHFK_EV(cudaEventRecord(E[kBwd], s_run))
        }
        DHFK_EV(cudaStreamWaitEvent(s_down, E[kFwd], 0))
        DHFK_CP(out_world_h + r0 * 48, d_world, rows * 48, cudaMemcpyDeviceToHost, s_down)
        DHFK_CP(out_uv_h + r0 * 32, d_uv, rows * 32, cudaMemcpyDeviceToHost, s_down)
        if (do_bwd) {
            DHFK_EV(cudaStreamWaitEvent(s_down, E[kBwd], 0))
            DHFK_CP(g_ang_h + r0 * 33, d_gang, rows * 33, cudaMemcpyDeviceToHost, s_down)
            DHFK_CP(g_grot_h + r0 * 3, d_ggrot, rows * 3, cudaMemcpyDeviceToHost, s_down)
            DHFK_CP(g_root_h + r0 * 3, d_groot, rows * 3, cudaMemcpyDeviceToHost, s_down)
        }
        DHFK_EV(cudaEventRecord(E[kDown], s_down))
    }
#undef DHFK_CP
#undef DHFK_EV
    for (cudaStream_t st : {s_up, s_run, s_down}) {
        if (!st) continue;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess && rc == DHFK_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
        cudaStreamDestroy(st);
    }
    for (int i = 0; i < n_ev; ++i) cudaEventDestroy(ev[i / kNumEv][i % kNumEv]);
    return rc;
}

}  // extern "C"
