// dhfk_device.cuh -- per-pose device math: degree-argument sincos, DH joint steps as signed
// axis permutation + planar rotation, pinhole projection forward/backward, and the
// compile-time tree walkers (forward: emit origins; backward: wrench accumulation).
//
// One thread owns one pose; everything here is straight-line register code after inlining.
#pragma once
#include "dhfk_topology.h"
#include "dhfk_sincos_table.h"

namespace dhfk {

struct V3 { float x, y, z; };
struct Frame { V3 X, Y, Z, O; };   // columns of the 3x4 world transform of a joint frame
struct Wrench { V3 F, M; };        // sum of forces / moments about the chain origin

#define DHFK_DI __device__ __forceinline__

DHFK_DI V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
// Packed fp32x2 arithmetic for the (x, y) lanes of a 3-vector (sm_100 FFMA2 / FMUL2 / FADD2: one instruction for both
// lanes, IEEE per lane -- bit-identical to the scalar forms below, operation for operation).  The tree walk is
// element-wise 3-vector arithmetic, and these kernels are issue / power-bound: 3 scalar instructions become 2.
#ifndef DHFK_PACKED_V3
#define DHFK_PACKED_V3 1
#endif
#if DHFK_PACKED_V3
DHFK_DI float2 xy_of(V3 a) { return make_float2(a.x, a.y); }
DHFK_DI V3 v3p(float2 p, float z) { return v3(p.x, p.y, z); }
DHFK_DI V3 operator+(V3 a, V3 b) { return v3p(__fadd2_rn(xy_of(a), xy_of(b)), a.z + b.z); }
DHFK_DI V3 operator-(V3 a, V3 b) {      // fma(b, -1, a) = a - b exactly
    return v3p(__ffma2_rn(xy_of(b), make_float2(-1.f, -1.f), xy_of(a)), a.z - b.z);
}
DHFK_DI V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
DHFK_DI V3 axpy(float s, V3 a, V3 b) { return v3p(__ffma2_rn(make_float2(s, s), xy_of(a), xy_of(b)), fmaf(s, a.z, b.z)); }
DHFK_DI V3 scale(float s, V3 a) { return v3p(__fmul2_rn(make_float2(s, s), xy_of(a)), s * a.z); }
#else
DHFK_DI V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
DHFK_DI V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
DHFK_DI V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
DHFK_DI V3 axpy(float s, V3 a, V3 b) { return v3(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
DHFK_DI V3 scale(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
#endif
DHFK_DI float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
DHFK_DI V3 cross(V3 a, V3 b) {
    return v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
// a - o x f
DHFK_DI V3 sub_cross(V3 a, V3 o, V3 f) {
    return v3(fmaf(o.z, f.y, fmaf(-o.y, f.z, a.x)), fmaf(o.x, f.z, fmaf(-o.z, f.x, a.y)),
              fmaf(o.y, f.x, fmaf(-o.x, f.y, a.z)));
}
// a + o x f
DHFK_DI V3 add_cross(V3 a, V3 o, V3 f) {
    return v3(fmaf(o.y, f.z, fmaf(-o.z, f.y, a.x)), fmaf(o.z, f.x, fmaf(-o.x, f.z, a.y)),
              fmaf(o.x, f.y, fmaf(-o.y, f.x, a.z)));
}
// 3x3 row-major matrix (array of 9) times vector, and transpose times vector
DHFK_DI V3 mat_vec(const float* m, V3 v) {
    return v3(fmaf(m[0], v.x, fmaf(m[1], v.y, m[2] * v.z)), fmaf(m[3], v.x, fmaf(m[4], v.y, m[5] * v.z)),
              fmaf(m[6], v.x, fmaf(m[7], v.y, m[8] * v.z)));
}
DHFK_DI V3 mat_vec_add(const float* m, V3 v, V3 b) {
    return v3(fmaf(m[0], v.x, fmaf(m[1], v.y, fmaf(m[2], v.z, b.x))),
              fmaf(m[3], v.x, fmaf(m[4], v.y, fmaf(m[5], v.z, b.y))),
              fmaf(m[6], v.x, fmaf(m[7], v.y, fmaf(m[8], v.z, b.z))));
}
DHFK_DI V3 matT_vec(const float* m, V3 v) {
    return v3(fmaf(m[0], v.x, fmaf(m[3], v.y, m[6] * v.z)), fmaf(m[1], v.x, fmaf(m[4], v.y, m[7] * v.z)),
              fmaf(m[2], v.x, fmaf(m[5], v.y, m[8] * v.z)));
}
DHFK_DI V3 matT_vec_add(const float* m, V3 v, V3 b) {
    return v3(fmaf(m[0], v.x, fmaf(m[3], v.y, fmaf(m[6], v.z, b.x))),
              fmaf(m[1], v.x, fmaf(m[4], v.y, fmaf(m[7], v.z, b.y))),
              fmaf(m[2], v.x, fmaf(m[5], v.y, fmaf(m[8], v.z, b.z))));
}

constexpr float kDegToRad = 0.017453292519943295f;

// ---------------------------------------------------------------------------------------
// sin/cos of (deg + 90*Q0) degrees.  Q0 is the compile-time theta0 quadrant of the joint.
//
// TRIG_ACCURATE: exact range reduction in degrees (r = deg - 90*rint(deg/90) is exact in
// fp32), degree-scaled minimax polynomials on |r| <= 45 (cephes sinf/cosf coefficients),
// quadrant fix-up by swap + sign-bit xor.  ~1 ulp; independent of |deg| up to ~3e8.
// TRIG_MUFU: exact reduction to |r| <= 180, then MUFU.SIN / MUFU.COS (abs err ~4e-7).
// The reference computes fl(fl(deg/180)*pi_f32) and then libm sinf: its own argument error
// is ~1e-7 relative, so either variant is at least as close to the true value as it is.
// ---------------------------------------------------------------------------------------
enum { TRIG_ACCURATE = 0, TRIG_MUFU = 1 };

// TRIG_ACCURATE implementation: 1 = table + remainder (default), 0 = polynomial on |r| <= 45 deg.
//   table: deg = k*2.8125 + d exactly (k = rint(deg*128/360), |d| <= 1.40625); {sin,cos}(k*2.8125) come from
//   a 128-entry fp32 table (1 KB, L1-resident, LDG.64), sin d = d*(D + S1 d^2), cos d = 1 - D^2 d^2/2, then
//   the angle-sum formulas.  Max abs error 1.0e-7 (fp32 emulation over [-720,720] deg; polynomial path
//   7.7e-8; the reference's own fl(fl(deg/180)*pi) argument rounding costs it 1.1e-6) at ~15 instead of ~23
//   instructions per sincos.  theta0 quadrants fold into the table index.
#ifndef DHFK_ACCURATE_TABLE
#define DHFK_ACCURATE_TABLE 1
#endif
// TRIG_MUFU: 1 (default) = reduce the angle exactly to |r| <= 180 deg before the two multiplies in front of MUFU.SIN /
// MUFU.COS (deg -> rad, rad -> turns), 0 = hand deg * pi/180 to them as it is.  The unit works on the fractional turn
// either way; what the reduction buys is that the two multiply roundings (1.8e-7 relative, together) act on |r| <= 180
// instead of on |deg|.  Measured and rejected (profiles/r2i1_ab_mufu_direct.txt): without it a sincos is 4 instead of 7
// instructions and the backward runs 0.5 % faster (generator mode 0.2120 -> 0.2096 ms), but the stress parity test
// (angles up to +-360 deg after theta0, clamp active) fails its 1e-5 bound.
#ifndef DHFK_MUFU_EXACT_REDUCTION
#define DHFK_MUFU_EXACT_REDUCTION 1
#endif
// TRIG_ACCURATE polynomial path: 1 = reduce by half turns (sign only), 0 = by quarter turns (swap + sign)
#ifndef DHFK_SINCOS_HALFTURN
#define DHFK_SINCOS_HALFTURN 1
#endif
static __device__ const float2 c_sincos_table[DHFK_SINCOS_TABLE_SIZE] = {DHFK_SINCOS_TABLE_VALUES};
static_assert(DHFK_SINCOS_TABLE_SIZE == 128, "index arithmetic below assumes 128 entries");

DHFK_DI void sincos_table_core(float deg, int q0, float& s, float& c) {
    const float kMagic = 12582912.0f;
    float t = fmaf(deg, 128.0f / 360.0f, kMagic);
    int idx = (__float_as_int(t) + 32 * q0) & 127;
    float kf = t - kMagic;
    float d = fmaf(kf, -2.8125f, deg);            // exact remainder in degrees
    float2 sc = __ldg(&c_sincos_table[idx]);
    constexpr double D = 3.14159265358979323846 / 180.0;
    float d2 = d * d;
    float sd = d * fmaf(d2, (float)(-D * D * D / 6.0), (float)D);
    float cd = fmaf(d2, (float)(-0.5 * D * D), 1.0f);
    s = fmaf(sc.x, cd, sc.y * sd);
    c = fmaf(sc.y, cd, -(sc.x * sd));
}

template <int TRIG, int Q0>
DHFK_DI void sincos_deg(float deg, float& s, float& c) {
    const float kMagic = 12582912.0f;  // 1.5 * 2^23: (x + kMagic) - kMagic == rint(x) for |x| < 2^22
    if (TRIG == TRIG_ACCURATE && DHFK_ACCURATE_TABLE) {
        sincos_table_core(deg, Q0, s, c);
    } else if (TRIG == TRIG_ACCURATE && DHFK_SINCOS_HALFTURN) {
        // Reduce by half turns: r = deg - 180*rint(deg/180) is exact, |r| <= 90, and sin / cos of the angle are those of
        // r up to ONE common sign (the parity of the half-turn count) -- no swap, no per-quadrant selects.  The price is
        // one more term per polynomial (least-squares fit on Chebyshev nodes over |r| <= 90 deg, absolute-error weighted,
        // tools/sincos_reduction_study.py): max abs error 1.23e-7 / 1.11e-7 (sin / cos, fp32 emulation over [-720, 720]
        // deg) against 8.0e-8 for the quarter-turn scheme below -- the reference's own fl(fl(deg/180)*pi) argument
        // rounding costs it 1.1e-6 -- for 13 instead of 18 instructions per sincos.  The compile-time theta0 quadrant
        // folds into which of the two values is the sine and into the constant part of the sign mask.
        float t = fmaf(deg, 1.0f / 180.0f, kMagic);
        const unsigned sign = (unsigned)__float_as_int(t) << 31;
        float q = t - kMagic;
        float r = fmaf(q, -180.0f, deg);
        float r2 = r * r;
        constexpr double D = 3.14159265358979323846 / 180.0;
        constexpr float S0 = (float)(9.99999976513755717e-01 * D);
        constexpr float S1 = (float)(-1.66666475934896696e-01 * D * D * D);
        constexpr float S2 = (float)(8.33289922283364168e-03 * D * D * D * D * D);
        constexpr float S3 = (float)(-1.98008653071941475e-04 * D * D * D * D * D * D * D);
        constexpr float S4 = (float)(2.59043003061907403e-06 * D * D * D * D * D * D * D * D * D);
        constexpr float C0 = (float)(-4.99999995353595350e-01 * D * D);
        constexpr float C1 = (float)(4.16666402579548498e-02 * D * D * D * D);
        constexpr float C2 = (float)(-1.38883983513082084e-03 * D * D * D * D * D * D);
        constexpr float C3 = (float)(2.47616556899794226e-05 * D * D * D * D * D * D * D * D);
        constexpr float C4 = (float)(-2.60734800845432612e-07 * D * D * D * D * D * D * D * D * D * D);
#if DHFK_PACKED_V3
        float2 pp = __ffma2_rn(make_float2(r2, r2), make_float2(S4, C4), make_float2(S3, C3));
        pp = __ffma2_rn(make_float2(r2, r2), pp, make_float2(S2, C2));
        pp = __ffma2_rn(make_float2(r2, r2), pp, make_float2(S1, C1));
        pp = __ffma2_rn(make_float2(r2, r2), pp, make_float2(S0, C0));
        const float ps = pp.x, pc = pp.y;
#else
        float ps = fmaf(r2, S4, S3);
        ps = fmaf(r2, ps, S2);
        ps = fmaf(r2, ps, S1);
        ps = fmaf(r2, ps, S0);
        float pc = fmaf(r2, C4, C3);
        pc = fmaf(r2, pc, C2);
        pc = fmaf(r2, pc, C1);
        pc = fmaf(r2, pc, C0);
#endif
        const float sv = r * ps;
        const float cv = fmaf(r2, pc, 1.0f);
        constexpr int Q = ((Q0 % 4) + 4) % 4;     // sin / cos (x + 90 Q): (s, c), (c, -s), (-s, -c), (-c, s)
        constexpr unsigned fs = (Q == 2 || Q == 3) ? 0x80000000u : 0u;
        constexpr unsigned fc = (Q == 1 || Q == 2) ? 0x80000000u : 0u;
        s = __int_as_float(__float_as_int((Q & 1) ? cv : sv) ^ (sign ^ fs));
        c = __int_as_float(__float_as_int((Q & 1) ? sv : cv) ^ (sign ^ fc));
    } else if (TRIG == TRIG_ACCURATE) {
        float t = fmaf(deg, 1.0f / 90.0f, kMagic);
        int n = __float_as_int(t) + Q0;           // low 2 bits: quadrant
        float q = t - kMagic;
        float r = fmaf(q, -90.0f, deg);           // exact, |r| <= 45 (+ 1 ulp of slack)
        float r2 = r * r;
        constexpr double D = 3.14159265358979323846 / 180.0;
        constexpr float S0 = (float)D;
        constexpr float S1 = (float)(-1.6666654611e-1 * D * D * D);
        constexpr float S2 = (float)(8.3321608736e-3 * D * D * D * D * D);
        constexpr float S3 = (float)(-1.9515295891e-4 * D * D * D * D * D * D * D);
        constexpr float C1 = (float)(-0.5 * D * D);
        constexpr float C2 = (float)(4.166664568298827e-2 * D * D * D * D);
        constexpr float C3 = (float)(-1.388731625493765e-3 * D * D * D * D * D * D);
        constexpr float C4 = (float)(2.443315711809948e-5 * D * D * D * D * D * D * D * D);
#if DHFK_PACKED_V3    // the two Horner chains side by side in the lanes of FFMA2: 3 instructions instead of 6
        float2 pp = __ffma2_rn(make_float2(r2, r2), make_float2(S3, C4), make_float2(S2, C3));
        pp = __ffma2_rn(make_float2(r2, r2), pp, make_float2(S1, C2));
        pp = __ffma2_rn(make_float2(r2, r2), pp, make_float2(S0, C1));
        float sv = r * pp.x;
        float cv = fmaf(r2, pp.y, 1.0f);
#else
        float ps = fmaf(r2, S3, S2);
        ps = fmaf(r2, ps, S1);
        ps = fmaf(r2, ps, S0);
        float sv = r * ps;
        float pc = fmaf(r2, C4, C3);
        pc = fmaf(r2, pc, C2);
        pc = fmaf(r2, pc, C1);
        float cv = fmaf(r2, pc, 1.0f);
#endif
        bool odd = (n & 1) != 0;
        float so = odd ? cv : sv;
        float co = odd ? sv : cv;
        s = __int_as_float(__float_as_int(so) ^ ((n << 30) & 0x80000000));
        c = __int_as_float(__float_as_int(co) ^ (((n + 1) << 30) & 0x80000000));
    } else {
#if DHFK_MUFU_EXACT_REDUCTION
        float t = fmaf(deg, 1.0f / 360.0f, kMagic);
        float q = t - kMagic;
        float r = fmaf(q, -360.0f, deg);          // exact, |r| <= 180
        float x = r * kDegToRad;
#else
        float x = deg * kDegToRad;                // MUFU takes the fractional turn itself, see DHFK_MUFU_EXACT_REDUCTION
#endif
        float sv = __sinf(x), cv = __cosf(x);
        constexpr int Q = ((Q0 % 4) + 4) % 4;
        if (Q == 0) { s = sv; c = cv; }
        else if (Q == 1) { s = cv; c = -sv; }
        else if (Q == 2) { s = -sv; c = -cv; }
        else { s = -cv; c = sv; }
    }
}

// Same with a runtime quadrant offset (used by the shared limb routine).
template <int TRIG>
DHFK_DI void sincos_deg_rt(float deg, int q0, float& s, float& c) {
    const float kMagic = 12582912.0f;
    if (TRIG == TRIG_ACCURATE && DHFK_ACCURATE_TABLE) {
        sincos_table_core(deg, q0, s, c);
    } else if (TRIG == TRIG_ACCURATE) {
        float t = fmaf(deg, 1.0f / 90.0f, kMagic);
        int n = __float_as_int(t) + q0;
        float q = t - kMagic;
        float r = fmaf(q, -90.0f, deg);
        float r2 = r * r;
        constexpr double D = 3.14159265358979323846 / 180.0;
        constexpr float S0 = (float)D;
        constexpr float S1 = (float)(-1.6666654611e-1 * D * D * D);
        constexpr float S2 = (float)(8.3321608736e-3 * D * D * D * D * D);
        constexpr float S3 = (float)(-1.9515295891e-4 * D * D * D * D * D * D * D);
        constexpr float C1 = (float)(-0.5 * D * D);
        constexpr float C2 = (float)(4.166664568298827e-2 * D * D * D * D);
        constexpr float C3 = (float)(-1.388731625493765e-3 * D * D * D * D * D * D);
        constexpr float C4 = (float)(2.443315711809948e-5 * D * D * D * D * D * D * D * D);
#if DHFK_PACKED_V3    // the two Horner chains side by side in the lanes of FFMA2: 3 instructions instead of 6
        float2 pp = __ffma2_rn(make_float2(r2, r2), make_float2(S3, C4), make_float2(S2, C3));
        pp = __ffma2_rn(make_float2(r2, r2), pp, make_float2(S1, C2));
        pp = __ffma2_rn(make_float2(r2, r2), pp, make_float2(S0, C1));
        float sv = r * pp.x;
        float cv = fmaf(r2, pp.y, 1.0f);
#else
        float ps = fmaf(r2, S3, S2);
        ps = fmaf(r2, ps, S1);
        ps = fmaf(r2, ps, S0);
        float sv = r * ps;
        float pc = fmaf(r2, C4, C3);
        pc = fmaf(r2, pc, C2);
        pc = fmaf(r2, pc, C1);
        float cv = fmaf(r2, pc, 1.0f);
#endif
        bool odd = (n & 1) != 0;
        float so = odd ? cv : sv;
        float co = odd ? sv : cv;
        s = __int_as_float(__float_as_int(so) ^ ((n << 30) & 0x80000000));
        c = __int_as_float(__float_as_int(co) ^ (((n + 1) << 30) & 0x80000000));
    } else {
        // the reference itself forms fl(theta0 + angle) in fp32 before sin/cos (:601 etc.)
        float d2 = fmaf(90.0f, (float)q0, deg);
#if DHFK_MUFU_EXACT_REDUCTION
        float t = fmaf(d2, 1.0f / 360.0f, kMagic);
        float q = t - kMagic;
        float r = fmaf(q, -360.0f, d2);
        float x = r * kDegToRad;
#else
        float x = d2 * kDegToRad;
#endif
        s = __sinf(x);
        c = __cosf(x);
    }
}

// ---------------------------------------------------------------------------------------
// DH joint step.  T_j = Rot_x(alpha) Trans_x(a) Rot_z(theta) Trans_z(d) (modified DH,
// forward_kinematics_DH_model.py:99-114):  W_j = W_parent * T_j, i.e.
//   o_j = o_p + a x_p - sin(alpha) d y_p + cos(alpha) d z_p
//   [x y z]_j = [x_p, ca y_p + sa z_p, -sa y_p + ca z_p] * Rz(theta)
// With alpha in {0,+-90} the twist is a signed axis permutation.
// ---------------------------------------------------------------------------------------
template <int J>
DHFK_DI void twist(const Frame& F, V3& y1, V3& z1) {
    constexpr int A = ALPHA_Q[J];
    if constexpr (A == 0) { y1 = F.Y; z1 = F.Z; }
    else if constexpr (A == -1) { y1 = -F.Z; z1 = F.Y; }
    else { y1 = F.Z; z1 = -F.Y; }
}
// direction (in the parent frame axes) along which this joint's bone length moves the origin
template <int J>
DHFK_DI V3 length_axis(const Frame& F) {
    constexpr int A = ALPHA_Q[J], KIND = LEN_KIND[J], SIGN = LEN_SIGN[J];
    V3 a;
    if constexpr (KIND == 1) a = F.X;
    else if constexpr (A == 0) a = F.Z;   // 'd': -sin(alpha) y_p + cos(alpha) z_p
    else if constexpr (A == -1) a = F.Y;
    else a = -F.Y;
    if constexpr (SIGN > 0) return a;
    else return -a;
}
template <int J>
DHFK_DI void advance_origin(Frame& F, const float* bone) {
    constexpr int KIND = LEN_KIND[J], BONE = LEN_BONE[J];
    if constexpr (KIND != 0) F.O = axpy(bone[BONE], length_axis<J>(F), F.O);
}
template <int J>
DHFK_DI void rotate_joint(Frame& F, float s, float c) {
    V3 y1, z1;
    twist<J>(F, y1, z1);
    V3 x = axpy(s, y1, scale(c, F.X));
    V3 y = axpy(c, y1, scale(-s, F.X));
    F.X = x; F.Y = y; F.Z = z1;
}
DHFK_DI Frame identity_frame() {
    Frame F;
    F.X = v3(1.f, 0.f, 0.f); F.Y = v3(0.f, 1.f, 0.f); F.Z = v3(0.f, 0.f, 1.f); F.O = v3(0.f, 0.f, 0.f);
    return F;
}

// ---------------------------------------------------------------------------------------
// Camera constants: passed by value as a kernel parameter (constant bank operands).
// M maps v = X_world - t to camera space exactly as qrot(conj(q), v) does
// (common/quaternion.py:6-24, common/camera.py:36-38), valid for non-unit q too.
// ---------------------------------------------------------------------------------------
struct alignas(8) CamConst {
    float M[9];
    float t[3];
    float2 Mc[3];      // columns of the top two rows of M: (M00,M10), (M01,M11), (M02,M12) -- packed operands of cam_mat_vec
    float2 txy;        // (t[0], t[1])
    float2 f, c, p;    // focal, principal point, tangential distortion as (x, y) pairs: FFMA2 operands
    float k[3];
    float k1x2, k2x3;  // 2*k[1], 3*k[2]
};

// X_cam = M (W - t) and M g + b with the camera constants, (x, y) lanes packed; same operation order as mat_vec /
// mat_vec_add, so the results are bit-identical to the scalar forms
DHFK_DI V3 cam_space(const CamConst& cc, V3 W) {
#if DHFK_PACKED_V3
    const float2 dxy = __fadd2_rn(xy_of(W), make_float2(-cc.txy.x, -cc.txy.y));
    const float dz = W.z - cc.t[2];
    float2 r = __fmul2_rn(cc.Mc[2], make_float2(dz, dz));
    r = __ffma2_rn(cc.Mc[1], make_float2(dxy.y, dxy.y), r);
    r = __ffma2_rn(cc.Mc[0], make_float2(dxy.x, dxy.x), r);
    return v3p(r, fmaf(cc.M[6], dxy.x, fmaf(cc.M[7], dxy.y, cc.M[8] * dz)));
#else
    return mat_vec(cc.M, v3(W.x - cc.t[0], W.y - cc.t[1], W.z - cc.t[2]));
#endif
}
DHFK_DI V3 cam_mat_vec_add(const CamConst& cc, V3 v, V3 b) {
#if DHFK_PACKED_V3
    float2 r = __ffma2_rn(cc.Mc[2], make_float2(v.z, v.z), xy_of(b));
    r = __ffma2_rn(cc.Mc[1], make_float2(v.y, v.y), r);
    r = __ffma2_rn(cc.Mc[0], make_float2(v.x, v.x), r);
    return v3p(r, fmaf(cc.M[6], v.x, fmaf(cc.M[7], v.y, fmaf(cc.M[8], v.z, b.z))));
#else
    return mat_vec_add(cc.M, v, b);
#endif
}

// Packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2): the projection works on (x, y) lanes, so every
// per-lane product/FMA is one instruction for both lanes; scalar operands broadcast for free (".F32").
#ifndef DHFK_PACKED_PROJ
#define DHFK_PACKED_PROJ 1
#endif
DHFK_DI float2 bc2(float s) { return make_float2(s, s); }

struct ProjAux { float rx, ry, x, y, r2, S, iz; };

// project_to_2d (common/camera.py:85-94) for one camera-space point
DHFK_DI float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // MUFU.RCP, <= 1 ulp; 1/0 = inf like the reference's x / 0
    return r;
}
// torch.clamp(x, -1, 1) (common/camera.py:85) propagates NaN (0/0 at x = z = 0, or NaN inputs); fminf / fmaxf would
// return the non-NaN operand.  min.NaN / max.NaN are the same single FMNMX instruction with NaN propagation.
DHFK_DI float clamp_unit_nan(float x) {
    float r;
    asm("max.NaN.f32 %0, %1, 0fBF800000;\n\tmin.NaN.f32 %0, %0, 0f3F800000;" : "=f"(r) : "f"(x));
    return r;
}
DHFK_DI void project_point(const CamConst& cc, V3 X, float& u, float& v, ProjAux& a) {
    float iz = rcp_approx(X.z);
    a.iz = iz;
#if DHFK_PACKED_PROJ
    const float2 r = __fmul2_rn(make_float2(X.x, X.y), bc2(iz));
    a.rx = r.x; a.ry = r.y;
    a.x = clamp_unit_nan(r.x);
    a.y = clamp_unit_nan(r.y);
    a.r2 = fmaf(a.x, a.x, a.y * a.y);
    float radial = fmaf(a.r2, fmaf(a.r2, fmaf(a.r2, cc.k[2], cc.k[1]), cc.k[0]), 1.f);
    float tan = fmaf(cc.p.x, a.x, cc.p.y * a.y);
    a.S = radial + tan;
    const float2 sxy = __ffma2_rn(make_float2(a.x, a.y), bc2(a.S), __fmul2_rn(cc.p, bc2(a.r2)));
    const float2 uv = __ffma2_rn(cc.f, sxy, cc.c);
    u = uv.x; v = uv.y;
#else
    a.rx = X.x * iz;
    a.ry = X.y * iz;
    a.x = clamp_unit_nan(a.rx);
    a.y = clamp_unit_nan(a.ry);
    a.r2 = fmaf(a.x, a.x, a.y * a.y);
    float radial = fmaf(a.r2, fmaf(a.r2, fmaf(a.r2, cc.k[2], cc.k[1]), cc.k[0]), 1.f);
    float tan = fmaf(cc.p.x, a.x, cc.p.y * a.y);
    a.S = radial + tan;
    float sx = fmaf(a.x, a.S, cc.p.x * a.r2);
    float sy = fmaf(a.y, a.S, cc.p.y * a.r2);
    u = fmaf(cc.f.x, sx, cc.c.x);
    v = fmaf(cc.f.y, sy, cc.c.y);
#endif
}
// gradient wrt the camera-space point; torch.clamp passes gradient where -1 <= x/z <= 1
DHFK_DI V3 project_point_bwd(const CamConst& cc, const ProjAux& a, float gu, float gv) {
#if DHFK_PACKED_PROJ
    const float2 a01 = __fmul2_rn(cc.f, make_float2(gu, gv));
    float drad = fmaf(a.r2, fmaf(a.r2, cc.k2x3, cc.k1x2), cc.k[0]);
    float ax = fmaf(a01.x, a.x, a01.y * a.y);
    float ap = fmaf(a01.x, cc.p.x, a01.y * cc.p.y);
    float t2 = 2.f * fmaf(ax, drad, ap);
    const float2 xy = make_float2(a.x, a.y);
    float2 g = __ffma2_rn(a01, bc2(a.S), __ffma2_rn(xy, bc2(t2), __fmul2_rn(cc.p, bc2(ax))));
    g.x = (a.x == a.rx) ? g.x : 0.f;
    g.y = (a.y == a.ry) ? g.y : 0.f;
    const float2 gi = __fmul2_rn(g, bc2(a.iz));
    return v3(gi.x, gi.y, -fmaf(g.x, a.rx, g.y * a.ry) * a.iz);
#else
    float a0 = cc.f.x * gu, a1 = cc.f.y * gv;
    float drad = fmaf(a.r2, fmaf(a.r2, cc.k2x3, cc.k1x2), cc.k[0]);
    float ax = fmaf(a0, a.x, a1 * a.y);
    float ap = fmaf(a0, cc.p.x, a1 * cc.p.y);
    float t2 = 2.f * fmaf(ax, drad, ap);
    float gx = fmaf(a0, a.S, fmaf(a.x, t2, ax * cc.p.x));
    float gy = fmaf(a1, a.S, fmaf(a.y, t2, ax * cc.p.y));
    gx = (a.x == a.rx) ? gx : 0.f;
    gy = (a.y == a.ry) ? gy : 0.f;
    return v3(gx * a.iz, gy * a.iz, -fmaf(gx, a.rx, gy * a.ry) * a.iz);
#endif
}

// tanh of two values at once: t = 1 - 2 / (1 + 2^(2x log2 e)), packed fp32x2 arithmetic around MUFU.EX2 / MUFU.RCP
// (7 instructions per pair).  Saturates cleanly (2^+big = inf -> rcp 0 -> 1; 2^-big = 0 -> -1), NaN stays NaN.
// Absolute error ~1.5e-7 (ex2.approx is 2^-22 relative on a value near 1, so no formula built on it is relatively
// accurate around 0; as an angle that is 3e-5 degrees at the widest slot range).
DHFK_DI float2 tanh2(float2 x) {
    const float2 a = __fmul2_rn(x, bc2(2.885390043f));
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
    const float2 d = __fadd2_rn(e, bc2(1.0f));
    const float2 r = make_float2(rcp_approx(d.x), rcp_approx(d.y));
    return __ffma2_rn(r, bc2(-2.0f), bc2(1.0f));
}
// sech^2 = d tanh / dx from the stored tanh value: 1 - t^2 (absolute error ~2e-7; the gradient it scales is judged
// against max(|ref|, 1), and half * d/d(angle) is O(1))
DHFK_DI float sech2_from_tanh(float t) { return fmaf(-t, t, 1.0f); }
// per-slot affine map of the generator epilogue (host-filled: half = (hi-lo)/2, mid = (hi+lo)/2, or
// 180 / 0 when GAN_whether_use_preAngle is off) and the root scale (10)
struct GenScale {
    float2 hm[GEN_NSLOT];     // (half, mid) per slot: one 64-bit constant load where the slot index is a run-time value
    float root_scale;
};

// ---------------------------------------------------------------------------------------
// Forward walker: depth-first over the compile-time tree.  Ctx provides
//   template<int J> float angle()  (joint angle, degrees), const float* bone (15 lengths)
//   template<int K> void emit(V3 origin_in_chain_frame)
// ---------------------------------------------------------------------------------------
template <int TRIG, int J, class Ctx> DHFK_DI void fwd_walk(Frame F, Ctx& ctx);

template <int TRIG, int J, int I, class Ctx>
DHFK_DI void fwd_children(const Frame& F, Ctx& ctx) {
    constexpr int C = nth_child(J, I);
    if constexpr (C >= 0) {
        fwd_walk<TRIG, C>(F, ctx);
        fwd_children<TRIG, J, I + 1>(F, ctx);
    }
}
template <int TRIG, int J, class Ctx>
DHFK_DI void fwd_walk(Frame F, Ctx& ctx) {
    advance_origin<J>(F, ctx.bone);
    constexpr int K = out_index_of_joint(J);
    if constexpr (K >= 0) ctx.template emit<K>(F.O);
    if constexpr (!is_leaf(J)) {
        float s, c;
        constexpr int Q0 = THETA0_Q[J];
        sincos_deg<TRIG, Q0>(ctx.template angle<J>(), s, c);
        rotate_joint<J>(F, s, c);
        fwd_children<TRIG, J, 0>(F, ctx);
    }
}

// ---------------------------------------------------------------------------------------
// Backward walker.  Returns the wrench (F, M about the chain origin, chain-frame
// coordinates) of every output at or below joint J.  For a rotating joint
//   dL/dtheta_j = (pi/180) * z_j . (M_j - o_j x F_j)
// where (F_j, M_j) sums the outputs strictly below j (the joint's own origin does not
// depend on its own theta), z_j is the joint axis after the alpha twist, o_j its origin.
// Ctx provides ang, bone and
//   template<int J> float angle()            -- body joints (compile-time index)
//   template<int K> float limb_angle(L)      -- joint K of limb L (run-time limb, compile-time position)
//   template<int K> V3 upstream(V3 origin)   -- total dL/d(origin) in the chain frame
//   template<int J> void grad_angle(float g), zero_grad_angle<J>();  limb_grad<K>(L, g), limb_zero_leaf(L)
//   void grad_bone(int b, float g)           -- only called when Ctx::kBoneGrad
// ---------------------------------------------------------------------------------------
template <int TRIG, int J, class Ctx> DHFK_DI Wrench bwd_walk(Frame F, Ctx& ctx);

// Limb descriptors live in constant memory: the limb loop index is warp-uniform, so they are read
// through the uniform datapath.
static __constant__ LimbDesc c_limbs[NLIMB] = {make_limb(0), make_limb(1), make_limb(2), make_limb(3)};

// One 5-joint limb, forward then reverse.  B = parent frame with the limb's alpha0 twist folded in.
// Extra Ctx members used:  void load_limb(const LimbDesc&),  template<int I> V3 upstream_limb(V3 origin)
template <int TRIG, class Ctx>
DHFK_DI Wrench bwd_limb(const Frame& B, const LimbDesc& L, Ctx& ctx) {
    const float sg = L.sigma;
    float s, c;
    ctx.load_limb(L);     // the limb's 3 upstream-gradient rows: 128-bit shared loads, no bank conflicts
    // joint 0: hip / shoulder offset along parent x, rotation about B.Z
    const V3 o0 = axpy(L.sgn0 * ctx.bone[L.b0], B.X, B.O);
    const V3 g0 = ctx.template upstream_limb<0>(o0);
    sincos_deg_rt<TRIG>(ctx.template limb_angle<0>(L), L.q0, s, c);
    const V3 X1 = axpy(s, B.Y, scale(c, B.X));
    const V3 Y1 = axpy(c, B.Y, scale(-s, B.X));
    // joint 1: alpha = sigma*90  =>  y' = sigma z_p, z' = -sigma y_p
    const V3 zj1 = scale(-sg, Y1);
    sincos_deg_rt<TRIG>(ctx.template limb_angle<1>(L), -1, s, c);
    const V3 X2 = axpy(s * sg, B.Z, scale(c, X1));
    const V3 Y2 = axpy(c * sg, B.Z, scale(-s, X1));
    // joint 2: same twist
    const V3 zj2 = scale(-sg, Y2);
    sincos_deg_rt<TRIG>(ctx.template limb_angle<2>(L), L.q2, s, c);
    const V3 X3 = axpy(s * sg, zj1, scale(c, X2));
    const V3 Y3 = axpy(c * sg, zj1, scale(-s, X2));
    // joint 3: knee / elbow, alpha 0, axis zj2
    const V3 o3 = axpy(ctx.bone[L.b3], X3, o0);
    const V3 g3 = ctx.template upstream_limb<1>(o3);
    sincos_deg_rt<TRIG>(ctx.template limb_angle<3>(L), 0, s, c);
    const V3 X4 = axpy(s, Y3, scale(c, X3));
    // joint 4: foot / wrist, leaf
    const V3 o4 = axpy(ctx.bone[L.b4], X4, o3);
    const V3 g4 = ctx.template upstream_limb<2>(o4);
    // reverse sweep
    Wrench w;
    w.F = g4;
    w.M = cross(o4, g4);
    ctx.limb_zero_leaf(L);
    ctx.template limb_grad<3>(L, kDegToRad * dot(zj2, sub_cross(w.M, o3, w.F)));
    if (Ctx::kBoneGrad) ctx.grad_bone(L.b4, dot(X4, w.F));
    w.F = w.F + g3;
    w.M = add_cross(w.M, o3, g3);
    if (Ctx::kBoneGrad) ctx.grad_bone(L.b3, dot(X3, w.F));
    const V3 tau = sub_cross(w.M, o0, w.F);
    ctx.template limb_grad<2>(L, kDegToRad * dot(zj2, tau));
    ctx.template limb_grad<1>(L, kDegToRad * dot(zj1, tau));
    ctx.template limb_grad<0>(L, kDegToRad * dot(B.Z, tau));
    w.F = w.F + g0;
    w.M = add_cross(w.M, o0, g0);
    if (Ctx::kBoneGrad) ctx.grad_bone(L.b0, L.sgn0 * dot(B.X, w.F));
    return w;
}

// All four limbs in ONE non-unrolled loop (called where the arms branch off the body; the legs, which
// hang off the chain origin, ride along so that the limb code exists exactly once).  Returns the
// arms' wrench; the legs' wrench is left in ctx.legs.
template <int TRIG, class Ctx>
DHFK_DI Wrench bwd_all_limbs(const Frame& P /*frame of the arms' parent joint*/, Ctx& ctx) {
    Wrench acc[2];          // [0] legs (hang off the chain root frame), [1] arms
    Frame B = ctx.base;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        if (half == 1) {    // arms: alpha0 = -90 folded into the base frame (y' = -z, z' = y)
            B.X = P.X; B.Y = -P.Z; B.Z = P.Y; B.O = P.O;
        }
        Wrench w0;
        w0.F = w0.M = v3(0.f, 0.f, 0.f);
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
#ifdef DHFK_EXPERIMENT_LIMBS      // sensitivity experiment only (wrong results): run fewer limbs
            if (2 * half + i >= DHFK_EXPERIMENT_LIMBS) break;
#endif
            const Wrench w = bwd_limb<TRIG>(B, c_limbs[2 * half + i], ctx);
            w0.F = w0.F + w.F;
            w0.M = w0.M + w.M;
        }
        if (half == 0) acc[0] = w0; else acc[1] = w0;
    }
    const Wrench legs = acc[0], arms = acc[1];
    ctx.legs = legs;
    return arms;
}

// children of joint J from index I on; limb roots are routed to the shared limb loop
template <int TRIG, int J, int I, bool HAVE, class Ctx>
DHFK_DI Wrench bwd_children_acc(const Frame& F, Ctx& ctx, Wrench acc) {
    constexpr int C = nth_child(J, I);
    if constexpr (C < 0) {
        static_assert(HAVE, "joint without children reached bwd_children_acc");
        return acc;
    } else {
        constexpr int LIMB = limb_of_root(C);
        if constexpr (LIMB == 3) {   // second arm: handled together with the first
            return bwd_children_acc<TRIG, J, I + 1, HAVE>(F, ctx, acc);
        } else {
            Wrench w;
            if constexpr (LIMB == 2) w = bwd_all_limbs<TRIG>(F, ctx);
            else w = bwd_walk<TRIG, C>(F, ctx);
            if constexpr (HAVE) { w.F = acc.F + w.F; w.M = acc.M + w.M; }
            return bwd_children_acc<TRIG, J, I + 1, true>(F, ctx, w);
        }
    }
}
template <int TRIG, int J, int I, class Ctx>
DHFK_DI Wrench bwd_children(const Frame& F, Ctx& ctx) {
    Wrench zero;
    zero.F = zero.M = v3(0.f, 0.f, 0.f);
    return bwd_children_acc<TRIG, J, I, false>(F, ctx, zero);
}
template <int TRIG, int J, class Ctx>
DHFK_DI Wrench bwd_walk(Frame F, Ctx& ctx) {
    V3 laxis = v3(0.f, 0.f, 0.f);
    constexpr bool kLen = Ctx::kBoneGrad && LEN_KIND[J] != 0;
    if constexpr (kLen) laxis = length_axis<J>(F);
    advance_origin<J>(F, ctx.bone);
    const V3 o = F.O;
    constexpr int K = out_index_of_joint(J);
    V3 gh = v3(0.f, 0.f, 0.f);
    if constexpr (K >= 0) gh = ctx.template upstream<K>(o);
    Wrench w;
    if constexpr (!is_leaf(J)) {
        V3 y1, zj;
        twist<J>(F, y1, zj);
        float s, c;
        constexpr int Q0 = THETA0_Q[J];
        sincos_deg<TRIG, Q0>(ctx.template angle<J>(), s, c);
        rotate_joint<J>(F, s, c);
        w = bwd_children<TRIG, J, 0>(F, ctx);
        constexpr bool kZero = origin_is_zero(J);
        V3 tau;
        if constexpr (kZero) tau = w.M;
        else tau = sub_cross(w.M, o, w.F);
        ctx.template grad_angle<J>(kDegToRad * dot(zj, tau));
        if constexpr (K >= 0) {
            w.F = w.F + gh;
            if constexpr (!kZero) w.M = add_cross(w.M, o, gh);
        }
    } else {
        static_assert(K >= 0, "every chain end is an output joint");
        ctx.template zero_grad_angle<J>();
        w.F = gh;
        w.M = cross(o, gh);
    }
    if constexpr (kLen) {
        constexpr int BONE = LEN_BONE[J];
        ctx.grad_bone(BONE, dot(laxis, w.F));
    }
    return w;
}

}  // namespace dhfk
