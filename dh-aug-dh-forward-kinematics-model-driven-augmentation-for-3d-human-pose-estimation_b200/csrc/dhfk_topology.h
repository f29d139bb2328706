// dhfk_topology.h -- compile-time kinematic tree of the 16-joint H36M DH skeleton.
//
// The tree is a set of constexpr tables.  The kernels walk it with template recursion
// (dhfk_walk.cuh), so every table lookup, every alpha in {0,+-90 deg} and every theta0
// quadrant is folded into the instruction stream / constant bank at compile time and dead
// frame columns are eliminated by the compiler.  dhfk_topology() (C ABI) exports the same
// tables so tests can compare them bit-exactly with the reference.
//
// Reference: models_Fk_GAN/forward_kinematics_DH_model.py:234-261 (alpha/theta tables),
// :571-589 (bone-length slots), :633,:648 (arms continue from body[8]),
// :751-817 + common/h36m_dataset.py:37-38 (output joints).
// Joint numbering = generator slot order (Fk_generator.py:179-184):
//   right leg 0-4 | left leg 5-9 | body 10-22 | right hand 23-27 | left hand 28-32
#pragma once

namespace dhfk {

constexpr int NJ = 33;     // local DH joints
constexpr int NOUT = 16;   // output joints (H36M 16-joint layout)
constexpr int NBONE = 15;  // bone lengths, used_16key_15bone_len_table order

// alpha in quarter turns (alpha_deg = 90 * ALPHA_Q)
constexpr int ALPHA_Q[NJ] = {0, -1, -1, 0, 0,  0, 1, 1, 0, 0,
                             0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 1,
                             -1, -1, -1, 0, 0,  -1, 1, 1, 0, 0};
// theta0 in quarter turns (theta0_deg = 90 * THETA0_Q)
constexpr int THETA0_Q[NJ] = {0, -1, 2, 0, 0,  2, -1, 0, 0, 0,
                              1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0,
                              -2, -1, 2, 0, 0,  0, -1, 0, 0, 0};
constexpr int PARENT[NJ] = {-1, 0, 1, 2, 3,  -1, 5, 6, 7, 8,
                            -1, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21,
                            18, 23, 24, 25, 26,  18, 28, 29, 30, 31};
// bone-length slot of each joint: 0 none, 1 = DH 'a' (along parent x), 2 = DH 'd'
constexpr int LEN_KIND[NJ] = {1, 0, 0, 1, 1,  1, 0, 0, 1, 1,
                              0, 0, 0, 2, 0, 0, 2, 0, 0, 0, 0, 0, 1,
                              1, 0, 0, 1, 1,  1, 0, 0, 1, 1};
constexpr int LEN_BONE[NJ] = {5, -1, -1, 3, 1,  4, -1, -1, 2, 0,
                              -1, -1, -1, 6, -1, -1, 7, -1, -1, -1, -1, -1, 14,
                              9, -1, -1, 11, 13,  8, -1, -1, 10, 12};
constexpr int LEN_SIGN[NJ] = {1, 0, 0, 1, 1,  -1, 0, 0, 1, 1,
                              0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1,
                              -1, 0, 0, 1, 1,  1, 0, 0, 1, 1};
// joint whose frame origin is output k
constexpr int OUT16[NOUT] = {10, 0, 3, 4, 5, 8, 9, 13, 16, 22, 28, 31, 32, 23, 26, 27};
constexpr int H36M_32_TO_16[NOUT] = {0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27};

constexpr int out_index_of_joint(int j) {
    for (int k = 0; k < NOUT; ++k)
        if (OUT16[k] == j) return k;
    return -1;
}
constexpr int num_children(int j) {
    int n = 0;
    for (int c = 0; c < NJ; ++c)
        if (PARENT[c] == j) ++n;
    return n;
}
// i-th child of joint j in ascending joint order, or -1.  j = -1 enumerates the chain roots.
constexpr int nth_child(int j, int i) {
    for (int c = 0; c < NJ; ++c)
        if (PARENT[c] == j) {
            if (i == 0) return c;
            --i;
        }
    return -1;
}
// A joint with no children needs no rotation: its theta cannot move any output.
constexpr bool is_leaf(int j) { return num_children(j) == 0; }
// True when the frame origin of joint j is exactly the chain origin (no length on the path).
constexpr bool origin_is_zero(int j) {
    while (j >= 0) {
        if (LEN_KIND[j] != 0) return false;
        j = PARENT[j];
    }
    return true;
}

// ---------------------------------------------------------------------------------------
// Generator epilogue (models_Fk_GAN/Fk_generator.py:121-168).  The network emits 35 columns per pose:
// 31 tanh'ed angle values scattered into a 37-slot vector (slots 4, 9, 22, 23, 28, 33 are forced to zero,
// :136; slots 34-36 are the global rotation), then each slot is mapped affinely to its range; column 31
// is never consumed and columns 32-34 are the root (tanh * 10, :122).
// Slot j is DH joint j for j < 33.
// ---------------------------------------------------------------------------------------
constexpr int GEN_NSLOT = 37;
constexpr int GEN_NCOL = 35;
constexpr int GEN_GROT_SLOT = 34;   // slots 34,35,36 = global rotation x,y,z
constexpr int GEN_ROOT_COL = 32;    // columns 32,33,34 = root x,y,z
constexpr int GEN_UNUSED_COL = 31;
constexpr bool gen_zero_slot(int i) { return i == 4 || i == 9 || i == 22 || i == 23 || i == 28 || i == 33; }
// network column feeding slot i, or -1 for a zero slot
constexpr int gen_src_col(int slot) {
    if (gen_zero_slot(slot)) return -1;
    int col = 0;
    for (int i = 0; i < slot; ++i)
        if (!gen_zero_slot(i)) ++col;
    return col;
}
static_assert(gen_src_col(0) == 0 && gen_src_col(5) == 4 && gen_src_col(32) == 27 && gen_src_col(34) == 28 &&
              gen_src_col(36) == 30, "generator slot map");

// ---------------------------------------------------------------------------------------
// Limbs.  The four 5-joint chains (right leg, left leg, right arm, left arm) have ONE structure:
//   joint 0: length along parent x (sign sgn0), output k0, alpha0 in {0,-90}, theta0 quadrant q0
//   joint 1: alpha = sigma*90, theta0 = -90          joint 2: alpha = sigma*90, theta0 quadrant q2
//   joint 3: length (+) along x, output k0+1, alpha 0, theta0 0
//   joint 4: length (+) along x, output k0+2, leaf
// so the backward kernel runs them through one runtime-parametrised routine (4x smaller hot code,
// which is what keeps the kernel resident in the instruction cache).  The descriptors below are
// DERIVED from the tables above and the pattern is checked at compile time.
// ---------------------------------------------------------------------------------------
constexpr int NLIMB = 4;
constexpr int LIMB_ROOT[NLIMB] = {0, 5, 23, 28};
constexpr int limb_of_root(int j) {
    for (int l = 0; l < NLIMB; ++l)
        if (LIMB_ROOT[l] == j) return l;
    return -1;
}
struct LimbDesc {
    int ang0;          // first joint / angle index
    int k0;            // output index of joint 0 (joints 3 and 4 are k0+1, k0+2)
    int b0, b3, b4;    // bone indices of the three lengths
    int q0, q2;        // theta0 quadrants of joints 0 and 2
    float sgn0;        // sign of the first length
    float sigma;       // alpha of joints 1 and 2 is sigma * 90 deg
    // where the limb's three upstream-gradient rows start inside a [16,3] / [16,2] pose row: first 16-byte chunk
    // and float offset inside it (the 9 / 6 floats are then read with 3 / 2 conflict-free 128-bit loads)
    int wc0, woff;     // [16,3] rows: floats 3*k0 .. 3*k0+8
    int uc0, uoff;     // [16,2] rows: floats 2*k0 .. 2*k0+5
    // generator mode: network column feeding joint 0 (-1: a fixed slot), joint 1 (joints 2 and 3 follow it) and the
    // leaf joint 4 (-1: a fixed slot)
    int gcol0, gcol1, gcol4;
};
constexpr LimbDesc make_limb(int l) {
    const int j = LIMB_ROOT[l];
    return LimbDesc{j, out_index_of_joint(j), LEN_BONE[j], LEN_BONE[j + 3], LEN_BONE[j + 4],
                    THETA0_Q[j], THETA0_Q[j + 2], (float)LEN_SIGN[j], (float)ALPHA_Q[j + 1],
                    (3 * out_index_of_joint(j)) / 4, (3 * out_index_of_joint(j)) % 4,
                    (2 * out_index_of_joint(j)) / 4, (2 * out_index_of_joint(j)) % 4,
                    gen_src_col(j), gen_src_col(j + 1), gen_src_col(j + 4)};
}
// the three inner joints of a limb are fed by consecutive network columns
constexpr bool limb_gen_cols_ok(int l) {
    const int j = LIMB_ROOT[l];
    return gen_src_col(j + 1) >= 0 && gen_src_col(j + 2) == gen_src_col(j + 1) + 1 &&
           gen_src_col(j + 3) == gen_src_col(j + 1) + 2;
}
static_assert(limb_gen_cols_ok(0) && limb_gen_cols_ok(1) && limb_gen_cols_ok(2) && limb_gen_cols_ok(3),
              "generator columns of a limb's inner joints are not consecutive");
// the 128-bit limb loads stay inside the 12- / 8-chunk rows and only the offsets handled below occur
constexpr bool limb_chunks_ok(int l) {
    const LimbDesc d = make_limb(l);
    return d.wc0 + 2 <= 11 && d.uc0 + 1 <= 7 && (d.woff == 0 || d.woff == 2 || d.woff == 3) &&
           (d.uoff == 0 || d.uoff == 2);
}
static_assert(limb_chunks_ok(0) && limb_chunks_ok(1) && limb_chunks_ok(2) && limb_chunks_ok(3),
              "limb upstream rows do not fit the 3-chunk / 2-chunk 128-bit load pattern");
constexpr bool limb_pattern_ok(int l) {
    const int j = LIMB_ROOT[l];
    const int k = out_index_of_joint(j);
    return PARENT[j + 1] == j && PARENT[j + 2] == j + 1 && PARENT[j + 3] == j + 2 && PARENT[j + 4] == j + 3 &&
           num_children(j) == 1 && num_children(j + 1) == 1 && num_children(j + 2) == 1 &&
           num_children(j + 3) == 1 && is_leaf(j + 4) &&
           LEN_KIND[j] == 1 && LEN_KIND[j + 1] == 0 && LEN_KIND[j + 2] == 0 && LEN_KIND[j + 3] == 1 &&
           LEN_KIND[j + 4] == 1 && LEN_SIGN[j + 3] == 1 && LEN_SIGN[j + 4] == 1 &&
           (ALPHA_Q[j] == 0 || ALPHA_Q[j] == -1) && (ALPHA_Q[j + 1] == 1 || ALPHA_Q[j + 1] == -1) &&
           ALPHA_Q[j + 2] == ALPHA_Q[j + 1] && ALPHA_Q[j + 3] == 0 && ALPHA_Q[j + 4] == 0 &&
           THETA0_Q[j + 1] == -1 && THETA0_Q[j + 3] == 0 &&
           k >= 0 && out_index_of_joint(j + 1) < 0 && out_index_of_joint(j + 2) < 0 &&
           out_index_of_joint(j + 3) == k + 1 && out_index_of_joint(j + 4) == k + 2;
}
static_assert(limb_pattern_ok(0) && limb_pattern_ok(1) && limb_pattern_ok(2) && limb_pattern_ok(3),
              "a limb does not follow the generic 5-joint pattern");
// legs hang off the chain origin with alpha0 = 0; both arms hang off the same body joint with alpha0 = -90
static_assert(PARENT[LIMB_ROOT[0]] == -1 && PARENT[LIMB_ROOT[1]] == -1 && ALPHA_Q[LIMB_ROOT[0]] == 0 &&
              ALPHA_Q[LIMB_ROOT[1]] == 0, "legs");
static_assert(PARENT[LIMB_ROOT[2]] == PARENT[LIMB_ROOT[3]] && PARENT[LIMB_ROOT[2]] >= 0 &&
              ALPHA_Q[LIMB_ROOT[2]] == -1 && ALPHA_Q[LIMB_ROOT[3]] == -1, "arms");

}  // namespace dhfk
