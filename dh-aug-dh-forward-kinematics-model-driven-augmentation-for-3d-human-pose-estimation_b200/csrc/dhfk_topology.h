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

}  // namespace dhfk
