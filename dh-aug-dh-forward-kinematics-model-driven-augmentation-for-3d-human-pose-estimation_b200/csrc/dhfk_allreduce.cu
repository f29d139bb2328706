// dhfk_allreduce.cu -- SURVEY 8 (e): the one exchange step of the data-parallel GAN iteration.
// The FK / projection path has no collective (poses are independent); what the ranks exchange is the gradient of the
// three small MLPs, five optimizer steps per iteration (models_Fk_GAN/model_fk_gan_train.py:314-341, 382-409, 415-482).
// Those gradients live in ONE flat buffer per rank (dhfk.parallel.FlatGradBuffer), a few MB: latency-bound.  NCCL's
// ring / tree launch moves 6.4 MB between 8 B200s in 74 us and holds 16-32 CTAs while it does; this kernel does the
// exchange in one pass over NVLink peer memory, in place, with a handful of CTAs:
//
//   rank r owns slice r of the buffer.  For every 16-byte element of its slice it
//     NVLS      : multimem.ld_reduce.add  from the multicast address  (the NVSwitch sums the R copies in flight)
//     peer path : R loads, one from every rank's buffer, summed in rank order
//   scales by 1/R (average) and writes the result into every rank's buffer
//     NVLS      : one multimem.st to the multicast address             (the switch replicates it)
//     peer path : R stores.
//   Each element is read and written by exactly one thread of one rank, so the exchange is in place without a
//   staging buffer, and every rank ends up with bit-identical values (replicas cannot drift apart).
//
// Two cross-GPU barriers bracket the pass (CTA b of every rank with CTA b of every other rank): "every rank's gradients
// are complete" before the first load, "every rank's stores are visible" before the kernel ends.  The flag words in
// peer memory carry a call counter and are never reset; the counter itself lives in device memory next to the flags
// (one word per CTA slot, advanced by the kernel), so the launch has no per-call argument and a captured CUDA graph
// can replay it.  A spin that outlasts `timeout_ns` gives up and raises *status instead of hanging the GPU.
#include "../../include/dhfk.h"
#include "dhfk_launch.h"

namespace dhfk {

constexpr int kArMaxWorld = DHFK_AR_MAX_WORLD;
constexpr int kArMaxCtas = DHFK_AR_MAX_CTAS;
constexpr int kArMaxThreads = 512;     // threads per CTA are a launch parameter (32..512)
// 16-byte elements in flight per thread.  NVLS path: 8 lets 8 CTAs reach the floor that 4 needs 16 CTAs for
// (profiles/r2u_unroll_n2.txt).  Peer path: every element is `world` loads, and 8 x 8 in flight measured 88 us for 6.4 MB
// on 8 GPUs where 4 x 8 takes 49 us (profiles/r2t2_peer_exchange_n8.json vs r2t_peer_exchange_n8.json).
constexpr int kArUnrollMc = 8;
constexpr int kArUnrollPeer = 4;

struct ArParams {
    float4* buf[kArMaxWorld];       // every rank's buffer range (peer-mapped addresses), index = rank
    float4* mc;                     // multicast address of the same range, or null
    unsigned* flags[kArMaxWorld];   // every rank's flag block: [kArMaxCtas][2][kArMaxWorld] flag words, then
                                    // [kArMaxCtas] call counters (only the local rank's counters are used)
    unsigned* status;               // local word: set to the call counter when a barrier timed out
    long long nvec;                 // 16-byte elements in the range
    unsigned long long timeout_ns;
    float scale;
    int rank, world;
};

DHFK_DI void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
DHFK_DI unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
DHFK_DI unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Strong system-scope accesses: never a stale L1 line of peer memory.  (Weak .cg loads / stores ordered by the barriers'
// fences measured the same: 41.7 vs 42.2 us for 6.4 MB on the peer path, profiles/r2r_*.)
DHFK_DI float4 ld_sys_v4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
DHFK_DI void st_sys_v4(float4* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
DHFK_DI float4 multimem_ld_reduce_v4(const float4* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
DHFK_DI void multimem_st_v4(float4* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// CTA `b` of this rank meets CTA `b` of every other rank.  Thread t < world signals rank t and waits for rank t.
DHFK_DI bool ar_barrier(const ArParams& p, int b, int phase, unsigned epoch) {
    __syncthreads();                      // every thread's stores of this CTA are ordered before the signal
    bool ok = true;
    if ((int)threadIdx.x < p.world) {
        const int peer = threadIdx.x;
        const int slot = (b * 2 + phase) * kArMaxWorld;
        if (phase) __threadfence_system();     // phase 0 follows no store of this kernel
        st_release_sys(p.flags[peer] + slot + p.rank, epoch);
        const unsigned* mine = p.flags[p.rank] + slot + peer;
        const unsigned long long t0 = global_ns();
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
            if (global_ns() - t0 > p.timeout_ns) { ok = false; break; }
        }
    }
    ok = __syncthreads_and(ok) != 0;
    if (!ok && threadIdx.x == 0) atomicExch(p.status, epoch ? epoch : 1u);
    return ok;
}

template <bool MC>
__global__ void __launch_bounds__(kArMaxThreads) dhfk_allreduce_kernel(const __grid_constant__ ArParams p) {
    constexpr int kArUnroll = MC ? kArUnrollMc : kArUnrollPeer;
    const int b = blockIdx.x;
    // this CTA slot's call counter: every rank has made the same calls, so the slots agree across ranks
    unsigned* counter = p.flags[p.rank] + kArMaxCtas * 2 * kArMaxWorld + b;
    const unsigned epoch = *counter + 1;
    if (!ar_barrier(p, b, 0, epoch)) return;
    const long long per_rank = (p.nvec + p.world - 1) / p.world;
    const long long lo = per_rank * p.rank < p.nvec ? per_rank * p.rank : p.nvec;
    const long long hi = lo + per_rank < p.nvec ? lo + per_rank : p.nvec;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = lo + (long long)b * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * kArUnroll) {
        float4 acc[kArUnroll];
        if (MC) {
#pragma unroll
            for (int u = 0; u < kArUnroll; ++u) {
                const long long i = i0 + u * stride;
                if (i < hi) acc[u] = multimem_ld_reduce_v4(p.mc + i);
            }
        } else {
#pragma unroll
            for (int u = 0; u < kArUnroll; ++u) {
                const long long i = i0 + u * stride;
                if (i < hi) {
                    acc[u] = ld_sys_v4(p.buf[0] + i);
                    for (int r = 1; r < p.world; ++r) {          // rank order: the same sum on whichever rank computes it
                        const float4 v = ld_sys_v4(p.buf[r] + i);
                        acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kArUnroll; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) {
                const float4 v = make_float4(acc[u].x * p.scale, acc[u].y * p.scale, acc[u].z * p.scale, acc[u].w * p.scale);
                if (MC) multimem_st_v4(p.mc + i, v);
                else
                    for (int r = 0; r < p.world; ++r) st_sys_v4(p.buf[r] + i, v);
            }
        }
    }
    if (ar_barrier(p, b, 1, epoch) && threadIdx.x == 0) *counter = epoch;
}

int launch_grad_allreduce(float* const* peer_bufs, float* mc_buf, unsigned* const* peer_flags, unsigned* status, int rank,
                          int world, long long n_floats, float scale, int max_ctas, int threads,
                          unsigned long long timeout_ns, cudaStream_t st, const char** where) {
    ArParams p = {};
    for (int r = 0; r < world; ++r) {
        p.buf[r] = reinterpret_cast<float4*>(peer_bufs[r]);
        p.flags[r] = peer_flags[r];
    }
    p.mc = reinterpret_cast<float4*>(mc_buf);
    p.status = status;
    p.nvec = n_floats / 4;
    p.timeout_ns = timeout_ns;
    p.scale = scale;
    p.rank = rank;
    p.world = world;
    // the same grid on every rank (CTA b pairs with CTA b): a function of the range and the world size only
    const long long per_rank = (p.nvec + world - 1) / world;
    const int unroll = mc_buf ? kArUnrollMc : kArUnrollPeer;
    long long want = (per_rank + (long long)threads * unroll - 1) / ((long long)threads * unroll);
    if (want < 1) want = 1;
    if (max_ctas < 1) max_ctas = 1;
    if (max_ctas > kArMaxCtas) max_ctas = kArMaxCtas;
    const unsigned blocks = (unsigned)(want < max_ctas ? want : max_ctas);
    *where = mc_buf ? "dhfk_allreduce_kernel<multimem>" : "dhfk_allreduce_kernel<peer>";
    if (mc_buf) dhfk_allreduce_kernel<true><<<blocks, threads, 0, st>>>(p);
    else dhfk_allreduce_kernel<false><<<blocks, threads, 0, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace dhfk
