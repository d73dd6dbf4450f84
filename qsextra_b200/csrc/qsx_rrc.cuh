// qsx_rrc.cuh -- CTA-wide register-resident kernel for large blocks (include/qsx.h, qsx_rrc_args; engine/rrc.py).
//
// One CTA of 4^NBITS threads owns one instance.  Thread (m_a, m_b) keeps the configuration tile v[i][j] and the
// per-element diagonal factors w[i][j] in registers for the whole evolution.  In-thread: diagonal multiply, exciton hops
// of both sides.  Across threads: pseudomode rotations (partner differs in one thread bit: warp shuffle for lane bits,
// double-buffered shared-memory exchange for warp bits) and amplitude damping (the (1,1) threads of a bit pair hand
// their tile to the (0,0) threads through shared memory).  The operation list is a compile-time constant (BigProg<PID>).
#pragma once
#ifdef __CUDACC_RTC__            // run-time specialisation (engine/jit.py): libcu++ instead of the host's <utility>
#include <cuda/std/utility>
namespace qstd = cuda::std;
#else
#include <utility>
namespace qstd = std;
#endif

namespace rrc {

template <int PID>
struct BigProg;
#ifdef QSX_RRC_EXTERNAL_PROGRAMS  // names the header that defines BigProg<0> and kNumBigProgs = 1 for this unit
#include QSX_RRC_EXTERNAL_PROGRAMS
#else
#include "qsx_rrc_programs.inc"
#endif

enum { OP_DIAG = 0, OP_HOPK = 1, OP_HOPB = 2, OP_ROTK = 3, OP_ROTB = 4, OP_AD = 5, OP_ADT = 6 };

struct Params {
    qsx_rrc_args a;
};

template <int PID>
struct Geo {
    using BP = BigProg<PID>;
    static constexpr int NA = BP::NA, NB = BP::NB, NBITS = BP::NBITS, M = 1 << NBITS, NT = 1 << (2 * NBITS);
    static constexpr int E = NA * NB * NT, DB = NB * M;
    static constexpr int popc(int v) { return v == 0 ? 0 : (v & 1) + popc(v >> 1); }
    // complex elements one exchange buffer must hold
    static constexpr int buffer_elems() {
        int need = NA * NB * (NT / 4);                                  // damping: a quarter of the threads write a tile
        for (int o = 0; o < BP::NOPS; ++o) {
            const int kind = BP::ops[o][0], bit = BP::ops[o][1], mask = BP::ops[o][3];
            if (kind == OP_ROTK && BP::ket_pos[bit] >= 5 && popc(mask) * NB * NT > need) need = popc(mask) * NB * NT;
            if (kind == OP_ROTB && BP::bra_pos[bit] >= 5 && popc(mask) * NA * NT > need) need = popc(mask) * NA * NT;
        }
        return need;
    }
    static constexpr int BUF = buffer_elems();
    // Exchanges only involve the two warps whose index differs in one warp bit (or one warp, for lane bits), so they
    // synchronise pairwise with named barriers.  Each warp bit (plus one slot for exchanges inside a warp) owns two
    // buffers used alternately: a warp can then never overwrite data its partner of an earlier exchange still reads.
    static constexpr int NWB = 2 * NBITS > 5 ? 2 * NBITS - 5 : 0;       // warp bits
    static constexpr int NGROUP = NWB + 1;
    static constexpr int PAIRS = NWB > 0 ? 1 << (NWB - 1) : 1;
    // The damping exchange synchronises the two warps that differ in the HIGHER thread bit of the (ket, bra) pair only:
    // correct as long as the lower bit is a lane bit.  Named barriers 1 .. NWB * PAIRS must exist (16 per CTA).
    static constexpr bool damping_pairs_ok() {
        for (int q = 0; q < NBITS; ++q)
            if ((BP::ket_pos[q] < BP::bra_pos[q] ? BP::ket_pos[q] : BP::bra_pos[q]) >= 5) return false;
        return true;
    }
    static_assert(damping_pairs_ok(), "a pseudomode with both thread bits above the lanes needs a 4-warp barrier");
    static_assert(1 + NWB * PAIRS <= 16, "named barrier ids exceed the 16 hardware barriers of a CTA");
    static constexpr size_t smem_bytes(int p_doubles) {
        return (size_t)2 * NGROUP * BUF * sizeof(double2) + (size_t)((p_doubles + 1) & ~1) * sizeof(double) +
               64 * sizeof(double2);
    }
};

// (c, s) rotation of a pair: x' = c x - i s y, y' = c y - i s x
__device__ __forceinline__ void pair_cs(double2& x, double2& y, double c, double s) {
    const double2 nx = make_double2(fma(c, x.x, s * y.y), fma(c, x.y, -s * y.x));
    const double2 ny = make_double2(fma(c, y.x, s * x.y), fma(c, y.y, -s * x.x));
    x = nx;
    y = ny;
}
// tangent form with a partner value: v' = v - i t p
__device__ __forceinline__ void tan_one(double2& v, const double2 p, double t) {
    v.x = fma(t, p.y, v.x);
    v.y = fma(-t, p.x, v.y);
}

__device__ __forceinline__ int remove_bit(int v, int pos) { return ((v >> (pos + 1)) << pos) | (v & ((1 << pos) - 1)); }

// synchronise the threads that exchange through thread bit POS: one warp (lane bit) or a pair of warps (warp bit)
template <int PID, int POS>
__device__ __forceinline__ void pair_sync(int tid) {
    if constexpr (POS < 5) {
        __syncwarp();
    } else {
        constexpr int kk = POS - 5;
        const int id = 1 + kk * Geo<PID>::PAIRS + remove_bit(tid >> 5, kk);
        asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
    }
}

template <int PID, int O>
__device__ __forceinline__ void big_op(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB],
                                       const double2 (&w)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P,
                                       double2* xbuf, int (&which)[Geo<PID>::NGROUP], int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
    constexpr int kind = BP::ops[O][0], a = BP::ops[O][1], b = BP::ops[O][2], mask = BP::ops[O][3], slot = BP::ops[O][4];
    if constexpr (kind == OP_DIAG) {
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) v[i][j] = cmul(v[i][j], w[i][j]);
    } else if constexpr (kind == OP_HOPK) {
        const double2 cs = *reinterpret_cast<const double2*>(P + slot);
#pragma unroll
        for (int j = 0; j < NB; ++j) pair_cs(v[a][j], v[b][j], cs.x, cs.y);
    } else if constexpr (kind == OP_HOPB) {
        const double2 cs = *reinterpret_cast<const double2*>(P + slot);
#pragma unroll
        for (int i = 0; i < NA; ++i) pair_cs(v[i][a], v[i][b], cs.x, -cs.y);
    } else if constexpr (kind == OP_ROTK || kind == OP_ROTB) {
        constexpr bool ket = kind == OP_ROTK;
        constexpr int pos = ket ? BP::ket_pos[a] : BP::bra_pos[a];
        const double t = ket ? P[slot + 1] : -P[slot + 1];
        if constexpr (pos < 5) {
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((mask >> (ket ? i : j)) & 1) {
                        double2 p;
                        p.x = __shfl_xor_sync(0xffffffffu, v[i][j].x, 1 << pos);
                        p.y = __shfl_xor_sync(0xffffffffu, v[i][j].y, 1 << pos);
                        tan_one(v[i][j], p, t);
                    }
        } else {
            constexpr int grp = pos - 5;
            double2* buf = xbuf + (2 * grp + which[grp]) * G::BUF;
            int e = 0;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((mask >> (ket ? i : j)) & 1) buf[(e++) * NT + tid] = v[i][j];
            pair_sync<PID, pos>(tid);
            e = 0;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((mask >> (ket ? i : j)) & 1) tan_one(v[i][j], buf[(e++) * NT + (tid ^ (1 << pos))], t);
            which[grp] ^= 1;
        }
    } else {   // amplitude damping of pseudomode `a`: ket thread bit pa, bra thread bit pb
        constexpr int pa = BP::ket_pos[a], pb = BP::bra_pos[a];
        constexpr int hi = pa > pb ? pa : pb, lo = pa > pb ? pb : pa;
        const double2 cs = *reinterpret_cast<const double2*>(P + slot);       // {cos phi, sin^2 phi}
        const int la = (tid >> pa) & 1, lb = (tid >> pb) & 1;
        const int cid = remove_bit(remove_bit(tid, hi), lo);
        constexpr int grp = hi >= 5 ? hi - 5 : G::NWB;              // exchanges inside a warp use the last buffer pair
        double2* buf = xbuf + (2 * grp + which[grp]) * G::BUF;
        // forward: the (1,1) threads of the bit pair hand their tile to the (0,0) threads; transposed (OP_ADT, the
        // read-out is propagated): the (0,0) threads hand theirs to the (1,1) threads
        constexpr bool fwd = kind == OP_AD;
        const bool giver = fwd ? (la & lb) : !(la | lb), taker = fwd ? !(la | lb) : (la & lb);
        if (giver) {
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) buf[(i * NB + j) * (NT / 4) + cid] = v[i][j];
        }
        pair_sync<PID, hi>(tid);
        const double cc = cs.x * cs.x;
        if (taker) {
            const double keep = fwd ? 1.0 : cc;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const double2 p = buf[(i * NB + j) * (NT / 4) + cid];
                    v[i][j].x = fma(cs.y, p.x, keep * v[i][j].x);
                    v[i][j].y = fma(cs.y, p.y, keep * v[i][j].y);
                }
        } else if (!giver || fwd) {
            const double f = giver ? cc : cs.x;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    v[i][j].x *= f;
                    v[i][j].y *= f;
                }
        }
        which[grp] ^= 1;
    }
}

template <int PID, int... Os>
__device__ __forceinline__ void big_step(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB],
                                         const double2 (&w)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P,
                                         double2* xbuf, int (&which)[Geo<PID>::NGROUP], int tid,
                                         qstd::integer_sequence<int, Os...>) {
    (big_op<PID, Os>(v, w, P, xbuf, which, tid), ...);
}

// Read-outs: per-site sums over the compile-time pairs em[][] = (i, j, site) of the threads with m_a == m_b --
// (KA, KB = KA -+ 1) blocks: the pairs whose configurations differ in `site` (emission Tr[(sum_s mu_s X_s) sigma]);
// (KA, KA) blocks: the diagonal pairs (i, i) with `site` occupied (populations).
template <int PID, int E>
__device__ __forceinline__ void trace_term(const double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB],
                                           double (&re)[BigProg<PID>::NBITS], double (&im)[BigProg<PID>::NBITS]) {
    constexpr int i = BigProg<PID>::em[E][0], j = BigProg<PID>::em[E][1], s = BigProg<PID>::em[E][2];
    re[s] += v[i][j].x;
    im[s] += v[i][j].y;
}
template <int PID, int... Es>
__device__ __forceinline__ void big_trace(const double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB],
                                          double (&re)[BigProg<PID>::NBITS], double (&im)[BigProg<PID>::NBITS],
                                          qstd::integer_sequence<int, Es...>) {
    (trace_term<PID, Es>(v, re, im), ...);
}

// (table entries are copied into constexpr locals: indexing a constexpr array with anything else odr-uses it in device code)
template <int PID, int Q>
__device__ __forceinline__ int ket_bit(int tid) {
    constexpr int p = BigProg<PID>::ket_pos[Q];
    return ((tid >> p) & 1) << Q;
}
template <int PID, int Q>
__device__ __forceinline__ int bra_bit(int tid) {
    constexpr int p = BigProg<PID>::bra_pos[Q];
    return ((tid >> p) & 1) << Q;
}
template <int PID, int... Qs>
__device__ __forceinline__ void thread_modes(int tid, int& ma, int& mb, qstd::integer_sequence<int, Qs...>) {
    ma = (ket_bit<PID, Qs>(tid) | ...);
    mb = (bra_bit<PID, Qs>(tid) | ...);
}

template <int PID>
__global__ void __launch_bounds__(1 << (2 * BigProg<PID>::NBITS)) evolve_rrc_kernel(const Params k) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB, NBITS = G::NBITS, M = G::M, NT = G::NT, DB = G::DB, E = G::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const qsx_rrc_args& a = k.a;
    double2* xbuf = reinterpret_cast<double2*>(smem_raw);
    double2* red = xbuf + 2 * G::NGROUP * G::BUF;
    double* P = reinterpret_cast<double*>(red + 64);
    const int tid = threadIdx.x;
    int ma, mb;
    thread_modes<PID>(tid, ma, mb, qstd::make_integer_sequence<int, NBITS>{});
    double mu[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) mu[s] = a.mu[s];

    for (int inst = blockIdx.x; inst < a.n_inst; inst += gridDim.x) {
        const int set = a.inst_set ? a.inst_set[inst] : 0;
        const int src = a.inst_src ? a.inst_src[inst] : inst;
        const double* Pg = a.params + (size_t)set * a.param_stride;
        __syncthreads();
        for (int i = tid; i < a.param_stride; i += NT) P[i] = Pg[i];
        __syncthreads();
        double2 w[NA][NB], v[NA][NB];
        {
            const double2* Wc = reinterpret_cast<const double2*>(P);
            const double2 gm = cmul_conj(reinterpret_cast<const double2*>(P + 2 * NA * NB)[ma],
                                         reinterpret_cast<const double2*>(P + 2 * NA * NB + 2 * M)[mb]);
            const double2* sin = reinterpret_cast<const double2*>(a.sigma_in) + (size_t)src * E;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    w[i][j] = cmul(Wc[i * NB + j], gm);
                    v[i][j] = sin[(size_t)((i << NBITS) | ma) * DB + ((j << NBITS) | mb)];
                }
        }
        int which[G::NGROUP];
#pragma unroll
        for (int g = 0; g < G::NGROUP; ++g) which[g] = 0;
        int slot = 0;
        int next_emit = a.n_emit > 0 ? a.emit_steps[0] : -1;
        int after_next = a.n_emit > 1 ? a.emit_steps[1] : -1;
        for (int step = 0; step <= a.n_steps && slot < a.n_emit; ++step) {
            if (step) big_step<PID>(v, w, P, xbuf, which, tid, qstd::make_integer_sequence<int, BP::NOPS>{});
            if (step == next_emit) {
                if (a.emit_trace) {
                    double re[NBITS], im[NBITS];                 // one pseudomode per site: NBITS = number of sites
#pragma unroll
                    for (int s = 0; s < NBITS; ++s) re[s] = im[s] = 0.0;
                    if constexpr (BP::NEM > 0) {
                        if (ma == mb) big_trace<PID>(v, re, im, qstd::make_integer_sequence<int, BP::NEM>{});
                    }
#pragma unroll
                    for (int s = 0; s < NBITS; ++s) {
#pragma unroll
                        for (int off = 16; off; off >>= 1) {
                            re[s] += __shfl_xor_sync(0xffffffffu, re[s], off);
                            im[s] += __shfl_xor_sync(0xffffffffu, im[s], off);
                        }
                        if ((tid & 31) == 0) red[(tid >> 5) * NBITS + s] = make_double2(re[s], im[s]);
                    }
                    __syncthreads();
                    if (tid == 0) {
                        double sr = 0.0, si = 0.0;
                        for (int s = 0; s < NBITS; ++s) {
                            double pr = 0.0, pi = 0.0;
                            for (int wv = 0; wv < (NT + 31) / 32; ++wv) {
                                pr += red[wv * NBITS + s].x;
                                pi += red[wv * NBITS + s].y;
                            }
                            if (a.emit_trace == 2) a.out_obs[((size_t)inst * a.n_emit + slot) * NBITS + s] = pr;
                            sr = fma(mu[s], pr, sr);
                            si = fma(mu[s], pi, si);
                        }
                        if (a.emit_trace == 1)
                            reinterpret_cast<double2*>(a.out_obs)[(size_t)inst * a.n_emit + slot] = make_double2(sr, si);
                    }
                    __syncthreads();
                }
                if (a.out_snap) {
                    double2* dst = reinterpret_cast<double2*>(a.out_snap) + ((size_t)inst * a.n_emit + slot) * E;
#pragma unroll
                    for (int i = 0; i < NA; ++i)
#pragma unroll
                        for (int j = 0; j < NB; ++j) dst[(size_t)((i << NBITS) | ma) * DB + ((j << NBITS) | mb)] = v[i][j];
                }
                ++slot;
                next_emit = after_next;
                after_next = slot + 1 < a.n_emit ? a.emit_steps[slot + 1] : -1;
            }
        }
    }
}

#ifndef __CUDACC_RTC__
template <int PID>
bool matches(const qsx_rrc_args& a) {
    using BP = BigProg<PID>;
    if (a.n_signature != BP::NSIG) return false;
    for (int i = 0; i < BP::NSIG; ++i)
        if (a.signature[i] != BP::sig[i]) return false;
    return true;
}

// index of the compile-time program with this signature, or -1
template <int PID>
int find(const qsx_rrc_args& a) {
    if constexpr (PID >= kNumBigProgs) {
        return -1;
    } else {
        return matches<PID>(a) ? PID : find<PID + 1>(a);
    }
}

// 0: launched, 1: no compile-time program matches, < 0: CUDA error code (negated)
template <int PID>
int launch(const qsx_rrc_args& a, int sms, int smem_max, cudaStream_t stream) {
    if constexpr (PID >= kNumBigProgs) {
        return 1;
    } else {
        if (!matches<PID>(a)) return launch<PID + 1>(a, sms, smem_max, stream);
        using G = Geo<PID>;
        const size_t smem = G::smem_bytes(a.param_stride);
        if (smem > (size_t)smem_max) return 1;
        cudaError_t e = cudaFuncSetAttribute(evolve_rrc_kernel<PID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
        int per_sm = (int)((size_t)(228 * 1024) / (smem + 1024));
        if (per_sm > 2048 / G::NT) per_sm = 2048 / G::NT;
        if (per_sm < 1) per_sm = 1;
        int grid = sms * per_sm;
        if (grid > a.n_inst) grid = a.n_inst;
        Params k;
        k.a = a;
        evolve_rrc_kernel<PID><<<grid, G::NT, smem, stream>>>(k);
        e = cudaGetLastError();
        return e == cudaSuccess ? 0 : -(int)e;
    }
}

#endif  // !__CUDACC_RTC__

}  // namespace rrc
