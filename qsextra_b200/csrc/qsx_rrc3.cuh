// qsx_rrc3.cuh -- third form of the CTA-wide register-resident kernel: amplitude damping without any exchange.
//
// The throughput form (qsx_rrc2.cuh) is bound by the shuffle / shared-memory data pipe, and more than half of its
// wavefronts are the four damping exchanges of a step ((1,1) -> (0,0) of every pseudomode's (ket bit, bra bit) pair).
// On the 4-dimensional space of one pair, damping  (0,0) += s^2 (1,1);  (0,1), (1,0) *= c;  (1,1) *= c^2  has the
// eigenvectors (0,0), (0,1), (1,0) and (1,1) - (0,0) with eigenvalues 1, c, c, c^2.  This kernel keeps, for every damped
// pseudomode q, the coordinates
//        u(0,0) = sigma(0,0) + sigma(1,1),   u(0,1), u(1,0), u(1,1) = sigma(0,1), sigma(1,0), sigma(1,1)
// (the transform T = prod_q T_q, T_q: (0,0) += (1,1)), in which EVERY damping is diagonal and folds into the per-thread
// factor of the next step's diagonal multiply.  Diagonal tables (the (0,0) and (1,1) entries of a pair carry the same
// phase), exciton hops (configurations only) commute with T and are unchanged.  A pseudomode rotation becomes
//   ket:  u(0,0) -= i t (u(1,0) + u(0,1));  u(1,0) -= i t (u(0,0) - u(1,1));  u(0,1) -= i t u(1,1);  u(1,1) -= i t u(0,1)
//   bra:  the same with the roles of the two bits exchanged and t -> -t
// i.e. next to the value across its own bit it needs the value across the partner bit of the pair (for half of the
// threads).  All ket rotations of a step share one exchange round, all bra rotations a second one (their rows /
// columns are disjoint for the eligible programs): 2 CTA barriers and 2816 wavefronts per step of the 64 x 64 block
// instead of 4 and 3367.  The transform is applied when an instance is loaded and undone (in place, then redone) around
// every emission, so this form serves levels that emit rarely (snapshots every 100 steps in the t2 level of a 2-D
// spectrum); programs that emit often keep the second form.
#pragma once

namespace rrc3 {

using rrc::BigProg;
using rrc::kNumBigProgs;
using rrc::OP_AD;
using rrc::OP_ADT;
using rrc::OP_DIAG;
using rrc::OP_HOPB;
using rrc::OP_HOPK;
using rrc::OP_ROTB;
using rrc::OP_ROTK;
using rrc::Params;

template <int PID>
struct Geo {
    using BP = BigProg<PID>;
    static constexpr int NA = BP::NA, NB = BP::NB, NBITS = BP::NBITS, M = 1 << NBITS, NT = 1 << (2 * NBITS);
    static constexpr int E = NA * NB * NT, DB = NB * M;
    static __host__ __device__ constexpr int popc(int v) { return v == 0 ? 0 : (v & 1) + popc(v >> 1); }
    static __host__ __device__ constexpr bool damped(int q) {
        for (int o = 0; o < BP::NOPS; ++o)
            if (BP::ops[o][0] == OP_AD && BP::ops[o][1] == q) return true;
        return false;
    }
    // rotations whose partner across a WARP bit is fetched through shared memory (after the round's barrier)
    static __host__ __device__ constexpr bool k_deferred(int o) {
        return BP::ops[o][0] == OP_ROTK && damped(BP::ops[o][1]) && BP::bra_pos[BP::ops[o][1]] >= 5;
    }
    static __host__ __device__ constexpr bool b_deferred(int o) {
        return BP::ops[o][0] == OP_ROTB && BP::bra_pos[BP::ops[o][1]] >= 5;
    }
    static __host__ __device__ constexpr int k_offset(int upto) {
        int off = 0;
        for (int o = 0; o < upto; ++o)
            if (k_deferred(o)) off += popc(BP::ops[o][3]) * NB * NT;
        return off;
    }
    static __host__ __device__ constexpr int b_offset(int upto) {
        int off = 0;
        for (int o = 0; o < upto; ++o)
            if (b_deferred(o)) off += popc(BP::ops[o][3]) * NA * NT;
        return off;
    }
    static constexpr int K_ELEMS = k_offset(BP::NOPS), B_ELEMS = b_offset(BP::NOPS), AD_ELEMS = NA * NB * (NT / 4);
    static constexpr int X_ELEMS = B_ELEMS > AD_ELEMS ? B_ELEMS : AD_ELEMS;     // bra round; load / emission exchanges
    static constexpr int BUF = K_ELEMS + X_ELEMS;
    // forward program [DIAG][hops, rotations][dampings], tile <= 16 elements, ket bits on lanes, and -- because the
    // rotations of one side share an exchange round -- pairwise disjoint rows among the ket rotations and pairwise
    // disjoint columns among the bra rotations (true for blocks with at most one exciton per side)
    static constexpr bool eligible() {
        if (NA * NB > 16 || BP::ops[0][0] != OP_DIAG) return false;
        bool in_tail = false, any_damped = false, rotating = false;
        int rows = 0, cols = 0;
        for (int o = 1; o < BP::NOPS; ++o) {
            const int kind = BP::ops[o][0];
            if (kind == OP_ADT || kind == OP_DIAG) return false;
            if (kind == OP_ROTK || kind == OP_ROTB) rotating = true;
            if ((kind == OP_HOPK || kind == OP_HOPB) && rotating) return false;     // hops before every rotation
            if (kind == OP_ROTK) {
                if (BP::ket_pos[BP::ops[o][1]] >= 5 || (rows & BP::ops[o][3])) return false;
                rows |= BP::ops[o][3];
            }
            if (kind == OP_ROTB) {
                if (cols & BP::ops[o][3]) return false;
                cols |= BP::ops[o][3];
            }
            if (kind == OP_AD) in_tail = any_damped = true;
            else if (in_tail) return false;
        }
        return any_damped;
    }
    static constexpr int MIN_CTAS = NT >= 256 ? 2 : (512 / NT > 8 ? 8 : 512 / NT);
    static constexpr size_t smem_bytes(int p_doubles) {
        return (size_t)BUF * sizeof(double2) + (size_t)((p_doubles + 1) & ~1) * sizeof(double) +
               64 * sizeof(double2) + (size_t)BP::NOPS * sizeof(double2);
    }
};

__device__ __forceinline__ double2 shfl_xor2(const double2 v, int mask) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}

// Round 1 of a step, in program order: diagonal multiply, hops, every ket rotation whose partners are all reachable by
// shuffles, every bra rotation over a lane bit; ket rotations that need a partner across a warp bit write their rows.
template <int PID, int O>
__device__ __forceinline__ void op_first(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double2 g, const double* P,
                                         const double2* H, double2* kbuf, int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
    constexpr int kind = BP::ops[O][0], a = BP::ops[O][1], b = BP::ops[O][2], mask = BP::ops[O][3], slot = BP::ops[O][4];
    if constexpr (kind == OP_DIAG) {
        const double2* Wc = reinterpret_cast<const double2*>(P);
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) v[i][j] = cmul(v[i][j], cmul(Wc[i * NB + j], g));
    } else if constexpr (kind == OP_HOPK) {
        const double2 ts = H[O];
#pragma unroll
        for (int j = 0; j < NB; ++j) rrc2::pair_shear(v[a][j], v[b][j], ts.x, ts.y);
    } else if constexpr (kind == OP_HOPB) {
        const double2 ts = H[O];
#pragma unroll
        for (int i = 0; i < NA; ++i) rrc2::pair_shear(v[i][a], v[i][b], -ts.x, -ts.y);
    } else if constexpr (kind == OP_ROTK) {
        constexpr int pk = BP::ket_pos[a], pb = BP::bra_pos[a];
        if constexpr (G::k_deferred(O)) {
            // only the threads with the partner bit CLEAR need the value across it, so only those with the bit SET
            // write (a warp bit: whole warps write, whole warps read)
            if ((tid >> pb) & 1) {
                double2* buf = kbuf + G::k_offset(O);
                int e = 0;
#pragma unroll
                for (int i = 0; i < NA; ++i)
#pragma unroll
                    for (int j = 0; j < NB; ++j)
                        if ((mask >> i) & 1) buf[(e++) * NT + tid] = v[i][j];
            }
        } else {
            const double t = P[slot + 1];
            const int la = (tid >> pk) & 1, lb = (tid >> pb) & 1;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((mask >> i) & 1) {
                        double2 p = shfl_xor2(v[i][j], 1 << pk);
                        if constexpr (G::damped(a)) {
                            const double2 q = shfl_xor2(v[i][j], 1 << pb);
                            const double w = lb ? 0.0 : (la ? -1.0 : 1.0);
                            p.x = fma(w, q.x, p.x);
                            p.y = fma(w, q.y, p.y);
                        }
                        rrc::tan_one(v[i][j], p, t);
                    }
        }
    }
}

// after the first barrier: the deferred ket rotations
template <int PID, int O>
__device__ __forceinline__ void op_k_deferred(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P,
                                              const double2* kbuf, int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    if constexpr (G::k_deferred(O)) {
        constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
        constexpr int a = BP::ops[O][1], mask = BP::ops[O][3], slot = BP::ops[O][4];
        constexpr int pk = BP::ket_pos[a], pb = BP::bra_pos[a];
        const double t = P[slot + 1];
        const int la = (tid >> pk) & 1, lb = (tid >> pb) & 1;
        const double w = lb ? 0.0 : (la ? -1.0 : 1.0);
        const double2* buf = kbuf + G::k_offset(O);
        int e = 0;
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if ((mask >> i) & 1) {
                    double2 p = shfl_xor2(v[i][j], 1 << pk);
                    if (!lb) {
                        const double2 q = buf[e * NT + (tid ^ (1 << pb))];
                        p.x = fma(w, q.x, p.x);
                        p.y = fma(w, q.y, p.y);
                    }
                    ++e;
                    rrc::tan_one(v[i][j], p, t);
                }
    }
}

// Round 2: bra rotations.  Over a lane bit: immediately; over a warp bit: write the columns, barrier, read.
template <int PID, int O>
__device__ __forceinline__ void op_b_first(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P, double2* bbuf,
                                           int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    if constexpr (BP::ops[O][0] == OP_ROTB) {
        constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
        constexpr int a = BP::ops[O][1], mask = BP::ops[O][3], slot = BP::ops[O][4];
        constexpr int pk = BP::ket_pos[a], pb = BP::bra_pos[a];
        if constexpr (G::b_deferred(O)) {
            double2* buf = bbuf + G::b_offset(O);
            int e = 0;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((mask >> j) & 1) buf[(e++) * NT + tid] = v[i][j];
        } else {
            const double t = -P[slot + 1];
            const int la = (tid >> pk) & 1, lb = (tid >> pb) & 1;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((mask >> j) & 1) {
                        double2 p = shfl_xor2(v[i][j], 1 << pb);
                        if constexpr (G::damped(a)) {
                            const double2 q = shfl_xor2(v[i][j], 1 << pk);
                            const double w = la ? 0.0 : (lb ? -1.0 : 1.0);
                            p.x = fma(w, q.x, p.x);
                            p.y = fma(w, q.y, p.y);
                        }
                        rrc::tan_one(v[i][j], p, t);
                    }
        }
    }
}

template <int PID, int O>
__device__ __forceinline__ void op_b_deferred(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P,
                                              const double2* bbuf, int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    if constexpr (G::b_deferred(O)) {
        constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
        constexpr int a = BP::ops[O][1], mask = BP::ops[O][3], slot = BP::ops[O][4];
        constexpr int pk = BP::ket_pos[a], pb = BP::bra_pos[a];
        const double t = -P[slot + 1];
        const int la = (tid >> pk) & 1, lb = (tid >> pb) & 1;
        const double w = la ? 0.0 : (lb ? -1.0 : 1.0);
        const double2* buf = bbuf + G::b_offset(O);
        int e = 0;
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if ((mask >> j) & 1) {
                    double2 p = buf[(e++) * NT + (tid ^ (1 << pb))];
                    if constexpr (G::damped(a)) {
                        const double2 q = shfl_xor2(v[i][j], 1 << pk);
                        p.x = fma(w, q.x, p.x);
                        p.y = fma(w, q.y, p.y);
                    }
                    rrc::tan_one(v[i][j], p, t);
                }
    }
}

template <int PID, int... Os>
__device__ __forceinline__ void step3(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double2 g, const double* P,
                                      const double2* H, double2* xbuf, int tid, qstd::integer_sequence<int, Os...>) {
    using G = Geo<PID>;
    double2* kbuf = xbuf;
    double2* bbuf = xbuf + G::K_ELEMS;
    (op_first<PID, Os>(v, g, P, H, kbuf, tid), ...);
    if constexpr (G::K_ELEMS > 0) {
        __syncthreads();
        (op_k_deferred<PID, Os>(v, P, kbuf, tid), ...);
    }
    (op_b_first<PID, Os>(v, P, bbuf, tid), ...);
    if constexpr (G::B_ELEMS > 0) {
        __syncthreads();
        (op_b_deferred<PID, Os>(v, P, bbuf, tid), ...);
    }
    // a region is rewritten in the next step: every reader of this step must have passed a barrier in between, which the
    // other round's barrier provides -- unless only one of the two rounds goes through shared memory
    if constexpr ((G::K_ELEMS > 0) != (G::B_ELEMS > 0)) __syncthreads();
}

// T_q (sign +1) or its inverse (sign -1) for every damped pseudomode: (0,0) += sign (1,1) of the bit pair, through the
// exchange region, CTA barriers on both sides of every round (load and emissions only)
template <int PID, int O>
__device__ __forceinline__ void pair_shift(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], double sign, double2* buf, int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    if constexpr (BP::ops[O][0] == OP_AD) {
        constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
        constexpr int a = BP::ops[O][1];
        constexpr int pa = BP::ket_pos[a], pb = BP::bra_pos[a];
        constexpr int hi = pa > pb ? pa : pb, lo = pa > pb ? pb : pa;
        const int la = (tid >> pa) & 1, lb = (tid >> pb) & 1;
        const int cid = rrc::remove_bit(rrc::remove_bit(tid, hi), lo);
        __syncthreads();
        if (la & lb) {
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) buf[(i * NB + j) * (NT / 4) + cid] = v[i][j];
        }
        __syncthreads();
        if (!(la | lb)) {
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const double2 p = buf[(i * NB + j) * (NT / 4) + cid];
                    v[i][j].x = fma(sign, p.x, v[i][j].x);
                    v[i][j].y = fma(sign, p.y, v[i][j].y);
                }
        }
    }
}
template <int PID, int... Os>
__device__ __forceinline__ void pair_shifts(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], double sign, double2* buf,
                                            int tid, qstd::integer_sequence<int, Os...>) {
    (pair_shift<PID, Os>(v, sign, buf, tid), ...);
}

template <int PID>
__global__ void __launch_bounds__(1 << (2 * BigProg<PID>::NBITS), Geo<PID>::MIN_CTAS) evolve_rrc3_kernel(const Params k) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB, NBITS = G::NBITS, M = G::M, NT = G::NT, DB = G::DB, E = G::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const qsx_rrc_args& a = k.a;
    double2* xbuf = reinterpret_cast<double2*>(smem_raw);
    double2* red = xbuf + G::BUF;
    double2* H = red + 64;
    double* P = reinterpret_cast<double*>(H + BP::NOPS);
    double2* shift_buf = xbuf + G::K_ELEMS;                      // the bra-round region doubles as the T_q exchange buffer
    const int tid = threadIdx.x;
    int ma, mb;
    rrc::thread_modes<PID>(tid, ma, mb, qstd::make_integer_sequence<int, NBITS>{});
    constexpr auto all_ops = qstd::make_integer_sequence<int, BP::NOPS>{};

    for (int inst = blockIdx.x; inst < a.n_inst; inst += gridDim.x) {
        const int set = a.inst_set ? a.inst_set[inst] : 0;
        const int src = a.inst_src ? a.inst_src[inst] : inst;
        const double* Pg = a.params + (size_t)set * a.param_stride;
        __syncthreads();
        for (int i = tid; i < a.param_stride; i += NT) P[i] = Pg[i];
        __syncthreads();
        if (tid == 0) rrc2::hop_table<PID>(P, H, all_ops);
        const double2 g0 = cmul_conj(reinterpret_cast<const double2*>(P + 2 * NA * NB)[ma],
                                     reinterpret_cast<const double2*>(P + 2 * NA * NB + 2 * M)[mb]);
        const double dd = rrc2::damp_scales<PID>(P, tid, all_ops);
        const double inv_dd = 1.0 / dd;
        const double2 g1 = make_double2(g0.x * dd, g0.y * dd);
        double2 v[NA][NB];
        {
            const double2* sin = reinterpret_cast<const double2*>(a.sigma_in) + (size_t)src * E;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) v[i][j] = sin[(size_t)((i << NBITS) | ma) * DB + ((j << NBITS) | mb)];
        }
        pair_shifts<PID>(v, +1.0, shift_buf, tid, all_ops);     // sigma -> u = T sigma
        __syncthreads();
        int slot = 0;
        int next_emit = a.n_emit > 0 ? a.emit_steps[0] : -1;
        int after_next = a.n_emit > 1 ? a.emit_steps[1] : -1;
        for (int step = 0; step <= a.n_steps && slot < a.n_emit; ++step) {
            if (step) step3<PID>(v, step == 1 ? g0 : g1, P, H, xbuf, tid, all_ops);
            if (step == next_emit) {
                // back to sigma in place: the pending damping scalings (diagonal in these coordinates), then T^-1
                const double out_scale = step ? dd : 1.0;
#pragma unroll
                for (int i = 0; i < NA; ++i)
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        v[i][j].x *= out_scale;
                        v[i][j].y *= out_scale;
                    }
                pair_shifts<PID>(v, -1.0, shift_buf, tid, all_ops);
                if (a.emit_trace) {
                    double re[NBITS], im[NBITS];
#pragma unroll
                    for (int s = 0; s < NBITS; ++s) re[s] = im[s] = 0.0;
                    if constexpr (BP::NEM > 0) {
                        if (ma == mb) rrc::big_trace<PID>(v, re, im, qstd::make_integer_sequence<int, BP::NEM>{});
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
                            sr = fma(a.mu[s], pr, sr);
                            si = fma(a.mu[s], pi, si);
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
                if (slot < a.n_emit) {                           // forward again: T, and the scalings stay pending
                    pair_shifts<PID>(v, +1.0, shift_buf, tid, all_ops);
                    const double back = step ? inv_dd : 1.0;
#pragma unroll
                    for (int i = 0; i < NA; ++i)
#pragma unroll
                        for (int j = 0; j < NB; ++j) {
                            v[i][j].x *= back;
                            v[i][j].y *= back;
                        }
                    __syncthreads();
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same propagation scheduled by SEGMENTS (the steps between two emissions) for launches that only write snapshots
// and hold more instances than the GPU keeps resident.  With one CTA per instance from start to end, 512 instances on
// 296 resident CTAs take two full rounds although the second round is only 73 % full.  Here the CTAs claim
// (instance, segment) pairs from one counter, in the order segment-major / instance-minor, and continue an instance
// from the snapshot its previous segment left in `out_snap` (the transform to the damping-diagonal coordinates is
// applied at every emission anyway, so a reload costs one pass over 64 KiB): every CTA ends up with the same number of
// segments.  A claimed pair waits for `progress[inst]` to reach its segment; pairs are claimed in dependency order by
// CTAs that are RUNNING, so the oldest unfinished pair never waits for a CTA that has not started (no deadlock when
// other kernels keep some SMs busy).  Same arithmetic per segment as the kernel above.
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int PID>
__global__ void __launch_bounds__(1 << (2 * BigProg<PID>::NBITS), Geo<PID>::MIN_CTAS) evolve_rrc3_segments_kernel(const Params k) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB, NBITS = G::NBITS, M = G::M, NT = G::NT, DB = G::DB, E = G::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int claimed;
    const qsx_rrc_args& a = k.a;
    double2* xbuf = reinterpret_cast<double2*>(smem_raw);
    double2* red = xbuf + G::BUF;
    double2* H = red + 64;
    double* P = reinterpret_cast<double*>(H + BP::NOPS);
    double2* shift_buf = xbuf + G::K_ELEMS;
    const int tid = threadIdx.x;
    int ma, mb;
    rrc::thread_modes<PID>(tid, ma, mb, qstd::make_integer_sequence<int, NBITS>{});
    constexpr auto all_ops = qstd::make_integer_sequence<int, BP::NOPS>{};
    int* progress = a.progress;
    int* counter = a.progress + a.n_inst;
    const long long n_pairs = (long long)a.n_inst * a.n_emit;

    for (;;) {
        __syncthreads();
        if (tid == 0) claimed = atomicAdd(counter, 1);
        __syncthreads();
        const int pair = claimed;
        if (pair >= n_pairs) break;
        const int inst = pair % a.n_inst, seg = pair / a.n_inst;
        const int set = a.inst_set ? a.inst_set[inst] : 0;
        const int src = a.inst_src ? a.inst_src[inst] : inst;
        const double* Pg = a.params + (size_t)set * a.param_stride;
        for (int i = tid; i < a.param_stride; i += NT) P[i] = Pg[i];
        if (seg > 0 && tid == 0)
            while (ld_acquire(progress + inst) < seg) __nanosleep(256);
        __syncthreads();
        if (tid == 0) rrc2::hop_table<PID>(P, H, all_ops);
        const double2 g0 = cmul_conj(reinterpret_cast<const double2*>(P + 2 * NA * NB)[ma],
                                     reinterpret_cast<const double2*>(P + 2 * NA * NB + 2 * M)[mb]);
        const double dd = rrc2::damp_scales<PID>(P, tid, all_ops);
        const double2 g1 = make_double2(g0.x * dd, g0.y * dd);
        double2* snaps = reinterpret_cast<double2*>(a.out_snap) + (size_t)inst * a.n_emit * E;
        double2 v[NA][NB];
        {
            // (written by another SM a moment ago: read through L2)
            const double2* from = seg == 0 ? reinterpret_cast<const double2*>(a.sigma_in) + (size_t)src * E
                                           : snaps + (size_t)(seg - 1) * E;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) v[i][j] = __ldcg(from + (size_t)((i << NBITS) | ma) * DB + ((j << NBITS) | mb));
        }
        pair_shifts<PID>(v, +1.0, shift_buf, tid, all_ops);     // sigma -> u = T sigma
        __syncthreads();
        const int first = seg == 0 ? 0 : a.emit_steps[seg - 1], last = a.emit_steps[seg];
        for (int step = first + 1; step <= last; ++step) step3<PID>(v, step == first + 1 ? g0 : g1, P, H, xbuf, tid, all_ops);
        const double out_scale = last > first ? dd : 1.0;       // the pending damping scalings, then T^-1
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                v[i][j].x *= out_scale;
                v[i][j].y *= out_scale;
            }
        pair_shifts<PID>(v, -1.0, shift_buf, tid, all_ops);
        {
            double2* dst = snaps + (size_t)seg * E;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) dst[(size_t)((i << NBITS) | ma) * DB + ((j << NBITS) | mb)] = v[i][j];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(progress + inst, seg + 1);
    }
}

#ifndef __CUDACC_RTC__
// 0: launched, 1: no eligible compile-time program, < 0: CUDA error code (negated)
template <int PID>
int launch(const qsx_rrc_args& a, int sms, int smem_max, cudaStream_t stream) {
    if constexpr (PID >= kNumBigProgs) {
        return 1;
    } else {
        if (!rrc::matches<PID>(a)) return launch<PID + 1>(a, sms, smem_max, stream);
        if constexpr (!Geo<PID>::eligible()) {
            return 1;
        } else {
            using G = Geo<PID>;
            const size_t smem = G::smem_bytes(a.param_stride);
            if (smem > (size_t)smem_max) return 1;
            cudaError_t e = cudaFuncSetAttribute(evolve_rrc3_kernel<PID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return -(int)e;
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, evolve_rrc3_kernel<PID>, G::NT, smem);
            if (e != cudaSuccess) return -(int)e;
            if (per_sm < 1) per_sm = 1;
            int grid = sms * per_sm;
            Params k;
            k.a = a;
            if (a.progress && a.out_snap && !a.emit_trace && a.n_emit > 1 && a.n_inst > grid) {
                // more instances than resident CTAs and nothing but snapshots to write: schedule by segments
                e = cudaFuncSetAttribute(evolve_rrc3_segments_kernel<PID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return -(int)e;
                evolve_rrc3_segments_kernel<PID><<<grid, G::NT, smem, stream>>>(k);
                e = cudaGetLastError();
                return e == cudaSuccess ? 0 : -(int)e;
            }
            if (grid > a.n_inst) grid = a.n_inst;
            evolve_rrc3_kernel<PID><<<grid, G::NT, smem, stream>>>(k);
            e = cudaGetLastError();
            return e == cudaSuccess ? 0 : -(int)e;
        }
    }
}
#endif  // !__CUDACC_RTC__

}  // namespace rrc3
