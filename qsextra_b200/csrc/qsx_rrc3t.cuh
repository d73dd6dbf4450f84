// qsx_rrc3t.cuh -- the damping-diagonal form (qsx_rrc3.cuh) for TRANSPOSED steps: the read-out propagations of the
// Heisenberg picture, one instance walking every step alone, where the dependent chain of a step is what counts.
//
// With respect to the pairing <W, sigma> the transposed damping of a (ket bit, bra bit) pair is
//        W'(1,1) = c^2 W(1,1) + s^2 W(0,0),   W'(0,1), W'(1,0) *= c,   W'(0,0) = W(0,0),
// diagonal in the coordinates  z(1,1) = W(1,1) - W(0,0),  z = W elsewhere  (z = T^-T W for the T of the forward form):
// eigenvalues 1, c, c, c^2 again.  A transposed step is  [dampings][rotations][hops][diagonal]; with v = dd . z kept in
// registers (dd = the product of the damping eigenvalues of the thread's bit pattern) every step is
//        v <- (dd D) H R v,
// i.e. the dampings fold into the diagonal multiply that ENDS the step, and an emission divides by dd on the way out.
// Diagonal tables and exciton hops commute with the transform.  A pseudomode rotation becomes (t = tan, old values on
// the right):
//   ket:  z(0,0) -= i t z(1,0);  z(1,0) -= i t z(0,0);  z(0,1) -= i t (z(1,1) + z(0,0));  z(1,1) -= i t (z(0,1) - z(1,0))
//   bra:  the roles of the two bits exchanged and t -> -t
// -- the threads with the PARTNER bit set also need the value across that bit.  All rotations of a step commute, so the
// bra rotations go first and share one exchange round (their columns are disjoint in the eligible programs); the ket
// rotations whose partner bit is a warp bit share a second round when their rows are disjoint (one exciton on the ket
// side), else each takes its own round in program order (two excitons: every row belongs to two rotations, and a
// rotation must see what the earlier one left).  2 - 4 CTA barriers and no damping exchange per step, where the first
// form (qsx_rrc.cuh) has 7 pairwise barriers and 4 tile exchanges.  z <- W at load, W <- z (in place, then back) around
// every emission: the read-outs emit every ~15 steps.  Shipped programs only.
#pragma once

namespace rrc3t {

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
using rrc3::shfl_xor2;

template <int PID>
struct Geo {
    using BP = BigProg<PID>;
    static constexpr int NA = BP::NA, NB = BP::NB, NBITS = BP::NBITS, M = 1 << NBITS, NT = 1 << (2 * NBITS);
    static constexpr int E = NA * NB * NT, DB = NB * M;
    static __host__ __device__ constexpr int popc(int v) { return v == 0 ? 0 : (v & 1) + popc(v >> 1); }
    static __host__ __device__ constexpr bool damped(int q) {
        for (int o = 0; o < BP::NOPS; ++o)
            if (BP::ops[o][0] == OP_ADT && BP::ops[o][1] == q) return true;
        return false;
    }
    // rotations that fetch a value across a WARP bit through shared memory: a ket rotation its partner (the bra bit of
    // the pair, damped pseudomodes only), a bra rotation the value across its own bit
    static __host__ __device__ constexpr bool k_deferred(int o) {
        return BP::ops[o][0] == OP_ROTK && damped(BP::ops[o][1]) && BP::bra_pos[BP::ops[o][1]] >= 5;
    }
    static __host__ __device__ constexpr bool b_deferred(int o) {
        return BP::ops[o][0] == OP_ROTB && BP::bra_pos[BP::ops[o][1]] >= 5;
    }
    static __host__ __device__ constexpr int k_offset(int upto) {          // only the threads with the bit CLEAR write
        int off = 0;
        for (int o = 0; o < upto; ++o)
            if (k_deferred(o)) off += popc(BP::ops[o][3]) * NB * (NT / 2);
        return off;
    }
    static __host__ __device__ constexpr int b_offset(int upto) {
        int off = 0;
        for (int o = 0; o < upto; ++o)
            if (b_deferred(o)) off += popc(BP::ops[o][3]) * NA * NT;
        return off;
    }
    static __host__ __device__ constexpr bool ket_rows_disjoint() {
        int rows = 0;
        for (int o = 0; o < BP::NOPS; ++o)
            if (BP::ops[o][0] == OP_ROTK) {
                if (rows & BP::ops[o][3]) return false;
                rows |= BP::ops[o][3];
            }
        return true;
    }
    static __host__ __device__ constexpr int n_k_deferred() {
        int n = 0;
        for (int o = 0; o < BP::NOPS; ++o) n += k_deferred(o) ? 1 : 0;
        return n;
    }
    static constexpr int K_ELEMS = k_offset(BP::NOPS), B_ELEMS = b_offset(BP::NOPS), AD_ELEMS = NA * NB * (NT / 4);
    static constexpr int X_ELEMS = B_ELEMS > AD_ELEMS ? B_ELEMS : AD_ELEMS;     // bra round; load / emission exchanges
    static constexpr int BUF = K_ELEMS + X_ELEMS;
    // CTA barriers of one step: one for the bra round, one per ket round
    static constexpr int K_ROUNDS = n_k_deferred() == 0 ? 0 : (ket_rows_disjoint() ? 1 : n_k_deferred());
    static constexpr int ROUNDS = (B_ELEMS > 0 ? 1 : 0) + K_ROUNDS;
    // transposed program [dampings][rotations][hops][DIAG]; ket bits on lanes; bra rotations on pairwise disjoint columns
    static constexpr bool eligible() {
        if (NA * NB > 24 || BP::ops[BP::NOPS - 1][0] != OP_DIAG) return false;
        bool rotating = false, hopping = false, any_damped = false;
        int cols = 0;
        for (int o = 0; o < BP::NOPS - 1; ++o) {
            const int kind = BP::ops[o][0];
            if (kind == OP_AD || kind == OP_DIAG) return false;
            if (kind == OP_ADT) {
                if (rotating || hopping) return false;
                any_damped = true;
            }
            if (kind == OP_ROTK || kind == OP_ROTB) {
                if (hopping) return false;
                rotating = true;
            }
            if (kind == OP_HOPK || kind == OP_HOPB) hopping = true;
            if (kind == OP_ROTK && BP::ket_pos[BP::ops[o][1]] >= 5) return false;
            if (kind == OP_ROTB) {
                if (BP::ket_pos[BP::ops[o][1]] >= 5 || (cols & BP::ops[o][3])) return false;
                cols |= BP::ops[o][3];
            }
        }
        return any_damped;
    }
    static constexpr size_t smem_bytes(int p_doubles) {
        return (size_t)BUF * sizeof(double2) + (size_t)BP::NOPS * sizeof(double2) + (size_t)((p_doubles + 1) & ~1) * sizeof(double);
    }
};

// bra rotations: over a lane bit by shuffles, over a warp bit the columns are written for the round's barrier
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
            const double w = la ? (lb ? -1.0 : 1.0) : 0.0;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((mask >> j) & 1) {
                        double2 p = shfl_xor2(v[i][j], 1 << pb);
                        if constexpr (G::damped(a)) {
                            const double2 q = shfl_xor2(v[i][j], 1 << pk);
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
        const double w = la ? (lb ? -1.0 : 1.0) : 0.0;
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

// ket rotation, everything by shuffles (the partner bit is a lane bit, or the pseudomode is not damped)
template <int PID, int O>
__device__ __forceinline__ void k_apply_lanes(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P, int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB;
    constexpr int a = BP::ops[O][1], mask = BP::ops[O][3], slot = BP::ops[O][4];
    constexpr int pk = BP::ket_pos[a], pb = BP::bra_pos[a];
    const double t = P[slot + 1];
    const int la = (tid >> pk) & 1, lb = (tid >> pb) & 1;
    const double w = lb ? (la ? -1.0 : 1.0) : 0.0;
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j)
            if ((mask >> i) & 1) {
                double2 p = shfl_xor2(v[i][j], 1 << pk);
                if constexpr (G::damped(a)) {
                    const double2 q = shfl_xor2(v[i][j], 1 << pb);
                    p.x = fma(w, q.x, p.x);
                    p.y = fma(w, q.y, p.y);
                }
                rrc::tan_one(v[i][j], p, t);
            }
}

// ket rotation whose partner bit is a warp bit: the threads with the bit CLEAR write their rows ...
template <int PID, int O>
__device__ __forceinline__ void k_write(const double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], double2* kbuf, int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
    constexpr int a = BP::ops[O][1], mask = BP::ops[O][3], pb = BP::bra_pos[a];
    if (!((tid >> pb) & 1)) {                                    // (a warp bit: whole warps write, whole warps read)
        double2* buf = kbuf + G::k_offset(O) + rrc::remove_bit(tid, pb);
        int e = 0;
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if ((mask >> i) & 1) buf[(e++) * (NT / 2)] = v[i][j];
    }
}

// ... and after the barrier the threads with the bit SET read them next to the shuffle across their own bit
template <int PID, int O>
__device__ __forceinline__ void k_read_apply(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P,
                                             const double2* kbuf, int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
    constexpr int a = BP::ops[O][1], mask = BP::ops[O][3], slot = BP::ops[O][4];
    constexpr int pk = BP::ket_pos[a], pb = BP::bra_pos[a];
    const double t = P[slot + 1];
    const int la = (tid >> pk) & 1, lb = (tid >> pb) & 1;
    const double w = la ? -1.0 : 1.0;
    const double2* buf = kbuf + G::k_offset(O) + rrc::remove_bit(tid, pb);
    int e = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j)
            if ((mask >> i) & 1) {
                double2 p = shfl_xor2(v[i][j], 1 << pk);
                if (lb) {
                    const double2 q = buf[e * (NT / 2)];
                    p.x = fma(w, q.x, p.x);
                    p.y = fma(w, q.y, p.y);
                }
                ++e;
                rrc::tan_one(v[i][j], p, t);
            }
}

// ket rotations, rows pairwise disjoint: everything reachable by shuffles now, the others write for ONE shared round
template <int PID, int O>
__device__ __forceinline__ void op_k_first(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P, double2* kbuf,
                                           int tid) {
    if constexpr (BigProg<PID>::ops[O][0] == OP_ROTK) {
        if constexpr (Geo<PID>::k_deferred(O)) k_write<PID, O>(v, kbuf, tid);
        else k_apply_lanes<PID, O>(v, P, tid);
    }
}
template <int PID, int O>
__device__ __forceinline__ void op_k_deferred(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P,
                                              const double2* kbuf, int tid) {
    if constexpr (Geo<PID>::k_deferred(O)) k_read_apply<PID, O>(v, P, kbuf, tid);
}
// ket rotations, overlapping rows: one after the other in program order, a round of its own for every deferred one
template <int PID, int O>
__device__ __forceinline__ void op_k_sequential(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P,
                                                double2* kbuf, int tid) {
    if constexpr (BigProg<PID>::ops[O][0] == OP_ROTK) {
        if constexpr (Geo<PID>::k_deferred(O)) {
            k_write<PID, O>(v, kbuf, tid);
            __syncthreads();
            k_read_apply<PID, O>(v, P, kbuf, tid);
        } else {
            k_apply_lanes<PID, O>(v, P, tid);
        }
    }
}

// hops and the diagonal multiply that ends the step (with the next step's dampings folded in), in program order
template <int PID, int O>
__device__ __forceinline__ void op_tail(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double2 g, const double* P,
                                        const double2* H) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB;
    constexpr int kind = BP::ops[O][0], a = BP::ops[O][1], b = BP::ops[O][2];
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
    }
}

template <int PID, int... Os>
__device__ __forceinline__ void step3t(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double2 g, const double* P,
                                       const double2* H, double2* xbuf, int tid, qstd::integer_sequence<int, Os...>) {
    using G = Geo<PID>;
    double2* kbuf = xbuf;
    double2* bbuf = xbuf + G::K_ELEMS;
    (op_b_first<PID, Os>(v, P, bbuf, tid), ...);
    if constexpr (G::B_ELEMS > 0) {
        __syncthreads();
        (op_b_deferred<PID, Os>(v, P, bbuf, tid), ...);
    }
    if constexpr (G::ket_rows_disjoint()) {
        (op_k_first<PID, Os>(v, P, kbuf, tid), ...);
        if constexpr (G::K_ELEMS > 0) {
            __syncthreads();
            (op_k_deferred<PID, Os>(v, P, kbuf, tid), ...);
        }
    } else {
        (op_k_sequential<PID, Os>(v, P, kbuf, tid), ...);
    }
    (op_tail<PID, Os>(v, g, P, H), ...);
    // every region is rewritten in the next step: its readers must have passed a barrier in between, which any other
    // round of the step provides -- a step with a single round needs one more
    if constexpr (G::ROUNDS == 1) __syncthreads();
}

// z(1,1) += sign W-or-z(0,0) for every damped pseudomode: the (0,0) threads of the bit pair hand their tile to the
// (1,1) threads through the exchange region (CTA barriers on both sides; load and emissions only)
template <int PID, int O>
__device__ __forceinline__ void pair_shift(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], double sign, double2* buf, int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    if constexpr (BP::ops[O][0] == OP_ADT) {
        constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
        constexpr int a = BP::ops[O][1];
        constexpr int pa = BP::ket_pos[a], pb = BP::bra_pos[a];
        constexpr int hi = pa > pb ? pa : pb, lo = pa > pb ? pb : pa;
        const int la = (tid >> pa) & 1, lb = (tid >> pb) & 1;
        const int cid = rrc::remove_bit(rrc::remove_bit(tid, hi), lo);
        __syncthreads();
        if (!(la | lb)) {
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) buf[(i * NB + j) * (NT / 4) + cid] = v[i][j];
        }
        __syncthreads();
        if (la & lb) {
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

// per-thread product of the damping eigenvalues: 1, cos, cos, cos^2 for (a_q, b_q) = (0,0), (0,1), (1,0), (1,1)
template <int PID, int O>
__device__ __forceinline__ double damp_scale(const double* P, int tid) {
    using BP = BigProg<PID>;
    constexpr int kind = BP::ops[O][0], a = BP::ops[O][1], slot = BP::ops[O][4];
    if constexpr (kind == OP_ADT) {
        constexpr int pa = BP::ket_pos[a], pb = BP::bra_pos[a];
        const double c = P[slot];
        const int n = ((tid >> pa) & 1) + ((tid >> pb) & 1);
        return n == 0 ? 1.0 : (n == 1 ? c : c * c);
    } else {
        return 1.0;
    }
}
template <int PID, int... Os>
__device__ __forceinline__ double damp_scales(const double* P, int tid, qstd::integer_sequence<int, Os...>) {
    return (damp_scale<PID, Os>(P, tid) * ...);
}

template <int PID>
__global__ void __launch_bounds__(1 << (2 * BigProg<PID>::NBITS), 1) evolve_rrc3t_kernel(const Params k) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB, NBITS = G::NBITS, M = G::M, NT = G::NT, DB = G::DB, E = G::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const qsx_rrc_args& a = k.a;
    double2* xbuf = reinterpret_cast<double2*>(smem_raw);
    double2* H = xbuf + G::BUF;
    double* P = reinterpret_cast<double*>(H + BP::NOPS);
    double2* shift_buf = xbuf + G::K_ELEMS;                      // the bra-round region doubles as the transform's buffer
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
        const double dd = damp_scales<PID>(P, tid, all_ops);
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
        // W -> v = dd . z,  z = T^-T W
        pair_shifts<PID>(v, -1.0, shift_buf, tid, all_ops);
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                v[i][j].x *= dd;
                v[i][j].y *= dd;
            }
        __syncthreads();
        int slot = 0;
        int next_emit = a.n_emit > 0 ? a.emit_steps[0] : -1;
        int after_next = a.n_emit > 1 ? a.emit_steps[1] : -1;
        for (int step = 0; step <= a.n_steps && slot < a.n_emit; ++step) {
            if (step) step3t<PID>(v, g1, P, H, xbuf, tid, all_ops);
            if (step == next_emit) {
                // back to W in place: the pre-applied dampings of the next step come off, then T^T
#pragma unroll
                for (int i = 0; i < NA; ++i)
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        v[i][j].x *= inv_dd;
                        v[i][j].y *= inv_dd;
                    }
                pair_shifts<PID>(v, +1.0, shift_buf, tid, all_ops);
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
                if (slot < a.n_emit) {                           // onwards: z again, dampings pre-applied
                    pair_shifts<PID>(v, -1.0, shift_buf, tid, all_ops);
#pragma unroll
                    for (int i = 0; i < NA; ++i)
#pragma unroll
                        for (int j = 0; j < NB; ++j) {
                            v[i][j].x *= dd;
                            v[i][j].y *= dd;
                        }
                    __syncthreads();
                }
            }
        }
    }
}

#ifndef __CUDACC_RTC__
// 0: launched, 1: no eligible compile-time program (or a read-out this form does not compute), < 0: CUDA error (negated)
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
            if (a.emit_trace) return 1;                          // snapshots only (what a read-out propagation asks for)
            const size_t smem = G::smem_bytes(a.param_stride);
            if (smem > (size_t)smem_max) return 1;
            cudaError_t e = cudaFuncSetAttribute(evolve_rrc3t_kernel<PID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return -(int)e;
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, evolve_rrc3t_kernel<PID>, G::NT, smem);
            if (e != cudaSuccess) return -(int)e;
            if (per_sm < 1) per_sm = 1;
            int grid = sms * per_sm;
            if (grid > a.n_inst) grid = a.n_inst;
            Params k;
            k.a = a;
            evolve_rrc3t_kernel<PID><<<grid, G::NT, smem, stream>>>(k);
            e = cudaGetLastError();
            return e == cudaSuccess ? 0 : -(int)e;
        }
    }
}
#endif  // !__CUDACC_RTC__

}  // namespace rrc3t
