// qsx_rr.cuh -- register-resident evolution kernel for small power-of-two blocks (see include/qsx.h, qsx_rr_args).
//
// A team of 2^L lanes owns one instance; each lane keeps 2^R elements of sigma AND the diagonal table(s) of the
// step in registers for the whole evolution.  Every operation is "combine each element with its xor-partner": a
// register renaming when the partner lives in the same thread, a warp shuffle when it lives on another lane of the
// team.  Shared memory only holds the per-instance angles (a few doubles) and the launch-uniform read-out weights;
// HBM is touched when an instance is loaded and when it emits.
#pragma once
#include <utility>

namespace rr {

template <int PID>
struct StaticProg;      // compile-time copies of host-built programs, specialised in qsx_rr_programs.inc
#include "qsx_rr_programs.inc"

struct Params {
    qsx_rr_args a;
    int32_t E, team, inst_per_cta, p_ops;     // p_ops: doubles of op parameters per set (after the diagonal tables)
};

// v <- c v - i s p  (s already carries the ket/bra sign).  TAN: v <- v - i t p (cos folded into the diagonal table).
template <bool TAN>
__device__ __forceinline__ void rot_one(double2& v, const double2 p, double c, double s) {
    if (TAN) {
        v.x = fma(s, p.y, v.x);
        v.y = fma(-s, p.x, v.y);
    } else {
        const double nx = fma(c, v.x, s * p.y), ny = fma(c, v.y, -s * p.x);
        v.x = nx;
        v.y = ny;
    }
}

// Hops and conditioned X-rotations.  The partner differs in the register bits op.reg_xor or in the lane bits
// op.lane_xor; op.reg_mask lists the registers that take part (set for both members of an in-thread pair).
template <int R, bool TAN>
__device__ __forceinline__ void op_rot(double2 (&v)[1 << R], const qsx_rr_op& op, double c, double s, unsigned wmask) {
    constexpr int NR = 1 << R;
    const unsigned am = op.reg_mask;
    const int mr = op.reg_xor;
    if (op.lane_xor) {                                  // partner on another lane of the team
#pragma unroll
        for (int r = 0; r < NR; ++r)
            if ((am >> r) & 1) {
                double2 p;
                p.x = __shfl_xor_sync(wmask, v[r].x, op.lane_xor);
                p.y = __shfl_xor_sync(wmask, v[r].y, op.lane_xor);
                rot_one<TAN>(v[r], p, c, s);
            }
    } else if ((mr & (mr - 1)) == 0) {                  // single register bit: in-thread pairs, static indices
#pragma unroll
        for (int B = 0; B < R; ++B)
            if (mr == (1 << B)) {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    if (!((r >> B) & 1) && ((am >> r) & 1)) {
                        const double2 x = v[r], y = v[r | (1 << B)];
                        rot_one<TAN>(v[r], y, c, s);
                        rot_one<TAN>(v[r | (1 << B)], x, c, s);
                    }
            }
    } else if (R <= 3) {                                // several register bits (configuration index of N = 4)
        double2 p[NR];
#pragma unroll
        for (int MR = 3; MR < NR; ++MR)
            if (mr == MR) {
#pragma unroll
                for (int r = 0; r < NR; ++r) p[r] = v[r ^ MR];
            }
#pragma unroll
        for (int r = 0; r < NR; ++r)
            if ((am >> r) & 1) rot_one<TAN>(v[r], p[r], c, s);
    }
}

// Amplitude damping of one pseudomode: its ket bit and bra bit sit in registers (reg_xor) and/or on lanes (aux0, aux1).
template <int R>
__device__ __forceinline__ void op_ad(double2 (&v)[1 << R], const qsx_rr_op& op, double c, double s2, int lane_in_team,
                                      unsigned wmask) {
    constexpr int NR = 1 << R;
    const int mr = op.reg_xor;
    const double cc = c * c;
    if (op.lane_xor == 0) {                             // both bits in registers: static quads
#pragma unroll
        for (int BA = 0; BA < R; ++BA)
#pragma unroll
            for (int BB = BA + 1; BB < R; ++BB)
                if (mr == ((1 << BA) | (1 << BB))) {
#pragma unroll
                    for (int r = 0; r < NR; ++r)
                        if (!(r & ((1 << BA) | (1 << BB)))) {
                            double2& v00 = v[r];
                            double2& v01 = v[r | (1 << BA)];
                            double2& v10 = v[r | (1 << BB)];
                            double2& v11 = v[r | (1 << BA) | (1 << BB)];
                            v00.x = fma(s2, v11.x, v00.x);
                            v00.y = fma(s2, v11.y, v00.y);
                            v01.x *= c; v01.y *= c;
                            v10.x *= c; v10.y *= c;
                            v11.x *= cc; v11.y *= cc;
                        }
                }
        return;
    }
    const bool la = (lane_in_team & op.aux0) != 0, lb = (lane_in_team & op.aux1) != 0;
    const double a_lane = (la ? c : 1.0) * (lb ? c : 1.0);
    const double b_lane = (la || lb) ? 0.0 : s2;
    if (mr == 0) {                                      // both bits on lanes
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double2 p;
            p.x = __shfl_xor_sync(wmask, v[r].x, op.lane_xor);
            p.y = __shfl_xor_sync(wmask, v[r].y, op.lane_xor);
            v[r].x = fma(b_lane, p.x, a_lane * v[r].x);
            v[r].y = fma(b_lane, p.y, a_lane * v[r].y);
        }
    } else {                                            // one bit in a register, the other on a lane
        const double a_set = a_lane * c;
#pragma unroll
        for (int B = 0; B < R; ++B)
            if (mr == (1 << B)) {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    if (!((r >> B) & 1)) {
                        double2& lo = v[r];
                        double2& hi = v[r | (1 << B)];
                        double2 p;
                        p.x = __shfl_xor_sync(wmask, hi.x, op.lane_xor);
                        p.y = __shfl_xor_sync(wmask, hi.y, op.lane_xor);
                        lo.x = fma(b_lane, p.x, a_lane * lo.x);
                        lo.y = fma(b_lane, p.y, a_lane * lo.y);
                        hi.x *= a_set;
                        hi.y *= a_set;
                    }
            }
    }
}

// Generic read-outs of one emission (any observable): a plain loop over the observables -- small code, used by the
// generic kernel and by compile-time programs whose read-out is not one of the two known kinds.
template <int R>
__device__ __forceinline__ void emit_generic(const double2 (&v)[1 << R], const qsx_rr_args& a, const double2* obs_w, int E,
                                          int phys0, int lt, int team, bool valid, int inst, int slot) {
    constexpr int NR = 1 << R;
    for (int o = 0; o < a.n_obs; ++o) {
        double re = 0.0, im = 0.0;
        if (a.n_entries[0] > 0) {
            for (int e = 0; e < a.n_entries[o]; ++e) {
                const unsigned rm = a.entry_reg_mask[o][e];
                double sr[4] = {0.0, 0.0, 0.0, 0.0}, si[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    if ((rm >> r) & 1) {
                        sr[r & 3] += v[r].x;
                        si[r & 3] += v[r].y;
                    }
                const double wgt = ((a.entry_lane_mask[o][e] >> lt) & 1) ? a.entry_weight[o][e] : 0.0;
                re = fma(wgt, (sr[0] + sr[1]) + (sr[2] + sr[3]), re);
                im = fma(wgt, (si[0] + si[1]) + (si[2] + si[3]), im);
            }
        } else {
            const unsigned rm = a.obs_reg_mask[o];
#pragma unroll
            for (int r = 0; r < NR; ++r)
                if ((rm >> r) & 1) {
                    const double2 wt = obs_w[o * E + phys0 + r];
                    re = fma(wt.x, v[r].x, fma(-wt.y, v[r].y, re));
                    im = fma(wt.x, v[r].y, fma(wt.y, v[r].x, im));
                }
        }
        for (int off = team >> 1; off; off >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, off);
            im += __shfl_xor_sync(0xffffffffu, im, off);
        }
        if (lt == 0 && valid) {
            const size_t at = ((size_t)inst * a.n_emit + slot) * a.n_obs + o;
            if (a.obs_real) a.out_obs[at] = re;
            else reinterpret_cast<double2*>(a.out_obs)[at] = make_double2(re, im);
        }
    }
}

template <int R, int ND>
__global__ void __launch_bounds__(32) evolve_rr_kernel(const Params k) {
    constexpr int NR = 1 << R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const qsx_rr_args& a = k.a;
    // shared: [read-out weights n_obs * E complex] [per team: op parameters]
    double2* obs_w = reinterpret_cast<double2*>(smem_raw);
    double* p_all = reinterpret_cast<double*>(obs_w + (size_t)a.n_obs * k.E);
    for (int i = threadIdx.x; i < a.n_obs * k.E; i += blockDim.x)
        obs_w[i] = reinterpret_cast<const double2*>(a.obs_weights)[i];
    __syncthreads();

    const int team_id = threadIdx.x >> a.lane_bits;
    const int lt = threadIdx.x & (k.team - 1);                 // lane in team
    double* P = p_all + (size_t)team_id * ((k.p_ops + 1) & ~1);
    const int phys0 = lt << R;
    constexpr unsigned wmask = 0xffffffffu;

    int logical[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) logical[r] = a.logical_index[phys0 + r];

    // All lanes of the warp stay in the loop until every team of the warp is done (teams past the end redo the last
    // instance without writing), so that every shuffle can use the constant full mask (plain SHFL, no WARPSYNC).
    for (int inst_raw = blockIdx.x * k.inst_per_cta + team_id;; inst_raw += gridDim.x * k.inst_per_cta) {
        const bool valid = inst_raw < a.n_inst;
        if (!__any_sync(0xffffffffu, valid)) break;
        const int inst = valid ? inst_raw : a.n_inst - 1;
        const int set = a.inst_set ? a.inst_set[inst] : 0;
        const int src = a.inst_src ? a.inst_src[inst] : inst;
        const double* Pg = a.params + (size_t)set * a.param_stride;
        double2 w[ND > 0 ? ND : 1][NR];
#pragma unroll
        for (int d = 0; d < ND; ++d)
#pragma unroll
            for (int r = 0; r < NR; ++r) w[d][r] = reinterpret_cast<const double2*>(Pg)[(size_t)d * k.E + phys0 + r];
        __syncwarp();
        for (int i = lt; i < k.p_ops; i += k.team) P[i] = Pg[(size_t)2 * ND * k.E + i];
        double2 v[NR];
        const double2* sin = reinterpret_cast<const double2*>(a.sigma_in) + (size_t)src * k.E;
#pragma unroll
        for (int r = 0; r < NR; ++r) v[r] = sin[logical[r]];
        __syncwarp();

        int slot = 0;
        int next_emit = a.n_emit > 0 ? a.emit_steps[0] : -1;
        int after_next = a.n_emit > 1 ? a.emit_steps[1] : -1;      // schedule is read one emission ahead
        for (int step = 0; step <= a.n_steps && slot < a.n_emit; ++step) {
            if (step) {
                for (int o = 0; o < a.n_ops; ++o) {
                    const qsx_rr_op& op = a.ops[o];
                    if (op.kind == QSX_RR_DIAG) {
#pragma unroll
                        for (int d = 0; d < ND; ++d)
                            if (op.aux0 == d) {
#pragma unroll
                                for (int r = 0; r < NR; ++r) v[r] = cmul(v[r], w[d][r]);
                            }
                    } else {
                        const double2 cs = *reinterpret_cast<const double2*>(P + op.p);
                        if (op.kind == QSX_RR_ROT) op_rot<R, false>(v, op, cs.x, op.sign > 0 ? cs.y : -cs.y, wmask);
                        else if (op.kind == QSX_RR_ROTT) op_rot<R, true>(v, op, cs.x, op.sign > 0 ? cs.y : -cs.y, wmask);
                        else op_ad<R>(v, op, cs.x, cs.y, lt, wmask);
                    }
                }
            }
            if (step == next_emit) {
                emit_generic<R>(v, a, obs_w, k.E, phys0, lt, k.team, valid, inst, slot);
                if (a.out_snap && valid) {
                    double2* dst = reinterpret_cast<double2*>(a.out_snap) + ((size_t)inst * a.n_emit + slot) * k.E;
#pragma unroll
                    for (int r = 0; r < NR; ++r) dst[logical[r]] = v[r];
                }
                ++slot;
                next_emit = after_next;
                after_next = slot + 1 < a.n_emit ? a.emit_steps[slot + 1] : -1;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// static variant: the operation list is a compile-time constant (StaticProg<PID>), so every mask test, dispatch and
// parameter slot folds away; the angles of the instance live in registers next to sigma and the diagonal table.
// ------------------------------------------------------------------------------------------------
template <int PID, int O>
__device__ __forceinline__ void static_op(double2 (&v)[1 << StaticProg<PID>::R],
                                          const double2 (&w)[StaticProg<PID>::ND > 0 ? StaticProg<PID>::ND : 1][1 << StaticProg<PID>::R],
                                          const double2 (&cs)[StaticProg<PID>::NOPS], int lt, unsigned wmask) {
    constexpr int R = StaticProg<PID>::R, NR = 1 << R;
    constexpr qsx_rr_op op = StaticProg<PID>::ops[O];
    if constexpr (op.kind == QSX_RR_DIAG) {
#pragma unroll
        for (int r = 0; r < NR; ++r) v[r] = cmul(v[r], w[op.aux0][r]);
    } else if constexpr (op.kind == QSX_RR_ROT) {
        op_rot<R, false>(v, op, cs[O].x, op.sign > 0 ? cs[O].y : -cs[O].y, wmask);
    } else if constexpr (op.kind == QSX_RR_ROTT) {
        op_rot<R, true>(v, op, cs[O].x, op.sign > 0 ? cs[O].y : -cs[O].y, wmask);
    } else {
        op_ad<R>(v, op, cs[O].x, cs[O].y, lt, wmask);
    }
}

template <int PID, int... Os>
__device__ __forceinline__ void static_step(double2 (&v)[1 << StaticProg<PID>::R],
                                            const double2 (&w)[StaticProg<PID>::ND > 0 ? StaticProg<PID>::ND : 1][1 << StaticProg<PID>::R],
                                            const double2 (&cs)[StaticProg<PID>::NOPS], int lt, unsigned wmask,
                                            std::integer_sequence<int, Os...>) {
    (static_op<PID, Os>(v, w, cs, lt, wmask), ...);
}

template <int PID, int... Os>
__device__ __forceinline__ void static_load_angles(double2 (&cs)[StaticProg<PID>::NOPS], const double* Pg,
                                                   std::integer_sequence<int, Os...>) {
    ((cs[Os] = StaticProg<PID>::ops[Os].kind == QSX_RR_DIAG
                   ? make_double2(0.0, 0.0)
                   : *reinterpret_cast<const double2*>(Pg + StaticProg<PID>::ops[Os].p)),
     ...);
}

// Read-outs with compile-time masks.  RO = 1: site populations (real); RO = 2: emission trace (complex, weights mu).
// (indices into the constexpr tables must be constant expressions in device code, hence the index sequences)
template <int PID, int O, int E>
__device__ __forceinline__ double pop_piece(const double2 (&v)[1 << StaticProg<PID>::R], int lt) {
    constexpr unsigned rm = StaticProg<PID>::pop_reg[O][E], lm = StaticProg<PID>::pop_lane[O][E];
    double sum = 0.0;
#pragma unroll
    for (int r = 0; r < (1 << StaticProg<PID>::R); ++r)
        if ((rm >> r) & 1) sum += v[r].x;
    return ((lm >> lt) & 1) ? sum : 0.0;
}
template <int PID, int O, int... Es>
__device__ __forceinline__ double pop_obs(const double2 (&v)[1 << StaticProg<PID>::R], int lt, std::integer_sequence<int, Es...>) {
    return (pop_piece<PID, O, Es>(v, lt) + ...);
}
template <int PID, int... Os>
__device__ __forceinline__ void pop_all(const double2 (&v)[1 << StaticProg<PID>::R], int lt, double (&p)[StaticProg<PID>::NPOP],
                                        std::integer_sequence<int, Os...>) {
    ((p[Os] = pop_obs<PID, Os>(v, lt, std::make_integer_sequence<int, StaticProg<PID>::POPE>{})), ...);
}

template <int PID, int S, int E>
__device__ __forceinline__ void em_piece(const double2 (&v)[1 << StaticProg<PID>::R], int lt, double mu, double& re, double& im) {
    constexpr unsigned rm = StaticProg<PID>::em_reg[S][E], lm = StaticProg<PID>::em_lane[S][E];
    if constexpr (rm != 0) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int r = 0; r < (1 << StaticProg<PID>::R); ++r)
            if ((rm >> r) & 1) {
                sr += v[r].x;
                si += v[r].y;
            }
        const double wgt = ((lm >> lt) & 1) ? mu : 0.0;
        re = fma(wgt, sr, re);
        im = fma(wgt, si, im);
    }
}
template <int PID, int S, int... Es>
__device__ __forceinline__ void em_site(const double2 (&v)[1 << StaticProg<PID>::R], int lt, double mu, double& re, double& im,
                                        std::integer_sequence<int, Es...>) {
    (em_piece<PID, S, Es>(v, lt, mu, re, im), ...);
}
template <int PID, int... Ss>
__device__ __forceinline__ void em_all(const double2 (&v)[1 << StaticProg<PID>::R], int lt, const qsx_rr_args& a, double& re,
                                       double& im, std::integer_sequence<int, Ss...>) {
    (em_site<PID, Ss>(v, lt, a.mu[Ss], re, im, std::make_integer_sequence<int, StaticProg<PID>::EME>{}), ...);
}

// Static read-outs are split in two so that the team reduction of emission t overlaps the arithmetic of step t+1:
// `partial` leaves one per-lane partial sum per read-out in `pend`; `flush` (branch-free: it runs at the start of every
// step, the compiler interleaves its shuffles with the step's independent FP64 work) reduces over the team and stores.
template <int PID, int RO>
struct Pending {
    static constexpr int N = RO == 1 ? StaticProg<PID>::NPOP : 2;
    double val[N > 0 ? N : 1];
    long long at;            // index of the first output value, or -1
};

template <int PID, int RO>
__device__ __forceinline__ void emit_partial(const double2 (&v)[1 << StaticProg<PID>::R], const qsx_rr_args& a, int lt,
                                             Pending<PID, RO>& pend) {
    using SP = StaticProg<PID>;
    if constexpr (RO == 1) {
        pop_all<PID>(v, lt, pend.val, std::make_integer_sequence<int, SP::NPOP>{});
    } else {
        pend.val[0] = 0.0;
        pend.val[1] = 0.0;
        em_all<PID>(v, lt, a, pend.val[0], pend.val[1], std::make_integer_sequence<int, SP::NEM>{});
    }
}

template <int PID, int RO>
__device__ __forceinline__ void emit_flush(const qsx_rr_args& a, int lt, Pending<PID, RO>& pend) {
    constexpr int TEAM = 1 << StaticProg<PID>::L, N = Pending<PID, RO>::N;
#pragma unroll
    for (int off = TEAM >> 1; off; off >>= 1)
#pragma unroll
        for (int o = 0; o < N; ++o) pend.val[o] += __shfl_xor_sync(0xffffffffu, pend.val[o], off);
    if (lt == 0 && pend.at >= 0) {
        double* dst = a.out_obs + pend.at;
        if constexpr (N == 2) {
            *reinterpret_cast<double2*>(dst) = make_double2(pend.val[0], pend.val[1]);    // 16-byte aligned: at is even
        } else {
#pragma unroll
            for (int o = 0; o < N; ++o) dst[o] = pend.val[o];
        }
    }
    pend.at = -1;
}

template <int PID, int RO>
__global__ void __launch_bounds__(32) evolve_rr_static_kernel(const Params k) {
    using SP = StaticProg<PID>;
    constexpr int R = SP::R, NR = 1 << R, ND = SP::ND, TEAM = 1 << SP::L, E = 1 << (SP::R + SP::L);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const qsx_rr_args& a = k.a;
    double2* obs_w = reinterpret_cast<double2*>(smem_raw);
    for (int i = threadIdx.x; i < a.n_obs * E; i += blockDim.x)
        obs_w[i] = reinterpret_cast<const double2*>(a.obs_weights)[i];
    __syncthreads();

    const int team_id = threadIdx.x / TEAM;
    const int lt = threadIdx.x % TEAM;
    const int phys0 = lt << R;
    constexpr unsigned wmask = 0xffffffffu;
    int logical[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) logical[r] = a.logical_index[phys0 + r];

    for (int inst_raw = blockIdx.x * (32 / TEAM) + team_id;; inst_raw += gridDim.x * (32 / TEAM)) {
        const bool valid = inst_raw < a.n_inst;
        if (!__any_sync(0xffffffffu, valid)) break;
        const int inst = valid ? inst_raw : a.n_inst - 1;
        const int set = a.inst_set ? a.inst_set[inst] : 0;
        const int src = a.inst_src ? a.inst_src[inst] : inst;
        const double* Pg = a.params + (size_t)set * a.param_stride;
        double2 w[ND > 0 ? ND : 1][NR];
#pragma unroll
        for (int d = 0; d < ND; ++d)
#pragma unroll
            for (int r = 0; r < NR; ++r) w[d][r] = reinterpret_cast<const double2*>(Pg)[d * E + phys0 + r];
        double2 cs[SP::NOPS];
        static_load_angles<PID>(cs, Pg, std::make_integer_sequence<int, SP::NOPS>{});
        double2 v[NR];
        const double2* sin = reinterpret_cast<const double2*>(a.sigma_in) + (size_t)src * E;
#pragma unroll
        for (int r = 0; r < NR; ++r) v[r] = sin[logical[r]];

        int slot = 0;
        int next_emit = a.n_emit > 0 ? a.emit_steps[0] : -1;
        int after_next = a.n_emit > 1 ? a.emit_steps[1] : -1;      // schedule is read one emission ahead
        Pending<PID, RO> pend;
        pend.at = -1;
#pragma unroll
        for (int o = 0; o < Pending<PID, RO>::N; ++o) pend.val[o] = 0.0;
        for (int step = 0; step <= a.n_steps && slot < a.n_emit; ++step) {
            if (step) {
                if constexpr (RO != 0) emit_flush<PID, RO>(a, lt, pend);          // read-out of the previous emission
                static_step<PID>(v, w, cs, lt, wmask, std::make_integer_sequence<int, SP::NOPS>{});
            }
            if (step == next_emit) {
                if constexpr (RO == 0) {
                    emit_generic<R>(v, a, obs_w, E, phys0, lt, TEAM, valid, inst, slot);
                } else {
                    emit_partial<PID, RO>(v, a, lt, pend);
                    pend.at = valid ? ((long long)inst * a.n_emit + slot) * Pending<PID, RO>::N : -1;
                }
                if (a.out_snap && valid) {
                    double2* dst = reinterpret_cast<double2*>(a.out_snap) + ((size_t)inst * a.n_emit + slot) * E;
#pragma unroll
                    for (int r = 0; r < NR; ++r) dst[logical[r]] = v[r];
                }
                ++slot;
                next_emit = after_next;
                after_next = slot + 1 < a.n_emit ? a.emit_steps[slot + 1] : -1;
            }
        }
        if constexpr (RO != 0) emit_flush<PID, RO>(a, lt, pend);
    }
}

// host: does `a` equal the compile-time program PID bit for bit?
template <int PID>
bool matches_static(const qsx_rr_args& a) {
    using SP = StaticProg<PID>;
    if (a.reg_bits != SP::R || a.lane_bits != SP::L || a.n_diag != SP::ND || a.n_ops != SP::NOPS) return false;
    for (int i = 0; i < SP::NOPS; ++i) {
        const qsx_rr_op &x = a.ops[i], &y = SP::ops[i];
        if (x.kind != y.kind || x.reg_xor != y.reg_xor || x.lane_xor != y.lane_xor || x.reg_mask != y.reg_mask ||
            x.p != y.p || x.sign != y.sign || x.aux0 != y.aux0 || x.aux1 != y.aux1)
            return false;
    }
    return true;
}

// host: do the site-resolved read-out pieces sent by the caller equal the compile-time tables of PID?
template <int PID, int RO>
bool matches_static_readout(const qsx_rr_args& a) {
    using SP = StaticProg<PID>;
    constexpr int NS = RO == 1 ? SP::NPOP : SP::NEM, NE = RO == 1 ? SP::POPE : SP::EME;
    for (int s = 0; s < 16; ++s)
        for (int e = 0; e < QSX_RR_MAX_ENTRIES; ++e) {
            unsigned rm = 0, lm = 0;
            if (s < NS && e < NE) {
                rm = RO == 1 ? SP::pop_reg[s < NS ? s : 0][e < NE ? e : 0] : SP::em_reg[s < NS ? s : 0][e < NE ? e : 0];
                lm = RO == 1 ? SP::pop_lane[s < NS ? s : 0][e < NE ? e : 0] : SP::em_lane[s < NS ? s : 0][e < NE ? e : 0];
            }
            if (rm == 0) lm = 0;
            if (a.site_reg_mask[s][e] != rm || a.site_lane_mask[s][e] != lm) return false;
        }
    return true;
}

template <int PID>
cudaError_t launch_static(const Params& k, int grid, size_t smem, cudaStream_t stream, bool* launched) {
    if constexpr (PID >= kNumStaticProgs) {
        *launched = false;
        return cudaSuccess;
    } else {
        if (matches_static<PID>(k.a)) {
            using SP = StaticProg<PID>;
            *launched = true;
#define QSX_RR_STATIC_LAUNCH(RO)                                                                                      \
    do {                                                                                                              \
        cudaError_t e = cudaFuncSetAttribute(evolve_rr_static_kernel<PID, RO>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem);                                                              \
        if (e != cudaSuccess) return e;                                                                               \
        evolve_rr_static_kernel<PID, RO><<<grid, 32, smem, stream>>>(k);                                              \
        return cudaGetLastError();                                                                                    \
    } while (0)
            if constexpr (SP::NPOP > 0) {
                if (k.a.readout_kind == 1 && k.a.n_obs == SP::NPOP && k.a.obs_real && matches_static_readout<PID, 1>(k.a))
                    QSX_RR_STATIC_LAUNCH(1);
            }
            if constexpr (SP::NEM > 0) {
                if (k.a.readout_kind == 2 && k.a.n_obs == 1 && !k.a.obs_real && matches_static_readout<PID, 2>(k.a))
                    QSX_RR_STATIC_LAUNCH(2);
            }
            QSX_RR_STATIC_LAUNCH(0);
#undef QSX_RR_STATIC_LAUNCH
        }
        return launch_static<PID + 1>(k, grid, smem, stream, launched);
    }
}

}  // namespace rr
