// qsx_rrc2.cuh -- throughput form of the CTA-wide register-resident kernel (qsx_rrc.cuh) for levels with MANY instances.
//
// Same compile-time programs (BigProg<PID>), same thread <-> (m_a, m_b) mapping and register tile v[i][j]; what changes
// is everything that kept the first form at one CTA per SM with its two data pipes (FP64, shuffle / shared memory)
// taking turns:
//   * <= 128 registers per thread, two small exchange buffers -> TWO CTAs (instances) per SM: while one instance waits
//     for an exchange the other one issues FP64 work;
//   * the per-element diagonal factors are not kept in registers: w[i][j] = Wcfg[i][j] (shared memory, broadcast) times
//     one per-thread factor g(m_a, m_b);
//   * amplitude damping AD_q = D_q . T_q with T_q: (0,0) += sin^2(phi) (1,1) and the diagonal scaling
//     D_q = diag(1, cos, cos, cos^2) on (a_q, b_q).  Every D_q commutes with the T of the other pseudomodes and the
//     dampings end the step, so all of them fold into g of the NEXT step's diagonal multiply (the state kept in
//     registers lacks the trailing scalings; emissions and snapshots multiply by the per-thread scalar on the way out);
//   * exciton hops as three shears (tan(theta/2), sin(theta), tan(theta/2)): 6 FMA per pair instead of 4 MUL + 4 FMA;
//   * the bra rotations over warp bits share ONE exchange round (they touch different columns and commute with the
//     ket side), dampings inside a warp need no CTA barrier: four CTA barriers per step instead of seven.
// Forward programs only (the transposed read-out propagations are single instances: latency, not throughput).
#pragma once

namespace rrc2 {

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
    // Forward program of the expected shape: DIAG first, then hops and pseudomode rotations, dampings last; nothing
    // transposed; a tile of at most 16 elements; ket rotations exchange over lane bits (warp shuffles).
    static constexpr bool eligible() {
        if (NA * NB > 16 || BP::ops[0][0] != OP_DIAG) return false;
        bool in_tail = false;
        for (int o = 1; o < BP::NOPS; ++o) {
            const int kind = BP::ops[o][0];
            if (kind == OP_ADT || kind == OP_DIAG) return false;
            if (kind == OP_ROTK && BP::ket_pos[BP::ops[o][1]] >= 5) return false;
            if (kind == OP_ROTB && BP::bra_pos[BP::ops[o][1]] >= 5)      // one exchange round: disjoint columns only
                for (int q = 1; q < o; ++q)
                    if (BP::ops[q][0] == OP_ROTB && BP::bra_pos[BP::ops[q][1]] >= 5 && (BP::ops[q][3] & BP::ops[o][3])) return false;
            if (kind == OP_AD) in_tail = true;
            else if (in_tail) return false;
        }
        return true;
    }
    // Shared-memory exchanges of one step and their buffers (complex elements):
    //   R  all bra rotations over warp bits, written together and separated from their reads by ONE CTA barrier
    //      (they act on different columns of the tile and commute with every ket-side operation);
    //   A  dampings whose two thread bits are lane bits: the exchange stays inside a warp (__syncwarp);
    //   B, C  the other dampings, alternating, one CTA barrier each.
    // A buffer is rewritten only after a CTA barrier that every reader of its previous contents has passed.
    static __host__ __device__ constexpr bool rotb_warp(int o) { return BP::ops[o][0] == OP_ROTB && BP::bra_pos[BP::ops[o][1]] >= 5; }
    static __host__ __device__ constexpr int rotb_offset(int upto) {
        int off = 0;
        for (int o = 0; o < upto; ++o)
            if (rotb_warp(o)) off += popc(BP::ops[o][3]) * NA * NT;
        return off;
    }
    static __host__ __device__ constexpr int ad_hi(int o) {
        const int pa = BP::ket_pos[BP::ops[o][1]], pb = BP::bra_pos[BP::ops[o][1]];
        return pa > pb ? pa : pb;
    }
    static __host__ __device__ constexpr int ad_rank(int upto) {            // dampings over a warp bit before operation `upto`
        int n = 0;
        for (int o = 0; o < upto; ++o)
            if (BP::ops[o][0] == OP_AD && ad_hi(o) >= 5) ++n;
        return n;
    }
    static constexpr int R_ELEMS = rotb_offset(BP::NOPS);
    static constexpr int AD_ELEMS = NA * NB * (NT / 4);
    static constexpr int BUF = R_ELEMS + 3 * AD_ELEMS;                     // R | A | B | C
    static constexpr int MIN_CTAS = NT >= 256 ? 2 : (512 / NT > 8 ? 8 : 512 / NT);
    static constexpr size_t smem_bytes(int p_doubles) {
        return (size_t)BUF * sizeof(double2) + (size_t)((p_doubles + 1) & ~1) * sizeof(double) +
               64 * sizeof(double2) + (size_t)BP::NOPS * sizeof(double2);
    }
};

// shear form of the pair rotation x' = c x - i s y, y' = c y - i s x:  x -= i th y;  y -= i s x;  x -= i th y
__device__ __forceinline__ void pair_shear(double2& x, double2& y, double th, double s) {
    x.x = fma(th, y.y, x.x);
    x.y = fma(-th, y.x, x.y);
    y.x = fma(s, x.y, y.x);
    y.y = fma(-s, x.x, y.y);
    x.x = fma(th, y.y, x.x);
    x.y = fma(-th, y.x, x.y);
}

// Phase 1 of a step, in program order: diagonal multiply, hops of both sides, ket rotations and the bra rotations over
// lane bits (shuffles); the bra rotations over warp bits are deferred to phase 2 (they commute with all of these).
template <int PID, int O>
__device__ __forceinline__ void op_first(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double2 g, const double* P,
                                         const double2* H, double2* xbuf, int tid) {
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
        for (int j = 0; j < NB; ++j) pair_shear(v[a][j], v[b][j], ts.x, ts.y);
    } else if constexpr (kind == OP_HOPB) {
        const double2 ts = H[O];
#pragma unroll
        for (int i = 0; i < NA; ++i) pair_shear(v[i][a], v[i][b], -ts.x, -ts.y);
    } else if constexpr (kind == OP_ROTK || kind == OP_ROTB) {
        constexpr bool ket = kind == OP_ROTK;
        constexpr int pos = ket ? BP::ket_pos[a] : BP::bra_pos[a];
        if constexpr (pos < 5) {
            const double t = ket ? P[slot + 1] : -P[slot + 1];
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((mask >> (ket ? i : j)) & 1) {
                        double2 p;
                        p.x = __shfl_xor_sync(0xffffffffu, v[i][j].x, 1 << pos);
                        p.y = __shfl_xor_sync(0xffffffffu, v[i][j].y, 1 << pos);
                        rrc::tan_one(v[i][j], p, t);
                    }
        }
    }
}

// Phase 2a (after EVERY ket-side operation and lane-bit rotation of the step: what is written must be the state the
// partner's rotation acts on): the bra rotations over warp bits write their columns into R.
template <int PID, int O>
__device__ __forceinline__ void op_rotb_write(const double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], double2* xbuf, int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    if constexpr (G::rotb_warp(O)) {
        constexpr int NA = G::NA, NB = G::NB, NT = G::NT, mask = BP::ops[O][3];
        double2* buf = xbuf + G::rotb_offset(O);
        int e = 0;
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if ((mask >> j) & 1) buf[(e++) * NT + tid] = v[i][j];
    }
}

// Phase 2b (after the CTA barrier): the bra rotations over warp bits read their partners' columns from R.
template <int PID, int O>
__device__ __forceinline__ void op_rotb_read(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P,
                                             const double2* xbuf, int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    if constexpr (G::rotb_warp(O)) {
        constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
        constexpr int a = BP::ops[O][1], mask = BP::ops[O][3], slot = BP::ops[O][4], pos = BP::bra_pos[a];
        const double t = -P[slot + 1];
        const double2* buf = xbuf + G::rotb_offset(O);
        int e = 0;
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if ((mask >> j) & 1) rrc::tan_one(v[i][j], buf[(e++) * NT + (tid ^ (1 << pos))], t);
    }
}

// Phase 3: dampings without their scalings -- the (1,1) threads of the bit pair hand their tile to the (0,0) threads,
// which add sin^2(phi) times it.
template <int PID, int O>
__device__ __forceinline__ void op_damp(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double* P, double2* xbuf,
                                        int tid) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    if constexpr (BP::ops[O][0] == OP_AD) {
        constexpr int NA = G::NA, NB = G::NB, NT = G::NT;
        constexpr int a = BP::ops[O][1], slot = BP::ops[O][4];
        constexpr int pa = BP::ket_pos[a], pb = BP::bra_pos[a];
        constexpr int hi = pa > pb ? pa : pb, lo = pa > pb ? pb : pa;
        constexpr int which = hi < 5 ? 0 : 1 + (G::ad_rank(O) & 1);          // A | B, C alternating
        const double ss = P[slot + 1];                                        // sin^2 phi
        const int la = (tid >> pa) & 1, lb = (tid >> pb) & 1;
        const int cid = rrc::remove_bit(rrc::remove_bit(tid, hi), lo);
        double2* buf = xbuf + G::R_ELEMS + which * G::AD_ELEMS;
        if (la & lb) {
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) buf[(i * NB + j) * (NT / 4) + cid] = v[i][j];
        }
        if constexpr (hi < 5) __syncwarp(); else __syncthreads();
        if (!(la | lb)) {
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const double2 p = buf[(i * NB + j) * (NT / 4) + cid];
                    v[i][j].x = fma(ss, p.x, v[i][j].x);
                    v[i][j].y = fma(ss, p.y, v[i][j].y);
                }
        }
        if constexpr (hi < 5) __syncwarp();              // the warp's region is rewritten by this warp only
    }
}

template <int PID, int... Os>
__device__ __forceinline__ void step2(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double2 g, const double* P,
                                      const double2* H, double2* xbuf, int tid, qstd::integer_sequence<int, Os...>) {
    (op_first<PID, Os>(v, g, P, H, xbuf, tid), ...);
    if constexpr (Geo<PID>::R_ELEMS > 0) {
        (op_rotb_write<PID, Os>(v, xbuf, tid), ...);
        __syncthreads();
        (op_rotb_read<PID, Os>(v, P, xbuf, tid), ...);
    }
    (op_damp<PID, Os>(v, P, xbuf, tid), ...);
}

// per-thread product of the folded damping scalings: 1, cos, cos, cos^2 for (a_q, b_q) = (0,0), (0,1), (1,0), (1,1)
template <int PID, int O>
__device__ __forceinline__ double damp_scale(const double* P, int tid) {
    using BP = BigProg<PID>;
    constexpr int kind = BP::ops[O][0], a = BP::ops[O][1], slot = BP::ops[O][4];
    if constexpr (kind == OP_AD) {
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

// hop parameters (cos, sin) -> (tan(theta / 2), sin), once per instance
template <int PID, int O>
__device__ __forceinline__ void hop_entry(const double* P, double2* H) {
    using BP = BigProg<PID>;
    constexpr int kind = BP::ops[O][0], slot = BP::ops[O][4];
    if constexpr (kind == OP_HOPK || kind == OP_HOPB) {
        const double c = P[slot], s = P[slot + 1];
        H[O] = make_double2(s / (1.0 + c), s);
    }
}
template <int PID, int... Os>
__device__ __forceinline__ void hop_table(const double* P, double2* H, qstd::integer_sequence<int, Os...>) {
    (hop_entry<PID, Os>(P, H), ...);
}

template <int PID>
__global__ void __launch_bounds__(1 << (2 * BigProg<PID>::NBITS), Geo<PID>::MIN_CTAS) evolve_rrc2_kernel(const Params k) {
    using BP = BigProg<PID>;
    using G = Geo<PID>;
    constexpr int NA = G::NA, NB = G::NB, NBITS = G::NBITS, M = G::M, NT = G::NT, DB = G::DB, E = G::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const qsx_rrc_args& a = k.a;
    double2* xbuf = reinterpret_cast<double2*>(smem_raw);
    double2* red = xbuf + G::BUF;
    double2* H = red + 64;
    double* P = reinterpret_cast<double*>(H + BP::NOPS);
    const int tid = threadIdx.x;
    int ma, mb;
    rrc::thread_modes<PID>(tid, ma, mb, qstd::make_integer_sequence<int, NBITS>{});

    for (int inst = blockIdx.x; inst < a.n_inst; inst += gridDim.x) {
        const int set = a.inst_set ? a.inst_set[inst] : 0;
        const int src = a.inst_src ? a.inst_src[inst] : inst;
        const double* Pg = a.params + (size_t)set * a.param_stride;
        __syncthreads();
        for (int i = tid; i < a.param_stride; i += NT) P[i] = Pg[i];
        __syncthreads();
        if (tid == 0) hop_table<PID>(P, H, qstd::make_integer_sequence<int, BP::NOPS>{});
        const double2 g0 = cmul_conj(reinterpret_cast<const double2*>(P + 2 * NA * NB)[ma],
                                     reinterpret_cast<const double2*>(P + 2 * NA * NB + 2 * M)[mb]);
        const double dd = damp_scales<PID>(P, tid, qstd::make_integer_sequence<int, BP::NOPS>{});
        const double2 g1 = make_double2(g0.x * dd, g0.y * dd);
        double2 v[NA][NB];
        {
            const double2* sin = reinterpret_cast<const double2*>(a.sigma_in) + (size_t)src * E;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) v[i][j] = sin[(size_t)((i << NBITS) | ma) * DB + ((j << NBITS) | mb)];
        }
        __syncthreads();
        int slot = 0;
        int next_emit = a.n_emit > 0 ? a.emit_steps[0] : -1;
        int after_next = a.n_emit > 1 ? a.emit_steps[1] : -1;
        for (int step = 0; step <= a.n_steps && slot < a.n_emit; ++step) {
            if (step) step2<PID>(v, step == 1 ? g0 : g1, P, H, xbuf, tid, qstd::make_integer_sequence<int, BP::NOPS>{});
            if (step == next_emit) {
                const double out_scale = step ? dd : 1.0;          // the trailing damping scalings the registers lack
                if (a.emit_trace) {
                    double re[NBITS], im[NBITS];
#pragma unroll
                    for (int s = 0; s < NBITS; ++s) re[s] = im[s] = 0.0;
                    if constexpr (BP::NEM > 0) {
                        if (ma == mb) rrc::big_trace<PID>(v, re, im, qstd::make_integer_sequence<int, BP::NEM>{});
                    }
#pragma unroll
                    for (int s = 0; s < NBITS; ++s) {
                        re[s] *= out_scale;
                        im[s] *= out_scale;
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
                        for (int j = 0; j < NB; ++j)
                            dst[(size_t)((i << NBITS) | ma) * DB + ((j << NBITS) | mb)] =
                                make_double2(v[i][j].x * out_scale, v[i][j].y * out_scale);
                }
                ++slot;
                next_emit = after_next;
                after_next = slot + 1 < a.n_emit ? a.emit_steps[slot + 1] : -1;
            }
        }
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
            cudaError_t e = cudaFuncSetAttribute(evolve_rrc2_kernel<PID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return -(int)e;
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, evolve_rrc2_kernel<PID>, G::NT, smem);
            if (e != cudaSuccess) return -(int)e;
            if (per_sm < 1) per_sm = 1;
            int grid = sms * per_sm;
            if (grid > a.n_inst) grid = a.n_inst;
            Params k;
            k.a = a;
            evolve_rrc2_kernel<PID><<<grid, G::NT, smem, stream>>>(k);
            e = cudaGetLastError();
            return e == cudaSuccess ? 0 : -(int)e;
        }
    }
}

#endif  // !__CUDACC_RTC__

}  // namespace rrc2
