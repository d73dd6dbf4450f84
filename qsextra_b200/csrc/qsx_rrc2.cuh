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
//   * shared-memory exchanges synchronise the CTA (bar.sync 0) and alternate between two buffers.
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
    static constexpr int popc(int v) { return v == 0 ? 0 : (v & 1) + popc(v >> 1); }
    // forward program of the expected shape: DIAG first, dampings last, nothing transposed, tile of at most 16 elements
    static constexpr bool eligible() {
        if (NA * NB > 16 || BP::ops[0][0] != OP_DIAG) return false;
        bool in_tail = false;
        for (int o = 1; o < BP::NOPS; ++o) {
            const int kind = BP::ops[o][0];
            if (kind == OP_ADT || kind == OP_DIAG) return false;
            if (kind == OP_AD) in_tail = true;
            else if (in_tail) return false;
        }
        return true;
    }
    static constexpr int buffer_elems() {
        int need = NA * NB * (NT / 4);
        for (int o = 0; o < BP::NOPS; ++o) {
            const int kind = BP::ops[o][0], bit = BP::ops[o][1], mask = BP::ops[o][3];
            if (kind == OP_ROTK && BP::ket_pos[bit] >= 5 && popc(mask) * NB * NT > need) need = popc(mask) * NB * NT;
            if (kind == OP_ROTB && BP::bra_pos[bit] >= 5 && popc(mask) * NA * NT > need) need = popc(mask) * NA * NT;
        }
        return need;
    }
    static constexpr int BUF = buffer_elems();
    static constexpr int MIN_CTAS = NT >= 256 ? 2 : (512 / NT > 8 ? 8 : 512 / NT);
    static constexpr size_t smem_bytes(int p_doubles) {
        return (size_t)2 * BUF * sizeof(double2) + (size_t)((p_doubles + 1) & ~1) * sizeof(double) +
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

template <int PID, int O>
__device__ __forceinline__ void op2(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double2 g, const double* P,
                                    const double2* H, double2* xbuf, int& cur, int tid) {
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
                        rrc::tan_one(v[i][j], p, t);
                    }
        } else {
            double2* buf = xbuf + cur * G::BUF;
            int e = 0;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((mask >> (ket ? i : j)) & 1) buf[(e++) * NT + tid] = v[i][j];
            __syncthreads();
            e = 0;
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((mask >> (ket ? i : j)) & 1) rrc::tan_one(v[i][j], buf[(e++) * NT + (tid ^ (1 << pos))], t);
            cur ^= 1;
        }
    } else {   // OP_AD without its scalings: the (1,1) threads of the bit pair hand sin^2(phi) x tile to the (0,0) threads
        constexpr int pa = BP::ket_pos[a], pb = BP::bra_pos[a];
        constexpr int hi = pa > pb ? pa : pb, lo = pa > pb ? pb : pa;
        const double ss = P[slot + 1];                                        // sin^2 phi
        const int la = (tid >> pa) & 1, lb = (tid >> pb) & 1;
        const int cid = rrc::remove_bit(rrc::remove_bit(tid, hi), lo);
        double2* buf = xbuf + cur * G::BUF;
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
                    v[i][j].x = fma(ss, p.x, v[i][j].x);
                    v[i][j].y = fma(ss, p.y, v[i][j].y);
                }
        }
        cur ^= 1;
    }
}

template <int PID, int... Os>
__device__ __forceinline__ void step2(double2 (&v)[BigProg<PID>::NA][BigProg<PID>::NB], const double2 g, const double* P,
                                      const double2* H, double2* xbuf, int& cur, int tid, qstd::integer_sequence<int, Os...>) {
    (op2<PID, Os>(v, g, P, H, xbuf, cur, tid), ...);
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
    double2* red = xbuf + 2 * G::BUF;
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
        int cur = 0;
        int slot = 0;
        int next_emit = a.n_emit > 0 ? a.emit_steps[0] : -1;
        int after_next = a.n_emit > 1 ? a.emit_steps[1] : -1;
        for (int step = 0; step <= a.n_steps && slot < a.n_emit; ++step) {
            if (step) step2<PID>(v, step == 1 ? g0 : g1, P, H, xbuf, cur, tid, qstd::make_integer_sequence<int, BP::NOPS>{});
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
