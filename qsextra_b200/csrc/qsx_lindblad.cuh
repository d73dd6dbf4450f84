// Lindblad limit (reference comparison path clevolve.py:16-108, QuTiP mesolve): sigma(t+h) = exp(hL) sigma by a scaled
// Taylor series.  One launch per term: y_k = (h/k) L y_{k-1}, a batched sparse matrix-vector product with the
// generator in ELL storage ([width][E], so the loads of a warp over consecutive rows coalesce) shared by all instances.
// Each thread owns one row for IPT instances, so an ELL entry is loaded once per IPT gathers.  The running sum
// acc += y_k is folded into the next term's launch (a thread adds its own row of the input vector), so a term costs
// one pass: E*width*(16 + 4) bytes of generator (L2-resident) + 3 vectors of 16*E bytes per instance.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lind {

constexpr int IPT = 4;

struct TermParams {
    int E, width, n_inst;
    int first, last;                 // first term reads acc itself (no accumulate), last term finishes acc
    double scale;
    const int* __restrict__ idx;
    const double2* __restrict__ val;
    const double2* x;                // first ? acc : y_{k-1}
    double2* y;                      // y_k (not written by the last term)
    double2* acc;
};

__global__ void __launch_bounds__(128) term_kernel(TermParams p) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.E) return;
    for (int base = blockIdx.y * IPT; base < p.n_inst; base += gridDim.y * IPT) {
        double2 sum[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) sum[q] = make_double2(0.0, 0.0);
        for (int k = 0; k < p.width; ++k) {
            const int j = p.idx[(size_t)k * p.E + r];
            const double2 v = p.val[(size_t)k * p.E + r];
#pragma unroll
            for (int q = 0; q < IPT; ++q) {
                if (base + q < p.n_inst) {
                    const double2 xv = p.x[(size_t)(base + q) * p.E + j];
                    sum[q].x = fma(v.x, xv.x, fma(-v.y, xv.y, sum[q].x));
                    sum[q].y = fma(v.x, xv.y, fma(v.y, xv.x, sum[q].y));
                }
            }
        }
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            if (base + q >= p.n_inst) continue;
            const size_t e = (size_t)(base + q) * p.E + r;
            const double2 s = make_double2(sum[q].x * p.scale, sum[q].y * p.scale);
            double2 a = p.acc[e];
            if (!p.first) {                          // the previous term's vector: only this thread touches acc[e]
                const double2 prev = p.x[e];
                a.x += prev.x;
                a.y += prev.y;
            }
            if (p.last) {
                a.x += s.x;
                a.y += s.y;
            } else {
                p.y[e] = s;
            }
            if (!p.first || p.last) p.acc[e] = a;
        }
    }
}

// sigma_in[inst_src[i]] -> acc[i]
__global__ void load_kernel(int E, int n_inst, const int* __restrict__ inst_src, const double2* __restrict__ in, double2* acc) {
    const size_t total = (size_t)E * n_inst;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / E), r = (int)(e % E);
        const int src = inst_src ? inst_src[i] : i;
        acc[e] = in[(size_t)src * E + r];
    }
}

// Emission k: snapshots and CSR read-outs (one warp per (instance, read-out)).
struct EmitParams {
    int E, n_inst, n_emit, k, n_obs, obs_real;
    const int* obs_ptr;
    const int* obs_idx;
    const double2* obs_w;
    const double2* acc;
    double* out_obs;
    double2* out_snap;
};

__global__ void emit_kernel(EmitParams p) {
    if (p.out_snap) {
        const size_t total = (size_t)p.E * p.n_inst;
        for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
            const size_t i = e / p.E, r = e % p.E;
            p.out_snap[(i * p.n_emit + p.k) * p.E + r] = p.acc[e];
        }
    }
    if (p.out_obs) {
        const int lane = threadIdx.x & 31;
        const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
        const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
        for (size_t u = warp; u < (size_t)p.n_inst * p.n_obs; u += n_warps) {
            const size_t i = u / p.n_obs;
            const int o = (int)(u % p.n_obs);
            double re = 0.0, im = 0.0;
            for (int t = p.obs_ptr[o] + lane; t < p.obs_ptr[o + 1]; t += 32) {
                const double2 w = p.obs_w[t];
                const double2 v = p.acc[i * p.E + p.obs_idx[t]];
                re += w.x * v.x - w.y * v.y;
                im += w.x * v.y + w.y * v.x;
            }
            for (int d = 16; d > 0; d >>= 1) {
                re += __shfl_xor_sync(0xffffffffu, re, d);
                im += __shfl_xor_sync(0xffffffffu, im, d);
            }
            if (lane == 0) {
                const size_t at = (i * p.n_emit + p.k) * p.n_obs + o;
                if (p.obs_real) p.out_obs[at] = re;
                else { p.out_obs[2 * at] = re; p.out_obs[2 * at + 1] = im; }
            }
        }
    }
}

}  // namespace lind
