/* TEST INFRASTRUCTURE (oracle) -- plain-C restatement of the reference's execution step.
 *
 * Applies the gate list that oracle/circuits.py derives from qsextra/qcomo/evolution/qevolve.py:74-348 to a
 * FULL-SPACE density matrix (all circuit qubits, collision ancilla included), literally gate by gate:
 *     kind 0: rho -> G rho G^dagger,  G = exp(-i theta P) = cos(theta) - i sin(theta) P,
 *             P = i^{nY} X^x Z^z  (P|v> = i^{nY} (-1)^{|v&z|} |v^x>)
 *     kind 1: reset of qubit x[k] (trace out, replace by |0>)
 * This is what AerSimulator computes in expectation (qevolve.py:546-562).  It shares no code with the CUDA engine
 * (which works on exciton-number sectors with fused operations).  Used by tests/ and by bench.py's cpu_baseline /
 * --impl reference legs only; never by the product.
 *
 * Build: make -C oracle/csrc   (gcc -O3 -fopenmp -shared -fPIC)
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

static inline int parity(unsigned v) { return __builtin_popcount(v) & 1; }

static void rotate(cplx* rho, int D, unsigned x, unsigned z, int ny, double theta) {
    const double c = cos(theta), s = sin(theta);
    static const cplx ipow[4] = {1, I, -1, -I};
    const cplx unit = ipow[ny & 3];
    if (x == 0) {
        if (z == 0) return;
        /* diagonal: rho[a][b] *= g(a) conj(g(b)),  g(v) = c - i s (-1)^{|v&z|} */
        for (int a = 0; a < D; ++a) {
            cplx ga = c - I * s * (parity(a & z) ? -1.0 : 1.0);
            for (int b = 0; b < D; ++b) {
                cplx gb = c - I * s * (parity(b & z) ? -1.0 : 1.0);
                rho[(size_t)a * D + b] *= ga * conj(gb);
            }
        }
        return;
    }
    const unsigned pivot = x & (~x + 1u);
    /* left: rho[a][:] <- c rho[a][:] - i s P[a][a^x] rho[a^x][:],  P[a][a^x] = unit (-1)^{|(a^x)&z|} */
    for (unsigned a = 0; a < (unsigned)D; ++a) {
        if (a & pivot) continue;
        unsigned a2 = a ^ x;
        cplx f_a = -I * s * unit * (parity(a2 & z) ? -1.0 : 1.0);
        cplx f_a2 = -I * s * unit * (parity(a & z) ? -1.0 : 1.0);
        cplx* ra = rho + (size_t)a * D;
        cplx* rb = rho + (size_t)a2 * D;
        for (int b = 0; b < D; ++b) {
            cplx u = ra[b], v = rb[b];
            ra[b] = c * u + f_a * v;
            rb[b] = c * v + f_a2 * u;
        }
    }
    /* right: rho[:][b] <- c rho[:][b] + conj(-i s P[b][b^x]) rho[:][b^x] */
    for (unsigned b = 0; b < (unsigned)D; ++b) {
        if (b & pivot) continue;
        unsigned b2 = b ^ x;
        cplx f_b = conj(-I * s * unit * (parity(b2 & z) ? -1.0 : 1.0));
        cplx f_b2 = conj(-I * s * unit * (parity(b & z) ? -1.0 : 1.0));
        for (int a = 0; a < D; ++a) {
            cplx* row = rho + (size_t)a * D;
            cplx u = row[b], v = row[b2];
            row[b] = c * u + f_b * v;
            row[b2] = c * v + f_b2 * u;
        }
    }
}

static void reset_qubit(cplx* rho, int D, int q) {
    const unsigned m = 1u << q;
    for (unsigned a = 0; a < (unsigned)D; ++a) {
        if (a & m) continue;
        for (unsigned b = 0; b < (unsigned)D; ++b) {
            if (b & m) continue;
            rho[(size_t)a * D + b] += rho[(size_t)(a | m) * D + (b | m)];
            rho[(size_t)a * D + (b | m)] = 0;
            rho[(size_t)(a | m) * D + b] = 0;
            rho[(size_t)(a | m) * D + (b | m)] = 0;
        }
    }
}

/* Apply one step (n_ops operations) in place. */
void qsxo_apply_ops(cplx* rho, int n_qubits, int n_ops, const int32_t* kind, const int32_t* x, const int32_t* z,
                    const int32_t* ny, const double* theta) {
    const int D = 1 << n_qubits;
    for (int k = 0; k < n_ops; ++k) {
        if (kind[k] == 0) rotate(rho, D, (unsigned)x[k], (unsigned)z[k], ny[k], theta[k]);
        else reset_qubit(rho, D, x[k]);
    }
}

/* Batched population dynamics: B parameter sets (theta[B][n_ops]) of one gate structure, common initial state.
 * out[B][n_emit][n_sys] = Prob(system qubit i = 1) after emit_steps[e] steps.  sys_qubits[n_sys] = qubit numbers.
 * Returns the number of OpenMP threads used. */
int qsxo_populations(int n_qubits, int n_sys, const int32_t* sys_qubits, int n_ops, const int32_t* kind,
                     const int32_t* x, const int32_t* z, const int32_t* ny, const double* theta, int B,
                     const cplx* rho0, int n_steps, int n_emit, const int32_t* emit_steps, double* out, int n_threads) {
    const int D = 1 << n_qubits;
    int used = 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    used = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        cplx* rho = (cplx*)malloc(sizeof(cplx) * (size_t)D * D);
        memcpy(rho, rho0, sizeof(cplx) * (size_t)D * D);
        const double* th = theta + (size_t)b * n_ops;
        int slot = 0;
        for (int step = 0; step <= n_steps && slot < n_emit; ++step) {
            if (step) qsxo_apply_ops(rho, n_qubits, n_ops, kind, x, z, ny, th);
            if (emit_steps[slot] == step) {
                for (int i = 0; i < n_sys; ++i) {
                    double p = 0.0;
                    for (int v = 0; v < D; ++v)
                        if ((v >> sys_qubits[i]) & 1) p += creal(rho[(size_t)v * D + v]);
                    out[((size_t)b * n_emit + slot) * n_sys + i] = p;
                }
                ++slot;
            }
        }
        free(rho);
    }
    return used;
}
