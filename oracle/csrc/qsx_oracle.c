/* TEST INFRASTRUCTURE (oracle) -- plain-C restatement of the reference's execution step.
 *
 * Applies the gate list that oracle/circuits.py derives from qsextra/qcomo/evolution/qevolve.py:74-348 to a
 * FULL-SPACE density matrix (all circuit qubits, collision ancilla included), literally gate by gate:
 *     kind 0: rho -> G rho G^dagger,  G = exp(-i theta P) = cos(theta) - i sin(theta) P,
 *             P = i^{nY} X^x Z^z  (P|v> = i^{nY} (-1)^{|v&z|} |v^x>)
 *     kind 1: reset of qubit x[k] (trace out, replace by |0>)
 *     kind 2: amplitude damping of qubit x[k], phi = theta[k]:  rho -> K0 rho K0^dagger + K1 rho K1^dagger,
 *             K0 = diag(1, cos phi), K1 = -i sin phi |0><1|
 *     kind 3: dephasing of qubit x[k], phi = theta[k]:  rho -> cos^2 phi rho + sin^2 phi Z rho Z
 *   Kinds 2 and 3 are the channels the collision gates of qevolve.py:292-348 induce on the system once their ancilla
 *   (prepared in |0>, reset afterwards) is traced out: exp(-i phi/2 X_m X_a) exp(-i phi/2 Y_m Y_a) + reset (d = 2
 *   pseudomodes) and rzx(2 phi) + reset.  oracle/fast.py substitutes them only when asked to (`kraus=True`: the state
 *   then has no ancilla qubit, a quarter of the memory); tests/ check the substituted list against the literal one.
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

/* (compiled for AVX2+FMA as well as baseline x86-64; the loader picks at run time, so the library runs on any host) */
__attribute__((target_clones("avx2,fma", "default")))
static void rotate_cs(cplx* rho, int D, unsigned x, unsigned z, int ny, double c, double s, int inner) {
    static const cplx ipow[4] = {1, I, -1, -I};
    const cplx unit = ipow[ny & 3];
    (void)inner;
    if (x == 0) {
        if (z == 0) return;
        /* diagonal: rho[a][b] *= g(a) conj(g(b)),  g(v) = c - i s (-1)^{|v&z|} */
#pragma omp parallel for if (inner > 1) num_threads(inner) schedule(static)
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
#pragma omp parallel for if (inner > 1) num_threads(inner) schedule(static)
    for (int ai = 0; ai < D; ++ai) {
        const unsigned a = (unsigned)ai;
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
    /* right: rho[:][b] <- c rho[:][b] + conj(-i s P[b][b^x]) rho[:][b^x]  (row by row: b and b^x share a row) */
#pragma omp parallel for if (inner > 1) num_threads(inner) schedule(static)
    for (int a = 0; a < D; ++a) {
        cplx* row = rho + (size_t)a * D;
        for (unsigned b = 0; b < (unsigned)D; ++b) {
            if (b & pivot) continue;
            unsigned b2 = b ^ x;
            cplx f_b = conj(-I * s * unit * (parity(b2 & z) ? -1.0 : 1.0));
            cplx f_b2 = conj(-I * s * unit * (parity(b & z) ? -1.0 : 1.0));
            cplx u = row[b], v = row[b2];
            row[b] = c * u + f_b * v;
            row[b2] = c * v + f_b2 * u;
        }
    }
}

static void reset_qubit_t(cplx* rho, int D, int q, int inner) {
    const unsigned m = 1u << q;
    (void)inner;
#pragma omp parallel for if (inner > 1) num_threads(inner) schedule(static)
    for (int ai = 0; ai < D; ++ai) {
        const unsigned a = (unsigned)ai;
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


static void damp_qubit(cplx* rho, int D, int q, double c, double s, int inner) {
    const unsigned m = 1u << q;
    const double cc = c * c, ss = s * s;
    (void)inner;
#pragma omp parallel for if (inner > 1) num_threads(inner) schedule(static)
    for (int ai = 0; ai < D; ++ai) {
        const unsigned a = (unsigned)ai;
        if (a & m) continue;
        cplx* r0 = rho + (size_t)a * D;
        cplx* r1 = rho + (size_t)(a | m) * D;
        for (unsigned b = 0; b < (unsigned)D; ++b) {
            if (b & m) continue;
            r0[b] += ss * r1[b | m];
            r1[b | m] *= cc;
            r0[b | m] *= c;
            r1[b] *= c;
        }
    }
}

static void dephase_qubit(cplx* rho, int D, int q, double c, double s, int inner) {
    const unsigned m = 1u << q;
    const double f = c * c - s * s;
    (void)inner;
#pragma omp parallel for if (inner > 1) num_threads(inner) schedule(static)
    for (int a = 0; a < D; ++a)
        for (int b = 0; b < D; ++b)
            if (((unsigned)a ^ (unsigned)b) & m) rho[(size_t)a * D + b] *= f;
}

/* one operation of a step; cs = {cos theta, sin theta} */
static inline void apply_op(cplx* rho, int D, int kind, int x, int z, int ny, const double* cs, int inner) {
    if (kind == 0) rotate_cs(rho, D, (unsigned)x, (unsigned)z, ny, cs[0], cs[1], inner);
    else if (kind == 1) reset_qubit_t(rho, D, x, inner);
    else if (kind == 2) damp_qubit(rho, D, x, cs[0], cs[1], inner);
    else dephase_qubit(rho, D, x, cs[0], cs[1], inner);
}

/* Apply one step (n_ops operations) in place. */
void qsxo_apply_ops(cplx* rho, int n_qubits, int n_ops, const int32_t* kind, const int32_t* x, const int32_t* z,
                    const int32_t* ny, const double* theta) {
    const int D = 1 << n_qubits;
    for (int k = 0; k < n_ops; ++k) {
        const double cs[2] = {cos(theta[k]), sin(theta[k])};
        apply_op(rho, D, kind[k], x[k], z[k], ny[k], cs, 1);
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
        double* cs = (double*)malloc(sizeof(double) * 2 * (size_t)(n_ops > 0 ? n_ops : 1));
        for (int k = 0; k < n_ops; ++k) {          /* the angles of a parameter set do not change from step to step */
            cs[2 * k] = cos(th[k]);
            cs[2 * k + 1] = sin(th[k]);
        }
        int slot = 0;
        for (int step = 0; step <= n_steps && slot < n_emit; ++step) {
            if (step) {
                for (int k = 0; k < n_ops; ++k) apply_op(rho, D, kind[k], x[k], z[k], ny[k], cs + 2 * k, 1);
            }
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
        free(cs);
    }
    return used;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Response function along delay-grid chains, operator form of qsextra/spectroscopy/simulation/qspectroscopy.py:117-183
 * (restated in oracle/reference_path.py: interaction_operators / qspectroscopy_operator): sigma starts as |0..0><0..0|;
 * interaction n applies the collective operator  sum_s mu_s O_n(s)  from the left (side 0, 'k') or from the right
 * (side 1, 'b'); O = X (kind 0), |0><1| (kind 1) or |1><0| (kind 2) on the system qubit of site s; between
 * interactions the block is evolved by the literal step (the gate list above); the last interaction is the emission:
 * signal = Tr[(sum_s mu_s X_s) sigma], read after each step count of `emit_steps` along the last delay.
 * No prefix is shared between chains: every chain is evolved from t = 0 (like the reference does, per grid point).
 * ------------------------------------------------------------------------------------------------------------- */
static void collective(const cplx* in, cplx* out, int D, int side, int kind, int n_sites, const int32_t* sys_qubits,
                       const double* mu, int inner) {
    (void)inner;
#pragma omp parallel for if (inner > 1) num_threads(inner) schedule(static)
    for (int a = 0; a < D; ++a) {
        for (int b = 0; b < D; ++b) {
            cplx acc = 0;
            for (int s = 0; s < n_sites; ++s) {
                const unsigned m = 1u << sys_qubits[s];
                if (side == 0) {            /* (O rho)[a][b] = sum_a' O[a][a'] rho[a'][b] */
                    const int set = (a & m) != 0;
                    if (kind == 0 || (kind == 1 && !set) || (kind == 2 && set)) acc += mu[s] * in[(size_t)(a ^ m) * D + b];
                } else {                    /* (rho O)[a][b] = sum_b' rho[a][b'] O[b'][b] */
                    const int set = (b & m) != 0;
                    if (kind == 0 || (kind == 1 && set) || (kind == 2 && !set)) acc += mu[s] * in[(size_t)a * D + (b ^ m)];
                }
            }
            out[(size_t)a * D + b] = acc;
        }
    }
}

/* n_inter interactions (the last one is the emission), n_inter - 1 delays.  chain_steps[c][d] = step count of delay d
 * (d < n_inter - 2) of chain c; the last delay emits after emit_steps[0..n_emit) (ascending).
 * out[c][e] complex.  Chains run in parallel (OpenMP).  Returns the number of threads used. */
int qsxo_response_chains(int n_qubits, int n_ops, const int32_t* kind, const int32_t* x, const int32_t* z,
                         const int32_t* ny, const double* theta, int n_inter, const int32_t* sides,
                         const int32_t* op_kinds, int n_sites, const int32_t* sys_qubits, const double* mu,
                         int n_chains, const int32_t* chain_steps, int n_emit, const int32_t* emit_steps, cplx* out,
                         int n_threads) {
    const int D = 1 << n_qubits;
    int used = 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    used = omp_get_max_threads();
#endif
    int outer = n_chains < used ? n_chains : used;
    if (outer < 1) outer = 1;
    const int inner = 1;      /* threading the row loops of one gate costs more than it saves at these sizes (measured) */
    double* cs = (double*)malloc(sizeof(double) * 2 * (size_t)(n_ops > 0 ? n_ops : 1));
    for (int k = 0; k < n_ops; ++k) {
        cs[2 * k] = cos(theta[k]);
        cs[2 * k + 1] = sin(theta[k]);
    }
    const int n_delays = n_inter - 1;
#pragma omp parallel for if (outer > 1) num_threads(outer) schedule(dynamic, 1)
    for (int c = 0; c < n_chains; ++c) {
        cplx* rho = (cplx*)calloc((size_t)D * D, sizeof(cplx));
        cplx* tmp = (cplx*)malloc(sizeof(cplx) * (size_t)D * D);
        rho[0] = 1.0;
        for (int n = 0; n < n_inter - 1; ++n) {
            collective(rho, tmp, D, sides[n], op_kinds[n], n_sites, sys_qubits, mu, inner);
            cplx* sw = rho; rho = tmp; tmp = sw;
            const int last_delay = n == n_delays - 1;
            const int steps = last_delay ? (n_emit ? emit_steps[n_emit - 1] : 0) : chain_steps[(size_t)c * (n_delays - 1) + n];
            int slot = 0;
            for (int step = 0; step <= steps; ++step) {
                if (step) {
                    for (int k = 0; k < n_ops; ++k) apply_op(rho, D, kind[k], x[k], z[k], ny[k], cs + 2 * k, inner);
                }
                if (last_delay) {
                    while (slot < n_emit && emit_steps[slot] == step) {
                        /* emission: Tr[(sum_s mu_s X_s) sigma] (the emission operator is X on either side) */
                        cplx acc = 0;
                        for (int s = 0; s < n_sites; ++s) {
                            const unsigned m = 1u << sys_qubits[s];
                            cplx part = 0;
                            for (int v = 0; v < D; ++v) part += rho[(size_t)(v ^ (int)m) * D + v];
                            acc += mu[s] * part;
                        }
                        out[(size_t)c * n_emit + slot] = acc;
                        ++slot;
                    }
                }
            }
        }
        free(rho);
        free(tmp);
    }
    free(cs);
    return outer;
}
