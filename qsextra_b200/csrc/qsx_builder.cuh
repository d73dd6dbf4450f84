// qsx_builder.cuh -- host-side lowering of one propagation step into the operation list of qsx_evolve
// (qsx_step_program_size / qsx_step_program_build of include/qsx.h).  Plain C++ on host memory: what
// qsextra_b200/engine/program.py does for the Python binding, restated for callers that are not Python, for the
// model family of the BASELINE configurations: 2-level pseudomodes (or none), damping / dephasing collisions, order-1
// Trotter formula.  Follows qsextra/qcomo/evolution/qevolve.py:74-105 (system block: Z_i with -eps_i/2, then X_iX_j and
// Y_iY_j with J_ij/2 for i < j), :197-253 and :264-289 (for k, for i: mode Z with -omega_k/2, then X_m and X_m Z_e with
// +-g_k/2), :292-348 (collisions: k outer, i inner) and :419-422 (Trotter_number Trotter steps, then one collision round).
// The exact rewritings are the ones of DESIGN.md section 1: Z-type rotations of one Trotter step commute back into one
// diagonal table, XX(th) YY(th) = hop, X_m(th) X_mZ_e(-th) = rotation by 2 th on an occupied site and identity on an
// empty one, mode-ancilla collision + reset = amplitude damping, rzx + reset = dephasing factors.
#pragma once
#include <cmath>
#include <vector>

namespace builder {

static inline int popcount32(unsigned v) {
    int n = 0;
    for (; v; v &= v - 1) ++n;
    return n;
}

struct Shape {
    int n_sites = 0, n_modes = 0, n_bits = 0, n_a = 0, n_b = 0, dim_a = 0, dim_b = 0;
    std::vector<int> cfg_a, cfg_b;
};

static std::vector<int> sector(int n_sites, int k) {
    std::vector<int> out;
    for (int m = 0; m < (1 << n_sites); ++m)
        if (popcount32((unsigned)m) == k) out.push_back(m);
    return out;
}

static int32_t shape_of(const qsx_system_desc* s, int ka, int kb, Shape& sh) {
    if (!s || s->n_sites < 1 || s->n_sites > 16 || s->n_modes < 0 || !s->energies || !s->couplings)
        return fail(QSX_ERR_INVALID, "qsx_system_desc: bad sizes or missing tables%s");
    if (s->n_modes > 0 && (!s->omega || !s->coupl_ep)) return fail(QSX_ERR_INVALID, "qsx_system_desc: pseudomode tables missing%s");
    if (s->trotter_number < 1 || !(s->dt > 0.0)) return fail(QSX_ERR_INVALID, "qsx_system_desc: dt and trotter_number must be positive%s");
    if (ka < 0 || kb < 0 || ka > s->n_sites || kb > s->n_sites) return fail(QSX_ERR_INVALID, "exciton numbers outside 0..N%s");
    sh.n_sites = s->n_sites;
    sh.n_modes = s->n_modes;
    sh.n_bits = s->n_sites * s->n_modes;
    if (sh.n_bits > 24) return fail(QSX_ERR_TOO_LARGE, "more than 24 pseudomode qubits%s");
    sh.cfg_a = sector(s->n_sites, ka);
    sh.cfg_b = sector(s->n_sites, kb);
    sh.n_a = (int)sh.cfg_a.size();
    sh.n_b = (int)sh.cfg_b.size();
    sh.dim_a = sh.n_a << sh.n_bits;
    sh.dim_b = sh.n_b << sh.n_bits;
    return QSX_OK;
}

// Writes the program when `ops` / `params` are given, counts otherwise.
static int32_t lower(const qsx_system_desc* s, const Shape& sh, qsx_op* ops, double* params, int* n_ops_out, int* n_params_out) {
    const int N = sh.n_sites, W = sh.n_modes, nb = sh.n_bits;
    const double dt_trotter = s->dt / s->trotter_number;
    int n_ops = 0, slot = 0;
    auto emit = [&](int type, int i0, int i1, int i2, int i3, int i4, int p) {
        if (ops) {
            qsx_op& o = ops[n_ops];
            o.type = type; o.i0 = i0; o.i1 = i1; o.i2 = i2; o.i3 = i3; o.i4 = i4; o.p = p; o.flags = 0;
        }
        ++n_ops;
    };
    // phase table of one side: exp(-i phase), phase = sum over Z terms of angle * eigenvalue (+1 on |0>, -1 on |1>)
    auto side_table = [&](const std::vector<int>& cfgs, bool with_phases) {
        const int at = slot;
        const int dim = (int)cfgs.size() << nb;
        if (params) {
            for (int c = 0; c < (int)cfgs.size(); ++c)
                for (int bits = 0; bits < (1 << nb); ++bits) {
                    double phase = 0.0;
                    if (with_phases) {
                        for (int i = 0; i < N; ++i) {
                            const double angle = -s->energies[i] / 2 * dt_trotter;
                            phase += ((cfgs[c] >> i) & 1) ? -angle : angle;
                        }
                        for (int i = 0; i < N; ++i)
                            for (int k = 0; k < W; ++k) {
                                const double angle = -s->omega[k] / 2 * dt_trotter;
                                phase += ((bits >> (i * W + k)) & 1) ? -angle : angle;
                            }
                    }
                    const int e = (c << nb) | bits;
                    params[at + 2 * e] = std::cos(phase);
                    params[at + 2 * e + 1] = -std::sin(phase);
                }
        }
        slot += 2 * dim;
        return at;
    };
    for (int rep = 0; rep < s->trotter_number; ++rep) {
        const int ga = side_table(sh.cfg_a, true), gb = side_table(sh.cfg_b, true);
        emit(QSX_OP_DIAG, ga, gb, -1, 0, 0, 0);
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j) {
                const double J = s->couplings[i * N + j];
                if (J == 0.0) continue;
                if (params) {
                    params[slot] = std::cos(J * dt_trotter);                  // 2 th, th = J dt / 2
                    params[slot + 1] = std::sin(J * dt_trotter);
                }
                emit(QSX_OP_HOP, i, j, 0, 0, 0, slot);
                slot += 2;
            }
        for (int k = 0; k < W; ++k) {
            const double g = s->coupl_ep[k];
            if (g == 0.0) continue;
            for (int i = 0; i < N; ++i) {
                if (params) {                                               // empty site: angle 0; occupied: g dt
                    params[slot] = 1.0;
                    params[slot + 1] = 0.0;
                    params[slot + 2] = std::cos(g * dt_trotter);
                    params[slot + 3] = std::sin(g * dt_trotter);
                }
                emit(QSX_OP_PROT, 1 << (i * W + k), 0, 0, i, 1, slot);
                slot += 4;
            }
        }
    }
    if (s->coll_rates) {
        if (W > 0) {
            for (int k = 0; k < W; ++k) {
                if (s->coll_rates[k] == 0.0) continue;
                const double phi = std::sqrt(s->coll_rates[k] * s->dt);
                for (int i = 0; i < N; ++i) {
                    if (params) {
                        params[slot] = std::cos(phi);
                        params[slot + 1] = std::sin(phi) * std::sin(phi);
                    }
                    emit(QSX_OP_AD, i * W + k, 0, 0, 0, 0, slot);
                    slot += 2;
                }
            }
        } else if (s->coll_rates[0] != 0.0) {                               // dephasing: cos^2 - sin^2 per differing site
            const double phi = std::sqrt(s->coll_rates[0]) * std::sqrt(s->dt);
            const double f = std::cos(phi) * std::cos(phi) - std::sin(phi) * std::sin(phi);
            const int ga = side_table(sh.cfg_a, false), gb = side_table(sh.cfg_b, false);
            const int t = slot;
            if (params)
                for (int a = 0; a < sh.n_a; ++a)
                    for (int b = 0; b < sh.n_b; ++b)
                        params[t + a * sh.n_b + b] = std::pow(f, popcount32((unsigned)(sh.cfg_a[a] ^ sh.cfg_b[b])));
            slot += sh.n_a * sh.n_b + ((sh.n_a * sh.n_b) & 1);
            emit(QSX_OP_DIAG, ga, gb, t, 0, 0, 0);
        }
    }
    *n_ops_out = n_ops;
    *n_params_out = slot;
    return QSX_OK;
}

}  // namespace builder

int32_t qsx_step_program_size(const qsx_system_desc* sys, int32_t ka, int32_t kb, int32_t* n_ops, int32_t* n_params,
                              int32_t* n_a, int32_t* n_b, int32_t* n_bits) {
    builder::Shape sh;
    int32_t rc = builder::shape_of(sys, ka, kb, sh);
    if (rc) return rc;
    int ops = 0, params = 0;
    rc = builder::lower(sys, sh, nullptr, nullptr, &ops, &params);
    if (rc) return rc;
    if (n_ops) *n_ops = ops;
    if (n_params) *n_params = params;
    if (n_a) *n_a = sh.n_a;
    if (n_b) *n_b = sh.n_b;
    if (n_bits) *n_bits = sh.n_bits;
    return QSX_OK;
}

int32_t qsx_step_program_build(const qsx_system_desc* sys, int32_t ka, int32_t kb, qsx_op* ops, double* params,
                               int32_t* cfg_a, int32_t* cfg_b, int32_t* rank_a, int32_t* rank_b) {
    if (!ops || !params || !cfg_a || !cfg_b || !rank_a || !rank_b) return fail(QSX_ERR_INVALID, "qsx_step_program_build: null output%s");
    builder::Shape sh;
    int32_t rc = builder::shape_of(sys, ka, kb, sh);
    if (rc) return rc;
    for (int m = 0; m < (1 << sh.n_sites); ++m) rank_a[m] = rank_b[m] = -1;
    for (int c = 0; c < sh.n_a; ++c) {
        cfg_a[c] = sh.cfg_a[c];
        rank_a[sh.cfg_a[c]] = c;
    }
    for (int c = 0; c < sh.n_b; ++c) {
        cfg_b[c] = sh.cfg_b[c];
        rank_b[sh.cfg_b[c]] = c;
    }
    int n_ops = 0, n_params = 0;
    return builder::lower(sys, sh, ops, params, &n_ops, &n_params);
}
