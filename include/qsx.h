/*
 * qsx.h -- C ABI of the B200 collision-model / Feynman-pathway engine (libqsx.so).
 *
 * The reference (federicogallina/qsextra) has no FFI: its hot path reaches the arithmetic through three
 * third-party call sites, which these entry points replace:
 *
 *   AerSimulator(method='statevector').run(qc_list).result()          qsextra/qcomo/evolution/qevolve.py:546-552
 *   AerSimulator().run(qc_list, shots=shots).result()                 qsextra/qcomo/evolution/qevolve.py:556-562
 *   Estimator(run_options={'shots': shots}).run(circuits, obs)...     qsextra/spectroscopy/simulation/qspectroscopy.py:114, :180-181
 *
 * Model.  One "instance" is the ket x bra block sigma[a][b] (complex128, row-major, b fastest) of the density
 * operator (qevolve) or of the ket-bra coherence (qspectroscopy) restricted to exciton-number sectors:
 *   a = cfgA_index << n_bits | bits_a ,  b = cfgB_index << n_bits | bits_b
 * where cfg* enumerate the N-bit occupation masks of the sector in increasing order and `bits` are the
 * pseudomode (and, for d>2 collisions, ancilla) qubits.  A step program is a list of operations
 * (qsx_op) applied as sigma -> G sigma G^dagger, or a channel; see the QSX_OP_* codes.
 *
 * Conventions: every pointer marked DEVICE is caller-allocated device memory on the current CUDA device;
 * the library allocates nothing persistent and keeps no global state besides the thread-local error string.
 * Calls only enqueue work on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 * Every function returns 0 on success, a negative QSX_ERR_* code otherwise; no exceptions cross the boundary.
 * There is no CPU fallback: without a usable CUDA device every compute entry point fails.
 */
#ifndef QSX_H
#define QSX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSX_ABI_VERSION 1

enum {
    QSX_OK = 0,
    QSX_ERR_INVALID = -1,   /* bad argument (message in qsx_last_error) */
    QSX_ERR_CUDA = -2,      /* CUDA runtime error */
    QSX_ERR_TOO_LARGE = -3  /* block does not fit shared memory and no workspace was given */
};

/* Operation codes (qsx_op.type).  Angles live in the per-parameter-set table `params` at slot `p`. */
enum {
    /* sigma[a][b] *= GA[a] * conj(GB[b]) * (T ? T[cfgA(a)][cfgB(b)] : 1)
     *   i0 = slot of GA (DA complex), i1 = slot of GB (DB complex), i2 = slot of T (nA*nB doubles) or -1.
     *   Merged Z-type rotations (qevolve.py:85-88, :210-217) and dephasing collisions (qevolve.py:302-304). */
    QSX_OP_DIAG = 0,
    /* exp(-i th (X_i X_j + Y_i Y_j)): i0 = i, i1 = j, p -> {cos 2th, sin 2th}   (qevolve.py:89-94) */
    QSX_OP_HOP = 1,
    /* exp(-i th P), P = i^{nY} X^x Z^z on the bits; i0 = xmask, i1 = zmask, i2 = nY & 3,
     *   i3 = conditioning site or -1, i4 = 1 to skip when the site is empty,
     *   p -> {cos th0, sin th0, cos th1, sin th1} (site empty / occupied)  (qevolve.py:239-252, :334-346) */
    QSX_OP_PROT = 2,
    /* amplitude damping of bit i0: K0 = diag(1, cos phi), K1 = -i sin phi |0><1|; p -> {cos phi, sin^2 phi}
     *   (2-level pseudomode collision + reset, qevolve.py:334-347) */
    QSX_OP_AD = 3,
    /* trace out bit i0 and replace it by |0>  (reset, qevolve.py:347) */
    QSX_OP_RESET = 4,
    /* fused step of the standard 2-level-pseudomode / Markovian program: reserved */
    QSX_OP_FUSED = 5
};

typedef struct qsx_op {
    int32_t type;
    int32_t i0, i1, i2, i3, i4;
    int32_t p;          /* slot (in doubles) into the parameter-set table */
    int32_t reserved;
} qsx_op;

/* Geometry of one ket x bra sector block. */
typedef struct qsx_block {
    int32_t n_sites;        /* N */
    int32_t n_bits;         /* pseudomode (+ancilla) qubits, identical on both sides */
    int32_t nA, nB;         /* number of occupation configurations on the ket / bra side */
    const int32_t* cfgA;    /* DEVICE [nA] occupation masks, increasing */
    const int32_t* cfgB;    /* DEVICE [nB] */
    const int32_t* rankA;   /* DEVICE [2^N] mask -> index in cfgA, or -1 */
    const int32_t* rankB;   /* DEVICE [2^N] */
} qsx_block;

/* One batched evolution launch: every instance runs `n_steps` steps of the same program, with its own
 * parameter set and initial block, and emits read-outs / snapshots after the step counts in `emit_steps`. */
typedef struct qsx_evolve_args {
    qsx_block block;
    int32_t n_ops;
    const qsx_op* ops;          /* DEVICE [n_ops] one propagation step */
    int32_t param_stride;       /* doubles per parameter set */
    const double* params;       /* DEVICE [n_sets][param_stride] */
    int32_t n_inst;
    const int32_t* inst_set;    /* DEVICE [n_inst] parameter set of each instance, or NULL (all use set 0) */
    const int32_t* inst_src;    /* DEVICE [n_inst] initial block of each instance, or NULL (instance i uses block i) */
    const double* sigma_in;     /* DEVICE [n_src][DA*DB] complex128 interleaved */
    int32_t n_steps;
    int32_t n_emit;
    const int32_t* emit_steps;  /* DEVICE [n_emit] strictly increasing step counts in [0, n_steps] */
    int32_t n_obs;              /* read-outs: out_obs[inst][emit][o] = sum_k obs_w[k] * sigma[obs_idx[k]] */
    const int32_t* obs_ptr;     /* DEVICE [n_obs+1] CSR offsets */
    const int32_t* obs_idx;     /* DEVICE element indices a*DB+b */
    const double* obs_w;        /* DEVICE complex128 weights */
    double* out_obs;            /* DEVICE [n_inst][n_emit][n_obs] complex128, or real parts only if obs_real */
    int32_t obs_real;
    double* out_snap;           /* DEVICE [n_inst][n_emit][DA*DB] complex128 snapshots, or NULL */
    double* workspace;          /* DEVICE scratch of qsx_evolve_workspace_bytes() bytes, or NULL */
    int32_t threads;            /* threads per instance (0 = library default) */
} qsx_evolve_args;

int32_t qsx_abi_version(void);

/* Message of the last failing call on this host thread ("" if none). */
const char* qsx_last_error(void);

/* sm_count, max opt-in dynamic shared memory per block (bytes), compute capability major*10+minor. */
int32_t qsx_device_info(int32_t* sm_count, int32_t* smem_optin, int32_t* cc);

/* Scratch needed by qsx_evolve when a block does not fit shared memory (0 if it fits). */
int64_t qsx_evolve_workspace_bytes(const qsx_evolve_args* args);

/* Batched collision-model evolution (replaces AerSimulator.run / Estimator.run, see file header). */
int32_t qsx_evolve(const qsx_evolve_args* args, void* stream);

/* Collective dipole operator of one light-matter interaction (qspectroscopy.py:65-80, :142-174):
 *   side 0 (ket):  out[a'][b] = sum_s mu[s] in[a' with site s flipped][b]
 *   side 1 (bra):  out[a][b'] = sum_s mu[s] in[a][b' with site s flipped]
 * restricted to source configurations inside `in_blk`'s sector (this selects X / sigma+ / sigma-).
 * `mu` is DEVICE [N] doubles; blocks are [n_inst][D*D'] complex128. */
int32_t qsx_apply_dipole(const qsx_block* in_blk, const qsx_block* out_blk, int32_t side, const double* mu,
                         int32_t n_inst, const double* sigma_in, double* sigma_out, void* stream);

/* Unrolled DFMA micro-benchmark: achieved FP64 TFLOP/s of the CUDA cores on the current device
 * (the roofline denominator of the on-chip-resident kernels; host pointer). Synchronises. */
int32_t qsx_fp64_peak(int32_t repeats, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* QSX_H */
