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
 * and, for the rows next to the hot path:
 *
 *   qutip.mesolve / sesolve(H, state, time, c_ops, e_ops)             qsextra/qcomo/evolution/clevolve.py:27-46, :88-107
 *   rotating frame, np.pad, fft / ifft, fftshift                      qsextra/spectroscopy/postprocessing.py:9-97
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

#ifdef __CUDACC_RTC__ /* run-time compilation of a kernel specialisation (engine/jit.py): no host headers */
typedef signed char int8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
#else
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define QSX_ABI_VERSION 20

enum {
    QSX_OK = 0,
    QSX_ERR_INVALID = -1,   /* bad argument (message in qsx_last_error) */
    QSX_ERR_CUDA = -2,      /* CUDA runtime error */
    QSX_ERR_TOO_LARGE = -3, /* block does not fit shared memory and no workspace was given */
    QSX_ERR_NO_PROGRAM = -4 /* qsx_evolve_rrc: no compile-time program has this signature (use qsx_evolve) */
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
    QSX_OP_RESET = 4
};

typedef struct qsx_op {
    int32_t type;
    int32_t i0, i1, i2, i3, i4;
    int32_t p;          /* slot (in doubles) into the parameter-set table */
    int32_t flags;      /* QSX_OPF_* */
} qsx_op;

/* The operation is replaced by its transpose with respect to the pairing <W, sigma> = sum_ab W[a][b] sigma[a][b]:
 * a step program run backwards with transposed operations propagates a read-out W instead of the state
 * (<W, Phi^n sigma> = <(Phi^T)^n W, sigma>; used for the last delay of a response function, where one propagation of
 * the emission read-out replaces one propagation per (t1, t2) chain).  DIAG and HOP are their own transposes; PROT
 * changes the sign of its angle when nY is odd; AD becomes W'[11] = cos^2 W[11] + sin^2 W[00], W'[00] = W[00],
 * W'[01], W'[10] scaled by cos (RESET: cos = 0, sin^2 = 1). */
#define QSX_OPF_TRANSPOSE 1

/* Fused passes.  A pass program is an equivalent regrouping of the step's operations into sweeps over sigma in
 * which every thread keeps a small tile in registers and applies several operations to it ("passes"):
 *
 *  QSX_PASS_CFG   tile = all (cfgA, cfgB) pairs at fixed bits.  f = {tile_off, n_tiles, nA, nB, slot GA, slot GB,
 *                 slot T_pre, slot T_post, hop_begin, hop_end}; slots are -1 when absent.  Order inside the pass:
 *                 sigma *= GA conj(GB) T_pre;  hops [hop_begin, hop_end) of `hops`;  sigma *= T_post.
 *                 tiles[2*t] = element offset of the tile, tiles[2*t+1] = bits_a | bits_b << 16.
 *                 hops[4*h] = {side (0 ket, 1 bra), p, q, param slot {cos 2th, sin 2th}}: rotate configurations p, q.
 *  QSX_PASS_MODE  tile = KB (1 or 2) pseudomode bits on both sides.  f = {tile_off, n_tiles, KB, bit0, bit1,
 *                 rot slot0, rot slot1, damp slot0, damp slot1, skip flags}.  Per bit: X-rotation conditioned on the
 *                 occupation of the bit's site (QSX_OP_PROT with a one-bit xmask, no zmask), then amplitude damping.
 *                 tiles[2*t] = element offset, tiles[2*t+1] = occupation flags
 *                 (bit0: site(bit0) in cfgA, bit1: in cfgB, bit2: site(bit1) in cfgA, bit3: in cfgB).
 *  QSX_PASS_OP    f = {index into `ops`}: the operation is applied by the generic sweep.
 */
enum { QSX_PASS_CFG = 0, QSX_PASS_MODE = 1, QSX_PASS_OP = 2 };

typedef struct qsx_pass {
    int32_t type;
    int32_t f[11];
} qsx_pass;

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
    const int32_t* emit_steps;  /* DEVICE [n_emit] strictly increasing step counts in [0, n_steps]: a PRECONDITION the
                                 * library cannot check without a device read-back (the Python binding asserts it on the
                                 * host array); entries out of order leave their output slots untouched */
    int32_t n_obs;              /* read-outs: out_obs[inst][emit][o] = sum_k obs_w[k] * sigma[obs_idx[k]] */
    const int32_t* obs_ptr;     /* DEVICE [n_obs+1] CSR offsets */
    const int32_t* obs_idx;     /* DEVICE element indices a*DB+b */
    const double* obs_w;        /* DEVICE complex128 weights */
    double* out_obs;            /* DEVICE [n_inst][n_emit][n_obs] complex128, or real parts only if obs_real */
    int32_t obs_real;
    double* out_snap;           /* DEVICE [n_inst][n_emit][DA*DB] complex128 snapshots, or NULL */
    double* workspace;          /* DEVICE scratch of qsx_evolve_workspace_bytes() bytes, or NULL */
    int32_t threads;            /* threads per instance (0 = library default); a power of two */
    int32_t n_passes;           /* 0: apply `ops` one by one; > 0: run the pass program below (same result) */
    const qsx_pass* passes;     /* DEVICE [n_passes] */
    int32_t n_tile_ints;        /* length of `tiles` in int32 */
    const int32_t* tiles;       /* DEVICE tile tables of all passes */
    int32_t n_hop_ints;         /* length of `hops` in int32 */
    const int32_t* hops;        /* DEVICE hop entries of all QSX_PASS_CFG passes (p < q) */
} qsx_evolve_args;

/* Register-resident evolution of SMALL blocks whose dimensions are powers of two ("RR kernel").
 * sigma, the diagonal table(s) and nothing else live in registers: every thread of a team of 2^L lanes holds 2^R
 * elements; the physical index of an element is lane << R | reg and `logical_index` maps it to a*DB + b.  An
 * operation combines every element with the element whose physical index differs by the xor mask (reg_xor, lane_xor):
 *   QSX_RR_DIAG  v *= W[table aux0]                      (W per parameter set, physical order, complex128)
 *   QSX_RR_ROT   v <- c v - i sign s partner   on the registers in `reg_mask`   (hops and conditioned X-rotations;
 *                p -> {cos, sin}; the participation never depends on lane bits)
 *   QSX_RR_ROTT  v <- v - i sign t partner, t = tan: the same rotation without its cos factor, which the host has
 *                folded into the diagonal table of the step (scalars commute with every operation); p -> {cos, tan}
 *   QSX_RR_AD    amplitude damping of the bit pair (ket, bra) whose register part is reg_xor and whose lane parts are
 *                aux0 (ket lane mask) and aux1 (bra lane mask); p -> {cos phi, sin^2 phi}
 * Parameter set layout: n_diag tables of E complex128 (physical order) followed by the op parameters (slots p are
 * relative to the start of the set). */
enum { QSX_RR_DIAG = 0, QSX_RR_ROT = 1, QSX_RR_AD = 2, QSX_RR_ROTT = 3 };
#define QSX_RR_MAX_OPS 24
#define QSX_RR_MAX_OBS 8
#define QSX_RR_MAX_ENTRIES 4

typedef struct qsx_rr_op {
    int32_t kind;
    int32_t reg_xor, lane_xor;
    uint32_t reg_mask;
    int32_t p;
    int32_t sign;            /* +1 ket side, -1 bra side */
    int32_t aux0, aux1;
} qsx_rr_op;

typedef struct qsx_rr_args {
    int32_t reg_bits;            /* R: 2^R elements per thread (0..4) */
    int32_t lane_bits;           /* L: team = 2^L lanes (0..5) */
    int32_t n_diag;              /* 0, 1 or 2 diagonal tables */
    int32_t n_ops;
    qsx_rr_op ops[QSX_RR_MAX_OPS];
    const int32_t* logical_index;   /* DEVICE [E] */
    int32_t param_stride;
    const double* params;        /* DEVICE [n_sets][param_stride] */
    int32_t n_inst;
    const int32_t* inst_set;     /* DEVICE or NULL */
    const int32_t* inst_src;     /* DEVICE or NULL */
    const double* sigma_in;      /* DEVICE [n_src][E] complex128, LOGICAL order (same as qsx_evolve) */
    int32_t n_steps, n_emit;
    const int32_t* emit_steps;   /* DEVICE */
    int32_t n_obs;               /* <= QSX_RR_MAX_OBS */
    const double* obs_weights;   /* DEVICE [n_obs][E] complex128 weights in PHYSICAL order */
    uint32_t obs_reg_mask[QSX_RR_MAX_OBS];  /* registers with a non-zero weight on some lane */
    /* Optional factorised form of the read-outs (n_entries[o] > 0 for every o, else the dense weights are used):
     * out[o] = sum_k weight[o][k] * [lane in lane_mask[o][k]] * sum_{r in reg_mask[o][k]} sigma(lane, r). */
    int32_t n_entries[QSX_RR_MAX_OBS];
    uint32_t entry_reg_mask[QSX_RR_MAX_OBS][QSX_RR_MAX_ENTRIES];
    uint32_t entry_lane_mask[QSX_RR_MAX_OBS][QSX_RR_MAX_ENTRIES];
    double entry_weight[QSX_RR_MAX_OBS][QSX_RR_MAX_ENTRIES];
    /* What the read-outs ARE, when the host knows it (lets a compile-time program use compile-time read-out masks):
     * 0 = unspecified (use the tables above), 1 = site populations (n_obs = N, obs_real),
     * 2 = emission trace Tr[(sum_s mu_s X_s) sigma] (n_obs = 1) with the dipole moments in `mu`. */
    int32_t readout_kind;
    double mu[16];
    /* site-resolved pieces of that read-out (site s: registers site_reg_mask[s][e] on lanes site_lane_mask[s][e]);
     * a compile-time program only uses its compile-time read-out masks when they equal these */
    uint32_t site_reg_mask[16][QSX_RR_MAX_ENTRIES];
    uint32_t site_lane_mask[16][QSX_RR_MAX_ENTRIES];
    double* out_obs;             /* DEVICE [n_inst][n_emit][n_obs] complex128 (or real parts if obs_real) */
    int32_t obs_real;
    double* out_snap;            /* DEVICE [n_inst][n_emit][E] complex128 in LOGICAL order, or NULL */
} qsx_rr_args;

/* Same contract as qsx_evolve, for programs the host could map onto the register-resident form. */
int32_t qsx_evolve_rr(const qsx_rr_args* args, void* stream);

/* CTA-wide register-resident evolution of LARGE blocks of systems with one 2-level pseudomode per site (DESIGN.md,
 * "RRC kernel").  Thread t of a CTA of 4^n_bits threads owns the configuration tile sigma[(i, m_a), (j, m_b)], i < nA,
 * j < nB, of ONE pair of pseudomode states for the whole evolution; pseudomode rotations and damping exchange with the
 * thread whose pair differs in one bit (pair), by warp shuffle or through a shared-memory buffer.  Only programs whose
 * `signature` (engine/rrc.py: RRCProgram.signature) equals a compile-time entry of csrc/qsx_rrc_programs.inc can run;
 * the call returns QSX_ERR_NO_PROGRAM otherwise and the caller uses qsx_evolve.
 * Parameter set: Wcfg (nA*nB complex) | gmA (2^n_bits complex) | gmB (2^n_bits complex) | op parameters. */
typedef struct qsx_rrc_args {
    int32_t n_signature;
    const int32_t* signature;    /* HOST */
    int32_t param_stride;
    const double* params;        /* DEVICE [n_sets][param_stride] */
    int32_t n_inst;
    const int32_t* inst_set;     /* DEVICE or NULL */
    const int32_t* inst_src;     /* DEVICE or NULL */
    const double* sigma_in;      /* DEVICE [n_src][DA*DB] complex128, logical order */
    int32_t n_steps, n_emit;
    const int32_t* emit_steps;   /* DEVICE */
    int32_t emit_trace;          /* 1: out_obs[inst][emit] = Tr[(sum_s mu_s X_s) sigma] (complex128), KB = KA -+ 1 blocks;
                                  * 2: out_obs[inst][emit][site] = population of the site (float64), KA = KB blocks */
    double mu[16];
    double* out_obs;             /* DEVICE [n_inst][n_emit] complex128 / [n_inst][n_emit][n_sites] float64, or NULL */
    double* out_snap;            /* DEVICE [n_inst][n_emit][DA*DB] complex128, logical order, or NULL */
    int32_t* progress;           /* DEVICE [n_inst + 1] ZEROED before the call, or NULL.  With it, a launch that only writes
                                  * snapshots and holds more instances than the GPU keeps resident is scheduled by SEGMENTS
                                  * (the steps between two emissions): CTAs claim (instance, segment) pairs in order and
                                  * continue an instance from its last snapshot, so the last round of instances no longer
                                  * leaves SMs idle.  Same results; [i] counts the emissions of instance i done, [n_inst]
                                  * is the claim counter. */
} qsx_rrc_args;

int32_t qsx_evolve_rrc(const qsx_rrc_args* args, void* stream);

/* Index of the compile-time program of qsx_evolve_rrc with this signature (HOST int32 vector), or -1 when the call
 * would answer "no compile-time program matches"; lets the host plan before it enqueues anything. */
int32_t qsx_rrc_find_program(const int32_t* signature, int32_t n_signature);

/* Host-side lowering of one propagation step into the operation list of qsx_evolve, for callers that are not Python
 * (the Python binding lowers in qsextra_b200/engine/program.py and covers more: d > 2 pseudomodes, Suzuki formulas,
 * batches of parameter sets).  Model family: N sites, W two-level pseudomodes per site (0: an excitonic system),
 * order-1 Trotter formula, `trotter_number` Trotter steps of dt / trotter_number followed by one collision round
 * (pseudomode damping with coll_rates[k], or dephasing with coll_rates[0] when n_modes == 0; NULL: no collisions).
 * Reference: qsextra/qcomo/evolution/qevolve.py:74-105, :197-253, :264-348, :419-422.  Everything is HOST memory; the
 * caller uploads `ops`, `params` (one parameter set, stride n_params) and the block tables and fills qsx_evolve_args. */
typedef struct qsx_system_desc {
    int32_t n_sites;
    int32_t n_modes;
    const double* energies;     /* [N] */
    const double* couplings;    /* [N*N] symmetric, zero diagonal */
    const double* omega;        /* [W] pseudomode frequencies, or NULL when n_modes == 0 */
    const double* coupl_ep;     /* [W] exciton-pseudomode couplings, or NULL */
    const double* coll_rates;   /* [W] damping rates / [1] dephasing rate / NULL */
    double dt;                  /* collision step */
    int32_t trotter_number;
} qsx_system_desc;

/* Sizes of the program of the (ka, kb) exciton-number block: operations, doubles per parameter set, configurations per
 * side and pseudomode qubits (block dimensions nA << n_bits, nB << n_bits).  Output pointers may be NULL. */
int32_t qsx_step_program_size(const qsx_system_desc* sys, int32_t ka, int32_t kb, int32_t* n_ops, int32_t* n_params,
                              int32_t* n_a, int32_t* n_b, int32_t* n_bits);

/* Fill ops [n_ops], params [n_params], cfg_a [nA], cfg_b [nB], rank_a / rank_b [2^N] (the tables of qsx_block). */
int32_t qsx_step_program_build(const qsx_system_desc* sys, int32_t ka, int32_t kb, qsx_op* ops, double* params,
                               int32_t* cfg_a, int32_t* cfg_b, int32_t* rank_a, int32_t* rank_b);

/* Add a kernel specialisation compiled at run time (engine/jit.py: NVRTC on csrc/qsx_rrc.cuh + csrc/qsx_rrc2.cuh around
 * the program's own BigProg<0> table) to the programs qsx_evolve_rrc can run: `cubin` is loaded into the current
 * device's primary context, `kernel_name` (and `kernel2_name`, the throughput form, or NULL) are the lowered names of
 * the kernels in it, `signature` (HOST) is what later calls are matched against.  Registering a signature twice is a
 * no-op.  This is what opens the compile-time menu: any eligible block structure runs the register-resident kernels. */
int32_t qsx_rrc_register(const void* cubin, int64_t cubin_bytes, const char* kernel_name, const char* kernel2_name,
                         const int32_t* signature, int32_t n_signature);

int32_t qsx_abi_version(void);

/* Message of the last failing call on this host thread ("" if none). */
const char* qsx_last_error(void);

/* sm_count, max opt-in dynamic shared memory per block (bytes), compute capability major*10+minor. */
int32_t qsx_device_info(int32_t* sm_count, int32_t* smem_optin, int32_t* cc);

/* Scratch needed by qsx_evolve when a block does not fit shared memory (0 if it fits).  Such a block is evolved by the
 * streaming path: sigma of one instance at a time lives in the scratch (L2 / HBM), every operation is one grid-wide
 * sweep, and the emission schedule is followed on the device.  That path synchronises `stream` once (it reads the
 * operation types back to decide how many sweeps to enqueue); everything else is enqueued. */
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

/* Parameter rows of a bound step program evaluated on the device from the raw angles of the step (what
 * StepProgram.bind computes on the host: the cos/sin pairs of every rotation and the phase tables of the merged Z
 * rotations).  Column c of every row:  phase = sum_{t in [ptr[c], ptr[c+1])} coef[t] * angles[set][idx[t]];
 *   kind 0: 0   1: cos(phase)   2: sin(phase)   3: -sin(phase)   4: sin(phase)^2
 *   kind 5: prod_t (cos^2 - sin^2)(angles[set][idx[t]])          (dephasing factors)
 * All pointers DEVICE; angles [n_sets][n_raw], out [n_sets][n_cols]. */
int32_t qsx_bind_params(int32_t n_sets, int32_t n_cols, int32_t n_raw, const int32_t* kind, const int32_t* ptr,
                        const int32_t* idx, const double* coef, const double* angles, double* out, void* stream);

/* Parameter sets of the register-resident kernel assembled on the device from the per-set operation parameters
 * (the angle tables StepProgram.bind produced): the diagonal tables in physical element order,
 *   w[e] = GA[a_of[e]] * conj(GB[b_of[e]]) * T[cfg(a_of[e])][cfg(b_of[e])] * prod_c cos_c ^ count_c[e]   (first table),
 * followed by the operation parameters with sin replaced by tan for the rotations run in tangent form.
 * One launch; replaces the gather/multiply sequence the host used to issue op by op. */
typedef struct {
    int32_t n_sets;
    int32_t n_elem;              /* E */
    int32_t n_bits;              /* pseudomode bits per side */
    int32_t n_cfg_b;             /* configurations on the bra side */
    int32_t n_diag;              /* <= 2 */
    int32_t slot_ga[2], slot_gb[2], slot_t[2];   /* into a row of op_params; slot_t = -1: no dephasing table */
    int32_t n_classes;           /* <= 8 cosine classes folded into the first table */
    int32_t class_slot[8];
    const int8_t* class_count;   /* DEVICE [n_classes][E] exponents */
    int32_t n_tangent;           /* <= 24 */
    int32_t tangent_slot[24];    /* row[slot + 1] = sin / cos */
    const int32_t* a_of;         /* DEVICE [E] */
    const int32_t* b_of;         /* DEVICE [E] */
    int32_t op_stride;           /* doubles per row of op_params */
    const double* op_params;     /* DEVICE [n_sets][op_stride] */
    double* out;                 /* DEVICE [n_sets][2*n_diag*E + op_stride] */
} qsx_rr_tables_args;

int32_t qsx_rr_tables(const qsx_rr_tables_args* args, void* stream);

/* Signal of the last delay from the propagated read-outs (engine/response.py, Heisenberg picture):
 *   out[i][k] = sum_e sigma[i][e] * readout[k][e]      (no conjugation)
 * sigma [n_inst][E], readout [n_emit][E], out [n_inst][n_emit], all complex128 DEVICE arrays.  It stands for the
 * accumulation `signal[*time_indices] += ...` over the estimator values at qspectroscopy.py:180-183.
 * The sums over e are split over several CTAs when the output alone is too small to fill the GPU; `workspace` (DEVICE,
 * qsx_readout_gemm_workspace_bytes bytes, may be NULL: no split) holds the partial sums, added in a fixed order. */
int64_t qsx_readout_gemm_workspace_bytes(int32_t n_inst, int32_t n_emit, int32_t n_elem);
int32_t qsx_readout_gemm(int32_t n_inst, int32_t n_emit, int32_t n_elem, const double* sigma, const double* readout,
                         double* out, double* workspace, void* stream);
/* The same for `n_batch` independent products in one launch -- the parameter sets of an ensemble (static-disorder
 * realisations of one aggregate), each with its own propagated read-outs:
 * sigma [n_batch][n_inst][E], readout [n_batch][n_emit][E], out [n_batch][n_inst][n_emit]. */
int64_t qsx_readout_gemm_batched_workspace_bytes(int32_t n_batch, int32_t n_inst, int32_t n_emit, int32_t n_elem);
int32_t qsx_readout_gemm_batched(int32_t n_batch, int32_t n_inst, int32_t n_emit, int32_t n_elem, const double* sigma,
                                 const double* readout, double* out, double* workspace, void* stream);

/* Lindblad limit of the collision model -- replaces the QuTiP solver calls of the comparison path
 * (qsextra/qcomo/evolution/clevolve.py:27-46 and :88-107: sesolve / mesolve(H, state, time, c_ops, e_ops)).
 * sigma' = L sigma with the sparse generator of one sector block in ELL storage; segment k propagates by seg_dt[k]
 * (scaled Taylor series of exp(hL), `terms` terms, sub-steps h <= theta / norm1) and is followed by emission k. */
typedef struct {
    int32_t E;                   /* elements of the block (DA*DB) */
    int32_t width;               /* ELL width (max entries per row) */
    const int32_t* ell_idx;      /* DEVICE [width][E] column of entry k of row r at [k*E + r] */
    const double* ell_val;       /* DEVICE [width][E] complex128, 0 for padding */
    double norm1;                /* upper bound of ||L||_1 */
    int32_t terms;               /* Taylor terms per sub-step (>= 2; 0: 20) */
    double theta;                /* largest ||hL||_1 of a sub-step (0: 1.0) */
    int32_t n_inst;
    const int32_t* inst_src;     /* DEVICE [n_inst] source block per instance, or NULL (= instance index) */
    const double* sigma_in;      /* DEVICE [n_src][E] complex128 */
    int32_t n_seg;               /* number of segments = number of emissions */
    const double* seg_dt;        /* HOST [n_seg] durations (>= 0; 0 emits the current state again) */
    int32_t n_obs;               /* CSR read-outs, as qsx_evolve_args */
    const int32_t* obs_ptr;
    const int32_t* obs_idx;
    const double* obs_w;
    int32_t obs_real;
    double* out_obs;             /* DEVICE [n_inst][n_seg][n_obs] float64 (obs_real) or complex128, or NULL */
    double* out_snap;            /* DEVICE [n_inst][n_seg][E] complex128, or NULL */
    double* workspace;           /* DEVICE qsx_lindblad_workspace_bytes(args) bytes */
} qsx_lindblad_args;

int64_t qsx_lindblad_workspace_bytes(const qsx_lindblad_args* args);
int32_t qsx_lindblad_evolve(const qsx_lindblad_args* args, void* stream);

/* Post-processing of the response function (qsextra/spectroscopy/postprocessing.py:9-37 rotating_frame, and the
 * zero padding of fourier_transform :57 / :77), one pass:
 *   linear (one delay axis):  out[i]    = R[i] e^{+i rf t1[i]} e^{-damping t1[i]}                     i < n1, else 0
 *   third order:              out[i][j] = R[i][t2_index][j] e^{-i rf (t1[i]-t3[j])} e^{-damping (t1[i]+t3[j])}, else 0
 * `out` is [n1*pad] or [n1*pad][n3*pad] complex128; all pointers DEVICE. */
typedef struct {
    int32_t n1, n2, n3;          /* signal [n1][n2][n3] complex128 (linear: n2 = n3 = 1) */
    int32_t t2_index;
    int32_t pad;                 /* pad_extension >= 1 */
    int32_t linear;
    double rf_freq, damping;
    const double* t1;
    const double* t3;            /* NULL when linear */
    const double* signal;
    double* out;
} qsx_frame_args;

int32_t qsx_rotating_frame(const qsx_frame_args* args, void* stream);

/* out[i][j] = scale * in[(i+s1) % m1][(j+s3) % m3] on [m1][m3] complex128 DEVICE arrays (in != out): the
 * fftshift (s = m - m/2) / ifftshift (s = m/2) and the prefactor of postprocessing.py:61 and :84-85. */
int32_t qsx_spectrum_shift(const double* in, double* out, int32_t m1, int32_t m3, int32_t s1, int32_t s3, double scale,
                           void* stream);

/* Unrolled DFMA micro-benchmark: achieved FP64 TFLOP/s of the CUDA cores on the current device
 * (the roofline denominator of the on-chip-resident kernels; host pointer). Synchronises. */
int32_t qsx_fp64_peak(int32_t repeats, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* QSX_H */
