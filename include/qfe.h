/*
 * qfe.h -- C ABI of libqfe.so, the B200-native Pauli-sum engine behind myqlm-fermion's
 * Hamiltonian hot path.  Plain pointers and sizes only; no C++ or torch types cross this line.
 *
 * The reference (pure Python, /root/reference) has no FFI today; each entry point below names the
 * reference interface it replaces (file:line relative to the reference checkout).  The ctypes
 * binding a maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions (all taken from the reference):
 *   - qubit q of an n-qubit register is mask bit (n-1-q): qubit 0 is the most significant
 *     bit of the amplitude index            (qat/fermion/hamiltonians.py:231-235)
 *   - |1> = occupied                         (qat/fermion/util.py:251,267)
 *   - Y = [[0,-i],[i,0]]                     (qat/fermion/hamiltonians.py:20-25)
 *   - a packed term is (x, z, c): the operator  c * i^{popc(x&z)} * X^x Z^z, i.e. ``c`` is the
 *     coefficient of the Pauli string with letters X (x only), Z (z only), Y (both) -- exactly
 *     ``Term.coeff`` of the reference.  Term sets are canonical: merged, ascending in (x, z),
 *     |c| <= drop_eps removed, identity folded into ``constant``.
 *   - complex numbers are interleaved (re, im) doubles (numpy complex128) unless a ``dtype``
 *     argument says QFE_C64 (interleaved floats, numpy complex64).
 *   - statevector pointers are DEVICE pointers owned by the caller (e.g. torch CUDA tensors);
 *     integral / term-list inputs and all scalar outputs are HOST pointers.
 *   - every function returns 0 on success or a negative qfe_status; qfe_last_error() gives the
 *     message (thread-local).  No C++ exception crosses the ABI.
 *   - one qfe_ctx per host thread and device; handles are not thread-safe.
 */
#ifndef QFE_H
#define QFE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qfe_ctx qfe_ctx;     /* device, stream, scratch buffers                        */
typedef struct qfe_terms qfe_terms; /* canonical packed Pauli-term set (one or a batch of operators) */
typedef struct qfe_plan qfe_plan;   /* sweep schedule of a term set for one shard geometry     */

typedef enum {
    QFE_OK = 0,
    QFE_ERR_ARG = -1,         /* bad argument (maps to ValueError / TypeError in the facade)  */
    QFE_ERR_CUDA = -2,        /* CUDA runtime error                                            */
    QFE_ERR_OOM = -3,         /* host or device allocation failed                              */
    QFE_ERR_UNSUPPORTED = -4, /* e.g. more than 64 qubits, unknown encoding (KeyError)         */
    QFE_ERR_STATE = -5,       /* handle used in the wrong state                                */
    QFE_ERR_COMM = -6         /* NCCL error / NCCL not available                               */
} qfe_status;

enum { QFE_ENC_JW = 0, QFE_ENC_BK = 1, QFE_ENC_PARITY = 2 };
enum { QFE_C128 = 0, QFE_C64 = 1 };

const char* qfe_last_error(void);
int qfe_version(void);

/* ---- context ---------------------------------------------------------------------------- */
int qfe_ctx_create(qfe_ctx** ctx, int device);
/* run on a caller-owned stream (cudaStream_t passed as void*; NULL = the ctx's own stream)   */
int qfe_ctx_set_stream(qfe_ctx* ctx, void* cuda_stream);
int qfe_ctx_sync(qfe_ctx* ctx);
int qfe_ctx_destroy(qfe_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches)                    */
int qfe_ctx_launch_count(const qfe_ctx* ctx, int64_t* count);

/* ---- kernel 1: fermion -> packed Pauli terms (integer expand + canonicalise) -------------- */
/* Replaces ElectronicStructureHamiltonian._get_fermionic_terms + to_spin(method)
 * (qat/fermion/hamiltonians.py:625-648, 427-454) and transform_to_{jw,bk,parity}_basis
 * (qat/fermion/transforms.py:82-257).
 *   hpq   : host, n*n complex128, C order                 (entries with |h| <= tol skipped)
 *   hpqrs : host, n^4 complex128, C order, or NULL        (term weight is 0.5*h, operator
 *           a+_p a+_q a_r a_s, as hamiltonians.py:646)
 *   tol   : the reference's TOL = 1e-12 (hamiltonians.py:541)
 *   drop_eps : canonicalisation threshold on merged Pauli coefficients                        */
int qfe_terms_from_integrals(qfe_ctx* ctx, int nqbits, const double* hpq, const double* hpqrs,
                             const double constant[2], double tol, double drop_eps, int encoding,
                             qfe_terms** out);
/* General FermionHamiltonian term list (qat/fermion/hamiltonians.py:295-312 objects):
 *   ops     : host, concatenated operator letters, 'C' (creation) or 'c' (annihilation)
 *   qbits   : host, mode index of every letter
 *   offsets : host, n_terms+1 prefix offsets into ops/qbits (a term has at most 12 letters)
 *   coeffs  : host, n_terms complex128                                                        */
int qfe_terms_from_fermion(qfe_ctx* ctx, int nqbits, int64_t n_terms, const uint8_t* ops,
                           const int32_t* qbits, const int64_t* offsets, const double* coeffs,
                           const double constant[2], double drop_eps, int encoding, qfe_terms** out);
/* SpinHamiltonian term list already in packed form (qat/fermion/hamiltonians.py:92-105);
 * duplicates are merged, so this is also Observable-style like-term collection.               */
int qfe_terms_from_pauli(qfe_ctx* ctx, int nqbits, int64_t n_terms, const uint64_t* x, const uint64_t* z,
                         const double* coeffs, const double constant[2], double drop_eps, qfe_terms** out);
/* Batch of K operators evaluated together (operator pool of AdaptVQEPlugin,
 * qat/plugins/adapt_vqe.py:50-62,181-185): output slot k collects operator k.                 */
int qfe_terms_concat(qfe_ctx* ctx, const qfe_terms* const* list, int64_t K, qfe_terms** out);
int qfe_terms_info(const qfe_terms* t, int* nqbits, int64_t* n_terms, int64_t* n_xmasks, int64_t* n_slots);
/* arrays sized from qfe_terms_info; owner may be NULL (slot of every term, 0 for a single operator);
 * constant = 2*n_slots doubles                                                                 */
int qfe_terms_export(const qfe_terms* t, uint64_t* x, uint64_t* z, double* coeffs, int32_t* owner, double* constant);
int qfe_terms_destroy(qfe_terms* t);

/* ---- kernel 2: matrix-free Pauli-sum apply / expectation ---------------------------------- */
/* Replaces SpinHamiltonian.get_matrix (qat/fermion/hamiltonians.py:163-250) followed by
 * matrix-vector products / <psi|H|psi> at the reference's call sites
 * (qat/plugins/adapt_vqe.py:182,206; qat/fermion/vqe.py:53-59;
 *  qat/plugins/sequential_optimization.py:117-139).
 *
 * Geometry: the statevector of n qubits is split by its (n - n_local) most significant index
 * bits (= qubits 0..n-n_local-1) into 2^(n-n_local) shards; ``shard`` is this rank's shard
 * index.  A term whose X mask has high part x_hi maps shard r to shard r^x_hi ("class x_hi").
 * tile_bits = 0 picks the default tile size.                                                   */
int qfe_plan_create(qfe_ctx* ctx, const qfe_terms* t, int n_local, int shard, int tile_bits, qfe_plan** out);
int qfe_plan_info(const qfe_plan* p, int64_t* n_sweeps, int64_t* n_groups, int64_t* n_terms, int32_t* n_classes, int32_t* tile_bits);
int qfe_plan_geometry(const qfe_plan* p, int* nqbits, int* n_local);
/* work of one <psi|H|psi> on this plan, for the roofline model of bench.py: groups (X masks) of the tiled layout that are visited on
 * half the rows of a tile (Hermitian pairs) and on all rows, and the staged terms of both kinds, summed over the classes      */
/* 1 when every term coefficient and constant of the planned set is real, i.e. every x_hi class is a Hermitian operator: <psi|H|psi>
 * then needs each pair of shards (r, r^x_hi) on one of the two ranks only                                                    */
int qfe_plan_hermitian(const qfe_plan* p, int* hermitian);
int qfe_plan_work(const qfe_plan* p, int64_t* half_groups, int64_t* full_groups, int64_t* half_terms, int64_t* full_terms);
/* operators (output slots) of the planned term set: 1, or K for a qfe_terms_concat batch        */
int qfe_plan_slots(const qfe_plan* p, int64_t* n_slots);
/* the distinct x_hi classes, ascending (class 0 first when present)                            */
int qfe_plan_classes(const qfe_plan* p, int32_t* x_hi, int32_t capacity);
int qfe_plan_destroy(qfe_plan* p);

/* out[2*s .. 2*s+1] += sum over terms of class x_hi in slot s of <bra_rows| c P |ket>, where
 * ``bra`` is this shard of the bra vector and ``ket`` the shard (shard ^ x_hi) of the ket vector.
 * ``out`` (host) has 2*n_slots doubles and is OVERWRITTEN.  For a single-slot set this is the
 * class's share of <bra|H|ket>; the constant contributes constant*<bra|ket> in class 0.        */
int qfe_expectation_class(qfe_ctx* ctx, const qfe_plan* p, int x_hi, const void* bra, const void* ket, int dtype, double* out);
/* the same restricted to every n_parts-th sweep of the class, starting with sweep `part`: the n_parts calls add up to
 * qfe_expectation_class.  Two ranks that hold the shards r and r^x_hi of a Hermitian operator each evaluate one half (their own shard
 * as the bra) and count it twice: <psi_r|G|psi_r'> + <psi_r'|G|psi_r> = 2 Re <psi_r|G_A|psi_r'> + 2 Re <psi_r'|G_B|psi_r>, G = G_A + G_B */
int qfe_expectation_class_part(qfe_ctx* ctx, const qfe_plan* p, int x_hi, const void* bra, const void* ket, int dtype, int part, int n_parts, double* out);
/* whole-vector convenience (n_local == n): <psi|H|psi> in out[0..1]                            */
int qfe_expectation(qfe_ctx* ctx, const qfe_plan* p, const void* psi, int dtype, double out[2]);
/* same through HOST memory (end-to-end path): psi_host is uploaded inside the call, in 16 pieces on a copy stream; the first sweeps
 * whose tiles do not involve the 4 top qubits run piece by piece behind the upload, which hides most of the copy                  */
int qfe_expectation_host(qfe_ctx* ctx, const qfe_plan* p, const void* psi_host, int dtype, double out[2]);
/* sharded form of the same: uploads this process's shard (2^n_local amplitudes) from psi_host into the caller's device buffer
 * psi_dev with the same pipeline and returns the class x_hi = 0 share of <psi|H|psi> in out[0..1] (0 if no term is local); the other
 * classes follow with qfe_expectation_class[_part] on psi_dev.  Single-operator term sets only.                                     */
int qfe_expectation_class0_host(qfe_ctx* ctx, const qfe_plan* p, const void* psi_host, void* psi_dev, int dtype, double out[2]);
/* y_rows (+)= H_class * ket ; accumulate = 0 overwrites y, 1 adds to it                        */
int qfe_apply_class(qfe_ctx* ctx, const qfe_plan* p, int x_hi, const void* ket, void* y, int dtype, int accumulate);
int qfe_apply(qfe_ctx* ctx, const qfe_plan* p, const void* psi, void* y, int dtype);

/* ADAPT-VQE gradient sweep (qat/plugins/adapt_vqe.py:181-185): g_k = -i<psi|[H,tau_k]|psi>
 * = 2 Im <H psi|tau_k|psi> for Hermitian H, tau_k.  ``scratch`` holds one vector (H psi).      */
int qfe_commutator_gradients(qfe_ctx* ctx, const qfe_plan* h_plan, const qfe_plan* pool_plan,
                             const void* psi, void* scratch, int dtype, double* grads);

/* Development aid, not part of the reference-facing surface: with QFE_PROF=1 in the environment the sweep kernels count, per warp of a
 * CTA (16) and phase of the tile loop (8 slots: wait for the tile, issue/own rows, group loops, barrier skew, loads), the cycles spent;
 * cycles = 16*8 values summed over the CTAs and launches since the last reset.  All zero without QFE_PROF.                         */
int qfe_dev_phase_cycles(qfe_ctx* ctx, uint64_t* cycles, int reset);

/* ---- kernel 3: Lanczos ground state on the matrix-free apply ------------------------------ */
/* Replaces eigvalsh / eigsh(get_matrix(sparse=True), which="SA") at the reference's call sites
 * (tests/integration/test_integration_models.py:102-103, test_integration_adaptvqe.py:97).
 *   v       : device, in: start vector (any norm), out: Ritz vector when want_vector != 0
 *   work    : device, 3 vectors of the same size as v
 *   eval    : host, lowest Ritz value; iters: iterations done; resid: |beta_m * s_m| estimate   */
int qfe_lanczos(qfe_ctx* ctx, const qfe_plan* p, void* v, void* work, int max_iter, double tol,
                int dtype, int want_vector, double* eval, int* iters, double* resid);

/* vector helpers on device shards (used by the sharded Lanczos driver of the Python host)     */
int qfe_vec_dot(qfe_ctx* ctx, const void* a, const void* b, int64_t n_amps, int dtype, double out[2]); /* <a|b> */
int qfe_vec_axpy(qfe_ctx* ctx, const double alpha[2], const void* x, void* y, int64_t n_amps, int dtype); /* y += alpha x */
int qfe_vec_scale(qfe_ctx* ctx, const double alpha[2], void* x, int64_t n_amps, int dtype);
/* G[i][j] = <bras[i]|kets[j]> for nb x nk device vectors of n_amps amplitudes (a multiple of 32) in ONE pass per panel of 32 x 32
 * vectors: out = 2*nb*nk doubles, row-major.  kets = NULL: the Gram matrix of the bras (Hermitian; nk is ignored, out is nb x nb).
 * bras / kets are HOST arrays of device pointers.  Replaces the per-entry overlap circuits of the natural-gradient optimizer
 * (qat/fermion/naturalgradient/qfim_plugin.py:159-199) and the dim^2 observables of the subspace expansion (chemistry/qse.py:191-214). */
int qfe_vec_gram(qfe_ctx* ctx, const void* const* bras, int nb, const void* const* kets, int nk, int64_t n_amps, int dtype, double* out);
int qfe_vec_fill_random(qfe_ctx* ctx, void* x, int64_t n_amps, int dtype, uint64_t seed, uint64_t offset); /* counter-based N(0,1)+iN(0,1) */
int qfe_vec_product_state(qfe_ctx* ctx, void* x, int n_local, int shard, int nqbits, const double* alpha_beta /* 4*n doubles */, int dtype);
/* computational-basis state on a buffer of n_amps amplitudes: all zero, amplitude `index` one (index = -1: all zero; a shard that does
 * not hold the occupied amplitude).  Replaces the X-gate layer the reference prepends to its ansatz circuits
 * (Hartree-Fock state, qat/fermion/chemistry/ucc.py:291-343)                                                                    */
int qfe_vec_basis_state(qfe_ctx* ctx, void* x, int64_t n_amps, int dtype, int64_t index);

/* ---- multi-GPU: one process per GPU, the statevector sharded by its high qubits -----------------------------------
 * The context owns an NCCL communicator.  Rank 0 makes a unique id, the host hands it to every rank by any channel (the Python host:
 * torch.distributed broadcast; an MPI or file rendezvous works as well), then every rank calls qfe_comm_init -- collectively.  Plans
 * are created with n_local = nqbits - log2(nranks) and shard = rank.  `partner` is device scratch owned by the caller for the shard of
 * the rank ^ x_hi partner: one shard (exchange, then sweeps) or two (partner_bytes >= 2 shards: the exchange of the next class runs
 * on a second stream while the sweeps of the current class run).  Every call below is collective over the communicator, and results
 * are identical on every rank.  NCCL is bound at run time (dlopen of libnccl.so.2; QFE_NCCL_LIB overrides): QFE_ERR_COMM if missing. */
int qfe_comm_unique_id(uint8_t id[128]);
int qfe_comm_init(qfe_ctx* ctx, const uint8_t id[128], int rank, int nranks);
int qfe_comm_info(const qfe_ctx* ctx, int* rank, int* nranks, int* nccl_version, int64_t* exchanges, int64_t* exchanged_bytes);
int qfe_comm_destroy(qfe_ctx* ctx);
/* attribution aid: the x_hi classes of the last qfe_expectation_sharded[_host] on this rank and the host milliseconds each took
 * (wait for the exchange + sweeps); n = number of classes, at most `capacity` are written                                          */
int qfe_comm_class_times(const qfe_ctx* ctx, int32_t* x_hi, double* ms, int32_t capacity, int32_t* n);
/* in-place sum of n host doubles over the ranks (one ncclAllReduce); identity without a communicator                               */
int qfe_allreduce_sum(qfe_ctx* ctx, double* values, int n);
/* partner <- the shard of rank ^ x_hi of the vector `shard` belongs to (ncclSend + ncclRecv in one group, on the ctx stream)        */
int qfe_exchange_shard(qfe_ctx* ctx, const void* shard, void* partner, int64_t n_amps, int dtype, int x_hi);
/* <bra|H|ket> of the whole register (bra_shard NULL: <ket|H|ket>; a Hermitian single-operator plan then splits every non-local class
 * between the two ranks of a pair); out = 2*n_slots doubles.  Replaces the same call sites as qfe_expectation_class.                */
int qfe_expectation_sharded(qfe_ctx* ctx, const qfe_plan* p, const void* bra_shard, const void* ket_shard, void* partner, int64_t partner_bytes,
                            int dtype, double* out);
/* the same with this rank's shard arriving from HOST memory: uploaded into shard_dev behind the sweeps of the local class             */
int qfe_expectation_sharded_host(qfe_ctx* ctx, const qfe_plan* p, const void* shard_host, void* shard_dev, void* partner, int64_t partner_bytes,
                                 int dtype, double out[2]);
/* y_shard = this rank's shard of H|ket>                                                                                             */
int qfe_apply_sharded(qfe_ctx* ctx, const qfe_plan* p, const void* ket_shard, void* y_shard, void* partner, int64_t partner_bytes, int dtype);
/* ADAPT-VQE gradient sweep on a sharded state (qat/plugins/adapt_vqe.py:181-185); scratch = one shard (H psi)                        */
int qfe_commutator_gradients_sharded(qfe_ctx* ctx, const qfe_plan* h_plan, const qfe_plan* pool_plan, const void* psi_shard, void* scratch,
                                     void* partner, int64_t partner_bytes, int dtype, double* grads);
int qfe_vec_dot_sharded(qfe_ctx* ctx, const void* a, const void* b, int64_t n_amps, int dtype, double out[2]); /* <a|b>, all-reduced */
/* qfe_lanczos on shards (work = 3 shards): the dots and norms of the recurrence are all-reduced                                     */
int qfe_lanczos_sharded(qfe_ctx* ctx, const qfe_plan* p, void* v, void* work, void* partner, int64_t partner_bytes, int max_iter, double tol,
                        int dtype, int want_vector, double* eval, int* iters, double* resid);

/* ---- state evolution ----------------------------------------------------------------------------------------------
 * In place psi <- exp(-i theta[K-1] P[K-1]) ... exp(-i theta[0] P[0]) psi on an unsharded state of nqbits qubits, with
 * P[k] = i^{|x&z|} X^x Z^z (the packed-term convention, coefficient 1).  One call = the reference's
 * make_spin_hamiltonian_trotter_slice(H, coeff) (qat/fermion/trotterisation.py:85-150) with theta[k] = coeff * Re(c_k) in the term
 * order of H, or one layer exp(-i theta_k O_k) of the ADAPT / UCC ansatz (term by term, _grow_ansatz) (qat/plugins/adapt_vqe.py:64-141).  Consecutive rotations
 * whose X masks fit 10 common qubits besides the 3 lowest are fused into one pass over the state; a rotation that flips more qubits than that
 * (long parity-encoding X strings) takes a pass of its own through global memory.                                               */
int qfe_pauli_rotations(qfe_ctx* ctx, int nqbits, int64_t K, const uint64_t* x, const uint64_t* z, const double* theta, void* psi, int dtype);

#ifdef __cplusplus
}
#endif
#endif /* QFE_H */
