/*
 * agf2_b200.h -- C-ABI of libagf2_b200.so: the B200 (sm_100a) replacement for the reference's
 * native moment-build library  auxgf/lib/libagf2/agf2.so  (reference paths are relative to the
 * obackhouse/auxgf source tree).
 *
 * Two layers:
 *
 *  (1) DROP-IN symbols with the reference's exact names and signatures (host pointers, `+=`
 *      semantics, [istart, iend) slice of the occupied-pole loop).  A build of auxgf can load this
 *      library in place of agf2.so without touching auxgf/lib/agf2.py:
 *          build_part_loop_rhf   replaces  auxgf/lib/libagf2/optragf2.c:32-43
 *          build_part_loop_uhf   replaces  auxgf/lib/libagf2/optuagf2.c:31-47
 *      They return void like the reference and there is NO CPU fallback: a failure (no CUDA device,
 *      CUDA error) leaves vv/vev untouched, prints the message to stderr and records a sticky status
 *      that agf2_b200_sticky_status() / agf2_b200_last_error() return (the caller then fails on the
 *      Cholesky of a zero vv, as the reference does on any non-SPD vv).  Set the environment variable
 *      AGF2_B200_ABORT_ON_ERROR=1 to abort() the process instead.
 *
 *  (2) STATUS-RETURNING entry points (`agf2_b200_*`): same arithmetic, 64-bit sizes, explicit
 *      SCS factors, device-resident variants (inputs already in HBM, no host<->device copies),
 *      the on-device pole compression (cuSOLVER potrf/trsm/syevd; replaces the NumPy
 *      cholesky/inv/eigh of auxgf/agf2/optragf2.py:270-278) and a dense symmetric eigensolver for
 *      the extended Fock matrix (replaces np.linalg.eigh in auxgf/aux/eig.py:23-33).
 *      Return value: 0 on success, an AGF2_B200_E* code otherwise; agf2_b200_last_error() gives
 *      the text (per calling thread).
 *
 *  (3) CONTEXTS and the COMMUNICATOR.  An agf2_b200_ctx owns the CUDA streams, scratch memory,
 *      instrumentation and -- optionally -- one rank of an NCCL communicator (one context per GPU;
 *      NCCL is loaded at run time, the library does not link it).  Several contexts per process are
 *      fine (one per device, or more); calls on one context are serialised by the library.  The
 *      entry points without a ctx argument use the DEFAULT context of the calling thread's current
 *      CUDA device (created on first use); pass NULL for `ctx` to name it explicitly.  Replaces the
 *      OpenMP team + mpi4py communicator of the reference (optragf2.c:95-115, auxgf/util/mpi.py:25-69).
 *      Threading contract: any thread may call; a context is used by one call at a time; scratch
 *      buffers are shared between calls of a context, so asynchronous device calls (sync = 0) on one
 *      context must all be given the same stream.
 *
 * Array conventions (identical to the reference, auxgf/lib/agf2.py:63-145):
 *      ixq   (nocc*nphys, naux)   row-major f64,  element [(i, x), Q]     = (i x | Q)
 *      qja   (naux, nocc*nvir)    row-major f64,  element [Q, (j, a)]     = (Q | j a)
 *      ei    (nocc)  energies of the poles playing the "occupied" role,  ea (nvir) the "virtual" role
 *      vv, vev (nphys, nphys)     row-major f64,  accumulated into (+=)
 * Arithmetic (auxgf/lib/libagf2/optragf2.c:117-295), with X[x,i,j,a] = sum_Q ixq[i,x,Q] qja[Q,j,a]:
 *      vv [x,y] += sum_{i in [istart,iend)} sum_{j,a} X[x,i,j,a] ( (os+ss) X[y,i,j,a] - ss X[y,j,i,a] )
 *      vev[x,y] += the same with the weight (ei[i] + ei[j] - ea[a])
 *   the reference hard-codes os = ss = 1 (optragf2.c:217,237); os/ss enter as in
 *   auxgf/aux/dfrmp2.py:82-84.
 */
#ifndef AGF2_B200_H
#define AGF2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    AGF2_B200_OK = 0,
    AGF2_B200_ENODEV = 1,    /* no CUDA device / not sm_100 */
    AGF2_B200_ECUDA = 2,     /* CUDA runtime / driver error */
    AGF2_B200_EINVAL = 3,    /* bad argument (sizes, alignment, null pointer) */
    AGF2_B200_ENOMEM = 4,    /* device allocation failed */
    AGF2_B200_ENOTSPD = 5,   /* vv is not positive definite (Cholesky failed; reference raises LinAlgError) */
    AGF2_B200_ESOLVER = 6,   /* cuSOLVER / cuBLAS failure */
    AGF2_B200_ECOMM = 7      /* NCCL failure / no communicator */
};

/* ---- (1) drop-in symbols: reference signatures, host memory -------------------------------- */

/* replaces auxgf/lib/libagf2/optragf2.c:32-43 */
void build_part_loop_rhf(double *ixq, double *qja, double *ei, double *ea,
                         uint32_t nphys, uint32_t nocc, uint32_t nvir, uint32_t naux,
                         uint32_t istart, uint32_t iend, double *vv, double *vev);

/* replaces auxgf/lib/libagf2/optuagf2.c:31-47 */
void build_part_loop_uhf(double *ixq_a, double *qja_a, double *qja_b,
                         double *ei_a, double *ei_b, double *ea_a, double *ea_b,
                         uint32_t nphys, uint32_t nocc_a, uint32_t nocc_b,
                         uint32_t nvir_a, uint32_t nvir_b, uint32_t naux,
                         uint32_t istart, uint32_t iend, double *vv, double *vev);

/* ---- (2) status-returning API --------------------------------------------------------------- */

const char *agf2_b200_last_error(void);
const char *agf2_b200_version(void);

/* First failure of a void drop-in call since the last clear (0 = none); agf2_b200_last_error() then
 * returns its text.  clear != 0 resets it. */
int agf2_b200_sticky_status(int clear);

/* cudaSetDevice for the calling thread: the entry points without a ctx argument then use that
 * device's default context. */
int agf2_b200_set_device(int device);

/* ---- (3) contexts ------------------------------------------------------------------------------ */
typedef struct agf2_b200_ctx agf2_b200_ctx;

/* A new context on `device` (< 0: the current device) with its own streams, scratch and options
 * (initialised from the process defaults). */
int agf2_b200_ctx_create(int device, agf2_b200_ctx **out);
int agf2_b200_ctx_destroy(agf2_b200_ctx *ctx);
int agf2_b200_ctx_get_device(agf2_b200_ctx *ctx, int *device);
int agf2_b200_ctx_synchronize(agf2_b200_ctx *ctx);

/* Options by name: "workspace_bytes", "overlap", "deterministic", "coulomb_factorization" (0/1/2),
 * "simple_kernels", "thin_tiles", "lpt_order", "tune_chunk", "profile", "visiting_bytes", "shard_step_limit".  ctx == NULL: the process
 * defaults, which the default contexts of all devices share. */
int agf2_b200_ctx_set_option(agf2_b200_ctx *ctx, const char *name, int64_t value);

/* Instrumentation of the last moments call on a context (NULL: default context): see the variants
 * without ctx below; comm bytes = what this rank sent over NVLink: [0] all-gather, [1] all-reduce. */
int agf2_b200_ctx_last_stage_ms(agf2_b200_ctx *ctx, double *ms4);
int agf2_b200_ctx_last_plan(agf2_b200_ctx *ctx, int64_t *info6);
int agf2_b200_ctx_last_comm_bytes(agf2_b200_ctx *ctx, int64_t *bytes2);
/* FP64 tensor-core flops the launches of this context EXECUTED since the last reset (tile padding included):
 * [0] k1_form, [1] k2_moment (pair moments, M build, Coulomb step), [2] general GEMM. */
int agf2_b200_ctx_flop_counters(agf2_b200_ctx *ctx, double *flops3, int reset);

/* ---- communicator: one rank per context, NCCL over NVLink / NVSwitch -------------------------------
 * Replaces comm.Reduce + comm.Bcast of auxgf/util/mpi.py:60-64 (auxgf/mpi/agf2/optragf2.py:394-395).
 * Rank 0 obtains the 128-byte id, the caller distributes it (file, socket, launcher env ...), every
 * rank calls comm_init (collective).  No MPI is involved. */
int agf2_b200_comm_get_unique_id(void *id128);
int agf2_b200_comm_init(agf2_b200_ctx *ctx, const void *id128, int rank, int nranks);
int agf2_b200_comm_info(agf2_b200_ctx *ctx, int *rank, int *nranks);
/* in-place collectives on DEVICE f64 buffers, asynchronous on `stream`; op: 0 sum, 1 max, 2 min;
 * all-gather: rank r's piece is buf[r * count_per_rank ...]. */
int agf2_b200_comm_allreduce(agf2_b200_ctx *ctx, double *buf, int64_t count, int op, void *stream);
int agf2_b200_comm_allgather(agf2_b200_ctx *ctx, double *buf, int64_t count_per_rank, void *stream);
int agf2_b200_comm_broadcast(agf2_b200_ctx *ctx, double *buf, int64_t count, int root, void *stream);
/* every rank has arrived and the device is idle */
int agf2_b200_comm_barrier(agf2_b200_ctx *ctx);
int agf2_b200_comm_destroy(agf2_b200_ctx *ctx);

/* Moment build on a context with a communicator: this rank builds the pair-work items
 * [rank, nranks) and ONE ncclAllReduce(sum, f64) of the packed [vv | vev] follows k3_finalize on the
 * same stream; every rank ends up with the full result added into its vv/vev.  Without a communicator:
 * identical to the plain device-resident calls.  ei_host (may be NULL): host copy of the pole energies
 * -- saves a device->host copy and a stream synchronisation at the start of the call. */
int agf2_b200_ctx_moments_rhf_dev(agf2_b200_ctx *ctx, const double *ixq, int64_t ld_ixq, const double *qja,
                                  int64_t ld_qja, const double *ei, const double *ea, const double *ei_host,
                                  int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux, int64_t istart,
                                  int64_t iend, double os_factor, double ss_factor, double *vv, double *vev,
                                  void *stream, int sync);
int agf2_b200_ctx_moments_uhf_dev(agf2_b200_ctx *ctx, const double *ixq_a, int64_t ld_ixq, const double *qja_a,
                                  int64_t ld_qja_a, const double *qja_b, int64_t ld_qja_b, const double *ei_a,
                                  const double *ei_b, const double *ea_a, const double *ea_b,
                                  const double *ei_a_host, const double *ei_b_host, int64_t nphys,
                                  int64_t nocc_a, int64_t nocc_b, int64_t nvir_a, int64_t nvir_b, int64_t naux,
                                  int64_t istart, int64_t iend, double os_factor, double ss_factor, double *vv,
                                  double *vev, void *stream, int sync);

/* POLE-SHARDED build, for shapes whose (ix|Q) tensor does not fit one GPU (BASELINE configs[4]: nocc=250,
 * nvir=2500, naux=10000, nphys=2750 -- the virtual-sector (ix|Q) is 550 GB).  Rank r of the context's communicator
 * holds only the poles [lo, hi) = agf2_b200_pole_block(nocc, r, nranks) of (ix|Q) (`ixq_block`, (hi - lo, nphys,
 * ld_ixq)); qja, ei, ea are complete on every rank.  Pairs inside a block are formed by its owner; for the pairs
 * between two blocks the poles of one block VISIT the other GPU in sub-blocks ("visiting_bytes" option, default 1 GiB,
 * two buffers) sent by ncclSend/ncclRecv over NVLink on a copy stream while the previous sub-block is being worked
 * on; the two owners of a block pair split its pairs by the parity of i + j.  The Coulomb-type terms go through the
 * M0/M1 factorisation (M rows built 1/N per rank and all-gathered, every rank treats its own poles), then the one
 * all-reduce of [vv | vev].  Same result as the replicated build to rounding; replaces the replicated host arrays of
 * auxgf/mpi/agf2/optragf2.py:370-374.  Whole pole range only (no [istart, iend) slice), RHF only. */
int agf2_b200_pole_block(int64_t nocc, int64_t rank, int64_t nranks, int64_t *lo, int64_t *hi);
int agf2_b200_ctx_moments_rhf_dev_sharded(agf2_b200_ctx *ctx, const double *ixq_block, int64_t ld_ixq,
                                          const double *qja, int64_t ld_qja, const double *ei, const double *ea,
                                          const double *ei_host, int64_t nphys, int64_t nocc, int64_t nvir,
                                          int64_t naux, double os_factor, double ss_factor, double *vv,
                                          double *vev, void *stream, int sync);
/* Host-only replay of the pole-sharded schedule (no GPU): cover (nocc x nocc int64, accumulated into) counts how
 * often the unordered pair {i, j} is formed by any of the `nranks` ranks at [max(i,j), min(i,j)] (must be exactly 1),
 * stats (nranks x 2) = pairs formed / poles sent per rank, for sub-blocks of `sub_block` poles.  limit > 0: the
 * sampled schedule of the "shard_step_limit" option (only the first `limit` sub-blocks of every ring step -- a
 * bounded benchmark sample of a build that would take minutes; the result is then NOT the full moments). */
int agf2_b200_shard_cover(int64_t nocc, int64_t nranks, int64_t sub_block, int64_t limit, int64_t *cover,
                          int64_t *stats);
/* poles per visiting sub-block a pole-sharded build of this shape uses on the context ("visiting_bytes" option) */
int64_t agf2_b200_ctx_shard_sub_block(agf2_b200_ctx *ctx, int64_t nocc, int64_t nphys, int64_t ld_ixq, int64_t nranks);
/* bytes this rank SENT point-to-point (visiting pole blocks) in the last call on the context */
int64_t agf2_b200_ctx_last_p2p_bytes(agf2_b200_ctx *ctx);

/* Host-pointer build on a context with a communicator -- the multi-GPU form of the drop-in call: every
 * rank passes the same host arrays; rank r copies 1/N of ixq (groups of poles, pipelined under the
 * pair loop) and of qja over PCIe, NCCL all-gathers replicate them over NVLink, the ranks build their
 * shares, one all-reduce, and every rank adds the full result into its host vv/vev. */
int agf2_b200_ctx_build_part_loop_rhf(agf2_b200_ctx *ctx, const double *ixq, const double *qja,
                                      const double *ei, const double *ea, int64_t nphys, int64_t nocc,
                                      int64_t nvir, int64_t naux, int64_t istart, int64_t iend,
                                      double os_factor, double ss_factor, double *vv, double *vev);
int agf2_b200_ctx_build_part_loop_uhf(agf2_b200_ctx *ctx, const double *ixq_a, const double *qja_a,
                                      const double *qja_b, const double *ei_a, const double *ei_b,
                                      const double *ea_a, const double *ea_b, int64_t nphys, int64_t nocc_a,
                                      int64_t nocc_b, int64_t nvir_a, int64_t nvir_b, int64_t naux,
                                      int64_t istart, int64_t iend, double os_factor, double ss_factor,
                                      double *vv, double *vev);

/* General FP64 GEMM on the library's own DMMA / TMA / stream-K pipeline (the kernel behind the Coulomb
 * factorisation's T = ixq [M0|M1], the DF transform and the exchange matrix):
 *     C (a_rows x b_rows, row-major, pitch ldc; transpose != 0: C^T is stored)  =  A (a_rows x k) B (b_rows x k)^T
 * (+= when accumulate != 0).  Device pointers, 16-byte aligned, k contiguous, lda/ldb even.  Bit-reproducible. */
int agf2_b200_ctx_gemm(agf2_b200_ctx *ctx, const double *a, int64_t a_rows, int64_t lda, const double *b,
                       int64_t b_rows, int64_t ldb, int64_t k, double *c, int64_t ldc, int transpose,
                       int accumulate, void *stream);

/* Optional: page-lock a caller-owned host buffer (cudaHostRegister) so that the uploads of the
 * host-pointer calls are asynchronous DMA transfers; pageable memory works too. */
int agf2_b200_host_register(const void *ptr, int64_t bytes);
int agf2_b200_host_unregister(const void *ptr);

/* Workspace for the (xi|ja) chunk that flows between the two kernels, in bytes per buffer
 * (a ring of buffers per stream is kept; a call allocates no more than its work items need).  Default 384 MiB:
 * long chunks amortise launch tails and the stream-K prologue/epilogue; the chunk streams through L2 either way
 * (measured: profiles/r02k_workspace_sweep.log, r02f_l2_k1_k2_nocachectl.csv). */
int agf2_b200_set_workspace_bytes(int64_t bytes);

/* Host-memory variants with SCS factors and 64-bit sizes (copies inputs to HBM, runs, adds the
 * result into the host vv/vev). */
int agf2_b200_build_part_loop_rhf(const double *ixq, const double *qja, const double *ei,
                                  const double *ea, int64_t nphys, int64_t nocc, int64_t nvir,
                                  int64_t naux, int64_t istart, int64_t iend, double os_factor,
                                  double ss_factor, double *vv, double *vev);

int agf2_b200_build_part_loop_uhf(const double *ixq_a, const double *qja_a, const double *qja_b,
                                  const double *ei_a, const double *ei_b, const double *ea_a,
                                  const double *ea_b, int64_t nphys, int64_t nocc_a,
                                  int64_t nocc_b, int64_t nvir_a, int64_t nvir_b, int64_t naux,
                                  int64_t istart, int64_t iend, double os_factor,
                                  double ss_factor, double *vv, double *vev);

/* The same two calls restricted to the pair-work items [part, nparts) (see the device-resident variants below):
 * one call per GPU / process, the caller sums the per-part vv/vev (replaces the per-rank [istart, iend) slices and
 * Reduce + Bcast of auxgf/mpi/agf2/optragf2.py:370-395 without the cost of non-symmetric slices). */
int agf2_b200_build_part_loop_rhf_sharded(const double *ixq, const double *qja, const double *ei,
                                          const double *ea, int64_t nphys, int64_t nocc, int64_t nvir,
                                          int64_t naux, int64_t istart, int64_t iend, double os_factor,
                                          double ss_factor, int64_t part, int64_t nparts, double *vv,
                                          double *vev);

int agf2_b200_build_part_loop_uhf_sharded(const double *ixq_a, const double *qja_a, const double *qja_b,
                                          const double *ei_a, const double *ei_b, const double *ea_a,
                                          const double *ea_b, int64_t nphys, int64_t nocc_a, int64_t nocc_b,
                                          int64_t nvir_a, int64_t nvir_b, int64_t naux, int64_t istart,
                                          int64_t iend, double os_factor, double ss_factor, int64_t part,
                                          int64_t nparts, double *vv, double *vev);

/* Device-resident variants: every pointer is DEVICE memory on the selected device.
 *   ixq : (nocc, nphys, ld_ixq) with ld_ixq >= naux, ld_ixq even   (row pitch in doubles)
 *   qja : (naux, nocc, ld_qja)  with ld_qja >= nvir, ld_qja even
 *   vv, vev : (nphys, nphys) contiguous, accumulated into (+=).
 * Work on the pole slice [istart, iend) can be restricted further to the pair-work items
 * [part, nparts) -- items are dealt round-robin -- which is how the moment build is sharded over
 * the GPUs of one box (replaces the MPI split of auxgf/mpi/agf2/optragf2.py:370-374); the caller
 * sums the per-GPU vv/vev (one all-reduce).  Use part = 0, nparts = 1 for everything.
 * `stream` is a cudaStream_t, used as given (0 = the CUDA default stream).  All work is ordered on that
 * stream; the library's scratch buffers are shared between calls, so concurrent calls must use the
 * same stream.  The call is asynchronous with respect to the host unless `sync` is non-zero. */
int agf2_b200_moments_rhf_dev(const double *ixq, int64_t ld_ixq, const double *qja, int64_t ld_qja,
                              const double *ei, const double *ea, int64_t nphys, int64_t nocc,
                              int64_t nvir, int64_t naux, int64_t istart, int64_t iend,
                              double os_factor, double ss_factor, int64_t part, int64_t nparts,
                              double *vv, double *vev, void *stream, int sync);

int agf2_b200_moments_uhf_dev(const double *ixq_a, int64_t ld_ixq, const double *qja_a,
                              int64_t ld_qja_a, const double *qja_b, int64_t ld_qja_b,
                              const double *ei_a, const double *ei_b, const double *ea_a,
                              const double *ea_b, int64_t nphys, int64_t nocc_a, int64_t nocc_b,
                              int64_t nvir_a, int64_t nvir_b, int64_t naux, int64_t istart,
                              int64_t iend, double os_factor, double ss_factor, int64_t part,
                              int64_t nparts, double *vv, double *vev, void *stream, int sync);

/* Pole compression on the device (replaces auxgf/agf2/optragf2.py:270-278):
 *   b = chol(vv)^T ; m = b^-T vev b^-1 ; (e, u) = eigh(m) ; c = b^T u
 * vv, vev: (nphys, nphys) f64 symmetric.  e: (nphys), c: (nphys, nphys) row-major with
 * c c^T = vv and c diag(e) c^T = vev.  `on_device` != 0: all four pointers are device memory. */
int agf2_b200_compress(const double *vv, const double *vev, int64_t nphys, double *e, double *c,
                       int on_device);

/* Dense symmetric eigensolver (cuSOLVER dsyevd) for the extended Fock matrix
 * (replaces np.linalg.eigh in auxgf/aux/eig.py:23-33).  a: (n, n) symmetric row-major, untouched;
 * w: (n) ascending; v: (n, n) row-major with eigenvectors in COLUMNS (NumPy convention). */
int agf2_b200_eigh(const double *a, int64_t n, double *w, double *v, int on_device);

/* Extended-Fock eigensolver with resident couplings -- the inner loop of auxgf/agf2/fock.py:119-143 and
 * auxgf/agf2/chempot.py:11-49 diagonalises  [[F, v], [v^T, diag(e - shift)]]  hundreds of times per AGF2 iteration
 * with the same couplings v.  set_couplings uploads v (host, (nphys, naux) row-major) once; every solve sends only
 * F (nphys^2) and e (naux), assembles the matrix on the device, runs cuSOLVER dsyevd in a reused workspace and returns
 * the eigenvalues w (nphys + naux, ascending) and the PHYSICAL rows of the eigenvectors v_phys ((nphys, nphys + naux)
 * row-major) -- all the Fock loop, the density matrix and the chemical-potential gradient use.  solve_async / wait:
 * solves on different handles (alpha / beta) overlap on the device. */
typedef struct agf2_b200_fock agf2_b200_fock;
int agf2_b200_fock_create(agf2_b200_fock **out);
int agf2_b200_fock_destroy(agf2_b200_fock *h);
int agf2_b200_fock_set_couplings(agf2_b200_fock *h, const double *v, int64_t nphys, int64_t naux);
int agf2_b200_fock_solve(agf2_b200_fock *h, const double *fock, const double *e, double shift, double *w,
                         double *v_phys);
int agf2_b200_fock_solve_async(agf2_b200_fock *h, const double *fock, const double *e, double shift, double *w,
                               double *v_phys);
int agf2_b200_fock_wait(agf2_b200_fock *h);

/* Galitskii-Migdal two-body energy term on the device (auxgf/aux/energy.py:43-62 restated as one GEMM + reduction):
 *     out = sum_{l < nl, k < nk} (sum_x gv[l, x] sv[k, x])^2 / (ge[l] - se[k])
 * gv (nl, nphys) and sv (nk, nphys): host arrays, pole-major (the transposes of the reference's (nphys, naux)
 * couplings), ge (nl), se (nk).  The overlap matrix runs on the library's DMMA pipeline. */
int agf2_b200_energy_2body(const double *gv, const double *ge, int64_t nl, const double *sv, const double *se,
                           int64_t nk, int64_t nphys, double *out);

/* Device-resident DF tensor ("staged once to HBM") and the per-sector DF transform on the device.
 *   agf2_b200_eri_create : eri is the Cholesky/DF tensor L, either tril-packed (naux, nphys(nphys+1)/2)
 *       as OptRAGF2 keeps it (auxgf/agf2/optragf2.py:159-160, packed != 0) or full (naux, nphys, nphys);
 *       host or device memory (on_device).  It is unpacked to (naux, nphys, nphys) in HBM.
 *   agf2_b200_ao2mo_sector : replaces the two OptRAGF2.ao2mo calls of auxgf/agf2/optragf2.py:243-244
 *       (pyscf _ao2mo.nr_e2, 's2' -> 's1'): from v_occ (nphys, nocc) and v_vir (nphys, nvir), row-major,
 *       forms ixq (nocc, nphys, ld_ixq) and qja (naux, nocc, ld_qja) in handle-owned HBM scratch, in
 *       exactly the layout agf2_b200_moments_*_dev consumes (valid until the next call on the handle). */
typedef struct agf2_b200_eri agf2_b200_eri;
int agf2_b200_eri_create(const double *eri, int64_t naux, int64_t nphys, int packed, int on_device,
                         agf2_b200_eri **out);
int agf2_b200_eri_destroy(agf2_b200_eri *h);
int agf2_b200_eri_full_ptr(agf2_b200_eri *h, double **full);
int agf2_b200_ao2mo_sector(agf2_b200_eri *h, const double *v_occ, int64_t nocc, const double *v_vir,
                           int64_t nvir, int on_device, void *stream, double **ixq, int64_t *ld_ixq,
                           double **qja, int64_t *ld_qja);

/* Coulomb / exchange matrices of a density from the resident DF tensor (replaces the blocked host loop
 * of auxgf/agf2/optragf2.py:358-398, optuagf2.py:317-366):  J = sum_Q L_Q tr(L_Q D),  K = sum_Q L_Q D L_Q.
 * rdm1, j, k: (nphys, nphys) row-major, host or device (on_device); blocking call. */
int agf2_b200_get_jk(agf2_b200_eri *h, const double *rdm1, double *j, double *k, int on_device, void *stream);

/* Instrumentation: number of kernels launched by this library since the last reset, and the
 * device time (ms, CUDA events) of the last moments call's two kernels summed over launches. */
int64_t agf2_b200_launch_count(int reset);
int agf2_b200_last_kernel_ms(double *k1_form_ms, double *k2_moment_ms);

/* Device time (ms) of the last moments call by stage: [0] k1_form, [1] k2_moment of the pair loop, [2] the
 * M0/M1 build and [3] the T = ixq M DGEMMs + moment accumulation of the Coulomb factorisation. */
int agf2_b200_last_stage_ms(double *ms4);

/* Plan of the last moments call: [0] Coulomb factorisation used, work items of this part by kind -- [1] PAIR (i>j: two
 * (xi|ja) blocks each), [2] DIAG, [3] OS (one block each), [4] ASYM (j outside the slice: two blocks) -- and
 * [5] the number of k1_form launches (= chunks). */
int agf2_b200_last_plan(int64_t *info6);

/* The work plan of a moments call, computed on the host without touching the GPU (same planner code the launches use):
 * cover (optional, nocc x (nocc [+ nocc_b]) int64, accumulated into): number of 8-column blocks of every (xi|ja) block
 * (i, j) -- same spin in columns [0, nocc), other spin after -- that this part's k1_form tiles form, summed over
 * x-blocks; stats8 = { Coulomb factorisation used, chunks, CTA tiles, thin tiles, widest chunk (columns), workspace
 * columns, most items in a chunk, 8-column blocks of one complete (xi|ja) block }.  uhf != 0: the UHF channel
 * convention of agf2_b200_moments_uhf_dev (nocc/nvir = same spin, nocc_b/nvir_b = other spin). */
int agf2_b200_plan_stats(int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux, int64_t nocc_b, int64_t nvir_b,
                         int64_t istart, int64_t iend, double os_factor, double ss_factor, int uhf, int64_t part,
                         int64_t nparts, int64_t *cover, int64_t *stats8);

/* Coulomb factorisation (default on): every Coulomb-type term  sum_ija w X_ija X_ija^T (e_i+e_j-e_a)^n  of
 * auxgf/lib/libagf2/optragf2.c:206-290 / optuagf2.c:211-314 is evaluated as  sum_i ixq_i (e_i^n M0 + n M1) ixq_i^T
 * with the naux x naux matrices M0 = sum_ja w q_ja q_ja^T, M1 = sum_ja w (e_j - e_a) q_ja q_ja^T, so that only the
 * exchange-type term  ss sum_{i>j} (X_ij - X_ji)(X_ij - X_ji)^T  goes through the pair loop.  Same result to
 * rounding.  mode 0: off, every term goes through the pair loop (one (xi|ja) block per pole pair and sign);
 * 1: always on; 2 (default): chosen per call from a flop estimate (it pays when nocc*nvir >> naux). */
int agf2_b200_set_coulomb_factorization(int mode);

/* Chunks alternate between two CUDA streams (default on) so that the tail of one kernel is filled by the
 * other chain's CTAs.  Switch it off to time k1_form / k2_moment in isolation (agf2_b200_last_kernel_ms). */
int agf2_b200_set_overlap(int on);

/* Stream-K partial tiles of k2_moment are summed in a fixed order (default on): repeating a build
 * gives bit-identical moments.  Off: partial tiles go straight into the accumulators with
 * red.global.add.f64 (order not fixed, results equal to ~1e-15 relative). */
int agf2_b200_set_deterministic(int on);

/* Debug: route the moment build through the naive one-thread-per-element CUDA kernels (same
 * planner, no tensor cores); used by the tests to separate planner bugs from kernel bugs. */
int agf2_b200_set_simple_kernels(int on);

#ifdef __cplusplus
}
#endif
#endif /* AGF2_B200_H */
