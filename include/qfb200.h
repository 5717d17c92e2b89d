/*
 * qfb200.h -- C ABI of libqfb200.so, the B200 (sm_100a) gate-application engine that sits
 * behind QuantumFlow's backend contract.
 *
 * Every entry point replaces one piece of arithmetic that the reference performs in Python/numpy.
 * Citations are relative to the reference tree (rigetti/quantumflow):
 *
 *   bk.tensormul            quantumflow/backend/numpybk.py:159-214   -> qfb_apply_dense / _diag / qfb_run_plan
 *   bk.inner                quantumflow/backend/numpybk.py:125-128   -> qfb_vdot
 *   bk.outer (numpy.outer)  quantumflow/backend/numpybk.py:17-21     -> qfb_outer
 *   bk.productdiag          quantumflow/backend/numpybk.py:150-156   -> qfb_density_diag
 *   bk.trace on [2^N,2^N]   quantumflow/qubits.py:185-197            -> qfb_density_trace
 *   bk.transpose (permute)  quantumflow/qubits.py:145-161            -> qfb_permute_bits
 *   State.norm              quantumflow/qubits.py:180-182            -> qfb_norm2
 *   State.normalize         quantumflow/states.py:108-111            -> qfb_scale / qfb_scale_rsqrt_dev
 *   State.probabilities     quantumflow/states.py:113-119            -> qfb_probs
 *   State.expectation       quantumflow/states.py:131-147            -> qfb_expect_diag
 *   Measure.run (P0/P1)     quantumflow/stdops.py:53-65              -> qfb_marginal + qfb_collapse
 *   Circuit.run/evolve loop quantumflow/circuits.py:87-109           -> qfb_run_plan (tiled multi-gate sweeps)
 *   autograd of tensormul   (TF in the reference, tensorflowbk.py:134-152) -> qfb_gate_grad
 *
 * Conventions
 *   - All state pointers are DEVICE pointers to complex128 (interleaved re,im doubles), C-order flat vectors
 *     of 2^nbits amplitudes. Tensor axis i of the reference's [2]*n tensor is flat-index bit (n-1-i).
 *   - `bits[]` are flat-index bit positions, gate qubit 0 first (gate qubit 0 is the MSB of the matrix index,
 *     quantumflow/qubits.py:70-80).
 *   - Small operators (`mat`, `diag`) are HOST pointers to row-major complex128; the launcher copies them into
 *     the kernel parameter block, so there is no hidden device allocation and no sync.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream). All calls are asynchronous on that
 *     stream unless stated; scalar results are written to DEVICE memory (`out_dev`).
 *   - dst == src means in place (race free: each thread owns a closed group of 2^k addresses).
 *   - Return value: QFB_OK (0) or an error code; qfb_last_error() returns a thread-local message.
 *   - `index_hi`: value of the index bits above `nbits` (the rank of a sharded state; 0 when not sharded).
 *     Diagonal entries and control masks may refer to bit positions >= nbits; they are resolved from index_hi.
 */
#ifndef QFB200_H
#define QFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QFB_OK 0
#define QFB_ERR_ARG 1
#define QFB_ERR_CUDA 2
#define QFB_ERR_UNSUPPORTED 3

#define QFB_MAX_DENSE_K 12   /* largest k accepted by qfb_apply_dense (generic path above k=4) */
#define QFB_MAX_DIAG_K 12
#define QFB_MAX_CTRL 8

/* ---- library ---- */
int qfb_version(void);
const char *qfb_last_error(void);
/* sm_count, compute capability, opt-in shared memory per block, total device memory of `device` */
int qfb_device_props(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin,
                     size_t *total_mem);

/* ---- gate application (bk.tensormul) ---- */
/* out[base|off[r]] = sum_c mat[r][c] * in[base|off[c]] for every group; optional controls: groups whose
 * control bits are not all 1 are copied (dst != src) or left untouched (dst == src). Control bit positions
 * >= nbits are resolved from index_hi. */
int qfb_apply_dense(void *dst, const void *src, int nbits, const double *mat_host, int k, const int *bits,
                    int nctrl, const int *ctrl_bits, uint64_t index_hi, void *stream);
/* out[i] = diag[sel(i)] * in[i], sel gathers bits[] (gate qubit 0 = MSB); positions >= nbits use index_hi */
int qfb_apply_diag(void *dst, const void *src, int nbits, const double *diag_host, int k, const int *bits,
                   uint64_t index_hi, void *stream);

/* ---- tiled multi-gate executor (Circuit.run / Circuit.evolve) ---- */
/* `plan_host` is the binary plan produced by the host planner (layout in quantumflow_b200/csrc/qfb_plan.h).
 * Executes every sweep of the plan in place on `state`. */
int qfb_run_plan(void *state, int nbits, uint64_t index_hi, const void *plan_host, size_t plan_bytes,
                 void *stream);
/* Structural validation of a plan (pure host code, no CUDA call): sizes, bit partitions, op records. */
int qfb_plan_validate(const void *plan_host, size_t plan_bytes);
/* Upload a plan once and replay it (plans are immutable); handle is freed with qfb_plan_destroy. */
int qfb_plan_upload(const void *plan_host, size_t plan_bytes, void **handle_out, void *stream);
int qfb_plan_launch(void *handle, void *state, int nbits, uint64_t index_hi, void *stream);
/* Sweeps [first_sweep, first_sweep + nsweeps) of an uploaded plan, optionally on a slice of the state: the amplitudes
 * whose index bits `fix_mask` equal `fix_value` (bits outside the tile of every launched sweep; needs the sweep-
 * specialised kernels, whose slice variants are built on first use); ctas_per_sm bounds the resident CTAs per SM of a
 * slice launch (0 = all that fit, -1 = one fewer: room for the exchange kernel of another slice). qfb_plan_sweep_info: number of sweeps, the
 * non-tile index bits of one sweep, whether the plan runs on sweep-specialised kernels (any pointer may be NULL). */
int qfb_plan_launch_part(void *handle, void *state, int nbits, uint64_t index_hi, int first_sweep, int nsweeps,
                         uint64_t fix_mask, uint64_t fix_value, int ctas_per_sm, void *stream);
int qfb_plan_sweep_info(void *handle, int sweep, int *nsweeps, uint64_t *nontile_mask, int *specialised);
int qfb_plan_destroy(void *handle);
/* Sweep-specialised kernels (csrc/qfb_jit.cu): qfb_plan_upload emits every sweep of the plan as straight-line PTX,
 * compiles it in-process for sm_100a and loads it (environment: QFB_JIT=0/1, QFB_JIT_MIN_BITS). The two entry
 * points below are pure host code (no GPU, no driver): the PTX text of one sweep (`needed` = bytes incl. the
 * terminator, `ncoef` = coefficients in the constant bank), and a compile check of every sweep of a plan
 * (`log` receives the PTX compiler's register / spill report). */
int qfb_jit_ptx(const void *plan_host, size_t plan_bytes, int sweep, char *buf, size_t cap, size_t *needed,
                size_t *ncoef);
int qfb_jit_check(const void *plan_host, size_t plan_bytes, char *log, size_t cap);
/* Everything a launch of one sweep's kernel consists of, for tools and for the CPU tests that EXECUTE the generated
 * code on a PTX emulator (tests/ptx_emulator.py): the PTX text (fix_mask != 0: the variant for launches over a slice
 * of the state, whose index bits fix_mask come from the kernel parameter p_fix), the coefficients in the order of the
 * module's constant bank qfb_coef, threads per CTA, dynamic shared memory, tiles a CTA works on side by side. Any
 * output pointer may be NULL. Host code only. */
int qfb_jit_source(const void *plan_host, size_t plan_bytes, int sweep, uint64_t fix_mask, char *ptx, size_t ptx_cap,
                   size_t *ptx_needed, double *coef, size_t coef_cap, size_t *ncoef, int *threads,
                   size_t *smem_bytes, int *groups);
/* Whole circuits on small states in ONE launch (csrc/qfb_small.cu; the QAOA gradient step of
 * examples/qaoa_maxcut.py:37-87, BASELINE.json configs[1]): the state (<= 13 qubits) stays in the shared memory of one
 * CTA for all gates; `batch` independent items (states [batch][2^nbits], matrices [batch][mats_stride]) run as one CTA
 * each. gates_dev: ngates records of 4 int32 {k (1|2), index bit of gate qubit 0, of gate qubit 1, offset of the
 * row-major 2^k x 2^k matrix in the item's matrix array (complex elements)}; all pointers are device pointers.
 * qfb_small_circuit_adjoint is the reverse sweep of the adjoint method for UNITARY gates (<= 12 qubits): from the final
 * state and grad_out = dL/dpsi_final it writes grad_in = dL/dpsi_initial and, per gate, grad[r][c] = sum over groups
 * of lambda_out[r] conj(psi_in[c]) (qfb_gate_grad's convention) into grad_mats_dev, laid out like mats_dev
 * ([batch][mats_stride], the gate's entries at its matrix offset); no intermediate state is stored. scratch_dev: qfb_small_circuit_scratch_doubles(batch, ngates) doubles. */
int qfb_small_circuit_run(void *out, const void *in, int nbits, int batch, int ngates, const void *gates_dev,
                          const void *mats_dev, size_t mats_stride, void *stream);
int qfb_small_circuit_adjoint(const void *psi_final, const void *grad_out, int nbits, int batch, int ngates,
                              const void *gates_dev, const void *mats_dev, size_t mats_stride, void *grad_mats_dev,
                              void *grad_in_dev, void *scratch_dev, void *stream);
size_t qfb_small_circuit_scratch_doubles(int batch, int ngates);
/* number of kernel launches performed by this library in this process (bench.py's gpu_launches) */
uint64_t qfb_launch_count(void);

/* ---- reductions / read-out ---- */
int qfb_vdot(const void *a, const void *b, uint64_t n, double *out_dev2, void *stream);   /* sum conj(a)*b */
int qfb_norm2(const void *a, uint64_t n, double *out_dev, void *stream);                  /* sum |a|^2 */
int qfb_probs(const void *a, uint64_t n, double *out_dev, void *stream);                  /* out[i]=|a[i]|^2 */
int qfb_expect_diag(const void *a, const double *diag_dev, uint64_t n, double *out_dev, void *stream);
/* out_dev2 = { sum_{bit=0} |a|^2 , sum_{bit=1} |a|^2 } */
int qfb_marginal(const void *a, int nbits, int bit, double *out_dev2, void *stream);
/* dst[i] = (bit(i)==value) ? scale*src[i] : 0 */
int qfb_collapse(void *dst, const void *src, int nbits, int bit, int value, double scale, void *stream);
int qfb_scale(void *dst, const void *src, uint64_t n, double scale_re, double scale_im, void *stream);
/* dst = src * rsqrt(*norm2_dev)  (State.normalize without a host round trip) */
int qfb_scale_rsqrt_dev(void *dst, const void *src, uint64_t n, const double *norm2_dev, void *stream);
/* dst = src / (complex at cdiv_dev2)  (Density.normalize: divide by trace) */
int qfb_scale_cdiv_dev(void *dst, const void *src, uint64_t n, const double *cdiv_dev2, void *stream);
/* dst = alpha*a + beta*b (complex scalars given as re,im) ; b may be NULL when beta == 0 */
int qfb_axpby(void *dst, const void *a, double alpha_re, double alpha_im, const void *b, double beta_re,
              double beta_im, uint64_t n, void *stream);
/* dst[i*nb + j] = a[i] * (conj_b ? conj(b[j]) : b[j]) */
int qfb_outer(void *dst, const void *a, uint64_t na, const void *b, uint64_t nb, int conj_b, void *stream);
int qfb_conj(void *dst, const void *src, uint64_t n, void *stream);
/* rho is a [2^nq, 2^nq] row-major matrix */
int qfb_density_diag(const void *rho, int nq, void *out_dev_c128, void *stream);
int qfb_density_trace(const void *rho, int nq, double *out_dev2, void *stream);
/* Partial trace of a [2]*nbits tensor (replaces the np.einsum with repeated subscripts of
 * quantumflow/qubits.py:201-227): dst has 2^nkeep elements, dst index bit b <- src index bit keep_pos[b];
 * trace_masks[t] = OR of the src index bits that hold traced qubit t in every rank block (ket bit | bra bit
 * of a density); dst[j] = sum over the 2^ntr settings in which all copies of each traced qubit agree.
 * keep_pos and trace_masks are host arrays and together cover every src index bit exactly once. */
int qfb_partial_trace(void *dst, const void *src, int nbits, int nkeep, const int *keep_pos, int ntr,
                      const uint64_t *trace_masks, void *stream);
/* dst index bit j <- src index bit perm[j]  (generalised transpose of a [2]*nbits tensor) */
int qfb_permute_bits(void *dst, const void *src, int nbits, const int *perm, int conj, void *stream);
/* cumulative search used by sampling: for each u[j] in [0,total) find smallest i with cdf(i) > u[j];
 * block-hierarchical, deterministic. probs_dev: float64[n]; u_host: nu uniforms already scaled to [0,1);
 * out_idx_host: nu indices. Synchronises the stream. */
int qfb_sample_search(const double *probs_dev, uint64_t n, const double *u_host, int nu, uint64_t *out_idx_host,
                      void *stream);

/* ---- host-side planner support (no GPU work; quantumflow_b200/planner.py) ---- */
/* Operators of a circuit in program order as parallel arrays: mix[i] / diag[i] = masks of the index bits operator
 * i mixes / only reads, cost[i] = planner work units, bytes[i] = upper bound of its records in a sweep.
 * count_executed: how many operators that touch a bit a sweep over the tile `tmask` executes (fmask = bits no
 * operator may mix: the rank bits of a sharded state; max_cost / room = work and size caps of one sweep).
 * refine_tile: local search over the tile -- exchange one tile bit outside `keep` for one bit outside the tile and
 * `fmask` while that raises the count (best exchange of a pass, up to `passes` passes). */
int qfb_plan_count_executed(const uint64_t *mix, const uint64_t *diag, const double *cost, const uint32_t *bytes,
                            int nops, uint64_t tmask, uint64_t fmask, double max_cost, int64_t room, int *count_out);
int qfb_plan_refine_tile(const uint64_t *mix, const uint64_t *diag, const double *cost, const uint32_t *bytes,
                         int nops, int nbits, uint64_t tmask, uint64_t fmask, uint64_t keep, double max_cost,
                         int64_t room, int passes, uint64_t *tmask_out, int *count_out);
/* refine_tile_lookahead: the same exchanges scored by the operators this sweep executes PLUS the operators the
 * best next sweep (greedy tile of low_bits + first-come bits up to tile_bits, then refine_tile) executes on
 * what is left; up to `lookahead_passes` passes, inner searches with `passes`. */
int qfb_plan_refine_tile_lookahead(const uint64_t *mix, const uint64_t *diag, const double *cost,
                                   const uint32_t *bytes, int nops, int nbits, int low_bits, int tile_bits,
                                   uint64_t tmask, uint64_t fmask, uint64_t keep, double max_cost, int64_t room,
                                   int passes, int lookahead_passes, uint64_t *tmask_out, int *score_out);

/* Threads the tile searches use for the candidates of one pass (results do not depend on it); 0 = back to the
 * default: QFB_PLAN_THREADS, else min(4, hardware threads / LOCAL_WORLD_SIZE). */
int qfb_plan_set_threads(int nthreads);
/* split_rounds: the greedy split of ONE sweep's operators into rounds of `reg_bits` register bits (the planner's
 * search over randomised variants calls it a few hundred times per sweep). posmask[i] = tile positions of the bits
 * operator i mixes (0xffffffff for a phase term), low_bits = tile positions that stay on the lanes in the rounds
 * that touch HBM, rnd[0..nrnd) = the random stream of the variant (NULL: plain greedy), backward != 0: the same on
 * the reversed list (latest-possible rounds). round_of / regs_of_round (bit masks of tile positions, room for
 * max_rounds) may be NULL; *tail_out = cost of the last round's operators; *consumed_out = random numbers used, -1
 * when the stream ran out (nothing else is valid then). */
int qfb_plan_split_rounds(const uint64_t *mix, const uint64_t *diag, const uint32_t *posmask, const double *cost,
                          int nops, int reg_bits, int low_bits, const double *rnd, int nrnd, double p_new,
                          int backward, int *round_of, uint32_t *regs_of_round, int max_rounds, int *nrounds_out,
                          double *tail_out, int *consumed_out);

/* ---- sharded states: the exchange half of a qubit remap (csrc/qfb_remap.cu; the reference has no distributed
 * path) ---- */
/* For i < npairs: the `nelems[i]` complex128 amplitudes at local_blocks[i] (this GPU) and at remote_blocks[i] (a
 * peer GPU's shard, mapped into this process, e.g. through CUDA IPC) trade places, in one kernel over NVLink.
 * Both ranks of a pair call it for disjoint halves of the pair's blocks. Stream ordered; the caller orders it
 * against the peers' work (a barrier before and after). Host arrays. */
int qfb_remap_swap(int npairs, void *const *local_blocks, void *const *remote_blocks, const uint64_t *nelems,
                   void *stream);
/* The same for a SLICE of every run: only the amplitudes whose offset inside the run has the bits selpos[0..nsel)
 * (ascending, nsel <= 4) equal to `selval` trade places; `ctas_per_sm` (0 = default) bounds the kernel's share of
 * every SM so that it can run beside a sweep on another stream. Sharded states pipeline a remap slice by slice:
 * last sweep of the stage on slice s+1 | exchange of slice s | first sweep of the next stage on slice s-1. */
int qfb_remap_swap_slice(int npairs, void *const *local_blocks, void *const *remote_blocks, const uint64_t *nelems,
                         int nsel, const int *selpos, uint64_t selval, int ctas_per_sm, void *stream);
/* Stream-ordered barrier across the GPUs of one box through peer memory (no host involvement, no NCCL kernel that
 * would need a free SM): flags_of_ranks[r] = rank r's array of `world` uint32 (flags_of_ranks[rank] = flags_local),
 * epochs must increase call by call; *error_dev (uint32) is set to the epoch if a peer does not arrive within two minutes. */
int qfb_peer_barrier(void *flags_local, void *const *flags_of_ranks, int world, int rank, uint32_t epoch,
                     void *error_dev, void *stream);

/* ---- batched stochastic trajectories (Kraus.run / UnitaryMixture.run, quantumflow/channels.py:70-77, 119-125, for
 * 2^nbatch_bits pure states of nstate qubits in one buffer: trajectory index = the top index bits) ---- */
/* out_dev[4 t .. 4 t + 3] = (sum |x|^2, sum |y|^2, Re sum x conj(y), Im sum x conj(y)) over the amplitude pairs
 * (x: bit = 0, y: bit = 1) of trajectory t: the 1-qubit reduced density from which the host gets every branch
 * probability w_k tr(K_k rho K_k^dagger). workspace_dev: qfb_batch_rho1_workspace() bytes. Deterministic. */
int qfb_batch_rho1(const void *state, int nstate, int nbatch_bits, int bit, double *out_dev, void *workspace_dev,
                   size_t workspace_bytes, void *stream);
size_t qfb_batch_rho1_workspace(int nstate, int nbatch_bits);
/* psi_t <- M_t psi_t in place, M_t = mats_dev[8 t .. 8 t + 7] (row-major 2x2 complex128, DEVICE memory, 16-byte
 * aligned): the Kraus branch drawn for trajectory t, already divided by its norm. */
int qfb_batch_apply1(void *state, int nstate, int nbatch_bits, int bit, const double *mats_dev, void *stream);

/* ---- autograd bridge ---- */
/* grad_mat[r][c] = sum_groups g[base|off[r]] * conj(psi[base|off[c]])   (k <= 3), written to out_dev (4^k c128) */
int qfb_gate_grad(const void *g, const void *psi, int nbits, int k, const int *bits, void *out_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QFB200_H */
