/*
 * damavand_b200.h -- C ABI of the B200-native statevector engine that sits behind damavand's
 * `gpu` and `distributed_gpu` apply methods.
 *
 * This is the boundary the reference's Rust host code (src/qubit_backend) binds through
 * `extern "C"` (reference: /root/reference/src/qubit_backend/circuit.rs:27-51,
 * circuit_gpu.rs:3-24, circuit_distributed_gpu.rs:6-33; implemented by
 * /root/reference/damavand-gpu/rust_communication.cu).  Differences from the reference ABI, all
 * deliberate (SURVEY.md "fact 3", section 8b):
 *   - 64-bit sizes and indices everywhere (the reference passes `int`, which overflows at 2^31);
 *   - an explicit handle instead of process-global state, with create/destroy (no leak on re-init);
 *   - every call returns 0 on success or a non-zero code; dvd_last_error() gives the message
 *     (the reference prints and exit()s inside checkCudaErrors);
 *   - amplitudes are interleaved complex128 on the device; the host-facing read/load calls keep the
 *     reference's split real / imaginary arrays;
 *   - gates are queued and applied in fused passes at the next dvd_flush() or observation;
 *   - one process drives ONE GPU; multi-GPU is one rank per GPU (dvd_create_distributed).
 * The reference's own 14 export names are provided on top of this by damavand_gpu_compat.h.
 *
 * Plain C types only; no CUDA or torch types cross this boundary.
 */
#ifndef DAMAVAND_B200_H
#define DAMAVAND_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dvd_state dvd_state;   /* opaque */

#define DVD_OK 0
#define DVD_ERR_CUDA 1
#define DVD_ERR_ARG 2
#define DVD_ERR_NCCL 3
#define DVD_ERR_INTERNAL 4

#define DVD_NCCL_ID_BYTES 128

/* ---- device queries ------------------------------------------------------------------------- */
/* replaces get_number_of_available_gpus, rust_communication.cu:24-29 */
int dvd_device_count(void);
/* replaces get_memory_for_gpu (MiB), rust_communication.cu:31-37; <0 on error */
double dvd_device_mem_mib(int device);
/* replaces peer_access_allowed, rust_communication.cu:39-47 */
int dvd_peer_access_allowed(int src_device, int dst_device);
/* message of the last failing call on this thread ("" if none) */
const char* dvd_last_error(void);

/* ---- state lifecycle -------------------------------------------------------------------------- */
/* replaces init_quantum_state, rust_communication.cu:143-198: allocate 2^n_qubits complex128
 * amplitudes on `device` and set |0...0>. */
int dvd_create(int n_qubits, int device, dvd_state** out);
/* One rank per GPU.  Rank r owns the contiguous chunk [r*2^n/world, (r+1)*2^n/world) -- the
 * reference's layout (circuit.rs:135-136,187-188).  world must be a power of two.
 * nccl_id: DVD_NCCL_ID_BYTES bytes obtained from dvd_nccl_unique_id() on rank 0 and broadcast by
 * the host (mpi4py, torch.distributed, ...).  Collective over all ranks. */
int dvd_create_distributed(int n_qubits, int device, int rank, int world, const void* nccl_id,
                           dvd_state** out);
int dvd_nccl_unique_id(void* out_id /* DVD_NCCL_ID_BYTES */);
int dvd_destroy(dvd_state* s);
/* reset to |0...0> (Circuit::reset_amplitudes, circuit.rs:262-302); drops queued gates */
int dvd_reset_zero_state(dvd_state* s);

/* ---- gate application ------------------------------------------------------------------------ */
/* replaces apply_one_qubit_gate_gpu_local / _distributed, rust_communication.cu:339-382, with the
 * reference's exact calling convention for the matrix: row-major 2x2 split into real and imaginary
 * parts, control = -1 when the gate is not controlled (circuit_gpu.rs:31-60).  The gate is queued. */
int dvd_apply_gate(dvd_state* s, const double m_re[4], const double m_im[4], int control, int target);
typedef struct {
    int32_t target;
    int32_t control;      /* -1 = none */
    double m[8];          /* m00.re m00.im m01.re m01.im m10.re m10.im m11.re m11.im */
} dvd_gate;
/* queue a whole circuit in one call */
int dvd_apply_circuit(dvd_state* s, const dvd_gate* gates, int64_t n_gates);
/* plan + launch everything queued (asynchronous on the state's stream) */
int dvd_flush(dvd_state* s);
/* wait for the device */
int dvd_synchronize(dvd_state* s);

/* ---- observation (each flushes first) ------------------------------------------------------- */
/* replaces measure_on_gpu, rust_communication.cu:330-337: out[i] = |amp[first+i]|^2 of the LOCAL chunk */
int dvd_probabilities(dvd_state* s, double* out, int64_t first, int64_t count);
/* sum of |amp|^2 over the whole (distributed) state, pairwise-tree order; allreduced */
int dvd_norm(dvd_state* s, double* out);
/* Sampling (Circuit::sample -> sample_local / sample_distributed, circuit.rs:434-485,
 * circuit_distributed.rs:42-129) with INJECTED uniforms in [0,1).
 *   world == 1: uniforms has `shots` entries; out[s] = smallest k with prefix(k) >= u[s]*total.
 *   world  > 1: uniforms has 2*shots entries: [0,shots) pick the rank from the per-rank totals,
 *               [shots,2*shots) pick the index inside that rank; out[s] is the global index and is
 *               identical on every rank.
 * Prefix sums use the pairwise-tree order documented in DESIGN.md. */
int dvd_sample(dvd_state* s, const double* uniforms, int64_t shots, uint64_t* out);
/* Circuit::extract_expectation_values, circuit.rs:494-513: out[s*n_obs+o] = +1 / -1 from bit
 * qubits[o] of samples[s].  Runs on s's device. */
int dvd_extract_expectation_values(dvd_state* s, const uint64_t* samples, int64_t shots,
                                   const int32_t* qubits, int32_t n_obs, double* out);
/* Summation order of the sampler's cumulative probabilities.  DVD_SAMPLER_TREE (default): fixed pairwise tree, any
 * size.  DVD_SAMPLER_SEQUENTIAL: the reference's strict left-to-right order (utils.rs:270-274) -- the indices are the
 * reference's for EVERY draw, at the price of one sequential walk over the chunk per state (<= 30 local qubits). */
#define DVD_SAMPLER_TREE 0
#define DVD_SAMPLER_SEQUENTIAL 1
int dvd_set_sampler(dvd_state* s, int order);
/* exact <Z_q> for q in [0, n_qubits) (extension; allreduced) */
int dvd_expectation_z(dvd_state* s, double* out_per_qubit);
/* replaces retrieve_amplitudes_on_host, rust_communication.cu:450-482: LOCAL chunk, split arrays */
int dvd_read_state(dvd_state* s, double* re, double* im, int64_t first, int64_t count);
/* replaces load_amplitudes_local_on_device, rust_communication.cu:384-398 (done properly) */
int dvd_load_state(dvd_state* s, const double* re, const double* im, int64_t first, int64_t count);
/* |<a|b>|^2 (circuit_metrics.rs:12-92); both states must have the same shape / communicator */
int dvd_fidelity(dvd_state* a, dvd_state* b, double* out);
/* copy src's amplitudes into dst (same shape) */
int dvd_copy_state(dvd_state* dst, dvd_state* src);
/* a second resident state with a copy of src's amplitudes (same shape, device, rank; distributed: src's communicator,
 * so destroy the snapshot before src).  First operand of dvd_fidelity in circuit.rs:753-769 /
 * circuit_metrics.rs:35-92 (distributed_dot). */
int dvd_snapshot(dvd_state* src, dvd_state** out);

/* ---- introspection --------------------------------------------------------------------------- */
int dvd_num_qubits(const dvd_state* s);
int dvd_num_local_qubits(const dvd_state* s);
int dvd_rank(const dvd_state* s);
int dvd_world(const dvd_state* s);
int dvd_device(const dvd_state* s);

typedef struct {
    int64_t gates_applied;        /* gates executed since creation / dvd_stats_reset */
    int64_t kernel_launches;      /* every kernel this library launched */
    int64_t tile_passes;          /* launches of the fused pass kernel */
    int64_t simple_passes;        /* launches of the one-gate kernel */
    int64_t stage_switches;
    int64_t global_swaps;         /* global<->local qubit swaps (NVLink exchanges) */
    int64_t swap_bytes_sent;      /* per rank */
    double pass_bytes;            /* HBM bytes the gate passes must move: 32 * 2^n_local per pass (16 * 2^n_local for
                                     the first pass after a reset, which does not read) */
    double gate_algorithmic_bytes;/* sum over gates of 32*2^n_local (16*2^n_local if controlled) */
    int64_t plan_cache_hits;      /* flushes that reused the previous plan (identical gate list) */
    int64_t jit_launches;         /* tile passes that ran as a structure-specialised (run-time compiled) kernel */
    int64_t remap_passes;         /* tile passes whose load (or store) carried global<->local swaps (fused remap over NVLink peer memory) */
    double remap_bytes_in;        /* bytes those passes pulled from partner ranks over NVLink, per rank (the same amount is
                                     served to the partners in the other direction) */
    double remap_ms;              /* device time of those passes (CUDA events on the state's stream) */
    double swap_ms;               /* device time of the stand-alone exchanges (in-place peer swap / staged NCCL path) */
    double pass_fp64_instr;       /* planner's estimate of the fp64 instructions (per lane: one DADD / DMUL / DFMA of one thread) the
                                     tile passes executed: the second roofline of gate-dense passes */
    int64_t store_remap_passes;   /* of remap_passes: passes whose STORE carried the swaps (the layout restore riding on the
                                     last gate pass: remote writes into the partners' second chunk) */
    double store_remap_ms;        /* of remap_ms: device time of those */
} dvd_stats;
int dvd_get_stats(const dvd_state* s, dvd_stats* out);
int dvd_stats_reset(dvd_state* s);
/* CUDA-event timer on the state's stream: begin records, end records + synchronises */
int dvd_timer_begin(dvd_state* s);
int dvd_timer_end(dvd_state* s, double* elapsed_ms);
/* 0 = fused tile passes (default), 1 = one kernel per gate (debug / baseline) */
int dvd_set_unfused(dvd_state* s, int unfused);
/* Structure-specialised pass kernels, compiled at run time with NVRTC (csrc/jit.h): 0 = off, 1 = compile in the
 * background and switch over when ready, 2 = compile on first use.  Without libnvrtc / libcuda the interpreter
 * kernels keep running.  dvd_jit_wait blocks until the background queue is empty. */
int dvd_set_jit(dvd_state* s, int mode);
int dvd_jit_wait(dvd_state* s);
int dvd_jit_info(dvd_state* s, int64_t* compiled, int64_t* failed, int64_t* pending, double* compile_seconds,
                 char* last_error, int64_t cap);
/* Kernel forms of the specialised kernels (csrc/jit.h: 0 = one tile per CTA at 2 CTAs/SM, 1 = the same at 3 CTAs/SM,
 * 2 = two-group persistent "ring" form).  The form is chosen per pass structure by timing the candidates on the first
 * dense launches.  out[0] = structures still being measured, out[1..3] = structures per chosen form,
 * out[4..6] = launches per form since the library was loaded. */
int dvd_jit_forms(int64_t out[7]);

/* ---- planner inspection (host only, no GPU needed) ----------------------------------------- */
/* Runs the pass planner on a gate list for a state of n_total qubits with n_local local qubits
 * (fuse != 0: after the diagonal-run fusion pre-pass) and writes a flat int32 description:
 *   [n_passes, then per pass: tile_q[12], n_switches, n_ops, then per op: gate_idx, opcode, group, table, regmask]
 * Returns the number of int32 written, or -(needed) if cap is too small, or INT64_MIN on error. */
int64_t dvd_plan_debug(int n_total, int n_local, const dvd_gate* gates, int64_t n_gates, int fuse,
                       int32_t* out, int64_t cap);
/* CUDA source of the structure-specialised kernel the engine would compile for pass `pass_index` of the same plan
 * (NUL-terminated text).  Returns its length, 0 if there is no such pass, -(needed) if cap is too small. */
int64_t dvd_jit_debug_source(int n_total, int n_local, const dvd_gate* gates, int64_t n_gates, int pass_index,
                             int form /* csrc/jit.h JitForm: 0, 1 or 2; + 16 = the store-side-remap variant */, char* out, int64_t cap);
/* NVRTC-compiles such a source for sm_100a (no GPU needed): cubin size, or -1 with the log in dvd_last_error(). */
int64_t dvd_jit_debug_compile(const char* source);
/* Runs the distributed planner: perm_io[logical] = physical (in/out).  Output:
 *   [n_steps, then per step: kind (0 local gates, 1 swap), a, b, n_gates, then per gate: gate_idx, target, control]
 * For swaps a = global physical qubit, b = local physical qubit, n_gates = 0. */
int64_t dvd_plan_distributed_debug(int n_total, int n_local, const dvd_gate* gates, int64_t n_gates,
                                   int32_t* perm_io, int restore_identity, int32_t* out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* DAMAVAND_B200_H */
