//! Raw bindings to include/damavand_b200.h.  One declaration per C entry point; the comments give
//! the reference `extern "C"` item each one supersedes (paths relative to the reference repo).
#![allow(non_camel_case_types)]
use libc::{c_char, c_double, c_int, c_void};

#[repr(C)]
pub struct dvd_state {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dvd_gate {
    pub target: i32,
    pub control: i32, // -1 = none (circuit_gpu.rs:36-41)
    pub m: [c_double; 8],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct dvd_stats {
    pub gates_applied: i64,
    pub kernel_launches: i64,
    pub tile_passes: i64,
    pub simple_passes: i64,
    pub stage_switches: i64,
    pub global_swaps: i64,
    pub swap_bytes_sent: i64,
    pub pass_bytes: c_double,
    pub gate_algorithmic_bytes: c_double,
    pub plan_cache_hits: i64,
    pub jit_launches: i64,
    pub remap_passes: i64,
    pub remap_bytes_in: c_double,
    pub remap_ms: c_double,
    pub swap_ms: c_double,
    pub pass_fp64_instr: c_double,
    pub store_remap_passes: i64,
    pub store_remap_ms: c_double,
}

pub const DVD_SAMPLER_TREE: c_int = 0;
pub const DVD_SAMPLER_SEQUENTIAL: c_int = 1;

pub const DVD_NCCL_ID_BYTES: usize = 128;

extern "C" {
    // src/qubit_backend/circuit.rs:30  get_number_of_available_gpus
    pub fn dvd_device_count() -> c_int;
    // circuit.rs:33  get_memory_for_gpu
    pub fn dvd_device_mem_mib(device: c_int) -> c_double;
    pub fn dvd_peer_access_allowed(src: c_int, dst: c_int) -> c_int;
    pub fn dvd_last_error() -> *const c_char;

    // circuit.rs:35-39  init_quantum_state
    pub fn dvd_create(n_qubits: c_int, device: c_int, out: *mut *mut dvd_state) -> c_int;
    pub fn dvd_create_distributed(
        n_qubits: c_int, device: c_int, rank: c_int, world: c_int,
        nccl_id: *const c_void, out: *mut *mut dvd_state,
    ) -> c_int;
    pub fn dvd_nccl_unique_id(out_id: *mut c_void) -> c_int;
    pub fn dvd_destroy(s: *mut dvd_state) -> c_int;
    pub fn dvd_reset_zero_state(s: *mut dvd_state) -> c_int;

    // circuit_gpu.rs:6-22  apply_one_qubit_gate_gpu_local / _distributed
    pub fn dvd_apply_gate(
        s: *mut dvd_state, m_re: *const c_double, m_im: *const c_double, control: c_int, target: c_int,
    ) -> c_int;
    pub fn dvd_apply_circuit(s: *mut dvd_state, gates: *const dvd_gate, n_gates: i64) -> c_int;
    pub fn dvd_flush(s: *mut dvd_state) -> c_int;
    pub fn dvd_synchronize(s: *mut dvd_state) -> c_int;

    // circuit.rs:41-44  measure_on_gpu
    pub fn dvd_probabilities(s: *mut dvd_state, out: *mut c_double, first: i64, count: i64) -> c_int;
    pub fn dvd_norm(s: *mut dvd_state, out: *mut c_double) -> c_int;
    // circuit.rs:434-485, circuit_distributed.rs:42-129 (sampling moves onto the device)
    pub fn dvd_sample(s: *mut dvd_state, uniforms: *const c_double, shots: i64, out: *mut u64) -> c_int;
    // circuit.rs:494-513
    pub fn dvd_extract_expectation_values(
        s: *mut dvd_state, samples: *const u64, shots: i64, qubits: *const i32, n_obs: i32, out: *mut c_double,
    ) -> c_int;
    pub fn dvd_expectation_z(s: *mut dvd_state, out_per_qubit: *mut c_double) -> c_int;
    // circuit.rs:46-50  retrieve_amplitudes_on_host
    pub fn dvd_read_state(s: *mut dvd_state, re: *mut c_double, im: *mut c_double, first: i64, count: i64) -> c_int;
    // circuit_distributed_gpu.rs:8-18 (commented out in the reference)  load_amplitudes_local_on_device
    pub fn dvd_load_state(s: *mut dvd_state, re: *const c_double, im: *const c_double, first: i64, count: i64) -> c_int;
    // circuit_metrics.rs:12-92
    pub fn dvd_fidelity(a: *mut dvd_state, b: *mut dvd_state, out: *mut c_double) -> c_int;
    pub fn dvd_copy_state(dst: *mut dvd_state, src: *mut dvd_state) -> c_int;
    // circuit.rs:753-769: the first state of get_fidelity_between_two_states_with_parameters stays resident
    pub fn dvd_snapshot(src: *mut dvd_state, out: *mut *mut dvd_state) -> c_int;
    // utils.rs:270-274: 1 = the reference's strictly sequential cumulative sums (<= 30 local qubits), 0 = pairwise tree
    pub fn dvd_set_sampler(s: *mut dvd_state, order: c_int) -> c_int;

    pub fn dvd_num_qubits(s: *const dvd_state) -> c_int;
    pub fn dvd_num_local_qubits(s: *const dvd_state) -> c_int;
    pub fn dvd_rank(s: *const dvd_state) -> c_int;
    pub fn dvd_world(s: *const dvd_state) -> c_int;
    pub fn dvd_device(s: *const dvd_state) -> c_int;
    pub fn dvd_get_stats(s: *const dvd_state, out: *mut dvd_stats) -> c_int;
    pub fn dvd_stats_reset(s: *mut dvd_state) -> c_int;
    pub fn dvd_timer_begin(s: *mut dvd_state) -> c_int;
    pub fn dvd_timer_end(s: *mut dvd_state, elapsed_ms: *mut c_double) -> c_int;
    pub fn dvd_set_unfused(s: *mut dvd_state, unfused: c_int) -> c_int;
    // run-time specialised pass kernels (NVRTC): 0 off, 1 background compile, 2 compile on first use
    pub fn dvd_set_jit(s: *mut dvd_state, mode: c_int) -> c_int;
    pub fn dvd_jit_wait(s: *mut dvd_state) -> c_int;
    pub fn dvd_jit_info(
        s: *mut dvd_state, compiled: *mut i64, failed: *mut i64, pending: *mut i64,
        compile_seconds: *mut c_double, last_error: *mut c_char, cap: i64,
    ) -> c_int;
    // out[0] structures still being measured, out[1..3] structures per chosen kernel form, out[4..6] launches per form
    pub fn dvd_jit_forms(out: *mut i64) -> c_int;
}
