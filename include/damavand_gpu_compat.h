/*
 * damavand_gpu_compat.h -- the reference's own 14 export names and `int` signatures
 * (/root/reference/damavand-gpu/rust_communication.cu), implemented on top of damavand_b200.h so
 * that the unmodified `extern "C"` blocks of the reference's Rust host
 * (src/qubit_backend/circuit.rs:27-51, circuit_gpu.rs:3-24, circuit_distributed_gpu.rs:6-33) link
 * against libdamavand_b200.so.  Limits inherited from the `int` ABI: at most 2^31-1 amplitudes per
 * process.  One process drives one GPU, so get_number_of_available_gpus() reports at most 1 and
 * multi-GPU runs use one rank per GPU (call dvd_compat_set_distributed before init_quantum_state).
 * Like the reference (checkCudaErrors), a failing call prints the error and exit()s.
 */
#ifndef DAMAVAND_GPU_COMPAT_H
#define DAMAVAND_GPU_COMPAT_H

#include "damavand_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

int get_number_of_available_gpus(void);                         /* rust_communication.cu:24 */
double get_memory_for_gpu(int local_gpu_rank);                  /* :31 */
int peer_access_allowed(int source_gpu_id, int target_gpu_id);  /* :39 */
void print_timers(void);                                        /* :49 */
void exchange_amplitudes_between_gpus(int current_gpu_rank, int partner_gpu_rank,
                                      int num_amplitudes_per_gpu);          /* :106 (no-op: swaps are planned internally) */
void init_quantum_state(int num_amplitudes_per_gpu, int num_gpus_per_node_required,
                        int is_first_node);                                   /* :143 */
void sequential_measure_on_gpu(int num_amplitudes_per_gpu, double* probabilities);   /* :200 */
void concurrent_measure_on_gpu(int num_amplitudes_per_gpu, double* probabilities);   /* :254 */
void measure_on_gpu(int num_amplitudes_per_gpu, double* probabilities);              /* :330 */
void apply_one_qubit_gate_gpu_local(double* gate_matrix_real, double* gate_matrix_imaginary,
                                    int num_qubits, int num_amplitudes_per_gpu, int control_qubit,
                                    int target_qubit);                        /* :339 */
void apply_one_qubit_gate_gpu_distributed(double* gate_matrix_real, double* gate_matrix_imaginary,
                                          int num_qubits, int num_amplitudes_per_gpu,
                                          int control_qubit, int target_qubit);   /* :361 */
void load_amplitudes_local_on_device(int num_amplitudes_per_gpu, double* local_amplitudes_real,
                                     double* local_amplitudes_imaginary);     /* :385 */
void split_amplitudes_between_gpus(int num_amplitudes_per_gpu, double* local_amplitudes_real,
                                   double* local_amplitudes_imaginary, double* partner_amplitudes_real,
                                   double* partner_amplitudes_imaginary);     /* :400 */
void retrieve_amplitudes_on_host(int num_amplitudes_per_gpu, double* local_amplitudes_real,
                                 double* local_amplitudes_imaginary);         /* :450 */

/* Not in the reference: tells the compat layer that this process is rank `rank` of `world`
 * (one GPU each) before init_quantum_state; nccl_id as in dvd_create_distributed. */
void dvd_compat_set_distributed(int rank, int world, int device, const void* nccl_id);
/* The handle behind the global state (NULL before init_quantum_state). */
dvd_state* dvd_compat_state(void);

#ifdef __cplusplus
}
#endif
#endif
