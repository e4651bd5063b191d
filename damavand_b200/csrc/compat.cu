// compat.cu -- the reference's export names (damavand_gpu_compat.h) over one process-global dvd_state.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/damavand_gpu_compat.h"

static dvd_state* g_state = nullptr;
static int g_rank = 0, g_world = 1, g_device = 0;
static unsigned char g_id[DVD_NCCL_ID_BYTES];
static bool g_have_id = false;

static void check(int rc, const char* what) {
    if (rc != DVD_OK) {   // mirrors checkCudaErrors: print + exit
        fprintf(stderr, "damavand-b200: %s failed: %s\n", what, dvd_last_error());
        exit(EXIT_FAILURE);
    }
}
static dvd_state* state(const char* what) {
    if (!g_state) { fprintf(stderr, "damavand-b200: %s called before init_quantum_state\n", what); exit(EXIT_FAILURE); }
    return g_state;
}

extern "C" {

int get_number_of_available_gpus(void) { return dvd_device_count() > 0 ? 1 : 0; }
double get_memory_for_gpu(int gpu) { return dvd_device_mem_mib(g_device + gpu); }
int peer_access_allowed(int a, int b) { return dvd_peer_access_allowed(a, b); }

void print_timers(void) {
    if (!g_state) return;
    dvd_stats st;
    if (dvd_get_stats(g_state, &st) != DVD_OK) return;
    printf("damavand-b200: gates %lld, kernel launches %lld, fused passes %lld, global swaps %lld\n",
           (long long)st.gates_applied, (long long)st.kernel_launches, (long long)st.tile_passes,
           (long long)st.global_swaps);
}

void dvd_compat_set_distributed(int rank, int world, int device, const void* nccl_id) {
    g_rank = rank; g_world = world; g_device = device;
    g_have_id = nccl_id != nullptr;
    if (nccl_id) memcpy(g_id, nccl_id, sizeof g_id);
}
dvd_state* dvd_compat_state(void) { return g_state; }

void init_quantum_state(int num_amplitudes_per_gpu, int num_gpus_per_node_required, int is_first_node) {
    (void)is_first_node;   // rank 0 owns amplitude 0, as in the reference
    if (num_gpus_per_node_required > 1) {
        fprintf(stderr, "damavand-b200: one process drives one GPU; launch one rank per GPU instead of %d GPUs per process\n",
                num_gpus_per_node_required);
        exit(EXIT_FAILURE);
    }
    int n_local = 0;
    while ((1ll << n_local) < (long long)num_amplitudes_per_gpu) ++n_local;
    int g = 0;
    while ((1 << g) < g_world) ++g;
    // The reference calls init_quantum_state from every reset_amplitudes() (circuit.rs:271-301): the same shape again is
    // a reset of the state that already exists (no re-allocation, and above all no second ncclCommInitRank with an id
    // that has been consumed); a different shape replaces it (the reference leaks here).
    if (g_state && dvd_num_qubits(g_state) == n_local + g && dvd_rank(g_state) == g_rank && dvd_world(g_state) == g_world &&
        dvd_device(g_state) == g_device) {
        check(dvd_reset_zero_state(g_state), "init_quantum_state");
        return;
    }
    if (g_state) { dvd_destroy(g_state); g_state = nullptr; }
    if (g_world > 1) check(dvd_create_distributed(n_local + g, g_device, g_rank, g_world, g_have_id ? g_id : nullptr, &g_state), "init_quantum_state");
    else check(dvd_create(n_local, g_device, &g_state), "init_quantum_state");
}

void apply_one_qubit_gate_gpu_local(double* re, double* im, int, int, int control, int target) {
    check(dvd_apply_gate(state("apply_one_qubit_gate_gpu_local"), re, im, control, target), "apply_one_qubit_gate_gpu_local");
}
void apply_one_qubit_gate_gpu_distributed(double* re, double* im, int, int, int control, int target) {
    check(dvd_apply_gate(state("apply_one_qubit_gate_gpu_distributed"), re, im, control, target), "apply_one_qubit_gate_gpu_distributed");
}
void exchange_amplitudes_between_gpus(int, int, int) {}

void measure_on_gpu(int n, double* probs) { check(dvd_probabilities(state("measure_on_gpu"), probs, 0, n), "measure_on_gpu"); }
void sequential_measure_on_gpu(int n, double* probs) { measure_on_gpu(n, probs); }
void concurrent_measure_on_gpu(int n, double* probs) { measure_on_gpu(n, probs); }

void load_amplitudes_local_on_device(int n, double* re, double* im) {
    check(dvd_load_state(state("load_amplitudes_local_on_device"), re, im, 0, n), "load_amplitudes_local_on_device");
}
void split_amplitudes_between_gpus(int n, double* lre, double* lim, double*, double*) {
    check(dvd_load_state(state("split_amplitudes_between_gpus"), lre, lim, 0, n), "split_amplitudes_between_gpus");
}
void retrieve_amplitudes_on_host(int n, double* re, double* im) {
    check(dvd_read_state(state("retrieve_amplitudes_on_host"), re, im, 0, n), "retrieve_amplitudes_on_host");
}

}  // extern "C"
