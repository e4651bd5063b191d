// engine.cu -- device state object + the C ABI of include/damavand_b200.h.
//
// Replaces /root/reference/damavand-gpu/rust_communication.cu (process-global state, OpenMP loop
// over GPUs, int32 sizes, exit() on error) and quantum_amplitudes.cu (SoA re/im arrays, one
// cudaDeviceSynchronize per gate).  One dvd_state drives one GPU; amplitudes are interleaved
// complex128 in one allocation; gates are queued and executed as fused passes.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/damavand_b200.h"
#include "kernels.h"
#include "nccl_dyn.h"
#include "planner.h"
#include "jit.h"
#include "jit_rt.h"

using namespace dvd;

static thread_local std::string g_last_error;
static NcclApi g_nccl;

static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(DVD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
    } while (0)
#define NC(call)                                                                              \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess)                                                               \
            return fail(DVD_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r__)); \
    } while (0)
#define TRY(call)                     \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != DVD_OK) return rc__; \
    } while (0)

struct dvd_state {
    int n_qubits = 0, n_local = 0, rank = 0, world = 1, device = 0;
    uint64_t n_amps = 0;        // local amplitudes
    uint64_t rank_bits = 0;     // rank << n_local
    cplx* amp = nullptr;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_pack[2] = {nullptr, nullptr}, ev_comm[2] = {nullptr, nullptr};
    std::vector<HostGate> pending;
    std::vector<int> perm;      // logical -> physical (identity between flushes)
    PassParams pass_params;     // staging for the kernel parameter block (copied at launch)
    // phase / offset tables of the queued passes: two staging sets used alternately, so that the host can plan
    // and upload flush f + 1 while the GPU still runs the passes of flush f
    struct TabSet { cplx* d = nullptr; cplx* h = nullptr; size_t cap = 0; cudaEvent_t done = nullptr; bool in_flight = false; };
    TabSet tabs[2];
    uint64_t n_flushes = 0;
    double* d_tree = nullptr; bool tree_valid = false;
    // reference-order sampler (dvd_set_sampler): sequential cumulative sums at the block boundaries
    int sampler = DVD_SAMPLER_TREE;
    double* d_seq_cum = nullptr; bool seq_valid = false;
    double* d_scratch = nullptr; size_t scratch_doubles = 0;
    double* h_stage = nullptr;                      // pinned staging of dvd_read_state / dvd_load_state (lazily allocated)
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    ncclComm_t comm = nullptr;
    bool comm_borrowed = false;   // a snapshot (dvd_snapshot) uses its source's communicator and must not destroy it
    cplx* swap_buf[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t swap_chunk = 0;
    // direct peer path: every other rank's state allocation mapped with CUDA IPC; d_bar backs the stream barriers
    bool peer_swap = false;
    std::vector<cplx*> peer_base;   // [world]: base of rank r's allocation (nullptr for this rank / not mapped)
    double* d_bar = nullptr;
    // Fused remap (world > 1): the allocation holds TWO chunks; a global<->local qubit swap is executed by the LOAD of
    // the next tile pass, which reads buffer `cur` of this rank and of its partners (NVLink peer memory) and writes
    // buffer 1 - cur of this rank.  Every rank flips `cur` at the same points of the same plan.
    cplx* buf[2] = {nullptr, nullptr};
    int cur = 0;
    bool fused_remap = false;
    struct RemapTimer { cudaEvent_t t0 = nullptr, t1 = nullptr; bool is_swap = false, is_store = false; };
    std::vector<RemapTimer> remap_timers;   // events around the fused passes since the last dvd_stats_reset
    size_t remap_timers_used = 0;
    std::unique_ptr<Pass> ident_pass;       // empty pass for swaps with no gate pass to ride on
    cplx* d_ident_tab = nullptr;
    cplx* peer_cur(int r) const { return peer_base[r] + (size_t)cur * n_amps; }
    cplx* peer_other(int r) const { return peer_base[r] + (size_t)(1 - cur) * n_amps; }
    int store_remap = 2;          // DVD_STORE_REMAP (planner.h DistPlan): 0 = swaps only ever ride on loads, 1 = the layout restore
                                  //   rides on the store of the last gate pass, 2 = every swap round rides on a store where it can
                                  //   (default: measured on 2 x B200, random32 300 -> 276 ms per dense forward)
    bool pushed_pending = false;  // a store-side pass has run: its remote writes are complete on every rank only after a barrier
    dvd_stats stats;
    bool unfused = false;
    // Support tracking: after a reset only amplitude 0 is stored; `support` holds the local qubits that a
    // non-diagonal gate has touched since.  Amplitudes with a bit outside `support` set are zero by construction
    // (and their memory is unwritten): tile passes neither read them nor launch the tiles that consist of them,
    // and anything else that looks at the buffer materialises the zeros first.
    uint64_t support = ~0ull;     // all ones = dense (nothing implied)
    bool lazy_zero = true;
    uint64_t zero_mask() const { return ~support & (n_amps - 1); }
    // the last few plans, reused when the same gate list is flushed again from the same kind of state (sampling loops
    // re-run one circuit; a reset + forward and a forward on the dense state it leaves are planned differently):
    // key = the queued gates, bit for bit, plus the mode they were planned in and what was known to be zero
    struct PlanCache {
        std::vector<uint64_t> key;
        std::vector<DistStep> steps;
        std::vector<std::vector<Pass>> plans;
        size_t total_tabs = 0;
        std::vector<std::vector<std::pair<int, int>>> store;   // planner.h DistPlan::store: swap rounds riding on the store of the
                                                               //   last pass of their step
        bool valid = false;
        uint64_t stamp = 0;      // last use (least recently used entry is replaced)
    };
    static constexpr size_t PLAN_CACHE_ENTRIES = 4;
    std::vector<PlanCache> caches;
    uint64_t cache_clock = 0;
    // the planners' portfolio winners per gate-list STRUCTURE (planner.h: PlanChoices): a variational loop flushes the
    // same circuit with new angles every iteration -- the plan cache misses, but the search need not be repeated
    ChoiceMemoTable memos;
    bool plan_cache = true;
    // structure-specialised kernels (jit_rt.h): off / background compile / compile on first use
    int jit_mode = JIT_OFF;
    int jit_min_qubits = 20;
    std::string jit_error;
    PlanOptions opt;
};

static int ensure_scratch(dvd_state* s, size_t doubles) {
    if (s->scratch_doubles >= doubles) return DVD_OK;
    if (s->d_scratch) CU(cudaFree(s->d_scratch));
    s->d_scratch = nullptr; s->scratch_doubles = 0;
    CU(cudaMalloc(&s->d_scratch, doubles * sizeof(double)));
    s->scratch_doubles = doubles;
    return DVD_OK;
}

// Reset to |0..0>: amplitude 0 lives on the first rank (circuit.rs:168-170, kernels.cu:62-81).  With support
// tracking only that one amplitude is written (16 B instead of 16 B per amplitude).
static int set_zero_state(dvd_state* s) {
    s->tree_valid = false; s->seq_valid = false;
    if (s->lazy_zero && !s->unfused && s->n_local >= TILE_BITS) {
        s->support = ~(s->n_amps - 1);    // no local qubit touched yet; rank-index bits always count as touched
    } else {
        CU(cudaMemsetAsync(s->amp, 0, s->n_amps * sizeof(cplx), s->stream));
        s->support = ~0ull;
    }
    CU(launch_set_basis_state(s->amp, 0, s->rank == 0 ? 1.0 : 0.0, s->stream));
    s->stats.kernel_launches++;
    return DVD_OK;
}
// Store the zeros that support tracking implied (no-op once every local qubit has been touched).
static int materialize(dvd_state* s) {
    const uint64_t zm = s->zero_mask();
    if (zm == 0) return DVD_OK;
    CU(launch_zero_outside_support(s->amp, s->n_local, zm, s->stream));
    s->stats.kernel_launches++;
    s->support = ~0ull;
    return DVD_OK;
}

// An event pair from the pool (dvd_get_stats turns the pairs into remap_ms / swap_ms).
static int timer_acquire(dvd_state* s, bool is_swap, dvd_state::RemapTimer** out) {
    if (s->remap_timers_used == s->remap_timers.size()) {
        if (s->remap_timers.size() >= 4096) { *out = nullptr; return DVD_OK; }   // nobody reset the stats: stop timing
        dvd_state::RemapTimer t;
        CU(cudaEventCreate(&t.t0)); CU(cudaEventCreate(&t.t1));
        s->remap_timers.push_back(t);
    }
    *out = &s->remap_timers[s->remap_timers_used++];
    (*out)->is_swap = is_swap;
    (*out)->is_store = false;
    return DVD_OK;
}

// Stream-ordered barrier over all ranks: a one-element allreduce completes on a rank only after every
// rank's stream has reached it, i.e. after all earlier kernels on every rank's stream have finished.
static int stream_barrier(dvd_state* s) {
    NC(g_nccl.AllReduce(s->d_bar, s->d_bar + 1, 1, ncclDouble, ncclSum, s->comm, s->stream));
    return DVD_OK;
}

// Map the partner ranks' state buffers into this process (CUDA IPC, one node) so that global<->local
// qubit swaps can run as one kernel over NVLink peer memory.  Every rank must take the same path, so
// the outcome is agreed on with an allreduce(min); DVD_SWAP=nccl forces the staged NCCL send/recv path
// (the only one that works across nodes).
static int map_peers(dvd_state* s) {
    CU(cudaMalloc(&s->d_bar, 4 * sizeof(double)));
    CU(cudaMemsetAsync(s->d_bar, 0, 4 * sizeof(double), s->stream));
    const char* mode = getenv("DVD_SWAP");
    double ok = (mode && std::string(mode) == "nccl") ? 0.0 : 1.0;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::vector<cudaIpcMemHandle_t> handles(s->world);
    unsigned char* d_h = nullptr;
    CU(cudaMalloc(&d_h, (size_t)s->world * 64));
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, s->buf[0]) != cudaSuccess) { cudaGetLastError(); ok = 0.0; std::memset(&mine, 0, sizeof mine); }
    CU(cudaMemcpyAsync(d_h + (size_t)s->rank * 64, &mine, 64, cudaMemcpyHostToDevice, s->stream));
    NC(g_nccl.AllGather(d_h + (size_t)s->rank * 64, d_h, 64, ncclChar, s->comm, s->stream));
    CU(cudaMemcpyAsync(handles.data(), d_h, (size_t)s->world * 64, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaFree(d_h));
    s->peer_base.assign(s->world, nullptr);
    if (ok != 0.0) {
        for (int r = 0; r < s->world; ++r) {
            if (r == s->rank) continue;
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError(); ok = 0.0; break;
            }
            s->peer_base[r] = static_cast<cplx*>(p);
        }
    }
    double agreed = 0.0;
    CU(cudaMemcpyAsync(s->d_bar + 2, &ok, sizeof ok, cudaMemcpyHostToDevice, s->stream));
    NC(g_nccl.AllReduce(s->d_bar + 2, s->d_bar + 3, 1, ncclDouble, ncclMin, s->comm, s->stream));
    CU(cudaMemcpyAsync(&agreed, s->d_bar + 3, sizeof agreed, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->peer_swap = agreed != 0.0;
    if (!s->peer_swap)
        for (auto& p : s->peer_base) if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
    // both chunks in place and every partner mapped: global<->local swaps ride on the next pass's load
    const char* fr = getenv("DVD_FUSED_REMAP");
    s->fused_remap = s->peer_swap && s->buf[1] != nullptr && !(fr && atoi(fr) == 0);
    const char* sr = getenv("DVD_STORE_REMAP");
    s->store_remap = sr ? std::max(0, std::min(2, atoi(sr))) : 2;
    return DVD_OK;
}

static int create_common(int n_qubits, int device, int rank, int world, const void* nccl_id, dvd_state** out) {
    if (!out) return fail(DVD_ERR_ARG, "out is null");
    *out = nullptr;
    if (world < 1 || (world & (world - 1))) return fail(DVD_ERR_ARG, "world must be a power of two");
    int g = 0; while ((1 << g) < world) ++g;
    if (n_qubits < 1 || n_qubits > 40) return fail(DVD_ERR_ARG, "n_qubits out of range [1,40]");
    if (n_qubits - g < 1) return fail(DVD_ERR_ARG, "need at least one local qubit per rank");
    if (rank < 0 || rank >= world) return fail(DVD_ERR_ARG, "bad rank");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return fail(DVD_ERR_CUDA, "Could not find any GPU.");   // circuit.rs:143-145
    if (device < 0 || device >= ndev) return fail(DVD_ERR_ARG, "bad device index");
    CU(cudaSetDevice(device));
    CU(kernels_init());
    dvd_state* s = new dvd_state();
    std::memset(&s->stats, 0, sizeof(s->stats));
    s->n_qubits = n_qubits; s->n_local = n_qubits - g; s->rank = rank; s->world = world; s->device = device;
    s->n_amps = 1ull << s->n_local;
    s->rank_bits = (uint64_t)rank << s->n_local;
    if (const char* e = getenv("DVD_PLAN_CANDIDATES")) { s->opt.candidates = std::max(1, atoi(e)); s->opt.portfolio = false; }
    if (const char* e = getenv("DVD_PLAN_PORTFOLIO")) s->opt.portfolio = atoi(e) != 0;
    if (const char* e = getenv("DVD_MACRO_OPS")) s->opt.macro_ops = atoi(e) != 0;
    if (const char* e = getenv("DVD_LAZY_ZERO")) s->lazy_zero = atoi(e) != 0;
    if (const char* e = getenv("DVD_BEST_GROUP")) s->opt.best_group = atoi(e) != 0;
    if (const char* e = getenv("DVD_RELABEL")) s->opt.relabel = atoi(e) != 0;
    if (const char* e = getenv("DVD_PLAN_CACHE")) s->plan_cache = atoi(e) != 0;
    if (const char* e = getenv("DVD_JIT")) s->jit_mode = std::string(e) == "sync" ? JIT_SYNC : std::max(0, std::min(2, atoi(e)));
    if (const char* e = getenv("DVD_JIT_MIN_QUBITS")) s->jit_min_qubits = atoi(e);
    s->perm.resize(n_qubits);
    for (int q = 0; q < n_qubits; ++q) s->perm[q] = q;
    auto cleanup = [&](int code) { dvd_destroy(s); return code; };
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && s->n_amps * sizeof(cplx) > free_b) {
        char buf[256];
        snprintf(buf, sizeof buf, "System requires more memory to simulate %d qubits: need %.1f GiB on device %d, %.1f GiB free",
                 n_qubits, s->n_amps * 16.0 / 1073741824.0, device, free_b / 1073741824.0);
        return cleanup(fail(DVD_ERR_ARG, buf));    // mirrors the panic at circuit.rs:123-128
    }
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&s->ev_t0) != cudaSuccess || cudaEventCreate(&s->ev_t1) != cudaSuccess)
        return cleanup(fail(DVD_ERR_CUDA, "stream/event creation failed"));
    for (int i = 0; i < 2; ++i)
        if (cudaEventCreateWithFlags(&s->ev_pack[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_comm[i], cudaEventDisableTiming) != cudaSuccess)
            return cleanup(fail(DVD_ERR_CUDA, "event creation failed"));
    // distributed states get room for two chunks when the device has it (fused remap, see dvd_state::buf): the second
    // chunk costs no more than the reference's own partner buffer (rust_communication.cu:178-197)
    cudaError_t e = cudaErrorMemoryAllocation;
    const char* fr = getenv("DVD_FUSED_REMAP");
    if (world > 1 && !(fr && atoi(fr) == 0) && 2 * s->n_amps * sizeof(cplx) + (1ull << 30) < free_b) {
        e = cudaMalloc(&s->buf[0], 2 * s->n_amps * sizeof(cplx));
        if (e == cudaSuccess) s->buf[1] = s->buf[0] + s->n_amps; else cudaGetLastError();
    }
    if (e != cudaSuccess) e = cudaMalloc(&s->buf[0], s->n_amps * sizeof(cplx));
    if (e != cudaSuccess) return cleanup(fail(DVD_ERR_CUDA, std::string("cudaMalloc(state): ") + cudaGetErrorString(e)));
    s->amp = s->buf[0];
    e = cudaMalloc(&s->d_tree, tree_size(s->n_local) * sizeof(double));
    if (e != cudaSuccess) return cleanup(fail(DVD_ERR_CUDA, std::string("cudaMalloc(tree): ") + cudaGetErrorString(e)));
    if (world > 1) {
        std::string why = g_nccl.load();
        if (!why.empty()) return cleanup(fail(DVD_ERR_NCCL, why));
        if (!nccl_id) return cleanup(fail(DVD_ERR_ARG, "nccl_id is null"));
        ncclUniqueId id;
        static_assert(sizeof(ncclUniqueId) == DVD_NCCL_ID_BYTES, "ncclUniqueId size");
        std::memcpy(&id, nccl_id, sizeof id);
        ncclResult_t r = g_nccl.CommInitRank(&s->comm, world, id, rank);
        if (r != ncclSuccess) return cleanup(fail(DVD_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)));
        int prc = map_peers(s);
        if (prc != DVD_OK) return cleanup(prc);
    }
    int rc = set_zero_state(s);
    if (rc != DVD_OK) return cleanup(rc);
    *out = s;
    return DVD_OK;
}

extern "C" {

int dvd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

double dvd_device_mem_mib(int device) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    return (double)p.totalGlobalMem / 1048576.0;
}

int dvd_peer_access_allowed(int a, int b) {
    int ok = 0;
    if (cudaDeviceCanAccessPeer(&ok, a, b) != cudaSuccess) { cudaGetLastError(); return 0; }
    return ok;
}

const char* dvd_last_error(void) { return g_last_error.c_str(); }

int dvd_create(int n_qubits, int device, dvd_state** out) { return create_common(n_qubits, device, 0, 1, nullptr, out); }

int dvd_create_distributed(int n_qubits, int device, int rank, int world, const void* nccl_id, dvd_state** out) {
    return create_common(n_qubits, device, rank, world, nccl_id, out);
}

int dvd_nccl_unique_id(void* out_id) {
    if (!out_id) return fail(DVD_ERR_ARG, "out_id is null");
    std::string why = g_nccl.load();
    if (!why.empty()) return fail(DVD_ERR_NCCL, why);
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    std::memcpy(out_id, &id, sizeof id);
    return DVD_OK;
}

int dvd_destroy(dvd_state* s) {
    if (!s) return DVD_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->comm_stream) cudaStreamSynchronize(s->comm_stream);
    if (s->comm && !s->comm_borrowed && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
    for (auto& p : s->peer_base) if (p) cudaIpcCloseMemHandle(p);
    if (s->buf[0]) cudaFree(s->buf[0]);
    for (auto& t : s->remap_timers) { if (t.t0) cudaEventDestroy(t.t0); if (t.t1) cudaEventDestroy(t.t1); }
    if (s->d_tree) cudaFree(s->d_tree);
    if (s->d_ident_tab) cudaFree(s->d_ident_tab);
    if (s->d_seq_cum) cudaFree(s->d_seq_cum);
    if (s->h_stage) cudaFreeHost(s->h_stage);
    for (auto& e : s->ev_stage) if (e) cudaEventDestroy(e);
    if (s->d_scratch) cudaFree(s->d_scratch);
    for (auto& t : s->tabs) {
        if (t.d) cudaFree(t.d);
        if (t.h) cudaFreeHost(t.h);
        if (t.done) cudaEventDestroy(t.done);
    }
    for (auto& b : s->swap_buf) if (b) cudaFree(b);
    if (s->d_bar) cudaFree(s->d_bar);
    for (int i = 0; i < 2; ++i) {
        if (s->ev_pack[i]) cudaEventDestroy(s->ev_pack[i]);
        if (s->ev_comm[i]) cudaEventDestroy(s->ev_comm[i]);
    }
    if (s->ev_t0) cudaEventDestroy(s->ev_t0);
    if (s->ev_t1) cudaEventDestroy(s->ev_t1);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
    delete s;
    return DVD_OK;
}

int dvd_reset_zero_state(dvd_state* s) {
    if (!s) return fail(DVD_ERR_ARG, "null state");
    CU(cudaSetDevice(s->device));
    s->pending.clear();
    return set_zero_state(s);
}

int dvd_apply_gate(dvd_state* s, const double m_re[4], const double m_im[4], int control, int target) {
    if (!s || !m_re || !m_im) return fail(DVD_ERR_ARG, "null argument");
    if (target < 0 || target >= s->n_qubits) return fail(DVD_ERR_ARG, "target qubit out of range");
    if (control < -1 || control >= s->n_qubits || control == target) return fail(DVD_ERR_ARG, "control qubit out of range");
    double m[8];
    for (int k = 0; k < 4; ++k) { m[2 * k] = m_re[k]; m[2 * k + 1] = m_im[k]; }
    HostGate g = make_gate(target, control, m, (int)s->pending.size());
    s->pending.push_back(g);
    return DVD_OK;
}

int dvd_apply_circuit(dvd_state* s, const dvd_gate* gates, int64_t n) {
    if (!s || (!gates && n > 0)) return fail(DVD_ERR_ARG, "null argument");
    for (int64_t i = 0; i < n; ++i) {
        const dvd_gate& in = gates[i];
        if (in.target < 0 || in.target >= s->n_qubits) return fail(DVD_ERR_ARG, "target qubit out of range");
        if (in.control < -1 || in.control >= s->n_qubits || in.control == in.target)
            return fail(DVD_ERR_ARG, "control qubit out of range");
    }
    s->pending.reserve(s->pending.size() + (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        HostGate g = make_gate(gates[i].target, gates[i].control, gates[i].m, (int)s->pending.size());
        s->pending.push_back(g);
    }
    return DVD_OK;
}

}  // extern "C"

// ---- flush ---------------------------------------------------------------------------------------
static SimpleOp simple_op(const HostGate& g) {
    SimpleOp op;
    std::memcpy(op.m, g.m, sizeof op.m);
    op.tbit = g.target();
    op.cbit = g.control();
    return op;
}

static int ensure_swap_buffers(dvd_state* s) {
    if (s->swap_buf[0]) return DVD_OK;
    const uint64_t half = s->n_amps / 2;
    uint64_t chunk = 1ull << 22;   // 64 MiB of complex128 per message
    if (const char* e = getenv("DVD_SWAP_CHUNK_LOG2")) { int v = atoi(e); if (v >= 4 && v <= 30) chunk = 1ull << v; }
    s->swap_chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(half, 1));
    for (auto& b : s->swap_buf) CU(cudaMalloc(&b, s->swap_chunk * sizeof(cplx)));
    return DVD_OK;
}

// Exchange physical rank-index qubit gq with physical local qubit lq: the half of the chunk whose bit
// lq differs from this rank's bit gq trades places with the partner rank's complementary half
// (partner = rank ^ 2^(gq-n_local) == Circuit::compute_partner_rank, circuit.rs:781-795).  Half a
// chunk crosses NVLink each way, against the reference's whole chunk both ways
// (rust_communication.cu:106-141).  Chunked and double buffered: pack k+1 / unpack k-1 on the compute
// stream overlap the ncclSend/ncclRecv of chunk k on the communication stream.
static int global_swap(dvd_state* s, int gq, int lq) {
    const int j = gq - s->n_local;
    const int b = (s->rank >> j) & 1;
    if (s->peer_swap) {
        // one kernel over NVLink peer memory: this rank moves the pairs of its half of the element range,
        // the partner the other half; barriers order it against the passes before and after on BOTH ranks
        const uint64_t half = s->n_amps / 2, share = half / 2;
        dvd_state::RemapTimer* tm = nullptr;
        TRY(timer_acquire(s, true, &tm));
        TRY(stream_barrier(s));
        if (tm) CU(cudaEventRecord(tm->t0, s->stream));
        CU(launch_swap_peer(s->amp, s->peer_cur(s->rank ^ (1 << j)), lq, b, b ? share : 0, b ? half : share, s->stream));
        if (tm) CU(cudaEventRecord(tm->t1, s->stream));
        TRY(stream_barrier(s));
        s->stats.kernel_launches++;
        s->stats.global_swaps++;
        s->stats.swap_bytes_sent += (int64_t)(half * sizeof(cplx));
        return DVD_OK;
    }
    TRY(ensure_swap_buffers(s));
    const int partner = s->rank ^ (1 << j);
    const int mybit = 1 - b;    // the half that leaves: local bit lq != my rank bit
    const uint64_t half = s->n_amps / 2, C = s->swap_chunk;
    const uint64_t nchunks = (half + C - 1) / C;
    cplx** sendb = &s->swap_buf[0];
    cplx** recvb = &s->swap_buf[2];
    dvd_state::RemapTimer* tm = nullptr;
    TRY(timer_acquire(s, true, &tm));
    if (tm) CU(cudaEventRecord(tm->t0, s->stream));
    for (uint64_t k = 0; k < nchunks; ++k) {
        const int slot = (int)(k & 1);
        const uint64_t first = k * C, cnt = std::min(C, half - first);
        CU(launch_pack_half(s->amp, lq, mybit, first, cnt, sendb[slot], s->stream));
        s->stats.kernel_launches++;
        CU(cudaEventRecord(s->ev_pack[slot], s->stream));
        CU(cudaStreamWaitEvent(s->comm_stream, s->ev_pack[slot], 0));
        NC(g_nccl.GroupStart());
        NC(g_nccl.Send(sendb[slot], cnt * 2, ncclDouble, partner, s->comm, s->comm_stream));
        NC(g_nccl.Recv(recvb[slot], cnt * 2, ncclDouble, partner, s->comm, s->comm_stream));
        NC(g_nccl.GroupEnd());
        CU(cudaEventRecord(s->ev_comm[slot], s->comm_stream));
        if (k >= 1) {
            const int ps = (int)((k - 1) & 1);
            const uint64_t pf = (k - 1) * C, pc = std::min(C, half - pf);
            CU(cudaStreamWaitEvent(s->stream, s->ev_comm[ps], 0));
            CU(launch_unpack_half(s->amp, lq, mybit, pf, pc, recvb[ps], s->stream));
            s->stats.kernel_launches++;
        }
    }
    {
        const uint64_t k = nchunks - 1;
        const int ps = (int)(k & 1);
        const uint64_t pf = k * C, pc = std::min(C, half - pf);
        CU(cudaStreamWaitEvent(s->stream, s->ev_comm[ps], 0));
        CU(launch_unpack_half(s->amp, lq, mybit, pf, pc, recvb[ps], s->stream));
        s->stats.kernel_launches++;
    }
    if (tm) CU(cudaEventRecord(tm->t1, s->stream));
    s->stats.global_swaps++;
    s->stats.swap_bytes_sent += (int64_t)(half * sizeof(cplx));
    return DVD_OK;
}

static int flush_impl(dvd_state* s) {
    if (s->pending.empty()) return DVD_OK;
    CU(cudaSetDevice(s->device));
    const bool tiled = !s->unfused && s->n_local >= TILE_BITS;
    const double chunk_bytes = (double)s->n_amps * 16.0;
    for (const HostGate& g : s->pending) {
        s->stats.gates_applied++;
        s->stats.gate_algorithmic_bytes += (g.cmask ? 1.0 : 2.0) * chunk_bytes;
    }
    // plan cache key: every field the planners read
    std::vector<uint64_t> key;
    if (s->plan_cache) {
        key.reserve(s->pending.size() * 11 + 2);
        key.push_back((tiled ? 1 : 0) | (s->zero_mask() << 1));   // (plans are chosen by the traffic they need given what is known to be zero)
        key.push_back((uint64_t)s->pending.size());
        for (const HostGate& g : s->pending) {
            key.push_back(g.tmask); key.push_back(g.cmask ^ (g.diag ? 1ull << 63 : 0));
            for (int k = 0; k < 8; ++k) { uint64_t b; std::memcpy(&b, &g.m[k], 8); key.push_back(b); }
            key.push_back((uint64_t)(int64_t)g.gate_idx);
        }
    }
    dvd_state::PlanCache* entry = nullptr;
    if (s->plan_cache)
        for (auto& c : s->caches) if (c.valid && c.key == key) entry = &c;
    const bool hit = entry != nullptr;
    if (hit) s->stats.plan_cache_hits++;
    if (!hit) {
        if (s->caches.size() < dvd_state::PLAN_CACHE_ENTRIES) { s->caches.emplace_back(); entry = &s->caches.back(); }
        else { entry = &s->caches[0]; for (auto& c : s->caches) if (!c.valid || c.stamp < entry->stamp) entry = &c; }
        entry->valid = false;
        entry->store.clear();
        // the portfolio winners of this gate-list structure, if it was planned before (new angles of the same circuit)
        PlanOptions popt = s->opt;
        popt.choices = s->memos.begin(s->pending, (tiled ? 1 : 0) | (s->zero_mask() << 1));
        std::vector<DistStep> steps;
        // every local step is planned up front so that all phase tables go to the device in one copy; the op
        // lists travel as kernel parameters
        std::vector<std::vector<Pass>> plans;
        size_t total_tabs = 0;
        try {
            // level 0 (fused mode only): CNOT-conjugated diagonal runs -> parity phases
            const std::vector<HostGate> fused = tiled ? fuse_diagonal_runs(s->pending) : s->pending;
            if (s->world > 1 && tiled) {
                // the schedule with the fewest passes among the tail-deferral thresholds, with its pass plans
                DistPlan dp = plan_distributed_tuned(fused, s->n_qubits, s->n_local, s->perm, /*restore_identity=*/true,
                                                     /*store_side=*/s->fused_remap ? s->store_remap : 0, popt, s->zero_mask());
                steps = std::move(dp.steps);
                plans = std::move(dp.plans);
                entry->store = std::move(dp.store);
            } else if (s->world > 1) {
                steps = plan_distributed(fused, s->n_qubits, s->n_local, s->perm, /*restore_identity=*/true);
            } else {
                DistStep st; st.kind = DistStep::LOCAL_GATES; st.gates = fused;
                steps.push_back(std::move(st));
            }
            plans.resize(steps.size());
            if (tiled)
                for (size_t i = 0; i < steps.size(); ++i)
                    if (steps[i].kind == DistStep::LOCAL_GATES) {
                        if (s->world == 1) {
                            PlanOptions o = popt;
                            o.zero_mask = s->zero_mask();      // after a reset: the plan whose early passes visit the fewest tiles
                            plans[i] = plan_local(steps[i].gates, s->n_local, s->n_qubits, o);
                        }
                        for (auto& p : plans[i]) total_tabs += p.tables.size();
                    }
        } catch (const std::exception& e) {
            return fail(DVD_ERR_INTERNAL, std::string("planner: ") + e.what());
        }
        entry->steps = std::move(steps);
        entry->plans = std::move(plans);
        entry->total_tabs = total_tabs;
        entry->key = std::move(key);
        entry->valid = s->plan_cache;
    }
    entry->stamp = ++s->cache_clock;
    std::vector<DistStep>& steps = entry->steps;
    std::vector<std::vector<Pass>>& plans = entry->plans;
    const size_t total_tabs = entry->total_tabs;
    if (tiled) {
        if (total_tabs) {
            dvd_state::TabSet& ts = s->tabs[s->n_flushes & 1];
            if (!ts.done) CU(cudaEventCreateWithFlags(&ts.done, cudaEventDisableTiming));
            if (ts.in_flight) { CU(cudaEventSynchronize(ts.done)); ts.in_flight = false; }   // the flush before last is over
            if (total_tabs > ts.cap) {
                if (ts.h) CU(cudaFreeHost(ts.h));
                if (ts.d) CU(cudaFree(ts.d));
                ts.h = nullptr; ts.d = nullptr; ts.cap = 0;
                const size_t cap = std::max<size_t>(total_tabs * 2, 64 * TABLE_ENTRIES);
                CU(cudaMallocHost(&ts.h, cap * sizeof(cplx)));
                CU(cudaMalloc(&ts.d, cap * sizeof(cplx)));
                ts.cap = cap;
            }
            size_t tat = 0;
            for (auto& pl : plans)
                for (auto& p : pl) {
                    if (!p.tables.empty()) std::memcpy(ts.h + tat, p.tables.data(), p.tables.size() * sizeof(cplx));
                    tat += p.tables.size();
                }
            CU(cudaMemcpyAsync(ts.d, ts.h, total_tabs * sizeof(cplx), cudaMemcpyHostToDevice, s->stream));
        }
    }
    cplx* const d_tabs = s->tabs[s->n_flushes & 1].d;
    size_t tat = 0;
    // Global<->local swaps waiting to ride on the next pass's load (fused remap): disjoint (gq, lq) pairs.
    std::vector<std::pair<int, int>> remap;
    bool fused_any = false;
    // store_swaps != nullptr: the pass STORES through the remap these swaps compose into (the layout restore riding on the
    // last gate pass): it reads this rank's buffer `cur` in place and writes buffer 1 - cur of this rank and of its partners
    auto run_pass = [&](Pass& p, const cplx* d_tab, const std::vector<std::pair<int, int>>* store_swaps = nullptr) -> int {
        RemapPlan sp;
        if (s->pushed_pending) {        // the pass in front of this one stored into other ranks' buffers: wait for all of them
            TRY(stream_barrier(s));
            s->pushed_pending = false;
        }
        if (store_swaps) {
            if (!remap.empty()) return fail(DVD_ERR_INTERNAL, "a pass cannot carry swaps on its load and on its store");
            if (!compose_remap(*store_swaps, s->n_local, s->rank, &sp, /*inverse=*/true)) return fail(DVD_ERR_INTERNAL, "store-side remap over too many positions");
            if (!sp.on) store_swaps = nullptr;      // the swaps cancel: a plain pass
            else TRY(materialize(s));               // a store-side pass writes every amplitude of the new layout: no implied zeros
        }
        PassParams& pp = s->pass_params;
        pp.pd = p.desc;
        pp.pd.rank_bits = s->rank_bits;
        // support tracking: implied zeros are not read, all-zero tiles are not launched
        const uint64_t zm = s->zero_mask();
        pp.pd.zero_mask = zm;
        uint64_t tile_mask = 0;
        for (int k = 0; k < TILE_BITS; ++k) tile_mask |= 1ull << pp.pd.tile_q[k];
        int zregs = 0;
        for (int k = 0; k < REG_BITS; ++k)
            if ((zm >> pp.pd.tile_q[IO_GROUP * REG_BITS + k]) & 1ull) zregs |= 1 << k;
        pp.pd.zero_regbits = (int8_t)zregs;
        cplx* out = s->amp;
        uint64_t lmask = 0;
        RemapPlan rp;
        if (!remap.empty()) {
            if (!compose_remap(remap, s->n_local, s->rank, &rp)) return fail(DVD_ERR_INTERNAL, "fused remap over too many positions");
            if (!rp.on) remap.clear();          // the swaps cancel: nothing moves
        }
        if (!remap.empty()) {
            // the pass reads buffer `cur` of this rank and of its partners, writes buffer 1 - cur of this rank
            apply_remap(rp, &pp.pd);
            lmask = rp.lmask;
            for (int sel = 0; sel < (1 << rp.n_sel); ++sel)
                pp.pd.remap_src[sel] = rp.src_rank[sel] == s->rank ? s->amp : s->peer_cur(rp.src_rank[sel]);
            out = s->buf[1 - s->cur];
            // every rank's buffer `cur` is complete (and nobody still reads the buffer this pass overwrites: the
            // previous fused pass, which read it, lies before this barrier on every rank)
            TRY(stream_barrier(s));
        }
        if (store_swaps) {
            apply_remap(sp, &pp.pd, /*store_side=*/true);
            for (int sel = 0; sel < (1 << sp.n_sel); ++sel)
                pp.pd.remap_src[sel] = sp.src_rank[sel] == s->rank ? s->buf[1 - s->cur] : s->peer_other(sp.src_rank[sel]);
            // nobody still reads the buffers this pass overwrites on every rank: the last pass whose load read them lies
            // before this barrier.  The barrier that ends the flush makes the stores of all ranks visible to their owners.
            TRY(stream_barrier(s));
        }
        // tiles whose fixed bits make every source amplitude zero are not launched (bits of the swapped-in
        // positions are rank-index bits of the source: never implied)
        if (pp.pd.remap_on || pp.pd.remap_st) {
            // the index bits that select the source rank become the lowest bits of the CTA index: tiles fetched over
            // NVLink and tiles fetched from local HBM alternate in launch order (on failure the default order stays)
            uint64_t sel_bits = 0;
            for (int k = 0; k < pp.pd.remap_n; ++k) sel_bits |= 1ull << pp.pd.remap_lq[k];
            fill_cta_runs_ex(pp.pd, zm & ~tile_mask & ~lmask, sel_bits);
        } else if (zm & ~tile_mask) {
            fill_cta_runs_sparse(pp.pd, zm & ~tile_mask);   // on failure the full grid stays
        }
        pp.pd.tables = d_tab;
        pp.pd.tid_off = reinterpret_cast<const uint64_t*>(d_tab + p.tid_off_slot);
        std::memcpy(pp.ops, p.ops.data(), p.ops.size() * sizeof(DevOp));
        dvd_state::RemapTimer* tm = nullptr;
        if (!remap.empty() || store_swaps) {
            TRY(timer_acquire(s, false, &tm));
            if (tm) { tm->is_store = store_swaps != nullptr; CU(cudaEventRecord(tm->t0, s->stream)); }
        }
        bool launched = false;
        if (s->jit_mode != JIT_OFF && s->n_local >= s->jit_min_qubits) {
            std::string jerr;
            launched = jit_launch(p, s->jit_mode, s->device, out, pp, s->stream, &jerr);
            if (launched) s->stats.jit_launches++;
            else if (!jerr.empty()) s->jit_error = jerr;
        }
        if (!launched) CU(launch_tile_pass(out, pp, s->stream));
        if (tm) CU(cudaEventRecord(tm->t1, s->stream));
        s->stats.kernel_launches++; s->stats.tile_passes++;
        s->stats.stage_switches += p.n_switches;
        s->stats.pass_fp64_instr += (double)p.fp64_per_thread * (double)(1ull << pp.pd.n_cta_bits) * NTHREADS;
        {   // HBM bytes this launch moves: its tiles are written in full, read where the input can be non-zero
            const double tiles_bytes = (double)(1ull << pp.pd.n_cta_bits) * TILE_AMPS * sizeof(cplx);
            s->stats.pass_bytes += tiles_bytes * (1.0 + 1.0 / (double)(1ull << __builtin_popcountll(zm & tile_mask & ~lmask)));
        }
        if (!remap.empty() || store_swaps) {
            const RemapPlan& mp = store_swaps ? sp : rp;
            const double chunk = (double)s->n_amps * sizeof(cplx);
            int local_sel = 0;
            for (int sel = 0; sel < (1 << mp.n_sel); ++sel) local_sel += mp.src_rank[sel] == s->rank;
            const double moved = chunk * (1.0 - (double)local_sel / (double)(1 << mp.n_sel));   // pulled (pushed) over NVLink
            if (store_swaps) {
                for (auto& sw : *store_swaps) s->stats.global_swaps += sw.first >= s->n_local;
                lmask = ~0ull; s->stats.store_remap_passes++; s->pushed_pending = true;
            }
            else s->stats.global_swaps += (int64_t)remap.size();
            s->stats.swap_bytes_sent += (int64_t)moved;
            s->stats.remap_passes++;
            s->stats.remap_bytes_in += moved;
            s->cur = 1 - s->cur;
            s->amp = s->buf[s->cur];
            s->support |= lmask;          // the rebuilt positions may hold former rank-index qubits: never implied zero
            remap.clear();
            fused_any = true;
        }
        s->support |= p.touch_mask;
        return DVD_OK;
    };
    // swaps with no pass to ride on (end of the gate list): an empty pass over the lowest tile, i.e. a plain
    // out-of-place gather through the same load path
    auto flush_remap = [&]() -> int {
        if (remap.empty()) return DVD_OK;
        if (!s->ident_pass) {
            s->ident_pass.reset(new Pass(make_identity_pass(s->n_local)));
            CU(cudaMalloc(&s->d_ident_tab, s->ident_pass->tables.size() * sizeof(cplx)));
            CU(cudaMemcpyAsync(s->d_ident_tab, s->ident_pass->tables.data(), s->ident_pass->tables.size() * sizeof(cplx),
                               cudaMemcpyHostToDevice, s->stream));
            CU(cudaStreamSynchronize(s->stream));     // the host copy lives in the Pass, but keep the upload simple
        }
        return run_pass(*s->ident_pass, s->d_ident_tab);
    };
    for (size_t i = 0; i < steps.size(); ++i) {
        DistStep& st = steps[i];
        if (st.kind == DistStep::GLOBAL_SWAP) {
            if (tiled && s->fused_remap) {
                // consecutive swaps compose into one remap as long as they stay within MAX_REMAP rank-index and
                // MAX_REMAP local positions (any permutation of those: shared positions included)
                remap.push_back({st.gq, st.lq});
                RemapPlan probe;
                if (!compose_remap(remap, s->n_local, s->rank, &probe)) {
                    remap.pop_back();
                    TRY(flush_remap());
                    remap.push_back({st.gq, st.lq});
                }
                continue;
            }
            TRY(materialize(s)); TRY(global_swap(s, st.gq, st.lq));
            continue;
        }
        if (tiled) {
            for (size_t k = 0; k < plans[i].size(); ++k) {
                Pass& p = plans[i][k];
                const bool store_here = k + 1 == plans[i].size() && i < entry->store.size() && !entry->store[i].empty();
                TRY(run_pass(p, d_tabs + tat, store_here ? &entry->store[i] : nullptr));
                tat += p.tables.size();
            }
        } else {
            TRY(materialize(s));
            for (const HostGate& g : st.gates) {
                CU(launch_simple_gate(s->amp, s->n_local, s->rank_bits, simple_op(g), s->stream));
                s->stats.kernel_launches++; s->stats.simple_passes++;
                s->stats.pass_bytes += (g.cmask ? 1.0 : 2.0) * chunk_bytes;
            }
        }
    }
    TRY(flush_remap());
    // every rank has finished reading this rank's buffers before anything after the flush (an observation, a load,
    // the destruction of the state) touches them
    if (fused_any) TRY(stream_barrier(s));
    s->pushed_pending = false;
    if (tiled && total_tabs) {
        dvd_state::TabSet& ts = s->tabs[s->n_flushes & 1];
        CU(cudaEventRecord(ts.done, s->stream));
        ts.in_flight = true;
    }
    s->n_flushes++;
    s->pending.clear();
    s->tree_valid = false; s->seq_valid = false;
    return DVD_OK;
}

// Everything queued has run and the buffer holds the state (observation points).
static int sync_state(dvd_state* s) {
    TRY(flush_impl(s));
    CU(cudaSetDevice(s->device));
    return materialize(s);
}

static int ensure_tree(dvd_state* s) {
    TRY(sync_state(s));
    if (s->tree_valid) return DVD_OK;
    CU(launch_build_tree(s->amp, s->n_local, s->d_tree, s->stream));
    s->stats.kernel_launches += 1 + std::max(0, s->n_local - BLK_BITS);
    s->tree_valid = true;
    return DVD_OK;
}

extern "C" {

int dvd_flush(dvd_state* s) {
    if (!s) return fail(DVD_ERR_ARG, "null state");
    return flush_impl(s);
}

int dvd_synchronize(dvd_state* s) {
    if (!s) return fail(DVD_ERR_ARG, "null state");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaStreamSynchronize(s->comm_stream));
    return DVD_OK;
}

int dvd_probabilities(dvd_state* s, double* out, int64_t first, int64_t count) {
    if (!s || (!out && count > 0)) return fail(DVD_ERR_ARG, "null argument");
    if (first < 0 || count < 0 || (uint64_t)(first + count) > s->n_amps) return fail(DVD_ERR_ARG, "range outside the local chunk");
    TRY(sync_state(s));
    const uint64_t CH = 1ull << 24;
    TRY(ensure_scratch(s, std::min<uint64_t>(CH, std::max<int64_t>(count, 1))));
    for (uint64_t done = 0; done < (uint64_t)count; done += CH) {
        const uint64_t c = std::min<uint64_t>(CH, count - done);
        CU(launch_probabilities(s->amp, first + done, c, s->d_scratch, s->stream));
        s->stats.kernel_launches++;
        CU(cudaMemcpyAsync(out + done, s->d_scratch, c * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    return DVD_OK;
}

int dvd_norm(dvd_state* s, double* out) {
    if (!s || !out) return fail(DVD_ERR_ARG, "null argument");
    TRY(ensure_tree(s));
    const double* root = s->d_tree + tree_level_offset(s->n_local, s->n_local);
    if (s->world > 1) {
        TRY(ensure_scratch(s, 1));
        NC(g_nccl.AllReduce(root, s->d_scratch, 1, ncclDouble, ncclSum, s->comm, s->stream));
        root = s->d_scratch;
    }
    CU(cudaMemcpyAsync(out, root, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return DVD_OK;
}

// Reference summation order: one sequential pass over the chunk per state (see k_seq_block_cum).
static const int SEQ_SAMPLER_MAX_LOCAL = 30;
static int ensure_seq_cum(dvd_state* s) {
    TRY(sync_state(s));
    if (s->seq_valid) return DVD_OK;
    const int nb = std::min(s->n_local, BLK_BITS);
    if (!s->d_seq_cum) CU(cudaMalloc(&s->d_seq_cum, (1ull << (s->n_local - nb)) * sizeof(double)));
    CU(launch_seq_block_cum(s->amp, s->n_local, s->d_seq_cum, s->stream));
    s->stats.kernel_launches++;
    s->seq_valid = true;
    return DVD_OK;
}

int dvd_set_sampler(dvd_state* s, int order) {
    if (!s) return fail(DVD_ERR_ARG, "null state");
    if (order != DVD_SAMPLER_TREE && order != DVD_SAMPLER_SEQUENTIAL) return fail(DVD_ERR_ARG, "sampler order must be 0 (tree) or 1 (sequential)");
    if (order == DVD_SAMPLER_SEQUENTIAL && s->n_local > SEQ_SAMPLER_MAX_LOCAL)
        return fail(DVD_ERR_ARG, "the reference-order sampler walks the chunk sequentially: at most 30 local qubits");
    s->sampler = order;
    return DVD_OK;
}

int dvd_sample(dvd_state* s, const double* uniforms, int64_t shots, uint64_t* out) {
    if (!s || ((!uniforms || !out) && shots > 0)) return fail(DVD_ERR_ARG, "null argument");
    if (shots < 0) return fail(DVD_ERR_ARG, "negative shot count");
    const bool seq = s->sampler == DVD_SAMPLER_SEQUENTIAL;
    if (seq) TRY(ensure_seq_cum(s)); else TRY(ensure_tree(s));
    if (shots == 0) return DVD_OK;
    const int nb_seq = std::min(s->n_local, BLK_BITS);
    // this chunk's total probability: the root of the pairwise tree, or the last sequential cumulative sum
    const double* d_total = seq ? s->d_seq_cum + ((1ull << (s->n_local - nb_seq)) - 1)
                                : s->d_tree + tree_level_offset(s->n_local, s->n_local);
    // scratch layout (in doubles): [u: shots][out: shots (u64)][sel: shots (i32, padded)][totals: world]
    const size_t need = (size_t)shots * 3 + s->world + 8;
    TRY(ensure_scratch(s, need));
    double* d_u = s->d_scratch;
    unsigned long long* d_out = reinterpret_cast<unsigned long long*>(s->d_scratch + shots);
    int32_t* d_sel = reinterpret_cast<int32_t*>(s->d_scratch + 2 * shots);
    double* d_tot = s->d_scratch + 3 * shots;
    if (s->world == 1) {
        CU(cudaMemcpyAsync(d_u, uniforms, shots * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        if (seq) CU(launch_sample_seq(s->amp, s->n_local, s->d_seq_cum, d_u, nullptr, 0, 0, shots, d_out, s->stream));
        else CU(launch_sample(s->amp, s->n_local, s->d_tree, d_u, nullptr, 0, 0, shots, d_out, s->stream));
        s->stats.kernel_launches++;
    } else {
        // sample_distributed, circuit_distributed.rs:42-129: per-rank totals -> rank per shot (first
        // draw) -> local index on that rank (second draw) -> index + amps_per_rank * rank.
        NC(g_nccl.AllGather(d_total, d_tot, 1, ncclDouble, s->comm, s->stream));
        std::vector<double> tot(s->world);
        CU(cudaMemcpyAsync(tot.data(), d_tot, s->world * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        std::vector<double> cum(s->world + 1, 0.0);
        for (int r = 0; r < s->world; ++r) cum[r + 1] = cum[r] + tot[r];        // utils.rs:270-274
        std::vector<int32_t> sel(shots);
        for (int64_t k = 0; k < shots; ++k) {
            const double xsi = uniforms[k] * cum[s->world];                       // utils.rs:260
            int idx = 0;
            while (idx <= s->world && !(xsi <= cum[idx])) ++idx;                   // utils.rs:262-266
            sel[k] = idx == 0 ? 0 : std::min(idx - 1, s->world - 1);
        }
        CU(cudaMemcpyAsync(d_u, uniforms + shots, shots * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemcpyAsync(d_sel, sel.data(), shots * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemsetAsync(d_out, 0, shots * sizeof(unsigned long long), s->stream));
        if (seq) CU(launch_sample_seq(s->amp, s->n_local, s->d_seq_cum, d_u, d_sel, s->rank, (uint64_t)s->rank * s->n_amps,
                                      shots, d_out, s->stream));
        else CU(launch_sample(s->amp, s->n_local, s->d_tree, d_u, d_sel, s->rank, (uint64_t)s->rank * s->n_amps,
                              shots, d_out, s->stream));
        s->stats.kernel_launches++;
        NC(g_nccl.AllReduce(d_out, d_out, shots, ncclUint64, ncclSum, s->comm, s->stream));
        CU(cudaStreamSynchronize(s->stream));   // sel must outlive the copy
    }
    CU(cudaMemcpyAsync(out, d_out, shots * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return DVD_OK;
}

int dvd_extract_expectation_values(dvd_state* s, const uint64_t* samples, int64_t shots,
                                   const int32_t* qubits, int32_t n_obs, double* out) {
    if (!s) return fail(DVD_ERR_ARG, "null state");
    if (shots < 0 || n_obs < 0) return fail(DVD_ERR_ARG, "negative size");
    if (shots == 0 || n_obs == 0) return DVD_OK;
    if (!samples || !qubits || !out) return fail(DVD_ERR_ARG, "null argument");
    for (int o = 0; o < n_obs; ++o)
        if (qubits[o] < 0 || qubits[o] >= 64) return fail(DVD_ERR_ARG, "observable qubit out of range");
    CU(cudaSetDevice(s->device));
    const size_t total = (size_t)shots * n_obs;
    const size_t need = total + shots + (n_obs + 1) / 2 + 4;
    TRY(ensure_scratch(s, need));
    double* d_out = s->d_scratch;
    unsigned long long* d_samples = reinterpret_cast<unsigned long long*>(s->d_scratch + total);
    int* d_q = reinterpret_cast<int*>(s->d_scratch + total + shots);
    CU(cudaMemcpyAsync(d_samples, samples, shots * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(d_q, qubits, n_obs * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CU(launch_extract_expectation(d_samples, shots, d_q, n_obs, d_out, s->stream));
    s->stats.kernel_launches++;
    CU(cudaMemcpyAsync(out, d_out, total * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return DVD_OK;
}

int dvd_expectation_z(dvd_state* s, double* out) {
    if (!s || !out) return fail(DVD_ERR_ARG, "null argument");
    TRY(sync_state(s));
    TRY(ensure_scratch(s, ez_partial_size() + 128));
    double* d_out = s->d_scratch + ez_partial_size();
    CU(launch_expectation_z(s->amp, s->n_local, s->n_qubits, s->rank_bits, s->d_scratch, d_out, s->stream));
    s->stats.kernel_launches += 2;
    if (s->world > 1) NC(g_nccl.AllReduce(d_out, d_out, s->n_qubits, ncclDouble, ncclSum, s->comm, s->stream));
    CU(cudaMemcpyAsync(out, d_out, s->n_qubits * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return DVD_OK;
}

// Pinned staging for dvd_read_state / dvd_load_state: two slots of [re | im], STAGE_AMPS amplitudes each.
static const uint64_t STAGE_AMPS = 1ull << 22;      // 32 MiB per array and slot
static int ensure_staging(dvd_state* s) {
    if (s->h_stage) return DVD_OK;
    CU(cudaMallocHost(&s->h_stage, 4 * STAGE_AMPS * sizeof(double)));
    for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&s->ev_stage[i], cudaEventDisableTiming));
    return DVD_OK;
}

// retrieve_amplitudes_on_host (rust_communication.cu:450-482 + the interleave loop of circuit.rs:396-401): the split into
// real and imaginary arrays runs on the device; chunks travel through two pinned slots, so that the copy of chunk k + 1
// over PCIe overlaps the host's memcpy of chunk k into the caller's (pageable) arrays.
int dvd_read_state(dvd_state* s, double* re, double* im, int64_t first, int64_t count) {
    if (!s || ((!re || !im) && count > 0)) return fail(DVD_ERR_ARG, "null argument");
    if (first < 0 || count < 0 || (uint64_t)(first + count) > s->n_amps) return fail(DVD_ERR_ARG, "range outside the local chunk");
    TRY(sync_state(s));
    if (count == 0) return DVD_OK;
    const uint64_t CH = STAGE_AMPS, n = (uint64_t)count;
    if (n <= 4096) {      // a handful of amplitudes: one small copy
        std::vector<cplx> tmp(n);
        CU(cudaMemcpyAsync(tmp.data(), s->amp + first, n * sizeof(cplx), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        for (uint64_t i = 0; i < n; ++i) { re[i] = tmp[i].x; im[i] = tmp[i].y; }
        return DVD_OK;
    }
    TRY(ensure_staging(s));
    TRY(ensure_scratch(s, 4 * std::min(CH, n)));
    const uint64_t cap = std::min(CH, n), n_chunks = (n + CH - 1) / CH;
    auto issue = [&](uint64_t k) -> int {
        const int slot = (int)(k & 1);
        const uint64_t off = k * CH, c = std::min(CH, n - off);
        double* d_re = s->d_scratch + (size_t)slot * 2 * cap;
        double* h_re = s->h_stage + (size_t)slot * 2 * CH;
        CU(launch_split_re_im(s->amp + first + off, c, d_re, d_re + cap, s->stream));
        s->stats.kernel_launches++;
        CU(cudaMemcpyAsync(h_re, d_re, c * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(h_re + CH, d_re + cap, c * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaEventRecord(s->ev_stage[slot], s->stream));
        return DVD_OK;
    };
    TRY(issue(0));
    for (uint64_t k = 0; k < n_chunks; ++k) {
        if (k + 1 < n_chunks) TRY(issue(k + 1));
        const int slot = (int)(k & 1);
        const uint64_t off = k * CH, c = std::min(CH, n - off);
        CU(cudaEventSynchronize(s->ev_stage[slot]));
        const double* h_re = s->h_stage + (size_t)slot * 2 * CH;
        std::memcpy(re + off, h_re, c * sizeof(double));
        std::memcpy(im + off, h_re + CH, c * sizeof(double));
    }
    return DVD_OK;
}

// load_amplitudes_local_on_device / split_amplitudes_between_gpus (rust_communication.cu:384-448), the same pipeline
// in the other direction: host memcpy of chunk k + 1 into a pinned slot overlaps the upload + join of chunk k.
int dvd_load_state(dvd_state* s, const double* re, const double* im, int64_t first, int64_t count) {
    if (!s || ((!re || !im) && count > 0)) return fail(DVD_ERR_ARG, "null argument");
    if (first < 0 || count < 0 || (uint64_t)(first + count) > s->n_amps) return fail(DVD_ERR_ARG, "range outside the local chunk");
    TRY(sync_state(s));
    s->tree_valid = false; s->seq_valid = false;
    if (count == 0) return DVD_OK;
    const uint64_t CH = STAGE_AMPS, n = (uint64_t)count;
    TRY(ensure_staging(s));
    TRY(ensure_scratch(s, 4 * std::min(CH, n)));
    const uint64_t cap = std::min(CH, n), n_chunks = (n + CH - 1) / CH;
    for (uint64_t k = 0; k < n_chunks; ++k) {
        const int slot = (int)(k & 1);
        const uint64_t off = k * CH, c = std::min(CH, n - off);
        double* d_re = s->d_scratch + (size_t)slot * 2 * cap;
        double* h_re = s->h_stage + (size_t)slot * 2 * CH;
        if (k >= 2) CU(cudaEventSynchronize(s->ev_stage[slot]));     // the upload that last used this slot is over
        std::memcpy(h_re, re + off, c * sizeof(double));
        std::memcpy(h_re + CH, im + off, c * sizeof(double));
        CU(cudaMemcpyAsync(d_re, h_re, c * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemcpyAsync(d_re + cap, h_re + CH, c * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CU(cudaEventRecord(s->ev_stage[slot], s->stream));
        CU(launch_join_re_im(s->amp + first + off, c, d_re, d_re + cap, s->stream));
        s->stats.kernel_launches++;
    }
    CU(cudaStreamSynchronize(s->stream));
    return DVD_OK;
}

int dvd_fidelity(dvd_state* a, dvd_state* b, double* out) {
    if (!a || !b || !out) return fail(DVD_ERR_ARG, "null argument");
    if (a->n_qubits != b->n_qubits || a->world != b->world || a->rank != b->rank || a->device != b->device)
        return fail(DVD_ERR_ARG, "states have different shapes");
    TRY(sync_state(a));
    TRY(sync_state(b));
    CU(cudaStreamSynchronize(b->stream));
    TRY(ensure_scratch(a, 148 * 4 * 2 + 8));
    double* d_out = a->d_scratch + 148 * 4 * 2;
    CU(launch_dot(a->amp, b->amp, a->n_amps, a->d_scratch, d_out, a->stream));
    a->stats.kernel_launches += 2;
    if (a->world > 1) NC(g_nccl.AllReduce(d_out, d_out, 2, ncclDouble, ncclSum, a->comm, a->stream));
    double h[2];
    CU(cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, a->stream));
    CU(cudaStreamSynchronize(a->stream));
    *out = h[0] * h[0] + h[1] * h[1];   // norm_sqr of the complex overlap, circuit_metrics.rs:24
    return DVD_OK;
}

// A second resident state holding a copy of src's amplitudes: same shape, device and rank, and -- for distributed
// states -- src's communicator (borrowed: the snapshot must be destroyed before src).  This is what
// get_fidelity_between_two_states_with_parameters needs for its first state (circuit.rs:753-769 keeps the first
// state's amplitudes on the host; circuit_metrics.rs:35-92 distributed_dot reduces the partial dots over the ranks).
int dvd_snapshot(dvd_state* src, dvd_state** out) {
    if (!src || !out) return fail(DVD_ERR_ARG, "null argument");
    *out = nullptr;
    TRY(sync_state(src));
    CU(cudaSetDevice(src->device));
    dvd_state* s = new dvd_state();
    std::memset(&s->stats, 0, sizeof(s->stats));
    s->n_qubits = src->n_qubits; s->n_local = src->n_local; s->rank = src->rank; s->world = src->world; s->device = src->device;
    s->n_amps = src->n_amps; s->rank_bits = src->rank_bits;
    s->opt = src->opt; s->lazy_zero = src->lazy_zero; s->plan_cache = src->plan_cache;
    s->jit_mode = src->jit_mode; s->jit_min_qubits = src->jit_min_qubits;
    s->perm.resize(s->n_qubits);
    for (int q = 0; q < s->n_qubits; ++q) s->perm[q] = q;
    s->comm = src->comm; s->comm_borrowed = true;
    auto cleanup = [&](int code) { dvd_destroy(s); return code; };
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&s->ev_t0) != cudaSuccess || cudaEventCreate(&s->ev_t1) != cudaSuccess)
        return cleanup(fail(DVD_ERR_CUDA, "stream/event creation failed"));
    for (int i = 0; i < 2; ++i)
        if (cudaEventCreateWithFlags(&s->ev_pack[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_comm[i], cudaEventDisableTiming) != cudaSuccess)
            return cleanup(fail(DVD_ERR_CUDA, "event creation failed"));
    cudaError_t e = cudaMalloc(&s->buf[0], s->n_amps * sizeof(cplx));
    if (e != cudaSuccess) return cleanup(fail(DVD_ERR_CUDA, std::string("cudaMalloc(snapshot): ") + cudaGetErrorString(e)));
    s->amp = s->buf[0];
    e = cudaMalloc(&s->d_tree, tree_size(s->n_local) * sizeof(double));
    if (e != cudaSuccess) return cleanup(fail(DVD_ERR_CUDA, std::string("cudaMalloc(tree): ") + cudaGetErrorString(e)));
    if (s->world > 1) {
        e = cudaMalloc(&s->d_bar, 4 * sizeof(double));
        if (e != cudaSuccess) return cleanup(fail(DVD_ERR_CUDA, std::string("cudaMalloc(barrier): ") + cudaGetErrorString(e)));
        cudaMemsetAsync(s->d_bar, 0, 4 * sizeof(double), s->stream);
    }
    CU(cudaStreamSynchronize(src->stream));
    e = cudaMemcpyAsync(s->amp, src->amp, s->n_amps * sizeof(cplx), cudaMemcpyDeviceToDevice, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return cleanup(fail(DVD_ERR_CUDA, std::string("snapshot copy: ") + cudaGetErrorString(e)));
    s->support = ~0ull;
    *out = s;
    return DVD_OK;
}

int dvd_copy_state(dvd_state* dst, dvd_state* src) {
    if (!dst || !src) return fail(DVD_ERR_ARG, "null argument");
    if (dst->n_qubits != src->n_qubits || dst->world != src->world || dst->device != src->device)
        return fail(DVD_ERR_ARG, "states have different shapes");
    TRY(sync_state(src));
    dst->pending.clear();
    CU(cudaStreamSynchronize(src->stream));
    CU(cudaMemcpyAsync(dst->amp, src->amp, src->n_amps * sizeof(cplx), cudaMemcpyDeviceToDevice, dst->stream));
    dst->tree_valid = false; dst->seq_valid = false;
    dst->support = ~0ull;
    return DVD_OK;
}

int dvd_num_qubits(const dvd_state* s) { return s ? s->n_qubits : -1; }
int dvd_num_local_qubits(const dvd_state* s) { return s ? s->n_local : -1; }
int dvd_rank(const dvd_state* s) { return s ? s->rank : -1; }
int dvd_world(const dvd_state* s) { return s ? s->world : -1; }
int dvd_device(const dvd_state* s) { return s ? s->device : -1; }

int dvd_get_stats(const dvd_state* cs, dvd_stats* out) {
    if (!cs || !out) return fail(DVD_ERR_ARG, "null argument");
    dvd_state* s = const_cast<dvd_state*>(cs);
    // event pairs around the fused-remap passes / stand-alone exchanges recorded since the last reset
    double ms_remap = 0.0, ms_swap = 0.0, ms_store = 0.0;
    for (size_t i = 0; i < s->remap_timers_used; ++i) {
        const dvd_state::RemapTimer& t = s->remap_timers[i];
        float f = 0.f;
        if (cudaEventSynchronize(t.t1) == cudaSuccess && cudaEventElapsedTime(&f, t.t0, t.t1) == cudaSuccess)
            { (t.is_swap ? ms_swap : ms_remap) += f; if (t.is_store) ms_store += f; }
        else cudaGetLastError();
    }
    s->stats.remap_ms = ms_remap;
    s->stats.swap_ms = ms_swap;
    s->stats.store_remap_ms = ms_store;
    *out = s->stats;
    return DVD_OK;
}
int dvd_stats_reset(dvd_state* s) {
    if (!s) return fail(DVD_ERR_ARG, "null state");
    std::memset(&s->stats, 0, sizeof s->stats);
    s->remap_timers_used = 0;
    return DVD_OK;
}
int dvd_timer_begin(dvd_state* s) {
    if (!s) return fail(DVD_ERR_ARG, "null state");
    CU(cudaSetDevice(s->device));
    CU(cudaEventRecord(s->ev_t0, s->stream));
    return DVD_OK;
}
int dvd_timer_end(dvd_state* s, double* ms) {
    if (!s || !ms) return fail(DVD_ERR_ARG, "null argument");
    CU(cudaEventRecord(s->ev_t1, s->stream));
    CU(cudaEventSynchronize(s->ev_t1));
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, s->ev_t0, s->ev_t1));
    *ms = f;
    return DVD_OK;
}
int dvd_set_jit(dvd_state* s, int mode) {
    if (!s) return fail(DVD_ERR_ARG, "null state");
    if (mode < 0 || mode > 2) return fail(DVD_ERR_ARG, "jit mode must be 0 (off), 1 (background) or 2 (on first use)");
    s->jit_mode = mode;
    return DVD_OK;
}
int dvd_jit_wait(dvd_state* s) {
    (void)s;
    jit_wait();
    return DVD_OK;
}
int dvd_jit_info(dvd_state* s, int64_t* compiled, int64_t* failed, int64_t* pending, double* compile_seconds,
                 char* last_error, int64_t cap) {
    const JitStats st = jit_stats();
    if (compiled) *compiled = st.compiled;
    if (failed) *failed = st.failed;
    if (pending) *pending = st.pending;
    if (compile_seconds) *compile_seconds = st.compile_seconds;
    if (last_error && cap > 0) {
        std::string msg = s ? s->jit_error : std::string();
        if (msg.empty()) msg = jit_available();
        snprintf(last_error, (size_t)cap, "%s", msg.c_str());
    }
    return DVD_OK;
}
int dvd_jit_forms(int64_t out[7]) {
    if (!out) return fail(DVD_ERR_ARG, "null argument");
    const JitStats st = jit_stats();
    out[0] = st.tuning;
    for (int f = 0; f < 3; ++f) { out[1 + f] = st.chosen[f]; out[4 + f] = st.launches[f]; }
    return DVD_OK;
}
int dvd_set_unfused(dvd_state* s, int unfused) {
    if (!s) return fail(DVD_ERR_ARG, "null state");
    s->unfused = unfused != 0;
    return DVD_OK;
}

// ---- planner inspection ---------------------------------------------------------------------
static std::vector<HostGate> to_host_gates(const dvd_gate* gates, int64_t n) {
    std::vector<HostGate> v;
    v.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (gates[i].target < 0 || gates[i].target > 62 || gates[i].control < -1 || gates[i].control > 62 ||
            gates[i].control == gates[i].target)
            throw std::runtime_error("bad qubit index");
        v.push_back(make_gate(gates[i].target, gates[i].control, gates[i].m, (int)i));
    }
    return v;
}

int64_t dvd_plan_debug(int n_total, int n_local, const dvd_gate* gates, int64_t n_gates, int fuse, int32_t* out, int64_t cap) {
    try {
        PlanOptions opt;
        if (const char* e = getenv("DVD_BEST_GROUP")) opt.best_group = atoi(e) != 0;
        if (const char* e = getenv("DVD_PLAN_CANDIDATES")) { opt.candidates = std::max(1, atoi(e)); opt.portfolio = false; }
        if (const char* e = getenv("DVD_PLAN_PORTFOLIO")) opt.portfolio = atoi(e) != 0;
        if (const char* e = getenv("DVD_RELABEL")) opt.relabel = atoi(e) != 0;
        std::vector<Pass> passes = plan_local(fuse ? fuse_diagonal_runs(to_host_gates(gates, n_gates)) : to_host_gates(gates, n_gates), n_local, n_total, opt);
        std::vector<int32_t> v;
        v.push_back((int32_t)passes.size());
        for (auto& p : passes) {
            for (int k = 0; k < TILE_BITS; ++k) v.push_back(p.desc.tile_q[k]);
            v.push_back(p.n_switches);
            v.push_back((int32_t)p.ops.size());
            for (auto& op : p.ops) {
                v.push_back(op.gate_idx); v.push_back(op.code); v.push_back(op.group);
                v.push_back(op.tab); v.push_back(op.regm);
            }
        }
        if ((int64_t)v.size() > cap) return -(int64_t)v.size();
        std::memcpy(out, v.data(), v.size() * sizeof(int32_t));
        return (int64_t)v.size();
    } catch (const std::exception& e) {
        g_last_error = std::string("planner: ") + e.what();
        return INT64_MIN;
    }
}

// Generated source of the structure-specialised kernel of pass `pass_index` (development / tests).  Returns the
// length written (without the terminator), -needed if cap is too small, INT64_MIN on a planner error.
int64_t dvd_jit_debug_source(int n_total, int n_local, const dvd_gate* gates, int64_t n_gates, int pass_index,
                             int form, char* out, int64_t cap) {
    try {
        PlanOptions opt;
        std::vector<Pass> passes = plan_local(fuse_diagonal_runs(to_host_gates(gates, n_gates)), n_local, n_total, opt);
        if (pass_index < 0 || pass_index >= (int)passes.size()) return 0;
        const bool st = form >= 16;      // + 16: the variant whose store goes through a remap (PassDesc::remap_st)
        if (st) form -= 16;
        const std::string src = generate_pass_source(passes[pass_index], "dvd_pass_static", form < 0 || form >= FORM_COUNT ? FORM_CLASSIC2 : form, st);
        if ((int64_t)src.size() + 1 > cap) return -(int64_t)(src.size() + 1);
        std::memcpy(out, src.c_str(), src.size() + 1);
        return (int64_t)src.size();
    } catch (const std::exception& e) {
        g_last_error = std::string("planner: ") + e.what();
        return INT64_MIN;
    }
}

// NVRTC-compile a generated source (no GPU needed).  Returns the cubin size, or -1 with the log in dvd_last_error().
int64_t dvd_jit_debug_compile(const char* source) {
    std::vector<char> cubin;
    const std::string log = jit_compile(source ? source : "", &cubin);
    if (!log.empty()) { g_last_error = log; return -1; }
    return (int64_t)cubin.size();
}

int64_t dvd_plan_distributed_debug(int n_total, int n_local, const dvd_gate* gates, int64_t n_gates,
                                   int32_t* perm_io, int restore_identity, int32_t* out, int64_t cap) {
    try {
        std::vector<int> perm(perm_io, perm_io + n_total);
        // the schedule the engine runs (tail deferral included; the restore stays in the step list)
        std::vector<DistStep> steps = n_local >= TILE_BITS
            ? plan_distributed_tuned(to_host_gates(gates, n_gates), n_total, n_local, perm, restore_identity != 0, /*store_side=*/0, PlanOptions()).steps
            : plan_distributed(to_host_gates(gates, n_gates), n_total, n_local, perm, restore_identity != 0);
        for (int q = 0; q < n_total; ++q) perm_io[q] = perm[q];
        std::vector<int32_t> v;
        v.push_back((int32_t)steps.size());
        for (auto& st : steps) {
            v.push_back((int32_t)st.kind);
            v.push_back(st.gq); v.push_back(st.lq);
            v.push_back((int32_t)st.gates.size());
            for (auto& g : st.gates) { v.push_back(g.gate_idx); v.push_back(g.target()); v.push_back(g.control()); }
        }
        if ((int64_t)v.size() > cap) return -(int64_t)v.size();
        std::memcpy(out, v.data(), v.size() * sizeof(int32_t));
        return (int64_t)v.size();
    } catch (const std::exception& e) {
        g_last_error = std::string("planner: ") + e.what();
        return INT64_MIN;
    }
}

}  // extern "C"
