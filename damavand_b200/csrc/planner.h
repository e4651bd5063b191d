// planner.h -- host-side gate scheduler (pure C++, no CUDA calls).
//
// Takes the reference's per-gate stream (one FFI call per gate in
// /root/reference/src/qubit_backend/circuit.rs:346-368) and turns it into a short list of passes
// over HBM.  Four levels:
//   0. fuse_diagonal_runs: runs of diagonal gates and CNOTs whose CNOTs cancel (e.g. the
//      RZ-CNOT-RZ-CNOT controlled-phase decomposition) become parity-phase ops: no data movement,
//      mutually commuting, foldable into one lazy scalar per thread;
//   1. distributed level (plan_distributed): logical->physical qubit map, global<->local qubit
//      swaps replacing the reference's per-gate full-chunk exchange
//      (src/qubit_backend/circuit_distributed_gpu.rs:40-147);
//   2. pass level (plan_local): choose TILE_BITS physical qubits per pass and pull every gate that
//      commutes its way to the front and fits the tile;
//   3. stage level: order the gates of a pass so that consecutive gates share a register group, and
//      compile its diagonal gates into a phase polynomial emitted as table / register-bit / pair ops.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

#include "tile_core.cuh"

namespace dvd {

// A gate on qubit masks.  Non-diagonal: tmask is one-hot (the target), cmask one-hot or 0.
// Diagonal: multiply by (parity(x & tmask) ? m11 : m00) where cmask == 0 or parity(x & cmask) == 1.
struct HostGate {
    uint64_t tmask = 0;
    uint64_t cmask = 0;
    double m[8];
    int gate_idx = -1;   // caller's index, -1 for fused / layout-restoring ops
    bool diag = false;
    int target() const { return __builtin_ctzll(tmask); }
    int control() const { return cmask ? __builtin_ctzll(cmask) : -1; }
};

HostGate make_gate(int target, int control, const double m[8], int gate_idx);

struct Pass {
    PassDesc desc;
    std::vector<DevOp> ops;
    std::vector<cplx> tables;  // the pass's table buffer (layout in tile_core.cuh), built by finish_tables()
    std::vector<TableDesc> tab_desc;   // one record per table op, in op order (ops[].tab indexes it)
    std::vector<cplx> tab_tile;        // TABLE_TILE_ENTRIES thread-index factors per table op
    std::vector<cplx> tab_bytes;       // byte sub-tables (TABLE_ENTRIES each)
    size_t tid_off_slot = 0;           // start (in cplx slots) of the [NGROUPS][NTHREADS] thread-offset table inside `tables`
    void finish_tables();              // concatenate [desc][tile][bytes][thread offsets] into `tables`, fill desc.run_*
    int n_switches = 0;        // stage switches inside the pass (shared-memory transposes)
    int fp64_per_thread = 0;   // estimate of the fp64 instructions (DADD / DMUL / DFMA) one thread executes for its 16
                               //   amplitudes in this pass (estimate_fp64): the passes of gate-dense circuits are bound by
                               //   the fp64 pipe, not by HBM, and bench.py reports that roofline next to the HBM one
    uint64_t touch_mask = 0;   // targets of the pass's non-diagonal gates (X / CNOT included): the only qubits whose
                               //   |0> can turn into a superposition in this pass (support tracking in the engine)
};

// The planners try a few variants and keep the cheapest (plan_local: tile candidates per pass x relabelling;
// plan_distributed_tuned: tail-deferral thresholds).  Which variant wins depends on the STRUCTURE of the gate list
// (qubits, controls, matrix classes), not on its angles, so a variational loop that flushes the same circuit with new
// parameters need not search again: the winners of one planning run are recorded here, in call order, and a later run
// with replay = true plans only those (any recorded choice yields a valid plan; a stale one is merely not the cheapest).
struct PlanChoices {
    std::vector<int> tape;     // plan_local: winning variant index per call
    size_t pos = 0;
    bool replay = false;
    int threshold = -1;        // plan_distributed_tuned: winning tail-deferral threshold
};

// The recorded choices of the last few gate-list structures (least recently used replaced).  begin() returns the entry
// of the structure of `gates` -- (qubits, controls, matrix classes) plus `flags`, whatever else the plans depend on --
// set up to replay if that structure was planned before, to record otherwise.  The pointer stays valid until the next
// begin().
class ChoiceMemoTable {
public:
    explicit ChoiceMemoTable(size_t capacity = 8) : capacity_(capacity) {}
    PlanChoices* begin(const std::vector<HostGate>& gates, uint64_t flags);
    size_t size() const { return entries_.size(); }
private:
    struct Entry { std::vector<uint64_t> skey; PlanChoices ch; uint64_t stamp = 0; };
    std::vector<Entry> entries_;
    size_t capacity_;
    uint64_t clock_ = 0;
};

struct PlanOptions {
    int min_low = 3;           // tile always contains physical qubits [0, min_low): 128 B segments
    int window = 16384;        // look-ahead (gates) when filling a pass
    int candidates = 12;       // tile candidates scored per pass (1 = first-come only)
    bool portfolio = true;     // plan_local also tries 4 and 2 candidates per pass and keeps the plan with the fewest passes
    int max_ops_per_pass = 1024;   // gates taken into one pass (halved and retried while the op stream exceeds MAX_OPS_PER_PASS)
    bool macro_ops = true;     // fuse 4-op runs on the four register bits into one dispatch (OC_REALPH4, OC_TWHAD4)
    uint64_t zero_mask = 0;    // local qubits still |0> in every populated basis state when the plan starts (the engine's support
                               //   tracking after a reset; 0 = dense): candidate plans are compared by the HBM traffic of their
                               //   passes (plan_traffic), and a pass neither launches nor reads what is zero by construction
    PlanChoices* choices = nullptr;   // record / replay of the portfolio winners (see PlanChoices)
    int defer_max_ops = -1;    // distributed schedule (plan_distributed_tuned): tail-deferral threshold; -1 = the best of a few
    bool best_group = false;   // stage order: group with the most runnable work (true) or group of the first waiting gate
    bool relabel = true;       // tile relabelling (measured on B200 in round 2: hea28 80 -> 55 passes, 236 -> 204 ms; DVD_RELABEL=0 turns it off): the pinned low tile positions are
                               //   physical qubits [0, min_low), present in every tile; at the end of a pass the logical
                               //   qubit held there may trade places with one of the tile's other qubits that the
                               //   following gates need sooner (three CNOTs = one more column swap of the permuting
                               //   transpose, no HBM traffic), so the next pass has 12 useful positions instead of 9.
                               //   The layout is back to identity when plan_local returns.
};

// Classify a 2x2 by exact zero / one tests on its entries.
void classify_gate(const double m[8], int32_t* kind, int8_t* d0_is_one);
inline bool is_diagonal(const double m[8]) { return m[2] == 0.0 && m[3] == 0.0 && m[4] == 0.0 && m[5] == 0.0; }

// Level 0.  Exact algebra on the gate list (no reordering across non-diagonal gates).
std::vector<HostGate> fuse_diagonal_runs(const std::vector<HostGate>& gates);

// Plan gates that are all executable locally: every non-diagonal gate has target < n_local.
// Qubits >= n_local (rank-index qubits) may appear in control masks and in diagonal gates.
// Requires n_local >= TILE_BITS.
// pass_of_gate (optional): for every input gate, the index of the pass that executes it.
std::vector<Pass> plan_local(const std::vector<HostGate>& gates, int n_local, int n_total,
                             const PlanOptions& opt, std::vector<int>* pass_of_gate = nullptr);

// HBM traffic of a plan in units of one full pass (read + write of every local amplitude), as the engine executes it
// with support tracking (engine.cu run_pass): a pass launches the tiles whose fixed bits avoid *zero_mask, reads the part
// of a tile that can be non-zero, writes the tile in full, and its non-diagonal targets leave the mask.  Dense: the
// number of passes.  zero_mask is updated to the state after the plan.
double plan_traffic(const std::vector<Pass>& passes, uint64_t* zero_mask);

// fp64 instructions per thread of an op list (per-kind costs of tile_core.cuh: general 2x2 = 16 per pair, real / RX-like
// = 8, Hadamard = 4, real + phase = 12, one complex multiply = 4).
int estimate_fp64(const std::vector<DevOp>& ops);

// A pass with no ops over the tile of the 12 lowest qubits: reads and writes every amplitude once.  The engine runs
// it when a fused remap (tile_core.cuh: PassDesc::remap_*) has no gate pass to ride on.
Pass make_identity_pass(int n_local);

// ---- distributed level ---------------------------------------------------------------------------
struct DistStep {
    enum Kind { LOCAL_GATES = 0, GLOBAL_SWAP = 1, LOCAL_SWAP = 2 } kind;
    // LOCAL_GATES: gates rewritten to physical qubits, all locally executable
    std::vector<HostGate> gates;
    // GLOBAL_SWAP: exchange physical global qubit `gq` (>= n_local) with physical local qubit `lq`
    // LOCAL_SWAP (only with local_swap_steps): exchange the physical LOCAL qubits `gq` and `lq` -- a transposition of the
    //   layout restore, which plan_distributed_tuned folds into the remap that rides on the last pass's store (it plans
    //   again with CNOT triples where it cannot: no executor runs a LOCAL_SWAP as a step)
    int gq = -1, lq = -1;
};

// A sequence of GLOBAL_SWAP steps composed into ONE fused remap (tile_core.cuh: PassDesc::remap_*), for rank `rank`.
// Returns false when the sequence touches more than MAX_REMAP rank-index or more than MAX_REMAP_LOCAL local positions (the
// caller then executes what it has and starts a new remap).  An identity composition gives on = false.
struct RemapPlan {
    bool on = false;
    int n_sel = 0;
    int sel_lq[MAX_REMAP];          // local positions of the index whose bits select the other rank
    int src_rank[1 << MAX_REMAP];   // the other rank per selector value
    int n_mv = 0;
    int mv_from[MAX_REMAP_LOCAL], mv_to[MAX_REMAP_LOCAL];
    uint64_t lmask = 0, rconst = 0;
    int n_global = 0, n_local_pos = 0;
};
// A swap is a pair of positions: (rank-index, local) or (local, local).  inverse = false: the LOAD-side form (where does
// the amplitude at new index i come from); inverse = true: the STORE-side form (where does the amplitude at old index i go).
bool compose_remap(const std::vector<std::pair<int, int>>& swaps, int n_local, int rank, RemapPlan* out, bool inverse = false);
// Fill the remap fields of a pass descriptor (all but remap_src, which the caller derives from src_rank).
void apply_remap(const RemapPlan& rp, PassDesc* pd, bool store_side = false);

// perm[logical] = physical, updated in place.  When `restore_identity` is set, trailing swaps bring
// the layout back to perm[q] = q (needed before measure / sample / readback, whose semantics are
// defined on the reference's contiguous-chunk layout, circuit.rs:135-136).
// local_swap_steps: the transpositions of LOCAL positions that the restore needs are emitted as LOCAL_SWAP steps instead
// of as CNOT triples inside a LOCAL_GATES step.
// tail_opt + defer_max_ops > 0 (tail deferral): a step's executable gates are planned into passes, and when the last of
// them has at most defer_max_ops ops its gates are kept back and run under the NEXT layout, after the swaps, where they
// share passes with that layout's gates (they stay in front of everything that waits, so the order remains legal).
std::vector<DistStep> plan_distributed(const std::vector<HostGate>& gates, int n_total, int n_local,
                                       std::vector<int>& perm, bool restore_identity, bool local_swap_steps = false,
                                       const PlanOptions* tail_opt = nullptr, int defer_max_ops = 0);
// The schedule the engine runs: plan_distributed with the tail-deferral threshold (0 = off, ...) that needs the fewest
// passes over HBM (swap rounds weighted in), together with the pass plan of every LOCAL_GATES step (plans[i] belongs to
// steps[i]; empty when n_local < TILE_BITS).  DVD_DEFER_TAILS=0 keeps the plain schedule.
// store_side: the executor can let the last pass of a LOCAL_GATES step STORE through a remap (PassDesc::remap_st: remote
// writes into the other chunk of this rank and of its partners) instead of the next pass LOADING through one.
//   0  never: every swap round rides on the load of the pass behind it (an empty pass if there is none);
//   1  the swaps that END the schedule -- the layout restore: rank-index swaps and transpositions of local positions,
//      which need a pass of their own otherwise -- ride on the store of the last gate pass;
//   2  every swap round rides on the store of the pass in front of it where it can.
// A round that is taken leaves the step list and is returned as store[i] for the LOCAL_GATES step i whose last pass
// carries it.  It can be taken when it composes into one remap (compose_remap) and the pass in front of it exists and
// is not the only pass of a step whose load already carries the previous round (one set of remap fields per pass).
struct DistPlan {
    std::vector<DistStep> steps;
    std::vector<std::vector<Pass>> plans;
    std::vector<std::vector<std::pair<int, int>>> store;     // per step; empty = a plain store
    int defer_max_ops = 0;     // the threshold that won
    int n_passes = 0;          // passes of the LOCAL_GATES steps
    int n_store = 0;           // swap rounds that ride on a store
};
// start_zero_mask: local qubits that are still |0> in every populated basis state when the schedule starts (the engine's
// support tracking after a reset; 0 = dense).  A storing pass writes every amplitude of the new layout, so the implied
// zeros must be stored first and the passes behind it run dense: a round in the middle of the schedule rides on a store
// only once every local qubit has been touched (measured on 2 x B200, random32 from a reset: 165 ms with the early
// round on a load, 219 ms with it on a store); until then it rides on the next load, which keeps the zeros implied.
DistPlan plan_distributed_tuned(const std::vector<HostGate>& gates, int n_total, int n_local, std::vector<int>& perm,
                                bool restore_identity, int store_side, const PlanOptions& opt, uint64_t start_zero_mask = 0);

}  // namespace dvd
