// emu.cpp -- CPU replay of the CUDA tile kernel and of the distributed swap schedule.
//
// TEST INFRASTRUCTURE ONLY (there is no GPU in the build container).  It includes the very same
// __host__ __device__ header the kernel is built from (damavand_b200/csrc/tile_core.cuh) and the
// real planner, and replays k_tile_pass thread by thread: same index arithmetic, same swizzle,
// same op decoding.  It cannot catch missing __syncthreads, but it does catch planner, commutation,
// tile-layout, stage-switch and control/target decoding bugs before a GPU is involved.
// Never loaded by the product.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <memory>
#include <stdexcept>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../damavand_b200/csrc/planner.cpp"
#include "../../include/damavand_b200.h"

using namespace dvd;

// Fused remap as the engine sets it up (engine.cu flush_impl / run_pass): `pairs` = (rank-index qubit, local qubit)
// swaps executed by this pass's load; `snapshot` = every rank's chunk before the pass (the buffers the ranks read).
struct RemapEmu {
    std::vector<std::pair<int, int>> pairs;
    const cplx* snapshot = nullptr;
    int rank = 0;
    uint64_t chunk = 0;
    cplx* store_all = nullptr;     // store-side remap (PassDesc::remap_st): every rank's OUTPUT chunk; the pass reads `amp` in place
};

static void run_pass(cplx* amp, const Pass& pass, uint64_t rank_bits, uint64_t zero_mask = 0, const RemapEmu* rm = nullptr) {
    PassDesc pd = pass.desc;
    uint64_t lmask = 0;
    if (rm && !rm->pairs.empty()) {
        RemapPlan rp;
        const bool st = rm->store_all != nullptr;
        if (!compose_remap(rm->pairs, pd.n_local, rm->rank, &rp, st)) throw std::runtime_error("emu: fused remap over too many positions");
        apply_remap(rp, &pd, st);
        if (!st) lmask = rp.lmask;
        for (int sel = 0; sel < (1 << rp.n_sel); ++sel)
            pd.remap_src[sel] = (st ? rm->store_all : rm->snapshot) + (uint64_t)rp.src_rank[sel] * rm->chunk;
        if (st && zero_mask) throw std::runtime_error("emu: store-side remap on a state with implied zeros");
        if (st && !rp.on) {      // the swaps cancel: a plain copy into the output chunk
            pd.remap_st = 1; pd.remap_n = 0; pd.remap_n_mv = 0; pd.remap_lmask = 0; pd.remap_const = 0;
            pd.remap_src[0] = rm->store_all + (uint64_t)rm->rank * rm->chunk;
        }
    }
    pd.rank_bits = rank_bits;
    pd.tables = pass.tables.data();
    pd.tid_off = reinterpret_cast<const uint64_t*>(pass.tables.data() + pass.tid_off_slot);
    // support tracking as the engine sets it up (engine.cu flush_impl): implied zeros are not read, tiles whose
    // fixed bits hit the mask are not visited
    uint64_t tile_mask = 0;
    for (int k = 0; k < TILE_BITS; ++k) tile_mask |= 1ull << pd.tile_q[k];
    pd.zero_mask = zero_mask;
    int zregs = 0;
    for (int k = 0; k < REG_BITS; ++k) if ((zero_mask >> pd.tile_q[IO_GROUP * REG_BITS + k]) & 1ull) zregs |= 1 << k;
    pd.zero_regbits = (int8_t)zregs;
    const uint64_t skip = zero_mask & ~tile_mask & ~lmask;
    bool sparse;
    if (pd.remap_on || pd.remap_st) {      // engine.cu run_pass: selector bits lowest in the CTA index
        uint64_t sel_bits = 0;
        for (int k = 0; k < pd.remap_n; ++k) sel_bits |= 1ull << pd.remap_lq[k];
        sparse = fill_cta_runs_ex(pd, skip, sel_bits);
    } else {
        sparse = skip != 0 && fill_cta_runs_sparse(pd, skip);
    }
    std::vector<bool> visited((size_t)1 << (pd.n_local > TILE_BITS ? pd.n_local - TILE_BITS : 0), false);
    const uint64_t ctas = 1ull << pd.n_cta_bits;
    if (!sparse && pd.n_cta_bits != pd.n_local - TILE_BITS) throw std::runtime_error("emu: n_cta_bits of the full grid");
    std::vector<cplx> tile(TILE_SLOTS);
    static cplx regs[NTHREADS][NREG];
    static ThreadCtx ctx[NTHREADS];
    for (uint64_t cta = 0; cta < ctas; ++cta) {
        const uint64_t base = sparse ? cta_base_runs(pd, cta) : cta_base(pd, cta);
        if (cta_base_runs(pd, cta) != base) throw std::runtime_error("emu: run-compressed CTA base differs");
        if (base & tile_mask) throw std::runtime_error("emu: CTA base overlaps the tile");
        {   // every tile is visited at most once, whatever the order of the CTA index bits
            uint64_t key = 0; int kb = 0;
            for (int q = 0; q < pd.n_local; ++q) if (!((tile_mask >> q) & 1ull)) key |= ((base >> q) & 1ull) << kb++;
            if (visited[key]) throw std::runtime_error("emu: a tile is visited twice");
            visited[key] = true;
        }
        if (base & zero_mask & ~lmask) { if (sparse) throw std::runtime_error("emu: sparse grid visits an all-zero tile"); continue; }
        const uint64_t gbase = base | pd.rank_bits;
        if (pd.n_tab > MAX_TABLE_OPS || pd.n_ops >= MAX_OPS_PER_PASS || (size_t)pd.n_ops != pass.ops.size())
            throw std::runtime_error("emu: pass exceeds the kernel parameter limits");
        std::vector<cplx> wcs(pd.n_tab > 0 ? pd.n_tab : 1);     // kernel prologue: per-CTA table constants
        for (int ti = 0; ti < pd.n_tab; ++ti) wcs[ti] = table_cta_const(pd.tables, ti, gbase);
        for (int tid = 0; tid < NTHREADS; ++tid) {
            const bool thread_zero = (tid_offset(pd, IO_GROUP, tid) & zero_mask) != 0;
            for (int j = 0; j < NREG; ++j) {
                const uint64_t i = base + tile_offset(pd, stage_idx(IO_GROUP, tid, j));
                if (!pd.remap_on) regs[tid][j] = (thread_zero || (j & pd.zero_regbits)) ? cplx{0.0, 0.0} : amp[i];
                else {      // tile_kernel.cuh tile_load, remap path
                    const uint64_t src = remap_index(pd, i);
                    regs[tid][j] = (src & zero_mask) ? cplx{0.0, 0.0} : pd.remap_src[remap_sel(pd, i)][src];
                }
            }
            ctx[tid].pidx = gbase | tile_offset(pd, stage_idx(IO_GROUP, tid, 0));
            for (int g = 0; g < NGROUPS; ++g)
                if (cta == 0 && tid_offset(pd, g, tid) != tile_offset(pd, stage_idx(g, tid, 0))) throw std::runtime_error("emu: thread offset table differs");
            ctx[tid].ph = cplx{1.0, 0.0};
            ctx[tid].ph_dirty = false;
            ctx[tid].tid = tid;
        }
        int cur = IO_GROUP;
        for (size_t oi = 0; oi < pass.ops.size(); ++oi) {
            const DevOp& op = pass.ops[oi];
            if (op.code < 0 || op.code >= OC_COUNT) throw std::runtime_error("emu: bad opcode");
            if (op.code >= OC_SWITCH) {
                const int from = (op.code - OC_SWITCH) / NGROUPS, to = (op.code - OC_SWITCH) % NGROUPS;
                if (from != cur || to != op.group) throw std::runtime_error("emu: switch does not start in the current group");
                if (from == to && !(op.flags & F_PERM)) throw std::runtime_error("emu: useless switch");
                for (int tid = 0; tid < NTHREADS; ++tid) flush_phase(regs[tid], ctx[tid]);
                std::vector<int> hits(TILE_SLOTS, 0);
                const unsigned v = (op.flags & F_PERM) ? perm_const(op, gbase) : 0;
                for (int tid = 0; tid < NTHREADS; ++tid)
                    for (int j = 0; j < NREG; ++j) {
                        unsigned idx = (unsigned)stage_idx(from, tid, j);
                        if (op.flags & F_PERM) idx = perm_index(op, v, idx);
                        if (idx >= (unsigned)TILE_AMPS) throw std::runtime_error("emu: permutation leaves the tile");
                        if (hits[smem_slot((int)idx)]++) throw std::runtime_error("emu: permutation is not a bijection");
                        tile[smem_slot((int)idx)] = regs[tid][j];
                    }
                for (int tid = 0; tid < NTHREADS; ++tid) {
                    for (int j = 0; j < NREG; ++j) regs[tid][j] = tile[smem_slot(stage_idx(to, tid, j))];
                    ctx[tid].pidx = thread_pidx(pd, gbase, to, tid);
                }
                cur = to;
                continue;
            }
            if (op.group != cur) throw std::runtime_error("emu: op not in its register group");
            if (is_table_op(op.code) && (op.flags & F_TABLE) && (op.tab < 0 || op.tab >= pd.n_tab)) throw std::runtime_error("emu: bad table index");
            int consumed = 0;
            if ((op.code == OC_REALPH4 || op.code == OC_TWHAD4) && oi + 3 >= pass.ops.size()) throw std::runtime_error("emu: truncated macro-op");
            for (int tid = 0; tid < NTHREADS; ++tid) consumed = apply_op(regs[tid], &op, op.code, op.flags, ctx[tid], pd.tables, pd.n_tab, wcs.data());
            if ((op.code == OC_REALPH4 || op.code == OC_TWHAD4) && consumed != 3) throw std::runtime_error("emu: macro-op did not run");
            for (int e = 1; e <= consumed; ++e) if (pass.ops[oi + e].group != cur) throw std::runtime_error("emu: macro-op crosses a stage");
            oi += (size_t)consumed;
        }
        for (int tid = 0; tid < NTHREADS; ++tid) flush_phase(regs[tid], ctx[tid]);
        if (cur != pd.io_out || (cur != IO_GROUP && cur != 1)) throw std::runtime_error("emu: pass does not end in its store layout");
        for (int tid = 0; tid < NTHREADS; ++tid)
            for (int j = 0; j < NREG; ++j) {
                const uint64_t i = base + tile_offset(pd, stage_idx(cur, tid, j));
                if (!pd.remap_st) amp[i] = regs[tid][j];
                else const_cast<cplx*>(pd.remap_src[remap_sel(pd, i)])[remap_index(pd, i)] = regs[tid][j];   // tile_kernel.cuh tile_store
            }
    }
}

static void run_simple(cplx* amp, int n_local, uint64_t rank_bits, const HostGate& g) {
    const uint64_t n = 1ull << n_local;
    const int t = g.target(), c = g.control();
    if (t < n_local) {
        for (uint64_t k = 0; k < n / 2; ++k) {
            const uint64_t i0 = ((k >> t) << (t + 1)) | (k & ((1ull << t) - 1)), i1 = i0 | (1ull << t);
            if (c >= 0 && !(((i0 | rank_bits) >> c) & 1ull)) continue;
            const cplx x = amp[i0], y = amp[i1];
            amp[i0] = cplx{x.x * g.m[0] - x.y * g.m[1] + y.x * g.m[2] - y.y * g.m[3], x.x * g.m[1] + x.y * g.m[0] + y.x * g.m[3] + y.y * g.m[2]};
            amp[i1] = cplx{x.x * g.m[4] - x.y * g.m[5] + y.x * g.m[6] - y.y * g.m[7], x.x * g.m[5] + x.y * g.m[4] + y.x * g.m[7] + y.y * g.m[6]};
        }
    } else {
        const int bit = (int)((rank_bits >> t) & 1ull);
        for (uint64_t i = 0; i < n; ++i) {
            if (c >= 0 && !(((i | rank_bits) >> c) & 1ull)) continue;
            amp[i] = cmul(amp[i], bit ? g.m[6] : g.m[0], bit ? g.m[7] : g.m[1]);
        }
    }
}

static std::vector<HostGate> conv(const dvd_gate* gates, int64_t n) {
    std::vector<HostGate> v;
    for (int64_t i = 0; i < n; ++i) v.push_back(make_gate(gates[i].target, gates[i].control, gates[i].m, (int)i));
    return v;
}

extern "C" {

// Whole state on `world` emulated ranks (world = 1: single GPU).  state: 2^n interleaved complex128,
// rank r's chunk is the contiguous slice r.  Returns 0, or -1 on planner error (message via emu_error).
static std::string g_err;
const char* emu_error() { return g_err.c_str(); }

// emu_run_sparse: the same, replaying the engine's support tracking: `state` must hold |0..0> as the engine's
// reset leaves it (amplitude 0 of every chunk stored, everything else arbitrary -- the tests fill it with NaN);
// implied zeros are materialised at the end, before every global swap and before the one-gate path.
static bool g_track_support = false;
static bool g_fused = true;      // global<->local swaps ride on the next pass's load (engine default when memory allows)
void emu_set_fused(int on) { g_fused = on != 0; }
static int g_store = 2;          // planner.h DistPlan: 0 = swaps ride on loads only, 1 = the layout restore rides on the last pass's store, 2 = every round where it can
static int g_defer = -1;         // tail-deferral threshold of the distributed schedule (-1: the planner picks)
static int g_last_defer = 0, g_last_store = 0;
void emu_set_store(int mode) { g_store = mode; }
void emu_set_defer(int th) { g_defer = th; }
int emu_last_defer() { return g_last_defer; }     // threshold of the last emu_run's schedule
int emu_last_store() { return g_last_store; }     // swap rounds of that schedule that rode on a store
int emu_run(int n_qubits, int world, const dvd_gate* gates, int64_t n_gates, int fuse, double* state, int64_t* stats /*[4]*/);
int emu_run_sparse(int n_qubits, int world, const dvd_gate* gates, int64_t n_gates, int fuse, double* state, int64_t* stats) {
    g_track_support = true;
    const int rc = emu_run(n_qubits, world, gates, n_gates, fuse, state, stats);
    g_track_support = false;
    return rc;
}

int emu_run(int n_qubits, int world, const dvd_gate* gates, int64_t n_gates, int fuse, double* state, int64_t* stats /*[4]*/) {
    try {
        int g = 0; while ((1 << g) < world) ++g;
        const int n_local = n_qubits - g;
        const uint64_t chunk = 1ull << n_local;
        cplx* amp = reinterpret_cast<cplx*>(state);
        std::vector<int> perm(n_qubits);
        for (int q = 0; q < n_qubits; ++q) perm[q] = q;
        std::vector<DistStep> steps;
        std::vector<HostGate> hg = conv(gates, n_gates);
        if (fuse && n_local >= TILE_BITS) hg = fuse_diagonal_runs(hg);
        PlanOptions opt;
        if (const char* e = getenv("DVD_RELABEL")) opt.relabel = atoi(e) != 0;   // experimental tile relabelling
        std::vector<std::vector<Pass>> plans;      // engine.cu flush_impl: the tuned schedule comes with its pass plans
        std::vector<std::vector<std::pair<int, int>>> store;
        g_last_defer = 0; g_last_store = 0;
        if (world > 1) {
            opt.defer_max_ops = g_defer;
            // engine.cu flush_impl: the schedule knows which local qubits are still |0> (every rank has the same mask)
            DistPlan dp = plan_distributed_tuned(hg, n_qubits, n_local, perm, true, g_fused ? g_store : 0, opt,
                                                 (g_track_support && n_local >= TILE_BITS) ? (chunk - 1) : 0);
            steps = std::move(dp.steps); plans = std::move(dp.plans); store = std::move(dp.store);
            g_last_defer = dp.defer_max_ops; g_last_store = dp.n_store;
        } else { DistStep st; st.kind = DistStep::LOCAL_GATES; st.gates = hg; steps.push_back(st); }
        for (int q = 0; q < n_qubits; ++q) if (perm[q] != q) throw std::runtime_error("emu: layout not restored");
        int64_t n_pass = 0, n_swap = 0, n_switch = 0, n_ops = 0;
        // per-rank support masks (engine.cu: dvd_state::support); dense when tracking is off
        const uint64_t local_mask = chunk - 1;
        std::vector<uint64_t> support(world, (g_track_support && n_local >= TILE_BITS) ? ~local_mask : ~0ull);
        auto materialize = [&](int r) {
            const uint64_t zm = ~support[r] & local_mask;
            if (zm) for (uint64_t i = 0; i < chunk; ++i) if (i & zm) amp[(uint64_t)r * chunk + i] = cplx{0.0, 0.0};
            support[r] = ~0ull;
        };
        // fused remap as in the engine: disjoint swaps wait for the next pass and ride on its load (out of place)
        std::vector<std::pair<int, int>> remap;
        std::vector<cplx> snapshot;
        std::unique_ptr<Pass> ident;
        auto fused_pass = [&](const Pass& p) {
            snapshot.assign(amp, amp + ((uint64_t)world << n_local));
            uint64_t lmask = 0;
            { RemapPlan rp0; compose_remap(remap, n_local, 0, &rp0); if (!rp0.on) remap.clear(); lmask = rp0.lmask; }
            for (int r = 0; r < world; ++r) {
                RemapEmu rm; rm.pairs = remap; rm.snapshot = snapshot.data(); rm.rank = r; rm.chunk = chunk;
                run_pass(amp + (uint64_t)r * chunk, p, (uint64_t)r << n_local, ~support[r] & local_mask, &rm);
                support[r] |= lmask | p.touch_mask;
            }
            remap.clear();
        };
        auto flush_remap = [&]() {
            if (remap.empty()) return;
            if (!ident) ident.reset(new Pass(make_identity_pass(n_local)));
            ++n_pass;
            fused_pass(*ident);
        };
        for (size_t si = 0; si < steps.size(); ++si) {
            DistStep& st = steps[si];
            if (st.kind == DistStep::GLOBAL_SWAP && g_fused && n_local >= TILE_BITS) {
                ++n_swap;
                if (st.gq < n_local || st.gq >= n_qubits || st.lq < 0 || st.lq >= n_local) throw std::runtime_error("emu: bad swap");
                remap.push_back({st.gq, st.lq});
                RemapPlan probe;
                if (!compose_remap(remap, n_local, 0, &probe)) { remap.pop_back(); flush_remap(); remap.push_back({st.gq, st.lq}); }
                continue;
            }
            if (st.kind == DistStep::GLOBAL_SWAP) {
                ++n_swap;
                const int j = st.gq - n_local;
                if (st.gq < n_local || st.gq >= n_qubits || st.lq < 0 || st.lq >= n_local) throw std::runtime_error("emu: bad swap");
                for (int r = 0; r < world; ++r) materialize(r);
                for (int r = 0; r < world; ++r) {
                    const int b = (r >> j) & 1, partner = r ^ (1 << j);
                    if (partner < r) continue;
                    // rank r sends its half with bit lq == 1-b; partner sends its half with bit lq == b
                    for (uint64_t h = 0; h < chunk / 2; ++h) {
                        cplx& mine = amp[(uint64_t)r * chunk + half_index(h, st.lq, 1 - b)];
                        cplx& theirs = amp[(uint64_t)partner * chunk + half_index(h, st.lq, b)];
                        std::swap(mine, theirs);
                    }
                }
                continue;
            }
            if (n_local >= TILE_BITS) {
                if (world == 1) opt.zero_mask = ~support[0] & local_mask;      // engine.cu flush_impl
                std::vector<Pass> passes = world > 1 ? std::move(plans[si]) : plan_local(st.gates, n_local, n_qubits, opt);
                for (auto& p : passes) { ++n_pass; n_switch += p.n_switches; n_ops += (int64_t)p.ops.size(); }
                size_t first = 0, end = passes.size();
                if (!remap.empty() && !passes.empty()) { fused_pass(passes[0]); first = 1; }
                const bool store_here = si < store.size() && !store[si].empty();
                const std::vector<std::pair<int, int>>& store_swaps = store_here ? store[si] : std::vector<std::pair<int, int>>();
                if (store_here) { if (end <= first) throw std::runtime_error("emu: no pass left for the store-side remap"); --end; }
                for (int r = 0; r < world; ++r)
                    for (size_t pi = first; pi < end; ++pi) {
                        run_pass(amp + (uint64_t)r * chunk, passes[pi], (uint64_t)r << n_local, ~support[r] & local_mask);
                        support[r] |= passes[pi].touch_mask;
                    }
                if (store_here) {
                    // engine.cu run_pass, store side: implied zeros are stored first, every rank writes the other buffers
                    for (int r = 0; r < world; ++r) materialize(r);
                    std::vector<cplx> out((uint64_t)world << n_local, cplx{NAN, NAN});
                    for (int r = 0; r < world; ++r) {
                        RemapEmu rm; rm.pairs = store_swaps; rm.rank = r; rm.chunk = chunk; rm.store_all = out.data();
                        run_pass(amp + (uint64_t)r * chunk, passes[end], (uint64_t)r << n_local, 0, &rm);
                    }
                    std::memcpy(amp, out.data(), out.size() * sizeof(cplx));
                    for (auto& sw : store_swaps) n_swap += sw.first >= n_local;
                }
            } else {
                for (int r = 0; r < world; ++r) materialize(r);
                for (int r = 0; r < world; ++r)
                    for (auto& gte : st.gates) run_simple(amp + (uint64_t)r * chunk, n_local, (uint64_t)r << n_local, gte);
                n_ops += (int64_t)st.gates.size();
            }
        }
        flush_remap();
        for (int r = 0; r < world; ++r) materialize(r);
        if (stats) { stats[0] = n_pass; stats[1] = n_swap; stats[2] = n_switch; stats[3] = n_ops; }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// shared-memory bank check of the three stage layouts: returns the worst conflict degree of a
// quarter-warp (8 lanes x 16 B) over all stages and registers; 1 = conflict free.
int emu_max_bank_conflict() {
    int worst = 0;
    for (int g = 0; g < NGROUPS; ++g)
        for (int j = 0; j < NREG; ++j)
            for (int q0 = 0; q0 < NTHREADS; q0 += 8) {
                int cnt[8] = {0};
                for (int l = 0; l < 8; ++l) cnt[smem_slot(stage_idx(g, q0 + l, j)) & 7]++;
                for (int b = 0; b < 8; ++b) worst = cnt[b] > worst ? cnt[b] : worst;
            }
    return worst;
}

}  // extern "C"

// Planning only (no state): pass count of every step of the distributed schedule, as the engine would run it.
// out: [n_steps, then per step: kind, gq, lq, n_gates, n_passes, then n_ops per pass]
extern "C" int64_t emu_plan_only(int n_qubits, int world, const dvd_gate* gates, int64_t n_gates, int32_t* out, int64_t cap) {
    try {
        int g = 0; while ((1 << g) < world) ++g;
        const int n_local = n_qubits - g;
        std::vector<int> perm(n_qubits);
        for (int q = 0; q < n_qubits; ++q) perm[q] = q;
        std::vector<HostGate> hg = fuse_diagonal_runs(conv(gates, n_gates));
        std::vector<DistStep> steps;
        PlanOptions opt;
        if (const char* e = getenv("DVD_RELABEL")) opt.relabel = atoi(e) != 0;
        DistPlan dp;
        if (world > 1) {      // (rounds that ride on a store are shown behind their step, as kind 3 / 4)
            opt.defer_max_ops = g_defer;
            dp = plan_distributed_tuned(hg, n_qubits, n_local, perm, true, g_fused ? g_store : 0, opt);
            for (size_t i = 0; i < dp.steps.size(); ++i) {
                steps.push_back(dp.steps[i]);
                for (auto& sw : dp.store[i]) { DistStep st; st.kind = (DistStep::Kind)(sw.first >= n_local ? 3 : 4); st.gq = sw.first; st.lq = sw.second; steps.push_back(st); }
            }
        } else { DistStep st; st.kind = DistStep::LOCAL_GATES; st.gates = hg; steps.push_back(st); }
        std::vector<int32_t> v;
        v.push_back((int32_t)steps.size());
        for (auto& st : steps) {
            v.push_back((int32_t)st.kind); v.push_back(st.gq); v.push_back(st.lq); v.push_back((int32_t)st.gates.size());
            if (st.kind == DistStep::LOCAL_GATES) {
                std::vector<Pass> passes = plan_local(st.gates, n_local, n_qubits, opt);
                v.push_back((int32_t)passes.size());
                for (auto& p : passes) v.push_back((int32_t)p.ops.size());
                if (getenv("DVD_PLAN_DUMP"))
                    for (auto& p : passes) {
                        fprintf(stderr, "  pass tile [");
                        for (int k = 0; k < TILE_BITS; ++k) fprintf(stderr, "%d ", p.desc.tile_q[k]);
                        fprintf(stderr, "] ops:");
                        for (auto& op : p.ops) fprintf(stderr, " %d(g%d)", op.code, op.gate_idx);
                        fprintf(stderr, "\n");
                    }
            } else v.push_back(0);
        }
        if ((int64_t)v.size() > cap) return -(int64_t)v.size();
        std::memcpy(out, v.data(), v.size() * sizeof(int32_t));
        return (int64_t)v.size();
    } catch (const std::exception& e) {
        g_err = e.what();
        return INT64_MIN;
    }
}

// Planning only: HBM traffic (in full passes over the local chunk) of the gate passes of a workload's schedule, from a
// reset (from_reset != 0: every local qubit still |0>) or on a dense state; *n_passes = its pass count.
extern "C" double emu_plan_traffic(int n_qubits, int world, const dvd_gate* gates, int64_t n_gates, int from_reset, int* n_passes) {
    try {
        int g = 0; while ((1 << g) < world) ++g;
        const int n_local = n_qubits - g;
        std::vector<int> perm(n_qubits);
        for (int q = 0; q < n_qubits; ++q) perm[q] = q;
        std::vector<HostGate> hg = fuse_diagonal_runs(conv(gates, n_gates));
        PlanOptions opt;
        if (const char* e = getenv("DVD_PLAN_PORTFOLIO")) opt.portfolio = atoi(e) != 0;
        uint64_t zm = from_reset ? (1ull << n_local) - 1 : 0;
        double traffic = 0.0;
        int np = 0;
        if (world > 1) {
            DistPlan dp = plan_distributed_tuned(hg, n_qubits, n_local, perm, true, g_fused ? g_store : 0, opt, zm);
            for (size_t i = 0; i < dp.steps.size(); ++i) {
                if (dp.steps[i].kind == DistStep::LOCAL_GATES) { traffic += plan_traffic(dp.plans[i], &zm); np += (int)dp.plans[i].size(); if (!dp.store[i].empty()) zm = 0; }
                else if (dp.steps[i].lq >= 0 && dp.steps[i].lq < n_local) zm &= ~(1ull << dp.steps[i].lq);
            }
        } else {
            opt.zero_mask = zm;
            std::vector<Pass> passes = plan_local(hg, n_local, n_qubits, opt);
            traffic = plan_traffic(passes, &zm);
            np = (int)passes.size();
        }
        if (n_passes) *n_passes = np;
        return traffic;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1.0;
    }
}

// Record / replay of the planners' portfolio choices (planner.h: PlanChoices): searches a schedule for `gates` (recording
// the winners), replays them on `gates2` (the same circuit with other angles; null = the same gates), and compares the
// replayed schedule op for op with the one a fresh search finds for `gates2`.  Returns 0 if they are identical, > 0
// otherwise, < 0 on a planner error; ms[0] / ms[1] = host time of the first search and of the replaying run.
extern "C" int emu_plan_replay_check(int n_qubits, int world, const dvd_gate* gates, const dvd_gate* gates2, int64_t n_gates, int from_reset, double* ms) {
    try {
        int g = 0; while ((1 << g) < world) ++g;
        const int n_local = n_qubits - g;
        const std::vector<HostGate> raw_a = conv(gates, n_gates), raw_b = conv(gates2 ? gates2 : gates, n_gates);
        const std::vector<HostGate> hg_a = fuse_diagonal_runs(raw_a), hg_b = fuse_diagonal_runs(raw_b);
        const uint64_t zm = from_reset ? (1ull << n_local) - 1 : 0;
        ChoiceMemoTable table(2), other(2);       // engine.cu flush_impl: dvd_state::memos
        {   // unrelated structures in between must not disturb the entry (and the table must not grow past its capacity)
            std::vector<HostGate> x = raw_a;
            for (int k = 0; k < 3; ++k) { x.pop_back(); other.begin(x, zm); }
            if (other.size() != 2) return 1001;
        }
        PlanChoices ch;
        std::vector<std::vector<Pass>> plans[3];
        std::vector<DistStep> steps[3];
        for (int run = 0; run < 3; ++run) {      // 0: search on A (record), 1: replay on B, 2: search on B
            const std::vector<HostGate>& hg = run == 0 ? hg_a : hg_b;
            PlanOptions opt;
            PlanChoices* c = run == 2 ? other.begin(raw_b, zm << 1 | 1) : table.begin(run == 0 ? raw_a : raw_b, zm << 1 | 1);
            if (c->replay != (run == 1)) return 1002;      // same structure, other angles: known; a fresh table: not
            opt.choices = c;
            if (run == 1) ch = *c;
            std::vector<int> perm(n_qubits);
            for (int q = 0; q < n_qubits; ++q) perm[q] = q;
            struct timespec t0, t1;
            clock_gettime(CLOCK_MONOTONIC, &t0);
            if (world > 1) {
                DistPlan dp = plan_distributed_tuned(hg, n_qubits, n_local, perm, true, g_fused ? g_store : 0, opt, zm);
                plans[run] = std::move(dp.plans); steps[run] = std::move(dp.steps);
            } else {
                opt.zero_mask = zm;
                plans[run].push_back(plan_local(hg, n_local, n_qubits, opt));
            }
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if (ms && run < 2) ms[run] = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
        }
        if (ch.tape.empty() || !ch.replay) return 1000;
        if (plans[2].size() != plans[1].size() || steps[2].size() != steps[1].size()) return 1;
        for (size_t i = 0; i < steps[2].size(); ++i)
            if (steps[2][i].kind != steps[1][i].kind || steps[2][i].gq != steps[1][i].gq || steps[2][i].lq != steps[1][i].lq) return 2;
        for (size_t i = 0; i < plans[2].size(); ++i) {
            if (plans[2][i].size() != plans[1][i].size()) return 3;
            for (size_t k = 0; k < plans[2][i].size(); ++k) {
                const Pass& a = plans[2][i][k]; const Pass& b = plans[1][i][k];
                if (a.ops.size() != b.ops.size() || std::memcmp(a.desc.tile_q, b.desc.tile_q, sizeof a.desc.tile_q)) return 4;
                for (size_t o = 0; o < a.ops.size(); ++o)
                    if (a.ops[o].code != b.ops[o].code || std::memcmp(a.ops[o].m, b.ops[o].m, sizeof a.ops[o].m)) return 5;
            }
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
