// planner.cpp -- see planner.h.  Pure host C++.
#include "planner.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <stdexcept>
#include <utility>

namespace dvd {

void classify_gate(const double m[8], int32_t* kind, int8_t* d0_is_one) {
    *d0_is_one = 0;
    const bool off_zero = m[2] == 0.0 && m[3] == 0.0 && m[4] == 0.0 && m[5] == 0.0;
    const bool diag_zero = m[0] == 0.0 && m[1] == 0.0 && m[6] == 0.0 && m[7] == 0.0;
    const bool all_real = m[1] == 0.0 && m[3] == 0.0 && m[5] == 0.0 && m[7] == 0.0;
    if (off_zero) {
        *kind = K_DIAG;
        *d0_is_one = (m[0] == 1.0 && m[1] == 0.0) ? 1 : 0;
    } else if (diag_zero) {
        *kind = (m[2] == 1.0 && m[3] == 0.0 && m[4] == 1.0 && m[5] == 0.0) ? K_SWAP : K_ANTIDIAG;
    } else if (all_real) {
        *kind = K_REAL;
    } else if (m[1] == 0.0 && m[7] == 0.0 && m[2] == 0.0 && m[4] == 0.0) {
        *kind = K_RXLIKE;
    } else {
        *kind = K_GENERAL;
    }
}

HostGate make_gate(int target, int control, const double m[8], int gate_idx) {
    HostGate g;
    g.tmask = 1ull << target;
    g.cmask = control >= 0 ? 1ull << control : 0;
    std::memcpy(g.m, m, sizeof g.m);
    g.gate_idx = gate_idx;
    g.diag = is_diagonal(m);
    return g;
}

// ---------------------------------------------------------------------------------------------------
// Level 0: diagonal-run fusion.
//
// Inside a run made only of diagonal gates and CNOTs the state is |x> -> phase(x) |A x> with A a
// GF(2)-linear map.  rows[q] tracks qubit q's current value as a parity of the run's input bits, so
// a diagonal gate met in the middle of the run is a phase on parity(x & rows[target]) (controlled by
// parity(x & rows[control])).  Whenever A is back to the identity the run so far equals the product
// of the collected parity phases -- all diagonal, all commuting, no amplitude moves.
// ---------------------------------------------------------------------------------------------------
std::vector<HostGate> fuse_diagonal_runs(const std::vector<HostGate>& gates) {
    std::vector<HostGate> out;
    out.reserve(gates.size());
    uint64_t rows[64];
    for (int q = 0; q < 64; ++q) rows[q] = 1ull << q;
    int nonid = 0;
    std::vector<HostGate> items, terms;
    size_t ck_items = 0, ck_terms = 0;
    bool ck_has_cnot = false, has_cnot = false;

    auto xor_rows = [&](uint64_t mask) {
        uint64_t r = 0;
        while (mask) { const int q = __builtin_ctzll(mask); mask &= mask - 1; r ^= rows[q]; }
        return r;
    };
    auto end_run = [&]() {
        if (items.empty()) return;
        if (ck_items > 0 && ck_has_cnot) {
            // fused prefix: merge terms with identical masks
            std::map<std::pair<uint64_t, uint64_t>, size_t> where;
            const size_t first = out.size();
            for (size_t i = 0; i < ck_terms; ++i) {
                const HostGate& t = terms[i];
                auto key = std::make_pair(t.tmask, t.cmask);
                auto it = where.find(key);
                if (it == where.end()) {
                    where[key] = out.size();
                    HostGate g = t;
                    g.gate_idx = -1;
                    out.push_back(g);
                } else {
                    HostGate& g = out[it->second];
                    const double a0 = g.m[0], b0 = g.m[1], a1 = g.m[6], b1 = g.m[7];
                    g.m[0] = a0 * t.m[0] - b0 * t.m[1]; g.m[1] = a0 * t.m[1] + b0 * t.m[0];
                    g.m[6] = a1 * t.m[6] - b1 * t.m[7]; g.m[7] = a1 * t.m[7] + b1 * t.m[6];
                }
            }
            (void)first;
            for (size_t i = ck_items; i < items.size(); ++i) out.push_back(items[i]);
        } else {
            for (const HostGate& g : items) out.push_back(g);
        }
        items.clear(); terms.clear();
        ck_items = ck_terms = 0; ck_has_cnot = has_cnot = false;
        if (nonid) { for (int q = 0; q < 64; ++q) rows[q] = 1ull << q; nonid = 0; }
    };

    for (const HostGate& g : gates) {
        int32_t kind; int8_t d0;
        classify_gate(g.m, &kind, &d0);
        const bool is_cnot = !g.diag && kind == K_SWAP && g.cmask != 0 && (g.cmask & (g.cmask - 1)) == 0;
        if (g.diag) {
            HostGate t = g;
            t.tmask = xor_rows(g.tmask);
            t.cmask = g.cmask ? xor_rows(g.cmask) : 0;
            terms.push_back(t);
            items.push_back(g);
        } else if (is_cnot) {
            const int t = g.target(), c = g.control();
            const bool was = rows[t] != (1ull << t);
            rows[t] ^= rows[c];
            const bool is = rows[t] != (1ull << t);
            nonid += (int)is - (int)was;
            items.push_back(g);
            has_cnot = true;
        } else {
            end_run();
            out.push_back(g);
            continue;
        }
        if (nonid == 0) { ck_items = items.size(); ck_terms = terms.size(); ck_has_cnot = has_cnot; }
    }
    end_run();
    return out;
}

namespace {

// Commutation bookkeeping for "pull a gate in front of the gates that were skipped".
// A qubit is used diagonally by a gate when it is in a control mask or in the mask of a diagonal
// gate, and non-diagonally when it is the target of a non-diagonal gate.  Two gates commute if on
// every shared qubit both uses are diagonal.
struct Blocked {
    uint64_t x = 0;  // qubits with a skipped non-diagonal use
    uint64_t z = 0;  // qubits with a skipped diagonal use
    bool can_pass(const HostGate& g) const {
        if (g.diag) return !(x & (g.tmask | g.cmask));
        return !((x | z) & g.tmask) && !(x & g.cmask);
    }
    void skip(const HostGate& g) {
        if (g.diag) z |= g.tmask | g.cmask;
        else { x |= g.tmask; z |= g.cmask; }
    }
};

}  // namespace

std::vector<Pass> plan_local(const std::vector<HostGate>& gates, int n_local, int n_total,
                             const PlanOptions& opt) {
    if (n_local < TILE_BITS) throw std::runtime_error("plan_local: n_local < TILE_BITS");
    if (n_total > 62) throw std::runtime_error("plan_local: too many qubits");
    std::vector<Pass> passes;
    const int G = (int)gates.size();
    const uint64_t all = n_total >= 64 ? ~0ull : ((1ull << n_total) - 1);
    for (int i = 0; i < G; ++i) {
        const HostGate& g = gates[i];
        if ((g.tmask | g.cmask) & ~all) throw std::runtime_error("plan_local: bad qubit index");
        if (!g.diag) {
            if (g.tmask == 0 || (g.tmask & (g.tmask - 1))) throw std::runtime_error("plan_local: bad target");
            if (g.cmask & (g.cmask - 1)) throw std::runtime_error("plan_local: multi-qubit control on a non-diagonal gate");
            if (g.cmask & g.tmask) throw std::runtime_error("plan_local: control == target");
            if (g.target() >= n_local) throw std::runtime_error("plan_local: non-diagonal gate on a rank-index qubit");
        }
    }
    std::vector<int> pending(G);
    for (int i = 0; i < G; ++i) pending[i] = i;
    const int min_low = std::min(opt.min_low, TILE_BITS);

    while (!pending.empty()) {
        // ---- pass level: grow the tile greedily, take everything that commutes to the front ----
        uint64_t tile = (1ull << min_low) - 1;
        int tile_n = min_low;
        Blocked blk;
        std::vector<int> taken, rest;
        const int limit = std::min<int>((int)pending.size(), opt.window);
        for (int k = 0; k < (int)pending.size(); ++k) {
            const int gi = pending[k];
            const HostGate& g = gates[gi];
            bool ok = k < limit && (int)taken.size() < opt.max_ops_per_pass && blk.can_pass(g);
            if (ok && !g.diag && !(tile & g.tmask)) {
                if (tile_n < TILE_BITS) { tile |= g.tmask; ++tile_n; }
                else ok = false;
            }
            if (ok) taken.push_back(gi);
            else { blk.skip(g); rest.push_back(gi); }
        }
        if (taken.empty()) throw std::runtime_error("plan_local: no progress");
        // pad the tile with the lowest unused local qubits (keeps segments long)
        for (int q = 0; q < n_local && tile_n < TILE_BITS; ++q)
            if (!((tile >> q) & 1)) { tile |= 1ull << q; ++tile_n; }

        // ---- tile positions: pinned low run, then by first non-diagonal use ----------------------
        Pass pass;
        std::memset(&pass.desc, 0, sizeof(pass.desc));
        pass.desc.n_local = n_local;
        std::vector<int> order;  // qubits by first non-diagonal target use
        uint64_t seen = (1ull << min_low) - 1;
        for (int gi : taken) {
            const HostGate& g = gates[gi];
            if (!g.diag && !(seen & g.tmask)) { seen |= g.tmask; order.push_back(g.target()); }
        }
        for (int q = 0; q < n_local; ++q)
            if (((tile >> q) & 1) && !((seen >> q) & 1)) { seen |= 1ull << q; order.push_back(q); }
        int pos_of[64];
        for (int q = 0; q < 64; ++q) pos_of[q] = -1;
        for (int p = 0; p < min_low; ++p) { pass.desc.tile_q[p] = p; pos_of[p] = p; }
        {
            // free positions, in the order they are handed out: IO group first (no switch needed
            // for the first gates), then the middle groups downwards, then the rest of group 0.
            std::vector<int> free_pos;
            for (int g = NGROUPS - 1; g >= 0; --g)
                for (int p = g * REG_BITS; p < (g + 1) * REG_BITS; ++p)
                    if (p >= min_low) free_pos.push_back(p);
            if (order.size() != free_pos.size()) throw std::runtime_error("plan_local: tile size");
            for (size_t i = 0; i < order.size(); ++i) {
                pass.desc.tile_q[free_pos[i]] = order[i];
                pos_of[order[i]] = free_pos[i];
            }
        }
        for (int p = 0; p < TILE_BITS; ++p) pass.desc.sorted_q[p] = pass.desc.tile_q[p];
        std::sort(pass.desc.sorted_q, pass.desc.sorted_q + TILE_BITS);

        // ---- stage level: sweep per register group ---------------------------------------------------
        std::vector<int> remaining = taken;
        int cur = IO_GROUP;
        while (!remaining.empty()) {
            // keep `cur` if the first remaining op can run there, else move to its group
            {
                const HostGate& g0 = gates[remaining[0]];
                if (!g0.diag) {
                    const int grp = pos_of[g0.target()] / REG_BITS;
                    if (grp != cur) { cur = grp; ++pass.n_switches; }
                }
            }
            uint64_t regphys = 0;          // physical qubits living in registers in this stage
            int regq[REG_BITS];
            for (int k = 0; k < REG_BITS; ++k) { regq[k] = pass.desc.tile_q[cur * REG_BITS + k]; regphys |= 1ull << regq[k]; }
            auto reg_mask = [&](uint64_t mask) {
                uint8_t r = 0;
                for (int k = 0; k < REG_BITS; ++k) if ((mask >> regq[k]) & 1) r |= (uint8_t)(1 << k);
                return r;
            };
            Blocked b2;
            std::vector<int> rem2;
            for (int gi : remaining) {
                const HostGate& g = gates[gi];
                const bool ok = b2.can_pass(g) && (g.diag || pos_of[g.target()] / REG_BITS == cur);
                if (!ok) { b2.skip(g); rem2.push_back(gi); continue; }
                DevOp op;
                std::memset(&op, 0, sizeof(op));
                std::memcpy(op.m, g.m, sizeof(op.m));
                classify_gate(g.m, &op.kind, &op.d0_is_one);
                op.gate_idx = g.gate_idx;
                op.has_ctrl = g.cmask != 0;
                op.cregm = reg_mask(g.cmask);
                op.cmask = g.cmask & ~regphys;
                if (g.diag) {
                    op.group = -1;
                    op.treg = -1;
                    op.tregm = reg_mask(g.tmask);
                    op.tmask = g.tmask & ~regphys;
                } else {
                    op.group = (int8_t)cur;
                    op.treg = (int8_t)(pos_of[g.target()] % REG_BITS);
                    op.tregm = 0;
                    op.tmask = 0;
                }
                pass.ops.push_back(op);
            }
            remaining.swap(rem2);
        }
        if (cur != IO_GROUP) ++pass.n_switches;
        pass.desc.n_ops = (int)pass.ops.size();
        passes.push_back(std::move(pass));
        pending.swap(rest);
    }
    return passes;
}

// ---------------------------------------------------------------------------------------------------
std::vector<DistStep> plan_distributed(const std::vector<HostGate>& gates, int n_total, int n_local,
                                       std::vector<int>& perm, bool restore_identity) {
    std::vector<DistStep> steps;
    const int G = (int)gates.size();
    std::vector<int> inv(n_total);  // physical -> logical
    for (int q = 0; q < n_total; ++q) inv[perm[q]] = q;

    auto local_step = [&]() -> DistStep& {
        if (steps.empty() || steps.back().kind != DistStep::LOCAL_GATES) {
            DistStep s; s.kind = DistStep::LOCAL_GATES; steps.push_back(std::move(s));
        }
        return steps.back();
    };
    auto emit_swap = [&](int gq, int lq) {
        DistStep s; s.kind = DistStep::GLOBAL_SWAP; s.gq = gq; s.lq = lq;
        steps.push_back(std::move(s));
        const int a = inv[gq], b = inv[lq];
        perm[a] = lq; perm[b] = gq; inv[gq] = b; inv[lq] = a;
    };
    auto emit_local_cnot = [&](int pc, int pt) {
        const double x[8] = {0, 0, 1, 0, 1, 0, 0, 0};
        local_step().gates.push_back(make_gate(pt, pc, x, -1));
    };
    auto map_mask = [&](uint64_t mask) {
        uint64_t r = 0;
        while (mask) { const int q = __builtin_ctzll(mask); mask &= mask - 1; r |= 1ull << perm[q]; }
        return r;
    };

    for (int i = 0; i < G; ++i) {
        const HostGate& g = gates[i];
        if (!g.diag && perm[g.target()] >= n_local) {
            // evict the local qubit whose next non-diagonal use is farthest away (Belady)
            std::vector<int> next_use(n_total, G + 1);
            for (int k = G - 1; k > i; --k)
                if (!gates[k].diag) next_use[gates[k].target()] = k;
            int victim = -1, best = -1;
            for (int p = n_local - 1; p >= 0; --p) {   // ties: prefer high local positions
                const int lq = inv[p];
                if (lq == g.control()) continue;        // keep this gate's control local
                if (next_use[lq] > best) { best = next_use[lq]; victim = p; }
            }
            emit_swap(perm[g.target()], victim);
        }
        HostGate pg = g;
        pg.tmask = map_mask(g.tmask);
        pg.cmask = map_mask(g.cmask);
        local_step().gates.push_back(pg);
    }

    if (restore_identity) {
        // 1. rank-index positions
        for (int gp = n_local; gp < n_total; ++gp) {
            if (perm[gp] == gp) continue;
            int x = perm[gp];                 // where logical gp lives now
            if (x >= n_local) {               // on another rank-index position: bounce through a local one
                const int l = n_local - 1;
                emit_swap(x, l);
                x = l;
            }
            emit_swap(gp, x);
        }
        // 2. local positions: transpositions as CNOT triples
        for (int p = 0; p < n_local; ++p) {
            if (perm[p] == p) continue;
            const int x = perm[p];            // logical p lives at local position x
            emit_local_cnot(p, x); emit_local_cnot(x, p); emit_local_cnot(p, x);
            const int other = inv[p];
            perm[p] = p; perm[other] = x; inv[p] = p; inv[x] = other;
        }
    }
    return steps;
}

}  // namespace dvd
