// planner.cpp -- see planner.h.  Pure host C++.
#include "planner.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>

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

namespace {

// Commutation bookkeeping for "pull a gate in front of the gates that were skipped".
// A qubit is used diagonally by a gate when it is a control or the target of a diagonal gate, and
// non-diagonally when it is the target of a non-diagonal gate.  Two gates commute if on every
// shared qubit both uses are diagonal.
struct Blocked {
    uint64_t x = 0;  // qubits with a skipped non-diagonal use
    uint64_t z = 0;  // qubits with a skipped diagonal use
    bool can_pass(const HostGate& g, bool diag) const {
        const uint64_t tb = 1ull << g.target;
        if (diag) { if (x & tb) return false; }
        else if ((x | z) & tb) return false;
        if (g.control >= 0 && (x & (1ull << g.control))) return false;
        return true;
    }
    void skip(const HostGate& g, bool diag) {
        const uint64_t tb = 1ull << g.target;
        if (diag) z |= tb; else x |= tb;
        if (g.control >= 0) z |= 1ull << g.control;
    }
};

}  // namespace

std::vector<Pass> plan_local(const std::vector<HostGate>& gates, int n_local, int n_total,
                             const PlanOptions& opt) {
    if (n_local < TILE_BITS) throw std::runtime_error("plan_local: n_local < TILE_BITS");
    if (n_total > 62) throw std::runtime_error("plan_local: too many qubits");
    std::vector<Pass> passes;
    const int G = (int)gates.size();
    std::vector<char> diag(G);
    for (int i = 0; i < G; ++i) {
        diag[i] = is_diagonal(gates[i].m);
        if (!diag[i] && gates[i].target >= n_local)
            throw std::runtime_error("plan_local: non-diagonal gate on a rank-index qubit");
        if (gates[i].target < 0 || gates[i].target >= n_total || gates[i].control >= n_total ||
            gates[i].control == gates[i].target)
            throw std::runtime_error("plan_local: bad qubit index");
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
            bool ok = k < limit && (int)taken.size() < opt.max_ops_per_pass && blk.can_pass(g, diag[gi]);
            if (ok && !diag[gi] && !((tile >> g.target) & 1)) {
                if (tile_n < TILE_BITS) { tile |= 1ull << g.target; ++tile_n; }
                else ok = false;
            }
            if (ok) taken.push_back(gi);
            else { blk.skip(g, diag[gi]); rest.push_back(gi); }
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
            const int t = gates[gi].target;
            if (!diag[gi] && !((seen >> t) & 1)) { seen |= 1ull << t; order.push_back(t); }
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
            if ((int)order.size() != (int)free_pos.size()) throw std::runtime_error("plan_local: tile size");
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
        bool first = true;
        while (!remaining.empty()) {
            // choose the group: keep `cur` if the first remaining op can run there, else move to it
            {
                const int gi0 = remaining[0];
                if (!diag[gi0]) {
                    const int g0 = pos_of[gates[gi0].target] / REG_BITS;
                    if (g0 != cur) { cur = g0; if (!first || g0 != IO_GROUP) ++pass.n_switches; }
                }
            }
            first = false;
            Blocked b2;
            std::vector<int> rem2;
            for (int gi : remaining) {
                const HostGate& g = gates[gi];
                bool ok = b2.can_pass(g, diag[gi]) &&
                          (diag[gi] || pos_of[g.target] / REG_BITS == cur);
                if (ok) {
                    DevOp op;
                    std::memset(&op, 0, sizeof(op));
                    std::memcpy(op.m, g.m, sizeof(op.m));
                    classify_gate(g.m, &op.kind, &op.d0_is_one);
                    op.tbit = (int8_t)g.target;
                    op.tpos = (int8_t)pos_of[g.target];
                    op.group = diag[gi] ? (int8_t)-1 : (int8_t)(op.tpos / REG_BITS);
                    op.cbit = (int8_t)g.control;
                    op.cpos = g.control >= 0 ? (int8_t)pos_of[g.control] : (int8_t)-1;
                    op.gate_idx = g.gate_idx;
                    if (g.control >= 0) ++pass.n_controlled;
                    pass.ops.push_back(op);
                } else {
                    b2.skip(g, diag[gi]);
                    rem2.push_back(gi);
                }
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
        HostGate g; g.target = pt; g.control = pc; g.gate_idx = -1;
        const double x[8] = {0, 0, 1, 0, 1, 0, 0, 0};
        std::memcpy(g.m, x, sizeof(x));
        local_step().gates.push_back(g);
    };

    for (int i = 0; i < G; ++i) {
        const HostGate& g = gates[i];
        const bool d = is_diagonal(g.m);
        if (!d && perm[g.target] >= n_local) {
            // evict the local qubit whose next non-diagonal use is farthest away (Belady)
            std::vector<int> next_use(n_total, G + 1);
            for (int k = G - 1; k > i; --k)
                if (!is_diagonal(gates[k].m)) next_use[gates[k].target] = k;
            int victim = -1, best = -1;
            for (int p = n_local - 1; p >= 0; --p) {   // ties: prefer high local positions
                const int lq = inv[p];
                if (lq == g.control) continue;          // keeping the control local is not required, but cheap
                if (next_use[lq] > best) { best = next_use[lq]; victim = p; }
            }
            emit_swap(perm[g.target], victim);
        }
        HostGate pg = g;
        pg.target = perm[g.target];
        pg.control = g.control >= 0 ? perm[g.control] : -1;
        local_step().gates.push_back(pg);
    }

    if (restore_identity) {
        // 1. rank-index positions
        for (int gp = n_local; gp < n_total; ++gp) {
            if (perm[gp] == gp) continue;
            int x = perm[gp];                 // where logical gp lives now
            if (x >= n_local) {               // on another rank-index position: bounce through a local one
                // pick a local position holding a qubit that belongs to a local position
                int l = n_local - 1;
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
