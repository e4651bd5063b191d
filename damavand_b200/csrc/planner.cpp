// planner.cpp -- see planner.h.  Pure host C++.
#include "planner.h"
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <climits>
#include <cmath>
#include <complex>
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
        *kind = (m[0] == m[2] && m[0] == m[4] && m[0] == -m[6]) ? K_HADAMARD : K_REAL;
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
// Level 0a: runs of uncontrolled single-qubit gates on one qubit (H RX RY RZ ...) collapse into one 2x2.
// A run is deferred while the gates in between do not touch its qubit (they commute with it) and is
// emitted just before the first gate that does.  Runs with at most one non-diagonal gate stay as they
// are: their diagonal gates are free in the phase polynomial (or fold into the gate as K_REALPH).
std::vector<HostGate> merge_single_qubit_runs(const std::vector<HostGate>& gates) {
    using cld = std::complex<long double>;
    std::vector<HostGate> out;
    out.reserve(gates.size());
    std::vector<std::vector<HostGate>> open(64);
    auto close = [&](int q) {
        std::vector<HostGate>& run = open[q];
        if (run.empty()) return;
        int n_nondiag = 0;
        for (const HostGate& g : run) n_nondiag += g.diag ? 0 : 1;
        // one non-diagonal gate: the diagonal gates around it are free or nearly so in the phase polynomial
        // (folded into the gate as K_REALPH, or into a twiddle that is emitted anyway)
        const bool keep = run.size() == 1 || n_nondiag <= 1;
        if (keep) {
            for (const HostGate& g : run) out.push_back(g);
        } else {
            cld m[4] = {cld(1, 0), cld(0, 0), cld(0, 0), cld(1, 0)};
            for (const HostGate& g : run) {          // later gates multiply from the left
                cld a[4];
                for (int k = 0; k < 4; ++k) a[k] = cld((long double)g.m[2 * k], (long double)g.m[2 * k + 1]);
                const cld r[4] = {a[0] * m[0] + a[1] * m[2], a[0] * m[1] + a[1] * m[3],
                                  a[2] * m[0] + a[3] * m[2], a[2] * m[1] + a[3] * m[3]};
                for (int k = 0; k < 4; ++k) m[k] = r[k];
            }
            double md[8];
            for (int k = 0; k < 4; ++k) { md[2 * k] = (double)m[k].real(); md[2 * k + 1] = (double)m[k].imag(); }
            out.push_back(make_gate(q, -1, md, -1));
        }
        run.clear();
    };
    for (const HostGate& g : gates) {
        const bool single = g.cmask == 0 && g.tmask != 0 && (g.tmask & (g.tmask - 1)) == 0;
        if (single) { open[g.target()].push_back(g); continue; }
        uint64_t touched = g.tmask | g.cmask;
        while (touched) { const int q = __builtin_ctzll(touched); touched &= touched - 1; close(q); }
        out.push_back(g);
    }
    for (int q = 0; q < 64; ++q) close(q);
    return out;
}

std::vector<HostGate> fuse_diagonal_runs(const std::vector<HostGate>& gates_in) {
    const std::vector<HostGate> gates = merge_single_qubit_runs(gates_in);
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

// Macro-ops: four consecutive ops of one shape on the four register bits run under one dispatch
// (the interpreter's per-op decode costs about as much as a Hadamard).  Done last: it reorders ops.
static void fuse_macro_ops(std::vector<DevOp>& ops) {
    // a macro-op only pays on the specialised kernel variant that holds it (kernels.cu: V_LAYERED, V_FOURIER);
    // a pass that needs other op classes runs on a general variant and keeps its ops one by one
    unsigned classes = 0;
    for (const DevOp& o : ops) classes |= op_class(o.code);
    const bool allow_r = !(classes & ~(C_REAL | C_HAD | C_DIAG));
    const bool allow_t = !(classes & ~(C_HAD | C_DIAG | C_TABLE));
    for (size_t i = 0; i + 3 < ops.size(); ++i) {
        // four uncontrolled real(+phase) gates on four different register bits: they commute, sort them by bit
        bool ok = true;
        unsigned bits = 0;
        for (int k = 0; k < 4 && ok; ++k) {
            const DevOp& o = ops[i + k];
            const int kind = (o.code - OC_GATE) / 4;
            ok = o.code >= OC_GATE && o.code < OC_CGEN && (kind == K_REAL || kind == K_REALPH) && !(o.flags & F_TCTRL);
            if (ok) bits |= 1u << ((o.code - OC_GATE) % 4);
        }
        if (ok && bits == 15u && allow_r) {
            DevOp sorted[4];
            for (int k = 0; k < 4; ++k) {
                DevOp o = ops[i + k];
                const int treg = (o.code - OC_GATE) % 4;
                if ((o.code - OC_GATE) / 4 == K_REAL) { o.m[1] = 1.0; o.m[3] = 0.0; o.code = OC_GATE + 4 * K_REALPH + treg; }
                sorted[treg] = o;
            }
            sorted[0].code = OC_REALPH4;
            for (int k = 0; k < 4; ++k) ops[i + k] = sorted[k];
            i += 3;
            continue;
        }
        ok = true;
        for (int k = 0; k < 4 && ok; ++k) ok = ops[i + k].code == OC_TWHAD + k && !(ops[i + k].flags & F_TCTRL);
        if (ok && allow_t) { ops[i].code = OC_TWHAD4; i += 3; }
    }
}

void Pass::finish_tables() {
    tables.clear();
    const size_t n_tab = tab_desc.size(), byte_base = n_tab + tab_tile.size();
    tid_off_slot = byte_base + tab_bytes.size();
    tables.resize(tid_off_slot + (size_t)NGROUPS * NTHREADS / 2);
    for (size_t i = 0; i < n_tab; ++i) {
        TableDesc d = tab_desc[i];
        d.byte_off += (uint32_t)byte_base;
        std::memcpy(&tables[i], &d, sizeof d);
    }
    if (!tab_tile.empty()) std::memcpy(&tables[n_tab], tab_tile.data(), tab_tile.size() * sizeof(cplx));
    if (!tab_bytes.empty()) std::memcpy(&tables[byte_base], tab_bytes.data(), tab_bytes.size() * sizeof(cplx));
    uint64_t* off = reinterpret_cast<uint64_t*>(&tables[tid_off_slot]);
    for (int g = 0; g < NGROUPS; ++g)
        for (int tid = 0; tid < NTHREADS; ++tid) off[g * NTHREADS + tid] = tile_offset(desc, stage_idx(g, tid, 0));
    fill_cta_runs(desc);
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

// ---------------------------------------------------------------------------------------------------
// Phase polynomial of the diagonal gates of a pass.
//
// Every 1- and 2-qubit diagonal term is rewritten as monomials over the index bits,
//     K * prod_q a_q^{x_q} * prod_{q<q'} b_qq'^{x_q x_q'},
// accumulated in extended precision and emitted as late as commutation allows: a term on qubits
// {q,q'} commutes with every gate that does not TARGET q or q' non-diagonally.  While a qubit is
// thread-level its terms cost a table lookup per thread; the terms of a register qubit are forced out
// only when a non-diagonal gate is about to hit that qubit (or at the end of the pass).
// ---------------------------------------------------------------------------------------------------
namespace {

using cl = std::complex<long double>;

bool is_one(const cl& v) { return std::abs(v - cl(1, 0)) < 1e-17L; }

struct DiagAcc {
    cl K{1, 0};
    std::map<int, cl> a;
    std::map<std::pair<int, int>, cl> b;

    void mul_a(int q, const cl& f) { auto it = a.find(q); if (it == a.end()) a[q] = f; else it->second *= f; }
    void mul_b(int q, int p, const cl& f) {
        auto key = std::make_pair(std::min(q, p), std::max(q, p));
        auto it = b.find(key); if (it == b.end()) b[key] = f; else it->second *= f;
    }
    // false: not a simple 1-/2-qubit phase (caller emits the generic op)
    bool add(const HostGate& g) {
        const cl d0((long double)g.m[0], (long double)g.m[1]), d1((long double)g.m[6], (long double)g.m[7]);
        const long double n0 = std::norm(d0), n1 = std::norm(d1);
        if (!(n0 > 1e-30L) || !(n1 > 1e-30L) || !std::isfinite((double)n0) || !std::isfinite((double)n1)) return false;
        const int nt = __builtin_popcountll(g.tmask), nc = __builtin_popcountll(g.cmask);
        if (g.tmask & g.cmask) return false;
        if (nc == 0 && nt == 0) { K *= d0; return true; }
        if (nc == 0 && nt == 1) { K *= d0; mul_a(__builtin_ctzll(g.tmask), d1 / d0); return true; }
        if (nc == 0 && nt == 2) {
            const int q = __builtin_ctzll(g.tmask), p = 63 - __builtin_clzll(g.tmask);
            const cl r = d1 / d0;
            K *= d0; mul_a(q, r); mul_a(p, r); mul_b(q, p, cl(1, 0) / (r * r));
            return true;
        }
        if (nc == 1 && nt == 1) {
            const int c = __builtin_ctzll(g.cmask), t = __builtin_ctzll(g.tmask);
            mul_a(c, d0); mul_b(c, t, d1 / d0);
            return true;
        }
        return false;
    }
};

cplx to_cplx(const cl& v) { return cplx{(double)v.real(), (double)v.imag()}; }

// Pending X / CNOT gates of a stage as an affine map over the tile index (see PermPayload).
struct PermAcc {
    uint16_t col[TILE_BITS];
    uint16_t v0 = 0;
    std::vector<std::pair<uint64_t, uint16_t>> cond;   // (physical control mask outside the tile, flipped tile bits)
    bool identity = true;
    PermAcc() { for (int p = 0; p < TILE_BITS; ++p) col[p] = (uint16_t)(1u << p); }
    void add_x(int t) { v0 ^= (uint16_t)(1u << t); identity = false; }
    void add_cnot(int c, int t) {     // tile positions
        for (int p = 0; p < TILE_BITS; ++p) if ((col[p] >> c) & 1) col[p] ^= (uint16_t)(1u << t);
        if ((v0 >> c) & 1) v0 ^= (uint16_t)(1u << t);
        for (auto& cd : cond) if ((cd.second >> c) & 1) cd.second ^= (uint16_t)(1u << t);
        identity = false;
    }
    bool can_add_ext(uint64_t mask) const {
        for (auto& cd : cond) if (cd.first == mask) return true;
        return (int)cond.size() < PERM_MAX_COND;
    }
    void add_cnot_ext(uint64_t mask, int t) {   // control outside the tile: a per-CTA conditional flip
        for (auto& cd : cond) if (cd.first == mask) { cd.second ^= (uint16_t)(1u << t); identity = false; return; }
        cond.push_back({mask, (uint16_t)(1u << t)});
        identity = false;
    }
};

struct StageEmitter {
    Pass& pass;
    DiagAcc& acc;
    int n_total;
    const int* pos_of;           // physical qubit -> tile position (-1: outside the tile)
    int group = IO_GROUP;
    int regq[REG_BITS];
    uint64_t regphys = 0;
    PermAcc perm;
    // last_real[q]: index in pass.ops of the latest uncontrolled real 2x2 (K_REAL / K_REALPH) on qubit q with
    // nothing non-diagonal on q since: a lone phase on q (RY then RZ) is folded into that op for free
    int last_real[64];
    void reset_last() { for (int q = 0; q < 64; ++q) last_real[q] = -1; }
    bool fold_into_last(int q, const cl& A) {
        const int k = last_real[q];
        if (k < 0) return false;
        DevOp& p = pass.ops[k];
        const int treg = (p.code - OC_GATE) % 4;
        cl w(1, 0);
        if ((p.code - OC_GATE) / 4 == K_REALPH) w = cl((long double)p.m[1], (long double)p.m[3]);
        w *= A;
        p.code = OC_GATE + 4 * K_REALPH + treg;
        p.m[1] = (double)w.real(); p.m[3] = (double)w.imag();
        return true;
    }

    void set_stage(int g) {
        group = g; regphys = 0;
        for (int k = 0; k < REG_BITS; ++k) { regq[k] = pass.desc.tile_q[g * REG_BITS + k]; regphys |= 1ull << regq[k]; }
    }
    int reg_of(int q) const { for (int k = 0; k < REG_BITS; ++k) if (regq[k] == q) return k; return -1; }
    uint8_t reg_mask(uint64_t mask) const {
        uint8_t r = 0;
        for (int k = 0; k < REG_BITS; ++k) if ((mask >> regq[k]) & 1) r |= (uint8_t)(1 << k);
        return r;
    }
    DevOp blank(int gate_idx) const {
        DevOp op; std::memset(&op, 0, sizeof(op));
        op.group = (int8_t)group; op.gate_idx = gate_idx; op.creg = -1;
        return op;
    }
    // Transpose to group `to`, executing every pending X / CNOT on the way.
    void emit_switch(int to) {
        DevOp op = blank(-1);
        op.code = OC_SWITCH + group * NGROUPS + to;
        op.group = (int8_t)to;
        if (!perm.identity) {
            op.flags |= F_PERM;
            PermPayload pp; std::memset(&pp, 0, sizeof(pp));
            for (int p = 0; p < TILE_BITS; ++p) pp.col[p] = perm.col[p];
            pp.v0 = perm.v0;
            pp.n_cond = (uint16_t)perm.cond.size();
            for (size_t k = 0; k < perm.cond.size(); ++k) {
                pp.cond_vec[k] = perm.cond[k].second;
                if (k == 0) op.tmask = perm.cond[k].first;
                else if (k == 1) op.cmask = perm.cond[k].first;
                else pp.cond_mask23[k - 2] = perm.cond[k].first;
            }
            std::memcpy(op.m, &pp, sizeof(pp));
        } else if (to == group) {
            return;
        }
        pass.ops.push_back(op);
        ++pass.n_switches;
        perm = PermAcc();
        set_stage(to);
    }
    // Table block (see tile_core.cuh) for scale * prod_c f_c^{x_c} over thread-level partners of the
    // current stage: tile qubits go to the two thread-index tables, the rest to per-byte tables of the
    // CTA's base index.
    void build_tables(const std::vector<std::pair<int, cl>>& partners, const cl& scale, DevOp* op) {
        std::vector<cl> tile_tab(TABLE_TILE_ENTRIES, cl(1, 0));
        for (int v = 0; v < 16; ++v) tile_tab[v] = scale;
        uint8_t mask = 0;
        const int sh = REG_BITS * group;
        for (auto& pr : partners) {
            const int p = pos_of[pr.first];
            if (p >= 0) {
                if (p >= sh && p < sh + REG_BITS) throw std::runtime_error("plan_local: register qubit in a thread table");
                const int tb = p < sh ? p : p - REG_BITS;          // bit of the thread index
                const int base = tb < 4 ? 0 : 16, bit = 1 << (tb & 3);
                for (int v = 0; v < 16; ++v) if (v & bit) tile_tab[base + v] *= pr.second;
            } else {
                mask |= (uint8_t)(1u << (pr.first / 8));
            }
        }
        op->tab = (int32_t)pass.tab_desc.size();
        op->flags |= F_TABLE;
        TableDesc d; std::memset(&d, 0, sizeof d);
        d.byte_off = (uint32_t)pass.tab_bytes.size();    // relocated by finish_tables()
        d.bytes = mask;
        pass.tab_desc.push_back(d);
        for (auto& e : tile_tab) pass.tab_tile.push_back(to_cplx(e));
        for (int by = 0; by < MAX_INDEX_BYTES; ++by) {
            if (!((mask >> by) & 1)) continue;
            std::vector<cl> e(TABLE_ENTRIES, cl(1, 0));
            for (auto& pr : partners) {
                if (pos_of[pr.first] >= 0 || pr.first / 8 != by) continue;
                const int bit = 1 << (pr.first % 8);
                for (int v = 0; v < TABLE_ENTRIES; ++v) if (v & bit) e[v] *= pr.second;
            }
            for (int v = 0; v < TABLE_ENTRIES; ++v) pass.tab_bytes.push_back(to_cplx(e[v]));
        }
    }
    // Emit every accumulated term that involves qubit q, in the current stage's layout.
    void flush_qubit(int q) {
        const int r = reg_of(q);
        cl A(1, 0);
        { auto it = acc.a.find(q); if (it != acc.a.end()) { A = it->second; acc.a.erase(it); } }
        std::vector<std::pair<int, cl>> tpart;          // thread-level partners
        std::vector<std::pair<int, cl>> rpart;          // register-level partners (register bit, factor)
        for (auto it = acc.b.begin(); it != acc.b.end();) {
            if (it->first.first != q && it->first.second != q) { ++it; continue; }
            const int c = it->first.first == q ? it->first.second : it->first.first;
            if (!is_one(it->second)) {
                const int rc = reg_of(c);
                if (rc >= 0) rpart.push_back({rc, it->second}); else tpart.push_back({c, it->second});
            }
            it = acc.b.erase(it);
        }
        if (tpart.empty() && rpart.empty()) {
            if (is_one(A)) return;
            if (fold_into_last(q, A)) return;
        }
        if (r < 0) {
            // q is thread-level: its own factor and its thread-level partners form a pivot table;
            // each register-level partner rc gets a one-partner table applied to the registers with bit rc
            if (!tpart.empty()) {
                DevOp op = blank(-1);
                op.code = OC_TABLE;
                op.tmask = 1ull << q;
                build_tables(tpart, A, &op);
                pass.ops.push_back(op);
            } else if (!is_one(A)) {
                DevOp op = blank(-1);
                op.code = OC_PHASE;
                op.tmask = 1ull << q;
                op.flags = F_D0_ONE;
                const cplx f = to_cplx(A);
                op.m[0] = 1.0; op.m[6] = f.x; op.m[7] = f.y;
                pass.ops.push_back(op);
            }
            for (auto& pr : rpart) {
                DevOp op = blank(-1);
                op.code = OC_TABLE_REG + pr.first;
                std::vector<std::pair<int, cl>> one{{q, pr.second}};
                build_tables(one, cl(1, 0), &op);
                pass.ops.push_back(op);
            }
            return;
        }
        if (tpart.empty() && rpart.empty()) {
            if (is_one(A)) return;
            DevOp op = blank(-1);
            op.code = OC_DIAG1 + r;
            op.flags = F_D0_ONE;
            const cplx f1 = to_cplx(A);
            op.m[0] = 1.0; op.m[6] = f1.x; op.m[7] = f1.y;
            pass.ops.push_back(op);
            return;
        }
        if (tpart.empty() && rpart.size() == 1 && is_one(A)) {
            DevOp op = blank(-1);
            const int rc = rpart[0].first;
            op.code = OC_PAIR + pair_id(std::min(r, rc), std::max(r, rc));
            const cplx f = to_cplx(rpart[0].second);
            op.m[0] = f.x; op.m[1] = f.y;
            pass.ops.push_back(op);
            return;
        }
        // general: (tables | constant A) x register-partner factors on the registers with bit r
        DevOp op = blank(-1);
        op.code = OC_TABLE_REG + r;
        if (!tpart.empty()) build_tables(tpart, A, &op);
        else { const cplx f = to_cplx(A); op.m[6] = f.x; op.m[7] = f.y; }
        for (auto& pr : rpart) {
            const int slot = pr.first < r ? pr.first : pr.first - 1;     // index among the other three bits, ascending
            const cplx f = to_cplx(pr.second);
            op.m[2 * slot] = f.x; op.m[2 * slot + 1] = f.y;
            op.flags |= (uint8_t)(1u << (F_PM_SHIFT + slot));
        }
        pass.ops.push_back(op);
    }
    // End of a pass that is not the last: lone single-qubit phases that fold into a gate of this pass for free.
    void fold_free_phases() {
        for (auto it = acc.a.begin(); it != acc.a.end();) {
            const int q = it->first;
            bool lone = last_real[q] >= 0 && !is_one(it->second);
            for (auto& kv : acc.b) if (lone && (kv.first.first == q || kv.first.second == q) && !is_one(kv.second)) lone = false;
            if (lone && fold_into_last(q, it->second)) it = acc.a.erase(it); else ++it;
        }
    }
    // End of the gate list: everything that is left.
    void flush_all() {
        for (int k = 0; k < REG_BITS; ++k) flush_qubit(regq[k]);
        // only thread-level qubits remain
        std::vector<std::pair<int, cl>> ones;
        for (auto& kv : acc.a) if (!is_one(kv.second)) ones.push_back({kv.first, kv.second});
        std::vector<std::pair<std::pair<int, int>, cl>> cross;
        for (auto& kv : acc.b) if (!is_one(kv.second)) cross.push_back({kv.first, kv.second});
        if (ones.size() <= 4) {
            bool k_done = is_one(acc.K);
            for (auto& pr : ones) {
                DevOp op = blank(-1);
                op.code = OC_PHASE;
                op.tmask = 1ull << pr.first;
                const cl d0 = k_done ? cl(1, 0) : acc.K;
                if (k_done) op.flags = F_D0_ONE;
                k_done = true;
                const cplx c0 = to_cplx(d0), c1 = to_cplx(d0 * pr.second);
                op.m[0] = c0.x; op.m[1] = c0.y; op.m[6] = c1.x; op.m[7] = c1.y;
                pass.ops.push_back(op);
            }
            if (!k_done) {
                DevOp op = blank(-1);
                op.code = OC_PHASE;
                const cplx c0 = to_cplx(acc.K);
                op.m[0] = c0.x; op.m[1] = c0.y; op.m[6] = c0.x; op.m[7] = c0.y;
                pass.ops.push_back(op);
            }
        } else {
            // one table op: K and all the 1-body factors
            DevOp op = blank(-1);
            op.code = OC_TABLE;
            build_tables(ones, acc.K, &op);
            pass.ops.push_back(op);
        }
        // thread-level pairs: pivot tables, greedy on the pair graph
        while (!cross.empty()) {
            std::map<int, int> deg;
            for (auto& pr : cross) { deg[pr.first.first]++; deg[pr.first.second]++; }
            int pivot = -1, best = 0;
            for (auto& kv : deg) if (kv.second > best) { best = kv.second; pivot = kv.first; }
            std::vector<std::pair<int, cl>> partners;
            std::vector<std::pair<std::pair<int, int>, cl>> rest;
            for (auto& pr : cross) {
                if (pr.first.first == pivot) partners.push_back({pr.first.second, pr.second});
                else if (pr.first.second == pivot) partners.push_back({pr.first.first, pr.second});
                else rest.push_back(pr);
            }
            DevOp op = blank(-1);
            op.code = OC_TABLE;
            op.tmask = 1ull << pivot;
            build_tables(partners, cl(1, 0), &op);
            pass.ops.push_back(op);
            cross.swap(rest);
        }
        acc = DiagAcc();
    }
    static bool is_perm_gate(const HostGate& g) {
        if (g.diag) return false;
        int32_t kind; int8_t d0;
        classify_gate(g.m, &kind, &d0);
        return kind == K_SWAP;
    }
    bool perm_capacity(const HostGate& g) const {
        const int c = g.control();
        if (c < 0 || pos_of[c] >= 0) return true;
        return perm.can_add_ext(g.cmask);
    }
    // X / CNOT: joins the pending permutation; executed by the next switch.
    void emit_perm(const HostGate& g) {
        const int t = g.target();
        flush_qubit(t);              // phases on t do not commute with a flip of t
        last_real[t] = -1;
        const int c = g.control();
        if (c < 0) perm.add_x(pos_of[t]);
        else if (pos_of[c] >= 0) perm.add_cnot(pos_of[c], pos_of[t]);
        else perm.add_cnot_ext(g.cmask, pos_of[t]);
    }
    void emit_gate(const HostGate& g) {
        if (g.diag) {
            if (acc.add(g)) return;
            DevOp op = blank(g.gate_idx);
            std::memcpy(op.m, g.m, sizeof(op.m));
            const uint8_t tregm = reg_mask(g.tmask), cregm = reg_mask(g.cmask);
            op.tmask = g.tmask & ~regphys;
            op.cmask = g.cmask & ~regphys;
            if (g.m[0] == 1.0 && g.m[1] == 0.0) op.flags |= F_D0_ONE;
            if (tregm == 0 && cregm == 0) {
                op.code = OC_PHASE;
                if (g.cmask) op.flags |= F_TCTRL;
            } else {
                op.code = OC_DIAGGEN; op.regm = (uint8_t)(tregm | (cregm << 4));
                if (g.cmask) op.flags |= F_HAS_CTRL;
            }
            pass.ops.push_back(op);
            return;
        }
        const int t = g.target();
        const int treg = reg_of(t);
        if (treg < 0) throw std::runtime_error("plan_local: gate outside its register group");
        const size_t n_before = pass.ops.size();
        flush_qubit(t);
        DevOp op = blank(g.gate_idx);
        std::memcpy(op.m, g.m, sizeof(op.m));
        int32_t kind; int8_t d0;
        classify_gate(g.m, &kind, &d0);
        if (kind == K_SWAP) kind = K_ANTIDIAG;   // X-like gate outside a permutation (never emitted by the sweep today)
        const int c = g.control();
        const int creg = c >= 0 ? reg_of(c) : -1;
        if (kind == K_HADAMARD) {
            if (c >= 0) kind = K_REAL;                   // the factor cannot leave a controlled gate
            else {
                acc.K *= cl((long double)g.m[0], 0);     // h [[1,1],[1,-1]]: the kernel adds / subtracts, h joins the pass constant
                if (pass.ops.size() > n_before && pass.ops.back().code == OC_TABLE_REG + treg) {
                    // the twiddle that flush_qubit just emitted and this butterfly run as one op
                    pass.ops.back().code = OC_TWHAD + treg;
                    last_real[t] = -1;
                    return;
                }
            }
        }
        if (creg >= 0) {
            op.code = OC_CGEN + treg; op.creg = (int8_t)creg;
        } else {
            if (g.cmask) { op.cmask = g.cmask; op.flags |= F_TCTRL; }
            op.code = OC_GATE + kind * 4 + treg;
        }
        pass.ops.push_back(op);
        last_real[t] = (op.code == OC_GATE + K_REAL * 4 + treg && !(op.flags & F_TCTRL)) ? (int)pass.ops.size() - 1 : -1;
    }
};

}  // namespace

static std::vector<Pass> plan_local_impl(const std::vector<HostGate>& gates_in, int n_local, int n_total, const PlanOptions& opt,
                                         std::vector<int>* pass_of_gate);

// A small portfolio instead of one greedy plan (host time: 1-10 ms per plan, cached with the plan).
// * Tile relabelling pays when it saves passes (chain-like circuits: hea28 80 -> 55); where it does not (random32, qft30:
//   the same pass count either way) its extra swap gates only add transposes and make untouched qubits look touched to
//   the support tracking -- measured on B200: random32 from a reset 274 -> 329 ms.  So it is kept only if it needs
//   strictly fewer passes.
// * Taking the highest-scoring tile for every pass is not the fewest passes overall: with fewer candidates per pass
//   (closer to first-come order) random32 needs 11 passes instead of 12.  The candidate counts below are tried in
//   order and a later one replaces the plan only if it needs strictly less HBM traffic (plan_traffic: the number of passes
//   on a dense state; after a reset the passes that run while qubits are still |0> only visit a fraction of the tiles, and
//   the plan whose early passes stay small wins).
PlanChoices* ChoiceMemoTable::begin(const std::vector<HostGate>& gates, uint64_t flags) {
    std::vector<uint64_t> skey;
    skey.reserve(gates.size() * 3 + 2);
    skey.push_back(flags);
    skey.push_back((uint64_t)gates.size());
    for (const HostGate& g : gates) {
        int32_t kind = 0;
        int8_t d0 = 0;
        classify_gate(g.m, &kind, &d0);
        skey.push_back(g.tmask);
        skey.push_back(g.cmask ^ (g.diag ? 1ull << 63 : 0));
        skey.push_back((uint64_t)(uint32_t)kind | ((uint64_t)(uint8_t)d0 << 32));
    }
    Entry* e = nullptr;
    for (Entry& x : entries_) if (x.skey == skey) e = &x;
    const bool known = e != nullptr;
    if (!e) {
        if (entries_.size() < capacity_ || entries_.empty()) { entries_.emplace_back(); e = &entries_.back(); }
        else { e = &entries_[0]; for (Entry& x : entries_) if (x.stamp < e->stamp) e = &x; }
        e->skey = std::move(skey);
        e->ch = PlanChoices();
    }
    e->stamp = ++clock_;
    e->ch.replay = known && !e->ch.tape.empty();
    e->ch.pos = 0;
    return &e->ch;
}

double plan_traffic(const std::vector<Pass>& passes, uint64_t* zero_mask) {
    double total = 0.0;
    uint64_t zm = *zero_mask;
    for (const Pass& p : passes) {
        uint64_t tile = 0;
        for (int k = 0; k < TILE_BITS; ++k) tile |= 1ull << p.desc.tile_q[k];
        const double launched = std::ldexp(1.0, -__builtin_popcountll(zm & ~tile));
        const double read = std::ldexp(1.0, -__builtin_popcountll(zm & tile));
        total += launched * 0.5 * (1.0 + read);
        zm &= ~p.touch_mask;
    }
    *zero_mask = zm;
    return total;
}

std::vector<Pass> plan_local(const std::vector<HostGate>& gates, int n_local, int n_total, const PlanOptions& opt,
                             std::vector<int>* pass_of_gate) {
    std::vector<Pass> best;
    std::vector<int> best_of;
    double best_cost = 0.0;
    const uint64_t zm0 = opt.zero_mask & ((1ull << n_local) - 1);
    bool have = false;
    const int cands[3] = {opt.candidates, 4, 2};
    PlanChoices* ch = opt.choices;
    int forced = -1, best_idx = 0;
    if (ch && ch->replay) forced = ch->pos < ch->tape.size() ? ch->tape[ch->pos++] : 0;
    for (int ci = 0; ci < (opt.portfolio ? 3 : 1); ++ci) {
        if (ci > 0 && cands[ci] >= opt.candidates) continue;
        for (int rl = 0; rl < (opt.relabel ? 2 : 1); ++rl) {
            if (forced >= 0 && ci * 2 + rl != forced) continue;
            if (have && best.size() <= 1) break;
            PlanOptions o = opt;
            o.candidates = cands[ci];
            o.relabel = rl == 1;
            std::vector<int> of;
            std::vector<Pass> cand = plan_local_impl(gates, n_local, n_total, o, pass_of_gate ? &of : nullptr);
            uint64_t zm = zm0;
            const double cost = plan_traffic(cand, &zm);      // dense: the number of passes
            if (!have || cost < best_cost - 1e-9) { best = std::move(cand); best_of.swap(of); best_cost = cost; have = true; best_idx = ci * 2 + rl; }
        }
    }
    if (!have) {      // a recorded choice that this option set does not offer: the plain plan
        PlanOptions o = opt;
        o.relabel = false;
        best = plan_local_impl(gates, n_local, n_total, o, pass_of_gate ? &best_of : nullptr);
    }
    if (ch && !ch->replay) ch->tape.push_back(best_idx);
    if (pass_of_gate) pass_of_gate->swap(best_of);
    return best;
}

static std::vector<Pass> plan_local_impl(const std::vector<HostGate>& gates_in, int n_local, int n_total,
                                         const PlanOptions& opt, std::vector<int>* pass_of_gate) {
    std::vector<HostGate> gates = gates_in;   // relabelling appends swap gates and renames qubits of the gates still to run
    if (n_local < TILE_BITS) throw std::runtime_error("plan_local: n_local < TILE_BITS");
    if (n_total > 62) throw std::runtime_error("plan_local: too many qubits");
    std::vector<Pass> passes;
    const int G = (int)gates.size();
    if (pass_of_gate) pass_of_gate->assign((size_t)G, -1);
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
    const int min_low = std::max(3, std::min(opt.min_low, TILE_BITS));   // >= 3: 128-byte segments, and the kernels' offset stash relies on it

    int gate_budget = opt.max_ops_per_pass;
    // The phase polynomial outlives a pass: a term is only emitted when a non-diagonal gate is about to hit
    // one of its qubits (or when the gate list ends), so a controlled phase between a qubit of this pass's
    // tile and one of a later pass's tile costs nothing here and folds into a per-CTA constant there.
    DiagAcc acc;
    // relabelling state: partner[s] = physical qubit whose logical content currently sits in pinned slot s (and whose
    // own position holds logical s); partner[s] == s: identity.  Always a product of disjoint transpositions.
    int partner[8];
    for (int sl = 0; sl < 8; ++sl) partner[sl] = sl;
    static const double kX[8] = {0, 0, 1, 0, 1, 0, 0, 0};
    auto swap_bits = [](uint64_t m, int x, int y) {
        const uint64_t bx = (m >> x) & 1ull, by = (m >> y) & 1ull;
        return bx == by ? m : m ^ ((1ull << x) | (1ull << y));
    };
    while (!pending.empty()) {
        const DiagAcc acc_start = acc;
        const size_t gates_start = gates.size();
        std::vector<std::pair<int, int>> swaps;      // physical-qubit transpositions appended to this pass, in order
        int trial[8];                                // partner[] if this pass is accepted
        for (int sl = 0; sl < 8; ++sl) trial[sl] = partner[sl];
        // ---- pass level: grow the tile greedily, take everything that commutes to the front ----
        // Candidate c hands the free tile positions to the qubits in first-come order but refuses the
        // first c newcomers: on layered circuits (entangler chains) that slides the window along the
        // dependency front, and the candidate that takes the most gates wins.  c = 0 is plain first-come.
        uint64_t tile = 0;
        int tile_n = 0;
        std::vector<int> taken, rest;
        const int limit = std::min<int>((int)pending.size(), opt.window);
        auto select = [&](int refuse, uint64_t& tile_o, int& tile_n_o, std::vector<int>* taken_o, std::vector<int>* rest_o) {
            uint64_t t = (1ull << min_low) - 1, refused = 0;
            int tn = min_low, n_taken = 0, score = 0;
            Blocked blk;
            for (int k = 0; k < (int)pending.size(); ++k) {
                const int gi = pending[k];
                const HostGate& g = gates[gi];
                bool ok = k < limit && n_taken < gate_budget && blk.can_pass(g);
                if (ok && !g.diag && !(t & g.tmask)) {
                    if (refused & g.tmask) ok = false;
                    else if (refuse > 0) { refused |= g.tmask; --refuse; ok = false; }
                    else if (tn < TILE_BITS) { t |= g.tmask; ++tn; }
                    else ok = false;
                }
                if (ok) { ++n_taken; score += g.diag ? 1 : 2; if (taken_o) taken_o->push_back(gi); }
                else {
                    blk.skip(g);
                    if (rest_o) rest_o->push_back(gi);
                    else if (k >= limit || (tn == TILE_BITS && (blk.x & t) == t)) break;   // scoring only: nothing more can be taken
                }
            }
            tile_o = t; tile_n_o = tn;
            return score;
        };
        {
            int best = 0, best_score = -1;
            for (int c = 0; c < std::max(1, opt.candidates); ++c) {
                uint64_t t; int tn;
                const int sc = select(c, t, tn, nullptr, nullptr);
                if (sc > best_score) { best_score = sc; best = c; }
                if (tn < TILE_BITS) break;      // the circuit ran out of new qubits: later candidates only refuse more
            }
            select(best, tile, tile_n, &taken, &rest);
        }
        if (taken.empty()) throw std::runtime_error("plan_local: no progress");
        // last pass of a relabelled plan: free tile positions go to the homes of the displaced qubits first, so that they
        // return in this pass's permuting transpose instead of in a pass of their own
        if (opt.relabel && rest.empty())
            for (int sl = 0; sl < min_low && tile_n < TILE_BITS; ++sl)
                if (partner[sl] != sl && !((tile >> partner[sl]) & 1)) { tile |= 1ull << partner[sl]; ++tile_n; }
        // pad the tile with the lowest unused local qubits (keeps segments long)
        for (int q = 0; q < n_local && tile_n < TILE_BITS; ++q)
            if (!((tile >> q) & 1)) { tile |= 1ull << q; ++tile_n; }

        if (opt.relabel) {
            // next_use[q]: how soon physical qubit q is the target of a non-diagonal gate among the gates left after this
            // pass (labels are physical, so next_use[slot] is the need of whatever the slot currently holds)
            std::vector<int> next_use(n_total, INT_MAX);
            {
                int k = 0;
                for (int gi : rest) {
                    const HostGate& g = gates[gi];
                    if (!g.diag && next_use[g.target()] == INT_MAX) next_use[g.target()] = k;
                    if (++k > opt.window) break;
                }
            }
            // SWAP(x, y) = CNOT(x,y) CNOT(y,x) CNOT(x,y), both inside the tile: joins the pass's last permuting transpose
            auto append_swap = [&](int x, int y) {
                const int pairs[3][2] = {{x, y}, {y, x}, {x, y}};
                for (auto& pr : pairs) { gates.push_back(make_gate(pr[1], pr[0], kX, -1)); taken.push_back((int)gates.size() - 1); }
                swaps.push_back({x, y});
                std::swap(next_use[x], next_use[y]);     // the gates written for x now find their data at y and vice versa
            };
            for (int sl = 0; sl < min_low; ++sl) trial[sl] = partner[sl];
            if (rest.empty()) {
                // last pass of this call: send every displaced qubit home that this tile can reach
                for (int sl = 0; sl < min_low; ++sl)
                    if (trial[sl] != sl && ((tile >> trial[sl]) & 1)) { append_swap(sl, trial[sl]); trial[sl] = sl; }
            } else {
                for (int sl = 0; sl < min_low; ++sl) {
                    const int cur = trial[sl];
                    if (cur != sl && !((tile >> cur) & 1)) continue;     // the resident's home is not in this tile
                    int best = -1;
                    for (int q = min_low; q < n_local; ++q) {
                        if (!((tile >> q) & 1) || q == cur) continue;
                        bool is_partner = false;
                        for (int s2 = 0; s2 < min_low; ++s2) if (trial[s2] == q) is_partner = true;
                        if (is_partner) continue;
                        if (next_use[q] < next_use[sl] && (best < 0 || next_use[q] < next_use[best])) best = q;
                    }
                    if (best < 0) continue;
                    if (cur != sl) append_swap(sl, cur);                 // the old resident goes home first (keeps the
                    append_swap(sl, best);                               //   layout a product of disjoint transpositions)
                    trial[sl] = best;
                }
            }
        }

        // ---- tile positions: pinned low run, then by first non-permutation target use -----------
        // Two hand-out orders of the free tile positions are planned and the one with fewer stage switches
        // wins: the group-0 positions go to the last-used qubits (a pass must end in the group-2 or group-1
        // layout to store coalesced, so ending in group 0 costs one more switch) or to the middle ones.
        auto build_pass = [&](int variant, DiagAcc& acc, Pass& pass) {
        std::memset(&pass.desc, 0, sizeof(pass.desc));
        pass.desc.n_local = n_local;
        std::vector<int> order;  // qubits by first use as the target of a gate that needs registers
        uint64_t seen = (1ull << min_low) - 1;
        for (int gi : taken) {
            const HostGate& g = gates[gi];
            if (!g.diag && !StageEmitter::is_perm_gate(g) && !(seen & g.tmask)) { seen |= g.tmask; order.push_back(g.target()); }
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
            static const int handout[2][NGROUPS] = {{2, 1, 0}, {2, 0, 1}};
            for (int gi = 0; gi < NGROUPS; ++gi) {
                const int g = handout[variant][gi];
                for (int p = g * REG_BITS; p < (g + 1) * REG_BITS; ++p)
                    if (p >= min_low) free_pos.push_back(p);
            }
            if (order.size() != free_pos.size()) throw std::runtime_error("plan_local: tile size");
            for (size_t i = 0; i < order.size(); ++i) {
                pass.desc.tile_q[free_pos[i]] = order[i];
                pos_of[order[i]] = free_pos[i];
            }
        }
        for (int p = 0; p < TILE_BITS; ++p) pass.desc.sorted_q[p] = pass.desc.tile_q[p];
        std::sort(pass.desc.sorted_q, pass.desc.sorted_q + TILE_BITS);

        // ---- stage level: sweep per register group ---------------------------------------------------
        // b2: gates skipped in this sweep; pp: X / CNOT gates accepted into the pending permutation.
        // A later gate may run in this stage only if it commutes past both sets; permutation gates
        // compose among themselves in program order, so they only need to pass b2.
        std::vector<int> remaining = taken;
        StageEmitter em{pass, acc, n_total, pos_of};
        em.reset_last();
        em.set_stage(IO_GROUP);
        // One sweep over `remaining` for register group `grp`.  X / CNOT gates whose target is a register
        // bit of the stage run at once as register exchanges; the others join the pending permutation that
        // the next switch executes.  em == nullptr: dry run, returns the weight of what would execute.
        auto sweep = [&](int grp, StageEmitter* emit, std::vector<int>* rem_out) {
            Blocked b2, pp;
            int n_cond = 0, score = 0;
            std::vector<uint64_t> cond_masks;
            for (int gi : remaining) {
                const HostGate& g = gates[gi];
                const bool is_perm = StageEmitter::is_perm_gate(g);
                const bool in_regs = !g.diag && pos_of[g.target()] / REG_BITS == grp;
                bool ok, as_perm = false;
                if (is_perm) {
                    as_perm = true;
                    ok = b2.can_pass(g);
                    if (ok) {      // capacity of the switch payload for controls outside the tile
                        const int c = g.control();
                        if (c >= 0 && pos_of[c] < 0) {
                            if (emit) ok = emit->perm_capacity(g);
                            else if (std::find(cond_masks.begin(), cond_masks.end(), g.cmask) == cond_masks.end()) {
                                if (n_cond < PERM_MAX_COND) { cond_masks.push_back(g.cmask); ++n_cond; } else ok = false;
                            }
                        }
                    }
                } else {
                    ok = b2.can_pass(g) && pp.can_pass(g) && (g.diag || in_regs);
                }
                if (!ok) { b2.skip(g); if (rem_out) rem_out->push_back(gi); continue; }
                if (as_perm) { if (emit) emit->emit_perm(g); pp.skip(g); }
                else {
                    if (emit) emit->emit_gate(g);
                    if (!g.diag) score += is_perm ? 1 : 4;
                }
            }
            return score;
        };
        while (!remaining.empty()) {
            {
                // next stage: the register group in which the most work can run (ties: stay, then the
                // group of the first waiting gate)
                int want = em.group, best = -1;
                int first_grp = em.group;
                for (int gi : remaining) {
                    const HostGate& g = gates[gi];
                    if (!g.diag && !StageEmitter::is_perm_gate(g)) { first_grp = pos_of[g.target()] / REG_BITS; break; }
                }
                int order[NGROUPS + 2], n_order = 0;
                order[n_order++] = em.group; order[n_order++] = first_grp;
                for (int g2 = 0; g2 < NGROUPS; ++g2) order[n_order++] = g2;
                bool seen_grp[NGROUPS] = {false, false, false};
                for (int k = 0; k < n_order && opt.best_group; ++k) {
                    const int grp = order[k];
                    if (seen_grp[grp]) continue;
                    seen_grp[grp] = true;
                    const int sc = sweep(grp, nullptr, nullptr);
                    if (sc > best) { best = sc; want = grp; }
                }
                if (best <= 0 || !opt.best_group) want = first_grp;
                em.emit_switch(want);    // no-op when nothing is pending and the group stays
            }
            std::vector<int> rem2;
            sweep(em.group, &em, &rem2);
            if (rem2.size() == remaining.size()) throw std::runtime_error("plan_local: stage made no progress");
            remaining.swap(rem2);
        }
        if (rest.empty()) em.flush_all();
        else em.fold_free_phases();
        // store layout: stay in group 1 or 2 (a pending permutation still needs its same-group switch)
        em.emit_switch(em.group == 0 ? IO_GROUP : em.group);
        pass.desc.io_out = em.group;
        if (opt.macro_ops) fuse_macro_ops(pass.ops);
        };
        Pass pass;
        {
            Pass cand[2];
            DiagAcc cacc[2] = {acc_start, acc_start};
            int best = 0;
            for (int v = 0; v < 2; ++v) {
                build_pass(v, cacc[v], cand[v]);
                if (v > 0 && (cand[v].n_switches < cand[best].n_switches ||
                              (cand[v].n_switches == cand[best].n_switches && cand[v].ops.size() < cand[best].ops.size()))) best = v;
            }
            pass = std::move(cand[best]);
            acc = cacc[best];
        }
        pass.desc.n_ops = (int)pass.ops.size();
        pass.desc.n_tab = (int)pass.tab_desc.size();
        pass.fp64_per_thread = estimate_fp64(pass.ops);
        pass.touch_mask = 0;
        for (int gi : taken) if (!gates[gi].diag) pass.touch_mask |= gates[gi].tmask;
        if (pass.desc.n_ops >= MAX_OPS_PER_PASS || pass.desc.n_tab > MAX_TABLE_OPS) {
            // the op list is a kernel parameter of bounded size: take fewer gates and plan this pass again
            if (taken.size() - 3 * swaps.size() <= 1) throw std::runtime_error("plan_local: one gate does not fit a pass");
            gate_budget = (int)(taken.size() - 3 * swaps.size()) / 2;
            acc = acc_start;
            gates.resize(gates_start);     // drop the swap gates of the rejected attempt
            continue;
        }
        gate_budget = opt.max_ops_per_pass;
        pass.finish_tables();
        passes.push_back(std::move(pass));
        if (pass_of_gate)
            for (int gi : taken) if (gi < G) (*pass_of_gate)[gi] = (int)passes.size() - 1;
        // the pass is accepted: its swaps are now part of the state, rename the qubits of every gate still to run
        for (auto& sw : swaps)
            for (int gi : rest) {
                gates[gi].tmask = swap_bits(gates[gi].tmask, sw.first, sw.second);
                gates[gi].cmask = swap_bits(gates[gi].cmask, sw.first, sw.second);
            }
        for (int sl = 0; sl < min_low; ++sl) partner[sl] = trial[sl];
        if (rest.empty()) {
            // qubits still away from home (their partner was outside the last tile): one more pass of swaps only
            for (int sl = 0; sl < min_low; ++sl)
                if (partner[sl] != sl) {
                    const int x = sl, y = partner[sl];
                    const int pairs[3][2] = {{x, y}, {y, x}, {x, y}};
                    for (auto& pr : pairs) { gates.push_back(make_gate(pr[1], pr[0], kX, -1)); rest.push_back((int)gates.size() - 1); }
                    partner[sl] = sl;
                }
        }
        pending.swap(rest);
    }
    return passes;
}

int estimate_fp64(const std::vector<DevOp>& ops) {
    int total = 0;
    bool scalar_dirty = false;
    for (const DevOp& raw : ops) {
        int code = raw.code;
        if (code == OC_REALPH4) code = OC_GATE + 4 * K_REALPH;
        if (code == OC_TWHAD4) code = OC_TWHAD;
        const unsigned pm = (raw.flags >> F_PM_SHIFT) & 7u;
        const int w4 = 4 * (int)((pm & 1u) + 2 * ((pm >> 1) & 1u) + 4 * ((pm >> 2) & 1u));     // twiddle variants over the other register bits
        if (code >= OC_SWITCH) { if (scalar_dirty) total += 4 * NREG; scalar_dirty = false; continue; }
        if (code < OC_CGEN) {
            static const int per_pair[] = {16, 8, 8, 8, 4, 12};     // K_GENERAL, K_REAL, K_RXLIKE, K_ANTIDIAG, K_HADAMARD, K_REALPH
            const int c = per_pair[(code - OC_GATE) / 4] * (NREG / 2);
            total += (raw.flags & F_TCTRL) ? c / 2 : c;             // thread-level control: half the threads on average
        } else if (code < OC_DIAG1) total += 16 * (NREG / 4);       // OC_CGEN: half the pairs
        else if (code < OC_PHASE) total += 4 * (NREG / 2);          // OC_DIAG1
        else if (code == OC_PHASE) { total += 4; scalar_dirty = true; }
        else if (code == OC_DIAGGEN) total += 4 * NREG;
        else if (code == OC_TABLE) { total += 12; scalar_dirty = true; }
        else if (code < OC_PAIR) total += ((raw.flags & F_TABLE) ? 8 : 0) + w4 + 4 * (NREG / 2);                    // OC_TABLE_REG
        else if (code < OC_TWHAD) total += 4 * (NREG / 4);          // OC_PAIR
        else total += ((raw.flags & F_TABLE) ? 8 : 0) + w4 + 4 * (NREG / 2) + 4 * (NREG / 2);                     // OC_TWHAD
    }
    if (scalar_dirty) total += 4 * NREG;
    return total;
}

bool compose_remap(const std::vector<std::pair<int, int>>& swaps, int n_local, int rank, RemapPlan* out, bool inverse) {
    // at[p] = the position whose (old) bit sits at position p after the swaps: new_bit[p] = old_bit[at[p]]
    std::vector<int> pos;      // positions involved, in order of first appearance
    auto idx_of = [&](int p) {
        for (size_t k = 0; k < pos.size(); ++k) if (pos[k] == p) return (int)k;
        pos.push_back(p);
        return (int)pos.size() - 1;
    };
    std::vector<int> at;
    for (auto& sw : swaps) {
        const int a = idx_of(sw.first), b = idx_of(sw.second);
        at.resize(pos.size(), -1);
        for (size_t k = 0; k < at.size(); ++k) if (at[k] < 0) at[k] = pos[k];
        std::swap(at[a], at[b]);
    }
    RemapPlan rp;
    std::vector<int> G, L;
    for (int p : pos) (p >= n_local ? G : L).push_back(p);
    rp.n_global = (int)G.size(); rp.n_local_pos = (int)L.size();
    if ((int)G.size() > MAX_REMAP || (int)L.size() > MAX_REMAP_LOCAL) return false;
    // load side: the old bit q ends up at position where(q), so the source of new index i' has old_bit[q] = new_bit[where(q)];
    // store side: the new bit at position p is old_bit[at[p]], so the destination of old index i has new_bit[p] = old_bit[at[p]]
    auto where = [&](int q) {
        if (inverse) { for (size_t k = 0; k < pos.size(); ++k) if (pos[k] == q) return at[k]; return q; }
        for (size_t k = 0; k < pos.size(); ++k) if (at[k] == q) return pos[k];
        return q;
    };
    bool identity = true;
    for (size_t k = 0; k < pos.size(); ++k) if (at[k] != pos[k]) identity = false;
    if (identity) { *out = rp; return true; }
    rp.on = true;
    // selector bits: the local positions of the new index that carry an old rank-index bit
    int sel_of_g[MAX_REMAP];
    for (size_t g = 0; g < G.size(); ++g) {
        const int w = where(G[g]);
        sel_of_g[g] = -1;
        if (w < n_local) { sel_of_g[g] = rp.n_sel; rp.sel_lq[rp.n_sel++] = w; }
    }
    for (int sel = 0; sel < (1 << rp.n_sel); ++sel) {
        int r = rank;
        for (size_t g = 0; g < G.size(); ++g) {
            const int j = G[g] - n_local, w = where(G[g]);
            const int bit = sel_of_g[g] >= 0 ? (sel >> sel_of_g[g]) & 1 : (rank >> (w - n_local)) & 1;
            r = (r & ~(1 << j)) | (bit << j);
        }
        rp.src_rank[sel] = r;
    }
    // the local positions of the source index: this rank's bits (constants) or other local bits of the new index
    for (int q : L) {
        rp.lmask |= 1ull << q;
        const int w = where(q);
        if (w >= n_local) rp.rconst |= (uint64_t)((rank >> (w - n_local)) & 1) << q;
        else { rp.mv_from[rp.n_mv] = w; rp.mv_to[rp.n_mv] = q; ++rp.n_mv; }
    }
    *out = rp;
    return true;
}

void apply_remap(const RemapPlan& rp, PassDesc* pd, bool store_side) {
    pd->remap_on = (rp.on && !store_side) ? 1 : 0;
    pd->remap_st = (rp.on && store_side) ? 1 : 0;
    pd->remap_n = (int8_t)rp.n_sel;
    for (int k = 0; k < rp.n_sel; ++k) pd->remap_lq[k] = (int8_t)rp.sel_lq[k];
    pd->remap_n_mv = (int8_t)rp.n_mv;
    for (int k = 0; k < rp.n_mv; ++k) { pd->remap_mv_from[k] = (int8_t)rp.mv_from[k]; pd->remap_mv_to[k] = (int8_t)rp.mv_to[k]; }
    pd->remap_lmask = rp.lmask;
    pd->remap_const = rp.rconst;
}

Pass make_identity_pass(int n_local) {
    if (n_local < TILE_BITS) throw std::runtime_error("make_identity_pass: n_local < TILE_BITS");
    Pass pass;
    std::memset(&pass.desc, 0, sizeof(pass.desc));
    pass.desc.n_local = n_local;
    for (int p = 0; p < TILE_BITS; ++p) pass.desc.tile_q[p] = pass.desc.sorted_q[p] = p;
    pass.desc.io_out = IO_GROUP;
    pass.finish_tables();
    return pass;
}

// ---------------------------------------------------------------------------------------------------
std::vector<DistStep> plan_distributed(const std::vector<HostGate>& gates, int n_total, int n_local,
                                       std::vector<int>& perm, bool restore_identity, bool local_swap_steps,
                                       const PlanOptions* tail_opt, int defer_max_ops) {
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
        if (getenv("DVD_PLAN_DUMP")) fprintf(stderr, "  swap: logical %d (at global position %d) comes in, logical %d (at local position %d) goes out\n", a, gq, b, lq);
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
    auto mapped = [&](const HostGate& g) {
        HostGate pg = g;
        pg.tmask = map_mask(g.tmask);
        pg.cmask = map_mask(g.cmask);
        return pg;
    };

    // List scheduling over LOGICAL qubits: run everything that is executable with the current layout
    // and commutes past what had to wait (same rule as the pass level), and only then pay for swaps.
    // On layered circuits this executes a whole light cone per layout instead of one layer.
    std::vector<int> pending(G);
    for (int i = 0; i < G; ++i) pending[i] = i;
    const int first_victim = n_local > 8 ? 5 : 0;    // keep swap segments >= 512 B when there is a choice
    const bool defer_tails = tail_opt && defer_max_ops > 0 && n_local >= TILE_BITS;
    while (!pending.empty()) {
        size_t n_deferred = 0;       // pending[0 .. n_deferred): executable now, but kept for the next layout (see below)
        {
            Blocked blk;
            std::vector<int> rest, exec;
            for (int gi : pending) {
                const HostGate& g = gates[gi];
                const bool local = g.diag || perm[g.target()] < n_local;
                if (local && blk.can_pass(g)) exec.push_back(gi);
                else { blk.skip(g); rest.push_back(gi); }
            }
            // Tail deferral: what is executable under this layout rarely fills its last pass (random32 on 8 ranks: passes
            // of 5, 3, 6 and 8 ops in front of the four swap rounds).  The gates of such a tail pass run just as well
            // AFTER the swaps, inside the passes of the next layout -- they commute past everything that waits (that is
            // how they got here) and the pass planner put the rest of this step in front of them -- provided their
            // qubits stay local, which the eviction rule below sees to (they are the next uses).
            if (defer_tails && !rest.empty() && exec.size() > 1) {
                std::vector<HostGate> phys;
                for (int gi : exec) phys.push_back(mapped(gates[gi]));
                std::vector<int> pass_of;       // (the last step is never a LOCAL_GATES step here: this list IS the step)
                const std::vector<Pass> passes = plan_local(phys, n_local, n_total, *tail_opt, &pass_of);
                if (passes.size() >= 2 && (int)passes.back().ops.size() <= defer_max_ops) {
                    const int last = (int)passes.size() - 1;
                    std::vector<int> keep, tail;
                    for (size_t k = 0; k < exec.size(); ++k) (pass_of[k] == last ? tail : keep).push_back(exec[k]);
                    if (!tail.empty() && !keep.empty()) {
                        exec.swap(keep);
                        n_deferred = tail.size();
                        std::vector<int> merged;
                        merged.reserve(tail.size() + rest.size());
                        for (int gi : tail) merged.push_back(gi);
                        for (int gi : rest) merged.push_back(gi);
                        rest.swap(merged);
                    }
                }
            }
            for (int gi : exec) local_step().gates.push_back(mapped(gates[gi]));
            pending.swap(rest);
        }
        if (pending.empty()) break;
        // the gates that wait for nothing but a rank-index target: bring those qubits in
        std::vector<int> want;
        {
            Blocked blk;
            for (size_t k = n_deferred; k < pending.size(); ++k) {     // (deferred gates block nothing: they are executable)
                const int gi = pending[k];
                const HostGate& g = gates[gi];
                if (!g.diag && perm[g.target()] >= n_local && blk.can_pass(g) &&
                    std::find(want.begin(), want.end(), g.target()) == want.end() &&
                    (int)want.size() < std::min(n_total - n_local, n_local))   // one evictable local position per swap-in
                    want.push_back(g.target());
                blk.skip(g);
            }
        }
        if (want.empty()) throw std::runtime_error("plan_distributed: no progress");
        // evict the local qubits whose next non-diagonal use is farthest away (Belady)
        std::vector<int> next_use(n_total, G + 1);
        for (int k = (int)pending.size() - 1; k >= 0; --k)
            if (!gates[pending[k]].diag) next_use[gates[pending[k]].target()] = k;
        // the rank-index positions this round vacates, and one victim per swap-in
        std::vector<int> vacated, victims;
        for (int lqbit : want) vacated.push_back(perm[lqbit]);
        std::vector<char> taken_pos(n_local, 0);
        for (size_t k = 0; k < want.size(); ++k) {
            int victim = -1, best = -1;
            bool best_home = false;
            for (int pass = 0; pass < 2 && victim < 0; ++pass)
                for (int p = n_local - 1; p >= (pass == 0 ? first_victim : 0); --p) {   // ties: prefer high local positions
                    const int lq = inv[p];
                    if (taken_pos[p] || std::find(want.begin(), want.end(), lq) != want.end()) continue;
                    // ... and, first of all, a displaced rank-index qubit whose home is vacated in this round: it goes
                    // home now and the layout restore has one swap less to do
                    const bool home = lq >= n_local && std::find(vacated.begin(), vacated.end(), lq) != vacated.end();
                    if (next_use[lq] > best || (next_use[lq] == best && home && !best_home)) { best = next_use[lq]; victim = p; best_home = home; }
                }
            if (victim < 0) throw std::runtime_error("plan_distributed: no local position left to evict");
            taken_pos[victim] = 1;
            victims.push_back(victim);
        }
        // pairing: a victim lands on the rank-index position it is swapped with -- its home where that is on offer
        std::vector<int> pos_of_victim(victims.size(), -1);
        std::vector<char> pos_used(vacated.size(), 0);
        for (size_t v = 0; v < victims.size(); ++v)
            for (size_t k = 0; k < vacated.size(); ++k)
                if (!pos_used[k] && vacated[k] == inv[victims[v]]) { pos_of_victim[v] = vacated[k]; pos_used[k] = 1; break; }
        for (size_t v = 0; v < victims.size(); ++v)
            if (pos_of_victim[v] < 0)
                for (size_t k = 0; k < vacated.size(); ++k)
                    if (!pos_used[k]) { pos_of_victim[v] = vacated[k]; pos_used[k] = 1; break; }
        for (size_t v = 0; v < victims.size(); ++v) emit_swap(pos_of_victim[v], victims[v]);
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
            if (local_swap_steps) {
                DistStep sw; sw.kind = DistStep::LOCAL_SWAP; sw.gq = p; sw.lq = x;
                steps.push_back(std::move(sw));
            } else {
                emit_local_cnot(p, x); emit_local_cnot(x, p); emit_local_cnot(p, x);
            }
            const int other = inv[p];
            perm[p] = p; perm[other] = x; inv[p] = p; inv[x] = other;
        }
    }
    return steps;
}

// Hand swap rounds to the store of the pass in front of them (planner.h: DistPlan::store).  Returns false when a round
// with LOCAL_SWAP steps -- which no executor runs as steps -- could not be taken: the caller plans again with CNOT triples.
static bool assign_store_side(DistPlan& dp, int n_local, int mode, uint64_t zero_mask) {
    std::vector<DistStep> steps;
    std::vector<std::vector<Pass>> plans;
    std::vector<std::vector<std::pair<int, int>>> store;
    int last_local = -1;            // index (new lists) of the last LOCAL_GATES step with passes
    bool first_pass_pulls = false;  //   its first pass loads through the swaps in front of it
    bool has_store = false;         //   its last pass already stores through a round
    bool pull_pending = false;      // swaps kept in the step list since that step
    const size_t n = dp.steps.size();
    size_t i = 0;
    while (i < n) {
        if (dp.steps[i].kind == DistStep::LOCAL_GATES) {
            steps.push_back(std::move(dp.steps[i]));
            plans.push_back(std::move(dp.plans[i]));
            store.emplace_back();
            if (!plans.back().empty()) { last_local = (int)steps.size() - 1; first_pass_pulls = pull_pending; has_store = false; pull_pending = false; }
            for (const Pass& p : plans.back()) zero_mask &= ~p.touch_mask;      // engine.cu run_pass: support |= touch_mask
            ++i;
            continue;
        }
        size_t j = i;
        std::vector<std::pair<int, int>> pairs;
        bool local_swaps = false;
        while (j < n && dp.steps[j].kind != DistStep::LOCAL_GATES) {
            pairs.push_back({dp.steps[j].gq, dp.steps[j].lq});
            local_swaps |= dp.steps[j].kind == DistStep::LOCAL_SWAP;
            ++j;
        }
        bool ends_schedule = true;
        for (size_t k = j; k < n; ++k) if (dp.steps[k].kind != DistStep::LOCAL_GATES || !dp.steps[k].gates.empty()) ends_schedule = false;
        RemapPlan rp;
        const bool take = mode >= 1 && (ends_schedule || (mode >= 2 && zero_mask == 0)) && last_local >= 0 && !has_store && !pull_pending &&
                          (!first_pass_pulls || plans[last_local].size() >= 2) &&
                          compose_remap(pairs, n_local, 0, &rp, /*inverse=*/true);
        if (take) {
            store[last_local] = std::move(pairs);
            has_store = true;
            zero_mask = 0;                       // the engine stores the implied zeros in front of a storing pass
            ++dp.n_store;
        } else {
            if (local_swaps) return false;
            for (size_t k = i; k < j; ++k) {
                if (dp.steps[k].lq >= 0 && dp.steps[k].lq < n_local) zero_mask &= ~(1ull << dp.steps[k].lq);   // rebuilt positions: never implied
                steps.push_back(std::move(dp.steps[k])); plans.emplace_back(); store.emplace_back();
            }
            pull_pending = true;
        }
        i = j;
    }
    dp.steps = std::move(steps);
    dp.plans = std::move(plans);
    dp.store = std::move(store);
    return true;
}

// Spread a round of k >= 2 swaps over the stores of the last k passes of the step in front of it, one swap each.
// A round moves (1 - 2^-k) of the chunk over NVLink in ONE pass, which is then bound by the links (measured on 4 x B200:
// 12.9 GB per direction = 22 ms in a pass that otherwise takes 13 ms), while a single swap moves half a chunk in about
// the time the pass needs anyway.  A swap can leave as soon as the last pass that targets its local qubit is over; the
// gates of the passes behind it are renamed to the layout it leaves (its local position now holds the qubit that came in,
// which nothing uses before the round is complete) and become steps of their own.  pass_of[i][g] = pass of gate g of step i.
static std::vector<DistStep> split_swap_rounds(const std::vector<DistStep>& in, const std::vector<std::vector<Pass>>& plans,
                                               const std::vector<std::vector<int>>& pass_of) {
    std::vector<DistStep> out;
    const size_t n = in.size();
    size_t i = 0;
    while (i < n) {
        const DistStep& st = in[i];
        size_t j = i + 1;
        bool pure = st.kind == DistStep::LOCAL_GATES;
        if (pure) while (j < n && in[j].kind != DistStep::LOCAL_GATES) { pure = pure && in[j].kind == DistStep::GLOBAL_SWAP; ++j; }
        const int k = (int)(j - i) - 1, m = (int)plans[i].size();
        struct Sw { int gq, lq, last, at; };
        std::vector<Sw> sw;
        bool mapped = pure && pass_of[i].size() == st.gates.size();
        if (mapped) for (int pg : pass_of[i]) if (pg < 0 || pg >= m) mapped = false;      // every gate has its pass
        if (mapped && k >= 2 && m >= 2) {
            for (size_t s2 = i + 1; s2 < j; ++s2) {
                Sw x{in[s2].gq, in[s2].lq, -1, 0};
                for (size_t g = 0; g < st.gates.size(); ++g)
                    if (!st.gates[g].diag && st.gates[g].target() == x.lq) x.last = std::max(x.last, pass_of[i][g]);
                sw.push_back(x);
            }
            std::stable_sort(sw.begin(), sw.end(), [](const Sw& a, const Sw& b) { return a.last > b.last; });
            int pos = m - 1;
            for (Sw& x : sw) { x.at = std::max(std::max(pos, x.last), 0); pos = x.at - 1; }
        }
        bool split = false;
        for (const Sw& x : sw) if (x.at != m - 1) split = true;
        if (!split) {
            for (size_t s2 = i; s2 < j; ++s2) out.push_back(in[s2]);
            i = j;
            continue;
        }
        // sub-steps: the gates of the passes up to each cut (in list order), then the swaps that leave there
        std::vector<HostGate> gates = st.gates;
        int from = 0;
        for (int cut = 0; cut < m; ++cut) {
            bool here = false;
            for (const Sw& x : sw) if (x.at == cut) here = true;
            if (!here && cut != m - 1) continue;
            DistStep sub;
            sub.kind = DistStep::LOCAL_GATES;
            for (size_t g = 0; g < gates.size(); ++g)
                if (pass_of[i][g] >= from && pass_of[i][g] <= cut) sub.gates.push_back(gates[g]);
            out.push_back(std::move(sub));
            for (const Sw& x : sw) {
                if (x.at != cut) continue;
                DistStep s3; s3.kind = DistStep::GLOBAL_SWAP; s3.gq = x.gq; s3.lq = x.lq;
                out.push_back(std::move(s3));
                for (size_t g = 0; g < gates.size(); ++g) {      // the layout behind this swap
                    if (pass_of[i][g] <= cut) continue;
                    for (uint64_t* mask : {&gates[g].tmask, &gates[g].cmask}) {
                        const uint64_t bl = (*mask >> x.lq) & 1ull, bg = (*mask >> x.gq) & 1ull;
                        if (bl != bg) *mask ^= (1ull << x.lq) | (1ull << x.gq);
                    }
                }
            }
            from = cut + 1;
        }
        i = j;
    }
    return out;
}

DistPlan plan_distributed_tuned(const std::vector<HostGate>& gates, int n_total, int n_local, std::vector<int>& perm,
                                bool restore_identity, int store_side, const PlanOptions& opt_in, uint64_t start_zero_mask) {
    static const int thresholds[] = {0, 6, 12, 20, 32};
    const char* env = getenv("DVD_DEFER_TAILS");
    const bool enabled = n_local >= TILE_BITS && !(env && atoi(env) == 0);
    const char* env_split = getenv("DVD_SPLIT_ROUNDS");
    const bool split_rounds = n_local >= TILE_BITS && store_side >= 2 && !(env_split && atoi(env_split) == 0);
    if (n_local < TILE_BITS) store_side = 0;
    DistPlan best;
    std::vector<int> best_perm;
    double best_cost = 0.0;
    bool have = false;
    PlanChoices* const ch = opt_in.choices;
    const bool replay = ch && ch->replay && ch->threshold >= 0;
    PlanChoices best_tape;
    for (int th : thresholds) {
        if (replay) { if (th != thresholds[0]) break; th = ch->threshold; }
        else if (opt_in.defer_max_ops >= 0) { if (th != thresholds[0]) break; th = opt_in.defer_max_ops; }
        else if (th > 0 && !enabled) break;
        // every threshold's run records (or replays) its own sequence of plan_local winners
        PlanChoices tape;
        if (replay) { tape = *ch; tape.pos = 0; }
        PlanOptions opt = opt_in;
        opt.choices = ch ? &tape : nullptr;
        // pass plans of every LOCAL_GATES step, with the zero mask as it evolves; returns the HBM traffic in full passes
        auto plan_steps = [&](DistPlan& d, std::vector<std::vector<int>>* pass_of) {
            double traffic = 0.0;
            d.plans.assign(d.steps.size(), {});
            d.n_passes = 0;
            if (pass_of) pass_of->assign(d.steps.size(), {});
            uint64_t zm = start_zero_mask & ((1ull << n_local) - 1);
            for (size_t i = 0; i < d.steps.size(); ++i) {
                const DistStep& st = d.steps[i];
                if (st.kind == DistStep::LOCAL_GATES && n_local >= TILE_BITS) {
                    PlanOptions o = opt;
                    o.zero_mask = zm;
                    d.plans[i] = plan_local(st.gates, n_local, n_total, o, pass_of ? &(*pass_of)[i] : nullptr);
                    d.n_passes += (int)d.plans[i].size();
                    traffic += plan_traffic(d.plans[i], &zm);
                } else if (st.kind != DistStep::LOCAL_GATES) {      // positions a swap rebuilds are never implied zero
                    if (st.lq >= 0 && st.lq < n_local) zm &= ~(1ull << st.lq);
                    if (st.gq >= 0 && st.gq < n_local) zm &= ~(1ull << st.gq);
                }
            }
            return traffic;
        };
        // Cost of the swap rounds in plain-pass units.  A pass takes about 0.45 of the time the links need for a whole chunk
        // (4 x B200: 13.2 ms against 17.2 GB / 587 GB/s; 8: 6.5 against 14.8; 2: 24.5 against 49), so a round of k swaps,
        // (1 - 2^-k) of a chunk, costs its pass max(0, 2.2 f - 1) passes extra: next to nothing for one swap, 0.65 for two,
        // 0.93 for three; pulling costs a little more than pushing (28.0 against 24.5 ms on 2 x B200), and a round with no
        // pass to ride on needs one of its own.
        auto extra = [](int k) { const double f = 1.0 - std::ldexp(1.0, -k); return std::max(0.0, 2.2 * f - 1.0) + 0.05; };
        auto rounds_cost = [&](const DistPlan& d) {
            double cost = 0.0;
            int waiting = 0;
            for (size_t i = 0; i < d.steps.size(); ++i) {
                const DistStep& st = d.steps[i];
                if (st.kind == DistStep::GLOBAL_SWAP) { ++waiting; continue; }
                if (st.kind != DistStep::LOCAL_GATES) continue;
                if (waiting && (!d.plans[i].empty() || n_local < TILE_BITS)) { cost += extra(waiting) + 0.15; waiting = 0; }
                int k = 0;
                for (auto& sw : d.store[i]) k += sw.first >= n_local;
                if (!d.store[i].empty()) cost += extra(std::max(k, 1));
            }
            if (waiting) cost += 1.0 + extra(waiting);
            return cost;
        };
        DistPlan cand;
        std::vector<int> p;
        double cost = 0.0;
        // with store_side the restore's local transpositions come as LOCAL_SWAP steps; if they cannot ride on the last
        // pass's store after all, the schedule is made again with CNOT triples
        for (int attempt = store_side ? 0 : 1; attempt < 2; ++attempt) {
            cand = DistPlan();
            p = perm;
            cand.steps = plan_distributed(gates, n_total, n_local, p, restore_identity, /*local_swap_steps=*/attempt == 0, &opt, th);
            cand.defer_max_ops = th;
            std::vector<std::vector<int>> pass_of;
            double traffic = plan_steps(cand, split_rounds ? &pass_of : nullptr);
            const int mode = attempt == 0 ? store_side : (store_side >= 2 ? 2 : 0);
            DistPlan alt;
            double alt_traffic = 0.0;
            bool have_alt = false;
            if (split_rounds) {      // the same schedule with its rounds of several swaps spread over several passes
                alt.steps = split_swap_rounds(cand.steps, cand.plans, pass_of);
                if (alt.steps.size() != cand.steps.size()) {
                    alt.defer_max_ops = th;
                    alt_traffic = plan_steps(alt, nullptr);
                    have_alt = assign_store_side(alt, n_local, mode, start_zero_mask);
                }
            }
            if (!assign_store_side(cand, n_local, mode, start_zero_mask)) continue;
            cost = traffic + rounds_cost(cand);
            if (have_alt && alt.n_passes <= cand.n_passes) {      // (never at the price of more passes: the model is too coarse for that trade)
                const double alt_cost = alt_traffic + rounds_cost(alt);
                if (alt_cost < cost - 1e-9) { cand = std::move(alt); cost = alt_cost; }
            }
            break;
        }
        if (!have || cost < best_cost - 1e-9) {
            best = std::move(cand); best_perm = p; best_cost = cost; have = true;
            best_tape = std::move(tape); best_tape.threshold = th;
        }
    }
    if (ch && !replay) { ch->tape = std::move(best_tape.tape); ch->threshold = best_tape.threshold; ch->pos = 0; }
    perm = best_perm;
    return best;
}

}  // namespace dvd
