// tile_core.cuh -- per-thread logic of the register/shared-memory tiled gate kernel.
//
// Everything here is __host__ __device__ so that the exact same index arithmetic and gate
// arithmetic can be replayed thread-by-thread on the CPU (tests/emu) where no GPU exists.
//
// Replaces apply_one_qubit_gate_kernel_local (reference damavand-gpu/kernels.cu:120-172): that
// kernel applies ONE gate per full pass over HBM, one thread per amplitude, in place (racy).
// Here a CTA owns a tile of 2^TILE_BITS amplitudes selected by TILE_BITS arbitrary physical
// qubits ("tile positions"), every thread owns 2^REG_BITS amplitudes in registers, and a whole
// run of gates is applied per pass.  One thread owns both amplitudes of every pair it updates,
// so there is no read/write race by construction.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define DVD_HD __host__ __device__ __forceinline__
#else
#define DVD_HD inline
#endif

namespace dvd {

struct alignas(16) cplx { double x, y; };

constexpr int TILE_BITS = 12;              // 2^12 amplitudes = 64 KiB per CTA tile
constexpr int REG_BITS = 4;                // 16 amplitudes per thread
constexpr int NREG = 1 << REG_BITS;
constexpr int THREAD_BITS = TILE_BITS - REG_BITS;
constexpr int NTHREADS = 1 << THREAD_BITS; // 256
constexpr int NGROUPS = TILE_BITS / REG_BITS;  // 3 register groups: tile positions [4g, 4g+4)
constexpr int IO_GROUP = NGROUPS - 1;      // global loads/stores use the group-2 layout (coalesced)
constexpr int TILE_AMPS = 1 << TILE_BITS;
static_assert(REG_BITS == 4, "register masks are 4 bits wide");

enum OpKind : int32_t {
    K_GENERAL = 0,   // arbitrary complex 2x2
    K_REAL = 1,      // all four entries real (H, RY)
    K_RXLIKE = 2,    // real diagonal, purely imaginary off-diagonal (RX)
    K_DIAG = 3,      // m01 = m10 = 0 (RZ, Z, S, T, and fused parity phases)
    K_ANTIDIAG = 4,  // m00 = m11 = 0 (Y)
    K_SWAP = 5,      // m01 = m10 = 1, m00 = m11 = 0 (X, CNOT): pure exchange
};

// One operation as the device sees it, fully decoded by the planner for the stage it runs in.
//
// Diagonal ops are "parity phases": the amplitude with physical index x is multiplied by
// (parity(x & T) ? m11 : m00) when the op is uncontrolled or parity(x & C) == 1.  T and C are
// given split into the bits that live in the executing thread's registers (tregm / cregm, 4-bit
// masks over the current register group) and all other physical bits (tmask / cmask).  A plain
// RZ / controlled-phase has one-hot masks; wider masks come from fusing CNOT-conjugated runs.
// Non-diagonal ops always target one register bit (treg); their control is one physical qubit,
// either another register bit (cregm one-hot) or a thread-level bit (cmask one-hot).
struct alignas(16) DevOp {
    double m[8];       // m00.re m00.im m01.re m01.im m10.re m10.im m11.re m11.im
    uint64_t tmask;    // diagonal: thread-level part of the target parity mask
    uint64_t cmask;    // thread-level part of the control parity mask (0 = none)
    int32_t kind;
    int8_t group;      // register group the op must run in (non-diagonal), -1 = any
    int8_t treg;       // non-diagonal: target register bit 0..3
    uint8_t tregm;     // diagonal: register-level part of the target parity mask
    uint8_t cregm;     // register-level part of the control parity mask
    int8_t has_ctrl;
    int8_t d0_is_one;  // diagonal with m00 == 1 exactly
    int8_t pad[2];
    int32_t gate_idx;  // caller's gate index (-1 for fused / layout ops)
};
static_assert(sizeof(DevOp) == 96, "DevOp layout");

// Per-launch description of a pass.
struct PassDesc {
    int32_t n_local;              // log2(local amplitudes)
    int32_t n_ops;
    int32_t tile_q[TILE_BITS];    // physical qubit of each tile position
    int32_t sorted_q[TILE_BITS];  // the same qubits in ascending order
    uint64_t rank_bits;           // this rank's value of the global (rank-index) qubits, in place
};

// ---- index helpers ---------------------------------------------------------------------------
// Tile index of register j of thread tid when group g's tile positions live in registers.
DVD_HD int stage_idx(int g, int tid, int j) {
    const int sh = REG_BITS * g;
    const int low = tid & ((1 << sh) - 1);
    const int high = tid >> sh;
    return (high << (sh + REG_BITS)) | (j << sh) | low;
}
// Shared-memory swizzle (16-byte units): makes the group-0 layout (stride-16 lanes) conflict free
// and keeps the other two layouts conflict free.
DVD_HD int swz(int idx) { return idx ^ ((idx >> 4) & 7); }

// Physical offset (in amplitudes) of tile index idx.
DVD_HD uint64_t tile_offset(const PassDesc& pd, int idx) {
    uint64_t off = 0;
#pragma unroll
    for (int p = 0; p < TILE_BITS; ++p) off |= (uint64_t)((idx >> p) & 1) << pd.tile_q[p];
    return off;
}
// Physical base index of CTA `cta`: its bits are deposited into the non-tile positions.
DVD_HD uint64_t cta_base(const PassDesc& pd, uint64_t cta) {
    uint64_t b = cta;
#pragma unroll
    for (int p = 0; p < TILE_BITS; ++p) {
        const int q = pd.sorted_q[p];
        b = ((b >> q) << (q + 1)) | (b & ((1ull << q) - 1));
    }
    return b;
}

// Local index of element h of the half-chunk whose bit lq equals bitval (global<->local qubit swap).
DVD_HD uint64_t half_index(uint64_t h, int lq, int bitval) {
    return ((h >> lq) << (lq + 1)) | ((uint64_t)bitval << lq) | (h & ((1ull << lq) - 1));
}

DVD_HD int parity64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __popcll(v) & 1;
#else
    return __builtin_parityll(v);
#endif
}
DVD_HD int parity4(int v) { return (0x6996 >> (v & 15)) & 1; }

// ---- arithmetic --------------------------------------------------------------------------------
DVD_HD cplx cmul(cplx a, double mr, double mi) { return cplx{a.x * mr - a.y * mi, a.x * mi + a.y * mr}; }

// Per-thread context of the current stage.
struct ThreadCtx {
    uint64_t pidx;   // physical index of register 0 (register bits zero), rank bits included
    cplx ph;         // lazily accumulated scalar phase common to all 16 registers
    bool ph_dirty;
};

DVD_HD void flush_phase(cplx (&a)[NREG], ThreadCtx& ctx) {
    if (!ctx.ph_dirty) return;
#pragma unroll
    for (int j = 0; j < NREG; ++j) a[j] = cmul(a[j], ctx.ph.x, ctx.ph.y);
    ctx.ph = cplx{1.0, 0.0};
    ctx.ph_dirty = false;
}

template <int KIND>
DVD_HD void pair_update(cplx& a0, cplx& a1, const double (&m)[8]) {
    const cplx x = a0, y = a1;
    if (KIND == K_GENERAL) {
        a0 = cplx{x.x * m[0] - x.y * m[1] + y.x * m[2] - y.y * m[3], x.x * m[1] + x.y * m[0] + y.x * m[3] + y.y * m[2]};
        a1 = cplx{x.x * m[4] - x.y * m[5] + y.x * m[6] - y.y * m[7], x.x * m[5] + x.y * m[4] + y.x * m[7] + y.y * m[6]};
    } else if (KIND == K_REAL) {
        a0 = cplx{x.x * m[0] + y.x * m[2], x.y * m[0] + y.y * m[2]};
        a1 = cplx{x.x * m[4] + y.x * m[6], x.y * m[4] + y.y * m[6]};
    } else if (KIND == K_RXLIKE) {   // m00, m11 real; m01 = i*m[3], m10 = i*m[5]
        a0 = cplx{x.x * m[0] - y.y * m[3], x.y * m[0] + y.x * m[3]};
        a1 = cplx{y.x * m[6] - x.y * m[5], y.y * m[6] + x.x * m[5]};
    } else if (KIND == K_ANTIDIAG) {
        a0 = cmul(y, m[2], m[3]);
        a1 = cmul(x, m[4], m[5]);
    } else {                          // K_SWAP
        a0 = y; a1 = x;
    }
}

// All 8 pairs along register bit B; creg >= 0 restricts to the pairs whose register bit creg is 1.
template <int B, int KIND>
DVD_HD void apply_pairs(cplx (&a)[NREG], const double (&m)[8], int creg) {
#pragma unroll
    for (int k = 0; k < NREG / 2; ++k) {
        const int j0 = ((k >> B) << (B + 1)) | (k & ((1 << B) - 1));
        const int j1 = j0 | (1 << B);
        if (creg >= 0 && !((j0 >> creg) & 1)) continue;
        pair_update<KIND>(a[j0], a[j1], m);
    }
}

template <int B>
DVD_HD void apply_kind(cplx (&a)[NREG], int kind, const double (&m)[8], int creg) {
    switch (kind) {
        case K_GENERAL: apply_pairs<B, K_GENERAL>(a, m, creg); break;
        case K_REAL: apply_pairs<B, K_REAL>(a, m, creg); break;
        case K_RXLIKE: apply_pairs<B, K_RXLIKE>(a, m, creg); break;
        case K_ANTIDIAG: apply_pairs<B, K_ANTIDIAG>(a, m, creg); break;
        default: apply_pairs<B, K_SWAP>(a, m, creg); break;
    }
}

// Multiply register j by (bit_j ? d1 : d0) where bit_j = compile-time bit B of j.
template <int B>
DVD_HD void diag_regbit(cplx (&a)[NREG], const double (&m)[8], bool flip, bool skip0) {
    // flip: the thread-level parity is odd, so register bit 0 <-> 1 trade factors
    const double r0 = flip ? m[6] : m[0], i0 = flip ? m[7] : m[1];
    const double r1 = flip ? m[0] : m[6], i1 = flip ? m[1] : m[7];
#pragma unroll
    for (int j = 0; j < NREG; ++j) {
        if ((j >> B) & 1) a[j] = cmul(a[j], r1, i1);
        else if (!skip0) a[j] = cmul(a[j], r0, i0);
    }
}

// Apply one op to the 16 register-resident amplitudes of a thread (stage already matches op.group).
DVD_HD void apply_op(cplx (&a)[NREG], const DevOp& op, ThreadCtx& ctx) {
    double m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = op.m[k];
    const int kind = op.kind;
    const int cregm = op.cregm;
    const bool cthread = op.has_ctrl ? (parity64(ctx.pidx & op.cmask) != 0) : true;   // thread-level control parity
    if (kind == K_DIAG) {
        const int tregm = op.tregm;
        const bool tpar = parity64(ctx.pidx & op.tmask) != 0;
        if (tregm == 0 && cregm == 0) {
            // the whole thread sees one factor: fold it into the lazy scalar phase (4 DFMA instead of 64)
            if (cthread && !(op.d0_is_one && !tpar)) {
                ctx.ph = cmul(ctx.ph, tpar ? m[6] : m[0], tpar ? m[7] : m[1]);
                ctx.ph_dirty = true;
            }
            return;
        }
        if (cregm == 0 && (tregm & (tregm - 1)) == 0) {
            // one register bit in the target parity, control (if any) thread-level: RZ on a register qubit
            if (!cthread) return;
            // with d0 == 1 the untouched half is the one whose overall parity is even
            const bool d0one = op.d0_is_one != 0;
            switch (tregm) {
                case 1: diag_regbit<0>(a, m, tpar, d0one && !tpar); break;
                case 2: diag_regbit<1>(a, m, tpar, d0one && !tpar); break;
                case 4: diag_regbit<2>(a, m, tpar, d0one && !tpar); break;
                default: diag_regbit<3>(a, m, tpar, d0one && !tpar); break;
            }
            return;
        }
        // general parity phase with register bits in target and/or control masks
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            const bool on = op.has_ctrl ? (cthread != (parity4(j & cregm) != 0)) : true;
            const bool bit = tpar != (parity4(j & tregm) != 0);
            if (on && !(op.d0_is_one && !bit)) a[j] = cmul(a[j], bit ? m[6] : m[0], bit ? m[7] : m[1]);
        }
        return;
    }
    // non-diagonal: one target register bit; control none / thread-level / one register bit
    int creg = -1;
    if (cregm) creg = cregm == 1 ? 0 : cregm == 2 ? 1 : cregm == 4 ? 2 : 3;
    else if (!cthread) return;
    switch (op.treg) {
        case 0: apply_kind<0>(a, kind, m, creg); break;
        case 1: apply_kind<1>(a, kind, m, creg); break;
        case 2: apply_kind<2>(a, kind, m, creg); break;
        default: apply_kind<3>(a, kind, m, creg); break;
    }
}

}  // namespace dvd
