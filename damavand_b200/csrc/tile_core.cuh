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
//
// The op stream is a small instruction set (OpCode) decoded once per op with ONE dense switch;
// every case is straight-line code over compile-time register indices.  Diagonal gates never
// run as gates: the planner folds them into a phase polynomial per pass
//     K * prod_q a_q^{x_q} * prod_{q<q'} b_qq'^{x_q x_q'}
// and emits it as byte-indexed phase tables (thread-level bits), per-register-bit factors and
// register-pair factors, at the latest point the commutation rules allow.
#pragma once
#ifdef __CUDACC_RTC__
// run-time compilation (jit_rt.cpp): no system headers; fixed-width types as on LP64 hosts (same sizes)
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <math.h>
#include <stdint.h>
#endif

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
constexpr int IO_GROUP = NGROUPS - 1;      // global loads use the group-2 layout; stores the group-2 or group-1 layout
                                           // (in both, lanes cover tile positions 0..2 = 128 contiguous bytes)
constexpr int TILE_AMPS = 1 << TILE_BITS;
constexpr int TABLE_ENTRIES = 256;         // one phase sub-table per byte of the physical index
constexpr int MAX_INDEX_BYTES = 8;
static_assert(REG_BITS == 4, "register masks are 4 bits wide");

// Matrix classes the planner distinguishes (exact zero / one tests on the entries).
enum OpKind : int32_t {
    K_GENERAL = 0,   // arbitrary complex 2x2
    K_REAL = 1,      // all four entries real (H, RY)
    K_RXLIKE = 2,    // real diagonal, purely imaginary off-diagonal (RX)
    K_ANTIDIAG = 3,  // m00 = m11 = 0 (Y)
    K_HADAMARD = 4,  // h * [[1, 1], [1, -1]], uncontrolled: add / subtract only, h joins the pass constant
    K_REALPH = 5,    // real 2x2 followed by diag(1, w) on the same qubit (RY then RZ): w = (m[1], m[3]); planner peephole
    K_SWAP = 6,      // m01 = m10 = 1, m00 = m11 = 0 (X, CNOT): pure exchange
    K_DIAG = 7,      // m01 = m10 = 0 (RZ, Z, S, T, fused parity phases)
};

// Device instruction set.
enum OpCode : int32_t {
    OC_GATE = 0,        // + kind*4 + treg (kind 0..5): 2x2 on register bit treg; control none / thread-level
    OC_CGEN = 28,       // + treg: general 2x2 on treg, control = register bit op.creg (rare)
    OC_DIAG1 = 36,      // + r: registers with bit r set *= (m[6], m[7])
    OC_PHASE = 40,      // thread-level parity phase -> lazy per-thread scalar
    OC_DIAGGEN = 41,    // generic parity phase with register bits in its masks (fallback)
    OC_TABLE = 42,      // product of phase-table lookups -> lazy scalar (optional thread-level pivot)
    OC_TABLE_REG = 43,  // + r: (table lookups | constant) x factors of the other register bits -> registers with bit r
    OC_PAIR = 47,       // + pair id: registers with both bits set *= (m[0], m[1])
    OC_TWHAD = 53,      // + r: OC_TABLE_REG + r immediately followed by the (uncontrolled) Hadamard on r: the
                        //   twiddle-then-butterfly step of a QFT, one dispatch instead of two
    OC_REALPH4 = 57,    // macro-op: the next four ops are uncontrolled K_REALPH gates on register bits 0,1,2,3 (one dispatch)
    OC_TWHAD4 = 58,     // macro-op: the next four ops are OC_TWHAD + 0,1,2,3 in this order (a radix-16 QFT butterfly)
    OC_SWITCH = 59,     // + from*3 + to: transpose through shared memory, optionally with a GF(2)-affine
                        //   permutation of the tile index (all pending X / CNOT gates)
    OC_COUNT = 68,
};
// Op classes: k_tile_pass is compiled for a few class subsets (fewer live registers, shorter decode,
// smaller code) and a pass runs on the smallest variant that covers the classes of its ops.
enum OpClass : unsigned {
    C_GENERAL = 1u << 0,   // K_GENERAL gates
    C_REAL = 1u << 1,      // K_REAL / K_REALPH gates
    C_RX = 1u << 2,        // K_RXLIKE gates
    C_HAD = 1u << 3,       // K_HADAMARD gates
    C_RARE = 1u << 4,      // K_ANTIDIAG, OC_CGEN, OC_DIAGGEN
    C_DIAG = 1u << 5,      // OC_DIAG1, OC_PHASE, OC_PAIR
    C_TABLE = 1u << 6,     // OC_TABLE, OC_TABLE_REG, OC_TWHAD
    C_MACRO_R = 1u << 7,   // OC_REALPH4
    C_MACRO_T = 1u << 8,   // OC_TWHAD4
    C_ALL = (1u << 9) - 1,
};
DVD_HD unsigned op_class(int code);

DVD_HD int pair_id(int r0, int r1) {   // r0 < r1
    return r0 == 0 ? r1 - 1 : r0 == 1 ? r1 + 1 : 5;
}

enum OpFlags : uint8_t {
    F_D0_ONE = 1,     // diagonal op with d0 == 1
    F_HAS_CTRL = 2,   // OC_DIAGGEN: controlled
    F_TCTRL = 4,      // thread-level control parity mask in cmask must be odd
    F_PERM = 8,       // OC_SWITCH: PermPayload in m[]
    F_PM_SHIFT = 4,   // OC_TABLE_REG: bits 4..6 = which of the other three register bits carry a factor
    F_TABLE = 128,    // table ops: the op has a phase table (else OC_TABLE_REG uses the constant m[6..7])
};

// One operation as the device sees it, fully decoded by the planner for the stage it runs in.
// The ops of a pass travel as a kernel PARAMETER (PassParams, constant bank): every warp reads the
// same op at the same time, so operands come through the constant cache / uniform datapath and
// never touch shared memory or its instruction queue.
struct alignas(16) DevOp {
    double m[8];       // m00.re m00.im m01.re m01.im m10.re m10.im m11.re m11.im  (or a PermPayload)
    uint64_t tmask;    // thread-level part of the target parity mask (diagonal ops, table pivot)
    uint64_t cmask;    // thread-level part of the control parity mask (0 = none)
    int32_t code;      // OpCode (+ operands)
    int32_t tab;       // table ops: index of the op's table record in the pass (see TableDesc)
    int32_t gate_idx;  // caller's gate index (-1 for fused / layout ops)
    int8_t group;      // register group the op runs in (OC_SWITCH: the group it switches to)
    int8_t creg;       // OC_CGEN: control register bit
    uint8_t regm;      // OC_DIAGGEN: tregm | cregm << 4
    uint8_t flags;     // OpFlags
};
static_assert(sizeof(DevOp) == 96, "DevOp layout");

// Phase tables of a pass (all entries complex128), one buffer per pass:
//   [n_tab TableDesc records, 16 B each]
//   [n_tab x TABLE_TILE_ENTRIES: factor by thread-index bits 0..3 (16 entries) and 4..7 (16 entries)]
//   [byte tables: 256 entries per set bit of TableDesc::bytes, indexed by byte b of the CTA's physical base]
// The thread-index factors depend only on the thread's position in the tile (read through L1, the
// same 512 B for every CTA); the byte tables only see index bits outside the tile, so they are
// reduced to ONE constant per CTA and table op in the kernel prologue.
constexpr int TABLE_TILE_ENTRIES = 32;
constexpr int MAX_TABLE_OPS = 128;         // per pass (shared-memory array of per-CTA constants)
struct alignas(16) TableDesc {
    uint32_t byte_off;   // start of the op's byte tables, in cplx entries from the start of the buffer
    uint32_t bytes;      // bit b: byte b of the physical index has a sub-table
    uint32_t pad[2];
};
static_assert(sizeof(TableDesc) == sizeof(cplx), "TableDesc is one table slot");

// Payload of a permuting OC_SWITCH, stored over DevOp::m.  The amplitude at tile index i moves to
//   i' = xor_{p : bit p of i} col[p]  ^  v0  ^  xor_k [parity(physical base & cond_mask_k)] cond_vec[k]
// (X on a tile qubit flips a bit of v0; CNOT between tile qubits adds a row of the matrix to another;
// CNOT controlled by a qubit outside the tile is a per-CTA conditional flip).  cond_mask 0 and 1
// live in DevOp::tmask / DevOp::cmask, 2 and 3 in the payload.
constexpr int PERM_MAX_COND = 4;
struct PermPayload {
    uint16_t col[TILE_BITS];
    uint16_t v0;
    uint16_t n_cond;
    uint16_t cond_vec[PERM_MAX_COND];
    uint32_t pad;
    uint64_t cond_mask23[2];
};
static_assert(sizeof(PermPayload) <= 64, "PermPayload must fit in DevOp::m");

// Per-launch description of a pass.
struct PassDesc {
    int32_t n_local;              // log2(local amplitudes)
    int32_t n_ops;
    int32_t tile_q[TILE_BITS];    // physical qubit of each tile position
    int32_t sorted_q[TILE_BITS];  // the same qubits in ascending order
    uint64_t rank_bits;           // this rank's value of the global (rank-index) qubits, in place
    const cplx* tables;           // phase tables of this pass (device pointer; host pointer in the replay)
    int32_t n_tab;                // table ops in this pass (<= MAX_TABLE_OPS)
    int32_t io_out;               // register group whose layout the tile is stored from (1 or 2; loads use IO_GROUP)
    // Index arithmetic precomputed by the planner (the bit-by-bit loops below cost ~17 % of the kernel's
    // instructions when every thread runs them at every stage switch):
    const uint64_t* tid_off;      // [NGROUPS][NTHREADS]: tile_offset(stage_idx(g, tid, 0)), device pointer
    int8_t n_runs;                // cta_base as runs of non-tile bits: bits [src, src+len) of the CTA index
    int8_t run_src[TILE_BITS + 1];    //   land at bit position dst of the physical index
    int8_t run_len[TILE_BITS + 1];
    int8_t run_dst[TILE_BITS + 1];
    // Set by the engine / launcher, not by the planner:
    int8_t n_cta_bits;            // log2(tiles the launch covers): non-tile bits that can be 1 on input (see zero_mask)
    int8_t zero_regbits;          // register bits of the load layout (IO_GROUP) whose qubit is in zero_mask
    int16_t last_switch;          // index of the last OC_SWITCH in ops (-1: none); after it the shared-memory tile is free
    uint64_t zero_mask;           // local qubits that are still |0> in every populated basis state (support tracking after
                                  //   a reset): amplitudes with one of these bits set are zero by construction, are never
                                  //   read, and tiles whose fixed bits hit the mask are not launched at all.  0 = dense state
                                  //   (with a fused remap: a mask over the SOURCE index, i.e. the layout before the swap)
    // Fused global<->local qubit remap (distributed_gpu, set by the engine from planner.h: compose_remap): a SEQUENCE of
    // swaps between rank-index qubits and local qubits -- any bit permutation over at most MAX_REMAP rank-index and
    // MAX_REMAP local positions -- is executed by this pass's LOAD instead of by exchanges of their own.  The pass then
    // runs out of place: the amplitude at (new) local index i is read from buffer remap_src[sel(i)], with
    //     sel(i) = sum_k bit(remap_lq[k] of i) << k                       (which rank held it: remap_n <= 3 index bits)
    // at local index
    //     (i & ~remap_lmask) | remap_const | sum_m bit(remap_mv_from[m] of i) << remap_mv_to[m]
    // (the local positions involved are rebuilt from this rank's own rank bits -- a constant -- or from other local
    // bits of i).  remap_src[sel] is this rank's own input buffer or a partner rank's (peer memory over NVLink, mapped
    // with CUDA IPC).  remap_on == 0: plain in-place pass.  Replaces exchange_amplitudes_between_gpus + the distributed
    // gate kernel (rust_communication.cu:106-141, kernels.cu:174-230): the exchange IS the next pass's read.
    // remap_st: the same fields describe the STORE instead (push): the amplitude this pass computes for (old) local index
    // i is written to buffer remap_src[sel(i)] -- the second chunk of this rank or of a partner -- at the index above.
    // Used for the swaps that END a gate list (the layout restore): they ride on the last gate pass instead of needing a
    // pass of their own, and remote writes need no response traffic on the links.  At most one of remap_on / remap_st.
    int8_t remap_on, remap_st;
    int8_t remap_n;
    int8_t remap_lq[3];
    int8_t remap_n_mv;
    int8_t remap_mv_from[6], remap_mv_to[6];
    uint64_t remap_lmask, remap_const;
    const cplx* remap_src[8];
};
constexpr int MAX_REMAP = 3;         // rank-index positions (and selector bits) of one fused remap
constexpr int MAX_REMAP_LOCAL = 6;   // local positions of one fused remap
// Kernel parameter block: the pass description and its whole op list (<= 32764 B of parameters).
constexpr int MAX_OPS_PER_PASS = 336;
struct PassParams {
    PassDesc pd;
    DevOp ops[MAX_OPS_PER_PASS];
};
static_assert(sizeof(PassParams) + 16 <= 32764, "kernel parameter space");

// ---- index helpers ---------------------------------------------------------------------------
// Tile index of register j of thread tid when group g's tile positions live in registers.
DVD_HD int stage_idx(int g, int tid, int j) {
    const int sh = REG_BITS * g;
    const int low = tid & ((1 << sh) - 1);
    const int high = tid >> sh;
    return (high << (sh + REG_BITS)) | (j << sh) | low;
}
// Shared-memory slot (16-byte units) of tile index idx: one pad slot per 16 makes the group-0 layout
// (lanes 16 slots apart) conflict free, keeps the other two conflict free, and -- unlike an XOR
// swizzle -- stays additive: slot(base + (j << 4g)) = slot(base) + const(g, j), so every register's
// address is an immediate offset from one per-thread base.
constexpr int TILE_SLOTS = TILE_AMPS + TILE_AMPS / 16;
DVD_HD int smem_slot(int idx) { return idx + (idx >> 4); }

// Two-group persistent form (tile_kernel.cuh): dynamic shared memory =
// [RING_BUFFERS tiles][per-group, double-buffered table constants][mbarriers]
constexpr int RING_BUFFERS = 3;
constexpr int RING_GROUPS = 2;
constexpr int RING_WC_BYTES = RING_GROUPS * 2 * MAX_TABLE_OPS * (int)sizeof(cplx);
constexpr int RING_SMEM_BYTES = RING_BUFFERS * TILE_SLOTS * (int)sizeof(cplx) + RING_WC_BYTES + 64 + NGROUPS * NTHREADS * 4;

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

// The same from the run-compressed form (PassDesc::run_*): a handful of shifts instead of TILE_BITS steps.
DVD_HD uint64_t cta_base_runs(const PassDesc& pd, uint64_t cta) {
    uint64_t b = 0;
    for (int r = 0; r < pd.n_runs; ++r)
        b |= ((cta >> pd.run_src[r]) & ((1ull << pd.run_len[r]) - 1)) << pd.run_dst[r];
    return b;
}
#ifndef __CUDACC_RTC__
// Host: fill PassDesc::run_* from sorted_q / n_local.
inline void fill_cta_runs(PassDesc& pd) {
    int n = 0, src = 0, prev = 0;    // prev: first physical bit not yet covered
    for (int p = 0; p <= TILE_BITS; ++p) {
        const int q = p < TILE_BITS ? pd.sorted_q[p] : pd.n_local;
        const int len = q - prev;    // non-tile bits [prev, q)
        if (len > 0) { pd.run_src[n] = (int8_t)src; pd.run_len[n] = (int8_t)len; pd.run_dst[n] = (int8_t)prev; ++n; src += len; }
        prev = q + 1;
    }
    pd.n_runs = (int8_t)n;
    pd.n_cta_bits = (int8_t)src;
}
// Host: the same for a launch that only covers the tiles whose fixed bits avoid `skip` (bits known to be 0 in every
// populated basis state).  Returns false when the run list would not fit (then the caller launches the full grid).
inline bool fill_cta_runs_sparse(PassDesc& pd, uint64_t skip) {
    uint64_t tile = 0;
    for (int p = 0; p < TILE_BITS; ++p) tile |= 1ull << pd.tile_q[p];
    int n = 0, src = 0, q = 0;
    int8_t rs[TILE_BITS + 1], rl[TILE_BITS + 1], rd[TILE_BITS + 1];
    while (q < pd.n_local) {
        if (((tile | skip) >> q) & 1ull) { ++q; continue; }
        int len = 0;
        while (q + len < pd.n_local && !(((tile | skip) >> (q + len)) & 1ull)) ++len;
        if (n > TILE_BITS) return false;   // run arrays hold TILE_BITS + 1 entries
        rs[n] = (int8_t)src; rl[n] = (int8_t)len; rd[n] = (int8_t)q; ++n; src += len; q += len;
    }
    for (int r = 0; r < n; ++r) { pd.run_src[r] = rs[r]; pd.run_len[r] = rl[r]; pd.run_dst[r] = rd[r]; }
    pd.n_runs = (int8_t)n;
    pd.n_cta_bits = (int8_t)src;
    return true;
}
// Host: the general form.  Bits in `skip` are fixed to zero (tiles that are zero by construction are not launched);
// the non-tile bits in `first` take the LOWEST bits of the CTA index.  A pass whose load carries a fused remap puts
// the bits that select the source rank there: tiles fetched over NVLink and tiles fetched from local HBM then
// alternate in launch order, so the links are busy for the whole pass instead of for its second half only.
inline bool fill_cta_runs_ex(PassDesc& pd, uint64_t skip, uint64_t first) {
    uint64_t tile = 0;
    for (int p = 0; p < TILE_BITS; ++p) tile |= 1ull << pd.tile_q[p];
    first &= ~(tile | skip) & ((1ull << pd.n_local) - 1);
    int n = 0, src = 0;
    int8_t rs[TILE_BITS + 1], rl[TILE_BITS + 1], rd[TILE_BITS + 1];
    for (int q = 0; q < pd.n_local; ++q)
        if ((first >> q) & 1ull) {
            if (n > TILE_BITS) return false;
            rs[n] = (int8_t)src; rl[n] = 1; rd[n] = (int8_t)q; ++n; ++src;
        }
    int q = 0;
    const uint64_t taken = tile | skip | first;
    while (q < pd.n_local) {
        if ((taken >> q) & 1ull) { ++q; continue; }
        int len = 0;
        while (q + len < pd.n_local && !((taken >> (q + len)) & 1ull)) ++len;
        if (n > TILE_BITS) return false;
        rs[n] = (int8_t)src; rl[n] = (int8_t)len; rd[n] = (int8_t)q; ++n; src += len; q += len;
    }
    for (int r = 0; r < n; ++r) { pd.run_src[r] = rs[r]; pd.run_len[r] = rl[r]; pd.run_dst[r] = rd[r]; }
    pd.n_runs = (int8_t)n;
    pd.n_cta_bits = (int8_t)src;
    return true;
}
#endif  // !__CUDACC_RTC__

// Source of the amplitude at (new) local index i under a fused remap (PassDesc::remap_*): buffer and local index.
DVD_HD unsigned remap_sel(const PassDesc& pd, uint64_t i) {
    unsigned sel = 0;
#pragma unroll
    for (int k = 0; k < MAX_REMAP; ++k)
        if (k < pd.remap_n) sel |= (unsigned)((i >> pd.remap_lq[k]) & 1ull) << k;
    return sel;
}
DVD_HD uint64_t remap_index(const PassDesc& pd, uint64_t i) {
    uint64_t src = (i & ~pd.remap_lmask) | pd.remap_const;
#pragma unroll
    for (int m = 0; m < MAX_REMAP_LOCAL; ++m)
        if (m < pd.remap_n_mv) src |= ((i >> pd.remap_mv_from[m]) & 1ull) << pd.remap_mv_to[m];
    return src;
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
DVD_HD cplx cmul(cplx a, cplx b) { return cmul(a, b.x, b.y); }

// Per-thread context of the current stage.
struct ThreadCtx {
    uint64_t pidx;   // physical index of register 0 (register bits zero), rank bits included
    cplx ph;         // lazily accumulated scalar phase common to all 16 registers
    bool ph_dirty;
    int tid;         // thread index inside the CTA
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
    if (KIND == K_HADAMARD) {        // the common factor h is folded into the pass constant by the planner
        // in place, no temporary: a1 = (x + y) - 2y with one rounding in the fma.  |error| <= 2^-53 (|x+y| + |x-y|),
        // and the register-to-register moves that the x - y form needs at every butterfly (14 % of the
        // instructions of a Fourier pass, profiles/r1_ncu_mix_qft30_v8.txt) disappear
        a0 = cplx{x.x + y.x, x.y + y.y};
        a1 = cplx{fma(-2.0, y.x, a0.x), fma(-2.0, y.y, a0.y)};
    } else if (KIND == K_REALPH) {   // real rows, then row 1 times w = (m[1], m[3])
        a0 = cplx{x.x * m[0] + y.x * m[2], x.y * m[0] + y.y * m[2]};
        const cplx t{x.x * m[4] + y.x * m[6], x.y * m[4] + y.y * m[6]};
        a1 = cmul(t, m[1], m[3]);
    } else if (KIND == K_GENERAL) {
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

// All 8 pairs along register bit B.
template <int B, int KIND>
DVD_HD void gate_all(cplx (&a)[NREG], const double* mp) {
    double m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = KIND == K_HADAMARD ? 0.0 : mp[k];
#pragma unroll
    for (int k = 0; k < NREG / 2; ++k) {
        const int j0 = ((k >> B) << (B + 1)) | (k & ((1 << B) - 1));
        pair_update<KIND>(a[j0], a[j0 | (1 << B)], m);
    }
}
// General 2x2 along B on the pairs whose register bit creg (runtime) is set.
template <int B>
DVD_HD void cgen(cplx (&a)[NREG], const double* mp, int creg) {
    double m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = mp[k];
#pragma unroll
    for (int k = 0; k < NREG / 2; ++k) {
        const int j0 = ((k >> B) << (B + 1)) | (k & ((1 << B) - 1));
        if ((j0 >> creg) & 1) pair_update<K_GENERAL>(a[j0], a[j0 | (1 << B)], m);
    }
}

// Multiply the registers with bit B set by d1 (the planner normalises d0 to 1 and keeps it in the pass constant).
template <int B>
DVD_HD void diag_regbit(cplx (&a)[NREG], const double* m) {
    const double r1 = m[6], i1 = m[7];
#pragma unroll
    for (int j = 0; j < NREG; ++j)
        if ((j >> B) & 1) a[j] = cmul(a[j], r1, i1);
}
template <int B0, int B1>
DVD_HD void scale_pair(cplx (&a)[NREG], double wr, double wi) {
#pragma unroll
    for (int j = 0; j < NREG; ++j) if (((j >> B0) & 1) && ((j >> B1) & 1)) a[j] = cmul(a[j], wr, wi);
}

// CTA-constant part of table op `ti`: product of its byte sub-tables at the CTA's physical base index.
DVD_HD cplx table_cta_const(const cplx* __restrict__ tables, int ti, uint64_t gbase) {
    const TableDesc d = reinterpret_cast<const TableDesc*>(tables)[ti];
    cplx w{1.0, 0.0};
    bool first = true;
    const cplx* t = tables + d.byte_off;
#pragma unroll
    for (int b = 0; b < MAX_INDEX_BYTES; ++b) {
        if ((d.bytes >> b) & 1u) {
            const cplx e = t[(unsigned)(gbase >> (8 * b)) & 255u];
            w = first ? e : cmul(w, e.x, e.y);
            first = false;
            t += TABLE_ENTRIES;
        }
    }
    return w;
}
// Thread-index factors of table op `ti` inside a pass with n_tab table ops.
DVD_HD const cplx* table_tile(const cplx* tables, int n_tab, int ti) {
    return tables + n_tab + (size_t)ti * TABLE_TILE_ENTRIES;
}
// Full table value of a thread: CTA constant x the two thread-index factors.
DVD_HD cplx table_value(const cplx* tl, cplx wc, int tid) {
#ifdef __CUDA_ARCH__
    const double2 lo = __ldg(reinterpret_cast<const double2*>(tl) + (tid & 15));
    const double2 hi = __ldg(reinterpret_cast<const double2*>(tl) + 16 + (tid >> 4));
#else
    const cplx lo = tl[tid & 15], hi = tl[16 + (tid >> 4)];
#endif
    return cmul(cmul(wc, lo.x, lo.y), hi.x, hi.y);
}

// Registers with bit B set *= W * prod_k f_k^{bit o_k of j}: W from the table (or the constant
// m[6..7] when the op has none), f_k = m[2k..2k+1] for the other three register bits o_0<o_1<o_2.
template <int B>
DVD_HD void table_reg(cplx (&a)[NREG], const DevOp& op, unsigned flags, const ThreadCtx& ctx, const cplx* tl, cplx wc) {
    constexpr int O0 = B == 0 ? 1 : 0, O1 = B <= 1 ? 2 : 1, O2 = B <= 2 ? 3 : 2;
    const unsigned pm = (flags >> F_PM_SHIFT) & 7u;
    const cplx w = (flags & F_TABLE) ? table_value(tl, wc, ctx.tid) : cplx{op.m[6], op.m[7]};
    cplx w4[4];
    w4[0] = w;
    w4[1] = (pm & 1u) ? cmul(w, op.m[0], op.m[1]) : w;
    w4[2] = (pm & 2u) ? cmul(w4[0], op.m[2], op.m[3]) : w4[0];
    w4[3] = (pm & 2u) ? cmul(w4[1], op.m[2], op.m[3]) : w4[1];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (h == 1 && (pm & 4u)) {
#pragma unroll
            for (int s = 0; s < 4; ++s) w4[s] = cmul(w4[s], op.m[4], op.m[5]);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int j = (1 << B) | ((s & 1) << O0) | (((s >> 1) & 1) << O1) | (h << O2);
            a[j] = cmul(a[j], w4[s].x, w4[s].y);
        }
    }
}

// ---- stage switch helpers (shared by the kernel and the CPU replay) ------------------------------
// Physical index of register 0 of thread tid in group g's layout (register bits zero).
DVD_HD uint64_t thread_pidx(const PassDesc& pd, uint64_t gbase, int g, int tid) {
    return gbase | tile_offset(pd, stage_idx(g, tid, 0));
}
// The same through the planner's table (device: read-only data path).
DVD_HD uint64_t tid_offset(const PassDesc& pd, int g, int tid) {
#ifdef __CUDA_ARCH__
    return __ldg(reinterpret_cast<const unsigned long long*>(pd.tid_off) + g * NTHREADS + tid);
#else
    return pd.tid_off[g * NTHREADS + tid];
#endif
}
// The same by arithmetic over the thread's 8 index bits (shift amounts from the constant bank, ~25 integer
// instructions, nothing waits on memory): used where the offset is needed at once -- in front of the tile's loads.
DVD_HD uint64_t tid_offset_arith(const PassDesc& pd, int g, int tid) {
    uint64_t off = 0;
    const int sh = REG_BITS * g;
#pragma unroll
    for (int k = 0; k < THREAD_BITS; ++k) {
        const int p = k < sh ? k : k + REG_BITS;      // tile position of thread-index bit k in group g's layout
        off |= (uint64_t)((tid >> k) & 1) << pd.tile_q[p];
    }
    return off;
}
// Constant part of a permuting switch for one CTA: v0 ^ conditional flips.
DVD_HD unsigned perm_const(const DevOp& op, uint64_t gbase) {
    const PermPayload& pp = *reinterpret_cast<const PermPayload*>(op.m);
    unsigned v = pp.v0;
    const int nc = pp.n_cond;
    if (nc > 0 && parity64(gbase & op.tmask)) v ^= pp.cond_vec[0];
    if (nc > 1 && parity64(gbase & op.cmask)) v ^= pp.cond_vec[1];
    if (nc > 2 && parity64(gbase & pp.cond_mask23[0])) v ^= pp.cond_vec[2];
    if (nc > 3 && parity64(gbase & pp.cond_mask23[1])) v ^= pp.cond_vec[3];
    return v;
}
// Image of tile index idx under the permutation (v = perm_const).
DVD_HD unsigned perm_index(const DevOp& op, unsigned v, unsigned idx) {
    const PermPayload& pp = *reinterpret_cast<const PermPayload*>(op.m);
    unsigned r = v;
#pragma unroll
    for (int p = 0; p < TILE_BITS; ++p) if ((idx >> p) & 1u) r ^= pp.col[p];
    return r;
}

#define DVD_CASE4(cls, base, STMT)        \
    case (base) + 0: if constexpr ((SET & (cls)) != 0) { constexpr int B = 0; STMT; } break; \
    case (base) + 1: if constexpr ((SET & (cls)) != 0) { constexpr int B = 1; STMT; } break; \
    case (base) + 2: if constexpr ((SET & (cls)) != 0) { constexpr int B = 2; STMT; } break; \
    case (base) + 3: if constexpr ((SET & (cls)) != 0) { constexpr int B = 3; STMT; } break;

// Twiddle on the registers with bit B (table_reg) followed by the Hadamard butterfly along B.
template <int B>
DVD_HD void twhad(cplx (&a)[NREG], const DevOp& op, unsigned flags, const ThreadCtx& ctx, const cplx* tables, int n_tab, const cplx* wcs) {
    table_reg<B>(a, op, flags, ctx, (flags & F_TABLE) ? table_tile(tables, n_tab, op.tab) : tables,
                 (flags & F_TABLE) ? wcs[op.tab] : cplx{1.0, 0.0});
    gate_all<B, K_HADAMARD>(a, op.m);
}

// Apply the op at opk[0] (anything but OC_SWITCH) to the 16 register-resident amplitudes of a thread.
// tables / n_tab: the pass's table buffer; wcs: per-CTA constants of its table ops (kernel prologue).
// Returns how many FOLLOWING ops the op consumed (macro-ops run opk[0..3] in one dispatch).
// SET: the op classes this instantiation can execute (the others compile to nothing).
// code / flags are opk[0].code / opk[0].flags: run-time values in the interpreter kernels, literals in the
// structure-specialised kernels (jit.cpp), where the switch and the flag tests fold away.
template <unsigned SET = C_ALL>
DVD_HD int apply_op(cplx (&a)[NREG], const DevOp* opk, int code, unsigned flags, ThreadCtx& ctx, const cplx* tables, int n_tab, const cplx* wcs) {
    const DevOp& op = opk[0];
    if ((flags & F_TCTRL) && !parity64(ctx.pidx & op.cmask)) return 0;   // thread-level control
    const double* m = op.m;
    switch (code) {
        DVD_CASE4(C_GENERAL, OC_GATE + 4 * K_GENERAL, (gate_all<B, K_GENERAL>(a, m)))
        DVD_CASE4(C_REAL, OC_GATE + 4 * K_REAL, (gate_all<B, K_REAL>(a, m)))
        DVD_CASE4(C_RX, OC_GATE + 4 * K_RXLIKE, (gate_all<B, K_RXLIKE>(a, m)))
        DVD_CASE4(C_RARE, OC_GATE + 4 * K_ANTIDIAG, (gate_all<B, K_ANTIDIAG>(a, m)))
        DVD_CASE4(C_HAD, OC_GATE + 4 * K_HADAMARD, (gate_all<B, K_HADAMARD>(a, m)))
        DVD_CASE4(C_REAL, OC_GATE + 4 * K_REALPH, (gate_all<B, K_REALPH>(a, m)))
        DVD_CASE4(C_RARE, OC_CGEN, (cgen<B>(a, m, op.creg)))
        DVD_CASE4(C_DIAG, OC_DIAG1, (diag_regbit<B>(a, m)))
        case OC_PHASE: if constexpr ((SET & C_DIAG) != 0) {
            const bool tpar = parity64(ctx.pidx & op.tmask) != 0;
            if (!((flags & F_D0_ONE) && !tpar)) {
                ctx.ph = cmul(ctx.ph, tpar ? m[6] : m[0], tpar ? m[7] : m[1]);
                ctx.ph_dirty = true;
            }
        } break;
        case OC_DIAGGEN: if constexpr ((SET & C_RARE) != 0) {   // control handled here: thread-level and register-level parts combine
            const int tregm = op.regm & 15, cregm = op.regm >> 4;
            const bool has_ctrl = (flags & F_HAS_CTRL) != 0;
            const bool cpar = op.cmask != 0 && parity64(ctx.pidx & op.cmask) != 0;
            const bool tpar = parity64(ctx.pidx & op.tmask) != 0;
            const bool d0one = (flags & F_D0_ONE) != 0;
#pragma unroll
            for (int j = 0; j < NREG; ++j) {
                const bool on = has_ctrl ? (cpar != (parity4(j & cregm) != 0)) : true;
                const bool bit = tpar != (parity4(j & tregm) != 0);
                if (on && !(d0one && !bit)) a[j] = cmul(a[j], bit ? m[6] : m[0], bit ? m[7] : m[1]);
            }
        } break;
        case OC_TABLE: if constexpr ((SET & C_TABLE) != 0) {
            if (op.tmask == 0 || parity64(ctx.pidx & op.tmask)) {
                const cplx w = table_value(table_tile(tables, n_tab, op.tab), wcs[op.tab], ctx.tid);
                ctx.ph = cmul(ctx.ph, w.x, w.y);
                ctx.ph_dirty = true;
            }
        } break;
        DVD_CASE4(C_TABLE, OC_TABLE_REG, (table_reg<B>(a, op, flags, ctx, (flags & F_TABLE) ? table_tile(tables, n_tab, op.tab) : tables,
                                              (flags & F_TABLE) ? wcs[op.tab] : cplx{1.0, 0.0})))
        DVD_CASE4(C_TABLE, OC_TWHAD, (twhad<B>(a, op, flags, ctx, tables, n_tab, wcs)))
        case OC_PAIR + 0: if constexpr ((SET & C_DIAG) != 0) scale_pair<0, 1>(a, m[0], m[1]); break;
        case OC_PAIR + 1: if constexpr ((SET & C_DIAG) != 0) scale_pair<0, 2>(a, m[0], m[1]); break;
        case OC_PAIR + 2: if constexpr ((SET & C_DIAG) != 0) scale_pair<0, 3>(a, m[0], m[1]); break;
        case OC_PAIR + 3: if constexpr ((SET & C_DIAG) != 0) scale_pair<1, 2>(a, m[0], m[1]); break;
        case OC_PAIR + 4: if constexpr ((SET & C_DIAG) != 0) scale_pair<1, 3>(a, m[0], m[1]); break;
        case OC_PAIR + 5: if constexpr ((SET & C_DIAG) != 0) scale_pair<2, 3>(a, m[0], m[1]); break;
        case OC_REALPH4: if constexpr ((SET & C_MACRO_R) != 0) {
            gate_all<0, K_REALPH>(a, opk[0].m); gate_all<1, K_REALPH>(a, opk[1].m);
            gate_all<2, K_REALPH>(a, opk[2].m); gate_all<3, K_REALPH>(a, opk[3].m);
            return 3;
        } break;
        case OC_TWHAD4: if constexpr ((SET & C_MACRO_T) != 0) {
            twhad<0>(a, opk[0], opk[0].flags, ctx, tables, n_tab, wcs); twhad<1>(a, opk[1], opk[1].flags, ctx, tables, n_tab, wcs);
            twhad<2>(a, opk[2], opk[2].flags, ctx, tables, n_tab, wcs); twhad<3>(a, opk[3], opk[3].flags, ctx, tables, n_tab, wcs);
            return 3;
        } break;
        default: break;
    }
    return 0;
}
DVD_HD unsigned op_class(int code) {
    if (code < OC_CGEN) {
        const int kind = (code - OC_GATE) / 4;
        return kind == K_GENERAL ? C_GENERAL : (kind == K_REAL || kind == K_REALPH) ? C_REAL : kind == K_RXLIKE ? C_RX
             : kind == K_HADAMARD ? C_HAD : C_RARE;
    }
    if (code < OC_DIAG1 || code == OC_DIAGGEN) return C_RARE;
    if (code < OC_DIAGGEN || (code >= OC_PAIR && code < OC_TWHAD)) return C_DIAG;   // OC_DIAG1, OC_PHASE, OC_PAIR
    if (code < OC_PAIR || (code >= OC_TWHAD && code < OC_REALPH4)) return C_TABLE;
    if (code == OC_REALPH4) return C_MACRO_R;
    if (code == OC_TWHAD4) return C_MACRO_T;
    return 0;   // OC_SWITCH
}
DVD_HD bool is_table_op(int code) { return (code >= OC_TABLE && code < OC_PAIR) || (code >= OC_TWHAD && code < OC_TWHAD + 4) || code == OC_TWHAD4; }

}  // namespace dvd
