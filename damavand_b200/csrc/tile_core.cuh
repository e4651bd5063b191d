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
static_assert(REG_BITS == 4, "apply_op decodes register groups with >> 2");

enum OpKind : int32_t {
    K_GENERAL = 0,   // arbitrary complex 2x2
    K_REAL = 1,      // all four entries real (H, RY)
    K_RXLIKE = 2,    // real diagonal, purely imaginary off-diagonal (RX)
    K_DIAG = 3,      // m01 = m10 = 0 (RZ, Z, S, T)
    K_ANTIDIAG = 4,  // m00 = m11 = 0 (Y)
    K_SWAP = 5,      // m01 = m10 = 1, m00 = m11 = 0 (X, CNOT): pure exchange
};

// One gate as the device sees it.  80 bytes, 16-byte aligned.
struct alignas(16) DevOp {
    double m[8];      // m00.re m00.im m01.re m01.im m10.re m10.im m11.re m11.im
    int32_t kind;
    int8_t group;     // register group of the target (0..NGROUPS-1), -1 = diagonal, any stage
    int8_t tpos;      // target tile position, -1 = not in the tile (diagonal gates only)
    int8_t tbit;      // target physical bit (always valid)
    int8_t cpos;      // control tile position, -1 = not in the tile
    int8_t cbit;      // control physical bit, -1 = no control
    int8_t d0_is_one; // diagonal gate with m00 == 1 exactly
    int8_t pad[2];
    int32_t gate_idx; // index of the gate in the caller's list (debug / plan inspection)
};
static_assert(sizeof(DevOp) == 80, "DevOp layout");

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

// ---- arithmetic --------------------------------------------------------------------------------
DVD_HD cplx cmul(cplx a, double mr, double mi) { return cplx{a.x * mr - a.y * mi, a.x * mi + a.y * mr}; }

template <int B>
DVD_HD void apply_pairs(cplx (&a)[NREG], const DevOp& op, int creg, bool active) {
    if (!active) return;
#pragma unroll
    for (int k = 0; k < NREG / 2; ++k) {
        const int j0 = ((k >> B) << (B + 1)) | (k & ((1 << B) - 1));
        const int j1 = j0 | (1 << B);
        if (creg >= 0 && !((j0 >> creg) & 1)) continue;
        const cplx x = a[j0], y = a[j1];
        switch (op.kind) {
            case K_GENERAL: {
                cplx nx, ny;
                nx.x = x.x * op.m[0] - x.y * op.m[1] + y.x * op.m[2] - y.y * op.m[3];
                nx.y = x.x * op.m[1] + x.y * op.m[0] + y.x * op.m[3] + y.y * op.m[2];
                ny.x = x.x * op.m[4] - x.y * op.m[5] + y.x * op.m[6] - y.y * op.m[7];
                ny.y = x.x * op.m[5] + x.y * op.m[4] + y.x * op.m[7] + y.y * op.m[6];
                a[j0] = nx; a[j1] = ny;
            } break;
            case K_REAL: {
                a[j0] = cplx{x.x * op.m[0] + y.x * op.m[2], x.y * op.m[0] + y.y * op.m[2]};
                a[j1] = cplx{x.x * op.m[4] + y.x * op.m[6], x.y * op.m[4] + y.y * op.m[6]};
            } break;
            case K_RXLIKE: {  // m00, m11 real; m01 = i*m[3], m10 = i*m[5]
                a[j0] = cplx{x.x * op.m[0] - y.y * op.m[3], x.y * op.m[0] + y.x * op.m[3]};
                a[j1] = cplx{y.x * op.m[6] - x.y * op.m[5], y.y * op.m[6] + x.x * op.m[5]};
            } break;
            case K_ANTIDIAG: {
                a[j0] = cmul(y, op.m[2], op.m[3]);
                a[j1] = cmul(x, op.m[4], op.m[5]);
            } break;
            case K_SWAP: {
                a[j0] = y; a[j1] = x;
            } break;
            default: break;
        }
    }
}

// Apply one op to the 16 register-resident amplitudes of a thread.
//   g      current register group (stage)
//   tbase  tile index of this thread's register 0 in stage g (register bits are zero)
//   gbase  physical index of the CTA's tile origin, OR-ed with the rank's global bits
DVD_HD void apply_op(cplx (&a)[NREG], const DevOp& op, int g, int tbase, uint64_t gbase) {
    bool active = true;
    int creg = -1;
    if (op.cbit >= 0) {
        if (op.cpos < 0) active = (gbase >> op.cbit) & 1ull;
        else if ((op.cpos >> 2) == g) creg = op.cpos & (REG_BITS - 1);
        else active = (tbase >> op.cpos) & 1;
    }
    if (op.kind == K_DIAG) {
        int treg = -1, tsel = 0;
        if (op.tpos < 0) tsel = (int)((gbase >> op.tbit) & 1ull);
        else if ((op.tpos >> 2) == g) treg = op.tpos & (REG_BITS - 1);
        else tsel = (tbase >> op.tpos) & 1;
        if (!active) return;
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            if (creg >= 0 && !((j >> creg) & 1)) continue;
            const int bit = treg >= 0 ? ((j >> treg) & 1) : tsel;
            if (bit) a[j] = cmul(a[j], op.m[6], op.m[7]);
            else if (!op.d0_is_one) a[j] = cmul(a[j], op.m[0], op.m[1]);
        }
        return;
    }
    // non-diagonal: the planner guarantees the target is in the current register group
    switch (op.tpos & (REG_BITS - 1)) {
        case 0: apply_pairs<0>(a, op, creg, active); break;
        case 1: apply_pairs<1>(a, op, creg, active); break;
        case 2: apply_pairs<2>(a, op, creg, active); break;
        default: apply_pairs<3>(a, op, creg, active); break;
    }
}

}  // namespace dvd
