// tile_kernel.cuh -- device-side helpers of the tiled pass kernel (register <-> shared-memory transposes, global
// tile I/O, cp.async staging).  Shared by the interpreter kernels in kernels.cu and by the structure-specialised
// kernels generated at run time (jit.cpp): one source of truth for the data movement.
#pragma once
#include "tile_core.cuh"

namespace dvd {

// =================================================================================================
// Tiled multi-gate pass
// =================================================================================================

// Register <-> shared-memory transposes.  Thanks to the additive padding (smem_slot) every
// register's slot is a compile-time offset from one per-thread base.
template <int G>
__device__ __forceinline__ void stage_store(cplx* tile, const cplx (&a)[NREG], int tid) {
    cplx* p = tile + smem_slot(stage_idx(G, tid, 0));
#pragma unroll
    for (int j = 0; j < NREG; ++j) p[smem_slot(j << (REG_BITS * G))] = a[j];
}
template <int G>
__device__ __forceinline__ void stage_load(const cplx* tile, cplx (&a)[NREG], int tid) {
    const cplx* p = tile + smem_slot(stage_idx(G, tid, 0));
#pragma unroll
    for (int j = 0; j < NREG; ++j) a[j] = p[smem_slot(j << (REG_BITS * G))];
}
// Store through a GF(2)-affine permutation of the tile index: every pending X / CNOT of the pass is
// executed here, as addressing, instead of as data movement of its own.
template <int G>
__device__ __forceinline__ void stage_store_perm(cplx* tile, const cplx (&a)[NREG], int tid, const DevOp& op, uint64_t gbase) {
    const PermPayload& pp = *reinterpret_cast<const PermPayload*>(op.m);
    const unsigned pb = perm_index(op, perm_const(op, gbase), (unsigned)stage_idx(G, tid, 0));
    const unsigned c0 = pp.col[REG_BITS * G + 0], c1 = pp.col[REG_BITS * G + 1];
    const unsigned c2 = pp.col[REG_BITS * G + 2], c3 = pp.col[REG_BITS * G + 3];
#pragma unroll
    for (int j = 0; j < NREG; ++j) {
        const unsigned x = pb ^ ((j & 1) ? c0 : 0u) ^ ((j & 2) ? c1 : 0u) ^ ((j & 4) ? c2 : 0u) ^ ((j & 8) ? c3 : 0u);
        tile[smem_slot((int)x)] = a[j];
    }
}

// Global <-> register layout = IO_GROUP stage: lanes run over tile positions 0..4, i.e. over
// >= 128 contiguous bytes (tile positions 0..2 are always physical qubits 0..2).  The addressing is
// recomputed for the write-back (opaque re-read of %tid / %ctaid) so that it does not occupy
// registers while the gates run.
struct IoAddr { cplx* p0; uint64_t hs[REG_BITS]; };
// Every amplitude is touched exactly once per pass: stream it past L1 so that the phase tables stay there.
__device__ __forceinline__ cplx ld_stream(const cplx* p) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
    return cplx{v.x, v.y};
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }
// `cbase` = cta_base_runs(tile index): the tile's physical base, computed once per tile by the caller.
template <int G>
__device__ __forceinline__ IoAddr io_addr(cplx* amp, const PassDesc& pd, uint64_t cbase) {
    unsigned tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    IoAddr io;
    io.p0 = amp + cbase + tid_offset(pd, G, (int)tid);
#pragma unroll
    for (int k = 0; k < REG_BITS; ++k) io.hs[k] = 1ull << pd.tile_q[G * REG_BITS + k];
    return io;
}
template <int G>
__device__ __forceinline__ void tile_store(cplx* amp, const PassDesc& pd, const cplx (&a)[NREG], uint64_t cbase) {
    const IoAddr io = io_addr<G>(amp, pd, cbase);
#pragma unroll
    for (int j = 0; j < NREG; ++j) {
        uint64_t off = 0;
#pragma unroll
        for (int k = 0; k < REG_BITS; ++k) if ((j >> k) & 1) off += io.hs[k];
        st_stream(io.p0 + off, a[j]);
    }
}
// Load the tile in the group-G layout.  Support tracking (PassDesc::zero_mask): amplitudes with a bit of zero_mask set
// are zero by construction and their memory is never read (after a reset it has not even been written).
template <int G>
__device__ __forceinline__ void tile_load(cplx* amp, const PassDesc& pd, cplx (&a)[NREG], uint64_t cbase, int tid) {
    const IoAddr io = io_addr<G>(amp, pd, cbase);
    const uint64_t zmask = pd.zero_mask;
    const bool thread_zero = (tid_offset(pd, G, tid) & zmask) != 0;
    const int zregs = pd.zero_regbits;
#pragma unroll
    for (int j = 0; j < NREG; ++j) {
        uint64_t off = 0;
#pragma unroll
        for (int k = 0; k < REG_BITS; ++k) if ((j >> k) & 1) off += io.hs[k];
        a[j] = (thread_zero || (j & zregs)) ? cplx{0.0, 0.0} : ld_stream(io.p0 + off);
    }
}

// Asynchronous global -> shared copies (LDGSTS): the next tile travels while the current one is computed on.
__device__ __forceinline__ void cp_async16(cplx* smem_dst, const cplx* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// Every thread fetches exactly the 16 amplitudes it will hold in the group-G layout, into the slots
// stage_load<G> reads them back from: the staging needs no barrier of its own.
template <int G>
__device__ __forceinline__ void tile_prefetch(cplx* tile, cplx* amp, const PassDesc& pd, uint64_t cbase, int tid) {
    const IoAddr io = io_addr<G>(amp, pd, cbase);
    cplx* sp = tile + smem_slot(stage_idx(G, tid, 0));
#pragma unroll
    for (int j = 0; j < NREG; ++j) {
        uint64_t off = 0;
#pragma unroll
        for (int k = 0; k < REG_BITS; ++k) if ((j >> k) & 1) off += io.hs[k];
        cp_async16(sp + smem_slot(j << (REG_BITS * G)), io.p0 + off);
    }
    cp_async_commit();
}

template <int FROM>
__device__ __forceinline__ void switch_store(cplx* tile, const cplx (&a)[NREG], int tid, const DevOp& op, unsigned flags, uint64_t gbase) {
    if (flags & F_PERM) stage_store_perm<FROM>(tile, a, tid, op, gbase);
    else stage_store<FROM>(tile, a, tid);
}

}  // namespace dvd
