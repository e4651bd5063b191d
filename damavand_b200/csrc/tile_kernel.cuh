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
// registers while the gates run.  The two-group persistent form runs 2 x NTHREADS threads per CTA: the thread's
// index inside its group is %tid.x & (NTHREADS - 1) in every form.
struct IoAddr { uint64_t i0; uint64_t hs[REG_BITS]; };
__device__ __forceinline__ int group_tid() {
    unsigned tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    return (int)(tid & (NTHREADS - 1));
}
// Per-thread tile offsets of the three register-group layouts, stashed in shared memory at the start of the kernel:
// they are needed again after every transpose (thread-level controls and parities) and for the write-back, when all
// registers hold amplitudes, and fetching them from the planner's table at that point put a dependent global load
// in front of the stores (7 % of the stall samples of a dense 30-qubit pass, profiles/r2_ncu_qft30_classic2.txt).
// 32 bits per entry: tile positions 0..2 are always qubits 0..2 (PlanOptions::min_low >= 3), so the low three bits of
// an offset are the low three bits of the thread index (layouts 1, 2) or zero (layout 0).  Every thread only reads
// the slots it wrote itself: no barrier.
constexpr int TOFF_WORDS = NGROUPS * NTHREADS;
struct Toff {
    const uint32_t* s;
    int tid;
    __device__ __forceinline__ uint64_t get(int g) const {
        return ((uint64_t)s[g * NTHREADS + tid] << 3) | (g == 0 ? 0u : (unsigned)(tid & 7));
    }
};
__device__ __forceinline__ Toff toff_fill(uint32_t* s, const PassDesc& pd, int tid) {
    uint64_t v[NGROUPS];
#pragma unroll
    for (int g = 0; g < NGROUPS; ++g) v[g] = tid_offset(pd, g, tid);
#pragma unroll
    for (int g = 0; g < NGROUPS; ++g) s[g * NTHREADS + tid] = (uint32_t)(v[g] >> 3);
    return Toff{s, tid};
}
// Every amplitude is touched exactly once per pass: stream it past L1 so that the phase tables stay there.
__device__ __forceinline__ cplx ld_stream(const cplx* p) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
    return cplx{v.x, v.y};
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }
// `cbase` = cta_base_runs(tile index): the tile's physical base, computed once per tile by the caller;
// `toff` = this thread's offset in the group-G layout.
template <int G>
__device__ __forceinline__ IoAddr io_addr(const PassDesc& pd, uint64_t cbase, uint64_t toff) {
    IoAddr io;
    io.i0 = cbase + toff;
#pragma unroll
    for (int k = 0; k < REG_BITS; ++k) io.hs[k] = 1ull << pd.tile_q[G * REG_BITS + k];
    return io;
}
__device__ __forceinline__ uint64_t io_reg_offset(const IoAddr& io, int j) {
    uint64_t off = 0;
#pragma unroll
    for (int k = 0; k < REG_BITS; ++k) if ((j >> k) & 1) off += io.hs[k];
    return off;
}
// ST: the store-side remap is a compile-time variant of the kernels (a run-time branch here costs the plain passes
// registers: 120 -> 128 and a spill on the 30-qubit Fourier passes); the engine launches it only with pd.remap_st set.
template <int G, bool ST = false>
__device__ __forceinline__ void tile_store(cplx* amp, const PassDesc& pd, const cplx (&a)[NREG], uint64_t cbase, const uint32_t* s_toff) {
    const Toff t{s_toff, group_tid()};       // (opaque re-read of %tid: nothing address-related stays live across the gates)
    const IoAddr io = io_addr<G>(pd, cbase, t.get(G));
    if constexpr (!ST) {
        cplx* p0 = amp + io.i0;
#pragma unroll
        for (int j = 0; j < NREG; ++j) st_stream(p0 + io_reg_offset(io, j), a[j]);
    } else {
        // store-side remap (PassDesc::remap_st): every amplitude goes to the rank and index that hold it after the
        // global<->local swaps -- this rank's second chunk or a partner's, written over NVLink
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            const uint64_t i = io.i0 + io_reg_offset(io, j);
            st_stream(const_cast<cplx*>(pd.remap_src[remap_sel(pd, i)]) + remap_index(pd, i), a[j]);
        }
    }
}
// Load the tile in the group-G layout.  Support tracking (PassDesc::zero_mask): amplitudes with a bit of zero_mask set
// are zero by construction and their memory is never read (after a reset it has not even been written).
// With a fused remap (PassDesc::remap_on) every amplitude comes from the buffer -- this rank's input or a partner
// rank's, over NVLink -- that held it before the global<->local qubit swap(s); zero_mask then applies to the source index.
// `toff`: the thread's offset in the group-G layout (tid_offset_arith at the start of a kernel, else the stash).
template <int G>
__device__ __forceinline__ void tile_load(const cplx* amp, const PassDesc& pd, cplx (&a)[NREG], uint64_t cbase, uint64_t toff) {
    const IoAddr io = io_addr<G>(pd, cbase, toff);
    const uint64_t zmask = pd.zero_mask;
    if (!pd.remap_on) {
        const cplx* p0 = amp + io.i0;
        const bool thread_zero = (toff & zmask) != 0;
        const int zregs = pd.zero_regbits;
#pragma unroll
        for (int j = 0; j < NREG; ++j)
            a[j] = (thread_zero || (j & zregs)) ? cplx{0.0, 0.0} : ld_stream(p0 + io_reg_offset(io, j));
    } else {
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            const uint64_t i = io.i0 + io_reg_offset(io, j);
            const uint64_t src = remap_index(pd, i);
            a[j] = (src & zmask) ? cplx{0.0, 0.0} : ld_stream(pd.remap_src[remap_sel(pd, i)] + src);
        }
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
// stage_load<G> reads them back from (dense states only: no zero_mask handling).  Fused remaps are honoured.
template <int G>
__device__ __forceinline__ void tile_prefetch_issue(cplx* tile, const cplx* amp, const PassDesc& pd, uint64_t cbase, int tid, uint64_t toff) {
    const IoAddr io = io_addr<G>(pd, cbase, toff);
    cplx* sp = tile + smem_slot(stage_idx(G, tid, 0));
    if (!pd.remap_on) {
        const cplx* p0 = amp + io.i0;
#pragma unroll
        for (int j = 0; j < NREG; ++j) cp_async16(sp + smem_slot(j << (REG_BITS * G)), p0 + io_reg_offset(io, j));
    } else {
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            const uint64_t i = io.i0 + io_reg_offset(io, j);
            cp_async16(sp + smem_slot(j << (REG_BITS * G)), pd.remap_src[remap_sel(pd, i)] + remap_index(pd, i));
        }
    }
}
template <int FROM>
__device__ __forceinline__ void switch_store(cplx* tile, const cplx (&a)[NREG], int tid, const DevOp& op, unsigned flags, uint64_t gbase) {
    if (flags & F_PERM) stage_store_perm<FROM>(tile, a, tid, op, gbase);
    else stage_store<FROM>(tile, a, tid);
}

// =================================================================================================
// Two-group persistent form ("ring"): ONE CTA of 2 x NTHREADS threads per SM, three shared-memory tile buffers.
// The CTA's tiles are numbered as slots s = 0, 1, 2, ...; slot s is computed by thread group s % 2 in buffer s % 3.
// A buffer only serves the transposes of its slot: as soon as the slot's last transpose has been read back the
// buffer is free, and the same threads fetch slot s + 3 into it with cp.async -- a full tile period before that
// slot is needed (by the OTHER group).  Completion travels through one mbarrier per buffer
// (cp.async.mbarrier.arrive.noinc), so no thread ever waits on a global load it issued itself, and the two
// groups run half a period apart: HBM reads of slot s + 3, fp64 work of slot s and shared-memory transposes of
// slot s + 1 overlap inside one SM.  The groups synchronise internally with named barriers (bar.sync id, 256).
// =================================================================================================

__device__ __forceinline__ void group_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(NTHREADS) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
// the executing thread's earlier cp.async copies arrive on `bar` when they have landed (counted in the init count)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
struct Ring {
    cplx* tiles;        // RING_BUFFERS x TILE_SLOTS
    cplx* wc;           // [RING_GROUPS][2][MAX_TABLE_OPS]
    uint64_t* full;     // RING_BUFFERS mbarriers
    uint32_t* toff;     // TOFF_WORDS: thread-offset stash (the same for both groups)
    __device__ __forceinline__ cplx* tile(unsigned slot) const { return tiles + (slot % RING_BUFFERS) * TILE_SLOTS; }
    __device__ __forceinline__ cplx* wcs(int g, unsigned k) const { return wc + (g * 2 + (k & 1)) * MAX_TABLE_OPS; }
};
__device__ __forceinline__ Ring ring_setup(unsigned char* smem_raw) {
    Ring r;
    r.tiles = reinterpret_cast<cplx*>(smem_raw);
    r.wc = r.tiles + RING_BUFFERS * TILE_SLOTS;
    r.full = reinterpret_cast<uint64_t*>(r.wc + RING_GROUPS * 2 * MAX_TABLE_OPS);
    r.toff = reinterpret_cast<uint32_t*>(r.full + 8);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < RING_BUFFERS; ++b) mbar_init(r.full + b, NTHREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    return r;
}
// Fetch slot `slot` (tile index t) into its buffer; all NTHREADS threads of one group call this.
__device__ __forceinline__ void ring_fetch(const Ring& r, unsigned slot, const cplx* amp, const PassDesc& pd, uint64_t t, int tid, uint64_t toff) {
    tile_prefetch_issue<IO_GROUP>(r.tile(slot), amp, pd, cta_base_runs(pd, t), tid, toff);
    cp_async_arrive(r.full + slot % RING_BUFFERS);
}
// Wait until slot `slot` has landed in its buffer (its use number slot / RING_BUFFERS gives the phase parity).
__device__ __forceinline__ void ring_wait(const Ring& r, unsigned slot) {
    mbar_wait(r.full + slot % RING_BUFFERS, (slot / RING_BUFFERS) & 1u);
}

}  // namespace dvd
