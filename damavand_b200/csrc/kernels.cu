// kernels.cu -- hand-written sm_100a kernels of the statevector hot path.
//
// Reference kernels these replace (all in /root/reference/damavand-gpu/kernels.cu):
//   apply_one_qubit_gate_kernel_local        :120-172  -> k_tile_pass (fused) / k_simple_gate
//   apply_one_qubit_gate_kernel_distributed  :174-230  -> k_pack_half / k_unpack_half around an NVLink
//                                                         exchange, then k_tile_pass on local qubits
//   measure_amplitudes_on_device_global      :44-60    -> k_probabilities, k_block_sums (+ tree), k_sample
//   init_zero_state_on_{first,other}_gpu     :62-96    -> cudaMemsetAsync + k_set_basis_state
// The path is HBM-bound complex128 streaming work: no tensor cores by design.
#include "kernels.h"
#include "tile_kernel.cuh"

#include <cstdlib>

namespace dvd {

// One launch = one pass: every amplitude is read once and written once; pp.ops is applied in between.
// The op list lives in the kernel's parameter space (constant bank): op fields are warp-uniform loads.
// SET: op classes compiled in (see OpClass); launch_tile_pass picks the smallest variant covering the pass.
// ST: the store goes through a remap (PassDesc::remap_st, tile_kernel.cuh: tile_store).
template <unsigned SET, bool ST = false>
__global__ void __launch_bounds__(NTHREADS, 2)
k_tile_pass(cplx* __restrict__ amp, const __grid_constant__ PassParams pp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);
    __shared__ cplx s_wc[MAX_TABLE_OPS];      // per-CTA constants of the pass's table ops
    __shared__ uint32_t s_toff[TOFF_WORDS];   // per-thread tile offsets of the three layouts (tile_kernel.cuh: Toff)
    const PassDesc& pd = pp.pd;

    const int tid = threadIdx.x;
    const uint64_t cbase = cta_base_runs(pd, (uint64_t)blockIdx.x);
    const uint64_t gbase = cbase | pd.rank_bits;
    const cplx* __restrict__ tables = pd.tables;
    const int n_tab = pd.n_tab;
    // byte tables only see index bits outside the tile: one constant per CTA and table op.  The two dependent
    // lookups start before the tile's own loads queue up in front of them.
    if (tid < n_tab) s_wc[tid] = table_cta_const(tables, tid, gbase);
    // a tile whose fixed bits hit the support mask is all zero on input, hence on output: nothing to do (the engine
    // normally does not even launch those)
    if (cbase & pd.zero_mask & ~pd.remap_lmask) return;
    cplx a[NREG];
    // with a fused remap (pd.remap_on) the input is pd.remap_src[..], out of place
    tile_load<IO_GROUP>(amp, pd, a, cbase, tid_offset_arith(pd, IO_GROUP, tid));
    const Toff toff = toff_fill(s_toff, pd, tid);   // table lookups in the shadow of the tile's loads

    if (n_tab > 0) __syncthreads();
    ThreadCtx ctx;
    ctx.pidx = gbase | toff.get(IO_GROUP);
    ctx.ph = cplx{1.0, 0.0};
    ctx.ph_dirty = false;
    ctx.tid = tid;
    const int n_ops = pd.n_ops;
    // (fetching the next op's code one iteration ahead was measured: the extra live register spills and
    // costs 2-9 %)
    for (int k = 0; k < n_ops; ++k) {
        const DevOp& op = pp.ops[k];
        const int code = op.code;
        if (code >= OC_SWITCH) {
            const int from = (code - OC_SWITCH) / NGROUPS, to = (code - OC_SWITCH) % NGROUPS;
            flush_phase(a, ctx);
            __syncthreads();   // the previous transpose's loads are done everywhere
            if (from == 0) switch_store<0>(tile, a, tid, op, op.flags, gbase);
            else if (from == 1) switch_store<1>(tile, a, tid, op, op.flags, gbase);
            else switch_store<2>(tile, a, tid, op, op.flags, gbase);
            __syncthreads();
            if (to == 0) stage_load<0>(tile, a, tid);
            else if (to == 1) stage_load<1>(tile, a, tid);
            else stage_load<2>(tile, a, tid);
            ctx.pidx = gbase | toff.get(to);
            continue;
        }
        k += apply_op<SET>(a, &op, code, op.flags, ctx, tables, n_tab, s_wc);
    }
    flush_phase(a, ctx);
    // the planner ends a pass in the group-2 or the group-1 layout: both store 128-byte segments per quarter warp
    // (rank bits lie above the local index bits, so gbase - rank_bits is the tile's base again)
    if (pd.io_out == IO_GROUP) tile_store<IO_GROUP, ST>(amp, pd, a, gbase - pd.rank_bits, s_toff);
    else tile_store<1, ST>(amp, pd, a, gbase - pd.rank_bits, s_toff);
}

// Two-group persistent form of the same pass (tile_kernel.cuh, "ring"): one CTA of 2 x 256 threads per SM, three
// shared-memory tile buffers; slot s is computed by group s % 2 in buffer s % 3, and as soon as its last transpose
// has been read back the same threads fetch slot s + 3 into the buffer with cp.async -- a full tile period ahead of
// its use by the other group.  No thread waits on a global load it issued itself; HBM reads, fp64 work and
// shared-memory transposes of three different tiles overlap inside one SM.  Dense states only (zero_mask == 0).
template <unsigned SET>
__global__ void __launch_bounds__(RING_GROUPS * NTHREADS, 1)
k_tile_pass_ring(cplx* __restrict__ amp, const __grid_constant__ PassParams pp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const PassDesc& pd = pp.pd;
    const cplx* __restrict__ tables = pd.tables;
    const int n_tab = pd.n_tab;
    const cplx* amp_in = amp;
    const Ring ring = ring_setup(smem_raw);
    const int grp = threadIdx.x / NTHREADS, tid = threadIdx.x % NTHREADS;
    const Toff toff = toff_fill(ring.toff, pd, tid);      // once per CTA: the offsets do not depend on the tile
    const unsigned n_tiles = 1u << pd.n_cta_bits, stride = gridDim.x;
    const int n_ops = pd.n_ops;
    const int last_switch = pd.last_switch;
    {   // prologue: slots 0 and 2 are fetched by group 0, slot 1 by group 1
        const unsigned t0 = blockIdx.x + grp * stride;
        if (t0 < n_tiles) ring_fetch(ring, grp, amp_in, pd, t0, tid, toff.get(IO_GROUP));
        if (grp == 0 && blockIdx.x + 2 * stride < n_tiles) ring_fetch(ring, 2, amp_in, pd, blockIdx.x + 2 * stride, tid, toff.get(IO_GROUP));
        if (tid < n_tab && t0 < n_tiles) ring.wcs(grp, 0)[tid] = table_cta_const(tables, tid, cta_base_runs(pd, t0) | pd.rank_bits);
    }
    unsigned kslot = 0;
    for (unsigned slot = grp; ; slot += RING_GROUPS, ++kslot) {
        const unsigned t = blockIdx.x + slot * stride;
        if (t >= n_tiles) break;
        cplx* tile = ring.tile(slot);
        const cplx* wcs = ring.wcs(grp, kslot);
        const uint64_t gbase = cta_base_runs(pd, t) | pd.rank_bits;
        ring_wait(ring, slot);
        group_sync(grp);   // this group's previous slot is over everywhere; its table constants are visible
        cplx a[NREG];
        stage_load<IO_GROUP>(tile, a, tid);
        // the slot's buffer is free once its last transpose has been read back: fetch slot + 3 into it
        auto release_buffer = [&]() {
            group_sync(grp);
            if (t + 3 * stride < n_tiles) ring_fetch(ring, slot + 3, amp_in, pd, t + 3 * stride, tid, toff.get(IO_GROUP));
            if (tid < n_tab && t + 2 * stride < n_tiles)
                ring.wcs(grp, kslot + 1)[tid] = table_cta_const(tables, tid, cta_base_runs(pd, t + 2 * stride) | pd.rank_bits);
        };
        if (last_switch < 0) release_buffer();
        ThreadCtx ctx;
        ctx.pidx = gbase | toff.get(IO_GROUP);
        ctx.ph = cplx{1.0, 0.0};
        ctx.ph_dirty = false;
        ctx.tid = tid;
        for (int k = 0; k < n_ops; ++k) {
            const DevOp& op = pp.ops[k];
            const int code = op.code;
            if (code >= OC_SWITCH) {
                const int from = (code - OC_SWITCH) / NGROUPS, to = (code - OC_SWITCH) % NGROUPS;
                flush_phase(a, ctx);
                group_sync(grp);
                if (from == 0) switch_store<0>(tile, a, tid, op, op.flags, gbase);
                else if (from == 1) switch_store<1>(tile, a, tid, op, op.flags, gbase);
                else switch_store<2>(tile, a, tid, op, op.flags, gbase);
                group_sync(grp);
                if (to == 0) stage_load<0>(tile, a, tid);
                else if (to == 1) stage_load<1>(tile, a, tid);
                else stage_load<2>(tile, a, tid);
                ctx.pidx = gbase | toff.get(to);
                if (k == last_switch) release_buffer();
                continue;
            }
            k += apply_op<SET>(a, &op, code, op.flags, ctx, tables, n_tab, wcs);
        }
        flush_phase(a, ctx);
        if (pd.io_out == IO_GROUP) tile_store<IO_GROUP>(amp, pd, a, gbase - pd.rank_bits, ring.toff);
        else tile_store<1>(amp, pd, a, gbase - pd.rank_bits, ring.toff);
    }
}

// =================================================================================================
// One gate per pass (small states, debug path)
// =================================================================================================
__global__ void __launch_bounds__(256)
k_simple_gate(cplx* __restrict__ amp, int n_local, uint64_t rank_bits, const __grid_constant__ SimpleOp op) {
    const uint64_t n = 1ull << n_local;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = op.tbit, c = op.cbit;
    if (t < n_local) {
        const uint64_t pairs = n >> 1;
        for (uint64_t k = gtid; k < pairs; k += stride) {
            const uint64_t i0 = ((k >> t) << (t + 1)) | (k & ((1ull << t) - 1));
            const uint64_t i1 = i0 | (1ull << t);
            if (c >= 0 && !(((i0 | rank_bits) >> c) & 1ull)) continue;
            const cplx x = amp[i0], y = amp[i1];
            cplx nx, ny;
            nx.x = x.x * op.m[0] - x.y * op.m[1] + y.x * op.m[2] - y.y * op.m[3];
            nx.y = x.x * op.m[1] + x.y * op.m[0] + y.x * op.m[3] + y.y * op.m[2];
            ny.x = x.x * op.m[4] - x.y * op.m[5] + y.x * op.m[6] - y.y * op.m[7];
            ny.y = x.x * op.m[5] + x.y * op.m[4] + y.x * op.m[7] + y.y * op.m[6];
            amp[i0] = nx; amp[i1] = ny;
        }
    } else {
        // rank-index target: only diagonal gates get here; the whole chunk sees one matrix entry
        const int bit = (int)((rank_bits >> t) & 1ull);
        const double dr = bit ? op.m[6] : op.m[0], di = bit ? op.m[7] : op.m[1];
        for (uint64_t i = gtid; i < n; i += stride) {
            if (c >= 0 && !(((i | rank_bits) >> c) & 1ull)) continue;
            amp[i] = cmul(amp[i], dr, di);
        }
    }
}

__global__ void k_set_basis_state(cplx* amp, uint64_t index, double value) { amp[index] = cplx{value, 0.0}; }

// Materialise the zeros that support tracking implied: every amplitude with a bit of zmask set becomes 0.0.
__global__ void __launch_bounds__(256)
k_zero_outside_support(cplx* __restrict__ amp, uint64_t n, uint64_t zmask) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (i & zmask) st_stream(amp + i, cplx{0.0, 0.0});
}

// =================================================================================================
// Probabilities, pairwise summation tree, sampler
// =================================================================================================
// |a|^2 exactly as the reference computes it (norm_sqr = re*re + im*im, no contraction):
// src/qubit_backend/circuit.rs:579-581
__device__ __forceinline__ double norm_sqr(cplx a) { return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y)); }

__global__ void __launch_bounds__(256)
k_probabilities(const cplx* __restrict__ amp, uint64_t first, uint64_t count, double* __restrict__ probs) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        probs[i] = norm_sqr(amp[first + i]);
}

// Level-BLK_BITS sums: one value per block of 1024 amplitudes, summed as a binary pairwise tree
// (xor-butterfly == pairwise tree because fp addition is commutative).
__global__ void __launch_bounds__(256)
k_block_sums(const cplx* __restrict__ amp, uint64_t n_blocks, double* __restrict__ out) {
    __shared__ double wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const cplx* src = amp + (b << BLK_BITS);
        double v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = norm_sqr(src[k * 256 + tid]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) v[k] = __dadd_rn(v[k], __shfl_xor_sync(0xffffffffu, v[k], off));
            if (lane == 0) wsum[k * 8 + warp] = v[k];
        }
        __syncthreads();
        if (warp == 0) {
            double w = wsum[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) w = __dadd_rn(w, __shfl_xor_sync(0xffffffffu, w, off));
            if (lane == 0) out[b] = w;
        }
        __syncthreads();
    }
}

// Whole tree of a state smaller than one block (n_local < BLK_BITS): root only.
__global__ void __launch_bounds__(256)
k_small_root(const cplx* __restrict__ amp, int n_local, double* __restrict__ out) {
    __shared__ double buf[2][1 << (BLK_BITS - 1)];
    const int n = 1 << n_local;
    if (n == 1) { if (threadIdx.x == 0) out[0] = norm_sqr(amp[0]); return; }
    for (int j = threadIdx.x; j < n / 2; j += blockDim.x)
        buf[0][j] = __dadd_rn(norm_sqr(amp[2 * j]), norm_sqr(amp[2 * j + 1]));
    __syncthreads();
    int cur = 0;
    for (int cnt = n / 4; cnt >= 1; cnt >>= 1) {
        for (int j = threadIdx.x; j < cnt; j += blockDim.x)
            buf[cur ^ 1][j] = __dadd_rn(buf[cur][2 * j], buf[cur][2 * j + 1]);
        __syncthreads();
        cur ^= 1;
    }
    if (threadIdx.x == 0) out[0] = buf[cur][0];
}

__global__ void __launch_bounds__(256)
k_tree_level(const double* __restrict__ in, double* __restrict__ out, uint64_t cnt) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += stride)
        out[j] = __dadd_rn(in[2 * j], in[2 * j + 1]);
}

uint64_t tree_level_offset(int n_local, int level) {
    const int nb = n_local < BLK_BITS ? n_local : BLK_BITS;
    uint64_t off = 0;
    for (int l = nb; l < level; ++l) off += 1ull << (n_local - l);
    return off;
}
uint64_t tree_size(int n_local) { return tree_level_offset(n_local, n_local) + 1; }

struct TreeOffsets { uint64_t off[64]; };

// One warp per shot.  Upper levels are read from the stored tree, the last nb levels are rebuilt
// from the amplitudes of the one block the shot lands in (same pairwise order as k_block_sums).
constexpr int SAMPLE_WARPS = 4;
__global__ void __launch_bounds__(SAMPLE_WARPS * 32)
k_sample(const cplx* __restrict__ amp, int n_local, int nb, const double* __restrict__ tree,
         const __grid_constant__ TreeOffsets offs, const double* __restrict__ u,
         const int32_t* __restrict__ sel, int32_t sel_value, uint64_t index_offset, uint64_t shots,
         unsigned long long* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* lv = reinterpret_cast<double*>(smem_raw) + (size_t)warp * (2ull << nb);
    const double total = tree[offs.off[n_local]];
    const uint64_t n_warps = (uint64_t)gridDim.x * SAMPLE_WARPS;
    const int bn = 1 << nb;
    for (uint64_t s = (uint64_t)blockIdx.x * SAMPLE_WARPS + warp; s < shots; s += n_warps) {
        if (sel != nullptr && sel[s] != sel_value) continue;
        const double xsi = __dmul_rn(u[s], total);
        double basev = 0.0;
        uint64_t j = 0;
        for (int l = n_local; l > nb; --l) {
            const double c = __dadd_rn(basev, tree[offs.off[l - 1] + 2 * j]);
            if (xsi <= c) j = 2 * j; else { basev = c; j = 2 * j + 1; }
        }
        const cplx* src = amp + (j << nb);
        for (int i = lane; i < bn; i += 32) lv[i] = norm_sqr(src[i]);
        __syncwarp();
        // level l of the block lives at lv + (2^(nb+1) - 2^(nb-l+1)), i.e. levels are packed back to back
        int in_off = 0;
        for (int l = 1; l <= nb; ++l) {
            const int cnt = bn >> l;
            const int out_off = in_off + (bn >> (l - 1));
            for (int i = lane; i < cnt; i += 32) lv[out_off + i] = __dadd_rn(lv[in_off + 2 * i], lv[in_off + 2 * i + 1]);
            __syncwarp();
            in_off = out_off;
        }
        // in_off now points at level nb (one entry).  Walk back down.
        uint32_t k = 0;
        int lvl_off = in_off;
        for (int l = nb; l >= 1; --l) {
            lvl_off -= bn >> (l - 1);  // offset of level l-1
            const double c = __dadd_rn(basev, lv[lvl_off + 2 * k]);
            if (xsi <= c) k = 2 * k; else { basev = c; k = 2 * k + 1; }
        }
        if (lane == 0) out[s] = index_offset + (j << nb) + k;
        __syncwarp();
    }
}

// ---- reference-order sampler (optional mode) ---------------------------------------------------
// The reference's cumulative array is a strict left-to-right fp64 sum (utils.rs:270-274); a parallel reduction
// rounds differently, and no parallel algorithm reproduces sequential rounding.  For states small enough to afford
// it this mode pays for the sequential order: ONE warp walks the whole chunk once per state (lanes fetch and square
// 32 amplitudes at a time, lane 0 adds them in index order) and stores the running sum at every block boundary
// (cum[(b+1) * 2^nb]); a shot then needs a binary search over the boundaries and a sequential scan of one block.
// The sums are exactly the reference's cum[] entries, so the sampled indices are the reference's for every draw.
__global__ void __launch_bounds__(32)
k_seq_block_cum(const cplx* __restrict__ amp, int nb, uint64_t n_blocks, double* __restrict__ block_cum) {
    __shared__ double p[2][32];
    const int lane = threadIdx.x;
    const uint64_t n = n_blocks << nb;
    double s = 0.0;
    const uint64_t per_block = 1ull << nb;
    // double-buffered: lanes fetch chunk c + 1 while lane 0 adds chunk c
    if (lane < (int)(n < 32 ? n : 32)) p[0][lane] = norm_sqr(amp[lane]);
    __syncwarp();
    const uint64_t n_chunks = (n + 31) / 32;
    for (uint64_t c = 0; c < n_chunks; ++c) {
        const int cur = (int)(c & 1);
        const uint64_t nxt = (c + 1) * 32 + lane;
        double v = 0.0;
        if (c + 1 < n_chunks && nxt < n) v = norm_sqr(amp[nxt]);
        if (lane == 0) {
            const uint64_t base = c * 32;
            const int cnt = (int)(n - base < 32 ? n - base : 32);
#pragma unroll 8
            for (int i = 0; i < cnt; ++i) {
                s = __dadd_rn(s, p[cur][i]);
                if (((base + i + 1) & (per_block - 1)) == 0) block_cum[(base + i) >> nb] = s;
            }
        }
        p[cur ^ 1][lane] = v;
        __syncwarp();
    }
}

// One thread per shot: out[s] = first j with xsi <= cum[j + 1] (utils.rs:258-268; xsi = 0 -> 0).
__global__ void __launch_bounds__(128)
k_sample_seq(const cplx* __restrict__ amp, int nb, uint64_t n_blocks, const double* __restrict__ block_cum,
             const double* __restrict__ u, const int32_t* __restrict__ sel, int32_t sel_value, uint64_t index_offset,
             uint64_t shots, unsigned long long* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const double total = block_cum[n_blocks - 1];
    for (uint64_t sh = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; sh < shots; sh += stride) {
        if (sel != nullptr && sel[sh] != sel_value) continue;
        const double xsi = __dmul_rn(u[sh], total);
        // first block b with xsi <= cum at its end (the boundary sums are non-decreasing)
        uint64_t lo = 0, hi = n_blocks - 1;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (xsi <= block_cum[mid]) hi = mid; else lo = mid + 1;
        }
        double sacc = lo > 0 ? block_cum[lo - 1] : 0.0;
        const cplx* src = amp + (lo << nb);
        const uint64_t cnt = 1ull << nb;
        uint64_t k = cnt - 1;
        for (uint64_t i = 0; i < cnt; ++i) {
            sacc = __dadd_rn(sacc, norm_sqr(src[i]));
            if (xsi <= sacc) { k = i; break; }
        }
        out[sh] = index_offset + (lo << nb) + k;
    }
}

__global__ void __launch_bounds__(256)
k_extract_expectation(const unsigned long long* __restrict__ samples, uint64_t shots, const int* __restrict__ qubits,
                      int n_obs, double* __restrict__ out) {
    const uint64_t total = shots * (uint64_t)n_obs;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const uint64_t s = e / n_obs;
        const int o = (int)(e - s * n_obs);
        out[e] = ((samples[s] >> qubits[o]) & 1ull) ? -1.0 : 1.0;   // circuit.rs:503-508
    }
}

// =================================================================================================
// Exact <Z_q> (extension): warp-shuffle + block reduction, deterministic second stage
// =================================================================================================
constexpr int EZ_CTAS = 148 * 4;
constexpr int EZ_Q = 64;
constexpr int EZ_LOW = 8;          // bits [0, EZ_LOW) of an index are fixed per thread (grid stride is a multiple of 256)
constexpr int EZ_HIGH = 32;        // accumulators for bits [EZ_LOW, EZ_LOW + EZ_HIGH): local qubits up to 40
uint64_t ez_partial_size() { return (uint64_t)EZ_CTAS * EZ_Q; }

// partial[cta][q] = sum over the CTA's amplitudes of p_i * (1 - 2 bit_q(i)) for q < 40, partial[cta][63] = sum p_i.
// The grid-stride loop advances by a multiple of 256, so bits 0..7 of every index a thread visits equal its thread
// index: those eight sums are +-(the thread's total) and need no accumulator of their own.
__global__ void __launch_bounds__(256)
k_expect_z_partial(const cplx* __restrict__ amp, int n_local, double* __restrict__ partial) {
    __shared__ double red[8][EZ_LOW + EZ_HIGH + 1];
    double acc[EZ_HIGH];
    double tot = 0.0;
#pragma unroll
    for (int q = 0; q < EZ_HIGH; ++q) acc[q] = 0.0;
    const uint64_t n = 1ull << n_local;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double p = norm_sqr(amp[i]);
        tot += p;
        const uint64_t hi = i >> EZ_LOW;
#pragma unroll
        for (int q = 0; q < EZ_HIGH; ++q) acc[q] += ((hi >> q) & 1ull) ? -p : p;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < EZ_LOW + EZ_HIGH + 1; ++q) {
        double v = q < EZ_LOW ? (((threadIdx.x >> q) & 1) ? -tot : tot) : q < EZ_LOW + EZ_HIGH ? acc[q < EZ_LOW ? 0 : q - EZ_LOW] : tot;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < EZ_LOW + EZ_HIGH + 1) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        partial[(uint64_t)blockIdx.x * EZ_Q + (threadIdx.x == EZ_LOW + EZ_HIGH ? EZ_Q - 1 : threadIdx.x)] = v;
    }
}

__global__ void k_expect_z_final(const double* __restrict__ partial, int n_ctas, int n_local, int n_total,
                                 uint64_t rank_bits, double* __restrict__ out) {
    const int q = threadIdx.x;
    if (q >= n_total) return;
    const int src = q < n_local ? q : EZ_Q - 1;     // rank-index qubits: +-(the chunk's total probability)
    double v = 0.0;
    for (int c = 0; c < n_ctas; ++c) v += partial[(uint64_t)c * EZ_Q + src];
    if (q >= n_local && ((rank_bits >> q) & 1ull)) v = -v;
    out[q] = v;
}

// =================================================================================================
// Swap staging and dot product
// =================================================================================================
__global__ void __launch_bounds__(256)
k_pack_half(const cplx* __restrict__ amp, int lq, int bitval, uint64_t first, uint64_t count, cplx* __restrict__ buf) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride)
        buf[e] = amp[half_index(first + e, lq, bitval)];
}
__global__ void __launch_bounds__(256)
k_unpack_half(cplx* __restrict__ amp, int lq, int bitval, uint64_t first, uint64_t count, const cplx* __restrict__ buf) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride)
        amp[half_index(first + e, lq, bitval)] = buf[e];
}

// Global<->local qubit swap straight over NVLink peer memory: ONE kernel reads this rank's leaving
// half and the partner's leaving half (a peer pointer mapped with CUDA IPC) and writes each into the
// other's place.  The same thread moves both elements of a pair, so nothing is staged and there is no
// race; the two ranks of a pair each handle half of the elements, which loads both NVLink directions
// equally (reads pull data towards this GPU, writes push it away).  Replaces
// exchange_amplitudes_between_gpus (rust_communication.cu:106-141: four serial full-chunk copies).
constexpr int SWAP_UNROLL = 4;
__global__ void __launch_bounds__(256)
k_swap_peer(cplx* __restrict__ mine, cplx* __restrict__ peer, int lq, int my_bit, uint64_t e_begin, uint64_t e_end) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t e = e_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // my leaving half: local bit lq == 1 - my_bit; the partner's leaving half: its bit lq == my_bit
    for (; e + (SWAP_UNROLL - 1) * stride < e_end; e += SWAP_UNROLL * stride) {
        double2 x[SWAP_UNROLL], y[SWAP_UNROLL];
#pragma unroll
        for (int u = 0; u < SWAP_UNROLL; ++u) {
            x[u] = *reinterpret_cast<const double2*>(mine + half_index(e + u * stride, lq, 1 - my_bit));
            y[u] = *reinterpret_cast<const double2*>(peer + half_index(e + u * stride, lq, my_bit));
        }
#pragma unroll
        for (int u = 0; u < SWAP_UNROLL; ++u) {
            *reinterpret_cast<double2*>(mine + half_index(e + u * stride, lq, 1 - my_bit)) = y[u];
            *reinterpret_cast<double2*>(peer + half_index(e + u * stride, lq, my_bit)) = x[u];
        }
    }
    for (; e < e_end; e += stride) {
        const uint64_t im = half_index(e, lq, 1 - my_bit), ip = half_index(e, lq, my_bit);
        const double2 x = *reinterpret_cast<const double2*>(mine + im);
        const double2 y = *reinterpret_cast<const double2*>(peer + ip);
        *reinterpret_cast<double2*>(mine + im) = y;
        *reinterpret_cast<double2*>(peer + ip) = x;
    }
}

// Split / join of the interleaved complex128 state into the separate real and imaginary arrays of the reference's
// boundary (retrieve_amplitudes_on_host / split_amplitudes_between_gpus, rust_communication.cu:400-482): done on the
// device, so that the host only moves bytes.
__global__ void __launch_bounds__(256)
k_split_re_im(const cplx* __restrict__ amp, uint64_t count, double* __restrict__ re, double* __restrict__ im) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const cplx v = amp[i];
        re[i] = v.x; im[i] = v.y;
    }
}
__global__ void __launch_bounds__(256)
k_join_re_im(cplx* __restrict__ amp, uint64_t count, const double* __restrict__ re, const double* __restrict__ im) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) amp[i] = cplx{re[i], im[i]};
}

constexpr int DOT_CTAS = 148 * 4;
__global__ void __launch_bounds__(256)
k_dot_partial(const cplx* __restrict__ a, const cplx* __restrict__ b, uint64_t count, double* __restrict__ partial) {
    __shared__ double red[8][2];
    double re = 0.0, im = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const cplx x = a[i], y = b[i];      // conj(x) * y
        re += x.x * y.x + x.y * y.y;
        im += x.x * y.y - x.y * y.x;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, off);
        im += __shfl_xor_sync(0xffffffffu, im, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[warp][0] = re; red[warp][1] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0, m = 0.0;
        for (int w = 0; w < 8; ++w) { r += red[w][0]; m += red[w][1]; }
        partial[2 * blockIdx.x] = r; partial[2 * blockIdx.x + 1] = m;
    }
}
__global__ void k_dot_final(const double* __restrict__ partial, int n, double* __restrict__ out) {
    double r = 0.0, m = 0.0;
    for (int c = 0; c < n; ++c) { r += partial[2 * c]; m += partial[2 * c + 1]; }
    out[0] = r; out[1] = m;
}

// =================================================================================================
// Launchers
// =================================================================================================
static inline unsigned grid_for(uint64_t work_items, int per_cta, unsigned cap) {
    uint64_t g = (work_items + per_cta - 1) / per_cta;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (unsigned)g;
}
constexpr unsigned STREAM_CAP = 148 * 32;  // grid-stride kernels: a multiple of the SM count

// Kernel variants by op-class set, most specific first; the last one runs everything.
constexpr unsigned V_LAYERED = C_REAL | C_HAD | C_DIAG | C_MACRO_R;                  // RY/RZ/CNOT ansatz layers
constexpr unsigned V_FOURIER = C_HAD | C_DIAG | C_TABLE | C_MACRO_T;                 // Hadamards + controlled phases
constexpr unsigned V_COMMON = C_GENERAL | C_REAL | C_RX | C_HAD | C_DIAG | C_TABLE;  // everything but rare ops / macros
constexpr unsigned VARIANTS[] = {V_LAYERED, V_FOURIER, V_COMMON, C_ALL};
typedef void (*TileKernel)(cplx*, const PassParams);
static TileKernel tile_kernel(int v, bool ring) {
    switch (v) {
        case 0: return ring ? k_tile_pass_ring<V_LAYERED> : k_tile_pass<V_LAYERED>;
        case 1: return ring ? k_tile_pass_ring<V_FOURIER> : k_tile_pass<V_FOURIER>;
        case 2: return ring ? k_tile_pass_ring<V_COMMON> : k_tile_pass<V_COMMON>;
        default: return ring ? k_tile_pass_ring<C_ALL> : k_tile_pass<C_ALL>;
    }
}

static int g_sm_count = 148;
static long g_ring_min_tiles = -1;   // DVD_RING_MIN_TILES: fewest tiles a pass needs for the ring form (default 8 per SM)
static int g_ring = 0;        // DVD_RING: 1 = dense passes with enough tiles run the two-group persistent form, 0 = one tile per CTA

long ring_min_tiles(int sm_count) { return g_ring_min_tiles >= 0 ? g_ring_min_tiles : 8l * sm_count; }

cudaError_t kernels_init() {
    for (int p = 0; p < 2; ++p)
        for (int v = 0; v < 4; ++v) {
            cudaError_t e = cudaFuncSetAttribute(tile_kernel(v, p != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 p ? RING_SMEM_BYTES : TILE_SLOTS * (int)sizeof(cplx));
            if (e != cudaSuccess) return e;
        }
    {
        cudaError_t e = cudaFuncSetAttribute(k_tile_pass<C_ALL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SLOTS * (int)sizeof(cplx));
        if (e != cudaSuccess) return e;
    }
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
        g_sm_count = sms;
    const char* e = getenv("DVD_RING");           // re-read at every dvd_create: absent means the default again
    g_ring = e ? atoi(e) != 0 : 0;
    e = getenv("DVD_RING_MIN_TILES");
    g_ring_min_tiles = e ? atol(e) : -1;
    return cudaFuncSetAttribute(k_sample, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SAMPLE_WARPS * (2 << BLK_BITS) * (int)sizeof(double));
}

cudaError_t launch_tile_pass(cplx* amp, PassParams& pp, cudaStream_t s) {
    const uint64_t ctas = 1ull << pp.pd.n_cta_bits;
    unsigned need = 0;
    int last_switch = -1;
    for (int k = 0; k < pp.pd.n_ops; ++k) {
        need |= op_class(pp.ops[k].code);
        if (pp.ops[k].code >= OC_SWITCH) last_switch = k;
    }
    pp.pd.last_switch = (int16_t)last_switch;
    int v = 0;
    while (v < 3 && (need & ~VARIANTS[v])) ++v;
    if (const char* e = getenv("DVD_KERNEL_VARIANT")) v = atoi(e) & 3;   // development: force a variant (3 = all ops)
    if (pp.pd.remap_st)      // the layout restore riding on this pass's store: one variant, all op classes
        k_tile_pass<C_ALL, true><<<(unsigned)ctas, NTHREADS, TILE_SLOTS * sizeof(cplx), s>>>(amp, pp);
    else if (g_ring && pp.pd.zero_mask == 0 && ctas >= (uint64_t)ring_min_tiles(g_sm_count))
        tile_kernel(v, true)<<<(unsigned)g_sm_count, RING_GROUPS * NTHREADS, RING_SMEM_BYTES, s>>>(amp, pp);
    else
        tile_kernel(v, false)<<<(unsigned)ctas, NTHREADS, TILE_SLOTS * sizeof(cplx), s>>>(amp, pp);
    return cudaGetLastError();
}

cudaError_t launch_simple_gate(cplx* amp, int n_local, uint64_t rank_bits, const SimpleOp& op, cudaStream_t s) {
    const uint64_t items = op.tbit < n_local ? (1ull << n_local) / 2 : (1ull << n_local);
    k_simple_gate<<<grid_for(items, 256, STREAM_CAP), 256, 0, s>>>(amp, n_local, rank_bits, op);
    return cudaGetLastError();
}

cudaError_t launch_set_basis_state(cplx* amp, uint64_t index, double value, cudaStream_t s) {
    k_set_basis_state<<<1, 1, 0, s>>>(amp, index, value);
    return cudaGetLastError();
}

cudaError_t launch_zero_outside_support(cplx* amp, int n_local, uint64_t zmask, cudaStream_t s) {
    if (zmask == 0) return cudaSuccess;
    k_zero_outside_support<<<grid_for(1ull << n_local, 256, STREAM_CAP), 256, 0, s>>>(amp, 1ull << n_local, zmask);
    return cudaGetLastError();
}

cudaError_t launch_probabilities(const cplx* amp, uint64_t first, uint64_t count, double* probs, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    k_probabilities<<<grid_for(count, 256, STREAM_CAP), 256, 0, s>>>(amp, first, count, probs);
    return cudaGetLastError();
}

cudaError_t launch_build_tree(const cplx* amp, int n_local, double* tree, cudaStream_t s) {
    if (n_local < BLK_BITS) {
        k_small_root<<<1, 256, 0, s>>>(amp, n_local, tree);
        return cudaGetLastError();
    }
    const uint64_t n_blocks = 1ull << (n_local - BLK_BITS);
    k_block_sums<<<grid_for(n_blocks, 1, 148 * 8 * 4), 256, 0, s>>>(amp, n_blocks, tree);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    for (int l = BLK_BITS + 1; l <= n_local; ++l) {
        const uint64_t cnt = 1ull << (n_local - l);
        k_tree_level<<<grid_for(cnt, 256, STREAM_CAP), 256, 0, s>>>(tree + tree_level_offset(n_local, l - 1),
                                                                  tree + tree_level_offset(n_local, l), cnt);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_sample(const cplx* amp, int n_local, const double* tree, const double* u,
                          const int32_t* sel, int32_t sel_value, uint64_t index_offset,
                          uint64_t shots, unsigned long long* out, cudaStream_t s) {
    if (shots == 0) return cudaSuccess;
    const int nb = n_local < BLK_BITS ? n_local : BLK_BITS;
    TreeOffsets offs;
    for (int l = 0; l < 64; ++l) offs.off[l] = (l >= nb && l <= n_local) ? tree_level_offset(n_local, l) : 0;
    const size_t smem = (size_t)SAMPLE_WARPS * (2ull << nb) * sizeof(double);
    k_sample<<<grid_for(shots, SAMPLE_WARPS, 148 * 3), SAMPLE_WARPS * 32, smem, s>>>(
        amp, n_local, nb, tree, offs, u, sel, sel_value, index_offset, shots, out);
    return cudaGetLastError();
}

cudaError_t launch_seq_block_cum(const cplx* amp, int n_local, double* block_cum, cudaStream_t s) {
    const int nb = n_local < BLK_BITS ? n_local : BLK_BITS;
    k_seq_block_cum<<<1, 32, 0, s>>>(amp, nb, 1ull << (n_local - nb), block_cum);
    return cudaGetLastError();
}
cudaError_t launch_sample_seq(const cplx* amp, int n_local, const double* block_cum, const double* u, const int32_t* sel,
                              int32_t sel_value, uint64_t index_offset, uint64_t shots, unsigned long long* out, cudaStream_t s) {
    if (shots == 0) return cudaSuccess;
    const int nb = n_local < BLK_BITS ? n_local : BLK_BITS;
    k_sample_seq<<<grid_for(shots, 128, 148 * 8), 128, 0, s>>>(amp, nb, 1ull << (n_local - nb), block_cum, u, sel, sel_value,
                                                               index_offset, shots, out);
    return cudaGetLastError();
}

cudaError_t launch_extract_expectation(const unsigned long long* samples, uint64_t shots, const int* qubits,
                                       int n_obs, double* out, cudaStream_t s) {
    const uint64_t total = shots * (uint64_t)n_obs;
    if (total == 0) return cudaSuccess;
    k_extract_expectation<<<grid_for(total, 256, STREAM_CAP), 256, 0, s>>>(samples, shots, qubits, n_obs, out);
    return cudaGetLastError();
}

cudaError_t launch_expectation_z(const cplx* amp, int n_local, int n_total, uint64_t rank_bits,
                                 double* partial, double* out, cudaStream_t s) {
    const int ctas = (int)grid_for(1ull << n_local, 256, EZ_CTAS);
    k_expect_z_partial<<<ctas, 256, 0, s>>>(amp, n_local, partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_expect_z_final<<<1, 64, 0, s>>>(partial, ctas, n_local, n_total, rank_bits, out);
    return cudaGetLastError();
}

cudaError_t launch_pack_half(const cplx* amp, int lq, int bitval, uint64_t first, uint64_t count, cplx* buf, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    k_pack_half<<<grid_for(count, 256, STREAM_CAP), 256, 0, s>>>(amp, lq, bitval, first, count, buf);
    return cudaGetLastError();
}
cudaError_t launch_unpack_half(cplx* amp, int lq, int bitval, uint64_t first, uint64_t count, const cplx* buf, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    k_unpack_half<<<grid_for(count, 256, STREAM_CAP), 256, 0, s>>>(amp, lq, bitval, first, count, buf);
    return cudaGetLastError();
}

cudaError_t launch_swap_peer(cplx* mine, cplx* peer, int lq, int my_bit, uint64_t e_begin, uint64_t e_end, cudaStream_t s) {
    if (e_end <= e_begin) return cudaSuccess;
    k_swap_peer<<<grid_for(e_end - e_begin, 256 * SWAP_UNROLL, 148 * 16), 256, 0, s>>>(mine, peer, lq, my_bit, e_begin, e_end);
    return cudaGetLastError();
}

cudaError_t launch_split_re_im(const cplx* amp, uint64_t count, double* re, double* im, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    k_split_re_im<<<grid_for(count, 256, STREAM_CAP), 256, 0, s>>>(amp, count, re, im);
    return cudaGetLastError();
}
cudaError_t launch_join_re_im(cplx* amp, uint64_t count, const double* re, const double* im, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    k_join_re_im<<<grid_for(count, 256, STREAM_CAP), 256, 0, s>>>(amp, count, re, im);
    return cudaGetLastError();
}

cudaError_t launch_dot(const cplx* a, const cplx* b, uint64_t count, double* partial, double* out, cudaStream_t s) {
    const int ctas = (int)grid_for(count, 256, DOT_CTAS);
    k_dot_partial<<<ctas, 256, 0, s>>>(a, b, count, partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_dot_final<<<1, 1, 0, s>>>(partial, ctas, out);
    return cudaGetLastError();
}

}  // namespace dvd
