// jit.cpp -- see jit.h.  Pure C++ (no CUDA calls): also compiled into the CPU replay library.
#include "jit.h"

#include <sstream>

namespace dvd {

namespace {

// The op a macro-op marker stands for (fuse_macro_ops overwrites the first code of a 4-run).
int demacro(int code) {
    if (code == OC_REALPH4) return OC_GATE + 4 * K_REALPH + 0;
    if (code == OC_TWHAD4) return OC_TWHAD + 0;
    return code;
}

}  // namespace

std::vector<uint32_t> pass_structure_key(const Pass& p, int form, bool store_remap) {
    std::vector<uint32_t> key;
    key.reserve(p.ops.size() + 1);
    key.push_back(0x80000000u | (store_remap ? 1u << 16 : 0u) | ((uint32_t)form << 8) | (uint32_t)p.desc.io_out);
    for (const DevOp& op : p.ops) {
        uint32_t k = (uint32_t)demacro(op.code) | ((uint32_t)op.flags << 8);
        if (op.code == OC_TABLE && op.tmask != 0) k |= 1u << 16;          // pivoted table op
        key.push_back(k);
    }
    return key;
}

std::string generate_pass_source(const Pass& p, const std::string& fn_name, int form, bool store_remap) {
    std::ostringstream o;
    const bool ring = form == FORM_RING;
    const bool has_tab = !p.tab_desc.empty();
    int last_switch = -1;
    for (size_t k = 0; k < p.ops.size(); ++k) if (demacro(p.ops[k].code) >= OC_SWITCH) last_switch = (int)k;
    const char* ind = ring ? "        " : "    ";
    const char* wcs = ring ? "wcs" : "s_wc";
    const char* sync = ring ? "group_sync(grp)" : "__syncthreads()";
    // ring form: the slot's buffer is free; the same threads fetch slot + 3 into it and prepare the table constants
    // of this group's next slot
    auto release_buffer = [&]() {
        o << ind << "group_sync(grp);   // every thread has read its registers back: the buffer is free\n"
          << ind << "if (t + 3 * stride < n_tiles) ring_fetch(ring, slot + 3, amp_in, pd, t + 3 * stride, tid, toff.get(IO_GROUP));\n";
        if (has_tab)
            o << ind << "if (tid < n_tab && t + 2 * stride < n_tiles)\n"
              << ind << "    ring.wcs(grp, k + 1)[tid] = table_cta_const(tables, tid, cta_base_runs(pd, t + 2 * stride) | pd.rank_bits);\n";
    };
    o << "#include \"tile_kernel.cuh\"\n"
         "using namespace dvd;\n"
         "extern \"C\" __global__ void __launch_bounds__("
      << (ring ? "RING_GROUPS * NTHREADS, 1" : form == FORM_CLASSIC3 ? "NTHREADS, 3" : "NTHREADS, 2") << ")\n"
      << fn_name << "(cplx* __restrict__ amp, const __grid_constant__ PassParams pp) {\n"
         "    extern __shared__ __align__(16) unsigned char smem_raw[];\n"
         "    const PassDesc& pd = pp.pd;\n"
         "    const cplx* __restrict__ tables = pd.tables;\n"
         "    const int n_tab = pd.n_tab;\n"
         "    const cplx* amp_in = amp;   // in place unless the load carries a fused remap (then pd.remap_src is read)\n";
    if (!ring) {
        o << "    cplx* tile = reinterpret_cast<cplx*>(smem_raw);\n"
             "    __shared__ cplx s_wc[MAX_TABLE_OPS];\n"
             "    __shared__ uint32_t s_toff[TOFF_WORDS];\n"
             "    const int tid = threadIdx.x;\n"
             "    const uint64_t cbase = cta_base_runs(pd, (uint64_t)blockIdx.x);\n"
             "    const uint64_t gbase = cbase | pd.rank_bits;\n";
        if (has_tab) o << "    if (tid < n_tab) s_wc[tid] = table_cta_const(tables, tid, gbase);\n";
        o << "    if (cbase & pd.zero_mask & ~pd.remap_lmask) return;\n"
             "    cplx a[NREG];\n"
             "    tile_load<IO_GROUP>(amp_in, pd, a, cbase, tid_offset_arith(pd, IO_GROUP, tid));\n"
             "    const Toff toff = toff_fill(s_toff, pd, tid);   // table lookups in the shadow of the tile's loads\n";
        if (has_tab) o << "    __syncthreads();\n";
    } else {
        o << "    const Ring ring = ring_setup(smem_raw);\n"
             "    const int grp = threadIdx.x / NTHREADS, tid = threadIdx.x % NTHREADS;\n"
             "    const uint32_t* s_toff = ring.toff;\n"
             "    const Toff toff = toff_fill(ring.toff, pd, tid);   // once per CTA: the offsets do not depend on the tile\n"
             "    const unsigned n_tiles = 1u << pd.n_cta_bits, stride = gridDim.x;   // dense states only (zero_mask == 0)\n"
             "    {   // prologue: slots 0 and 2 are fetched by group 0, slot 1 by group 1\n"
             "        const unsigned t0 = blockIdx.x + grp * stride;\n"
             "        if (t0 < n_tiles) ring_fetch(ring, grp, amp_in, pd, t0, tid, toff.get(IO_GROUP));\n"
             "        if (grp == 0 && blockIdx.x + 2 * stride < n_tiles) ring_fetch(ring, 2, amp_in, pd, blockIdx.x + 2 * stride, tid, toff.get(IO_GROUP));\n";
        if (has_tab) o << "        if (tid < n_tab && t0 < n_tiles) ring.wcs(grp, 0)[tid] = table_cta_const(tables, tid, cta_base_runs(pd, t0) | pd.rank_bits);\n";
        o << "    }\n"
             "    unsigned k = 0;\n"
             "    for (unsigned slot = grp; ; slot += RING_GROUPS, ++k) {\n"
             "        const unsigned t = blockIdx.x + slot * stride;\n"
             "        if (t >= n_tiles) break;\n"
             "        cplx* tile = ring.tile(slot);\n"
             "        const cplx* wcs = ring.wcs(grp, k);\n"
             "        const uint64_t gbase = cta_base_runs(pd, t) | pd.rank_bits;\n"
             "        ring_wait(ring, slot);\n"
             "        group_sync(grp);   // this group's previous slot is over everywhere; its table constants are visible\n"
             "        cplx a[NREG];\n"
             "        stage_load<IO_GROUP>(tile, a, tid);\n";
        if (last_switch < 0) release_buffer();
    }
    o << ind << "ThreadCtx ctx;\n"
      << ind << "ctx.pidx = gbase | toff.get(IO_GROUP);\n"
      << ind << "ctx.ph = cplx{1.0, 0.0};\n"
      << ind << "ctx.ph_dirty = false;\n"
      << ind << "ctx.tid = tid;\n";
    for (size_t k = 0; k < p.ops.size(); ++k) {
        const DevOp& op = p.ops[k];
        const int code = demacro(op.code);
        const unsigned flags = op.flags;
        if (code >= OC_SWITCH) {
            const int from = (code - OC_SWITCH) / NGROUPS, to = (code - OC_SWITCH) % NGROUPS;
            o << ind << "flush_phase(a, ctx); " << sync << ";\n"
              << ind << "switch_store<" << from << ">(tile, a, tid, pp.ops[" << k << "], " << flags << "u, gbase); " << sync << ";\n"
              << ind << "stage_load<" << to << ">(tile, a, tid); ctx.pidx = gbase | toff.get(" << to << ");\n";
            if (ring && (int)k == last_switch) release_buffer();
        } else {
            o << ind << "apply_op<C_ALL>(a, &pp.ops[" << k << "], " << code << ", " << flags << "u, ctx, tables, n_tab, " << wcs << ");\n";
        }
    }
    o << ind << "flush_phase(a, ctx);\n"
      << ind << "tile_store<" << (int)p.desc.io_out << (store_remap ? ", true" : "") << ">(amp, pd, a, gbase - pd.rank_bits, s_toff);\n";
    if (ring) o << "    }\n";
    o << "}\n";
    return o.str();
}

}  // namespace dvd
