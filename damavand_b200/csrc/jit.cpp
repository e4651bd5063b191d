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

std::vector<uint32_t> pass_structure_key(const Pass& p, bool persistent) {
    std::vector<uint32_t> key;
    key.reserve(p.ops.size() + 1);
    key.push_back(0x80000000u | (persistent ? 0x100u : 0u) | (uint32_t)p.desc.io_out);
    for (const DevOp& op : p.ops) {
        uint32_t k = (uint32_t)demacro(op.code) | ((uint32_t)op.flags << 8);
        if (op.code == OC_TABLE && op.tmask != 0) k |= 1u << 16;          // pivoted table op
        key.push_back(k);
    }
    return key;
}

std::string generate_pass_source(const Pass& p, const std::string& fn_name, bool persistent) {
    std::ostringstream o;
    const bool has_tab = !p.tab_desc.empty();
    int last_switch = -1;
    for (size_t k = 0; k < p.ops.size(); ++k) if (demacro(p.ops[k].code) >= OC_SWITCH) last_switch = (int)k;
    const char* ind = persistent ? "        " : "    ";
    const char* wcs = persistent ? "wcs" : "s_wc";
    // persistent form: fetch tile t + gridDim.x into the (now free) shared-memory tile, and its table constants
    auto prefetch_next = [&]() {
        o << ind << "if (tn < n_tiles) {\n"
          << ind << "    __syncthreads();   // every thread has read its registers back: the tile buffer is free\n"
          << ind << "    const uint64_t cb = cta_base_runs(pd, (uint64_t)tn);\n"
          << ind << "    tile_prefetch<IO_GROUP>(tile, amp, pd, cb, tid);\n";
        if (has_tab) o << ind << "    if (tid < n_tab) s_wc[buf ^ 1][tid] = table_cta_const(tables, tid, cb | pd.rank_bits);\n";
        o << ind << "}\n";
    };
    o << "#include \"tile_kernel.cuh\"\n"
         "using namespace dvd;\n"
         "extern \"C\" __global__ void __launch_bounds__(NTHREADS, 2)\n"
      << fn_name << "(cplx* __restrict__ amp, const __grid_constant__ PassParams pp) {\n"
         "    extern __shared__ __align__(16) unsigned char smem_raw[];\n"
         "    cplx* tile = reinterpret_cast<cplx*>(smem_raw);\n"
      << (persistent ? "    __shared__ cplx s_wc[2][MAX_TABLE_OPS];\n" : "    __shared__ cplx s_wc[MAX_TABLE_OPS];\n")
      << "    const PassDesc& pd = pp.pd;\n"
         "    const int tid = threadIdx.x;\n"
         "    const cplx* __restrict__ tables = pd.tables;\n"
         "    const int n_tab = pd.n_tab;\n";
    if (!persistent) {
        o << "    const uint64_t cbase = cta_base_runs(pd, (uint64_t)blockIdx.x);\n"
             "    const uint64_t gbase = cbase | pd.rank_bits;\n";
        if (has_tab) o << "    if (tid < n_tab) s_wc[tid] = table_cta_const(tables, tid, gbase);\n";
        o << "    if (cbase & pd.zero_mask) return;\n"
             "    cplx a[NREG];\n"
             "    tile_load<IO_GROUP>(amp, pd, a, cbase, tid);\n";
        if (has_tab) o << "    __syncthreads();\n";
    } else {
        o << "    const unsigned n_tiles = 1u << pd.n_cta_bits;   // dense states only (zero_mask == 0)\n"
             "    unsigned t = blockIdx.x;\n"
             "    if (t >= n_tiles) return;\n"
             "    {\n"
             "        const uint64_t cb = cta_base_runs(pd, (uint64_t)t);\n"
             "        tile_prefetch<IO_GROUP>(tile, amp, pd, cb, tid);\n";
        if (has_tab) o << "        if (tid < n_tab) s_wc[0][tid] = table_cta_const(tables, tid, cb | pd.rank_bits);\n";
        o << "    }\n"
             "    int buf = 0;\n"
             "    for (; t < n_tiles; t += gridDim.x, buf ^= 1) {\n"
             "        cplx a[NREG];\n"
             "        cp_async_wait_all();\n"
             "        __syncthreads();   // s_wc[buf] is visible; nobody still reads the previous tile's transposes\n"
             "        stage_load<IO_GROUP>(tile, a, tid);\n"
             "        const unsigned tn = t + gridDim.x;\n"
             "        const uint64_t gbase = cta_base_runs(pd, (uint64_t)t) | pd.rank_bits;\n"
             "        const cplx* wcs = s_wc[buf];\n";
        if (last_switch < 0) prefetch_next();
    }
    o << ind << "ThreadCtx ctx;\n"
      << ind << "ctx.pidx = gbase | tid_offset(pd, IO_GROUP, tid);\n"
      << ind << "ctx.ph = cplx{1.0, 0.0};\n"
      << ind << "ctx.ph_dirty = false;\n"
      << ind << "ctx.tid = tid;\n";
    for (size_t k = 0; k < p.ops.size(); ++k) {
        const DevOp& op = p.ops[k];
        const int code = demacro(op.code);
        const unsigned flags = op.flags;
        if (code >= OC_SWITCH) {
            const int from = (code - OC_SWITCH) / NGROUPS, to = (code - OC_SWITCH) % NGROUPS;
            o << ind << "flush_phase(a, ctx); __syncthreads();\n"
              << ind << "switch_store<" << from << ">(tile, a, tid, pp.ops[" << k << "], " << flags << "u, gbase); __syncthreads();\n"
              << ind << "stage_load<" << to << ">(tile, a, tid); ctx.pidx = gbase | tid_offset(pd, " << to << ", tid);\n";
            if (persistent && (int)k == last_switch) prefetch_next();
        } else {
            o << ind << "apply_op<C_ALL>(a, &pp.ops[" << k << "], " << code << ", " << flags << "u, ctx, tables, n_tab, " << wcs << ");\n";
        }
    }
    o << ind << "flush_phase(a, ctx);\n"
      << ind << "tile_store<" << (int)p.desc.io_out << ">(amp, pd, a, gbase - pd.rank_bits);\n";
    if (persistent) o << "    }\n";
    o << "}\n";
    return o.str();
}

}  // namespace dvd
