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

std::vector<uint32_t> pass_structure_key(const Pass& p) {
    std::vector<uint32_t> key;
    key.reserve(p.ops.size() + 1);
    key.push_back(0x80000000u | (uint32_t)p.desc.io_out);
    for (const DevOp& op : p.ops) {
        uint32_t k = (uint32_t)demacro(op.code) | ((uint32_t)op.flags << 8);
        if (op.code == OC_TABLE && op.tmask != 0) k |= 1u << 16;          // pivoted table op
        key.push_back(k);
    }
    return key;
}

std::string generate_pass_source(const Pass& p, const std::string& fn_name) {
    std::ostringstream o;
    o << "#include \"tile_kernel.cuh\"\n"
         "using namespace dvd;\n"
         "extern \"C\" __global__ void __launch_bounds__(NTHREADS, 2)\n"
      << fn_name << "(cplx* __restrict__ amp, const __grid_constant__ PassParams pp) {\n"
         "    extern __shared__ __align__(16) unsigned char smem_raw[];\n"
         "    cplx* tile = reinterpret_cast<cplx*>(smem_raw);\n"
         "    __shared__ cplx s_wc[MAX_TABLE_OPS];\n"
         "    const PassDesc& pd = pp.pd;\n"
         "    const int tid = threadIdx.x;\n"
         "    const uint64_t cbase = cta_base_runs(pd, (uint64_t)blockIdx.x);\n"
         "    const uint64_t gbase = cbase | pd.rank_bits;\n"
         "    const cplx* __restrict__ tables = pd.tables;\n"
         "    const int n_tab = pd.n_tab;\n";
    const bool has_tab = !p.tab_desc.empty();
    if (has_tab) o << "    if (tid < n_tab) s_wc[tid] = table_cta_const(tables, tid, gbase);\n";
    o << "    const uint64_t zmask = pd.zero_mask;\n"
         "    if (cbase & zmask) return;\n"
         "    cplx a[NREG];\n"
         "    tile_load<IO_GROUP>(amp, pd, a, cbase, tid);\n";
    if (has_tab) o << "    __syncthreads();\n";
    o << "    ThreadCtx ctx;\n"
         "    ctx.pidx = gbase | tid_offset(pd, IO_GROUP, tid);\n"
         "    ctx.ph = cplx{1.0, 0.0};\n"
         "    ctx.ph_dirty = false;\n"
         "    ctx.tid = tid;\n";
    for (size_t k = 0; k < p.ops.size(); ++k) {
        const DevOp& op = p.ops[k];
        const int code = demacro(op.code);
        const unsigned flags = op.flags;
        if (code >= OC_SWITCH) {
            const int from = (code - OC_SWITCH) / NGROUPS, to = (code - OC_SWITCH) % NGROUPS;
            o << "    flush_phase(a, ctx); __syncthreads();\n"
              << "    switch_store<" << from << ">(tile, a, tid, pp.ops[" << k << "], " << flags << "u, gbase); __syncthreads();\n"
              << "    stage_load<" << to << ">(tile, a, tid); ctx.pidx = gbase | tid_offset(pd, " << to << ", tid);\n";
        } else if (code == OC_TABLE) {
            // the pivot test reads op.tmask: make the unpivoted form a literal too
            o << "    apply_op<C_ALL>(a, &pp.ops[" << k << "], " << code << ", " << flags << "u, ctx, tables, n_tab, s_wc);\n";
        } else {
            o << "    apply_op<C_ALL>(a, &pp.ops[" << k << "], " << code << ", " << flags << "u, ctx, tables, n_tab, s_wc);\n";
        }
    }
    o << "    flush_phase(a, ctx);\n"
      << "    tile_store<" << (int)p.desc.io_out << ">(amp, pd, a, gbase - pd.rank_bits);\n"
         "}\n";
    return o.str();
}

}  // namespace dvd
