// jit.h -- structure-specialised pass kernels (source generation; compilation lives in jit_rt.cu).
//
// k_tile_pass is an interpreter: every op costs a constant-bank fetch of its code, a branch tree, and a join at
// which all 64 amplitude registers must sit in canonical places.  The STRUCTURE of a pass -- the sequence of
// (opcode, flags, register group) -- is known on the host when the pass is planned; the operands (matrices,
// masks, tables) are not part of it.  generate_pass_source() writes a kernel whose body is that sequence as
// straight-line calls of the same apply_op / transpose helpers with the code and flags as literals, so the
// switch, the flag tests and the joins fold away at compile time, while angles stay run-time data in the
// constant bank: one compiled kernel serves every parameter set of a variational circuit.
#pragma once
#include <string>
#include <vector>

#include "planner.h"

namespace dvd {

// Kernel forms a pass structure can be compiled into (jit_rt picks per structure, by measurement):
//   FORM_CLASSIC2  one tile per CTA, 256 threads, two CTAs per SM (<= 128 registers);
//   FORM_CLASSIC3  the same with __launch_bounds__(256, 3): three CTAs = 24 warps per SM if it fits 80 registers
//                  (kept only where ptxas needed no or next to no local memory);
//   FORM_RING      persistent, one CTA of 2 x 256 threads per SM, three shared-memory tile buffers filled by cp.async a
//                  full tile period ahead (tile_kernel.cuh, "two-group persistent form"); dense states only.
enum JitForm : int { FORM_CLASSIC2 = 0, FORM_CLASSIC3 = 1, FORM_RING = 2, FORM_COUNT = 3 };

// Structure key of a pass: equal keys <=> identical generated source.
// store_remap: the variant whose store goes through a remap (PassDesc::remap_st; tile_kernel.cuh: tile_store<G, true>).
std::vector<uint32_t> pass_structure_key(const Pass& p, int form = FORM_CLASSIC2, bool store_remap = false);

// CUDA source of `extern "C" __global__ void <fn_name>(cplx*, const PassParams)`; expects tile_kernel.cuh to be
// includable under that name.
std::string generate_pass_source(const Pass& p, const std::string& fn_name, int form = FORM_CLASSIC2, bool store_remap = false);

}  // namespace dvd
