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

// Structure key of a pass: equal keys <=> identical generated source.
std::vector<uint32_t> pass_structure_key(const Pass& p, bool persistent = false);

// CUDA source of `extern "C" __global__ void <fn_name>(cplx*, const PassParams)`; expects tile_kernel.cuh to be
// includable under that name.  persistent: one CTA per resident slot loops over the tiles and fetches the next
// tile with cp.async once the last transpose of the current one has been read back (the specialised form of
// k_tile_pass_persist; dense states only).  Experimental: not yet measured on a GPU, off unless DVD_JIT_PERSIST=1.
std::string generate_pass_source(const Pass& p, const std::string& fn_name, bool persistent = false);

}  // namespace dvd
