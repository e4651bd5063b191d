// jit_rt.h -- run-time compilation and launch of the structure-specialised pass kernels (source: jit.h).
//
// NVRTC (libnvrtc.so.12) and the driver API (libcuda.so.1) are resolved with dlopen: the library still loads and
// runs (interpreter kernels) where either is missing.  Compilation runs on a background thread and produces a
// cubin for sm_100a -- no CUDA context involved; the module is loaded by the launching thread, in the primary
// context of the state's device, the first time the kernel is wanted after its cubin is ready.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "planner.h"

namespace dvd {

enum JitMode : int {
    JIT_OFF = 0,     // interpreter kernels only
    JIT_ASYNC = 1,   // compile in the background, use a kernel once it is ready (dvd_jit_wait to block)
    JIT_SYNC = 2,    // compile on first use (tests)
};

// Empty string when NVRTC and the driver API are usable, else the reason.
std::string jit_available();

// Launch pass `p` through its specialised kernel if one is ready (or, in JIT_SYNC mode, as soon as it has been
// compiled).  Returns true when the launch was issued; false = caller falls back to the interpreter kernel
// (JIT_ASYNC queues the compilation on the first miss).  `err` receives a message when a compiled kernel failed
// to load or launch (the key is then blacklisted).
bool jit_launch(const Pass& p, int mode, int device, cplx* amp, const PassParams& pp, cudaStream_t stream, std::string* err);

// Block until every queued compilation has finished.
void jit_wait();

struct JitStats {
    long compiled = 0, failed = 0, pending = 0;
    double compile_seconds = 0.0;
    long tuning = 0;          // pass structures whose kernel form is still being measured
    long chosen[3] = {0, 0, 0};    // structures per chosen form (jit.h: JitForm)
    long launches[3] = {0, 0, 0};  // launches per form
};
JitStats jit_stats();

// Compile a source with NVRTC to a cubin (used by jit_launch and by the CPU-side test of the generator).
// Returns empty string on success, else the compiler log.
std::string jit_compile(const std::string& source, std::vector<char>* cubin);

}  // namespace dvd
