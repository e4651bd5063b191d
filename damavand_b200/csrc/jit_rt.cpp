// jit_rt.cpp -- see jit_rt.h.
#include "jit_rt.h"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "jit.h"
#include "jit_headers.inc"   // k_src_tile_core, k_src_tile_kernel: the two device headers as text (build.py)

namespace dvd {

namespace {

struct Api {
    void* h_nvrtc = nullptr;
    void* h_cuda = nullptr;
    std::string why;      // non-empty: unusable
    // NVRTC
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    // driver API
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;

    static void* open_first(const char* const* names) {
        for (const char* const* n = names; *n; ++n)
            if (void* h = dlopen(*n, RTLD_NOW | RTLD_GLOBAL)) return h;
        return nullptr;
    }
    // need_driver == false: NVRTC only (the generator test runs where no driver exists)
    const std::string& load(bool need_driver) {
        if (!h_nvrtc) {
            static const char* const names[] = {"libnvrtc.so.12", "libnvrtc.so", nullptr};
            h_nvrtc = open_first(names);
            if (!h_nvrtc) { why = std::string("dlopen(libnvrtc.so.12) failed: ") + dlerror(); return why; }
#define DVD_SYM(h, field, name)                                                  \
    field = reinterpret_cast<decltype(field)>(dlsym(h, name));                   \
    if (!field) { why = std::string("symbol missing: ") + name; h = nullptr; return why; }   /* retried next call */
            DVD_SYM(h_nvrtc, CreateProgram, "nvrtcCreateProgram")
            DVD_SYM(h_nvrtc, CompileProgram, "nvrtcCompileProgram")
            DVD_SYM(h_nvrtc, GetCUBINSize, "nvrtcGetCUBINSize")
            DVD_SYM(h_nvrtc, GetCUBIN, "nvrtcGetCUBIN")
            DVD_SYM(h_nvrtc, GetProgramLogSize, "nvrtcGetProgramLogSize")
            DVD_SYM(h_nvrtc, GetProgramLog, "nvrtcGetProgramLog")
            DVD_SYM(h_nvrtc, DestroyProgram, "nvrtcDestroyProgram")
        }
        if (need_driver && !h_cuda) {
            static const char* const names[] = {"libcuda.so.1", "libcuda.so", nullptr};
            h_cuda = open_first(names);
            if (!h_cuda) { why = std::string("dlopen(libcuda.so.1) failed: ") + dlerror(); return why; }
            DVD_SYM(h_cuda, ModuleLoadData, "cuModuleLoadData")
            DVD_SYM(h_cuda, ModuleGetFunction, "cuModuleGetFunction")
            DVD_SYM(h_cuda, FuncSetAttribute, "cuFuncSetAttribute")
            DVD_SYM(h_cuda, LaunchKernel, "cuLaunchKernel")
            DVD_SYM(h_cuda, GetErrorString, "cuGetErrorString")
#undef DVD_SYM
        }
        why.clear();
        return why;
    }
    std::string cu_error(CUresult r) const {
        const char* s = nullptr;
        if (GetErrorString && GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
        return "CUresult " + std::to_string((int)r);
    }
};

struct Entry {
    enum State { QUEUED, COMPILED, FAILED } state = QUEUED;
    std::vector<char> cubin;
    std::string log;
    std::map<int, CUfunction> fn;    // per device (module loaded in that device's primary context)
};

struct Runtime {
    std::mutex mu;
    std::condition_variable cv;
    Api api;
    std::map<std::vector<uint32_t>, std::shared_ptr<Entry>> entries;
    std::deque<std::pair<std::shared_ptr<Entry>, std::string>> queue;
    std::vector<std::thread> workers;
    std::map<int, int> sm_count;     // per device
    int busy = 0;
    bool stop = false;
    JitStats stats;

    ~Runtime() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
            queue.clear();
        }
        cv.notify_all();
        for (auto& t : workers) if (t.joinable()) t.join();
    }
    void start_workers() {   // mu held
        if (!workers.empty()) return;
        unsigned n = std::thread::hardware_concurrency();
        n = n == 0 ? 2 : (n > 4 ? 4 : n);
        for (unsigned i = 0; i < n; ++i) workers.emplace_back([this] { work(); });
    }
    void work() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [this] { return stop || !queue.empty(); });
            if (stop) return;
            auto item = std::move(queue.front());
            queue.pop_front();
            ++busy;
            lk.unlock();
            std::vector<char> cubin;
            const auto t0 = std::chrono::steady_clock::now();
            std::string log = jit_compile(item.second, &cubin);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            lk.lock();
            --busy;
            stats.compile_seconds += dt;
            if (log.empty()) { item.first->cubin = std::move(cubin); item.first->state = Entry::COMPILED; ++stats.compiled; }
            else { item.first->log = std::move(log); item.first->state = Entry::FAILED; ++stats.failed; }
            cv.notify_all();
        }
    }
};

Runtime& rt() {
    static Runtime r;
    return r;
}

}  // namespace

namespace {

// Disk cache of compiled cubins, opt-in: the directory named by $DVD_JIT_CACHE_DIR (nothing is written anywhere
// unless it is set).  File name = FNV-1a hash of the generated source and of both embedded headers, so a rebuilt
// library with different device code never picks up a stale kernel.
std::string cache_dir() {
    const char* e = getenv("DVD_JIT_CACHE_DIR");
    return e ? std::string(e) : std::string();
}
uint64_t fnv1a(uint64_t h, const char* p, size_t n) {
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
    return h;
}
std::string cache_path(const std::string& source) {
    const std::string dir = cache_dir();
    if (dir.empty()) return "";
    uint64_t h = 1469598103934665603ull;
    h = fnv1a(h, source.data(), source.size());
    h = fnv1a(h, k_src_tile_core, sizeof k_src_tile_core);
    h = fnv1a(h, k_src_tile_kernel, sizeof k_src_tile_kernel);
    char name[64];
    snprintf(name, sizeof name, "/%016llx.sm_100a.cubin", (unsigned long long)h);
    return dir + name;
}
void mkdirs(const std::string& dir) {
    for (size_t i = 1; i <= dir.size(); ++i)
        if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0755);
}
bool cache_read(const std::string& path, std::vector<char>* out) {
    if (path.empty()) return false;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    bool ok = n > 64;
    if (ok) { out->resize((size_t)n); ok = fread(out->data(), 1, (size_t)n, f) == (size_t)n; }
    fclose(f);
    return ok && (*out)[0] == 0x7f && (*out)[1] == 'E' && (*out)[2] == 'L' && (*out)[3] == 'F';   // a cubin is an ELF image
}
void cache_write(const std::string& path, const std::vector<char>& data) {
    if (path.empty()) return;
    mkdirs(path.substr(0, path.rfind('/')));
    const std::string tmp = path + ".tmp." + std::to_string((long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());   // rename: readers never see a partial file
}

}  // namespace

std::string jit_compile(const std::string& source, std::vector<char>* cubin) {
    const std::string cpath = cache_path(source);
    if (cache_read(cpath, cubin)) return "";
    Api* api;
    {
        Runtime& r = rt();
        std::lock_guard<std::mutex> lk(r.mu);
        const std::string why = r.api.load(false);
        if (!why.empty()) return why;
        api = &r.api;
    }
    const char* hdr_src[] = {k_src_tile_core, k_src_tile_kernel};
    const char* hdr_name[] = {"tile_core.cuh", "tile_kernel.cuh"};
    nvrtcProgram prog = nullptr;
    if (api->CreateProgram(&prog, source.c_str(), "dvd_pass.cu", 2, hdr_src, hdr_name) != NVRTC_SUCCESS) return "nvrtcCreateProgram failed";
    // same code generation as the in-tree build: sm_100a, C++17, fma contraction on (nvcc's default)
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=true"};
    const nvrtcResult rc = api->CompileProgram(prog, 3, opts);
    std::string log;
    if (rc != NVRTC_SUCCESS) {
        size_t n = 0;
        api->GetProgramLogSize(prog, &n);
        log.resize(n);
        if (n) api->GetProgramLog(prog, &log[0]);
        if (log.empty()) log = "nvrtcCompileProgram failed";
    } else {
        size_t n = 0;
        if (api->GetCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) log = "nvrtcGetCUBINSize failed";
        else {
            cubin->resize(n);
            if (api->GetCUBIN(prog, cubin->data()) != NVRTC_SUCCESS) log = "nvrtcGetCUBIN failed";
        }
    }
    api->DestroyProgram(&prog);
    if (log.empty()) cache_write(cpath, *cubin);
    return log;
}

std::string jit_available() {
    Runtime& r = rt();
    std::lock_guard<std::mutex> lk(r.mu);
    return r.api.load(true);
}

bool jit_launch(const Pass& p, int mode, int device, cplx* amp, const PassParams& pp, cudaStream_t stream, std::string* err) {
    if (mode == JIT_OFF) return false;
    Runtime& r = rt();
    CUfunction fn = nullptr;
    // experimental persistent (cp.async prefetch) form: dense states only, off unless DVD_JIT_PERSIST=1
    const char* pe = getenv("DVD_JIT_PERSIST");
    bool persistent = pe && atoi(pe) != 0 && pp.pd.zero_mask == 0;
    unsigned resident = 0;
    {
        std::unique_lock<std::mutex> lk(r.mu);
        if (!r.api.load(true).empty()) return false;
        if (persistent && r.sm_count.find(device) == r.sm_count.end()) {
            int sms = 0;
            if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 0; }
            r.sm_count[device] = sms;
        }
        if (persistent) {
            resident = 2u * (unsigned)r.sm_count[device];      // __launch_bounds__(NTHREADS, 2): two CTAs per SM
            if (resident == 0 || (1u << pp.pd.n_cta_bits) <= resident) persistent = false;
        }
        const std::vector<uint32_t> key = pass_structure_key(p, persistent);
        std::shared_ptr<Entry>& slot = r.entries[key];
        if (!slot) {
            slot = std::make_shared<Entry>();
            r.queue.emplace_back(slot, generate_pass_source(p, "dvd_pass_static", persistent));
            ++r.stats.pending;
            r.start_workers();
            r.cv.notify_all();
        }
        std::shared_ptr<Entry> e = slot;
        if (mode == JIT_SYNC) r.cv.wait(lk, [&] { return e->state != Entry::QUEUED; });
        if (e->state != Entry::COMPILED) {
            if (e->state == Entry::FAILED && err && !e->log.empty()) { *err = "jit compile: " + e->log; e->log.clear(); }
            return false;
        }
        auto it = e->fn.find(device);
        if (it == e->fn.end()) {
            CUmodule mod = nullptr;
            CUresult rc = r.api.ModuleLoadData(&mod, e->cubin.data());
            if (rc == CUDA_SUCCESS) rc = r.api.ModuleGetFunction(&fn, mod, "dvd_pass_static");
            if (rc == CUDA_SUCCESS)
                rc = r.api.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, TILE_SLOTS * (int)sizeof(cplx));
            if (rc != CUDA_SUCCESS) {
                if (err) *err = "jit module load: " + r.api.cu_error(rc);
                e->state = Entry::FAILED;
                return false;
            }
            e->fn[device] = fn;
        } else {
            fn = it->second;
        }
    }
    const unsigned ctas = persistent ? resident : 1u << pp.pd.n_cta_bits;
    void* params[] = {(void*)&amp, (void*)&pp};
    const CUresult rc = r.api.LaunchKernel(fn, ctas, 1, 1, NTHREADS, 1, 1, TILE_SLOTS * (unsigned)sizeof(cplx), (CUstream)stream, params, nullptr);
    if (rc != CUDA_SUCCESS) {
        if (err) *err = "jit launch: " + r.api.cu_error(rc);
        return false;
    }
    return true;
}

void jit_wait() {
    Runtime& r = rt();
    std::unique_lock<std::mutex> lk(r.mu);
    r.cv.wait(lk, [&] { return r.stop || (r.queue.empty() && r.busy == 0); });
}

JitStats jit_stats() {
    Runtime& r = rt();
    std::lock_guard<std::mutex> lk(r.mu);
    JitStats s = r.stats;
    s.pending = (long)r.queue.size() + r.busy;
    return s;
}

}  // namespace dvd
