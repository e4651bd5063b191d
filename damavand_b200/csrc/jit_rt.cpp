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
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
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
            DVD_SYM(h_cuda, FuncGetAttribute, "cuFuncGetAttribute")
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
    int form = FORM_CLASSIC2;
    std::vector<char> cubin;
    std::string log;
    struct Loaded { CUfunction fn = nullptr; int local_bytes = 0; int regs = 0; };
    std::map<int, Loaded> fn;    // per device (module loaded in that device's primary context)
};

// One pass structure: its kernel forms and the choice between them.  The choice is MEASURED: once every candidate
// has been compiled, dense full-grid launches of the structure rotate through the candidates bracketed by CUDA
// events until each has TUNE_SAMPLES timings of the same grid; the fastest is kept (per device).
constexpr int TUNE_SAMPLES = 2;
struct Tuner {
    std::shared_ptr<Entry> e[FORM_COUNT];
    struct Sample { cudaEvent_t t0 = nullptr, t1 = nullptr; int form = 0; };
    struct PerDevice {
        int best_dense = -1;          // decided form for dense launches (-1: still measuring)
        int best_classic = -1;        // decided classic form (launches with a support mask, or too few tiles for the ring)
        int sig = -1;                 // n_cta_bits of the launches being compared
        std::vector<float> ms[FORM_COUNT];
        std::vector<Sample> pending;
        bool dropped[FORM_COUNT] = {false, false, false};   // form failed to load / needs too much local memory
    };
    std::map<int, PerDevice> dev;
};

struct Runtime {
    std::mutex mu;
    std::condition_variable cv;
    Api api;
    std::map<std::vector<uint32_t>, std::shared_ptr<Tuner>> tuners;   // key: pass_structure_key(p, FORM_CLASSIC2)
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
        n = n == 0 ? 2 : (n > 12 ? 12 : n);
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

namespace {

// DVD_JIT_FORM: classic2 | classic3 | ring force one kernel form (development, parity tests); anything else = measure.
int forced_form() {
    const char* e = getenv("DVD_JIT_FORM");
    if (!e) return -1;
    const std::string v(e);
    if (v == "classic2") return FORM_CLASSIC2;
    if (v == "classic3") return FORM_CLASSIC3;
    if (v == "ring") return FORM_RING;
    return -1;
}
// a classic3 kernel that had to spill more than this is not worth measuring
constexpr int CLASSIC3_MAX_LOCAL_BYTES = 256;

// Load `e` on `device` if it has not been yet.  mu held.  Returns nullptr when the kernel cannot be used.
const Entry::Loaded* loaded_fn(Runtime& r, Entry& e, int device, std::string* err) {
    if (e.state != Entry::COMPILED) return nullptr;
    auto it = e.fn.find(device);
    if (it != e.fn.end()) return &it->second;
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    CUresult rc = r.api.ModuleLoadData(&mod, e.cubin.data());
    if (rc == CUDA_SUCCESS) rc = r.api.ModuleGetFunction(&fn, mod, "dvd_pass_static");
    const int smem = e.form == FORM_RING ? RING_SMEM_BYTES : TILE_SLOTS * (int)sizeof(cplx);
    if (rc == CUDA_SUCCESS) rc = r.api.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem);
    if (rc != CUDA_SUCCESS) {
        if (err) *err = "jit module load: " + r.api.cu_error(rc);
        e.state = Entry::FAILED;
        return nullptr;
    }
    Entry::Loaded l;
    l.fn = fn;
    r.api.FuncGetAttribute(&l.local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, fn);
    r.api.FuncGetAttribute(&l.regs, CU_FUNC_ATTRIBUTE_NUM_REGS, fn);
    return &(e.fn[device] = l);
}

// Collect the tuning launches whose events have completed.  mu held.
void harvest(Tuner::PerDevice& d) {
    for (size_t i = 0; i < d.pending.size();) {
        Tuner::Sample& sm = d.pending[i];
        if (cudaEventQuery(sm.t1) != cudaSuccess) { cudaGetLastError(); ++i; continue; }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sm.t0, sm.t1) == cudaSuccess) d.ms[sm.form].push_back(ms); else cudaGetLastError();
        cudaEventDestroy(sm.t0); cudaEventDestroy(sm.t1);
        d.pending.erase(d.pending.begin() + (long)i);
    }
}

}  // namespace

bool jit_launch(const Pass& p, int mode, int device, cplx* amp, const PassParams& pp, cudaStream_t stream, std::string* err) {
    if (mode == JIT_OFF) return false;
    Runtime& r = rt();
    CUfunction fn = nullptr;
    int form = FORM_CLASSIC2;
    unsigned sms = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::shared_ptr<Tuner> tuner;
    {
        std::unique_lock<std::mutex> lk(r.mu);
        if (!r.api.load(true).empty()) return false;
        if (r.sm_count.find(device) == r.sm_count.end()) {
            int n = 0;
            if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 0; }
            r.sm_count[device] = n;
        }
        sms = (unsigned)r.sm_count[device];
        const bool st = pp.pd.remap_st != 0;       // the store-side remap is a kernel variant of its own (never in the ring form)
        std::shared_ptr<Tuner>& slot = r.tuners[pass_structure_key(p, FORM_CLASSIC2, st)];
        if (!slot) slot = std::make_shared<Tuner>();
        {   // queue the forms this launch may want and that nobody asked for yet (classic2 always: it can run every launch)
            const int forced = forced_form();
            bool queued = false;
            for (int f = 0; f < FORM_COUNT; ++f) {
                if (slot->e[f] || (forced >= 0 && f != forced && f != FORM_CLASSIC2) || (st && f == FORM_RING)) continue;
                slot->e[f] = std::make_shared<Entry>();
                slot->e[f]->form = f;
                r.queue.emplace_back(slot->e[f], generate_pass_source(p, "dvd_pass_static", f, st));
                ++r.stats.pending;
                queued = true;
            }
            if (queued) { r.start_workers(); r.cv.notify_all(); }
        }
        tuner = slot;
        Tuner& t = *tuner;
        auto all_done = [&] {
            for (auto& e : t.e) if (e && e->state == Entry::QUEUED) return false;
            return true;
        };
        if (mode == JIT_SYNC) r.cv.wait(lk, all_done);
        // the ring form needs a dense state and enough tiles to keep every SM's ring turning
        const unsigned n_tiles = 1u << pp.pd.n_cta_bits;
        const char* mt = getenv("DVD_RING_MIN_TILES");
        const bool ring_ok = pp.pd.zero_mask == 0 && sms > 0 && (long)n_tiles >= (mt ? atol(mt) : 8l * (long)sms);
        Tuner::PerDevice& d = t.dev[device];
        auto usable = [&](int f) -> const Entry::Loaded* {
            if (!t.e[f] || d.dropped[f]) return nullptr;
            std::string lerr;
            const Entry::Loaded* l = loaded_fn(r, *t.e[f], device, &lerr);
            if (!l) {
                if (t.e[f]->state == Entry::FAILED) {
                    d.dropped[f] = true;
                    if (err && !t.e[f]->log.empty()) { *err = "jit compile: " + t.e[f]->log; t.e[f]->log.clear(); }
                    else if (err && !lerr.empty()) *err = lerr;
                }
                return nullptr;
            }
            if (f == FORM_CLASSIC3 && l->local_bytes > CLASSIC3_MAX_LOCAL_BYTES) { d.dropped[f] = true; return nullptr; }
            return l;
        };
        const int forced = forced_form();
        const Entry::Loaded* pick = nullptr;
        if (forced >= 0) {
            form = (forced == FORM_RING && !ring_ok) ? FORM_CLASSIC2 : forced;
            pick = usable(form);
            if (!pick && form != FORM_CLASSIC2) { form = FORM_CLASSIC2; pick = usable(form); }
        } else if (!all_done()) {
            form = FORM_CLASSIC2;          // the other candidates are still compiling
            pick = usable(form);
        } else {
            harvest(d);
            const bool dense_full = ring_ok;
            // (a pass whose load carries a fused remap pulls half its tile over NVLink: not a fair timing sample)
            if (dense_full && !pp.pd.remap_on && !pp.pd.remap_st && d.best_dense < 0 && (d.sig < 0 || d.sig == pp.pd.n_cta_bits)) {
                // measuring: the usable candidate with the fewest timings (finished or in flight) runs next
                int n_have[FORM_COUNT];
                for (int f = 0; f < FORM_COUNT; ++f) n_have[f] = (int)d.ms[f].size();
                for (auto& sm : d.pending) ++n_have[sm.form];
                int cand = -1;
                bool complete = true;
                for (int f = 0; f < FORM_COUNT; ++f) {
                    if (!usable(f)) continue;
                    if ((int)d.ms[f].size() < TUNE_SAMPLES) complete = false;
                    if (n_have[f] < TUNE_SAMPLES && (cand < 0 || n_have[f] < n_have[cand])) cand = f;
                }
                if (complete) {
                    float best = 0.f, best_c = 0.f;
                    for (int f = 0; f < FORM_COUNT; ++f) {
                        if (!usable(f)) continue;
                        float m = d.ms[f][0];
                        for (float v : d.ms[f]) m = v < m ? v : m;
                        if (d.best_dense < 0 || m < best) { d.best_dense = f; best = m; }
                        if (f != FORM_RING && (d.best_classic < 0 || m < best_c)) { d.best_classic = f; best_c = m; }
                    }
                } else if (cand >= 0) {
                    form = cand;
                    pick = usable(form);
                    d.sig = pp.pd.n_cta_bits;
                    if (pick && (cudaEventCreate(&ev0) != cudaSuccess || cudaEventCreate(&ev1) != cudaSuccess)) {
                        cudaGetLastError();
                        if (ev0) cudaEventDestroy(ev0);
                        ev0 = ev1 = nullptr;
                    }
                }
            }
            if (!pick) {
                // (the ring form is the wrong tool over NVLink: 55 against 28 ms per fused pass on 2 x B200)
                form = dense_full && !pp.pd.remap_on && !pp.pd.remap_st ? (d.best_dense >= 0 ? d.best_dense : FORM_CLASSIC2)
                                                                         : (d.best_classic >= 0 ? d.best_classic : FORM_CLASSIC2);
                pick = usable(form);
                if (!pick && form != FORM_CLASSIC2) { form = FORM_CLASSIC2; pick = usable(form); }
            }
        }
        if (!pick) return false;
        fn = pick->fn;
    }
    const bool ring = form == FORM_RING;
    const unsigned ctas = ring ? sms : 1u << pp.pd.n_cta_bits;
    const unsigned threads = ring ? RING_GROUPS * NTHREADS : NTHREADS;
    const unsigned smem = ring ? (unsigned)RING_SMEM_BYTES : TILE_SLOTS * (unsigned)sizeof(cplx);
    void* params[] = {(void*)&amp, (void*)&pp};
    if (ev0) cudaEventRecord(ev0, stream);
    const CUresult rc = r.api.LaunchKernel(fn, ctas, 1, 1, threads, 1, 1, smem, (CUstream)stream, params, nullptr);
    if (ev0) {
        cudaEventRecord(ev1, stream);
        std::lock_guard<std::mutex> lk(r.mu);
        Tuner::Sample sm; sm.t0 = ev0; sm.t1 = ev1; sm.form = form;
        tuner->dev[device].pending.push_back(sm);
    }
    if (rc != CUDA_SUCCESS) {
        if (err) *err = "jit launch: " + r.api.cu_error(rc);
        std::lock_guard<std::mutex> lk(r.mu);
        tuner->dev[device].dropped[form] = true;
        return false;
    }
    {
        std::lock_guard<std::mutex> lk(r.mu);
        ++r.stats.launches[form];
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
    s.tuning = 0;
    for (int f = 0; f < FORM_COUNT; ++f) s.chosen[f] = 0;
    for (auto& kv : r.tuners)
        for (auto& dv : kv.second->dev) {
            if (dv.second.sig >= 0 && dv.second.best_dense < 0) ++s.tuning;
            if (dv.second.best_dense >= 0) ++s.chosen[dv.second.best_dense];
        }
    return s;
}

}  // namespace dvd
