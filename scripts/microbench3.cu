// microbench3.cu -- round 2: what limits the high-stride tile I/O of the fused pass (30 qubits, tile = qubits 0..2 +
// 21..29, i.e. 512 segments of 128 B, 32 MiB apart: every segment of a tile lies in a different 2 MiB page).
//   A  one tile per CTA (the k_tile_pass access pattern), CTAs/SM limited to 1..4 by shared memory
//   B  one CTA handles the two tiles adjacent in bit 3 one after the other (same pages, 256 B per page in total)
//   C  one CTA (512 threads) handles both at once (= 256 B segments)
//   D  persistent CTAs: tiles c, c+G, ... (strided) against a contiguous range of tile indices per CTA
//   E  cache hints: plain ld/st, ld.cs/st.cs (what the pass kernel uses), ld.nc.L1::no_allocate / st.cg
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/microbench3 scripts/microbench3.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int TB = 12;
template <int HINT> __device__ __forceinline__ double2 ld(const double2* p) {
    if (HINT == 1) return __ldcs(p);
    if (HINT == 2) { double2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p)); return v; }
    return *p;
}
template <int HINT> __device__ __forceinline__ void st(double2* p, double2 v) {
    if (HINT == 1) __stcs(p, v);
    else if (HINT == 2) __stcg(p, v);
    else *p = v;
}
__device__ __forceinline__ uint64_t tile_base(uint64_t b, int seg_bits, int hi_start) {
    const int hi_bits = TB - seg_bits;
    const uint64_t lowmask = (1ull << (hi_start - seg_bits)) - 1;
    return ((b & lowmask) << seg_bits) | ((b >> (hi_start - seg_bits)) << (hi_start + hi_bits));
}
__device__ __forceinline__ uint64_t tile_off(int idx, int seg_bits, int hi_start) {
    return (uint64_t)(idx & ((1 << seg_bits) - 1)) | ((uint64_t)(idx >> seg_bits) << hi_start);
}
template <int HINT>
__device__ __forceinline__ void rmw_tile(double2* a, uint64_t base, int seg_bits, int hi_start, double c) {
    double2 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = ld<HINT>(a + base + tile_off(k * 256 + (threadIdx.x & 255), seg_bits, hi_start));
#pragma unroll
    for (int k = 0; k < 16; ++k) { v[k].x *= c; v[k].y *= c; st<HINT>(a + base + tile_off(k * 256 + (threadIdx.x & 255), seg_bits, hi_start), v[k]); }
}
// A / E: one tile per CTA
template <int HINT>
__global__ void __launch_bounds__(256) k_one(double2* a, int seg_bits, int hi_start, double c) {
    extern __shared__ unsigned char pad[];
    rmw_tile<HINT>(a, tile_base(blockIdx.x, seg_bits, hi_start), seg_bits, hi_start, c);
    if (c == 123.456) pad[threadIdx.x] = 0;
}
// B: two adjacent tiles, one after the other
__global__ void __launch_bounds__(256) k_pair_seq(double2* a, int seg_bits, int hi_start, double c) {
    extern __shared__ unsigned char pad[];
    rmw_tile<1>(a, tile_base(2ull * blockIdx.x, seg_bits, hi_start), seg_bits, hi_start, c);
    rmw_tile<1>(a, tile_base(2ull * blockIdx.x + 1, seg_bits, hi_start), seg_bits, hi_start, c);
    if (c == 123.456) pad[threadIdx.x] = 0;
}
// C: two adjacent tiles at once, 512 threads (group g takes tile 2b + g)
__global__ void __launch_bounds__(512) k_pair_joint(double2* a, int seg_bits, int hi_start, double c) {
    extern __shared__ unsigned char pad[];
    rmw_tile<1>(a, tile_base(2ull * blockIdx.x + (threadIdx.x >> 8), seg_bits, hi_start), seg_bits, hi_start, c);
    if (c == 123.456) pad[threadIdx.x] = 0;
}
// D: persistent; order 0 = strided (t = cta + k * grid), 1 = contiguous range per CTA
__global__ void __launch_bounds__(256) k_persist(double2* a, int seg_bits, int hi_start, uint64_t n_tiles, int order, double c) {
    extern __shared__ unsigned char pad[];
    if (order == 0) {
        for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) rmw_tile<1>(a, tile_base(t, seg_bits, hi_start), seg_bits, hi_start, c);
    } else {
        const uint64_t per = (n_tiles + gridDim.x - 1) / gridDim.x;
        const uint64_t t0 = per * blockIdx.x, t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
        for (uint64_t t = t0; t < t1; ++t) rmw_tile<1>(a, tile_base(t, seg_bits, hi_start), seg_bits, hi_start, c);
    }
    if (c == 123.456) pad[threadIdx.x] = 0;
}
// G: one tile per CTA + L2 prefetch of the tile `dist` CTAs ahead (each 128-byte line once: lanes with (tid & 7) == 0)
template <int MODE>   // 0: prefetch.global.L2, 1: cp.async.bulk.prefetch.L2 (128 B), 2: prefetch.global.L2::evict_last
__global__ void __launch_bounds__(256) k_one_pf(double2* a, int seg_bits, int hi_start, uint64_t n_tiles, int dist, double c) {
    extern __shared__ unsigned char pad[];
    const uint64_t tn = (uint64_t)blockIdx.x + dist;
    if (tn < n_tiles && (threadIdx.x & 7) == 0) {
        const uint64_t base = tile_base(tn, seg_bits, hi_start);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const double2* p = a + base + tile_off(k * 256 + threadIdx.x, seg_bits, hi_start);
            if (MODE == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            else if (MODE == 1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], 128;" ::"l"(p));
            else asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
        }
    }
    rmw_tile<1>(a, tile_base(blockIdx.x, seg_bits, hi_start), seg_bits, hi_start, c);
    if (c == 123.456) pad[threadIdx.x] = 0;
}
// H: one tile per CTA + "prefetch by load": cp.async (LDGSTS) of 4 bytes per SECTORS-th 32-byte sector of every line of
// the tile `dist` CTAs ahead into a scratch word of shared memory (no register, never waited for)
template <int PER_LINE>   // 1: one word per 128-byte line, 2: one per 64 bytes, 4: one per 32-byte sector
__global__ void __launch_bounds__(256) k_one_pfld(double2* a, int seg_bits, int hi_start, uint64_t n_tiles, int dist, double c) {
    extern __shared__ unsigned char pad[];
    __shared__ unsigned scratch[256];
    const uint64_t tn = (uint64_t)blockIdx.x + dist;
    if (tn < n_tiles) {
        const uint64_t base = tile_base(tn, seg_bits, hi_start);
        const unsigned d = (unsigned)__cvta_generic_to_shared(&scratch[threadIdx.x]);
        // 512 lines x PER_LINE words per tile, 256 threads: 2 * PER_LINE copies per thread
#pragma unroll
        for (int k = 0; k < 2 * PER_LINE; ++k) {
            const int w = k * 256 + threadIdx.x;               // word index in [0, 512 * PER_LINE)
            const int line = w / PER_LINE, sub = w % PER_LINE;
            const char* p = reinterpret_cast<const char*>(a + base + tile_off(line * 8, seg_bits, hi_start)) + sub * (128 / PER_LINE);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(p) : "memory");
        }
    }
    rmw_tile<1>(a, tile_base(blockIdx.x, seg_bits, hi_start), seg_bits, hi_start, c);
    if (c == 123.456) pad[threadIdx.x] = (unsigned char)scratch[threadIdx.x];
}
// F: read-only and write-only halves of the traffic (where is the loss: reads or writes?)
__global__ void __launch_bounds__(256) k_read_only(const double2* a, int seg_bits, int hi_start, double* sink) {
    const uint64_t base = tile_base(blockIdx.x, seg_bits, hi_start);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { const double2 v = __ldcs(a + base + tile_off(k * 256 + threadIdx.x, seg_bits, hi_start)); acc += v.x + v.y; }
    if (acc == 123.456) *sink = acc;
}
__global__ void __launch_bounds__(256) k_write_only(double2* a, int seg_bits, int hi_start, double c) {
    const uint64_t base = tile_base(blockIdx.x, seg_bits, hi_start);
#pragma unroll
    for (int k = 0; k < 16; ++k) __stcs(a + base + tile_off(k * 256 + threadIdx.x, seg_bits, hi_start), make_double2(c, c));
}

template <typename F>
static float time_it(F&& launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 30;
    const uint64_t N = 1ull << n;
    const double bytes = 32.0 * (double)N;
    double2* a = nullptr;
    CK(cudaMalloc(&a, N * sizeof(double2)));
    CK(cudaMemset(a, 0, N * sizeof(double2)));
    double* sink = nullptr;
    CK(cudaMalloc(&sink, 8));
    const uint64_t n_tiles = N >> TB;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto report = [&](const char* name, float ms, double b) { printf("%-64s %7.3f ms  %6.0f GB/s\n", name, ms, b / ms / 1e6); fflush(stdout); };
    struct Shape { const char* name; int seg_bits, hi_start; };
    const Shape shapes[] = {{"seg128B hi[n-9..]", 3, n - 9}, {"seg128B mid[12..20]", 3, 12}, {"seg256B hi[n-8..]", 4, n - 8},
                            {"seg128B hi[n-10..n-2]", 3, n - 10}, {"contig[0..11]", 12, 12}};
    CK(cudaFuncSetAttribute(k_one<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CK(cudaFuncSetAttribute(k_one<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CK(cudaFuncSetAttribute(k_one<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CK(cudaFuncSetAttribute(k_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    char name[160];
    for (const Shape& sh : shapes) {
        if (sh.seg_bits == 12) {   // contiguous tiles: base = b << 12
            report("A one tile/CTA contig ld.cs/st.cs", time_it([&] { k_one<1><<<(unsigned)n_tiles, 256>>>(a, 12, 12, 1.0); }), bytes);
            continue;
        }
        for (int hint = 0; hint < 3; ++hint) {
            snprintf(name, sizeof name, "A %s hint=%s", sh.name, hint == 0 ? "plain" : hint == 1 ? "cs" : "nc.noalloc/cg");
            float ms = hint == 0 ? time_it([&] { k_one<0><<<(unsigned)n_tiles, 256>>>(a, sh.seg_bits, sh.hi_start, 1.0); })
                     : hint == 1 ? time_it([&] { k_one<1><<<(unsigned)n_tiles, 256>>>(a, sh.seg_bits, sh.hi_start, 1.0); })
                                 : time_it([&] { k_one<2><<<(unsigned)n_tiles, 256>>>(a, sh.seg_bits, sh.hi_start, 1.0); });
            report(name, ms, bytes);
        }
        if (sh.seg_bits != 3 || sh.hi_start != n - 9) continue;
        for (int per_sm = 1; per_sm <= 6; ++per_sm) {
            const size_t smem = per_sm >= 5 ? 0 : (size_t)(220 * 1024 / per_sm) - 2048;
            snprintf(name, sizeof name, "A %s cs, CTAs/SM <= %d (smem %zu KB)", sh.name, per_sm >= 5 ? 8 : per_sm, smem / 1024);
            report(name, time_it([&] { k_one<1><<<(unsigned)n_tiles, 256, smem>>>(a, sh.seg_bits, sh.hi_start, 1.0); }), bytes);
        }
        CK(cudaFuncSetAttribute(k_one_pf<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        CK(cudaFuncSetAttribute(k_one_pf<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        CK(cudaFuncSetAttribute(k_one_pf<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        CK(cudaFuncSetAttribute(k_one_pfld<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 230000));
        CK(cudaFuncSetAttribute(k_one_pfld<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 230000));
        CK(cudaFuncSetAttribute(k_one_pfld<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 230000));
        for (size_t gran : {(size_t)0, (size_t)128, (size_t)32}) {
            if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set L2 fetch granularity %zu: %s\n", gran, cudaGetErrorString(e)); }
            size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity now %zu\n", got);
            for (int per_sm = 2; per_sm <= 3; ++per_sm)
                for (int pl : {1, 2, 4})
                    for (int dist : {296, 592, 1184}) {
                        const size_t smem = (size_t)(216 * 1024 / per_sm) - 2048;
                        snprintf(name, sizeof name, "H prefetch by LDGSTS.32 x%d/line dist %4d, CTAs/SM <= %d", pl, dist, per_sm);
                        float ms = pl == 1 ? time_it([&] { k_one_pfld<1><<<(unsigned)n_tiles, 256, smem>>>(a, sh.seg_bits, sh.hi_start, n_tiles, dist, 1.0); })
                                 : pl == 2 ? time_it([&] { k_one_pfld<2><<<(unsigned)n_tiles, 256, smem>>>(a, sh.seg_bits, sh.hi_start, n_tiles, dist, 1.0); })
                                           : time_it([&] { k_one_pfld<4><<<(unsigned)n_tiles, 256, smem>>>(a, sh.seg_bits, sh.hi_start, n_tiles, dist, 1.0); });
                        report(name, ms, bytes);
                    }
            snprintf(name, sizeof name, "A (no prefetch) CTAs/SM <= 2, granularity %zu", got);
            report(name, time_it([&] { k_one<1><<<(unsigned)n_tiles, 256, (size_t)(216 * 1024 / 2) - 2048>>>(a, sh.seg_bits, sh.hi_start, 1.0); }), bytes);
        }
        report("B pair of adjacent tiles, sequential", time_it([&] { k_pair_seq<<<(unsigned)(n_tiles / 2), 256>>>(a, sh.seg_bits, sh.hi_start, 1.0); }), bytes);
        report("C pair of adjacent tiles, joint (512 thr)", time_it([&] { k_pair_joint<<<(unsigned)(n_tiles / 2), 512>>>(a, sh.seg_bits, sh.hi_start, 1.0); }), bytes);
        for (int order = 0; order < 2; ++order)
            for (int per_sm = 2; per_sm <= 8; per_sm *= 2) {
                snprintf(name, sizeof name, "D persistent %s, %d CTAs/SM", order ? "contiguous ranges" : "strided", per_sm);
                report(name, time_it([&] { k_persist<<<sms * per_sm, 256>>>(a, sh.seg_bits, sh.hi_start, n_tiles, order, 1.0); }), bytes);
            }
        report("F read only", time_it([&] { k_read_only<<<(unsigned)n_tiles, 256>>>(a, sh.seg_bits, sh.hi_start, sink); }), bytes / 2);
        report("F write only", time_it([&] { k_write_only<<<(unsigned)n_tiles, 256>>>(a, sh.seg_bits, sh.hi_start, 0.0); }), bytes / 2);
    }
    report("F read only contig", time_it([&] { k_read_only<<<(unsigned)n_tiles, 256>>>(a, 12, 12, sink); }), bytes / 2);
    report("F write only contig", time_it([&] { k_write_only<<<(unsigned)n_tiles, 256>>>(a, 12, 12, 0.0); }), bytes / 2);
    printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
