// microbench2.cu -- HBM ceilings for the access patterns of the fused gate pass on B200.
//   1. in-place contiguous scale (LDG.128/STG.128): the read+write ceiling
//   2. the real k_tile_pass with zero ops, for several tile shapes (segment size / stride)
//   3. a persistent TMA (cp.async.bulk, 1-D) ring: gather tile -> smem -> regs -> smem -> scatter
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench2 \
//        scripts/microbench2.cu damavand_b200/csrc/kernels.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../damavand_b200/csrc/kernels.h"

using namespace dvd;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_scale(double2* __restrict__ a, uint64_t n, double c) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; i < n; i += stride) {
        double2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = a[i + k * 256];
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k].x *= c; v[k].y *= c; a[i + k * 256] = v[k]; }
    }
}

// one CTA per tile, no persistence: thread owns 16 amps (like k_tile_pass), UNROLLED loads then stores
template <int NT, int PER>
__global__ void __launch_bounds__(NT) k_tile_rmw(double2* __restrict__ a, int seg_bits, int hi_start, double c) {
    extern __shared__ unsigned char pad_smem[];
    constexpr int TB = 12;
    const int hi_bits = TB - seg_bits;
    // cta base: insert zeros at [0,seg_bits) and [hi_start, hi_start+hi_bits)
    uint64_t b = blockIdx.x;
    uint64_t lowmask = (1ull << (hi_start - seg_bits)) - 1;
    uint64_t base = ((b & lowmask) << seg_bits) | ((b >> (hi_start - seg_bits)) << (hi_start + hi_bits));
    double2 v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int idx = k * NT + threadIdx.x;          // tile index: low seg_bits contiguous
        const uint64_t off = (uint64_t)(idx & ((1 << seg_bits) - 1)) | ((uint64_t)(idx >> seg_bits) << hi_start);
        v[k] = a[base + off];
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int idx = k * NT + threadIdx.x;
        const uint64_t off = (uint64_t)(idx & ((1 << seg_bits) - 1)) | ((uint64_t)(idx >> seg_bits) << hi_start);
        v[k].x *= c; v[k].y *= c;
        a[base + off] = v[k];
    }
    if (c == 123.456) pad_smem[threadIdx.x] = 0;
}

// ---- TMA bulk ring ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int STAGES, int TB>
__global__ void __launch_bounds__(256, 1) k_bulk_ring(double2* __restrict__ a, int seg_bits, int hi_start, uint64_t n_tiles, double c, int touch) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int TILE = 1 << TB;
    double2* buf = reinterpret_cast<double2*>(smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * TILE * 16);
    const int tid = threadIdx.x;
    const int hi_bits = TB - seg_bits;
    const int nseg = 1 << hi_bits;
    const uint32_t seg_bytes = 16u << seg_bits;
    const uint64_t lowmask = (1ull << (hi_start - seg_bits)) - 1;
    auto tile_base = [&](uint64_t b) {
        return ((b & lowmask) << seg_bits) | ((b >> (hi_start - seg_bits)) << (hi_start + hi_bits));
    };
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue_load = [&](uint64_t t, int slot) {
        const uint64_t base = tile_base(t);
        if (tid == 0) mbar_expect_tx(&full[slot], TILE * 16);
        for (int s = tid; s < nseg; s += 256)
            bulk_g2s(buf + (size_t)slot * TILE + ((size_t)s << seg_bits), a + base + ((uint64_t)s << hi_start), seg_bytes, &full[slot]);
    };
    // my tiles: blockIdx.x, +grid, ...
    const uint64_t first = blockIdx.x, step = gridDim.x;
    uint64_t n_my = first < n_tiles ? (n_tiles - first + step - 1) / step : 0;
    for (int p = 0; p < STAGES - 1 && (uint64_t)p < n_my; ++p) issue_load(first + p * step, p);
    for (uint64_t k = 0; k < n_my; ++k) {
        const int slot = (int)(k % STAGES);
        mbar_wait(&full[slot], (uint32_t)((k / STAGES) & 1));
        double2* b = buf + (size_t)slot * TILE;
        if (touch) {
#pragma unroll
            for (int j = 0; j < TILE / 256; ++j) {
                double2 v = b[j * 256 + tid];
                v.x *= c; v.y *= c;
                b[j * 256 + tid] = v;
            }
            fence_async();
        }
        __syncthreads();
        const uint64_t base = tile_base(first + k * step);
        for (int s = tid; s < nseg; s += 256)
            bulk_s2g(a + base + ((uint64_t)s << hi_start), b + ((size_t)s << seg_bits), seg_bytes);
        bulk_commit();
        // refill the slot used by tile k-1 (its stores were committed one iteration ago)
        if (k + STAGES - 1 < n_my) {
            bulk_wait_read<1>();
            __syncthreads();
            issue_load(first + (k + STAGES - 1) * step, (int)((k + STAGES - 1) % STAGES));
        }
    }
    bulk_wait_read<0>();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- fp64 / smem ceilings ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_lds(double* out, int iters) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = make_double2(i, -i);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) { double2 v = sm[(idx + k * 256) & 4095]; acc.x += v.x; acc.y += v.y; }
        idx = (idx + 1) & 4095;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

template <class F>
static float time_ms(F f, int reps = 3) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 30;
    const uint64_t N = 1ull << n;
    const double bytes = 32.0 * N;
    double2* a; CK(cudaMalloc(&a, N * 16));
    CK(cudaMemset(a, 0, N * 16));
    CK(kernels_init());
    printf("n=%d state %.1f GiB, bytes per pass %.1f GiB\n", n, N * 16.0 / (1 << 30), bytes / (1 << 30));

    for (int g : {148 * 8, 148 * 16, 148 * 32}) {
        float ms = time_ms([&] { k_scale<<<g, 256>>>(a, N, 1.0); });
        printf("scale in-place grid=%5d: %.3f ms  %.0f GB/s\n", g, ms, bytes / ms / 1e6);
    }
    {
        float ms = time_ms([&] { cudaMemcpyAsync(a, a + N / 2, N * 8, cudaMemcpyDeviceToDevice); });
        printf("cudaMemcpy D2D half->half: %.3f ms  %.0f GB/s (r+w)\n", ms, (double)N * 16 / ms / 1e6);
    }
    // real tile pass, zero ops
    struct Shape { const char* name; std::vector<int> q; };
    std::vector<Shape> shapes;
    auto mk = [&](const char* nm, int seg, int hi) { Shape s; s.name = nm; for (int i = 0; i < seg; ++i) s.q.push_back(i); for (int i = 0; i < 12 - seg; ++i) s.q.push_back(hi + i); shapes.push_back(s); };
    mk("contig[0..11]", 12, 12);
    mk("seg128B+hi[n-9..]", 3, n - 9);
    mk("seg128B+mid[12..20]", 3, 12);
    mk("seg64B+hi[n-10..]", 2, n - 10);
    mk("seg256B+hi[n-8..]", 4, n - 8);
    mk("seg512B+hi[n-7..]", 5, n - 7);
    for (auto& s : shapes) {
        PassDesc pd; memset(&pd, 0, sizeof pd);
        pd.n_local = n; pd.n_ops = 0;
        for (int p = 0; p < 12; ++p) pd.tile_q[p] = pd.sorted_q[p] = s.q[p];
        std::sort(pd.sorted_q, pd.sorted_q + 12);
        float ms = time_ms([&] { launch_tile_pass(reinterpret_cast<cplx*>(a), nullptr, pd, 0); });
        printf("k_tile_pass 0 ops %-22s: %.3f ms  %.0f GB/s\n", s.name, ms, bytes / ms / 1e6);
    }
    // standalone rmw, occupancy sweep via dynamic smem padding
    for (int seg : {3, 2}) {
        for (int smem_kb : {0, 48, 64, 100}) {
            const int hi = n - (12 - seg);
            CK(cudaFuncSetAttribute(k_tile_rmw<256, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            float ms = time_ms([&] { k_tile_rmw<256, 16><<<(unsigned)(N >> 12), 256, smem_kb * 1024>>>(a, seg, hi, 1.0); });
            printf("tile_rmw<256,16> seg=%dB smem=%3dKB: %.3f ms  %.0f GB/s\n", 16 << seg, smem_kb, ms, bytes / ms / 1e6);
        }
        {
            const int hi = n - (12 - seg);
            float ms = time_ms([&] { k_tile_rmw<512, 8><<<(unsigned)(N >> 12), 512>>>(a, seg, hi, 1.0); });
            printf("tile_rmw<512,8>  seg=%dB         : %.3f ms  %.0f GB/s\n", 16 << seg, ms, bytes / ms / 1e6);
        }
    }
    // TMA bulk ring
    {
        auto run = [&](auto kern, int stages, int tb, int seg, int touch, const char* nm) {
            const size_t smem = (size_t)stages * (16u << tb) + 64;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int hi = n - (tb - seg);
            float ms = time_ms([&] { kern<<<148, 256, smem>>>(a, seg, hi, N >> tb, 1.0, touch); });
            printf("bulk_ring %-10s seg=%4dB touch=%d: %.3f ms  %.0f GB/s\n", nm, 16 << seg, touch, ms, bytes / ms / 1e6);
        };
        for (int seg : {3, 4, 6}) {
            run(k_bulk_ring<3, 12>, 3, 12, seg, 0, "S3 T12");
            run(k_bulk_ring<3, 12>, 3, 12, seg, 1, "S3 T12");
            run(k_bulk_ring<2, 12>, 2, 12, seg, 1, "S2 T12");
            run(k_bulk_ring<3, 11>, 3, 11, seg, 1, "S3 T11");
        }
    }
    // fp64 + smem
    double* out; CK(cudaMalloc(&out, 148 * 16 * 256 * 8));
    for (int cps : {1, 2, 4}) {
        int grid = 148 * cps, iters = 20000;
        float ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 2);
        double fl = 16.0 * iters * 256.0 * grid;
        printf("DFMA ctas/SM=%d: %.2f TFLOP/s  %.1f DFMA/clk/SM @1.965GHz\n", cps, 2 * fl / ms / 1e9, fl / (ms * 1e-3) / 148 / 1.965e9);
    }
    CK(cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int cps : {1, 2, 3}) {
        int grid = 148 * cps, iters = 2000;
        float ms = time_ms([&] { k_lds<<<grid, 256, 65536>>>(out, iters); }, 2);
        double by = 16.0 * 16 * iters * 256.0 * grid;
        printf("LDS.128 ctas/SM=%d: %.1f B/clk/SM\n", cps, by / (ms * 1e-3) / 148 / 1.965e9);
    }
    printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
