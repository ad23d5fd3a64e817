// l2_probe_bench — what a B200 delivers for the access pattern of the node-table probes: random, independent
// 32-byte records (one LDG.E.ENL2.256 each) out of a table that stays resident in L2 (or does not), and random
// single-byte reads from a shared-memory seed table.  Standalone; build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/l2_probe_bench profiles/tools/l2_probe_bench.cu
//   gpurun_out/l2_probe_bench > gpurun_out/l2_probe_bench.json
// Output: one JSON object per line {"test": ..., "table_mb": ..., "ilp": ..., "gprobes_s": ..., "gbs": ...}.
// The numbers give the second roofline of the scoring kernel (bound "l2_random", BASELINE.md / bench.py).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

struct Rec32 {
    uint32_t v[8];
};

__device__ __forceinline__ Rec32 load_record(const void* base, uint32_t slot) {
    Rec32 r;
    const char* p = static_cast<const char*>(base) + (size_t(slot) << 5);
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                   "=r"(r.v[7])
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x *= 0x9E3779B1u;
    x ^= x >> 15;
    x *= 0x85EBCA77u;
    x ^= x >> 13;
    return x;
}

// kIlp independent random record loads in flight per thread, `iters` rounds; kDependent chains the next index on
// the loaded data (latency-bound variant: one probe depends on the previous, as the 3->2->1 fallback chain does).
template <int kIlp, bool kDependent>
__global__ void __launch_bounds__(1024) k_probe(const void* table, uint32_t nslots, int iters, uint32_t* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    uint32_t idx[kIlp];
#pragma unroll
    for (int j = 0; j < kIlp; ++j) idx[j] = mix(tid * kIlp + j + 1);
    for (int i = 0; i < iters; ++i) {
        Rec32 r[kIlp];
#pragma unroll
        for (int j = 0; j < kIlp; ++j) r[j] = load_record(table, uint32_t((uint64_t(idx[j]) * nslots) >> 32));
#pragma unroll
        for (int j = 0; j < kIlp; ++j) {
            const uint32_t s = r[j].v[0] ^ r[j].v[3] ^ r[j].v[7];
            acc += s;
            idx[j] = mix(idx[j] + (kDependent ? s : 0u) + 0x632BE5ABu);
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// Variants that tell WHAT the limit is: kBytes per lane (8 / 16 / 32 / 128) and kShare lanes per 128-byte line
// (1 = every lane its own random line, 2 / 4 = neighbouring lanes read different records of the same line).
template <int kBytes, int kShare>
__global__ void __launch_bounds__(1024) k_probe_var(const void* table, uint32_t nlines, int iters, uint32_t* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    uint32_t idx = mix((tid / kShare) + 1);
    const uint32_t sub = (tid % kShare) * (128 / kShare);
    for (int i = 0; i < iters; ++i) {
        const uint32_t line = uint32_t((uint64_t(idx) * nlines) >> 32);
        const char* p = static_cast<const char*>(table) + (size_t(line) << 7) + sub;
        uint32_t s;
        if (kBytes == 8) {
            uint32_t a, b;
            asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
            s = a ^ b;
        } else if (kBytes == 16) {
            uint32_t a, b, c, d;
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
            s = a ^ b ^ c ^ d;
        } else if (kBytes == 32) {
            const Rec32 r = load_record(p, 0);
            s = r.v[0] ^ r.v[3] ^ r.v[7];
        } else {
            const Rec32 r0 = load_record(p, 0), r1 = load_record(p, 1), r2 = load_record(p, 2), r3 = load_record(p, 3);
            s = r0.v[0] ^ r1.v[3] ^ r2.v[7] ^ r3.v[1];
        }
        acc += s;
        idx = mix(idx + 0x632BE5ABu);
    }
    if (acc == 0x12345678u) *sink = acc;
}

// random byte reads from a 37 KB shared-memory table (the perfect-hash seeds): LDS.U8 with random bank pattern
__global__ void __launch_bounds__(1024) k_smem_seed(int iters, uint32_t* sink) {
    __shared__ uint8_t s_seed[37632];
    for (int i = threadIdx.x; i < 37632; i += blockDim.x) s_seed[i] = uint8_t(i * 7);
    __syncthreads();
    uint32_t x = mix(blockIdx.x * blockDim.x + threadIdx.x + 1), acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc += s_seed[(uint64_t(x) * 37632u) >> 32];
            x = x * 0x9E3779B1u + 0x7F4A7C15u;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// pure issue-rate reference: dependent-free integer work (IMAD + LOP3 mix), 8 chains per thread
__global__ void __launch_bounds__(1024) k_issue(int iters, uint32_t* sink) {
    uint32_t a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = a[j] * 0x9E3779B1u + 12345u;  // IMAD
            a[j] ^= a[j] >> 7;                   // SHF + LOP3
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc ^= a[j];
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
static float time_ms(F&& launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
    return ms / reps;
}

int main() {
    int dev = 0, n_sm = 0, clk_khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    uint32_t* sink;
    CK(cudaMalloc(&sink, 4));
    const size_t max_bytes = size_t(1024) << 20;
    void* table;
    CK(cudaMalloc(&table, max_bytes));
    CK(cudaMemset(table, 0x5A, max_bytes));
    const int threads = 1024, blocks_per_sm = 2;
    const int grid = n_sm * blocks_per_sm;
    const double nthreads = double(grid) * threads;
    printf("{\"test\": \"device\", \"sms\": %d, \"clock_mhz\": %.0f}\n", n_sm, clk_khz / 1000.0);

    {   // what the limit is made of: bytes per lane and lanes per line, L2-resident 23 MB table and an L1-sized 64 KB one
        const int iters = 64;
        const double sizes_mb[] = {0.0625, 23};
        for (double mb : sizes_mb) {
            const uint32_t nlines = uint32_t(mb * 1048576.0 / 128.0);
            struct V { const char* name; float ms; } v[7];
            v[0] = {"8B_share1", time_ms([&] { k_probe_var<8, 1><<<grid, threads>>>(table, nlines, iters, sink); }, 5)};
            v[1] = {"16B_share1", time_ms([&] { k_probe_var<16, 1><<<grid, threads>>>(table, nlines, iters, sink); }, 5)};
            v[2] = {"32B_share1", time_ms([&] { k_probe_var<32, 1><<<grid, threads>>>(table, nlines, iters, sink); }, 5)};
            v[3] = {"128B_share1", time_ms([&] { k_probe_var<128, 1><<<grid, threads>>>(table, nlines, iters, sink); }, 5)};
            v[4] = {"32B_share2", time_ms([&] { k_probe_var<32, 2><<<grid, threads>>>(table, nlines, iters, sink); }, 5)};
            v[5] = {"32B_share4", time_ms([&] { k_probe_var<32, 4><<<grid, threads>>>(table, nlines, iters, sink); }, 5)};
            v[6] = {"8B_share4", time_ms([&] { k_probe_var<8, 4><<<grid, threads>>>(table, nlines, iters, sink); }, 5)};
            for (const V& x : v) {
                const double probes = nthreads * iters;
                printf("{\"test\": \"variant\", \"variant\": \"%s\", \"table_mb\": %.4f, \"ms\": %.4f, \"glane_loads_s\": %.2f, "
                       "\"per_clk_per_sm\": %.3f}\n",
                       x.name, mb, x.ms, probes / x.ms / 1e6, probes / x.ms / 1e3 / n_sm / clk_khz);
            }
            fflush(stdout);
        }
    }
    const double table_mb[] = {0.0625, 0.5, 2, 8, 23, 48, 96, 256, 1024};
    for (double mb : table_mb) {
        const uint32_t nslots = uint32_t(mb * 1048576.0 / 32.0);
        const int iters = 64;
        struct V { const char* name; int ilp; bool dep; float ms; } v[] = {
            {"ilp1", 1, false, 0}, {"ilp2", 2, false, 0}, {"ilp4", 4, false, 0}, {"dep1", 1, true, 0}, {"dep2", 2, true, 0}};
        v[0].ms = time_ms([&] { k_probe<1, false><<<grid, threads>>>(table, nslots, iters, sink); }, 5);
        v[1].ms = time_ms([&] { k_probe<2, false><<<grid, threads>>>(table, nslots, iters, sink); }, 5);
        v[2].ms = time_ms([&] { k_probe<4, false><<<grid, threads>>>(table, nslots, iters, sink); }, 5);
        v[3].ms = time_ms([&] { k_probe<1, true><<<grid, threads>>>(table, nslots, iters, sink); }, 5);
        v[4].ms = time_ms([&] { k_probe<2, true><<<grid, threads>>>(table, nslots, iters, sink); }, 5);
        for (const V& x : v) {
            const double probes = nthreads * iters * x.ilp;
            printf("{\"test\": \"record32\", \"variant\": \"%s\", \"table_mb\": %.0f, \"ms\": %.4f, \"gprobes_s\": %.2f, "
                   "\"gbs\": %.1f, \"per_clk_per_sm\": %.3f}\n",
                   x.name, mb, x.ms, probes / x.ms / 1e6, probes * 32.0 / x.ms / 1e6, probes / x.ms / 1e3 / n_sm / clk_khz);
        }
        fflush(stdout);
    }
    {
        const int iters = 256;
        const float ms = time_ms([&] { k_smem_seed<<<n_sm, 1024>>>(iters, sink); }, 5);
        const double reads = double(n_sm) * 1024 * iters * 4;
        printf("{\"test\": \"smem_seed_u8\", \"ms\": %.4f, \"greads_s\": %.2f, \"reads_per_clk_per_sm\": %.2f}\n", ms,
               reads / ms / 1e6, reads / ms / 1e3 / n_sm / double(clk_khz));
    }
    {
        const int iters = 2048;
        const float ms = time_ms([&] { k_issue<<<grid, threads>>>(iters, sink); }, 5);
        const double warp_instr = nthreads / 32.0 * iters * 8 * 3;
        printf("{\"test\": \"issue\", \"ms\": %.4f, \"gwarp_instr_s\": %.1f, \"per_clk_per_sm\": %.2f}\n", ms,
               warp_instr / ms / 1e6, warp_instr / ms / 1e3 / n_sm / double(clk_khz));
    }
    CK(cudaFree(table));
    CK(cudaFree(sink));
    return 0;
}
