// Microbenchmark: what does one random 16-byte table access cost on B200 (DRAM bytes, accesses/s)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o random_access random_access.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix64(u64 x) { x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31; return x; }

template <int MODE>
__global__ void __launch_bounds__(256) k_access(u64* table, u64 slots, u64 n, u64 seed, u64* out) {
    u64 acc = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 s = __umul64hi(mix64(i + seed), slots);
        u64* p = table + 2 * s;
        if (MODE == 0) { ulonglong2 v = *reinterpret_cast<ulonglong2*>(p); acc ^= v.x ^ v.y; }                       // plain ld (L1)
        if (MODE == 1) { u64 a, b; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p)); acc ^= a ^ b; }
        if (MODE == 2) { u64 a, b; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); acc ^= a ^ b; }
        if (MODE == 3) { atomicAdd(p + 1, 1ull); }                                                                    // RED only
        if (MODE == 4) { u64 a, b; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); acc ^= a; if (a != 12345) atomicAdd(p + 1, 1ull); }  // ld then RED
        if (MODE == 5) { u64 a, b; asm volatile("ld.global.cg.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p)); acc ^= a ^ b; }
        if (MODE == 6) { u64 a, b; asm volatile("ld.global.cg.L2::128B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p)); acc ^= a ^ b; }
        if (MODE == 7) { u64 old = atomicAdd(p + 1, 1ull); acc ^= old; }                                               // ATOM with return
        if (MODE == 8) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
        if (MODE == 9) { u64 a, b; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p)); acc ^= a; p[1] = b + 1; }  // ld + plain store (non-atomic RMW)
    }
    if (acc == 0x1234567) out[0] = acc;
}

int main(int argc, char** argv) {
    const u64 slots = argc > 1 ? strtoull(argv[1], 0, 10) : 160000000ull;  // 16 B each -> 2.56 GB
    const u64 n = argc > 2 ? strtoull(argv[2], 0, 10) : 64000000ull;
    const int gran = argc > 3 ? atoi(argv[3]) : 0;
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set gran %d -> %s\n", gran, cudaGetErrorString(e)); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit = %zu\n", g);
    u64 *table, *out;
    cudaMalloc(&table, slots * 16); cudaMalloc(&out, 8);
    cudaMemset(table, 0, slots * 16);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char* names[] = {"ld plain", "ld.cg", "ld.relaxed.gpu", "RED.add", "ld.relaxed+RED", "ld.cg.L2::64B", "ld.cg.L2::128B", "ATOM.add ret", "prefetch.L2", "ld+st"};
    for (int mode = 0; mode < 10; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            const int grid = 148 * 8;
            switch (mode) {
                case 0: k_access<0><<<grid, 256>>>(table, slots, n, rep * 7919 + mode, out); break;
                case 1: k_access<1><<<grid, 256>>>(table, slots, n, rep * 7919 + mode, out); break;
                case 2: k_access<2><<<grid, 256>>>(table, slots, n, rep * 7919 + mode, out); break;
                case 3: k_access<3><<<grid, 256>>>(table, slots, n, rep * 7919 + mode, out); break;
                case 4: k_access<4><<<grid, 256>>>(table, slots, n, rep * 7919 + mode, out); break;
                case 5: k_access<5><<<grid, 256>>>(table, slots, n, rep * 7919 + mode, out); break;
                case 6: k_access<6><<<grid, 256>>>(table, slots, n, rep * 7919 + mode, out); break;
                case 7: k_access<7><<<grid, 256>>>(table, slots, n, rep * 7919 + mode, out); break;
                case 8: k_access<8><<<grid, 256>>>(table, slots, n, rep * 7919 + mode, out); break;
                case 9: k_access<9><<<grid, 256>>>(table, slots, n, rep * 7919 + mode, out); break;
            }
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep == 1) printf("mode %d %-16s %8.3f ms  %7.2f G access/s  (%s)\n", mode, names[mode], ms, n / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
