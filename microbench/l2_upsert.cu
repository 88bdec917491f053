// Microbenchmark: how fast can B200 apply hash-table upserts when all CTAs work on one L2-sized table region at a time?
// Mimics upsert_regions_kernel (genomix_b200/csrc/gx_split.cu): records are streamed (evict-first), sorted by region;
// each record does one 16-byte slot load plus atomics on that slot. Sweeps region size and the atomic mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_upsert l2_upsert.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__host__ __device__ __forceinline__ u64 mix64(u64 x) { x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31; return x; }

// records: slot index (global), generated region-sorted on the device
__global__ void gen_records(u64* rec, u64 n, u64 slots, u32 n_regions, u64 per_region, u64 distinct_per_region) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 r = i / per_region;
        const u64 lo = (u64)(((unsigned __int128)r * slots) / n_regions), hi = (u64)(((unsigned __int128)(r + 1) * slots) / n_regions);
        // draw one of `distinct_per_region` keys of the region, then place it pseudo-randomly in the region's slot range
        const u64 key = mix64(i * 0x9e3779b97f4a7c15ull + 1) % distinct_per_region;
        rec[i] = lo + mix64(key + r * 1000003ull) % (hi - lo);
    }
}

__device__ __forceinline__ void ld_v2(const u64* p, u64& a, u64& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ bool cas128(u64* addr, u64 c0, u64 c1, u64 n0, u64 n1, u64& o0, u64& o1) {
    asm volatile("{\n\t.reg .b128 c, n, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
                 "atom.relaxed.gpu.global.cas.b128 o, [%6], c, n;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(o0), "=l"(o1) : "l"(c0), "l"(c1), "l"(n0), "l"(n1), "l"(addr) : "memory");
    return o0 == c0 && o1 == c1;
}

// MODE 0: ld only   1: ld + RED.add   2: ld + RED.add + RED.or   3: real upsert (empty -> CAS64 + add + or; else add, or if new bits)
//      4: real upsert with a 128-bit CAS claiming key and value at once   5: RED.add only (no load)
//      6: ld + 32-bit RED.add   7: 32-bit RED.add only
template <int MODE, int PER>
__global__ void __launch_bounds__(512) k_upsert(u64* table, const u64* __restrict__ rec, u64 n, u64 slots, u32 n_regions, u64 per_region,
                                                int prefetch, u64* out) {
    const int tid = threadIdx.x;
    constexpr int TILE = 512 * PER;
    const u64 n_tiles = (n + TILE - 1) / TILE;
    u64 acc = 0;
    for (u64 t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const u64 first = t * TILE;
        if (prefetch) {
            const u64 r = first / per_region;
            if (r + 1 < n_regions) {
                const u64 lo = (u64)(((unsigned __int128)(r + 1) * slots) / n_regions), hi = (u64)(((unsigned __int128)(r + 2) * slots) / n_regions);
                const u64 lines = (hi - lo) * 16 / 128, tiles_per_region = per_region / TILE;
                const u64 part = t % tiles_per_region;
                const u64 l0 = lines * part / tiles_per_region, l1 = lines * (part + 1) / tiles_per_region;
                const char* base = (const char*)(table + 2 * lo);
                for (u64 l = l0 + tid; l < l1; l += 512) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + l * 128));
            }
        }
        u64 s[PER], k[PER], v[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const u64 idx = first + tid + (u64)i * 512;
            s[i] = idx < n ? __ldcs(rec + idx) : ~0ull;
            if (MODE != 5 && MODE != 7 && s[i] != ~0ull) ld_v2(table + 2 * s[i], k[i], v[i]);
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            if (s[i] == ~0ull) continue;
            u64* p = table + 2 * s[i];
            const u64 key = s[i] * 2 + 1, m = (mix64(s[i]) & 3) + 1;   // the slot's key; a few distinct masks per key
            if (MODE == 0) acc ^= k[i] ^ v[i];
            if (MODE == 1 || MODE == 5) { atomicAdd(p + 1, 1ull); if (MODE == 1) acc ^= k[i]; }
            if (MODE == 6 || MODE == 7) { atomicAdd(reinterpret_cast<u32*>(p + 1), 1u); if (MODE == 6) acc ^= k[i]; }
            if (MODE == 2) { atomicAdd(p + 1, 1ull); atomicOr(p + 1, m << 48); acc ^= k[i]; }
            if (MODE == 3) {
                u64 cur = k[i], seen = v[i];
                if (cur == 0) { cur = atomicCAS(p, 0ull, key); if (cur == 0) cur = key; seen = 0; }
                if (cur == key) { atomicAdd(p + 1, 1ull); if ((seen & (m << 48)) != (m << 48)) atomicOr(p + 1, m << 48); }
            }
            if (MODE == 4) {
                u64 cur = k[i], seen = v[i];
                bool done = false;
                if (cur == 0) {
                    u64 o0, o1;
                    if (cas128(p, 0ull, 0ull, key, 1ull | (m << 48), o0, o1)) done = true; else { cur = o0; seen = o1; }
                }
                if (!done && cur == key) { atomicAdd(p + 1, 1ull); if ((seen & (m << 48)) != (m << 48)) atomicOr(p + 1, m << 48); }
            }
        }
    }
    if (acc == 0x1234567) out[0] = acc;
}

int main(int argc, char** argv) {
    const u64 table_mb = argc > 1 ? strtoull(argv[1], 0, 10) : 2048;
    const u64 n = argc > 2 ? strtoull(argv[2], 0, 10) : 184000000ull;
    const double distinct_frac = argc > 3 ? atof(argv[3]) : 0.28;     // distinct keys / records
    const u64 slots = table_mb * 1024 * 1024 / 16;
    u64 *table, *rec, *out;
    cudaMalloc(&table, slots * 16); cudaMalloc(&rec, n * 8); cudaMalloc(&out, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char* names[] = {"ld only", "ld+RED.add", "ld+RED.add+RED.or", "upsert CAS64", "upsert CAS128", "RED.add only", "ld+RED.add.32", "RED.add.32 only"};
    printf("table %llu MB, %llu records, distinct fraction %.2f\n", table_mb, n, distinct_frac);
    const int region_mb[] = {4, 32, 128};
    for (int ri = 0; ri < 3; ++ri) {
        const u32 n_regions = (u32)(table_mb / region_mb[ri]);
        if (n_regions == 0) continue;
        const u64 per_region = (n / n_regions) / 8192 * 8192, nn = per_region * n_regions;
        gen_records<<<148 * 8, 256>>>(rec, nn, slots, n_regions, per_region, (u64)(per_region * distinct_frac) + 1);
        for (int mode = 0; mode < 8; ++mode) {
            for (int cfg = 0; cfg < 3; cfg += 2) {   // 0: 2 CTA/SM no prefetch, 1: 2 CTA/SM prefetch, 2: 4 CTA/SM... (1 CTA of 512 = 16 warps)
                const int grid = 148 * (cfg == 2 ? 4 : 2), prefetch = cfg == 1;
                float best = 1e9;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaMemset(table, 0, slots * 16);
                    cudaEventRecord(a);
                    switch (mode) {
                        case 0: k_upsert<0, 4><<<grid, 512>>>(table, rec, nn, slots, n_regions, per_region, prefetch, out); break;
                        case 1: k_upsert<1, 4><<<grid, 512>>>(table, rec, nn, slots, n_regions, per_region, prefetch, out); break;
                        case 2: k_upsert<2, 4><<<grid, 512>>>(table, rec, nn, slots, n_regions, per_region, prefetch, out); break;
                        case 3: k_upsert<3, 4><<<grid, 512>>>(table, rec, nn, slots, n_regions, per_region, prefetch, out); break;
                        case 4: k_upsert<4, 4><<<grid, 512>>>(table, rec, nn, slots, n_regions, per_region, prefetch, out); break;
                        case 5: k_upsert<5, 4><<<grid, 512>>>(table, rec, nn, slots, n_regions, per_region, prefetch, out); break;
                        case 6: k_upsert<6, 4><<<grid, 512>>>(table, rec, nn, slots, n_regions, per_region, prefetch, out); break;
                        case 7: k_upsert<7, 4><<<grid, 512>>>(table, rec, nn, slots, n_regions, per_region, prefetch, out); break;
                    }
                    cudaEventRecord(b); cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b);
                    if (ms < best) best = ms;
                }
                printf("region %4d MB x%5u  mode %d %-18s cfg %d  %7.3f ms  %6.1f G/s  (%s)\n", region_mb[ri], n_regions, mode, names[mode], cfg,
                       best, nn / best / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
