// Microbenchmark (developer tool, GPU box): what does a 64-byte record fetch cost per lane when every lane of a warp reads a
// different record (random, L2-resident 96 MB array), as the trace kernel's node fetch does?
//   mode 0  each lane: two 256-bit loads of ITS record                       (2 instructions x 32 distinct 128-B lines)
//   mode 1  lane pairs: instruction 1 = both halves of the even lane's record, instruction 2 = of the odd lane's
//           (2 instructions x 16 distinct lines); each lane keeps 32 B of either record
//   mode 2  each lane: four 128-bit loads of its record                      (4 x 32 lines)
//   mode 3  mode 0 + a prefetch.global.L1 of the next record's two sectors
// The next record index depends on the loaded data (pointer chase), 7 CTAs x 128 threads per SM like k_wf_trace.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pair_fetch tools/micro/pair_fetch.cu && ./pair_fetch
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
    float x0, x1, x2, x3, x4, x5, x6, x7;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3), "=f"(x4), "=f"(x5), "=f"(x6), "=f"(x7) : "l"(p));
    a = make_float4(x0, x1, x2, x3); b = make_float4(x4, x5, x6, x7);
}
struct Rec { uint32_t next[16]; }; // 64 B; every word holds a pseudo-random successor
__device__ __forceinline__ uint32_t pick(uint32_t r, uint32_t n, uint32_t hot, uint32_t hot_eighths) { return ((r >> 28) & 7u) < hot_eighths ? (r % hot) : (r % n); }
template <int MODE>
__global__ void __launch_bounds__(128, 7) k(const Rec* __restrict__ recs, uint32_t n, int steps, uint32_t* out, uint32_t hot, uint32_t he) {
    uint32_t cur = (blockIdx.x * 128u + threadIdx.x) * 2654435761u % n;
    const int lane = threadIdx.x & 31;
    const bool even = (lane & 1) == 0;
    uint32_t acc = 0;
    for (int s = 0; s < steps; ++s) {
        float4 a, b, c, d;
        if (MODE == 0 || MODE == 3) {
            ldg256(&recs[cur], a, b);
            ldg256(reinterpret_cast<const char*>(&recs[cur]) + 32, c, d);
            const uint32_t nxt = pick(__float_as_uint(a.x) ^ __float_as_uint(d.w), n, hot, he);
            if (MODE == 3) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(&recs[nxt]));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(&recs[nxt]) + 32));
                // ~60 dependent instructions of "box test" work
                float f = a.y;
                for (int i = 0; i < 60; ++i) f = f * 1.0001f + 0.5f;
                acc += __float_as_uint(f) & 1u;
            }
            cur = nxt;
        } else if (MODE == 1) {
            const uint32_t pcur = __shfl_xor_sync(0xffffffffu, cur, 1);
            const uint32_t c0 = even ? cur : pcur, c1 = even ? pcur : cur;
            ldg256(reinterpret_cast<const char*>(&recs[c0]) + (even ? 0 : 32), a, b); // even's record: even lane first half, odd lane second half
            ldg256(reinterpret_cast<const char*>(&recs[c1]) + (even ? 32 : 0), c, d); // odd's record
            const uint32_t mine = even ? __float_as_uint(a.x) : __float_as_uint(c.x);        // first word of my record
            const uint32_t help = even ? __float_as_uint(d.w) : __float_as_uint(b.w);        // last word of the partner's record
            const uint32_t got = __shfl_xor_sync(0xffffffffu, help, 1);                      // last word of MY record
            cur = pick(mine ^ got, n, hot, he);
        } else {
            const float4* p = reinterpret_cast<const float4*>(&recs[cur]);
            a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); d = __ldg(p + 3);
            cur = pick(__float_as_uint(a.x) ^ __float_as_uint(d.w), n, hot, he);
        }
        acc += cur;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
int main() {
    const uint32_t n = 96u << 20 >> 6; // 96 MB of records
    std::vector<Rec> h(n);
    uint64_t x = 88172645463325252ull;
    for (uint32_t i = 0; i < n; ++i) for (int j = 0; j < 16; ++j) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i].next[j] = (uint32_t)(x >> 16); }
    Rec* d; uint32_t* o;
    cudaMalloc(&d, sizeof(Rec) * (size_t)n); cudaMalloc(&o, 4);
    cudaMemcpy(d, h.data(), sizeof(Rec) * (size_t)n, cudaMemcpyHostToDevice);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int steps = 2000, grid = sms * 7;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (uint32_t he = 0; he <= 8; he += (he == 0 ? 3 : 5))
    for (int mode = 0; mode < 4; ++mode) {
        const uint32_t hot = 512;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<grid, 128>>>(d, n, steps, o, hot, he);
            else if (mode == 1) k<1><<<grid, 128>>>(d, n, steps, o, hot, he);
            else if (mode == 2) k<2><<<grid, 128>>>(d, n, steps, o, hot, he);
            else k<3><<<grid, 128>>>(d, n, steps, o, hot, he);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double fetches = (double)grid * 128 * steps;
        printf("{\"hot_eighths\": %u, \"mode\": %d, \"ms\": %.3f, \"gfetch_per_s\": %.3f, \"cycles_per_warp_fetch_per_sm\": %.1f}\n", he, mode, best, fetches / best / 1e6,
               best * 1e-3 * 1.965e9 / (fetches / 32 / sms));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
