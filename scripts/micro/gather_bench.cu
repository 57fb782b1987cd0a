// Random-gather microbenchmark for B200: what bounds a walk step (row info -> neighbour column)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
// Modes: single array of S bytes gathered with element width W and cache hint H, or the walk pattern
// (8-byte row info array of R bytes, then a 4-byte column array of S bytes, dependent).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }

template <int H>
__device__ __forceinline__ uint32_t ld4(const uint32_t *p, uint64_t pl, uint64_t pf) {
    uint32_t v;
    if (H == 0) v = __ldg(p);
    else if (H == 1) asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pl));
    else if (H == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pf));
    else asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <int H>
__device__ __forceinline__ uint2 ld8(const uint2 *p, uint64_t pl, uint64_t pf) {
    uint2 v;
    if (H == 0) v = __ldg(p);
    else if (H == 1) asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pl));
    else if (H == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pf));
    else asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// independent gathers: UN per thread in flight
template <int W, int H, int UN>
__global__ void gather_kernel(const uint32_t *a, uint64_t n_elems, int iters, uint32_t *out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t pl = pol_last(), pf = pol_first();
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        uint32_t v[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) {
            const uint64_t idx = ((uint64_t)mix(tid * 2654435761u + it * UN + u) * n_elems) >> 32;
            if (W == 4) v[u] = ld4<H>(a + idx, pl, pf);
            else if (W == 8) { uint2 q = ld8<H>((const uint2 *)a + idx, pl, pf); v[u] = q.x ^ q.y; }
            else { uint4 q = __ldg((const uint4 *)a + idx); v[u] = q.x ^ q.y ^ q.z ^ q.w; }
        }
#pragma unroll
        for (int u = 0; u < UN; u++) acc += v[u];
    }
    if (acc == 0x12345678u) out[tid] = acc;
}

// walk pattern: node -> rowinfo[node] (8 B) -> col[start + r % deg] (4 B) -> node ...
template <int HR, int HC, int UN>
__global__ void walk_kernel(const uint2 *rowinfo, const uint32_t *col, uint32_t n_nodes, int hops, int iters, uint32_t *out,
                            uint32_t *sink, int write_bytes) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t pl = pol_last(), pf = pol_first();
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        uint32_t cur[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) cur[u] = (uint32_t)(((uint64_t)mix(tid * 2654435761u + it * UN + u) * n_nodes) >> 32);
        for (int h = 0; h < hops; h++) {
            uint2 ri[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) ri[u] = ld8<HR>(rowinfo + cur[u], pl, pf);
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const uint32_t off = __umulhi(mix(cur[u] + h + it), ri[u].y);
                cur[u] = ld4<HC>(col + ri[u].x + off, pl, pf);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; u++) acc += cur[u];
        // optional streaming output (emulates the SpG rows being written)
        for (int b = 0; b < write_bytes; b += 4) sink[((size_t)tid * iters + it) * (write_bytes / 4) + b / 4] = acc;
    }
    if (acc == 0x12345678u) out[tid] = acc;
}

template <typename F>
static float time_it(F f, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char **argv) {
    const int blocks = 148 * 8, threads = 256;
    const int64_t nthreads = (int64_t)blocks * threads;
    uint32_t *out; cudaMalloc(&out, nthreads * 4);
    const char *only = argc > 1 ? argv[1] : "all";
    // ---- single-array gathers
    if (!strcmp(only, "all") || !strcmp(only, "gather")) {
        const size_t sizes[] = {8u << 20, 64u << 20, 243u << 20, 1024u << 20};
        for (size_t S : sizes) {
            uint32_t *a; cudaMalloc(&a, S); cudaMemset(a, 1, S);
            const int iters = 32;
            auto report = [&](const char *name, float ms, int un) {
                printf("gather %-22s S=%5zu MB  %7.3f ms  %7.1f G acc/s\n", name, S >> 20, ms, nthreads * (double)iters * un / ms / 1e6);
            };
            report("4B ldg UN4", time_it([&] { gather_kernel<4, 0, 4><<<blocks, threads>>>(a, S / 4, iters, out); }, 5), 4);
            report("4B ldg UN8", time_it([&] { gather_kernel<4, 0, 8><<<blocks, threads>>>(a, S / 4, iters, out); }, 5), 8);
            report("4B evict_last UN8", time_it([&] { gather_kernel<4, 1, 8><<<blocks, threads>>>(a, S / 4, iters, out); }, 5), 8);
            report("4B noalloc+first UN8", time_it([&] { gather_kernel<4, 2, 8><<<blocks, threads>>>(a, S / 4, iters, out); }, 5), 8);
            report("4B noalloc UN8", time_it([&] { gather_kernel<4, 3, 8><<<blocks, threads>>>(a, S / 4, iters, out); }, 5), 8);
            report("8B ldg UN8", time_it([&] { gather_kernel<8, 0, 8><<<blocks, threads>>>(a, S / 8, iters, out); }, 5), 8);
            report("16B ldg UN8", time_it([&] { gather_kernel<16, 0, 8><<<blocks, threads>>>(a, S / 16, iters, out); }, 5), 8);
            cudaFree(a);
        }
    }
    // ---- walk pattern (ppa-like: 576k nodes, 60.6M edges)
    if (!strcmp(only, "all") || !strcmp(only, "walk")) {
        const uint32_t N = 576289; const size_t E = 60637674;
        uint2 *ri; uint32_t *col, *sink;
        cudaMalloc(&ri, (size_t)N * 8); cudaMalloc(&col, E * 4);
        const size_t sink_bytes = (size_t)nthreads * 16 * 64;
        cudaMalloc(&sink, sink_bytes);
        // uniform degrees (E/N each) and uniformly random neighbours are enough for the memory behaviour
        uint2 *h_ri = (uint2 *)malloc((size_t)N * 8); uint32_t *h_col = (uint32_t *)malloc(E * 4);
        const uint32_t deg = (uint32_t)(E / N);
        for (uint32_t i = 0; i < N; i++) h_ri[i] = make_uint2(i * deg, deg);
        uint64_t s = 88172645463325252ull;
        for (size_t e = 0; e < E; e++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h_col[e] = (uint32_t)(s % N); }
        cudaMemcpy(ri, h_ri, (size_t)N * 8, cudaMemcpyHostToDevice); cudaMemcpy(col, h_col, E * 4, cudaMemcpyHostToDevice);
        const int iters = 16, hops = 2;
        auto report = [&](const char *name, float ms, int un) {
            printf("walk   %-34s %7.3f ms  %7.1f G hops/s\n", name, ms, nthreads * (double)iters * un * hops / ms / 1e6);
        };
        report("ri ldg, col ldg", time_it([&] { walk_kernel<0, 0, 8><<<blocks, threads>>>(ri, col, N, hops, iters, out, sink, 0); }, 5), 8);
        report("ri evict_last, col ldg", time_it([&] { walk_kernel<1, 0, 8><<<blocks, threads>>>(ri, col, N, hops, iters, out, sink, 0); }, 5), 8);
        report("ri evict_last, col noalloc+first", time_it([&] { walk_kernel<1, 2, 8><<<blocks, threads>>>(ri, col, N, hops, iters, out, sink, 0); }, 5), 8);
        report("ri evict_last, col noalloc", time_it([&] { walk_kernel<1, 3, 8><<<blocks, threads>>>(ri, col, N, hops, iters, out, sink, 0); }, 5), 8);
        report("ri ldg, col ldg + 64B writes", time_it([&] { walk_kernel<0, 0, 8><<<blocks, threads>>>(ri, col, N, hops, iters, out, sink, 64); }, 5), 8);
        report("ri evict_last, col ldg + 64B writes", time_it([&] { walk_kernel<1, 0, 8><<<blocks, threads>>>(ri, col, N, hops, iters, out, sink, 64); }, 5), 8);
        report("ri ldg, col ldg UN4", time_it([&] { walk_kernel<0, 0, 4><<<blocks, threads>>>(ri, col, N, hops, iters, out, sink, 0); }, 5), 4);
    }
    return 0;
}
