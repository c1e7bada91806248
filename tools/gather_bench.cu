// Micro-benchmark: dependent 16-byte gathers over a table, thread-per-chain.
//
// Maps out what the B200 memory system gives a pointer-chasing kernel shaped like
// spb_pml_fast_kernel (one 16-byte record per step, next address taken from the record):
// steps/s as a function of table size (L2-resident .. HBM), L2 fetch granularity
// (cudaLimitMaxL2FetchGranularity), load flavour and occupancy.  Development tool only.
//
//   gather_bench <table MiB> <fetch granularity 0|32|64|128> <flavour 0..4> <threads/SM> <steps>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint4 ld_flavour(const uint4* p, int flavour) {
    uint4 r;
    switch (flavour) {
        case 0: asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
        case 1: asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
        case 2: asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
        case 3: {  // L2 evict_last through a cache-hint policy
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
            break;
        }
        case 5: asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
        case 6: asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
        case 7: asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
        case 8: asm volatile("ld.global.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
        case 9: asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
        default: {  // one 32-byte load (sm_100: ld.global.v8.b32), whole sector
            uint32_t a, b, c, d;
            asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w), "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"((const void*)((uintptr_t)p & ~(uintptr_t)31)));
            if ((uintptr_t)p & 16) { r.x = a; r.y = b; r.z = c; r.w = d; }
            break;
        }
    }
    return r;
}

__global__ void fill(uint4* t, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        t[i] = make_uint4((uint32_t)(h % n), (uint32_t)i, 1u, (uint32_t)i * 3u);
    }
}

template <int FL>
__global__ void __launch_bounds__(128) chase(const uint4* __restrict__ t, uint64_t n, int steps, uint32_t* out) {
    uint32_t k = (uint32_t)((blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 2654435761u % n);
    uint32_t acc = 0, salt = k * 0x27D4EB2Fu + 1u;
    for (int s = 0; s < steps; s++) {
        const uint4 r = ld_flavour(t + k, FL);
        // next index depends on the loaded record (r.y == k) and a per-thread salt: a random walk that never merges with other threads'
        uint32_t h = ((r.y + r.w + (r.x & 1u)) ^ salt) * 0x9E3779B1u;  // uses every word: the compiler keeps ONE 16-byte load
        h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
        salt += 0x632BE5ABu;
        k = (uint32_t)(((uint64_t)h * n) >> 32);
        acc += r.z;
    }
    if (acc == 0xFFFFFFFFu) out[0] = k;  // keep the chain alive
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = acc;
}

int main(int argc, char** argv) {
    const double mib = argc > 1 ? atof(argv[1]) : 640;
    const int gran = argc > 2 ? atoi(argv[2]) : 0;
    const int fl = argc > 3 ? atoi(argv[3]) : 0;
    const int tpsm = argc > 4 ? atoi(argv[4]) : 1280;
    const int steps = argc > 5 ? atoi(argv[5]) : 2000;
    size_t def = 0;
    CK(cudaDeviceGetLimit(&def, cudaLimitMaxL2FetchGranularity));
    if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran));
    size_t now = 0;
    CK(cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint64_t n = (uint64_t)(mib * 1048576.0 / 16.0);
    uint4* t;
    uint32_t* out;
    CK(cudaMalloc(&t, n * 16));
    CK(cudaMalloc(&out, 64));
    fill<<<sms * 8, 256>>>(t, n);
    CK(cudaDeviceSynchronize());
    const int ctas = sms * (tpsm / 128);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(a));
        switch (fl) {
            case 0: chase<0><<<ctas, 128>>>(t, n, steps, out); break;
            case 1: chase<1><<<ctas, 128>>>(t, n, steps, out); break;
            case 2: chase<2><<<ctas, 128>>>(t, n, steps, out); break;
            case 3: chase<3><<<ctas, 128>>>(t, n, steps, out); break;
            case 5: chase<5><<<ctas, 128>>>(t, n, steps, out); break;
            case 6: chase<6><<<ctas, 128>>>(t, n, steps, out); break;
            case 7: chase<7><<<ctas, 128>>>(t, n, steps, out); break;
            case 8: chase<8><<<ctas, 128>>>(t, n, steps, out); break;
            case 9: chase<9><<<ctas, 128>>>(t, n, steps, out); break;
            default: chase<4><<<ctas, 128>>>(t, n, steps, out); break;
        }
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep && ms < best) best = ms;
    }
    const double total = (double)ctas * 128 * steps;
    printf("table_mib=%.0f gran_default=%zu gran=%zu flavour=%d threads_per_sm=%d steps=%d ms=%.3f gsteps_per_s=%.2f\n", mib, def, now, fl, tpsm,
           steps, best, total / best / 1e6);
    return 0;
}
