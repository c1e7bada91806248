// Micro-benchmark 2: what the memory system gives a thread-per-read walk shaped like spb_move_kernel.
// Each thread chases 16-byte rows of a table (one random gather per step, optionally a neighbour row
// with probability ~0.6) WHILE it streams its own read (8 bytes every 8 steps) and writes its own
// output (16 bytes every 16 steps for 1-byte values / every 4 steps for 4-byte values).  The knobs
// select the cache policy of each of the three streams, so that an ncu run shows which one costs the
// L2 its hit rate.  Development tool only.
//
//   gather_bench2 <table MiB> <threads/SM> <steps (multiple of 128)> <nbr 0|1>
//                 <rd 0 none|1 plain 8 B|3 nc.no_allocate+evict_first 8 B|32/64/128 = bytes staged per thread in shared memory>
//                 <wr 0 none|1 plain 16 B|2 .cs 16 B|32/64/128 = bytes staged per thread in shared memory, .cs> <out bytes 1|4>
//                 <row 0 nc.no_allocate|1 +L2 evict_last|2 plain>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void fill(uint4* t, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        t[i] = make_uint4((uint32_t)(h % n), (uint32_t)i, 1u, (uint32_t)i * 3u);
    }
}
template <int ROW>
__device__ __forceinline__ uint4 ld_row(const uint4* p, uint64_t pol) {
    uint4 r;
    if (ROW == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (ROW == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    else asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int RD>
__device__ __forceinline__ uint64_t ld_read(const uint8_t* p, uint64_t pol) {
    uint64_t v;
    if (RD == 1) asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (RD == 2) asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
template <int WR>
__device__ __forceinline__ void st_out(void* p, uint4 v, uint64_t pol) {
    if (WR == 1) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (WR == 2) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (WR == 3) asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
    else asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int NBR, int RD, int WR, int OB, int ROW>
__global__ void __launch_bounds__(128) walk(const uint4* __restrict__ t, uint64_t n, int steps, const uint8_t* __restrict__ reads, uint8_t* __restrict__ outp, uint32_t* sink) {
    // staging buffers: chunk j (16 bytes) of thread x lives at [j * 128 + x] (conflict-free for 16-byte lane accesses)
    constexpr int RCH = RD >= 32 ? RD / 16 : 0, WCH = WR >= 32 ? WR / 16 : 0;
    __shared__ uint4 s_rd[RCH ? RCH * 128 : 1];
    __shared__ uint4 s_wr[WCH ? WCH * 128 : 1];
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t k = (uint32_t)(tid * 2654435761u % n);
    uint32_t acc = 0, salt = k * 0x27D4EB2Fu + 1u;
    uint64_t pol_first, pol_last;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    const uint8_t* rp = reads + tid * (uint64_t)steps;   // this thread's read (steps bytes), walked backwards
    uint8_t* op = outp + tid * (uint64_t)steps * OB;
    uint64_t chunk = 0;
    uint4 ob = make_uint4(0, 0, 0, 0);
    for (int s = steps - 1; s >= 0; s--) {
        uint32_t c = 0;
        if (RD >= 32) {
            if ((s & (RD - 1)) == RD - 1) {  // refill the staged segment: RCH independent 16-byte loads
                const uint4* src = reinterpret_cast<const uint4*>(rp + (s & ~(RD - 1)));
#pragma unroll
                for (int j = 0; j < RCH; j++) {
                    uint4 a;
                    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(src + j), "l"(pol_first));
                    s_rd[j * 128 + threadIdx.x] = a;
                }
            }
            if ((s & 7) == 7) chunk = reinterpret_cast<const uint64_t*>(&s_rd[((s & (RD - 1)) >> 4) * 128 + threadIdx.x])[(s >> 3) & 1];
            c = (uint32_t)(chunk >> ((s & 7) * 8)) & 0xFFu;
        } else if (RD) {
            if ((s & 7) == 7) chunk = ld_read<RD>(rp + (s & ~7), pol_first);
            c = (uint32_t)(chunk >> ((s & 7) * 8)) & 0xFFu;
        }
        uint4 r = ld_row<ROW>(t + k, ROW == 1 ? pol_last : pol_first);
        uint32_t h = ((r.y + r.w + (r.x & 1u)) ^ salt ^ c) * 0x9E3779B1u;
        h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
        if (NBR && (h & 0xFFu) < 150u) {  // ~0.59 of the steps read a neighbour row too (k + 1..4)
            const uint32_t k2 = k + 1 + ((h >> 8) & 3u);
            const uint4 r2 = ld_row<ROW>(t + (k2 < n ? k2 : k), ROW == 1 ? pol_last : pol_first);
            h += r2.y * 0x632BE5ABu;
            h ^= h >> 15;
        }
        salt += 0x632BE5ABu;
        k = (uint32_t)(((uint64_t)h * n) >> 32);
        acc += r.z;
        if (WR) {
            constexpr int VPW = 16 / OB;  // values per 16-byte word
            const uint32_t sl = s & (VPW - 1), v = h & 0x3Fu;
            if (OB == 1) {
                const uint32_t w = sl >> 2, sh = (sl & 3) * 8;
                if (w == 0) ob.x |= v << sh; else if (w == 1) ob.y |= v << sh; else if (w == 2) ob.z |= v << sh; else ob.w |= v << sh;
            } else {
                if (sl == 0) ob.x = v; else if (sl == 1) ob.y = v; else if (sl == 2) ob.z = v; else ob.w = v;
            }
            if (sl == 0) {
                const size_t byte0 = (size_t)(s & ~(VPW - 1)) * OB;  // byte offset of this word in the thread's output
                if (WR >= 32) {
                    s_wr[((byte0 & (WR - 1)) >> 4) * 128 + threadIdx.x] = ob;
                    if ((byte0 & (WR - 1)) == 0) {  // segment complete (walking downwards): flush WCH consecutive words
#pragma unroll
                        for (int j = 0; j < WCH; j++) st_out<2>(op + byte0 + 16 * j, s_wr[j * 128 + threadIdx.x], pol_first);
                    }
                } else st_out<WR>(op + byte0, ob, pol_first);
                ob = make_uint4(0, 0, 0, 0);
            }
        }
    }
    if (acc == 0xFFFFFFFFu) sink[0] = k;
    if (threadIdx.x == 0 && blockIdx.x == 0) sink[1] = acc;
}

template <int NBR, int RD, int WR, int OB>
void launch_row(int row, int ctas, const uint4* t, uint64_t n, int steps, const uint8_t* rd, uint8_t* out, uint32_t* sink) {
    if (row == 0) walk<NBR, RD, WR, OB, 0><<<ctas, 128>>>(t, n, steps, rd, out, sink);
    else if (row == 1) walk<NBR, RD, WR, OB, 1><<<ctas, 128>>>(t, n, steps, rd, out, sink);
    else walk<NBR, RD, WR, OB, 2><<<ctas, 128>>>(t, n, steps, rd, out, sink);
}
template <int NBR, int RD, int WR>
void launch_ob(int ob, int row, int ctas, const uint4* t, uint64_t n, int steps, const uint8_t* rd, uint8_t* out, uint32_t* sink) {
    if (ob == 1) launch_row<NBR, RD, WR, 1>(row, ctas, t, n, steps, rd, out, sink); else launch_row<NBR, RD, WR, 4>(row, ctas, t, n, steps, rd, out, sink);
}
template <int NBR, int RD>
void launch_wr(int wr, int ob, int row, int ctas, const uint4* t, uint64_t n, int steps, const uint8_t* rd, uint8_t* out, uint32_t* sink) {
    switch (wr) {
        case 0: launch_ob<NBR, RD, 0>(ob, row, ctas, t, n, steps, rd, out, sink); break;
        case 1: launch_ob<NBR, RD, 1>(ob, row, ctas, t, n, steps, rd, out, sink); break;
        case 2: launch_ob<NBR, RD, 2>(ob, row, ctas, t, n, steps, rd, out, sink); break;
        case 32: launch_ob<NBR, RD, 32>(ob, row, ctas, t, n, steps, rd, out, sink); break;
        case 64: launch_ob<NBR, RD, 64>(ob, row, ctas, t, n, steps, rd, out, sink); break;
        default: launch_ob<NBR, RD, 128>(ob, row, ctas, t, n, steps, rd, out, sink); break;
    }
}
template <int NBR>
void launch_rd(int rdm, int wr, int ob, int row, int ctas, const uint4* t, uint64_t n, int steps, const uint8_t* rd, uint8_t* out, uint32_t* sink) {
    switch (rdm) {
        case 0: launch_wr<NBR, 0>(wr, ob, row, ctas, t, n, steps, rd, out, sink); break;
        case 1: launch_wr<NBR, 1>(wr, ob, row, ctas, t, n, steps, rd, out, sink); break;
        case 3: launch_wr<NBR, 3>(wr, ob, row, ctas, t, n, steps, rd, out, sink); break;
        case 32: launch_wr<NBR, 32>(wr, ob, row, ctas, t, n, steps, rd, out, sink); break;
        case 64: launch_wr<NBR, 64>(wr, ob, row, ctas, t, n, steps, rd, out, sink); break;
        default: launch_wr<NBR, 128>(wr, ob, row, ctas, t, n, steps, rd, out, sink); break;
    }
}

int main(int argc, char** argv) {
    const double mib = argc > 1 ? atof(argv[1]) : 151;
    const int tpsm = argc > 2 ? atoi(argv[2]) : 1280;
    const int steps = argc > 3 ? atoi(argv[3]) : 1024;
    const int nbr = argc > 4 ? atoi(argv[4]) : 0;
    const int rdm = argc > 5 ? atoi(argv[5]) : 0;
    const int wr = argc > 6 ? atoi(argv[6]) : 0;
    const int ob = argc > 7 ? atoi(argv[7]) : 1;
    const int row = argc > 8 ? atoi(argv[8]) : 0;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint64_t n = (uint64_t)(mib * 1048576.0 / 16.0);
    const int ctas = sms * (tpsm / 128);
    const uint64_t threads = (uint64_t)ctas * 128;
    uint4* t; uint32_t* sink; uint8_t *rd, *out;
    CK(cudaMalloc(&t, n * 16));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMalloc(&rd, threads * steps + 64));
    CK(cudaMalloc(&out, threads * steps * (uint64_t)ob + 64));
    CK(cudaMemset(rd, 7, threads * steps + 64));
    fill<<<sms * 8, 256>>>(t, n);
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(a));
        if (nbr) launch_rd<1>(rdm, wr, ob, row, ctas, t, n, steps, rd, out, sink); else launch_rd<0>(rdm, wr, ob, row, ctas, t, n, steps, rd, out, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep && ms < best) best = ms;
    }
    printf("table_mib=%.0f tpsm=%d steps=%d nbr=%d rd=%d wr=%d ob=%d row=%d ms=%.3f gsteps_per_s=%.2f\n", mib, tpsm, steps, nbr, rdm, wr, ob, row, best, (double)threads * steps / best / 1e6);
    return 0;
}
