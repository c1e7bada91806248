// GPU construction of what SPUMONI's build pipeline gets from its PFP toolchain.
//
// `spumoni build` shells out to Big-BWT and pfp_thresholds (/root/reference/src/spumoni.cpp:442-545) and then reads
// their outputs back (compute_ms_pml.cpp:44-82, :357-439; thresholds_ds.hpp:393-426):
//     <ref>.bwt.heads   one byte per run of the BWT            <ref>.bwt.len   5-byte run lengths
//     <ref>.thr_pos     5-byte threshold per run               <ref>.ssa/.esa  5 + 5 bytes (BWT position, SA value) at run starts / ends
// This file computes exactly those arrays on the device from the text itself, for texts below 2^31 - 16 characters
// (32-bit suffix indices and CUB's 32-bit item counts; the 3 Gb configurations need the 40-bit variant):
//   1. suffix array by prefix doubling: sort (8-character key, suffix), then repeatedly sort (rank[i], rank[i + h]) with h
//      doubling until every suffix has its own rank -- O(log(max LCP)) radix sorts of n (key, suffix) pairs, which is
//      what a B200 is good at (each sort streams 12 n bytes a few times; the 10^8-character pangenome of config 2 takes
//      about 13 rounds).  The text ends with its unique smallest terminator, so no comparison runs past the end;
//   2. BWT and, through the Phi array, PLCP / LCP (one thread per 512-character chunk of the text, each restarting
//      from h = 0: the classic h - 1 carry works inside a chunk);
//   3. run boundaries by a prefix sum over "BWT character differs from its predecessor";
//   4. MONI thresholds: for consecutive runs of a character ending at e and starting at s, the LEFTMOST position in
//      (e, s] that minimises LCP (0 for the first run of a character): range minima over LCP by 256-element blocks plus a
//      sparse table over the block minima, one thread per run, the previous run of the same character found by a stable
//      radix sort of the run indices by head;
//   5. SA samples at run starts / ends.
// Sorting and scanning use CUB's device-wide primitives (library code, like cuBLAS for a GEMM -- this is the build side,
// not the query hot path); everything specific to the problem is in the kernels below.  tests/build_emulator.py restates
// the same steps in numpy and is checked against the suffix-array construction of tools/mkindex on the CPU.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "build_kernels.h"

namespace spb {

namespace {

#define SPB_BUILD_CUDA(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(_e); return false; } \
    } while (0)

constexpr int kThreads = 256;
constexpr uint32_t kRmqBlock = 256;   // LCP values per block minimum
constexpr uint32_t kPlcpChunk = 512;  // text positions per PLCP thread
constexpr uint32_t kNone = 0xFFFFFFFFu;

inline unsigned grid_for(uint64_t n, int per_thread = 1) {
    const uint64_t g = (n + (uint64_t)kThreads * per_thread - 1) / ((uint64_t)kThreads * per_thread);
    return (unsigned)(g < 1 ? 1 : (g > 1048576 ? 1048576 : g));
}

// key = the 8 characters at i, big-endian (the text is followed by 8 zero bytes); sa = identity
__global__ void __launch_bounds__(kThreads) kb_init_keys(const uint8_t* __restrict__ T, uint32_t n, uint64_t* __restrict__ key, uint32_t* __restrict__ sa) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t k = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) k = (k << 8) | T[i + b];
        key[i] = k;
        sa[i] = (uint32_t)i;
    }
}
// head[j] = j where a new group of equal keys starts, else 0 (an inclusive max-scan turns it into every element's group start);
// *groups += number of groups
__global__ void __launch_bounds__(kThreads) kb_group_heads(const uint64_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ head, unsigned long long* groups) {
    unsigned long long mine = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
        const bool first = j == 0 || key[j] != key[j - 1];
        head[j] = first ? (uint32_t)j : 0u;
        mine += first;
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(groups, mine);
}
__global__ void __launch_bounds__(kThreads) kb_scatter_rank(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ start, uint32_t n, uint32_t* __restrict__ rank) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) rank[sa[j]] = start[j];
}
// key = (rank of the suffix, 1 + rank of the suffix h characters later | 0 past the end), packed into low_bits + rank bits
__global__ void __launch_bounds__(kThreads) kb_double_keys(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank, uint32_t n, uint32_t h, uint32_t low_bits, uint64_t* __restrict__ key) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = sa[j];
        const uint64_t nx = (uint64_t)s + h;
        const uint64_t lo = nx < n ? (uint64_t)rank[nx] + 1 : 0;
        key[j] = ((uint64_t)rank[s] << low_bits) | lo;
    }
}
// phi[sa[j]] = sa[j - 1] (kNone for j == 0)
__global__ void __launch_bounds__(kThreads) kb_phi(const uint32_t* __restrict__ sa, uint32_t n, uint32_t* __restrict__ phi) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) phi[sa[j]] = j ? sa[j - 1] : kNone;
}
// PLCP in place over phi: one thread per chunk of the text
__global__ void __launch_bounds__(kThreads) kb_plcp(const uint8_t* __restrict__ T, uint32_t n, uint32_t* __restrict__ phi_plcp) {
    const uint64_t chunks = ((uint64_t)n + kPlcpChunk - 1) / kPlcpChunk;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t lo = c * kPlcpChunk, hi = lo + kPlcpChunk < n ? lo + kPlcpChunk : n;
        uint64_t h = 0;
        for (uint64_t i = lo; i < hi; i++) {
            const uint32_t j = phi_plcp[i];
            if (j == kNone) { phi_plcp[i] = 0; h = 0; continue; }
            while (i + h < n && (uint64_t)j + h < n && T[i + h] == T[(uint64_t)j + h]) h++;
            phi_plcp[i] = (uint32_t)h;
            if (h) h--;
        }
    }
}
// lcp[j] = plcp[sa[j]], bwt[j] = T[sa[j] - 1] (the terminator T[n - 1] for sa[j] == 0)
__global__ void __launch_bounds__(kThreads) kb_lcp_bwt(const uint8_t* __restrict__ T, const uint32_t* __restrict__ sa, const uint32_t* __restrict__ plcp, uint32_t n,
                                                       uint32_t* __restrict__ lcp, uint8_t* __restrict__ bwt) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = sa[j];
        lcp[j] = plcp[s];
        bwt[j] = s ? T[s - 1] : T[n - 1];
    }
}
__global__ void __launch_bounds__(kThreads) kb_run_flags(const uint8_t* __restrict__ bwt, uint32_t n, uint32_t* __restrict__ flag) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) flag[j] = (j == 0 || bwt[j] != bwt[j - 1]) ? 1u : 0u;
}
// run_id = exclusive prefix sum of flag; starts[run_id[j]] = j at run starts
__global__ void __launch_bounds__(kThreads) kb_run_starts(const uint8_t* __restrict__ bwt, const uint32_t* __restrict__ run_id, uint32_t n, uint32_t* __restrict__ starts) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x)
        if (j == 0 || bwt[j] != bwt[j - 1]) starts[run_id[j]] = (uint32_t)j;
}
__global__ void __launch_bounds__(kThreads) kb_run_fields(const uint8_t* __restrict__ bwt, const uint32_t* __restrict__ sa, const uint32_t* __restrict__ starts, uint32_t r,
                                                          uint8_t* __restrict__ heads, uint32_t* __restrict__ ssa, uint32_t* __restrict__ esa, uint32_t* __restrict__ run_idx, uint8_t* __restrict__ head_key) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < r; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = starts[k], e = starts[k + 1] - 1;  // starts[r] = n
        const uint8_t c = bwt[s];
        heads[k] = c;
        head_key[k] = c;
        ssa[k] = sa[s];
        esa[k] = sa[e];
        run_idx[k] = (uint32_t)k;
    }
}
// leftmost minimum of every block of kRmqBlock LCP values: one warp per block
__global__ void __launch_bounds__(kThreads) kb_block_min(const uint32_t* __restrict__ lcp, uint32_t n, uint32_t nblk, uint32_t* __restrict__ lvl0) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t b = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nblk; b += warps) {
        uint32_t best_v = 0xFFFFFFFFu, best_p = 0xFFFFFFFFu;
        for (uint32_t t = lane; t < kRmqBlock; t += 32) {
            const uint64_t p = b * kRmqBlock + t;
            if (p < n) { const uint32_t v = lcp[p]; if (v < best_v) { best_v = v; best_p = (uint32_t)p; } }  // a lane sees its positions in increasing order
        }
        for (int o = 16; o; o >>= 1) {
            const uint32_t ov = __shfl_xor_sync(0xFFFFFFFFu, best_v, o), op = __shfl_xor_sync(0xFFFFFFFFu, best_p, o);
            if (ov < best_v || (ov == best_v && op < best_p)) { best_v = ov; best_p = op; }
        }
        if (lane == 0) lvl0[b] = best_p;
    }
}
// level j from level j - 1: window of 2^j blocks starting at block a = better of the two half windows (ties: the left one)
__global__ void __launch_bounds__(kThreads) kb_sparse_level(const uint32_t* __restrict__ lcp, const uint32_t* __restrict__ prev, uint32_t* __restrict__ cur, uint32_t count, uint32_t half) {
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < count; a += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = prev[a], y = prev[a + half];
        cur[a] = lcp[y] < lcp[x] ? y : x;
    }
}
__device__ __forceinline__ uint32_t better(const uint32_t* __restrict__ lcp, uint32_t a, uint32_t b) { return lcp[b] < lcp[a] ? b : a; }  // a lies left of b
__device__ __forceinline__ uint32_t scan_min(const uint32_t* __restrict__ lcp, uint64_t l, uint64_t r) {  // leftmost argmin of lcp[l..r]
    uint32_t bv = lcp[l], bp = (uint32_t)l;
    for (uint64_t p = l + 1; p <= r; p++) { const uint32_t v = lcp[p]; if (v < bv) { bv = v; bp = (uint32_t)p; } }
    return bp;
}
// thr[k] for run sorted[t]: runs sorted (stably) by head, so the previous run of the same character is sorted[t - 1]
__global__ void __launch_bounds__(kThreads) kb_thresholds(const uint32_t* __restrict__ lcp, const uint32_t* __restrict__ starts, const uint8_t* __restrict__ heads,
                                                          const uint32_t* __restrict__ sorted, uint32_t r, const uint32_t* __restrict__ levels, const uint64_t* __restrict__ level_off,
                                                          uint32_t* __restrict__ thr) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < r; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = sorted[t];
        if (t == 0 || heads[sorted[t - 1]] != heads[k]) { thr[k] = 0; continue; }
        const uint64_t l = (uint64_t)starts[sorted[t - 1] + 1], rr = starts[k];  // (end of the previous run of the character, start of this one]
        const uint64_t bl = l / kRmqBlock, br = rr / kRmqBlock;
        uint32_t best;
        if (bl == br) best = scan_min(lcp, l, rr);
        else {
            best = scan_min(lcp, l, (bl + 1) * kRmqBlock - 1);
            if (br - bl > 1) {
                const uint64_t a = bl + 1, b = br - 1;
                const uint32_t j = 31u - (uint32_t)__clz((uint32_t)(b - a + 1));
                const uint32_t* lv = levels + level_off[j];
                best = better(lcp, best, better(lcp, lv[a], lv[b - (1ull << j) + 1]));
            }
            best = better(lcp, best, scan_min(lcp, br * kRmqBlock, rr));
        }
        thr[k] = best;
    }
}

template <typename T>
struct DeviceArray {
    T* p = nullptr;
    ~DeviceArray() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
};

uint32_t bits_for(uint64_t v) { uint32_t b = 1; while ((v >> b) != 0) b++; return b; }

}  // namespace

bool build_rlbwt_device(const uint8_t* text, uint64_t n64, int device, BuildParts& out, std::string& err, bool verbose) {
    if (n64 < 2 || n64 >= 0x7FFFFFF0ull) { err = "text length must be in [2, 2^31 - 16) for the 32-bit device builder"; return false; }
    if (text[n64 - 1] != 0) { err = "the text must end with its terminator (a zero byte)"; return false; }
    const uint32_t n = (uint32_t)n64;
    SPB_BUILD_CUDA(cudaSetDevice(device));
    cudaEvent_t ev0, ev1;
    SPB_BUILD_CUDA(cudaEventCreate(&ev0));
    SPB_BUILD_CUDA(cudaEventCreate(&ev1));
    auto lap = [&](const char* what) {
        if (!verbose) return;
        cudaEventRecord(ev1);
        cudaEventSynchronize(ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, ev0, ev1);
        std::fprintf(stderr, "[build] %-28s %9.2f ms\n", what, ms);
        cudaEventRecord(ev0);
    };
    DeviceArray<uint8_t> dT, d_bwt;
    DeviceArray<uint64_t> key_a, key_b;
    DeviceArray<uint32_t> sa_a, sa_b, rank, aux;
    DeviceArray<unsigned long long> d_groups;
    SPB_BUILD_CUDA(dT.alloc((size_t)n + 16));
    SPB_BUILD_CUDA(key_a.alloc(n));
    SPB_BUILD_CUDA(key_b.alloc(n));
    SPB_BUILD_CUDA(sa_a.alloc(n));
    SPB_BUILD_CUDA(sa_b.alloc(n));
    SPB_BUILD_CUDA(rank.alloc(n));
    SPB_BUILD_CUDA(aux.alloc((size_t)n + 1));
    SPB_BUILD_CUDA(d_groups.alloc(1));
    SPB_BUILD_CUDA(cudaMemset(dT.p + n, 0, 16));
    SPB_BUILD_CUDA(cudaMemcpy(dT.p, text, n, cudaMemcpyHostToDevice));
    cudaEventRecord(ev0);

    // ---- 1. suffix array
    cub::DoubleBuffer<uint64_t> keys(key_a.p, key_b.p);
    cub::DoubleBuffer<uint32_t> sas(sa_a.p, sa_b.p);
    size_t tmp_sort = 0, tmp_scan = 0;
    SPB_BUILD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys, sas, (int)n, 0, 64));
    SPB_BUILD_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tmp_scan, aux.p, aux.p, cub::Max(), (int)n));
    {
        size_t t2 = 0;
        SPB_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, aux.p, aux.p, (int)n));
        if (t2 > tmp_scan) tmp_scan = t2;
    }
    DeviceArray<uint8_t> d_tmp;
    const size_t tmp_bytes = tmp_sort > tmp_scan ? tmp_sort : tmp_scan;
    SPB_BUILD_CUDA(d_tmp.alloc(tmp_bytes + 256));
    kb_init_keys<<<grid_for(n), kThreads>>>(dT.p, n, keys.Current(), sas.Current());
    {
        size_t tb = tmp_bytes;
        SPB_BUILD_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tb, keys, sas, (int)n, 0, 64));
    }
    const uint32_t low_bits = bits_for((uint64_t)n + 1), rank_bits = bits_for(n);
    int rounds = 0;
    for (uint64_t h = 8;; h *= 2) {
        SPB_BUILD_CUDA(cudaMemset(d_groups.p, 0, sizeof(unsigned long long)));
        kb_group_heads<<<grid_for(n), kThreads>>>(keys.Current(), n, aux.p, d_groups.p);
        {
            size_t tb = tmp_bytes;
            SPB_BUILD_CUDA(cub::DeviceScan::InclusiveScan(d_tmp.p, tb, aux.p, aux.p, cub::Max(), (int)n));
        }
        unsigned long long groups = 0;
        SPB_BUILD_CUDA(cudaMemcpy(&groups, d_groups.p, sizeof(groups), cudaMemcpyDeviceToHost));
        if (groups == n) break;
        if (h >= n64 * 2) { err = "prefix doubling did not converge (is the terminator unique?)"; return false; }
        kb_scatter_rank<<<grid_for(n), kThreads>>>(sas.Current(), aux.p, n, rank.p);
        kb_double_keys<<<grid_for(n), kThreads>>>(sas.Current(), rank.p, n, (uint32_t)(h > 0xFFFFFFFFull ? 0xFFFFFFFFu : h), low_bits, keys.Current());
        {
            size_t tb = tmp_bytes;
            SPB_BUILD_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tb, keys, sas, (int)n, 0, (int)(low_bits + rank_bits)));
        }
        rounds++;
    }
    SPB_BUILD_CUDA(cudaGetLastError());
    if (verbose) std::fprintf(stderr, "[build] prefix doubling: %d rounds\n", rounds);
    lap("suffix array");
    uint32_t* sa = sas.Current();
    uint32_t* spare = sas.Alternate();  // n uint32: becomes lcp

    // ---- 2. BWT, PLCP (in `rank`), LCP (in `spare`)
    SPB_BUILD_CUDA(d_bwt.alloc((size_t)n + 1));
    kb_phi<<<grid_for(n), kThreads>>>(sa, n, rank.p);
    kb_plcp<<<grid_for(((uint64_t)n + kPlcpChunk - 1) / kPlcpChunk), kThreads>>>(dT.p, n, rank.p);
    kb_lcp_bwt<<<grid_for(n), kThreads>>>(dT.p, sa, rank.p, n, spare, d_bwt.p);
    uint32_t* lcp = spare;
    SPB_BUILD_CUDA(cudaGetLastError());
    lap("bwt + plcp + lcp");

    // ---- 3. runs
    kb_run_flags<<<grid_for(n), kThreads>>>(d_bwt.p, n, aux.p);
    {
        size_t tb = tmp_bytes;
        SPB_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, aux.p, aux.p, (int)n));
    }
    uint32_t last_id = 0;
    SPB_BUILD_CUDA(cudaMemcpy(&last_id, aux.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
    // the run id of the last position is r - 1 (exclusive sum + the flag of that position: its id counts the flags before it)
    // exclusive sum: aux[j] = number of run starts before j; if the last position is itself a run start it opens run aux[n-1],
    // else it belongs to run aux[n-1] - 1
    uint8_t last2[2] = {0, 0};
    SPB_BUILD_CUDA(cudaMemcpy(last2, d_bwt.p + (n - 2), 2, cudaMemcpyDeviceToHost));
    const uint32_t runs = last2[0] != last2[1] ? last_id + 1 : last_id;
    DeviceArray<uint32_t> d_starts, d_ssa, d_esa, d_thr, d_idx_a, d_idx_b;
    DeviceArray<uint8_t> d_heads, d_hk_a, d_hk_b;
    SPB_BUILD_CUDA(d_starts.alloc((size_t)runs + 1));
    SPB_BUILD_CUDA(d_ssa.alloc(runs));
    SPB_BUILD_CUDA(d_esa.alloc(runs));
    SPB_BUILD_CUDA(d_thr.alloc(runs));
    SPB_BUILD_CUDA(d_idx_a.alloc(runs));
    SPB_BUILD_CUDA(d_idx_b.alloc(runs));
    SPB_BUILD_CUDA(d_heads.alloc(runs));
    SPB_BUILD_CUDA(d_hk_a.alloc(runs));
    SPB_BUILD_CUDA(d_hk_b.alloc(runs));
    // run ids are "number of starts before j": a start at j has id = index of its run
    kb_run_starts<<<grid_for(n), kThreads>>>(d_bwt.p, aux.p, n, d_starts.p);
    SPB_BUILD_CUDA(cudaMemcpy(d_starts.p + runs, &n, sizeof(uint32_t), cudaMemcpyHostToDevice));
    kb_run_fields<<<grid_for(runs), kThreads>>>(d_bwt.p, sa, d_starts.p, runs, d_heads.p, d_ssa.p, d_esa.p, d_idx_a.p, d_hk_a.p);
    SPB_BUILD_CUDA(cudaGetLastError());
    lap("runs + samples");

    // ---- 4. thresholds
    cub::DoubleBuffer<uint8_t> hk(d_hk_a.p, d_hk_b.p);
    cub::DoubleBuffer<uint32_t> idx(d_idx_a.p, d_idx_b.p);
    {
        size_t tb = 0;
        SPB_BUILD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, hk, idx, (int)runs, 0, 8));
        DeviceArray<uint8_t> t2;
        SPB_BUILD_CUDA(t2.alloc(tb + 256));
        SPB_BUILD_CUDA(cub::DeviceRadixSort::SortPairs(t2.p, tb, hk, idx, (int)runs, 0, 8));  // stable: runs of a character stay in BWT order
        SPB_BUILD_CUDA(cudaDeviceSynchronize());
    }
    const uint32_t nblk = (uint32_t)(((uint64_t)n + kRmqBlock - 1) / kRmqBlock);
    std::vector<uint64_t> level_off;
    uint64_t total_lv = 0;
    for (uint32_t j = 0; (1ull << j) <= nblk; j++) { level_off.push_back(total_lv); total_lv += (uint64_t)nblk - (1ull << j) + 1; }
    DeviceArray<uint32_t> d_levels;
    DeviceArray<uint64_t> d_level_off;
    SPB_BUILD_CUDA(d_levels.alloc((size_t)total_lv));
    SPB_BUILD_CUDA(d_level_off.alloc(level_off.size()));
    SPB_BUILD_CUDA(cudaMemcpy(d_level_off.p, level_off.data(), level_off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    kb_block_min<<<grid_for((uint64_t)nblk * 32), kThreads>>>(lcp, n, nblk, d_levels.p);
    for (uint32_t j = 1; j < level_off.size(); j++) {
        const uint32_t count = nblk - (1u << j) + 1;
        kb_sparse_level<<<grid_for(count), kThreads>>>(lcp, d_levels.p + level_off[j - 1], d_levels.p + level_off[j], count, 1u << (j - 1));
    }
    kb_thresholds<<<grid_for(runs), kThreads>>>(lcp, d_starts.p, d_heads.p, idx.Current(), runs, d_levels.p, d_level_off.p, d_thr.p);
    SPB_BUILD_CUDA(cudaGetLastError());
    SPB_BUILD_CUDA(cudaDeviceSynchronize());
    lap("thresholds");

    // ---- 5. back to the host
    out.n = n64;
    out.heads.resize(runs);
    std::vector<uint32_t> starts((size_t)runs + 1), thr(runs), ssa(runs), esa(runs);
    SPB_BUILD_CUDA(cudaMemcpy(out.heads.data(), d_heads.p, runs, cudaMemcpyDeviceToHost));
    SPB_BUILD_CUDA(cudaMemcpy(starts.data(), d_starts.p, ((size_t)runs + 1) * 4, cudaMemcpyDeviceToHost));
    SPB_BUILD_CUDA(cudaMemcpy(thr.data(), d_thr.p, (size_t)runs * 4, cudaMemcpyDeviceToHost));
    SPB_BUILD_CUDA(cudaMemcpy(ssa.data(), d_ssa.p, (size_t)runs * 4, cudaMemcpyDeviceToHost));
    SPB_BUILD_CUDA(cudaMemcpy(esa.data(), d_esa.p, (size_t)runs * 4, cudaMemcpyDeviceToHost));
    out.lens.resize(runs); out.thr.resize(runs); out.ssa.resize(runs); out.esa.resize(runs);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)runs; k++) {
        out.lens[(size_t)k] = (uint64_t)starts[(size_t)k + 1] - starts[(size_t)k];
        out.thr[(size_t)k] = thr[(size_t)k];
        out.ssa[(size_t)k] = ssa[(size_t)k];
        out.esa[(size_t)k] = esa[(size_t)k];
    }
    lap("copy back");
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return true;
}

}  // namespace spb
