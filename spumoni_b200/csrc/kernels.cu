// Hand-written sm_100a kernels for SPUMONI's per-read MS/PML backward search.
//
//   spb_walk_kernel<G, MS>   K1/K2 (+K3, K4 fused)
//       PML : pml_pointers::_query            /root/reference/src/compute_ms_pml.cpp:237-286
//       MS  : ms_pointers::_query             /root/reference/src/compute_ms_pml.cpp:570-623
//             ms_t::matching_statistics       /root/reference/src/compute_ms_pml.cpp:795-812
//       bin-max classifier (epilogue)         /root/reference/src/compute_ms_pml.cpp:969-995
//
// Work shape: integer pointer chasing, HBM/L2-latency bound; no tensor cores.  A group of G lanes
// (G = 4/8/16/32, a template parameter) walks one read right-to-left.  Every step needs the run
// containing the new BWT position; instead of a rank/select cascade the group issues ONE
// coalesced load of G consecutive 16-byte run records starting at a precomputed destination run
// and resolves the predecessor with a ballot + popc; the winning lane's record is broadcast with
// shuffles and becomes the state of the next step (see flat_index.h).  Reads are staged through
// shared memory in 16*G-byte chunks with a register prefetch of the next chunk; outputs are
// buffered one value per lane and written as G consecutive words.  Groups pull reads from a
// global atomic counter, so ragged read lengths balance themselves.
#include <cuda_runtime.h>
#include <stdint.h>
#include "flat_index.h"
#include "kernels.h"

namespace spb {

__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_nc4(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t rec_start(const uint4& r) { return (uint64_t)r.x | ((uint64_t)((r.w >> 8) & 0xFFu) << 32); }
__device__ __forceinline__ uint64_t rec_lf(const uint4& r) { return (uint64_t)r.y | ((uint64_t)((r.w >> 16) & 0xFFu) << 32); }
__device__ __forceinline__ uint32_t rec_head(const uint4& r) { return r.w & 0xFFu; }
__device__ __forceinline__ uint64_t crec_thr(const uint4& r) { return (uint64_t)r.x | ((uint64_t)(r.w & 0xFFu) << 32); }
__device__ __forceinline__ uint64_t crec_lf(const uint4& r) { return (uint64_t)r.y | ((uint64_t)((r.w >> 8) & 0xFFu) << 32); }

template <int G>
struct Grp {
    static constexpr unsigned LOW = (G == 32) ? 0xFFFFFFFFu : ((1u << G) - 1u);
};

// Find the last run whose start <= target, scanning forward from run w0 (start[w0] <= target).
template <int G>
__device__ __forceinline__ void locate_run(const uint4* __restrict__ runs, uint64_t R, uint64_t target, uint32_t w0,
                                           unsigned gmask, int gl, int gbase, uint32_t& k_out, uint4& rec_out) {
    for (;;) {
        uint4 rec = ldg_nc16(runs + (size_t)w0 + gl);
        bool le = rec_start(rec) <= target;
        unsigned bal = (__ballot_sync(gmask, le) >> gbase) & Grp<G>::LOW;
        int cnt = __popc(bal);
        if (cnt < G) {
            int src = cnt > 0 ? cnt - 1 : 0;
            rec_out.x = __shfl_sync(gmask, rec.x, src, G);
            rec_out.y = __shfl_sync(gmask, rec.y, src, G);
            rec_out.z = __shfl_sync(gmask, rec.z, src, G);
            rec_out.w = __shfl_sync(gmask, rec.w, src, G);
            k_out = w0 + src;
            return;
        }
        // Window exhausted: the target lies >= G-1 runs past w0.  Rare (long LF images); fall back to
        // a binary search bounded by the fact that every run holds at least one character.
        uint64_t s_last = __shfl_sync(gmask, rec_start(rec), G - 1, G);
        uint64_t lo = (uint64_t)w0 + (G - 1);
        uint64_t hi = lo + (target - s_last) + 1;
        if (hi > R + 1) hi = R + 1;  // record R+1 is padding (start = 2^40-1)
        while (hi - lo > (uint64_t)(G - 1)) {
            uint64_t mid = (lo + hi) >> 1;
            uint4 mr = ldg_nc16(runs + mid);
            if (rec_start(mr) <= target) lo = mid; else hi = mid;
        }
        w0 = (uint32_t)lo;
    }
}

// number of runs with head byte c among the runs [32*(k>>5), k)  +  cnt[cidx] of the block header
__device__ __forceinline__ uint32_t crun_ordinal(const uint8_t* __restrict__ blocks, uint32_t blk_stride, uint32_t k,
                                                 uint32_t c, uint32_t cidx) {
    const uint8_t* hdr = blocks + (size_t)(k >> 5) * blk_stride;
    uint4 h0 = ldg_nc16(hdr), h1 = ldg_nc16(hdr + 16);
    uint32_t base = ldg_nc4(hdr + 32 + 4 * cidx);
    uint32_t t = k & 31u, pat = c * 0x01010101u;
    uint32_t w[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int nb = (int)t - 4 * i;
        nb = nb < 0 ? 0 : (nb > 4 ? 4 : nb);
        uint32_t mask = nb == 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
        acc += __popc(__vcmpeq4(w[i], pat) & mask);
    }
    return base + (acc >> 3);
}

template <int G, bool MS>
__global__ void __launch_bounds__(SPB_CTA_THREADS) spb_walk_kernel(const WalkParams p) {
    constexpr int GROUPS = SPB_CTA_THREADS / G;
    constexpr int CH = (16 * G > 256) ? 256 : 16 * G;  // read bytes staged per group and refill
    constexpr int LB = CH / G;                         // bytes fetched per lane per refill (16 or 8)
    __shared__ uint4 s_chartab[256];
    __shared__ __align__(16) uint8_t s_chunk[GROUPS][CH];

    for (int i = threadIdx.x; i < 256; i += SPB_CTA_THREADS) s_chartab[i] = reinterpret_cast<const uint4*>(p.chartab)[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const int gbase = lane - gl;
    const unsigned gmask = Grp<G>::LOW << gbase;
    uint8_t* chunk = s_chunk[threadIdx.x / G];
    const uint4* __restrict__ runs = p.runs;
    const uint64_t n = p.n, R = p.R;
    const bool do_stats = p.bin_width != 0 && p.stats != nullptr;

    for (;;) {
        unsigned long long r = 0;
        if (gl == 0) r = atomicAdd(p.work_counter, 1ULL);
        r = __shfl_sync(gmask, r, 0, G);
        if (r >= p.n_reads) break;
        const uint64_t base = p.offsets[r];
        const uint64_t m = p.offsets[r + 1] - base;
        if (m == 0) {
            if (do_stats && gl == 0) { spb_bin_stats z; z.sum_max = 0; z.above = 0; z.below = 0; p.stats[r] = z; }
            continue;
        }

        // ---- stage the right-most chunk of the read, prefetch the one before it
        uint64_t q = (base + m - 1) / CH;  // chunk index in absolute buffer coordinates
        uint4 pre = make_uint4(0, 0, 0, 0);
        {
            const uint8_t* src = p.reads + q * CH + (size_t)gl * LB;
            __syncwarp(gmask);
            if (LB == 16) *reinterpret_cast<uint4*>(chunk + gl * 16) = *reinterpret_cast<const uint4*>(src);
            else *reinterpret_cast<uint2*>(chunk + gl * 8) = *reinterpret_cast<const uint2*>(src);
            if (q * CH > base) {
                const uint8_t* s2 = src - CH;
                if (LB == 16) pre = *reinterpret_cast<const uint4*>(s2);
                else { uint2 t2 = *reinterpret_cast<const uint2*>(s2); pre.x = t2.x; pre.y = t2.y; }
            }
            __syncwarp(gmask);
        }

        // ---- start with the empty string: pos = n-1 (its run is R-1)
        uint64_t pos = n - 1;
        uint32_t k = (uint32_t)(R - 1);
        uint4 krec = ldg_nc16(runs + k);
        uint32_t length = 0;
        uint64_t sample = p.last_run_sample;
        uint32_t out32 = 0;
        uint64_t out64 = 0;
        // classifier state (PML: bins are closed while walking right-to-left)
        const uint64_t bw = p.bin_width;
        uint64_t nbins = do_stats ? (m / bw > 0 ? m / bw : 1) : 1;
        uint64_t bin_lb = (nbins - 1) * bw;
        uint32_t bin_max = 0, above = 0, below = 0;
        uint64_t sum_max = 0;

        for (uint64_t i = 0; i < m; ++i) {
            const uint64_t pidx = m - 1 - i;
            const uint64_t gi = base + pidx;
            if (gi < q * CH) {  // crossed into the previous chunk: publish the prefetched bytes
                --q;
                __syncwarp(gmask);
                if (LB == 16) *reinterpret_cast<uint4*>(chunk + gl * 16) = pre;
                else *reinterpret_cast<uint2*>(chunk + gl * 8) = make_uint2(pre.x, pre.y);
                if (q * CH > base) {
                    const uint8_t* s2 = p.reads + (q - 1) * CH + (size_t)gl * LB;
                    if (LB == 16) pre = *reinterpret_cast<const uint4*>(s2);
                    else { uint2 t2 = *reinterpret_cast<const uint2*>(s2); pre.x = t2.x; pre.y = t2.y; }
                }
                __syncwarp(gmask);
            }
            uint32_t c = chunk[gi - q * CH];
            if ((p.flags & SPB_FLAG_TOUPPER) && c >= 'a' && c <= 'z') c -= 32;

            const uint4 ct = s_chartab[c];
            uint64_t npos;
            uint32_t w0;
            if (ct.w == 0xFFFFFFFFu) {
                // character absent from the index (:249 / :581): reset, then still LF with rank 0
                length = 0;
                if (MS) sample = 0;
                npos = (uint64_t)ct.x | ((uint64_t)ct.y << 32);  // F[c]
                w0 = ct.z;
            } else if (k < R && rec_head(krec) == c && c < 0x80u) {
                // bwt[pos] == c (:250 / :582).  c >= 0x80 never compares equal there (signed char)
                length++;
                if (MS) sample--;
                npos = rec_lf(krec) + (pos - rec_start(krec));
                w0 = krec.z;
            } else {
                // jump to the next / previous run of c according to the threshold (:251-279 / :583-616)
                const uint32_t j = (k < R) ? crun_ordinal(p.blocks, p.blk_stride, k, c, ct.w) : ct.y;
                const uint4 cr = ldg_nc16(p.crec + (size_t)ct.x + j);
                const bool down = pos < crec_thr(cr);
                const bool same = (k < R) && rec_head(krec) == c;  // only reachable for c >= 0x80
                length = 0;
                if (MS) {
                    const uint4 cs = ldg_nc16(p.csamp + (size_t)ct.x + j);
                    if (j < ct.y) sample = (uint64_t)cs.x | ((uint64_t)cs.y << 32);
                    if (down) sample = (uint64_t)cs.z | ((uint64_t)cs.w << 32);
                }
                const uint64_t lf = crec_lf(cr);
                npos = down ? lf - 1 : (same ? lf + (pos - rec_start(krec)) : lf);
                w0 = (down && cr.z > 0) ? cr.z - 1 : cr.z;
            }

            // ---- emit
            if (MS) {
                if ((int)(gi & (G - 1)) == gl) out64 = sample;
                if ((gi & (G - 1)) == 0 || pidx == 0) {
                    const uint64_t idx = (gi & ~(uint64_t)(G - 1)) + gl;
                    if (idx >= gi && idx < base + m) p.pointers[idx] = out64;
                }
            } else {
                if ((int)(gi & (G - 1)) == gl) out32 = length;
                if ((gi & (G - 1)) == 0 || pidx == 0) {
                    const uint64_t idx = (gi & ~(uint64_t)(G - 1)) + gl;
                    if (idx >= gi && idx < base + m) p.lengths[idx] = out32;
                }
                if (do_stats) {
                    if (pidx < bin_lb) {  // left the bin [bin_lb, ...): close it
                        if (bin_max >= p.max_value_thr) above++; else below++;
                        sum_max += bin_max;
                        bin_max = 0;
                        bin_lb -= bw;
                    }
                    bin_max = length > bin_max ? length : bin_max;
                }
            }

            // ---- one backward step: pos = LF(pos, c), and find its run
            pos = npos;
            if (pos >= n) {  // pos == n (App. C-2): virtual run R
                k = (uint32_t)R;
                krec = ldg_nc16(runs + R);
            } else {
                locate_run<G>(runs, R, pos, w0, gmask, gl, gbase, k, krec);
            }
        }

        if (!MS) {
            if (do_stats && gl == 0) {
                if (bin_max >= p.max_value_thr) above++; else below++;
                sum_max += bin_max;
                spb_bin_stats st; st.sum_max = sum_max; st.above = above; st.below = below;
                p.stats[r] = st;
            }
            continue;
        }

        // ---- MS: pointers -> lengths (left to right, carry l), :795-812
        __threadfence_block();
        __syncwarp(gmask);
        {
            const uint64_t n_text = p.n_text;
            uint64_t l = 0, prev = 0;
            uint64_t bin_ub = bw;  // exclusive upper bound of the current bin
            uint64_t bins_closed = 0;
            for (uint64_t i = 0; i < m; ++i) {
                const uint64_t ptr = __ldcg(p.pointers + base + i);
                if (i == 0 || ptr != prev + 1) {
                    for (;;) {
                        const uint64_t t = l + (uint64_t)gl;
                        const uint64_t tp = ptr + t;  // wraps like size_t (App. C-4)
                        bool ok = (i + t < m) && (tp < n_text);
                        if (ok) {
                            uint32_t rc = p.reads[base + i + t];
                            if ((p.flags & SPB_FLAG_TOUPPER) && rc >= 'a' && rc <= 'z') rc -= 32;
                            ok = rc == (uint32_t)p.text[tp];
                        }
                        const unsigned bal = (__ballot_sync(gmask, ok) >> gbase) & Grp<G>::LOW;
                        if (bal == Grp<G>::LOW) { l += G; continue; }
                        l += (uint64_t)(__ffs(~bal) - 1);
                        break;
                    }
                }
                const uint64_t gi = base + i;
                if ((int)(gi & (G - 1)) == gl) out32 = (uint32_t)l;
                if ((gi & (G - 1)) == (uint64_t)(G - 1) || i == m - 1) {
                    const uint64_t idx = (gi & ~(uint64_t)(G - 1)) + gl;
                    if (idx <= gi && idx >= base) p.lengths[idx] = out32;
                }
                if (do_stats) {
                    if (i >= bin_ub && bins_closed + 1 < nbins) {
                        if (bin_max >= p.max_value_thr) above++; else below++;
                        sum_max += bin_max;
                        bin_max = 0;
                        bin_ub += bw;
                        bins_closed++;
                    }
                    bin_max = (uint32_t)l > bin_max ? (uint32_t)l : bin_max;
                }
                l = l ? l - 1 : 0;
                prev = ptr;
            }
            if (do_stats && gl == 0) {
                if (bin_max >= p.max_value_thr) above++; else below++;
                sum_max += bin_max;
                spb_bin_stats st; st.sum_max = sum_max; st.above = above; st.below = below;
                p.stats[r] = st;
            }
        }
    }
}

template <int G, bool MS>
static cudaError_t launch_t(const WalkParams& p, int num_sms, cudaStream_t stream) {
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spb_walk_kernel<G, MS>, SPB_CTA_THREADS, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    // persistent grid: a whole number of CTAs per SM; reads are pulled from the work counter
    unsigned long long groups_needed = (p.n_reads + (SPB_CTA_THREADS / G) - 1) / (SPB_CTA_THREADS / G);
    unsigned long long grid = (unsigned long long)num_sms * occ;
    if (groups_needed < grid) grid = groups_needed ? groups_needed : 1;
    spb_walk_kernel<G, MS><<<(unsigned)grid, SPB_CTA_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_walk(const WalkParams& p, bool ms, int lanes, int num_sms, cudaStream_t stream) {
    switch (lanes) {
        case 4: return ms ? launch_t<4, true>(p, num_sms, stream) : launch_t<4, false>(p, num_sms, stream);
        case 8: return ms ? launch_t<8, true>(p, num_sms, stream) : launch_t<8, false>(p, num_sms, stream);
        case 16: return ms ? launch_t<16, true>(p, num_sms, stream) : launch_t<16, false>(p, num_sms, stream);
        case 32: return ms ? launch_t<32, true>(p, num_sms, stream) : launch_t<32, false>(p, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace spb
