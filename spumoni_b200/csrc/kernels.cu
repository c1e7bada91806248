// Hand-written sm_100a kernels for SPUMONI's per-read MS/PML backward search.
//
//   spb_move_kernel<MS, STATS, OW, LD, RW, DOC, MSI>       one thread per read over the move rows (default path);
//       RW = bytes per row (8 compact / 16 wide / 32 full), OW = bytes per length (4 / 2 / 1, or 0 = one bit per base)
//       PML : pml_pointers::_query            /root/reference/src/compute_ms_pml.cpp:237-286 (+ doc numbers :288-340)
//       MS  : ms_pointers::_query             /root/reference/src/compute_ms_pml.cpp:570-623 (+ doc numbers :625-682)
//       bin-max classifier (PML, fused)       /root/reference/src/compute_ms_pml.cpp:969-995
//   spb_ms_lengths_kernel<STATS, OW>                       MS pointers -> lengths, one thread per read
//       ms_t::matching_statistics             /root/reference/src/compute_ms_pml.cpp:795-812, classifier :1150-1176
//   spb_walk_kernel<G, MS>                                 G lanes per read over the generic run records
//       (indexes without fast symbols; every symbol and size; both MS phases in one launch)
//
// Work shape: integer pointer chasing, one dependent chain per read; no tensor cores, no GEMM-shaped work.
// What bounds it on a B200 and why the tables look the way they do: DESIGN.md section 4.3.
// The generic kernel comes first: a group of G lanes (G = 4/8/16/32) walks one read right-to-left.  Every step
// needs the run containing the new BWT position; instead of a rank/select cascade the group issues ONE
// coalesced load of G consecutive 16-byte run records starting at a precomputed destination run
// and resolves the predecessor with a ballot + popc; the winning lane's record is broadcast with
// shuffles and becomes the state of the next step (see flat_index.h).  Reads are staged through
// shared memory in 16*G-byte chunks with a register prefetch of the next chunk; outputs are
// buffered one value per lane and written as G consecutive words.  Groups pull reads from a
// global atomic counter, so ragged read lengths balance themselves.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstring>
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
                    if (idx >= gi && idx < base + m) reinterpret_cast<uint32_t*>(p.lengths)[idx] = out32;
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
                    if (idx <= gi && idx >= base) reinterpret_cast<uint32_t*>(p.lengths)[idx] = out32;
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


// ==========================================================================================
// Thread-per-read kernels over the move rows (flat_index.h): PML and the MS pointer walk.
//
// Every lane is a small state machine that issues exactly ONE 16-byte row load per loop iteration:
//   STEP : row[k]  -> normalise (off >= len: walk to the next row), then match or decide the jump
//   JUMP2: row[k2] -> the next run headed by the read character; its (dest, delta) is the landing state
// so the 32 independent pointer chases of a warp stay in lock step at the load and their latencies
// overlap.  The only random access of a step is row[k]; k2 and the normalisation walk touch the
// neighbouring rows of the 128-byte line that access just brought in.  Lanes pull their next read from
// an atomic work counter; rare cases (non-fast symbols, saturated fields, far jump targets, pos == n)
// take the generic absolute-position step.
// ==========================================================================================
#define SPB_MV_THREADS 128
/* (prefetching the following sectors of row[k]'s line into L1 was measured: -5 % / -15 % for one / two sectors) */
#ifndef SPB_MV_MINB
#define SPB_MV_MINB 9  /* 128-thread CTAs per SM: 56 registers, no spills (at 10 the 48-register cap spills inside the loop) */
#endif
#ifndef SPB_MV_MINB_BITS
#define SPB_MV_MINB_BITS 10  /* PML with bit output: 4 KB of output staging + 4 KB of read staging, 48 registers without spills: 78.6 against
                                76.0 Gbases/s on config 2 (profiles/r2c_experiments.json, c2_stage32_10) */
#endif
#ifndef SPB_STAGE_BITS_RD
#define SPB_STAGE_BITS_RD 32
#endif

template <int LD>
__device__ __forceinline__ uint4 ld_row(const uint4* p) {
    uint4 r;
    if (LD == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <int LD>
__device__ __forceinline__ uint2 ld_row8(const uint64_t* p) {
    uint2 r;
    if (LD == 0) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    else asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

struct ThreadLoc { uint32_t k; uint4 rec; };

// thread-level predecessor search over the generic run records (generic step only)
__device__ __noinline__ ThreadLoc locate_thread(const uint4* __restrict__ runs, uint64_t R, uint64_t target, uint64_t w0) {
    uint64_t lo = w0;
    uint4 cur = ldg_nc16(runs + lo);
    // gallop, then bisect: start[lo] <= target < start[hi]
    uint64_t step = 1, hi = lo + 1;
    for (;;) {
        if (hi > R + 1) { hi = R + 1; break; }
        uint4 t = ldg_nc16(runs + hi);
        if (rec_start(t) > target) break;
        lo = hi; cur = t;
        hi = lo + step;
        step <<= 1;
    }
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        uint4 t = ldg_nc16(runs + mid);
        if (rec_start(t) <= target) { lo = mid; cur = t; } else hi = mid;
    }
    ThreadLoc out;
    out.k = (uint32_t)lo;
    out.rec = cur;
    return out;
}

// One reference step in absolute coordinates (same statements as spb_walk_kernel's body) for a single
// thread.  In: state (k, off) -- off may be un-normalised (>= len[k]) or SPB_MV_OFF_LAST --, byte c.
// Out: the normalised new state; PML length / MS sample updated in place.
struct SlowOut { uint32_t k, off, length, doc; uint64_t sample; };
template <bool MS, bool DOC>
__device__ __noinline__ SlowOut slow_step(const WalkParams& p, uint32_t c, uint32_t k, uint32_t off, uint32_t length, uint64_t sample, uint32_t doc) {
    const uint4* __restrict__ runs = p.runs;
    const uint64_t R = p.R, n = p.n;
    const uint4 ct = reinterpret_cast<const uint4*>(p.chartab)[c];
    uint64_t pos = n;
    uint4 krec = ldg_nc16(runs + (k < R ? k : R));
    if (k < R) {
        if (off == SPB_MV_OFF_LAST) pos = rec_start(ldg_nc16(runs + k + 1)) - 1;
        else {
            pos = rec_start(krec) + off;
            const ThreadLoc L = locate_thread(runs, R, pos, k);  // un-normalised offsets: find the true run
            k = L.k; krec = L.rec;
        }
    } else k = (uint32_t)R;
    uint64_t npos, w0;
    if (ct.w == 0xFFFFFFFFu) {
        length = 0;
        if (MS) sample = 0;
        if (MS && DOC) doc = p.doc_absent_ms;  // :640-642; the PML variant leaves the document alone (:303)
        npos = (uint64_t)ct.x | ((uint64_t)ct.y << 32);
        w0 = ct.z;
    } else if (k < R && rec_head(krec) == c && c < 0x80u) {
        length++;
        if (MS) sample--;
        npos = rec_lf(krec) + (pos - rec_start(krec));
        w0 = krec.z;
    } else {
        const uint32_t j = (k < R) ? crun_ordinal(p.blocks, p.blk_stride, k, c, ct.w) : ct.y;
        const uint4 cr = ldg_nc16(p.crec + (size_t)ct.x + j);
        const bool down = pos < crec_thr(cr);
        const bool same = (k < R) && rec_head(krec) == c;
        length = 0;
        if (MS) {
            const uint4 cs = ldg_nc16(p.csamp + (size_t)ct.x + j);
            if (j < ct.y) sample = (uint64_t)cs.x | ((uint64_t)cs.y << 32);
            if (down) sample = (uint64_t)cs.z | ((uint64_t)cs.w << 32);
        }
        if (DOC) {
            const uint2 cd = __ldg(p.cdoc + (size_t)ct.x + j);
            if (j < ct.y) doc = cd.x;  // start_runs_doc of the next c-run (:317 / :658)
            if (down) doc = cd.y;      // end_runs_doc of the previous c-run (:327 / :669)
        }
        const uint64_t lf = crec_lf(cr);
        npos = down ? lf - 1 : (same ? lf + (pos - rec_start(krec)) : lf);
        w0 = (down && cr.z > 0) ? cr.z - 1 : cr.z;
    }
    SlowOut o;
    o.length = length;
    o.sample = sample;
    o.doc = doc;
    if (npos >= n) {
        o.k = (uint32_t)R;
        o.off = 0;
    } else {
        const ThreadLoc L = locate_thread(runs, R, npos, w0);
        o.k = L.k;
        o.off = (uint32_t)(npos - rec_start(L.rec));
    }
    return o;
}

// ---- per-thread staging of the read and output streams in shared memory -----------------------
// A lane consumes one read byte and produces one value per step, so a 128-byte line of either stream
// stays "live" for > 100 steps of > 100 000 resident lanes: left to the L2, these half-used lines evict
// the move rows (measured: tools/gather_bench2.cu, profiles/r1_gather_bench2_streams.txt).  Each lane
// therefore moves SPB_STAGE bytes at a time: reads are fetched as whole 64-byte blocks (evict-first, no
// L1) into shared memory and consumed from there one byte per step; values are stored into shared
// memory one per step and written out as whole blocks.  Chunk j (16 bytes) of lane x lives at
// [j * SPB_MV_THREADS + x].  Both per-step accesses are branch-free (the lanes of a warp sit at unrelated
// positions of their reads, so any "every n-th step" branch would run for a few lanes almost every
// iteration); only the block refill / flush branches, once per SPB_STAGE bytes.
#ifndef SPB_STAGE
#define SPB_STAGE 64
#endif
#define SPB_STAGE_CH (SPB_STAGE / 16)
#define SPB_STAGE_PLANE (SPB_MV_THREADS * 16)  /* bytes between chunk j and chunk j+1 of a lane */

// one whole 32-byte sector per instruction (sm_100 256-bit load), evict-first in L2, not kept in L1
__device__ __forceinline__ void ld_stream32(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
// (re)fill the lane's staged block of the read buffer: the RS-aligned block containing byte gi (RS = 64, or 32 in the
// bit-output kernels: with their 4 KB of output staging that takes the CTA to 8.4 KB of shared memory and 10 CTAs per SM)
template <int RS>
__device__ __forceinline__ void stage_reads(uint4* s_rd, const uint8_t* reads, uint64_t gi) {
    const uint8_t* src = reads + (gi & ~(uint64_t)(RS - 1));
    uint4 v[RS / 16];
#pragma unroll
    for (int j = 0; j < RS / 16; j += 2) ld_stream32(src + 16 * j, v[j], v[j + 1]);
#pragma unroll
    for (int j = 0; j < RS / 16; j++) s_rd[j * SPB_MV_THREADS + threadIdx.x] = v[j];
}
// Asynchronous variant (SPB_ASYNC_READS): the lane's NEXT 32-byte block is copied global -> shared by the copy engine
// (cp.async / LDGSTS, evict-first in L2) while the lane walks the current one -- two 32-byte slots per lane, the same 64
// bytes of shared memory as the synchronous staging -- so that crossing a block boundary does not put an HBM round trip on
// the warp's critical path (any lane of the warp that refills stalls all 32).
#ifndef SPB_ASYNC_READS
#define SPB_ASYNC_READS 0
#endif
__device__ __forceinline__ void prefetch_block32(uint4* s_rd, uint32_t slot, const uint8_t* src) {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));  // not volatile: hoisted / rematerialised by the compiler
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rd + (slot * 2 + j) * SPB_MV_THREADS + threadIdx.x);
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src + 16 * j), "l"(pol) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void prefetch_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t staged_offset(uint32_t byte_in_block) {  // byte offset inside a lane's staging area
    return (byte_in_block >> 4) * SPB_STAGE_PLANE + threadIdx.x * 16 + (byte_in_block & 15u);
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& w) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
}
// Slow half of the output flush: only bytes [lo, hi) of the staged block starting at array byte sb are valid
// (the block holds the beginning or the end of a read).  EB = element size in bytes.
template <int EB>
__device__ __noinline__ void flush_partial(uint8_t* out, const uint4* s_out, uint64_t sb, uint64_t lo, uint64_t hi) {
    const uint8_t* sp = reinterpret_cast<const uint8_t*>(s_out) + threadIdx.x * 16;
    for (uint32_t j = 0; j < SPB_STAGE_CH; j++) {
        const uint64_t wb = sb + 16 * j;
        if (wb + 16 <= lo || wb >= hi) continue;
        if (wb >= lo && wb + 16 <= hi) { st_stream16(out + wb, s_out[j * SPB_MV_THREADS + threadIdx.x]); continue; }
        for (uint32_t b = 0; b < 16; b += EB) {
            if (wb + b < lo || wb + b >= hi) continue;
            const uint8_t* e = sp + j * SPB_STAGE_PLANE + b;
            if (EB == 1) out[wb + b] = *e;
            else if (EB == 2) *reinterpret_cast<uint16_t*>(out + wb + b) = *reinterpret_cast<const uint16_t*>(e);
            else if (EB == 4) *reinterpret_cast<uint32_t*>(out + wb + b) = *reinterpret_cast<const uint32_t*>(e);
            else *reinterpret_cast<uint64_t*>(out + wb + b) = *reinterpret_cast<const uint64_t*>(e);
        }
    }
}
// Stage value v of element gi (EB bytes) of the read [base, end) walked downwards; write the block out when it is complete.
template <int EB>
__device__ __forceinline__ void stage_value(uint4* s_out, void* out, uint64_t gi, uint64_t v, uint64_t end, bool last) {
    const uint64_t bo = gi * EB;
    const uint32_t so = (uint32_t)bo & (SPB_STAGE - 1);
    uint8_t* sp = reinterpret_cast<uint8_t*>(s_out) + staged_offset(so);
    if (EB == 1) *sp = (uint8_t)v;
    else if (EB == 2) *reinterpret_cast<uint16_t*>(sp) = (uint16_t)v;
    else if (EB == 4) *reinterpret_cast<uint32_t*>(sp) = (uint32_t)v;
    else *reinterpret_cast<uint64_t*>(sp) = v;
    if (so == 0 || last) {
        const uint64_t sb = bo - so, hi = end * EB;
        if (so == 0 && sb + SPB_STAGE <= hi) {
#pragma unroll
            for (int j = 0; j < SPB_STAGE_CH; j++) st_stream16(reinterpret_cast<uint8_t*>(out) + sb + 16 * j, s_out[j * SPB_MV_THREADS + threadIdx.x]);
        } else
            flush_partial<EB>(reinterpret_cast<uint8_t*>(out), s_out, sb, bo, hi < sb + SPB_STAGE ? hi : sb + SPB_STAGE);
    }
}

// Same for a read walked upwards (element gi of [base, end)): the block is complete when its last element arrives.
template <int EB>
__device__ __forceinline__ void stage_value_up(uint4* s_out, void* out, uint64_t gi, uint64_t v, uint64_t base, bool last) {
    const uint64_t bo = gi * EB;
    const uint32_t so = (uint32_t)bo & (SPB_STAGE - 1);
    uint8_t* sp = reinterpret_cast<uint8_t*>(s_out) + staged_offset(so);
    if (EB == 1) *sp = (uint8_t)v;
    else if (EB == 2) *reinterpret_cast<uint16_t*>(sp) = (uint16_t)v;
    else if (EB == 4) *reinterpret_cast<uint32_t*>(sp) = (uint32_t)v;
    else *reinterpret_cast<uint64_t*>(sp) = v;
    if (so + EB == SPB_STAGE || last) {
        const uint64_t sb = bo - so, lo = base * EB;
        if (so + EB == SPB_STAGE && sb >= lo) {
#pragma unroll
            for (int j = 0; j < SPB_STAGE_CH; j++) st_stream16(reinterpret_cast<uint8_t*>(out) + sb + 16 * j, s_out[j * SPB_MV_THREADS + threadIdx.x]);
        } else
            flush_partial<EB>(reinterpret_cast<uint8_t*>(out), s_out, sb, lo > sb ? lo : sb, bo + EB);
    }
}
// The first `need` (1..8) bytes at an arbitrary byte address, in the low bytes of the result: two aligned loads +
// funnel shift; the second word is only touched when one of the needed bytes lies in it.
__device__ __forceinline__ uint64_t ld_unaligned8(const uint8_t* p, uint32_t need) {
    const uint64_t* a = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)7);
    const uint32_t sh = ((uint32_t)reinterpret_cast<uintptr_t>(p) & 7u) * 8u;
    const uint64_t lo = __ldg(a);
    if (sh + 8u * need <= 64u) return lo >> sh;
    return (lo >> sh) | (__ldg(a + 1) << (64 - sh));
}
__device__ __forceinline__ uint64_t toupper8(uint64_t x) {  // ::toupper on 8 packed bytes ('a'..'z' only, C locale)
    const uint64_t t = x & 0x7F7F7F7F7F7F7F7Full;
    const uint64_t ge_a = t + 0x1F1F1F1F1F1F1F1Full, gt_z = t + 0x0505050505050505ull;
    const uint64_t m = ge_a & ~gt_z & ~x & 0x8080808080808080ull;
    return x - (m >> 2);
}

struct BinState {
    uint64_t bound = 0, sum_max = 0;
    uint32_t bin_max = 0, above = 0, below = 0;
};

// A decoded move row (any layout).  raw0..2 keep the words the jump fields are extracted from.
struct RowView {
    uint32_t dest, len, delta, hfid;
    bool slow;
    uint32_t raw0, raw1, raw2;
    uint2 smp;      // MSI only: {ss, sl}
    uint4 land;     // RW == 32 only: dest2[0..2], delta2[1] | delta2[2] << 16   (delta2[0] = raw2 >> 16)
};
__device__ __forceinline__ void ld_row32(const void* p, uint4& a, uint4& b, int ld) {  // one whole sector per instruction
    if (ld == 0) asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                              : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
    else asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                      : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
// MSI (MS over compact rows only): every row slot is 16 bytes, {row, u32 ss, u32 sl}: the samples come with the row load
template <int LD, int RW, bool MSI>
__device__ __forceinline__ RowView load_row(const WalkParams& p, uint32_t k) {
    RowView r;
    r.smp = make_uint2(0, 0);
    r.land = make_uint4(0, 0, 0, 0);
    if (RW == 8) {
        uint2 v;  // x = lo, y = hi (flat_index.h)
        if (MSI) {
            const uint4 q = ld_row<LD>(reinterpret_cast<const uint4*>(p.rows8) + k);
            v = make_uint2(q.x, q.y);
            r.smp = make_uint2(q.z, q.w);
        } else v = ld_row8<LD>(p.rows8 + k);
        r.dest = v.x & 0xFFFFFFu; r.len = (v.x >> 24) & 0x7Fu; r.slow = (v.x >> 31) != 0;
        r.delta = v.y & 0xFFu; r.hfid = (v.y >> 8) & 7u;
        r.raw0 = v.y; r.raw1 = 0; r.raw2 = 0;
    } else if (RW == 32) {
        uint4 a, b;
        ld_row32(reinterpret_cast<const uint4*>(p.rows32) + 2 * (size_t)k, a, b, LD);
        r.dest = a.x; r.len = a.y & SPB_MV_LEN_MASK; r.delta = a.z & SPB_MV_LEN_MASK;
        r.hfid = (a.w >> 12) & 7u; r.slow = ((a.w >> 15) & 1u) != 0;
        r.raw0 = a.y; r.raw1 = a.z; r.raw2 = a.w;
        r.land = b;
    } else {
        const uint4 q = ld_row<LD>(p.rows + k);
        r.dest = q.x; r.len = q.y & SPB_MV_LEN_MASK; r.delta = q.z & SPB_MV_LEN_MASK;
        r.hfid = (q.w >> 27) & 7u; r.slow = ((q.w >> 30) & 1u) != 0;
        r.raw0 = q.y; r.raw1 = q.z; r.raw2 = q.w;
    }
    return r;
}

// ---- one bit per base (OW == 0): bit (gi + bit_shift) of the output is set iff the PML at base gi is non-zero, i.e. iff
// the step was a match (lengths[j] = bit ? lengths[j + 1] + 1 : 0, compute_ms_pml.cpp:249-281: the bits determine every
// length).  A lane collects 32 bits in a register and parks the word in shared memory every step (branch-free, like the
// byte stores of the other widths); a whole 32-byte block (256 bases) is written with two 16-byte stores.  Blocks that
// hold the first or the last bases of a read are shared with the neighbouring reads: their words go out one by one,
// whole words as plain stores, the one or two shared words as atomicOr of the read's own bits -- the bit map must be
// zero before the launch.
#define SPB_BITS_WORDS 8
__device__ __noinline__ void flush_bits_partial(uint32_t* out, const uint32_t* s_bits, uint64_t blk, uint64_t lo, uint64_t hi) {
    for (uint32_t j = 0; j < SPB_BITS_WORDS; j++) {
        const uint64_t wlo = blk * 256 + 32 * j, whi = wlo + 32;
        if (whi <= lo || wlo >= hi) continue;
        const uint32_t v = s_bits[j * SPB_MV_THREADS + threadIdx.x];
        if (wlo >= lo && whi <= hi) { out[blk * SPB_BITS_WORDS + j] = v; continue; }
        const uint32_t a = (uint32_t)((lo > wlo ? lo : wlo) - wlo), b = (uint32_t)((hi < whi ? hi : whi) - wlo);  // bits [a, b) of the word, b - a < 32
        const uint32_t mask = ((1u << (b - a)) - 1u) << a;
        if (v & mask) atomicOr(out + blk * SPB_BITS_WORDS + j, v & mask);
    }
}
// bi = gi + bit_shift: position of the base in the bit map; [blo, bhi) = the read's bit range; acc = the lane's open word
__device__ __forceinline__ void stage_bit(uint32_t* s_bits, uint32_t* out, uint64_t bi, uint32_t bit, uint64_t blo, uint64_t bhi, uint32_t& acc, bool last) {
    acc |= bit << ((uint32_t)bi & 31u);
    s_bits[(((uint32_t)bi >> 5) & (SPB_BITS_WORDS - 1u)) * SPB_MV_THREADS + threadIdx.x] = acc;
    if (((uint32_t)bi & 31u) == 0) acc = 0;
    if (((uint32_t)bi & 255u) == 0 || last) {
        const uint64_t blk = bi >> 8, b0 = blk << 8;
        if (b0 >= blo && b0 + 256 <= bhi) {
            uint4 w0, w1;
            w0.x = s_bits[0 * SPB_MV_THREADS + threadIdx.x]; w0.y = s_bits[1 * SPB_MV_THREADS + threadIdx.x];
            w0.z = s_bits[2 * SPB_MV_THREADS + threadIdx.x]; w0.w = s_bits[3 * SPB_MV_THREADS + threadIdx.x];
            w1.x = s_bits[4 * SPB_MV_THREADS + threadIdx.x]; w1.y = s_bits[5 * SPB_MV_THREADS + threadIdx.x];
            w1.z = s_bits[6 * SPB_MV_THREADS + threadIdx.x]; w1.w = s_bits[7 * SPB_MV_THREADS + threadIdx.x];
            st_stream16(out + blk * SPB_BITS_WORDS, w0);
            st_stream16(out + blk * SPB_BITS_WORDS + 4, w1);
        } else
            flush_bits_partial(out, s_bits, blk, blo > b0 ? blo : b0, bhi < b0 + 256 ? bhi : b0 + 256);
    }
}

// One thread per read, ONE BACKWARD-SEARCH STEP PER LOOP ITERATION for every lane of the warp: load row[k] (walking
// forward while off >= len), classify the read character (match / threshold jump / generic), for a jump load the row
// of the next run headed by that character (probing consecutive rows with the compact layout), land, emit.  The
// jump's second load is predicated, not a separate iteration: an earlier version advanced one *load* per iteration
// (lanes in different modes) and ran at 14 of 32 lanes per issued instruction; the step-per-iteration form keeps
// the whole warp on the same instructions and is issue-efficient (profiles/).
template <bool MS, bool STATS, int OW, int LD, int RW, bool DOC, bool MSI = false>
__global__ void __launch_bounds__(SPB_MV_THREADS, ((STATS || DOC) && SPB_MV_MINB > 8) ? 8 : ((!MS && OW == 0) ? SPB_MV_MINB_BITS : SPB_MV_MINB)) spb_move_kernel(const __grid_constant__ WalkParams p) {
    constexpr bool ROW8 = RW == 8;
    constexpr bool BITS = !MS && OW == 0;
    __shared__ uint8_t s_fid[256];  // 0..3 fast id, 4 = present but not fast, 5 = absent
    __shared__ uint32_t s_end_k[4], s_end_off[4];
    __shared__ uint64_t s_end_sl[4];
    constexpr bool ASYNC = SPB_ASYNC_READS != 0;
    constexpr int RS = ASYNC ? 32 : (BITS ? SPB_STAGE_BITS_RD : SPB_STAGE);  // read bytes staged per lane (and per slot)
    __shared__ uint4 s_rd[(ASYNC ? 4 : RS / 16) * SPB_MV_THREADS];  // staged read bytes, see stage_reads (ASYNC: two slots)
    __shared__ uint4 s_out[BITS ? 1 : SPB_STAGE_CH * SPB_MV_THREADS];  // staged output values (lengths, or MS pointers)
    __shared__ uint32_t s_bits[BITS ? SPB_BITS_WORDS * SPB_MV_THREADS : 1];  // staged match bits (OW == 0)
    __shared__ uint4 s_doc[DOC ? SPB_STAGE_CH * SPB_MV_THREADS : 1];  // staged document numbers
    __shared__ uint32_t s_end_doc[4];
    for (int i = threadIdx.x; i < 256; i += SPB_MV_THREADS) {
        const SpbCharTab t = p.chartab[i];
        s_fid[i] = (uint8_t)(t.cidx == 0xFFFFFFFFu ? 5u : t.dest_f);
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (threadIdx.x == i) { s_end_k[i] = p.mv_end_k[i]; s_end_off[i] = p.mv_end_off[i]; s_end_sl[i] = p.mv_end_sl[i]; s_end_doc[i] = p.mv_end_doc[i]; }
    __syncthreads();

    const uint32_t R32 = (uint32_t)p.R;
    const bool upper = (p.flags & SPB_FLAG_TOUPPER) != 0;
    const bool do_stats = !MS && STATS && p.bin_width != 0 && p.stats != nullptr;
    const uint8_t* rd_bytes = reinterpret_cast<const uint8_t*>(s_rd);
    constexpr uint32_t VMAX = OW == 4 ? 0xFFFFFFFFu : (OW == 2 ? 0xFFFFu : 0xFFu);

    // lane state
    bool have = false;
    unsigned long long r = 0;
    uint64_t base = 0, end = 0, gi = 0;
    uint32_t k = 0, off = 0, length = 0;
    uint64_t sample = 0;
    uint32_t nsat = 0, doc = 0, bit_acc = 0, rslot = 0;
    BinState bs;

    for (;;) {
        if (!have) {  // ---- pull the next read (rare, divergent)
            r = atomicAdd(p.work_counter, 1ULL);
            if (r >= p.n_reads) break;
            if (p.order) r = p.order[r];  // longest reads first (the host sorts mixed-length batches): shorter tail
            base = p.offsets[r];
            const uint64_t m = p.offsets[r + 1] - base;
            if (m == 0) {
                if (do_stats) { spb_bin_stats z; z.sum_max = 0; z.above = 0; z.below = 0; p.stats[r] = z; }
                continue;
            }
            end = base + m;
            gi = end - 1;
            k = R32 - 1; off = SPB_MV_OFF_LAST; length = 0;  // empty string: pos = n-1
            if (BITS) bit_acc = 0;
            if (MS) sample = p.last_run_sample;
            if (DOC) doc = p.doc_init;
            stage_reads<RS>(s_rd, p.reads, gi);
            if (ASYNC) {  // slot 0 holds the block of gi; the block below it (if the read reaches into it) is on its way into slot 1
                rslot = 0;
                const uint64_t blk = gi & ~(uint64_t)(RS - 1);
                if (blk > base) prefetch_block32(s_rd, 1, p.reads + blk - RS);
            }
            have = true;
            if (do_stats) {
                const uint64_t bw = p.bin_width;
                const uint64_t nbins = m / bw > 0 ? m / bw : 1;
                bs.bound = base + (nbins - 1) * bw;  // lower bound of the open bin, in global index coordinates
                bs.bin_max = 0; bs.above = 0; bs.below = 0; bs.sum_max = 0;
            }
        }

        // ---- the row of the current position: walk forward while off >= len (the LF image of a run may span several runs)
        uint32_t c = rd_bytes[staged_offset(((uint32_t)gi & (RS - 1)) + (ASYNC ? rslot * 32u : 0u))];
        if (upper && c >= 'a' && c <= 'z') c -= 32;
        RowView row = load_row<LD, RW, MSI>(p, k);
        bool slow = false;
        for (;;) {
            if (row.slow) { slow = true; break; }
            if (off == SPB_MV_OFF_LAST) off = row.len - 1;
            if (off < row.len) break;
            off -= row.len;
            k++;
            row = load_row<LD, RW, MSI>(p, k);
        }
        // ---- classify: match (bwt[pos] == c, :250 / :582), threshold jump (:251-279 / :583-616), or the generic step
        const uint32_t s = s_fid[c];
        bool jump = false, endcase = false, down = false;
        uint32_t k2 = k, slot = 0;
        if (!slow) {
            if (s >= SPB_SYM_SLOW) slow = true;
            else if (s != row.hfid) {
                slot = s - (s > row.hfid ? 1u : 0u);
                uint32_t d;
                bool gen;
                if (ROW8) {
                    const uint32_t cls = (row.raw0 >> (11 + 2 * slot)) & 3u;
                    d = (row.raw0 >> (23 + 3 * slot)) & 7u;
                    gen = cls == 3;
                    down = (cls == 1) || (cls == 2 && off < ((row.raw0 >> 17) & 0x3Fu));
                } else {
                    const uint32_t v = slot == 0 ? (row.raw0 >> 20) : (slot == 1 ? (row.raw1 >> 20) : (row.raw2 & 0xFFFu));
                    d = RW == 32 ? 1u : ((row.raw2 >> (12u + 5u * slot)) & 31u);
                    gen = v == SPB_MV_THR_SLOW || (RW != 32 && d == SPB_MV_DIST_FAR);
                    down = (v == SPB_MV_THR_ALWAYS) || (off < v);
                }
                if (gen) slow = true;
                else if (d == 0) endcase = true;  // no later c-run: thr = n+1, the previous c-run is the last one
                else { jump = true; k2 = k + d; }
            }
        }
        // ---- a jump lands through the row of the next run headed by c (compact rows: row k2 or a later one: probe);
        // full rows carry the landing state themselves
        if (jump && RW != 32) {
            row = load_row<LD, RW, MSI>(p, k2);
            if (ROW8) {
                uint32_t probes = SPB_MV8_PROBES;
                while (row.hfid != s) {
                    if (--probes == 0) { slow = true; break; }
                    k2++;
                    row = load_row<LD, RW, MSI>(p, k2);
                }
            }
            if (row.slow) slow = true;  // landing state not encodable: generic step from (k, off)
        }
        // ---- land
        if (slow) {  // generic step in absolute coordinates
            const SlowOut so = slow_step<MS, DOC>(p, c, k, off, length, sample, doc);  // by value: no state lives in local memory
            k = so.k; off = so.off; length = so.length;
            if (MS) sample = so.sample;
            if (DOC) doc = so.doc;
        } else if (endcase) {
            k = s_end_k[s]; off = s_end_off[s];
            length = 0;
            if (MS) sample = s_end_sl[s];
            if (DOC) doc = s_end_doc[s];
        } else if (!jump) {
            length++;
            if (MS) sample--;
            off += row.delta;
            k = row.dest;
        } else {
            if (MS && MSI) sample = down ? row.smp.y : row.smp.x;
            else if (MS) {
                if (p.msamp_stride == 8) {  // uniform branch: compact sample table, read on jumps only
                    const uint2 sm = ld_row8<0>(reinterpret_cast<const uint64_t*>(p.msamp) + k2);
                    sample = down ? sm.y : sm.x;
                } else {
                    const uint4 sm = ldg_nc16(p.msamp + k2);
                    sample = down ? ((uint64_t)sm.z | ((uint64_t)sm.w << 32)) : ((uint64_t)sm.x | ((uint64_t)sm.y << 32));
                }
            }
            if (DOC) { const uint2 md = __ldg(p.mdoc + k2); doc = down ? md.y : md.x; }
            // first position of the c-run's LF image, or (down) one position before it
            uint32_t ldest = row.dest, ldelta = row.delta;
            if (RW == 32) {
                ldest = slot == 0 ? row.land.x : (slot == 1 ? row.land.y : row.land.z);
                ldelta = slot == 0 ? (row.raw2 >> 16) : (slot == 1 ? (row.land.w & 0xFFFFu) : (row.land.w >> 16));
            }
            k = ldest - ((down && ldelta == 0) ? 1u : 0u);
            off = ldelta - (down ? 1u : 0u);  // down, delta == 0: 0xFFFFFFFF == SPB_MV_OFF_LAST (last position of the previous row)
            length = 0;
        }

        // ---- emit
        {
            const bool last = (gi == base);
            if (DOC) stage_value<4>(s_doc, p.docs, gi, doc, end, last);
            if (MS) {
                if (p.ptr32) {  // uniform branch: 32-bit pointer output, anything that does not fit is clamped and counted
                    uint32_t v = (uint32_t)sample;
                    if (sample > 0xFFFFFFFEull) { v = 0xFFFFFFFFu; nsat++; }
                    stage_value<4>(s_out, p.pointers, gi, v, end, last);
                } else stage_value<8>(s_out, p.pointers, gi, sample, end, last);
            }
            else {
                if (BITS) stage_bit(s_bits, reinterpret_cast<uint32_t*>(p.lengths), gi + p.bit_shift, length != 0 ? 1u : 0u, base + p.bit_shift, end + p.bit_shift, bit_acc, last);
                else {
                    uint32_t v = length;
                    if (OW < 4 && v > VMAX) { v = VMAX; nsat++; }
                    stage_value<(OW ? OW : 4)>(s_out, p.lengths, gi, v, end, last);
                }
                if (do_stats) {
                    if (gi < bs.bound) {  // left the bin [bound, ...): close it
                        if (bs.bin_max >= p.max_value_thr) bs.above++; else bs.below++;
                        bs.sum_max += bs.bin_max;
                        bs.bin_max = 0;
                        bs.bound -= p.bin_width;
                    }
                    bs.bin_max = length > bs.bin_max ? length : bs.bin_max;
                }
            }
            if (last) {
                if (do_stats) {
                    if (bs.bin_max >= p.max_value_thr) bs.above++; else bs.below++;
                    bs.sum_max += bs.bin_max;
                    spb_bin_stats st; st.sum_max = bs.sum_max; st.above = bs.above; st.below = bs.below;
                    p.stats[r] = st;
                }
                have = false;
            } else {
                --gi;
                if (((uint32_t)gi & (RS - 1)) == RS - 1) {
                    if (ASYNC) {  // the block of the new gi was prefetched into the other slot; start fetching the one below it
                        prefetch_wait();
                        rslot ^= 1u;
                        const uint64_t blk = gi & ~(uint64_t)(RS - 1);
                        if (blk > base) prefetch_block32(s_rd, rslot ^ 1u, p.reads + blk - RS);
                    } else stage_reads<RS>(s_rd, p.reads, gi);
                }
            }
        }
    }
    if (nsat && p.saturated) atomicAdd(p.saturated, (unsigned long long)nsat);
}

// MS pointers -> lengths (ms_t::matching_statistics, compute_ms_pml.cpp:795-812) + bin-max classifier
// (:1150-1176); one thread per read, left to right, carrying l.  Each iteration emits one length; where the pointer
// is not the previous one + 1 it first extends l, eight characters per round (read vs text at pointer + l,
// word-wise); lengths are staged in shared memory and written as whole blocks like the walk kernel's outputs.
template <bool STATS, int OW>
__global__ void __launch_bounds__(SPB_MV_THREADS) spb_ms_lengths_kernel(const __grid_constant__ WalkParams p) {
    __shared__ uint4 s_out[SPB_STAGE_CH * SPB_MV_THREADS];
    const bool upper = (p.flags & SPB_FLAG_TOUPPER) != 0;
    const bool do_stats = STATS && p.bin_width != 0 && p.stats != nullptr;
    const uint64_t n_text = p.n_text;
    constexpr uint32_t VMAX = OW == 4 ? 0xFFFFFFFFu : (OW == 2 ? 0xFFFFu : 0xFFu);
    bool have = false, ext = false;
    unsigned long long r = 0;
    uint64_t base = 0, m = 0, i = 0, l = 0, ptr = 0, prev = 0, nbins = 1, bins_closed = 0;
    uint32_t nsat = 0;
    BinState bs;
    for (;;) {
        if (!have) {
            r = atomicAdd(p.work_counter + 1, 1ULL);
            if (r >= p.n_reads) break;
            base = p.offsets[r];
            m = p.offsets[r + 1] - base;
            if (m == 0) {
                if (do_stats) { spb_bin_stats z; z.sum_max = 0; z.above = 0; z.below = 0; p.stats[r] = z; }
                continue;
            }
            i = 0; l = 0; prev = 0;
            ptr = p.ptr32 ? (uint64_t)reinterpret_cast<const uint32_t*>(p.pointers)[base] : p.pointers[base];
            ext = true;
            have = true;
            if (do_stats) {
                const uint64_t bw = p.bin_width;
                nbins = m / bw > 0 ? m / bw : 1;
                bs.bound = bw;  // exclusive upper bound of the open bin (read coordinates)
                bins_closed = 0;
                bs.bin_max = 0; bs.above = 0; bs.below = 0; bs.sum_max = 0;
            }
        }
        if (ext) {
            // while (i + l < m && ptr + l < n && read[i + l] == text[ptr + l]) ++l;   eight characters per round; the whole
            // extension runs inside this iteration so that every lane emits one length per iteration
            for (;;) {
                const uint64_t tp = ptr + l;  // wraps like size_t (App. C-4)
                uint64_t room = (i + l < m) ? m - (i + l) : 0;
                if (tp < n_text) { if (n_text - tp < room) room = n_text - tp; } else room = 0;
                uint32_t adv = 0;
                if (room) {
                    const uint32_t need = room < 8 ? (uint32_t)room : 8u;
                    uint64_t a = ld_unaligned8(p.reads + base + i + l, need);
                    if (upper) a = toupper8(a);
                    const uint64_t x = a ^ ld_unaligned8(p.text + tp, need);
                    adv = x ? (uint32_t)(__ffsll((long long)x) - 1) >> 3 : 8u;
                    if (adv > room) adv = (uint32_t)room;
                }
                l += adv;
                if (adv < 8) break;
            }
            ext = false;
        }
        // ---- emit lengths[i] = l
        {
            uint32_t v = (uint32_t)l;
            if (OW < 4 && v > VMAX) { v = VMAX; nsat++; }
            stage_value_up<OW>(s_out, p.lengths, base + i, v, base, i == m - 1);
            if (do_stats) {
                if (i >= bs.bound && bins_closed + 1 < nbins) {
                    if (bs.bin_max >= p.max_value_thr) bs.above++; else bs.below++;
                    bs.sum_max += bs.bin_max;
                    bs.bin_max = 0;
                    bs.bound += p.bin_width;
                    bins_closed++;
                }
                bs.bin_max = (uint32_t)l > bs.bin_max ? (uint32_t)l : bs.bin_max;
            }
        }
        l = l ? l - 1 : 0;
        prev = ptr;
        if (++i == m) {
            if (do_stats) {
                if (bs.bin_max >= p.max_value_thr) bs.above++; else bs.below++;
                bs.sum_max += bs.bin_max;
                spb_bin_stats st; st.sum_max = bs.sum_max; st.above = bs.above; st.below = bs.below;
                p.stats[r] = st;
            }
            have = false;
        } else {
            ptr = p.ptr32 ? (uint64_t)reinterpret_cast<const uint32_t*>(p.pointers)[base + i] : p.pointers[base + i];
            ext = ptr != prev + 1;
        }
    }
    if (OW < 4 && nsat && p.saturated) atomicAdd(p.saturated, (unsigned long long)nsat);
}

template <typename K>
static cudaError_t launch_persistent(K kernel, const WalkParams& p, int num_sms, cudaStream_t stream) {
    int occ = 0;
    // The L1 / shared-memory split is left to the driver: the L1 serves the neighbour-row loads, and forcing the
    // maximum shared-memory carve-out halved the throughput (profiles/r1_move_kernel_variants.json).
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, SPB_MV_THREADS, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    // persistent grid: a whole number of CTAs per SM; reads are pulled from the work counter
    unsigned long long need = (p.n_reads + SPB_MV_THREADS - 1) / SPB_MV_THREADS;
    unsigned long long grid = (unsigned long long)num_sms * occ;
    if (need < grid) grid = need ? need : 1;
    if (p.persist_bytes) {
        // the move rows get the persisting part of the L2 for this launch (the read / output streams are evict-first already);
        // an opt-in experiment: SPUMONI_B200_L2_PERSIST, capi.cu
        cudaLaunchConfig_t cfg;
        std::memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(SPB_MV_THREADS);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        std::memset(attr, 0, sizeof(attr));
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(p.persist_base);
        attr[0].val.accessPolicyWindow.num_bytes = (size_t)p.persist_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = p.persist_ratio;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel, p);
    }
    kernel<<<(unsigned)grid, SPB_MV_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

template <bool STATS, int LD, int RW>
static cudaError_t launch_move_pml(const WalkParams& p, int num_sms, cudaStream_t stream) {
    switch (p.out_width) {
        case 0: return launch_persistent(spb_move_kernel<false, STATS, 0, LD, RW, false>, p, num_sms, stream);
        case 1: return launch_persistent(spb_move_kernel<false, STATS, 1, LD, RW, false>, p, num_sms, stream);
        case 2: return launch_persistent(spb_move_kernel<false, STATS, 2, LD, RW, false>, p, num_sms, stream);
        default: return launch_persistent(spb_move_kernel<false, STATS, 4, LD, RW, false>, p, num_sms, stream);
    }
}
template <int LD, int RW>
static cudaError_t launch_move(const WalkParams& p, bool ms, bool stats, int num_sms, cudaStream_t stream) {
    constexpr bool ROW8 = RW == 8;
    if (RW == 32) {  // full rows exist for PML without documents only
        if (ms || p.docs) return cudaErrorInvalidValue;
        return stats ? launch_move_pml<true, LD, 32>(p, num_sms, stream) : launch_move_pml<false, LD, 32>(p, num_sms, stream);
    }
    constexpr int RWX = RW == 32 ? 16 : RW;  // (keeps the MS / -d instantiations below out of the RW == 32 case)
    if (p.docs) {  // document numbers: u32 lengths (keeps the number of kernel variants down)
        if (p.out_width != 4) return cudaErrorInvalidValue;
        if (ms && ROW8 && p.rows8_stride == 16) return launch_persistent(spb_move_kernel<true, false, 4, 1, RWX, true, ROW8>, p, num_sms, stream);
        if (ms) return launch_persistent(spb_move_kernel<true, false, 4, 1, RWX, true>, p, num_sms, stream);
        return stats ? launch_persistent(spb_move_kernel<false, true, 4, 1, RWX, true>, p, num_sms, stream)
                     : launch_persistent(spb_move_kernel<false, false, 4, 1, RWX, true>, p, num_sms, stream);
    }
    if (ms && ROW8 && p.rows8_stride == 16) return launch_persistent(spb_move_kernel<true, false, 4, 1, RWX, false, ROW8>, p, num_sms, stream);
    if (ms) return launch_persistent(spb_move_kernel<true, false, 4, 1, RWX, false>, p, num_sms, stream);
    return stats ? launch_move_pml<true, 1, RWX>(p, num_sms, stream) : launch_move_pml<false, 1, RWX>(p, num_sms, stream);
}
template <bool STATS>
static cudaError_t launch_ms_lengths(const WalkParams& p, int num_sms, cudaStream_t stream) {
    switch (p.out_width) {
        case 1: return launch_persistent(spb_ms_lengths_kernel<STATS, 1>, p, num_sms, stream);
        case 2: return launch_persistent(spb_ms_lengths_kernel<STATS, 2>, p, num_sms, stream);
        default: return launch_persistent(spb_ms_lengths_kernel<STATS, 4>, p, num_sms, stream);
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

cudaError_t launch_walk(const WalkParams& p, bool ms, int lanes, int num_sms, cudaStream_t stream, int* launches) {
    const bool stats = p.bin_width && p.stats;
    if (lanes == 1) {
        if ((!p.rows && !p.rows8 && !p.rows32) || (ms && !p.msamp && p.rows8_stride != 16)) return cudaErrorInvalidValue;
        if (p.out_width == 0 && (ms || p.docs)) return cudaErrorInvalidValue;  // one bit per base describes PML lengths only
        cudaError_t e;
        // rows allocate in L1 (the neighbour rows of a normalisation walk / compact-row probe hit there: +12 % on config 2);
        // the full rows are one sector each and may bypass it (row_ld = 0, development knob)
        if (p.rows8) e = launch_move<1, 8>(p, ms, stats, num_sms, stream);
        else if (p.rows32) e = p.row_ld ? launch_move<1, 32>(p, ms, stats, num_sms, stream) : launch_move<0, 32>(p, ms, stats, num_sms, stream);
        else e = launch_move<1, 16>(p, ms, stats, num_sms, stream);
        if (launches) *launches = 1;
        if (ms && e == cudaSuccess) {
            e = stats ? launch_ms_lengths<true>(p, num_sms, stream) : launch_ms_lengths<false>(p, num_sms, stream);
            if (launches) *launches = 2;
        }
        return e;
    }
    if (p.out_width != 4 || p.docs || p.ptr32 || p.order) return cudaErrorInvalidValue;  // the cooperative kernels write u32 lengths / u64 pointers only, no document numbers
    if (launches) *launches = 1;
    switch (lanes) {
        case 4: return ms ? launch_t<4, true>(p, num_sms, stream) : launch_t<4, false>(p, num_sms, stream);
        case 8: return ms ? launch_t<8, true>(p, num_sms, stream) : launch_t<8, false>(p, num_sms, stream);
        case 16: return ms ? launch_t<16, true>(p, num_sms, stream) : launch_t<16, false>(p, num_sms, stream);
        case 32: return ms ? launch_t<32, true>(p, num_sms, stream) : launch_t<32, false>(p, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

// ==========================================================================================
// 2-bit packed reads (spb_pack_reads, host_util.cpp) -> one byte per base, in place of the H2D copy of the bytes:
// a quarter of the PCIe traffic; the walk kernels then read the unpacked buffer as usual (2.25 bytes of HBM traffic
// per base, noise next to the walk).  Base i of the sub-batch is packed base i + shift (a sub-batch starts anywhere in
// the caller's packed array); every thread turns one 32-bit window into 16 bytes.  Bytes that are not ACGT come from
// the caller's sorted exception list and are scattered over the unpacked buffer afterwards.
// ==========================================================================================
__global__ void __launch_bounds__(256) spb_unpack_kernel(const uint32_t* __restrict__ packed, uint32_t shift, uint64_t n_bases, uint4* __restrict__ reads) {
    const uint64_t n16 = (n_bases + 15) / 16;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n16; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w0 = packed[t], w1 = shift ? packed[t + 1] : 0u;  // bases 16 t + shift .. + 15
        const uint32_t v = __funnelshift_r(w0, w1, 2u * shift);
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t x = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) x |= ((0x54474341u >> (8u * ((v >> (8 * q + 2 * b)) & 3u))) & 0xFFu) << (8 * b);  // "ACGT"
            o[q] = x;
        }
        reads[t] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
__global__ void __launch_bounds__(256) spb_unpack_exceptions_kernel(const uint64_t* __restrict__ pos, const uint8_t* __restrict__ val, uint64_t n, uint64_t b0, uint8_t* __restrict__ reads) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) reads[pos[t] - b0] = val[t];
}
cudaError_t launch_unpack(const uint32_t* d_packed, uint32_t shift, uint64_t n_bases, uint8_t* d_reads, const uint64_t* d_exc_pos,
                          const uint8_t* d_exc_val, uint64_t n_exc, uint64_t b0, int num_sms, cudaStream_t stream, int* launches) {
    if (launches) *launches = 0;
    if (n_bases == 0) return cudaSuccess;
    const uint64_t need = ((n_bases + 15) / 16 + 255) / 256, cap = (uint64_t)num_sms * 8;
    spb_unpack_kernel<<<(unsigned)(need < cap ? need : cap), 256, 0, stream>>>(d_packed, shift, n_bases, reinterpret_cast<uint4*>(d_reads));
    if (launches) *launches = 1;
    if (n_exc) {
        const uint64_t need2 = (n_exc + 255) / 256;
        spb_unpack_exceptions_kernel<<<(unsigned)(need2 < cap ? need2 : cap), 256, 0, stream>>>(d_exc_pos, d_exc_val, n_exc, b0, d_reads);
        if (launches) *launches = 2;
    }
    return cudaGetLastError();
}

// ==========================================================================================
// K5: on-device decimal formatting (one warp per read).
// ==========================================================================================
template <int EB>
__device__ __forceinline__ uint64_t load_value(const void* v, uint64_t i) {
    if (EB == 1) return reinterpret_cast<const uint8_t*>(v)[i];
    if (EB == 2) return reinterpret_cast<const uint16_t*>(v)[i];
    if (EB == 4) return reinterpret_cast<const uint32_t*>(v)[i];
    return reinterpret_cast<const uint64_t*>(v)[i];
}
__device__ __forceinline__ uint32_t dec_digits(uint64_t v) {
    uint32_t d = 1;
    while (v >= 10) { v /= 10; d++; }
    return d;
}

template <int EB>
__global__ void __launch_bounds__(256) spb_text_sizes_kernel(const TextParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < p.n_reads; r += warps) {
        const uint64_t a = p.offsets[r], b = p.offsets[r + 1];
        uint64_t bytes = 0;
        for (uint64_t i = a + lane; i < b; i += 32) bytes += dec_digits(load_value<EB>(p.values, i)) + 1;  // digits + ' '
        for (int o = 16; o; o >>= 1) bytes += __shfl_xor_sync(0xFFFFFFFFu, bytes, o);
        if (lane == 0) p.starts[r + 1] = bytes + (p.id_offsets[r + 1] - p.id_offsets[r]) + 3;  // '>' id '\n' ... '\n'
    }
}

// exclusive scan in place: starts[0] = 0, starts[r + 1] = sum of the sizes of reads 0..r.  One block; each thread owns
// a contiguous slice (a sub-batch has at most a few hundred thousand reads).
__global__ void __launch_bounds__(1024) spb_text_scan_kernel(uint64_t* starts, uint64_t n) {
    __shared__ uint64_t part[1024];
    const uint64_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint64_t lo = (uint64_t)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    uint64_t sum = 0;
    for (uint64_t i = lo; i < hi; i++) sum += starts[i + 1];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < blockDim.x; o <<= 1) {  // Hillis-Steele over the 1024 slice sums
        const uint64_t v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint64_t run = threadIdx.x ? part[threadIdx.x - 1] : 0;
    if (threadIdx.x == 0) starts[0] = 0;
    for (uint64_t i = lo; i < hi; i++) { run += starts[i + 1]; starts[i + 1] = run; }
}

template <int EB>
__global__ void __launch_bounds__(256) spb_text_format_kernel(const TextParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < p.n_reads; r += warps) {
        const uint64_t a = p.offsets[r], b = p.offsets[r + 1];
        const uint64_t ia = p.id_offsets[r], idl = p.id_offsets[r + 1] - ia;
        char* out = p.text + p.starts[r];
        if (lane == 0) { out[0] = '>'; out[1 + idl] = '\n'; }
        for (uint64_t i = lane; i < idl; i += 32) out[1 + i] = (char)p.ids[ia + i];
        uint64_t pos = idl + 2;  // bytes of this read's text written so far (uniform across the warp)
        for (uint64_t i0 = a; i0 < b; i0 += 32) {
            const uint64_t i = i0 + lane;
            uint64_t v = i < b ? load_value<EB>(p.values, i) : 0;
            const uint32_t nd = i < b ? dec_digits(v) : 0, len = i < b ? nd + 1 : 0;
            uint32_t incl = len;  // inclusive warp scan of the text lengths
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)lane >= o) incl += t; }
            if (i < b) {
                char* q = out + pos + (incl - len);
                for (uint32_t d = nd; d-- > 0;) { q[d] = (char)('0' + v % 10); v /= 10; }
                q[nd] = ' ';
            }
            pos += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (lane == 0) out[pos] = '\n';
    }
}

static unsigned text_grid(uint64_t n_reads, int num_sms) {
    const uint64_t need = (n_reads + 7) / 8, cap = (uint64_t)num_sms * 8;  // 8 warps per 256-thread CTA
    return (unsigned)(need < cap ? (need ? need : 1) : cap);
}
cudaError_t launch_text_sizes(const TextParams& p, int num_sms, cudaStream_t stream) {
    const unsigned grid = text_grid(p.n_reads, num_sms);
    switch (p.elem_bytes) {
        case 1: spb_text_sizes_kernel<1><<<grid, 256, 0, stream>>>(p); break;
        case 2: spb_text_sizes_kernel<2><<<grid, 256, 0, stream>>>(p); break;
        case 4: spb_text_sizes_kernel<4><<<grid, 256, 0, stream>>>(p); break;
        case 8: spb_text_sizes_kernel<8><<<grid, 256, 0, stream>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    spb_text_scan_kernel<<<1, 1024, 0, stream>>>(p.starts, p.n_reads);
    return cudaGetLastError();
}
cudaError_t launch_text_format(const TextParams& p, int num_sms, cudaStream_t stream) {
    const unsigned grid = text_grid(p.n_reads, num_sms);
    switch (p.elem_bytes) {
        case 1: spb_text_format_kernel<1><<<grid, 256, 0, stream>>>(p); break;
        case 2: spb_text_format_kernel<2><<<grid, 256, 0, stream>>>(p); break;
        case 4: spb_text_format_kernel<4><<<grid, 256, 0, stream>>>(p); break;
        case 8: spb_text_format_kernel<8><<<grid, 256, 0, stream>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace spb
