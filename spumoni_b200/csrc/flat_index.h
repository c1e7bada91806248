// Flat, position-independent device image of a SPUMONI index (PML or MS).
//
// The reference answers every step of the backward search with 2-4 succinct-structure queries
// (rank/select on sd_vectors + a Huffman wavelet tree; /root/reference/src/compute_ms_pml.cpp:
// 246-285).  This layout is decoded once from the same on-disk bytes and reformulates the step
// as (verified equivalent):
//   match    : pos' = lf[k] + (pos - start[k])                       -> 0 extra loads
//   mismatch : j = #c-runs before run k;  rec = crec[c][j];
//              down = pos < rec.thr;  pos' = rec.lf - down           -> block header + crec
//   locate   : run k' containing pos' = last run with start <= pos' in a window of records that
//              begins at a precomputed destination run (dest)       -> one coalesced window load
// All multi-byte values are little-endian; positions are 40-bit (the on-disk formats use 5-byte
// integers, THRBYTES/SSABYTES, /root/reference/include/spumoni_main.hpp:63-64).
#pragma once
#include <stdint.h>

#define SPB_FLAT_MAGIC 0x3030324250535F42ULL /* "B_SPB200" */
#define SPB_FLAT_VERSION 9u
#define SPB_RUN_PAD 64u  /* sentinel + padding run records past R (start = 2^40-1) */
#define SPB_POS_MAX 0xFFFFFFFFFFULL

// 16-byte run record, one per BWT run k (AoS so a window of records is one coalesced load):
//   x = start[k] low 32          first BWT position of the run
//   y = lf[k] low 32             F[head] + #head chars before the run  (LF of the run's first char)
//   z = dest[k]                  index of the run containing BWT position lf[k]
//   w = head | start_hi<<8 | lf_hi<<16
// Record R is a sentinel with start = n (so "pos == n", SURVEY App. C-2, resolves to k = R);
// records R+1 .. R+SPB_RUN_PAD-1 have start = 2^40-1.
struct SpbRunRec { uint32_t x, y, z, w; };

// 16-byte per-character run record, for character c and ordinal j among the c-runs
// (j = 0 .. numRunsC[c]; the last one is a sentinel with thr = n+1, lf = F[c]+numC[c]):
//   x = thr low 32               thresholds[run] as a BWT position, 0 for j == 0
//   y = lf low 32                F[c] + #c before the j-th c-run
//   z = dest                     index of the run containing BWT position lf (R if lf == n)
//   w = thr_hi | lf_hi<<8
struct SpbCharRec { uint32_t x, y, z, w; };

// MS only, same indexing as SpbCharRec: SA samples re-ordered per character
//   ss = samples_start[j-th c-run]  (0 for the sentinel),  sl = samples_last[(j-1)-th c-run] (0 for j == 0)
struct SpbCharSamp { uint64_t ss, sl; };
// document numbers, same indexing as SpbCharSamp (and, per move row, as msamp)
struct SpbDocPair { uint32_t sd, ed; };

// Block header, one per 32 consecutive runs b: cnt[s] = number of runs with compact symbol s in
// runs [0, 32b); then the 32 raw head bytes of the block.  Stride = 32 + 4*sigma_pad bytes.
// (heads first so that both parts sit in the first sectors)

struct SpbCharTab {      // per byte value c; 16 bytes
    uint32_t crec_base;  // index of crec[c][0]                 | absent: F[c] low 32
    uint32_t num_runs;   // numRunsC[c]                         | absent: F[c] high
    uint32_t dest_f;     // fast symbol id 0..3, or SPB_SYM_SLOW | absent: run containing F[c] (R if F[c] == n)
    uint32_t cidx;       // compact symbol index (0..sigma-1)   | absent: 0xFFFFFFFF
};

// ---- move rows (thread-per-read kernel, PML and MS) -------------------------------------------
// Up to four "fast" symbols (the most frequent run heads below 0x80: A,C,G,T for DNA) are walked over
// a table of 16-byte rows, one per BWT run, that is the ONLY index memory a regular step touches.  The
// state of a read is (k, off): run index and offset inside the run; no absolute BWT position is carried.
//
//   x = dest                      run containing lf[k]  (lf[k] = LF of the run's first character)
//   y = len   (20) | thr0 << 20   run length                       | threshold code of jump slot 0
//   z = delta (20) | thr1 << 20   lf[k] - start[dest]              | threshold code of jump slot 1
//   w = thr2 (12) | dist0 << 12 | dist1 << 17 | dist2 << 22 | fid << 27 | slow << 30
//
//   match  (c == head[k])         (k, off) -> (dest, off + delta); the next row load normalises the state
//                                 by walking forward while off >= len (the LF image of a run can span
//                                 several runs; neighbours share the 128-byte line just fetched).
//   jump   (c != head[k]); the jump slot of fast symbol s in a run whose head has fast id h is s - (s > h):
//          thr code v: 0 never down | 1..SPB_MV_THR_MAX: down iff off < v  (v = thr - start[k]) |
//                      SPB_MV_THR_SLOW: use the generic step | SPB_MV_THR_ALWAYS: always down
//          dist d    : the next run headed by c is k2 = k + d (1..SPB_MV_DIST_FAR-1); 0 = there is none
//                      (state SpbFlatHeader::mv_end[s], always down); SPB_MV_DIST_FAR = too far, generic step
//          up   -> (dest[k2], delta[k2])                  first position of the next c-run, then LF
//          down -> that position - 1: (dest[k2], delta[k2] - 1), or (dest[k2] - 1, SPB_MV_OFF_LAST) when
//                  delta[k2] == 0     (last position of the previous c-run, then LF)
//   slow = 1: len or delta does not fit its field, or the head is not a fast symbol, or k == R (pos == n):
//          every step through this row is done by the generic absolute-position code.
// MS adds one 16-byte SpbCharSamp per row k2: ss = samples_start[k2], sl = samples_last[previous run with
// the same head] -- the two values a jump through k2 can assign to the SA sample.
#define SPB_SYM_SLOW 4u
#ifndef SPB_MV_LEN_BITS
#define SPB_MV_LEN_BITS 20      /* the test build lowers these limits to exercise the fall-backs */
#endif
#ifndef SPB_MV_THR_MAX
#define SPB_MV_THR_MAX 0xFFDu
#endif
#ifndef SPB_MV_DIST_FAR
#define SPB_MV_DIST_FAR 31u
#endif
#define SPB_MV_LEN_MASK ((1u << SPB_MV_LEN_BITS) - 1u)
#define SPB_MV_THR_SLOW 0xFFEu
#define SPB_MV_THR_ALWAYS 0xFFFu
#define SPB_MV_OFF_LAST 0xFFFFFFFFu
#define SPB_MV_ROW_PAD 16u
struct SpbMoveRow { uint32_t x, y, z, w; };

// ---- compact move rows (8 bytes per run) --------------------------------------------------------
// The walk is bound by how many DRAM lines it fetches per step, i.e. by the L2 hit rate of the row table
// (a B200 keeps ~60 MB of a randomly gathered table resident: tools/gather_bench*, profiles/).  Indexes
// with fewer than 2^24 runs therefore get rows half the size; the semantics are those of SpbMoveRow.
// Every field lies inside one 32-bit half so that decoding needs no 64-bit shifts:
//   lo:  bits  0-23 dest | 24-30 len (1..SPB_MV8_LEN_MAX) | 31 slow
//   hi:  bits  0-7  delta (0..SPB_MV8_DELTA_MAX) | 8-10 fid (SPB_SYM_SLOW for other heads; then slow = 1)
//        bits 11-16 threshold class of jump slot 0/1/2: 0 never down | 1 always down | 2 down iff off < T | 3 generic step
//        bits 17-22 T    the one in-run threshold of this row (two slots with different T: the second is class 3)
//        bits 23-31 distance of jump slot 0/1/2 (3 bits each): 0 no later c-run | 1..SPB_MV8_DIST_PROBE-1 the next
//                   c-run is row k + d | SPB_MV8_DIST_PROBE it is row k + SPB_MV8_DIST_PROBE or later: the kernel
//                   probes consecutive rows (fid == c?) from there, SPB_MV8_PROBES rows at most, then the generic step
#ifndef SPB_MV8_LEN_MAX
#define SPB_MV8_LEN_MAX 127u     /* the test build lowers these limits to exercise the fall-backs */
#endif
#ifndef SPB_MV8_DELTA_MAX
#define SPB_MV8_DELTA_MAX 255u
#endif
#ifndef SPB_MV8_T_MAX
#define SPB_MV8_T_MAX 63u
#endif
#ifndef SPB_MV8_DIST_PROBE
#define SPB_MV8_DIST_PROBE 7u
#endif
#ifndef SPB_MV8_PROBES
#define SPB_MV8_PROBES 24u
#endif
#define SPB_MV8_MAX_RUNS (1ull << 24)
#define SPB_MV8_ROW_PAD 64u      /* slow-flagged padding rows past R (probing may run past the last run) */

// ---- full move rows (32 bytes per run = one DRAM sector; PML without document numbers) ------------
// Indexes whose row table does not fit the L2 (r >= 2^24: every step is a DRAM access, and the machine sustains a
// fixed number of random DRAM accesses per second whatever their size up to a sector pair) get
// rows that hold everything a step can need, so that a threshold jump costs no second gather: next to the 16-byte
// row's match fields (dest, len, delta) and threshold codes, the LANDING state of the jump for each of the three other
// fast symbols -- (dest, delta) of the next run headed by that symbol, i.e. the LF image of its first position; a
// "down" jump lands one position before it.  "No later run of that symbol" (App. C-6) is folded into the same
// fields: threshold code ALWAYS and a landing one position past SpbFlatHeader::mv_end.
//   w0 = dest
//   w1 = len   (20) | thr0 << 20
//   w2 = delta (20) | thr1 << 20
//   w3 = thr2 (12) | fid << 12 | slow << 15 | delta2[0] << 16
//   w4..w6 = dest2[0..2]                 run holding the landing position of jump slot 0..2
//   w7 = delta2[1] | delta2[2] << 16     offset of the landing position inside that run (SPB_MV32_D2_MAX at most; a
//                                         landing that does not fit sets the slot's threshold code to SPB_MV_THR_SLOW)
//   up   -> (dest2, delta2);   down -> (dest2, delta2 - 1), or (dest2 - 1, SPB_MV_OFF_LAST) when delta2 == 0
#ifndef SPB_MV32_D2_MAX
#define SPB_MV32_D2_MAX 0xFFFFu  /* the test build lowers this limit to exercise the fall-back */
#endif
struct SpbMoveRow32 { uint32_t w[8]; };

struct SpbFlatHeader {
    uint64_t magic;
    uint32_t version;
    uint32_t mode;  // 0 PML, 1 MS
    uint64_t n;     // BWT length (text + terminator)
    uint64_t R;     // number of runs
    uint64_t n_text;           // MS: length of the text (ra.getLen(), compute_ms_pml.cpp:774)
    uint64_t last_run_sample;  // MS: get_last_run_sample() == (samples_last[R-1]+1) % n
    uint32_t sigma, sigma_pad, blk_stride, num_blocks;
    uint64_t num_crec;
    uint64_t off_runs, off_blocks, off_crec, off_csamp, off_text;  // byte offsets from blob start
    uint64_t total_bytes;
    uint64_t terminator_position;
    uint64_t off_mrows, off_msamp;  // move rows ((R + SPB_MV_ROW_PAD) * 16 bytes) and, MS only, per-row samples; 0 if no fast symbols
    uint32_t num_fast;           // number of fast symbols (0..4)
    uint32_t fast_sym[4];        // byte value of each fast id
    uint32_t mv_end_k[4];        // state of BWT position F[c] + numC[c] - 1 per fast symbol (jump target when no later c-run exists)
    uint32_t mv_end_off[4];
    uint32_t pad0[3];
    uint64_t mv_end_sl[4];       // MS: samples_last of the last c-run
    uint64_t off_mrows8;         // compact rows ((R + SPB_MV8_ROW_PAD) * 8 bytes); 0 if not applicable (then off_mrows != 0)
    // document array (-d; /root/reference/src/doc_array.cpp): present iff num_docs != 0
    uint64_t off_cdoc;           // SpbDocPair per SpbCharRec (generic step): start_runs_doc[j-th c-run], end_runs_doc[(j-1)-th c-run]
    uint64_t off_mdoc;           // SpbDocPair per move row k: start_runs_doc[k], end_runs_doc[previous run with the same head]
    uint32_t num_docs;           // 1 + largest document number
    uint32_t doc_init;           // end_runs_doc[R-1]: document of the empty match (compute_ms_pml.cpp:298, :634)
    uint32_t doc_absent_ms;      // start_runs_doc[run_of_position(0)]: MS mode, character absent from the index (:640-642)
    uint32_t mv_end_doc[4];      // end_runs_doc of the last c-run per fast symbol
    uint32_t mrows8_stride;      // 8, or 16 (MS, n < 2^32): every compact row is followed by {u32 ss, u32 sl} -- the SpbCharSamp of
                                 // that row in 32 bits -- so that a jump gets its SA sample from the row load itself (no msamp table)
    uint64_t off_mrows32;        // full 32-byte rows ((R + SPB_MV_ROW_PAD) * 32 bytes, 32-byte aligned); 0 if not built (then
                                 // off_mrows or off_mrows8 is set).  PML images without a document array only.
    uint64_t R_file;             // runs of the BWT as stored on disk (what the reference logs); R counts rows: >= R_file when
                                 // long runs were cut into several rows (index_host.cpp, balance_runs)
    uint32_t msamp_stride;       // bytes per entry of the per-row sample table at off_msamp: 16 (SpbCharSamp, u64 pairs) or 8
                                 // ({u32 ss, u32 sl}: n < 2^32, compact rows kept at 8 bytes so that the walk's own table stays in the L2)
    uint32_t pad3;               // keeps chartab 16-byte aligned (the kernels load its entries as uint4)
    uint64_t F[256];
    SpbCharTab chartab[256];
};
#ifdef __cplusplus
#include <cstddef>
static_assert(offsetof(SpbFlatHeader, chartab) % 16 == 0, "chartab entries are loaded as 16-byte vectors");
#endif
