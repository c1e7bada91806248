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
#define SPB_FLAT_VERSION 3u
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

// Block header, one per 32 consecutive runs b: cnt[s] = number of runs with compact symbol s in
// runs [0, 32b); then the 32 raw head bytes of the block.  Stride = 32 + 4*sigma_pad bytes.
// (heads first so that both parts sit in the first sectors)

struct SpbCharTab {      // per byte value c; 16 bytes
    uint32_t crec_base;  // index of crec[c][0]                 | absent: F[c] low 32
    uint32_t num_runs;   // numRunsC[c]                         | absent: F[c] high
    uint32_t dest_f;     // fast symbol id 0..3, or SPB_SYM_SLOW | absent: run containing F[c] (R if F[c] == n)
    uint32_t cidx;       // compact symbol index (0..sigma-1)   | absent: 0xFFFFFFFF
};

// ---- PML fast path (thread-per-read kernel) --------------------------------------------------
// Up to four "fast" symbols (the most frequent run heads below 0x80: A,C,G,T for DNA) get a fully
// resolved transition table so that one backward-search step is ONE 16-byte gather.  The state of a
// read is (k, off, h): run index, offset inside the run, fast id of the run's head (SPB_SYM_SLOW if
// the head is not a fast symbol).  Row k (64 bytes) = { SpbFastMatch, SpbFastJump[3] }; the jump
// slot of fast symbol s in a run whose head has fast id h is  s - (s > h).
//
//   SpbFastMatch (c == head[k]):   pos' = lf[k] + off lies in run dest, dest+1, or further
//     x = dlen   = start[dest+1] - lf[k]        off <  dlen        -> (dest,   off + delta, hd)
//     y = dest                                   off - dlen < len1  -> (dest+1, off - dlen,  hd1)
//     z = delta  = lf[k] - start[dest]           otherwise          -> scan forward from dest+2 (SpbRunLen)
//     w = len1 (26 bits, saturating) | hd << 26 | hd1 << 29
//   SpbFastJump (c != head[k], c the slot's symbol; j = #c-runs before k, (thr, lf) = crec[c][j]):
//     x = thr_rel = clamp(thr - start[k], 0, len[k])     down = off < thr_rel   (<=> pos < thr)
//     y = kA      run containing lf                        !down -> (kA, offA, hA)
//     z = offA (26) | hA << 26 | flagB << 29               down  -> (kA - flagB, offB, hB)   (position lf - 1)
//     w = offB (26) | hB << 26
//   offsets/lengths that do not fit 26 bits are stored as SPB_FAST_SAT and send the step down the
//   generic (absolute position) path.
// SpbRunLen[k] = len[k] (29 bits, saturating) | fast id of head[k] << 29, one word per run (+ padding).
#define SPB_SYM_SLOW 4u
#ifndef SPB_FAST_SAT
#define SPB_FAST_SAT 0x3FFFFFFu  /* 26-bit fields; the test build lowers these to exercise the fallbacks */
#endif
#ifndef SPB_RL_SAT
#define SPB_RL_SAT 0x1FFFFFFFu   /* 29-bit run length */
#endif
struct SpbFastRec { uint32_t x, y, z, w; };

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
    uint64_t off_fast, off_rl;   // fast rows (R * 64 bytes) and SpbRunLen words; 0 if no fast symbols
    uint32_t num_fast;           // number of fast symbols (0..4)
    uint32_t init_off;           // state of the empty string: k = R-1, off = len[R-1]-1, h = init_h
    uint32_t init_h;
    uint32_t fast_sym[4];        // byte value of each fast id
    uint32_t pad0;
    uint64_t F[256];
    SpbCharTab chartab[256];
};
