// Launch interface between the C-ABI layer (capi.cu) and the kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/spumoni_b200.h"
#include "flat_index.h"

#define SPB_CTA_THREADS 256
#define SPB_FLAG_TOUPPER 1u

namespace spb {

struct WalkParams {
    // index (device pointers into the flat image)
    const uint4* runs;
    const uint8_t* blocks;
    const uint4* crec;
    const uint4* csamp;
    const uint8_t* text;
    const SpbCharTab* chartab;
    uint64_t n, R, n_text, last_run_sample;
    uint32_t blk_stride;
    uint32_t flags;
    // batch
    const uint8_t* reads;      // 256-byte aligned
    const uint64_t* offsets;   // n_reads + 1
    uint64_t n_reads;
    void* lengths;             // u32, or u16 / u8 (thread-per-read kernels, out_width 2 / 1)
    uint64_t* pointers;        // MS only; uint32_t[] when ptr32 (thread-per-read kernels)
    uint32_t ptr32;            // MS: write / read 32-bit pointers; values above 2^32 - 2 are clamped and counted
    spb_bin_stats* stats;      // nullable
    uint32_t bin_width;
    uint64_t max_value_thr;
    unsigned long long* work_counter;  // zeroed before launch
    // thread-per-read kernels over the move rows; rows == nullptr when the index has no fast symbols
    const uint4* rows;         // 16-byte move rows, or
    const uint64_t* rows8;     // 8-byte compact rows, or
    const void* rows32;        // 32-byte full rows (PML without documents)   (exactly one of rows / rows8 / rows32 is set)
    uint32_t rows8_stride;     // 8, or 16 when every row carries its two 32-bit SA samples (MS; then msamp == nullptr)
    const uint4* msamp;        // MS only: per-row samples, 16-byte entries (u64 ss, u64 sl) ...
    uint32_t msamp_stride;     // ... or 8-byte entries {u32 ss, u32 sl} when msamp_stride == 8
    uint32_t mv_end_k[4], mv_end_off[4];
    uint64_t mv_end_sl[4];
    uint32_t out_width;        // bytes per length value: 4, 2 or 1 (values above the maximum are clamped and counted), or
                               // 0 (PML, thread-per-read kernels): one bit per base, `lengths` is a zero-initialised bit map
    uint64_t bit_shift;        // out_width 0: base i of the batch is bit i + bit_shift of the map
    const uint32_t* order;     // nullable: the i-th read pulled from the work counter is read order[i] (thread-per-read kernels)
    // host-only launch hint (not read by the kernels): L2 access-policy window over the move rows; persist_bytes == 0 = none
    const void* persist_base;
    uint64_t persist_bytes;
    float persist_ratio;
    uint32_t row_ld;           // 0 = rows bypass L1, 1 = rows allocate in L1 (default)
    unsigned long long* saturated;  // nullable; += number of clamped length values
    // document numbers (-d), thread-per-read kernels only; docs == nullptr when not requested
    uint32_t* docs;            // one per base, like lengths
    const uint2* cdoc;         // SpbDocPair per SpbCharRec
    const uint2* mdoc;         // SpbDocPair per move row
    uint32_t doc_init, doc_absent_ms;
    uint32_t mv_end_doc[4];
};

// lanes: 4/8/16/32 = cooperative walk kernel (PML or MS); 1 = thread-per-read kernels over the move rows.
// work_counter must point at two zeroed counters.  *launches (nullable) receives the number of kernels launched.
cudaError_t launch_walk(const WalkParams& p, bool ms, int lanes, int num_sms, cudaStream_t stream, int* launches);

// 2-bit packed reads -> bytes (d_packed 4-byte aligned, readable up to the word after the last base; d_reads 16-byte
// aligned, writable up to the next multiple of 16).  Base i of the output is packed base i + shift (shift 0..3 whole bases
// plus the caller's word alignment, 0..15).  Exceptions: d_reads[pos[i] - b0] = val[i].
cudaError_t launch_unpack(const uint32_t* d_packed, uint32_t shift, uint64_t n_bases, uint8_t* d_reads, const uint64_t* d_exc_pos,
                          const uint8_t* d_exc_val, uint64_t n_exc, uint64_t b0, int num_sms, cudaStream_t stream, int* launches);

// K5: decimal formatting of a batch of per-base values on the device -- the text the reference writes with
// std::copy(..., ostream_iterator<size_t>(f, " ")) (/root/reference/src/compute_ms_pml.cpp:1007-1010):
//     '>' id '\n' v0 ' ' v1 ' ' ... '\n'       per read.
// Three launches on `stream`: exact text size per read, exclusive scan over the reads (starts[n_reads] = total bytes),
// then the bytes.  elem_bytes = size of one value (1, 2, 4 or 8).
struct TextParams {
    const void* values;        // one value per base (device)
    const uint64_t* offsets;   // n_reads + 1 (device)
    const uint8_t* ids;        // read ids back to back (device)
    const uint64_t* id_offsets;  // n_reads + 1 (device)
    uint64_t n_reads;
    uint64_t* starts;          // out: n_reads + 1 byte offsets of every read's text (device)
    char* text;                // out (device); written by launch_text_format only
    uint32_t elem_bytes;
};
cudaError_t launch_text_sizes(const TextParams& p, int num_sms, cudaStream_t stream);   // sizes + scan -> starts
cudaError_t launch_text_format(const TextParams& p, int num_sms, cudaStream_t stream);  // needs starts

}  // namespace spb
