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
    uint32_t* lengths;
    uint64_t* pointers;        // MS only
    spb_bin_stats* stats;      // nullable
    uint32_t bin_width;
    uint64_t max_value_thr;
    unsigned long long* work_counter;  // zeroed before launch
    // PML fast path (thread per read); rows == nullptr when the index has no fast symbols
    const uint4* rows;
    const uint32_t* rl;
    uint32_t init_off, init_h;
};

// lanes: 4/8/16/32 = cooperative walk kernel (PML or MS); 1 = PML fast path (thread per read)
cudaError_t launch_walk(const WalkParams& p, bool ms, int lanes, int num_sms, cudaStream_t stream);

}  // namespace spb
