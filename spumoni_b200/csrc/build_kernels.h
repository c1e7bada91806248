// Device construction of the run-length BWT, thresholds and SA samples of a text (build_kernels.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace spb {

// what Big-BWT / pfp_thresholds hand to `spumoni build` (/root/reference/src/spumoni.cpp:442-545), per run of the BWT
struct BuildParts {
    uint64_t n = 0;               // BWT length (text + terminator)
    std::vector<uint8_t> heads;   // <ref>.bwt.heads
    std::vector<uint64_t> lens;   // <ref>.bwt.len
    std::vector<uint64_t> thr;    // <ref>.thr_pos (BWT positions; 0 for the first run of a character)
    std::vector<uint64_t> ssa;    // <ref>.ssa: SA value at the first position of the run
    std::vector<uint64_t> esa;    // <ref>.esa: SA value at the last position of the run
};

// text: n bytes, the last one the terminator 0, every other byte > 1 (ms_rle_string.hpp:249-253 folds bytes <= 1).
bool build_rlbwt_device(const uint8_t* text, uint64_t n, int device, BuildParts& out, std::string& err, bool verbose);

}  // namespace spb
