// Host-side index handling: on-disk format -> decoded arrays -> flat device image.
#pragma once
#include <sys/mman.h>

#include <cstdint>
#include <cstdlib>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "flat_index.h"

namespace spb {

struct IoError : std::runtime_error { explicit IoError(const std::string& m) : std::runtime_error(m) {} };
struct LimitError : std::runtime_error { explicit LimitError(const std::string& m) : std::runtime_error(m) {} };

// vector whose resize()/sized constructor leaves trivially-constructible elements uninitialised: value-initialising a
// 400 MB image (or an 80 MB per-run table) is half a second of serial page faults before the first useful store
// Large blocks are 2 MB-aligned and advised into transparent huge pages: first touch of a multi-GB image is otherwise
// millions of 4 KB faults (half of the flatten time of a 7.5 GB image was kernel time).
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
    static constexpr size_t kHugeFrom = size_t(4) << 20;
    template <class U> struct rebind { using other = DefaultInitAlloc<U>; };
    T* allocate(size_t n) {
        const size_t bytes = n * sizeof(T);
        if (bytes < kHugeFrom) return std::allocator<T>::allocate(n);
        void* p = nullptr;
        if (posix_memalign(&p, size_t(2) << 20, bytes) != 0) throw std::bad_alloc();
        madvise(p, bytes, MADV_HUGEPAGE);  // advisory: ignored where THP is off
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t n) {
        if (n * sizeof(T) < kHugeFrom) std::allocator<T>::deallocate(p, n);
        else std::free(p);
    }
    template <class U, class... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U;
        else ::new ((void*)p) U(std::forward<A>(a)...);
    }
};
template <class T> using RawVec = std::vector<T, DefaultInitAlloc<T>>;
using FlatBlob = RawVec<uint8_t>;

// pml_pointers / ms_pointers after decoding every succinct component into plain arrays
struct DecodedIndex {
    uint64_t n = 0, R = 0, terminator_position = 0;
    uint64_t R_file = 0;                  // runs of the BWT as stored on disk (R grows when balance_runs cuts runs into rows)
    std::vector<uint64_t> F;              // 256 entries
    RawVec<uint8_t> heads;                // R run heads (bytes <= 1 already stored as 1)
    RawVec<uint64_t> lens;                // R run lengths
    RawVec<uint64_t> thr;                 // R thresholds (BWT positions; 0 for the first run of a letter)
    RawVec<uint64_t> samples_start;       // MS: R text positions
    RawVec<uint64_t> samples_last;        // MS: R text positions
    RawVec<uint32_t> doc_start, doc_end;  // -d: R document numbers each (DocumentArray), empty otherwise
};

struct NullDb {
    uint64_t num_values = 0;
    double ks_stat_threshold = 0, mean_null_stat = 0, percentile_value = 0;
};

DecodedIndex parse_index_file(const std::string& path, bool ms);
// Cut runs into rows so that the LF image of a row overlaps few rows and no row is too long (index_host.cpp); call after
// parse_index_file / load_doc_array and before build_flat_image.
struct BalanceParams {
    uint32_t trigger = 4;   // cut a row whose LF image overlaps more than this many rows ... (0 = leave the runs alone)
    uint32_t target = 2;    // ... into pieces that overlap this many (measured on the r = 1.5e8 stand-in of config 4: 8/4 ->
                            // 30 Gbases/s with 1.15 x the rows, 4/2 -> 35 with 1.43 x, 3/1 and 2/1 -> 38-39 with 4.8-6.6 x: a target
                            // of 1 does not converge, every cut makes a new boundary; profiles/r2c_*)
    uint64_t lcap = 4093;   // and any row longer than this (SPB_MV_THR_MAX: in-run thresholds stay exact); 0 = no cap
};
BalanceParams default_balance_params();
void balance_runs(DecodedIndex& d, const BalanceParams& bp);
// The indexed text for MS mode: the sequence lines of <prefix>.fa, or -- for an index built with `spumoni build -g` -- the file
// verbatim; expected_length (n - 1, 0 = unknown) decides.
std::string load_reference_text(const std::string& fasta_path, uint64_t expected_length = 0);
NullDb load_null_db(const std::string& path);
// EmpNullDatabase constructor tail (/root/reference/src/emp_null_database.cpp:42-79) and serialize (:82-102)
NullDb summarize_null_stats(const uint32_t* stats, uint64_t n);
void save_null_db(const std::string& path, const NullDb& db, const uint32_t* stats);
// DocumentArray::load (/root/reference/src/doc_array.cpp:197-201) into d.doc_start / d.doc_end
void load_doc_array(const std::string& path, DecodedIndex& d);
FlatBlob build_flat_image(const DecodedIndex& d, bool ms, const std::string& text);
void validate_flat_image(const void* blob, uint64_t bytes);

}  // namespace spb
