// Host-side planning helpers shared by the C-ABI layer (capi.cu) and their unit tests (exported for tests as spb_plan_*).
#pragma once
#include <cstdint>
#include <vector>

namespace spb {

// Contiguous read ranges balanced by BASES, not read count (mixed-length inputs; SURVEY section 8 e): cuts[s] .. cuts[s+1]
// is the range of shard s; the ranges tile [0, n_reads).
std::vector<uint64_t> shard_by_bases(const uint64_t* offsets, uint64_t n_reads, uint32_t n_shards);

// One sub-batch of a host query call: reads [r0, r1), bases [b0, b0 + nb) of the caller's buffers.
struct Chunk {
    uint64_t r0, r1, b0, nb;
    uint64_t min_len, max_len;
};
// Cuts reads [r_begin, r_end) into sub-batches.  One read is one GPU thread, so what fills the machine is the number of
// READS in flight, not the number of bases: a sub-batch is closed once it holds `want_reads` reads and at least
// max(`min_bases`, 32 768 x its longest read) bases -- a read of m bases occupies one lane for m steps, so a sub-batch with
// long reads needs that much work per lane or its kernel is all tail -- or when the next read would take it past
// `max_bases` (device-buffer bound; a single longer read still gets a sub-batch of its own).  Throws std::invalid_argument on decreasing offsets or a read of 2^32 bases or more.
std::vector<Chunk> plan_chunks(const uint64_t* offsets, uint64_t r_begin, uint64_t r_end, uint64_t want_reads,
                               uint64_t min_bases, uint64_t max_bases);

// Longest-first processing order of the reads of a chunk (quarter-octave length classes, stable inside a class); the
// persistent kernels pull reads in this order, so the long chains start first and the tail of a launch is short reads.
// order[i] is a chunk-local read index.  Returns false (order untouched) when the lengths are uniform enough not to matter.
bool longest_first_order(const uint64_t* offsets, const Chunk& c, uint32_t* order);

}  // namespace spb
