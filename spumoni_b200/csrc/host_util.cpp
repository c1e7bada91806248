// Host-side helpers of the C ABI that need no GPU: the compact transports of the query calls (2-bit packed reads in,
// one bit per base out) and the read partition used for multi-GPU runs.  Declared in include/spumoni_b200.h.
//
// Why they exist: end to end the PML path is bound by the PCIe link, not by the kernels -- one
// byte per base in and one to four bytes per base out.  Both directions compress losslessly:
//   * reads: DNA is 2 bits per base; everything else (N, IUPAC codes, lower case when ::toupper is not requested) is
//     carried verbatim in a sorted exception list, so arbitrary bytes stay legal input (compute_ms_pml.cpp:916-917
//     passes whatever the FASTA holds);
//   * PML lengths: lengths[j] = match(j) ? lengths[j + 1] + 1 : 0 (compute_ms_pml.cpp:249-281), so the match bits
//     determine every length, for any read length.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/spumoni_b200.h"
#include "host_util.h"

namespace spb {

std::vector<uint64_t> shard_by_bases(const uint64_t* offsets, uint64_t n_reads, uint32_t n_shards) {
    std::vector<uint64_t> cuts((size_t)n_shards + 1, 0);
    if (n_shards == 0) return cuts;
    const uint64_t b0 = n_reads ? offsets[0] : 0, total = n_reads ? offsets[n_reads] - b0 : 0;
    for (uint32_t s = 1; s < n_shards; s++) {
        const uint64_t want = b0 + (uint64_t)((unsigned __int128)total * s / n_shards);
        // first read that starts at or after the s-th fraction of the bases
        uint64_t c = (uint64_t)(std::lower_bound(offsets, offsets + n_reads + 1, want) - offsets);
        if (c > n_reads) c = n_reads;
        cuts[s] = std::max(c, cuts[s - 1]);
    }
    cuts[n_shards] = n_reads;
    return cuts;
}

constexpr uint64_t kLanesPerSubBatch = 32768;

std::vector<Chunk> plan_chunks(const uint64_t* offsets, uint64_t r_begin, uint64_t r_end, uint64_t want_reads,
                               uint64_t min_bases, uint64_t max_bases) {
    std::vector<Chunk> out;
    uint64_t r0 = r_begin;
    while (r0 < r_end) {
        Chunk c;
        c.r0 = r0; c.b0 = offsets[r0];
        c.min_len = ~0ull; c.max_len = 0;
        uint64_t r1 = r0;
        for (;;) {
            if (offsets[r1 + 1] < offsets[r1]) throw std::invalid_argument("offsets are not non-decreasing");
            const uint64_t len = offsets[r1 + 1] - offsets[r1];
            if (len > 0xFFFFFFFFull) throw std::length_error("read longer than 2^32-1 bases");
            if (r1 > r0 && offsets[r1 + 1] - c.b0 > max_bases) break;
            c.min_len = std::min(c.min_len, len); c.max_len = std::max(c.max_len, len);
            r1++;
            if (r1 == r_end) break;
            // enough reads, enough bases -- and enough bases for the longest read seen so far: a read of m bases holds its lane
            // for m steps however empty the rest of the machine is, so a sub-batch must carry about that much work for each of
            // the lanes it can count on (a share of the ~170 000 resident ones), or its kernel is all tail
            if (r1 - r0 >= want_reads && offsets[r1] - c.b0 >= std::max<uint64_t>(min_bases, c.max_len * kLanesPerSubBatch)) break;
        }
        c.r1 = r1; c.nb = offsets[r1] - c.b0;
        out.push_back(c);
        r0 = r1;
    }
    return out;
}

bool longest_first_order(const uint64_t* offsets, const Chunk& c, uint32_t* order) {
    const uint64_t nr = c.r1 - c.r0;
    if (nr < 2 || c.max_len < 2 * std::max<uint64_t>(c.min_len, 32)) return false;
    auto cls = [](uint64_t len) -> uint32_t {  // 4 classes per octave, 0 for len < 4
        if (len < 4) return 0;
        const uint32_t e = 63u - (uint32_t)__builtin_clzll(len);
        return 4 * e + (uint32_t)((len >> (e - 2)) & 3u);
    };
    uint32_t cnt[260] = {0};
    for (uint64_t i = 0; i < nr; i++) cnt[cls(offsets[c.r0 + i + 1] - offsets[c.r0 + i])]++;
    uint32_t at[260], run = 0;
    for (int k = 259; k >= 0; k--) { at[k] = run; run += cnt[k]; }  // longest class first
    for (uint64_t i = 0; i < nr; i++) order[at[cls(offsets[c.r0 + i + 1] - offsets[c.r0 + i])]++] = (uint32_t)i;
    return true;
}

}  // namespace spb

namespace {
template <typename T>
void expand_bits_to(const uint64_t* bits, const uint64_t* offsets, uint64_t n_reads, T* out) {
    const uint64_t vmax = (uint64_t)(T)~(T)0;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < (int64_t)n_reads; r++) {
        uint64_t l = 0;
        for (uint64_t j = offsets[r + 1]; j-- > offsets[r];) {
            l = ((bits[j >> 6] >> (j & 63)) & 1u) ? l + 1 : 0;
            out[j] = (T)(l > vmax ? vmax : l);
        }
    }
}
template <typename T>
void widen_to(const T* narrow, const uint64_t* offsets, uint64_t n_reads, uint32_t* out) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < (int64_t)n_reads; r++) {
        uint32_t l = 0;
        for (uint64_t j = offsets[r + 1]; j-- > offsets[r];) {
            l = narrow[j] ? l + 1 : 0;
            out[j] = l;
        }
    }
}
}  // namespace

extern "C" {

uint64_t spb_bits_bytes(uint64_t n_bases) { return (n_bases + 255) / 256 * 32; }
uint64_t spb_packed_bytes(uint64_t n_bases) { return (n_bases + 3) / 4 + 8; }

int spb_shard_by_bases(const uint64_t* offsets, uint64_t n_reads, uint32_t n_shards, uint64_t* cuts) {
    if (!cuts || n_shards == 0 || (n_reads && !offsets)) return SPB_ERR_ARG;
    const std::vector<uint64_t> c = spb::shard_by_bases(offsets, n_reads, n_shards);
    std::memcpy(cuts, c.data(), c.size() * sizeof(uint64_t));
    return SPB_OK;
}

int spb_pack_reads(const uint8_t* reads, uint64_t n_bases, uint32_t flags, uint8_t* packed, uint64_t* exc_pos,
                   uint8_t* exc_val, uint64_t exc_capacity, uint64_t* n_exc) {
    if (!n_exc || (n_bases && (!reads || !packed))) return SPB_ERR_ARG;
    uint8_t code[256];
    std::memset(code, 0xFF, sizeof(code));
    code['A'] = 0; code['C'] = 1; code['G'] = 2; code['T'] = 3;
    if (flags & SPB_QUERY_TOUPPER) { code['a'] = 0; code['c'] = 1; code['g'] = 2; code['t'] = 3; }
    const uint64_t kBlock = 1u << 20;  // bases per work item (a multiple of 4: blocks own whole packed bytes)
    const int64_t nblk = (int64_t)((n_bases + kBlock - 1) / kBlock);
    std::vector<uint64_t> cnt((size_t)nblk + 1, 0);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t b = 0; b < nblk; b++) {
        const uint64_t lo = (uint64_t)b * kBlock, hi = std::min(n_bases, lo + kBlock);
        uint64_t c = 0;
        for (uint64_t i = lo; i < hi; i += 4) {
            uint32_t v = 0;
            const uint64_t e = std::min<uint64_t>(4, hi - i);
            for (uint64_t t = 0; t < e; t++) {
                const uint8_t k = code[reads[i + t]];
                c += k == 0xFF;
                v |= (uint32_t)(k & 3u) << (2 * t);
            }
            packed[i >> 2] = (uint8_t)v;
        }
        cnt[(size_t)b + 1] = c;
    }
    for (int64_t b = 0; b < nblk; b++) cnt[(size_t)b + 1] += cnt[(size_t)b];
    *n_exc = cnt[(size_t)nblk];
    std::memset(packed + ((n_bases + 3) >> 2), 0, 8);  // the device unpacks whole 32-bit words
    if (*n_exc > exc_capacity || (*n_exc && (!exc_pos || !exc_val))) return SPB_ERR_LIMIT;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t b = 0; b < nblk; b++) {
        if (cnt[(size_t)b + 1] == cnt[(size_t)b]) continue;
        const uint64_t lo = (uint64_t)b * kBlock, hi = std::min(n_bases, lo + kBlock);
        uint64_t w = cnt[(size_t)b];
        for (uint64_t i = lo; i < hi; i++)
            if (code[reads[i]] == 0xFF) { exc_pos[w] = i; exc_val[w] = reads[i]; w++; }
    }
    return SPB_OK;
}


int spb_pml_expand_bits(const void* bits, const uint64_t* offsets, uint64_t n_reads, uint32_t length_bytes, void* lengths) {
    if (n_reads == 0) return SPB_OK;
    if (!bits || !offsets || !lengths) return SPB_ERR_ARG;
    const uint64_t* b = static_cast<const uint64_t*>(bits);
    switch (length_bytes) {
        case 1: expand_bits_to(b, offsets, n_reads, static_cast<uint8_t*>(lengths)); return SPB_OK;
        case 2: expand_bits_to(b, offsets, n_reads, static_cast<uint16_t*>(lengths)); return SPB_OK;
        case 0: case 4: expand_bits_to(b, offsets, n_reads, static_cast<uint32_t*>(lengths)); return SPB_OK;
        default: return SPB_ERR_ARG;
    }
}

int spb_pml_widen_lengths(const void* narrow, uint32_t narrow_bytes, const uint64_t* offsets, uint64_t n_reads, uint32_t* lengths) {
    if (n_reads == 0) return SPB_OK;
    if (!narrow || !offsets || !lengths) return SPB_ERR_ARG;
    switch (narrow_bytes) {
        case 1: widen_to(static_cast<const uint8_t*>(narrow), offsets, n_reads, lengths); return SPB_OK;
        case 2: widen_to(static_cast<const uint16_t*>(narrow), offsets, n_reads, lengths); return SPB_OK;
        case 4: widen_to(static_cast<const uint32_t*>(narrow), offsets, n_reads, lengths); return SPB_OK;
        default: return SPB_ERR_ARG;
    }
}

}  // extern "C"
