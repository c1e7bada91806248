// Host side of the index: parse the reference's on-disk structures and flatten them.
//
// Reads, byte for byte, what `spumoni build` writes:
//   pml_pointers::serialize / load   /root/reference/src/compute_ms_pml.cpp:192-213, :222-229
//   ms_pointers::serialize / load    /root/reference/src/compute_ms_pml.cpp:517-542, :551-562
//   my_serialize / my_load (F)       /root/reference/include/common.hpp:458-465, :487-495
//   thr_bv::serialize / load         /root/reference/include/thresholds_ds.hpp:502-523
//   EmpNullDatabase::load            /root/reference/src/emp_null_database.cpp:104-110
// and the third-party containers inside them (ri::rle_string, ri::sparse_sd_vector,
// ri::huff_string = sdsl::wt_huff<>, sdsl::sd_vector<>, sdsl::int_vector<>).
// Nothing is queried in succinct form: every structure is decoded ONCE into plain arrays
// (heads / run lengths / thresholds / samples), cross-validated against the redundant
// information in the file, and re-encoded into the flat device image of flat_index.h.
#include "index_host.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace spb {

namespace {

struct FormatError : std::runtime_error {
    explicit FormatError(const std::string& m) : std::runtime_error(m) {}
};

class ByteReader {
public:
    ByteReader(const uint8_t* b, size_t n) : cur_(b), end_(b + n) {}
    uint64_t u64() { return scalar<uint64_t>(); }
    uint16_t u16() { return scalar<uint16_t>(); }
    uint8_t u8() { return scalar<uint8_t>(); }
    double f64() { return scalar<double>(); }
    const uint8_t* bytes(size_t n) {
        if ((size_t)(end_ - cur_) < n) throw FormatError("unexpected end of index file");
        const uint8_t* p = cur_;
        cur_ += n;
        return p;
    }
    bool at_end() const { return cur_ == end_; }
    const uint8_t* pos() const { return cur_; }
    size_t remaining() const { return (size_t)(end_ - cur_); }
private:
    template <typename T> T scalar() {
        T v;
        std::memcpy(&v, bytes(sizeof(T)), sizeof(T));
        return v;
    }
    const uint8_t* cur_;
    const uint8_t* end_;
};

// sdsl::int_vector<w>: bit length, [width byte if w == 0], ceil(bits/64) words. A view into the file buffer (sections
// are not 8-byte aligned: words are fetched with memcpy).
struct PackedInts {
    uint64_t bits = 0;
    uint32_t width = 0;
    const uint8_t* raw = nullptr;
    uint64_t count() const { return width ? bits / width : 0; }
    uint64_t num_words() const { return (bits + 63) / 64; }
    uint64_t word(uint64_t w) const {
        uint64_t v;
        std::memcpy(&v, raw + w * 8, 8);
        return v;
    }
    uint64_t at(uint64_t i) const {
        const uint64_t b = i * width;
        const uint32_t sh = (uint32_t)(b & 63);
        uint64_t v = word(b >> 6) >> sh;
        if (sh + width > 64) v |= word((b >> 6) + 1) << (64 - sh);
        return width >= 64 ? v : v & ((uint64_t(1) << width) - 1);
    }
    bool bit(uint64_t i) const { return (word(i >> 6) >> (i & 63)) & 1u; }
};

PackedInts read_packed(ByteReader& in, uint32_t fixed_width) {
    PackedInts v;
    v.bits = in.u64();
    v.width = fixed_width ? fixed_width : in.u8();
    if (v.width == 0 || v.width > 64) throw FormatError("int_vector with invalid width");
    const uint64_t nwords = (v.bits + 63) / 64;
    if (nwords > in.remaining() / 8) throw FormatError("int_vector longer than file");
    v.raw = in.bytes((size_t)nwords * 8);
    return v;
}

// slices of [0, total) for the passes that run on all host cores; the slice count is fixed, so nothing that is
// computed depends on the number of threads
constexpr int kSlices = 64;
inline uint64_t slice_lo(uint64_t total, int c) { return (uint64_t)((unsigned __int128)total * (unsigned)c / kSlices); }

// ---- deep verification of the redundant parts of the file (SPUMONI_B200_VERIFY_INDEX=1 / spb_index_verify) -----------
// The succinct containers carry acceleration structures that the loader does not need: select_support_mcl (twice per
// Elias-Fano vector, twice per wavelet tree) and rank_support_v (once per wavelet tree).  Normally only their extent is
// parsed.  In verify mode they are RE-DERIVED from the decoded bits, following the published construction of sdsl-lite
// (select_support_mcl<b,1>::init_slow + serialize: superblocks of 4096 arguments, 64-spaced mini blocks or, past
// log^4 n, long superblocks; rank_support_v<1,1>: per 512 bits an absolute count and seven packed 9-bit relative counts),
// and compared byte for byte with the file -- so an index written by a real `spumoni build` either matches this
// loader's understanding of every field or fails with the field's name.
bool g_verify = false;
std::string g_field;  // the field being parsed (for error messages)

inline uint32_t bits_hi(uint64_t x) { return x ? 63u - (uint32_t)__builtin_clzll(x) : 0u; }  // sdsl::bits::hi

struct BitWriter {  // int_vector<0> / bit_vector image: header + words
    std::string out;
    void u64(uint64_t v) { out.append((const char*)&v, 8); }
    void u8(uint8_t v) { out.push_back((char)v); }
    // int_vector<0> of `count` entries of `width` bits
    void packed(const std::vector<uint64_t>& vals, uint64_t count, uint8_t width, bool with_width) {
        const uint64_t bits = count * width;
        u64(bits);
        if (with_width) u8(width);
        std::vector<uint64_t> w((size_t)((bits + 63) / 64), 0);
        for (uint64_t i = 0; i < vals.size(); i++) {
            const uint64_t b = i * width;
            const uint32_t sh = (uint32_t)(b & 63);
            const uint64_t v = width >= 64 ? vals[i] : (vals[i] & ((uint64_t(1) << width) - 1));
            w[(size_t)(b >> 6)] |= v << sh;
            if (sh + width > 64) w[(size_t)(b >> 6) + 1] |= v >> (64 - sh);
        }
        out.append((const char*)w.data(), w.size() * 8);
    }
};

// what select_support_mcl<ones,1> serialises for the bit vector bv
std::string expected_select_support(const PackedInts& bv, bool ones) {
    BitWriter o;
    const uint64_t SB = 4096, nbits = bv.bits;
    const uint64_t capacity = ((nbits + 63) >> 6) << 6;
    const uint64_t logn = bits_hi(capacity) + 1, logn4 = logn * logn * logn * logn;
    uint64_t arg_cnt = 0;
    for (uint64_t w = 0; w < bv.num_words(); w++) {
        uint64_t x = bv.word(w);
        if (w == bv.num_words() - 1 && (nbits & 63)) x &= (uint64_t(1) << (nbits & 63)) - 1;
        arg_cnt += (uint64_t)__builtin_popcountll(x);
    }
    if (!ones) arg_cnt = nbits - arg_cnt;
    o.u64(arg_cnt);
    if (!arg_cnt) return o.out;
    const uint64_t sb = (arg_cnt + SB - 1) / SB;
    std::vector<uint64_t> superblock((size_t)sb, 0);
    struct Blk { bool is_long = false; uint8_t width = 0; std::vector<uint64_t> vals; };
    std::vector<Blk> blk((size_t)sb);
    bool any_long = false;
    std::vector<uint64_t> argpos((size_t)SB);
    uint64_t cnt = 0, sb_cnt = 0;
    for (uint64_t i = 0; i < nbits; i++) {
        if (bv.bit(i) != ones) continue;
        argpos[(size_t)(cnt % SB)] = i;
        ++cnt;
        if (cnt % SB == 0 || cnt == arg_cnt) {
            superblock[(size_t)sb_cnt] = argpos[0];
            const uint64_t last = (cnt - 1) % SB, pos_diff = argpos[(size_t)last] - argpos[0];
            Blk& b = blk[(size_t)sb_cnt];
            if (pos_diff > logn4) {
                any_long = true;
                b.is_long = true;
                b.width = (uint8_t)(bits_hi(argpos[(size_t)last]) + 1);
                b.vals.assign((size_t)SB, 0);
                for (uint64_t j = 0; j <= last; j++) b.vals[(size_t)j] = argpos[(size_t)j];
            } else {
                b.width = (uint8_t)(bits_hi(pos_diff) + 1);
                b.vals.assign(64, 0);
                for (uint64_t j = 0; j <= last; j += 64) b.vals[(size_t)(j / 64)] = argpos[(size_t)j] - argpos[0];
            }
            ++sb_cnt;
        }
    }
    o.packed(superblock, sb, (uint8_t)logn, true);
    std::vector<uint64_t> mol((size_t)(any_long ? sb : 0), 0);
    for (uint64_t i = 0; i < mol.size(); i++) mol[(size_t)i] = blk[(size_t)i].is_long ? 0 : 1;
    o.packed(mol, mol.size(), 1, false);
    for (const Blk& b : blk) o.packed(b.vals, b.vals.size(), b.width, true);
    return o.out;
}

// what rank_support_v<1,1> serialises for bv: one int_vector<64>
std::string expected_rank_support(const PackedInts& bv) {
    const uint64_t capacity = ((bv.bits + 63) >> 6) << 6, nblk = ((capacity >> 9) + 1) << 1, words = capacity >> 6;
    std::vector<uint64_t> bb((size_t)nblk, 0);
    auto pop = [&](uint64_t w) {
        uint64_t x = bv.word(w);
        if (w == bv.num_words() - 1 && (bv.bits & 63)) x &= (uint64_t(1) << (bv.bits & 63)) - 1;
        return (uint64_t)__builtin_popcountll(x);
    };
    uint64_t j = 0, sum = words ? pop(0) : 0, second = 0, i;
    for (i = 1; i < words; ++i) {
        if (!(i & 7)) {
            j += 2;
            bb[(size_t)j - 1] = second;
            bb[(size_t)j] = bb[(size_t)j - 2] + sum;
            second = sum = 0;
        } else second |= sum << (63 - 9 * (i & 7));
        sum += pop(i);
    }
    if (i & 7) {
        second |= sum << (63 - 9 * (i & 7));
        bb[(size_t)j + 1] = second;
    } else {
        j += 2;
        bb[(size_t)j - 1] = second;
        bb[(size_t)j] = bb[(size_t)j - 2] + sum;
        bb[(size_t)j + 1] = 0;
    }
    BitWriter o;
    o.packed(bb, nblk, 64, false);
    return o.out;
}

void compare_section(const uint8_t* from, const uint8_t* to, const std::string& want, const char* what) {
    const size_t got = (size_t)(to - from);
    if (got != want.size() || std::memcmp(from, want.data(), got) != 0) {
        size_t at = 0;
        while (at < got && at < want.size() && from[at] == (uint8_t)want[at]) at++;
        throw FormatError(g_field + ": " + what + " in the file differs from the one derived from the decoded bits (" + std::to_string(got) +
                          " bytes in the file, " + std::to_string(want.size()) + " expected, first difference at byte " + std::to_string(at) + ")");
    }
}

// sdsl::select_support_mcl: only its extent matters (it is an accelerator, not data) -- unless verifying
void skip_select_support(ByteReader& in, const PackedInts* bv = nullptr, bool ones = true) {
    const uint8_t* from = in.pos();
    const uint64_t args = in.u64();
    if (args != 0) {
        read_packed(in, 0);  // superblock
        read_packed(in, 1);  // mini_or_long
        const uint64_t superblocks = (args + 4095) >> 12;
        for (uint64_t i = 0; i < superblocks; i++) read_packed(in, 0);
    }
    if (g_verify && bv) compare_section(from, in.pos(), expected_select_support(*bv, ones), ones ? "select_support_mcl<1>" : "select_support_mcl<0>");
}

// ri::sparse_sd_vector -> sorted positions of the set bits
struct SparseBits {
    uint64_t universe = 0;
    RawVec<uint64_t> ones;
};
SparseBits read_sparse_bits(ByteReader& in) {
    SparseBits s;
    s.universe = in.u64();
    const uint64_t declared = in.u64();
    if (s.universe == 0) {
        if (declared) throw FormatError("sparse_sd_vector: ones in an empty universe");
        return s;
    }
    const uint64_t size = in.u64();
    const uint32_t wl = in.u8();
    PackedInts low = read_packed(in, 0);
    PackedInts high = read_packed(in, 1);
    skip_select_support(in, &high, true);
    skip_select_support(in, &high, false);
    if (size != s.universe) throw FormatError("sparse_sd_vector: size != universe");
    if (declared && low.width != wl) throw FormatError("sparse_sd_vector: low width != wl");
    // Elias-Fano: the i-th set bit of `high` at index b encodes bucket b - i. Slices of `high` decode independently once
    // the number of set bits before each slice is known.
    const uint64_t hw = high.num_words();
    auto hword = [&](uint64_t w) {
        uint64_t x = high.word(w);
        if (w == hw - 1 && (high.bits & 63)) x &= (uint64_t(1) << (high.bits & 63)) - 1;
        return x;
    };
    uint64_t before[kSlices + 1] = {0};
#pragma omp parallel for schedule(static) if (hw > 8192)
    for (int c = 0; c < kSlices; c++) {
        uint64_t cnt = 0;
        for (uint64_t w = slice_lo(hw, c), e = slice_lo(hw, c + 1); w < e; w++) cnt += (uint64_t)__builtin_popcountll(hword(w));
        before[c + 1] = cnt;
    }
    for (int c = 0; c < kSlices; c++) before[c + 1] += before[c];
    if (before[kSlices] > low.count()) throw FormatError("sparse_sd_vector: more high bits than low entries");
    if (before[kSlices] != declared) throw FormatError("sparse_sd_vector: declared count != decoded count");
    s.ones.resize((size_t)declared);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) if (hw > 8192)
    for (int c = 0; c < kSlices; c++) {
        uint64_t seen = before[c];
        bool ok = true;
        for (uint64_t w = slice_lo(hw, c), e = slice_lo(hw, c + 1); w < e; w++) {
            uint64_t word = hword(w);
            while (word) {
                const uint64_t b = w * 64 + (uint32_t)__builtin_ctzll(word);
                word &= word - 1;
                const uint64_t v = ((b - seen) << wl) | low.at(seen);
                ok &= v < s.universe && (seen == before[c] || v > s.ones[(size_t)seen - 1]);
                s.ones[(size_t)seen++] = v;
            }
        }
        bad |= !ok;
    }
    for (int c = 1; c < kSlices; c++)  // the first position of a non-empty slice against the last one before it
        if (before[c] && before[c] < before[c + 1] && s.ones[(size_t)before[c]] <= s.ones[(size_t)before[c] - 1]) bad = 1;
    if (bad) throw FormatError("sparse_sd_vector: positions not increasing / out of universe");
    return s;
}

// ri::huff_string (sdsl::wt_huff<>) -> the plain byte string
RawVec<uint8_t> read_wavelet_string(ByteReader& in) {
    const uint64_t size = in.u64();
    const uint64_t sigma = in.u64();
    PackedInts bv = read_packed(in, 1);
    {
        const uint8_t* from = in.pos();
        read_packed(in, 64);  // rank_support_v blocks
        if (g_verify) compare_section(from, in.pos(), expected_rank_support(bv), "rank_support_v");
    }
    skip_select_support(in, &bv, true);
    skip_select_support(in, &bv, false);
    const uint64_t nnodes = in.u64();
    if (nnodes > 1024) throw FormatError("wt_huff: implausible node count");
    struct Node { uint64_t bv_pos, bv_pos_rank; uint16_t parent, child0, child1; };
    std::vector<Node> nodes((size_t)nnodes);
    for (auto& nd : nodes) {
        nd.bv_pos = in.u64();
        nd.bv_pos_rank = in.u64();
        nd.parent = in.u16();
        nd.child0 = in.u16();
        nd.child1 = in.u16();
    }
    uint16_t c_to_leaf[256];
    for (int i = 0; i < 256; i++) c_to_leaf[i] = in.u16();
    in.bytes(256 * 8);  // m_path (derivable from the tree)
    const uint16_t UNDEF = 0xFFFF;
    std::vector<int> leaf_symbol((size_t)nnodes, -1);
    uint64_t leaves = 0;
    for (int c = 0; c < 256; c++)
        if (c_to_leaf[c] != UNDEF) {
            if (c_to_leaf[c] >= nnodes || nodes[c_to_leaf[c]].child0 != UNDEF) throw FormatError("wt_huff: bad c_to_leaf");
            leaf_symbol[c_to_leaf[c]] = c;
            leaves++;
        }
    if (leaves != sigma) throw FormatError("wt_huff: sigma != number of leaves");
    RawVec<uint8_t> out((size_t)size);
    if (size == 0) return out;
    if (nnodes == 0) throw FormatError("wt_huff: empty tree for non-empty string");
    // rank directory: ones before each 64-bit word
    const uint64_t nwords = (bv.bits + 63) / 64;
    std::vector<uint64_t> before((size_t)nwords + 1, 0);
    for (uint64_t w = 0; w < nwords; w++) before[w + 1] = before[w] + (uint64_t)__builtin_popcountll(bv.word(w));
    auto rank1 = [&](uint64_t i) {
        uint64_t r = before[i >> 6];
        if (i & 63) r += (uint64_t)__builtin_popcountll(bv.word(i >> 6) & ((uint64_t(1) << (i & 63)) - 1));
        return r;
    };
    std::vector<uint64_t> node_rank((size_t)nnodes, 0), node_len((size_t)nnodes, 0);
    for (size_t v = 0; v < nodes.size(); v++)
        if (nodes[v].child0 != UNDEF) {
            if (nodes[v].bv_pos > bv.bits) throw FormatError("wt_huff: node outside bit vector");
            node_rank[v] = rank1(nodes[v].bv_pos);
        }
    bool bad = false;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)size; i++) {
        uint32_t v = 0;
        uint64_t idx = (uint64_t)i;
        int guard = 0;
        while (nodes[v].child0 != UNDEF) {
            const uint64_t at = nodes[v].bv_pos + idx;
            if (at >= bv.bits || ++guard > 300) { bad = true; break; }
            const uint64_t ones = rank1(at) - node_rank[v];
            if (bv.bit(at)) { idx = ones; v = nodes[v].child1; }
            else { idx -= ones; v = nodes[v].child0; }
            if (v >= nodes.size()) { bad = true; break; }
        }
        if (bad) continue;
        if (leaf_symbol[v] < 0) { bad = true; continue; }
        out[(size_t)i] = (uint8_t)leaf_symbol[v];
    }
    if (bad) throw FormatError("wt_huff: inconsistent tree / bit vector");
    return out;
}

RawVec<uint8_t> slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw IoError("cannot open " + path);
    const std::streamoff sz = f.tellg();
    f.seekg(0);
    RawVec<uint8_t> b((size_t)sz);
    if (sz && !f.read((char*)b.data(), sz)) throw IoError("cannot read " + path);
    return b;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// ------------------------------------------------------------------------------------------
DecodedIndex parse_index_file(const std::string& path, bool ms) {
    RawVec<uint8_t> buf = slurp(path);
    ByteReader in(buf.data(), buf.size());
    DecodedIndex d;
    g_verify = std::getenv("SPUMONI_B200_VERIFY_INDEX") != nullptr;
    auto letter = [](int c) { char b[16]; std::snprintf(b, sizeof(b), c > 32 && c < 127 ? "['%c']" : "[0x%02x]", c); return std::string(b); };
    try {
        g_field = "terminator_position / F";
        d.terminator_position = in.u64();
        const uint64_t fsize = in.u64();
        if (fsize != 256) throw FormatError("F does not have 256 entries");
        d.F.resize(256);
        for (auto& f : d.F) f = in.u64();
        // ri::rle_string
        d.n = in.u64();
        d.R = in.u64();
        const uint64_t B = in.u64();
        if (d.n == 0 || d.R == 0 || B == 0) throw FormatError("empty BWT");
        g_field = "bwt.runs (ri::sparse_sd_vector)";
        SparseBits runs = read_sparse_bits(in);
        std::vector<SparseBits> per_letter(256);
        for (int c = 0; c < 256; c++) { g_field = "bwt.runs_per_letter" + letter(c) + " (ri::sparse_sd_vector)"; per_letter[c] = read_sparse_bits(in); }
        g_field = "bwt.run_heads (ri::huff_string = sdsl::wt_huff<>)";
        d.heads = read_wavelet_string(in);
        g_field = "run-length BWT";
        if (d.heads.size() != d.R) throw FormatError("run_heads.size() != R");
        const uint64_t R = d.R;
        // ordinal of every run among the runs of its letter: per-slice letter counts, then an exclusive scan
        std::vector<uint64_t> ord0((size_t)(kSlices + 1) * 256, 0);
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < kSlices; c++) {
            uint64_t* cnt = &ord0[(size_t)(c + 1) * 256];
            for (uint64_t k = slice_lo(R, c), e = slice_lo(R, c + 1); k < e; k++) cnt[d.heads[(size_t)k]]++;
        }
        for (int c = 0; c < kSlices; c++)
            for (int a = 0; a < 256; a++) ord0[(size_t)(c + 1) * 256 + a] += ord0[(size_t)c * 256 + a];
        const uint64_t* runs_of = &ord0[(size_t)kSlices * 256];  // runs per letter
        for (int a = 0; a < 256; a++) {
            if (runs_of[a] > per_letter[a].ones.size()) throw FormatError("more runs of a letter than runs_per_letter bits");
            if (runs_of[a] != per_letter[a].ones.size()) throw FormatError("runs_per_letter has unused bits");
            if (per_letter[a].ones.size() && per_letter[a].ones.back() + 1 != per_letter[a].universe)
                throw FormatError("runs_per_letter universe != letter count");
        }
        // run lengths: the j-th c-run ends at c-local index per_letter[c].ones[j]
        d.lens.resize((size_t)R);
        std::vector<uint64_t> slice_pos((size_t)kSlices + 1, 0), slice_cnt((size_t)kSlices * 256, 0);
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < kSlices; c++) {
            uint64_t ord[256];
            std::memcpy(ord, &ord0[(size_t)c * 256], sizeof(ord));
            uint64_t* cnt = &slice_cnt[(size_t)c * 256];
            uint64_t tot = 0;
            for (uint64_t k = slice_lo(R, c), e = slice_lo(R, c + 1); k < e; k++) {
                const uint8_t a = d.heads[(size_t)k];
                const RawVec<uint64_t>& ones = per_letter[a].ones;
                const uint64_t j = ord[a]++;
                const uint64_t len = j ? ones[(size_t)j] - ones[(size_t)j - 1] : ones[0] + 1;
                d.lens[(size_t)k] = len;
                cnt[a] += len;
                tot += len;
            }
            slice_pos[(size_t)c + 1] = tot;
        }
        for (int c = 0; c < kSlices; c++) slice_pos[(size_t)c + 1] += slice_pos[(size_t)c];
        if (slice_pos[(size_t)kSlices] != d.n) throw FormatError("sum of run lengths != n");
        {   // `runs` marks the last position of every B-th run (ms_rle_string.hpp:255-256)
            if (runs.universe != d.n) throw FormatError("runs bit vector universe != n");
            if (runs.ones.size() < R / B) throw FormatError("runs bit vector disagrees with run lengths");
            if (runs.ones.size() > R / B) throw FormatError("runs bit vector has extra bits");
            int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : bad)
            for (int c = 0; c < kSlices; c++) {
                uint64_t pos = slice_pos[(size_t)c];
                for (uint64_t k = slice_lo(R, c), e = slice_lo(R, c + 1); k < e; k++) {
                    pos += d.lens[(size_t)k];
                    if (k % B == B - 1) bad |= runs.ones[(size_t)(k / B)] != pos - 1;
                }
            }
            if (bad) throw FormatError("runs bit vector disagrees with run lengths");
        }
        {   // F must equal build_F_ of the decoded runs (compute_ms_pml.cpp:119-147)
            std::vector<uint64_t> cnt(256, 0), F(256, 0);
            for (int c = 0; c < kSlices; c++)
                for (int a = 0; a < 256; a++) cnt[a] += slice_cnt[(size_t)c * 256 + a];
            uint64_t acc = 0;
            for (int c = 0; c < 256; c++) { F[c] = acc; acc += cnt[c]; }
            if (F != d.F) throw FormatError("F array disagrees with the run-length BWT");
        }
        auto unpack = [&](const PackedInts& v, RawVec<uint64_t>& out) {
            out.resize((size_t)R);
#pragma omp parallel for schedule(static)
            for (int64_t k = 0; k < (int64_t)R; k++) out[(size_t)k] = v.at((uint64_t)k);
        };
        if (ms) {
            g_field = "samples_last (sdsl::int_vector<>)";
            PackedInts sl = read_packed(in, 0);
            if (sl.count() != R) throw FormatError("samples_last has != R entries");
            unpack(sl, d.samples_last);
        }
        // thr_bv: thresholds of the j-th c-run (j >= 1) is the (j-1)-th bit of thresholds_per_letter[c]
        {
            std::vector<SparseBits> thr(256);
            for (int c = 0; c < 256; c++) { g_field = "thresholds_per_letter" + letter(c) + " (ri::sparse_sd_vector)"; thr[c] = read_sparse_bits(in); }
            g_field = "thresholds";
            for (int a = 0; a < 256; a++) {
                const uint64_t want = runs_of[a] ? runs_of[a] - 1 : 0;
                if (thr[a].ones.size() < want) throw FormatError("fewer thresholds than runs for a letter");
                if (thr[a].ones.size() != want) throw FormatError("threshold count != runs-1 for a letter");
            }
            d.thr.resize((size_t)R);
#pragma omp parallel for schedule(dynamic, 1)
            for (int c = 0; c < kSlices; c++) {
                uint64_t ord[256];
                std::memcpy(ord, &ord0[(size_t)c * 256], sizeof(ord));
                for (uint64_t k = slice_lo(R, c), e = slice_lo(R, c + 1); k < e; k++) {
                    const uint8_t a = d.heads[(size_t)k];
                    const uint64_t j = ord[a]++;
                    d.thr[(size_t)k] = j ? thr[a].ones[(size_t)j - 1] : 0;
                }
            }
        }
        if (ms) {
            g_field = "samples_start (sdsl::int_vector<>)";
            PackedInts ss = read_packed(in, 0);
            if (ss.count() != R) throw FormatError("samples_start has != R entries");
            unpack(ss, d.samples_start);
        }
        if (!in.at_end()) throw FormatError("trailing bytes after the index");
    } catch (FormatError& e) {
        const std::string msg = e.what();
        throw FormatError(path + ": " + (msg.compare(0, g_field.size(), g_field) == 0 ? msg : "[" + g_field + "] " + msg));
    }
    return d;
}

// The indexed text: every line of <prefix>.fa that is not a '>' header, trailing white space stripped, concatenated.
// The file is cut into slices at line starts; each slice is measured, then copied to its final place.
std::string load_reference_text(const std::string& fasta_path, uint64_t expected_length) {
    const RawVec<uint8_t> buf = slurp(fasta_path);
    // `spumoni build -g` indexes the bytes of a general-text file as they are (/root/reference/src/spumoni.cpp:589-591: the file
    // itself is the parse input) and `run -M -g` reads them back through the SLP; a FASTA reference is indexed without its header
    // lines and line ends (refbuilder.cpp:349-379).  The BWT length tells the two apart.
    auto fasta_like = [&]() { for (size_t i = 0; i < buf.size(); i++) if (buf[i] == '>' && (i == 0 || buf[i - 1] == '\n')) return true; return false; };
    if (expected_length && buf.size() == expected_length && !fasta_like()) return std::string((const char*)buf.data(), buf.size());
    const uint8_t* b = buf.data();
    const size_t N = buf.size();
    size_t lo[kSlices + 1];
    lo[0] = 0;
    for (int c = 1; c < kSlices; c++) {
        const size_t p = std::max<size_t>((size_t)slice_lo(N, c), lo[c - 1]);
        const void* nl = p == 0 ? b : (p - 1 < N ? std::memchr(b + p - 1, '\n', N - (p - 1)) : nullptr);
        lo[c] = p == 0 ? 0 : (nl ? (size_t)((const uint8_t*)nl - b) + 1 : N);
    }
    lo[kSlices] = N;
    // calls emit(first, length) for the kept part of every line in [from, to)
    auto for_each_line = [&](size_t from, size_t to, auto&& emit) {
        for (size_t s0 = from; s0 < to;) {
            const void* nl = std::memchr(b + s0, '\n', to - s0);
            size_t e = nl ? (size_t)((const uint8_t*)nl - b) : to;
            const size_t next = e + 1;
            if (!(e > s0 && b[s0] == '>')) {
                while (e > s0 && std::isspace(b[e - 1])) e--;
                if (e > s0) emit(s0, e - s0);
            }
            s0 = next;
        }
    };
    size_t at[kSlices + 1] = {0};
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < kSlices; c++) {
        size_t kept = 0;
        for_each_line(lo[c], lo[c + 1], [&](size_t, size_t len) { kept += len; });
        at[c + 1] = kept;
    }
    for (int c = 0; c < kSlices; c++) at[c + 1] += at[c];
    if (expected_length && at[kSlices] != expected_length && buf.size() == expected_length)  // general text that happens to hold '>' lines
        return std::string((const char*)buf.data(), buf.size());
    std::string text(at[kSlices], '\0');
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < kSlices; c++) {
        char* dst = &text[0] + at[c];
        for_each_line(lo[c], lo[c + 1], [&](size_t first, size_t len) { std::memcpy(dst, b + first, len); dst += len; });
    }
    return text;
}

NullDb load_null_db(const std::string& path) {
    RawVec<uint8_t> buf = slurp(path);
    ByteReader in(buf.data(), buf.size());
    NullDb db;
    try {
        db.num_values = in.u64();
        db.ks_stat_threshold = in.f64();
        db.mean_null_stat = in.f64();
        db.percentile_value = in.f64();
        PackedInts v = read_packed(in, 0);
        if (v.count() != db.num_values) throw FormatError("null_stats length != num_values");
    } catch (FormatError& e) {
        throw FormatError(path + ": " + e.what());
    }
    return db;
}

NullDb summarize_null_stats(const uint32_t* stats, uint64_t n) {
    NullDb db;
    db.num_values = n;
    double sum = 0.0;
    for (uint64_t i = 0; i < n; i++) sum += stats[i];
    db.mean_null_stat = sum / (double)n;  // 0/0 for an empty input, like the reference
    std::vector<uint32_t> sorted(stats, stats + n);
    std::sort(sorted.begin(), sorted.end());
    uint64_t largest = 0, cur = n ? sorted[0] : 0, occ = 0;
    for (uint32_t x : sorted) {
        if (x == cur) occ++;
        else {
            if (occ >= 5) largest = cur;
            cur = x;
            occ = 1;
        }
    }
    if (occ >= 5) largest = cur;
    db.percentile_value = (double)largest;
    return db;
}

void save_null_db(const std::string& path, const NullDb& db, const uint32_t* stats) {
    uint32_t mx = 0;
    for (uint64_t i = 0; i < db.num_values; i++) mx = std::max(mx, stats[i]);
    // max(ceil(log2(max)), 1): a max that is an exact power of two does not fit its own width (reference quirk)
    uint32_t width = 1;
    if (mx > 1) {
        uint32_t c = 0;
        while ((1ull << c) < mx) c++;
        width = c < 1 ? 1 : c;
    }
    const uint64_t bits = db.num_values * width, mask = width >= 64 ? ~0ull : ((1ull << width) - 1);
    std::vector<uint64_t> words((size_t)((bits + 63) / 64), 0);
    for (uint64_t i = 0; i < db.num_values; i++) {
        const uint64_t v = stats[i] & mask, b = i * width;
        words[(size_t)(b >> 6)] |= v << (b & 63);
        if ((b & 63) + width > 64) words[(size_t)(b >> 6) + 1] |= v >> (64 - (b & 63));
    }
    std::ofstream o(path, std::ios::binary);
    if (!o) throw IoError("cannot write " + path);
    o.write((const char*)&db.num_values, 8);
    o.write((const char*)&db.ks_stat_threshold, 8);
    o.write((const char*)&db.mean_null_stat, 8);
    o.write((const char*)&db.percentile_value, 8);
    o.write((const char*)&bits, 8);
    const uint8_t w8 = (uint8_t)width;
    o.write((const char*)&w8, 1);
    o.write((const char*)words.data(), (std::streamsize)(words.size() * 8));
    if (!o) throw IoError("write failed: " + path);
}

void load_doc_array(const std::string& path, DecodedIndex& d) {
    RawVec<uint8_t> buf = slurp(path);
    ByteReader in(buf.data(), buf.size());
    try {
        const uint64_t num_entries = in.u64();
        PackedInts sd = read_packed(in, 0), ed = read_packed(in, 0);
        if (in.remaining()) throw FormatError("trailing bytes");
        if (num_entries != d.R || sd.count() != d.R || ed.count() != d.R) throw FormatError("document array does not have one entry per BWT run");
        d.doc_start.resize((size_t)d.R);
        d.doc_end.resize((size_t)d.R);
        int too_big = 0;
#pragma omp parallel for schedule(static) reduction(| : too_big)
        for (int64_t k = 0; k < (int64_t)d.R; k++) {
            const uint64_t a = sd.at((uint64_t)k), b = ed.at((uint64_t)k);
            too_big |= a >= 0xFFFFFFFFull || b >= 0xFFFFFFFFull;
            d.doc_start[(size_t)k] = (uint32_t)a;
            d.doc_end[(size_t)k] = (uint32_t)b;
        }
        if (too_big) throw LimitError("document numbers >= 2^32-1 are not supported");
    } catch (FormatError& e) {
        throw FormatError(path + ": " + e.what());
    }
}

// ------------------------------------------------------------------------------------------
// Run splitting ("balancing", Nishimoto & Tabei's move structure).  The walk normalises a state (k, off) with
// off >= len[k] by stepping to the following rows, i.e. a match step costs one row per destination run that the LF image
// of its run overlaps before the landing position.  On average that is < 2 rows, but a long run whose image covers many
// short runs makes every step through it a long, divergent walk (one lane of the warp working, hundreds of rows) --
// and long runs are where the positions are.  A run is therefore cut into sub-runs ("rows") with the same head until
//   * the LF image of every row overlaps at most `trigger` rows (cut every `target` boundaries), and
//   * no row is longer than `lcap` (in-run thresholds and offsets then always fit their fields).
// A sub-run inherits threshold, SA samples and documents of its run: the query semantics do not change (the first
// sub-run of a run is "the next run of its letter" for everything before it, the last one "the previous run" for
// everything after it; inside a run only the match step and the signed-char quirk of App. C-1 look at the row, and
// both read per-run values).  Cutting adds boundaries, so it is iterated to a fixed point (geometric: a fresh piece
// overlaps `target` rows and is cut again only beyond `trigger`), eight passes at most -- the bound is a performance
// property, the kernels' loop is correct for any walk length.
void balance_runs(DecodedIndex& d, const BalanceParams& bp) {
    d.R_file = d.R;
    if (bp.trigger == 0 || d.R == 0) return;
    const uint32_t trigger = bp.trigger, target = std::max<uint32_t>(1, std::min(bp.target, bp.trigger));
    // the length cap is soft: it may add at most max(R / 2, 2^20) rows (an index of a few runs of 10^10 characters is not
    // turned into millions of rows; its rows then keep the generic threshold code where an offset does not fit)
    uint64_t lcap = bp.lcap ? bp.lcap : ~0ull;
    if (bp.lcap) lcap = std::max<uint64_t>(lcap, d.n / std::max<uint64_t>(d.R / 2, 1ull << 20) + 1);
    const bool ms = !d.samples_start.empty(), doc = !d.doc_start.empty();
    for (int pass = 0; pass < 8; pass++) {
        const uint64_t R = d.R;
        RawVec<uint64_t> start((size_t)R + 1), lf((size_t)R);
        {
            std::vector<uint64_t> sliceC((size_t)(kSlices + 1) * 256, 0), slicePos((size_t)kSlices + 1, 0);
#pragma omp parallel for schedule(dynamic, 1)
            for (int c = 0; c < kSlices; c++) {
                uint64_t* cc = &sliceC[(size_t)(c + 1) * 256];
                uint64_t tot = 0;
                for (uint64_t k = slice_lo(R, c), e = slice_lo(R, c + 1); k < e; k++) { cc[d.heads[(size_t)k]] += d.lens[(size_t)k]; tot += d.lens[(size_t)k]; }
                slicePos[(size_t)c + 1] = tot;
            }
            for (int c = 0; c < kSlices; c++) {
                slicePos[(size_t)c + 1] += slicePos[(size_t)c];
                for (int a = 0; a < 256; a++) sliceC[(size_t)(c + 1) * 256 + a] += sliceC[(size_t)c * 256 + a];
            }
#pragma omp parallel for schedule(dynamic, 1)
            for (int c = 0; c < kSlices; c++) {
                uint64_t seen[256];
                for (int a = 0; a < 256; a++) seen[a] = d.F[a] + sliceC[(size_t)c * 256 + a];
                uint64_t pos = slicePos[(size_t)c];
                for (uint64_t k = slice_lo(R, c), e = slice_lo(R, c + 1); k < e; k++) {
                    start[(size_t)k] = pos;
                    lf[(size_t)k] = seen[d.heads[(size_t)k]];
                    seen[d.heads[(size_t)k]] += d.lens[(size_t)k];
                    pos += d.lens[(size_t)k];
                }
            }
            start[(size_t)R] = d.n;
        }
        // cut offsets per row, slice by slice: (row, offset inside the row), ascending
        std::vector<std::vector<std::pair<uint64_t, uint64_t>>> cuts(kSlices);
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < kSlices; c++) {
            uint64_t cur[256];
            for (int a = 0; a < 256; a++) cur[a] = ~0ull;
            std::vector<uint64_t> offs;
            auto& out = cuts[(size_t)c];
            for (uint64_t k = slice_lo(R, c), e = slice_lo(R, c + 1); k < e; k++) {
                const uint8_t a = d.heads[(size_t)k];
                const uint64_t p = lf[(size_t)k], len = d.lens[(size_t)k];
                uint64_t i = cur[a];
                if (i == ~0ull) i = (uint64_t)(std::upper_bound(start.begin(), start.end(), p) - start.begin()) - 1;
                else while (start[(size_t)i + 1] <= p) i++;
                cur[a] = i;
                if (len == 1) continue;
                offs.clear();
                // boundaries of the row table strictly inside the image (p, p + len)
                uint64_t j = i + 1, inside = 0;
                while (j <= R && start[(size_t)j] < p + len) { inside++; j++; }
                if (inside + 1 > trigger)
                    for (uint64_t t = target; t <= inside; t += target) offs.push_back(start[(size_t)(i + t)] - p);
                if (len > lcap) {
                    const uint64_t pieces = (len + lcap - 1) / lcap;
                    for (uint64_t q = 1; q < pieces; q++) offs.push_back((uint64_t)((unsigned __int128)len * q / pieces));
                    std::sort(offs.begin(), offs.end());
                    offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
                }
                for (uint64_t o : offs) if (o > 0 && o < len) out.emplace_back(k, o);
            }
        }
        uint64_t total_cuts = 0;
        std::vector<uint64_t> cut_before((size_t)kSlices + 1, 0);
        for (int c = 0; c < kSlices; c++) { total_cuts += cuts[(size_t)c].size(); cut_before[(size_t)c + 1] = total_cuts; }
        if (total_cuts == 0) break;
        if (R + total_cuts >= 0xFFFFFFFFull - SPB_RUN_PAD) break;  // would leave the 32-bit row index space: keep what we have
        const uint64_t R2 = R + total_cuts;
        RawVec<uint8_t> heads2((size_t)R2);
        RawVec<uint64_t> lens2((size_t)R2), thr2((size_t)R2), ss2(ms ? (size_t)R2 : 0), sl2(ms ? (size_t)R2 : 0);
        RawVec<uint32_t> ds2(doc ? (size_t)R2 : 0), de2(doc ? (size_t)R2 : 0);
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < kSlices; c++) {
            const auto& cs = cuts[(size_t)c];
            size_t ci = 0;
            uint64_t w = slice_lo(R, c) + cut_before[(size_t)c];
            for (uint64_t k = slice_lo(R, c), e = slice_lo(R, c + 1); k < e; k++) {
                uint64_t prev = 0;
                for (;;) {
                    const bool more = ci < cs.size() && cs[ci].first == k;
                    const uint64_t upto = more ? cs[ci].second : d.lens[(size_t)k];
                    heads2[(size_t)w] = d.heads[(size_t)k];
                    lens2[(size_t)w] = upto - prev;
                    thr2[(size_t)w] = d.thr[(size_t)k];
                    if (ms) { ss2[(size_t)w] = d.samples_start[(size_t)k]; sl2[(size_t)w] = d.samples_last[(size_t)k]; }
                    if (doc) { ds2[(size_t)w] = d.doc_start[(size_t)k]; de2[(size_t)w] = d.doc_end[(size_t)k]; }
                    w++;
                    prev = upto;
                    if (!more) break;
                    ci++;
                }
            }
        }
        d.heads.swap(heads2); d.lens.swap(lens2); d.thr.swap(thr2);
        if (ms) { d.samples_start.swap(ss2); d.samples_last.swap(sl2); }
        if (doc) { d.doc_start.swap(ds2); d.doc_end.swap(de2); }
        d.R = R2;
    }
}

BalanceParams default_balance_params() {
    BalanceParams bp;  // SPUMONI_B200_BALANCE="trigger,target,lcap" ("0" disables)
    if (const char* e = std::getenv("SPUMONI_B200_BALANCE")) {
        unsigned long long a = 0, b = 0, c = 0;
        const int got = std::sscanf(e, "%llu,%llu,%llu", &a, &b, &c);
        if (got >= 1) bp.trigger = (uint32_t)a;
        if (got >= 2) bp.target = (uint32_t)b;
        if (got >= 3) bp.lcap = c;
    }
    return bp;
}

// ------------------------------------------------------------------------------------------
static inline uint32_t pack_hi(uint64_t a, int shift) { return (uint32_t)((a >> 32) & 0xFF) << shift; }

FlatBlob build_flat_image(const DecodedIndex& d, bool ms, const std::string& text) {
    const uint64_t n = d.n, R = d.R;
    if (n >= SPB_POS_MAX - 1) throw LimitError("BWT length >= 2^40-2 is not supported");
    if (R >= 0xFFFFFFFFull - SPB_RUN_PAD) throw LimitError("more than 2^32-64 runs are not supported");
    if (ms && text.size() + 1 != n) throw FormatError("reference text length + 1 != BWT length (is <prefix>.fa the indexed text?)");

    // per-run tables, in kSlices contiguous slices of the run range so that every O(R) pass below runs on all host cores
    const int kChunks = kSlices;
    auto chunk_lo = [&](int c) -> uint64_t { return slice_lo(R, c); };
    RawVec<uint64_t> start((size_t)R + 1), lf((size_t)R);
    std::vector<uint64_t> numC(256, 0);
    std::vector<uint32_t> numRunsC(256, 0);
    std::vector<uint32_t> ordBefore((size_t)(kChunks + 1) * 256, 0);  // runs of each letter before a slice
    {
        // pass 1: per-slice length sums per head; exclusive scan; pass 2: fill start / lf from the slice's prefix
        std::vector<uint64_t> sliceC((size_t)(kChunks + 1) * 256, 0), slicePos((size_t)kChunks + 1, 0);
        std::vector<uint32_t> sliceRuns((size_t)kChunks * 256, 0);
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < kChunks; c++) {
            uint64_t* cc = &sliceC[(size_t)(c + 1) * 256];
            uint32_t* cr = &sliceRuns[(size_t)c * 256];
            uint64_t tot = 0;
            for (uint64_t k = chunk_lo(c), e = chunk_lo(c + 1); k < e; k++) {
                const uint64_t l = d.lens[(size_t)k];
                cc[d.heads[(size_t)k]] += l;
                cr[d.heads[(size_t)k]]++;
                tot += l;
            }
            slicePos[(size_t)c + 1] = tot;
        }
        for (int c = 0; c < kChunks; c++) {
            slicePos[(size_t)c + 1] += slicePos[(size_t)c];
            for (int a = 0; a < 256; a++) {
                sliceC[(size_t)(c + 1) * 256 + a] += sliceC[(size_t)c * 256 + a];
                numRunsC[a] += sliceRuns[(size_t)c * 256 + a];
                ordBefore[(size_t)(c + 1) * 256 + a] = numRunsC[a];
            }
        }
        for (int a = 0; a < 256; a++) numC[a] = sliceC[(size_t)kChunks * 256 + a];
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < kChunks; c++) {
            uint64_t seen[256];
            for (int a = 0; a < 256; a++) seen[a] = d.F[a] + sliceC[(size_t)c * 256 + a];
            uint64_t pos = slicePos[(size_t)c];
            for (uint64_t k = chunk_lo(c), e = chunk_lo(c + 1); k < e; k++) {
                start[(size_t)k] = pos;
                const uint8_t a = d.heads[(size_t)k];
                lf[(size_t)k] = seen[a];
                seen[a] += d.lens[(size_t)k];
                pos += d.lens[(size_t)k];
            }
        }
        start[(size_t)R] = n;
    }
    auto run_of = [&](uint64_t p) -> uint64_t {  // run containing BWT position p (R if p == n)
        return (uint64_t)(std::upper_bound(start.begin(), start.end(), p) - start.begin()) - 1;
    };
    // dest[k] = run holding LF(start of run k). Within one head LF is increasing in k, so after one binary search per
    // (slice, head) the destination only ever moves forward.
    RawVec<uint32_t> dest((size_t)R);
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < kChunks; c++) {
        uint64_t cur[256];
        for (int a = 0; a < 256; a++) cur[a] = ~0ull;
        for (uint64_t k = chunk_lo(c), e = chunk_lo(c + 1); k < e; k++) {
            const uint8_t a = d.heads[(size_t)k];
            const uint64_t p = lf[(size_t)k];
            uint64_t i = cur[a];
            if (i == ~0ull) i = run_of(p);
            else while (start[(size_t)i + 1] <= p) i++;
            cur[a] = i;
            dest[(size_t)k] = (uint32_t)i;
        }
    }

    // compact alphabet
    uint32_t sigma = 0;
    uint32_t cidx[256];
    for (int c = 0; c < 256; c++) cidx[c] = numRunsC[c] ? sigma++ : 0xFFFFFFFFu;
    const uint32_t sigma_pad = (uint32_t)align_up(sigma ? sigma : 1, 8);
    const uint32_t blk_stride = 32 + 4 * sigma_pad;
    const uint32_t num_blocks = (uint32_t)((R + 31) / 32 + 1);
    uint64_t num_crec = 0;
    uint32_t crec_base[256];
    for (int c = 0; c < 256; c++) {
        crec_base[c] = (uint32_t)num_crec;
        if (numRunsC[c]) num_crec += (uint64_t)numRunsC[c] + 1;
    }
    if (num_crec >= 0xFFFFFFFFull) throw LimitError("per-character record table exceeds 2^32 entries");

    // layout
    SpbFlatHeader h;
    std::memset(&h, 0, sizeof(h));
    h.magic = SPB_FLAT_MAGIC;
    h.version = SPB_FLAT_VERSION;
    h.mode = ms ? 1 : 0;
    h.n = n;
    h.R = R;
    h.R_file = d.R_file ? d.R_file : R;
    h.n_text = ms ? text.size() : 0;
    h.last_run_sample = ms ? (d.samples_last[(size_t)R - 1] + 1) % n : 0;
    h.sigma = sigma;
    h.sigma_pad = sigma_pad;
    h.blk_stride = blk_stride;
    h.num_blocks = num_blocks;
    h.num_crec = num_crec;
    h.terminator_position = d.terminator_position;
    size_t off = align_up(sizeof(SpbFlatHeader), 256);
    h.off_runs = off;
    off = align_up(off + ((size_t)R + SPB_RUN_PAD) * sizeof(SpbRunRec), 256);
    h.off_blocks = off;
    off = align_up(off + (size_t)num_blocks * blk_stride, 256);
    h.off_crec = off;
    off = align_up(off + ((size_t)num_crec + 1) * sizeof(SpbCharRec), 256);
    if (ms) {
        h.off_csamp = off;
        off = align_up(off + ((size_t)num_crec + 1) * sizeof(SpbCharSamp), 256);
        h.off_text = off;
        off = align_up(off + text.size() + 64, 256);
    }
    // fast symbols: up to four most frequent run heads with 1 < byte < 0x80 (bytes >= 0x80 never
    // "match" in the reference, App. C-1; the terminator never occurs in a read that matters)
    uint32_t fid[256];
    for (int c = 0; c < 256; c++) fid[c] = SPB_SYM_SLOW;
    uint32_t num_fast = 0;
    {
        std::vector<int> cand;
        for (int c = 2; c < 0x80; c++) if (numRunsC[c]) cand.push_back(c);
        std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return numRunsC[a] > numRunsC[b]; });
        if (cand.size() > 4) cand.resize(4);
        std::sort(cand.begin(), cand.end());
        for (int c : cand) { h.fast_sym[num_fast] = (uint32_t)c; fid[c] = num_fast++; }
        int too_long = 0;  // 32-bit in-run offsets would overflow: generic kernels only
#pragma omp parallel for schedule(static) reduction(| : too_long)
        for (int64_t k = 0; k < (int64_t)R; k++) too_long |= d.lens[(size_t)k] >= 0xFFFFFFFEull;
        if (too_long) num_fast = 0;
        if (!num_fast) for (int c = 0; c < 256; c++) fid[c] = SPB_SYM_SLOW;
    }
    h.num_fast = num_fast;
    // row width: 8-byte rows whenever run indices fit 24 bits (the table then has a chance to live in the L2); otherwise
    // 32-byte full rows for PML without documents (one DRAM sector per step, jumps included) and 16-byte rows for MS / -d
    // (their jumps read the samples / documents of the target row anyway).  SPUMONI_B200_ROWS=16|32 forces a layout (tests).
    const bool use_doc = !d.doc_start.empty();
    bool rows8 = num_fast && R + SPB_MV8_ROW_PAD < SPB_MV8_MAX_RUNS;
    bool rows32 = num_fast && !rows8 && !ms && !use_doc;
    if (const char* e = std::getenv("SPUMONI_B200_ROWS")) {
        if (std::atoi(e) == 16) { rows8 = false; rows32 = false; }
        if (std::atoi(e) == 32 && num_fast && !ms && !use_doc) { rows8 = false; rows32 = true; }
    }
    if (num_fast) {
        // MS over compact rows: with 32-bit text positions the two samples of a row ride in the same 16 bytes
        // (SPUMONI_B200_MSAMP = inline (default) | split32: 8-byte rows + a separate table of {u32 ss, u32 sl} read on jumps only |
        //  wide: 8-byte rows + the 16-byte SpbCharSamp table, what indexes with n >= 2^32 get)
        bool ms_inline = ms && rows8 && n < (1ull << 32);
        bool ms_split32 = false;
        if (const char* e = std::getenv("SPUMONI_B200_MSAMP")) {
            if (std::string(e) == "wide") ms_inline = false;
            if (std::string(e) == "split32" && ms_inline) { ms_inline = false; ms_split32 = true; }
        }
        h.msamp_stride = ms ? (ms_split32 ? 8u : 16u) : 0u;
        h.mrows8_stride = rows8 ? (ms_inline ? 16u : 8u) : 0u;
        if (rows8) {
            h.off_mrows8 = off;
            off = align_up(off + ((size_t)R + SPB_MV8_ROW_PAD) * h.mrows8_stride, 256);
        } else if (rows32) {
            h.off_mrows32 = off;
            off = align_up(off + ((size_t)R + SPB_MV_ROW_PAD) * sizeof(SpbMoveRow32), 256);
        } else {
            h.off_mrows = off;
            off = align_up(off + ((size_t)R + SPB_MV_ROW_PAD) * sizeof(SpbMoveRow), 256);
        }
        if (ms && !ms_inline) {
            h.off_msamp = off;
            off = align_up(off + ((size_t)R + SPB_MV8_ROW_PAD) * h.msamp_stride, 256);
        }
    }
    if (use_doc) {
        h.off_cdoc = off;
        off = align_up(off + ((size_t)num_crec + 1) * sizeof(SpbDocPair), 256);
        if (num_fast) {
            h.off_mdoc = off;
            off = align_up(off + ((size_t)R + SPB_MV8_ROW_PAD) * sizeof(SpbDocPair), 256);
        }
        uint32_t mx = 0;
#pragma omp parallel for schedule(static) reduction(max : mx)
        for (int64_t k = 0; k < (int64_t)R; k++) mx = std::max(mx, std::max(d.doc_start[(size_t)k], d.doc_end[(size_t)k]));
        h.num_docs = mx + 1;
        h.doc_init = d.doc_end[(size_t)R - 1];
        h.doc_absent_ms = d.doc_start[0];
    }
    h.total_bytes = off;
    for (int c = 0; c < 256; c++) {
        h.F[c] = d.F[c];
        SpbCharTab& t = h.chartab[c];
        if (numRunsC[c]) {
            t.crec_base = crec_base[c];
            t.num_runs = numRunsC[c];
            t.dest_f = fid[c];
            t.cidx = cidx[c];
        } else {
            t.crec_base = (uint32_t)(d.F[c] & 0xFFFFFFFFu);
            t.num_runs = (uint32_t)(d.F[c] >> 32);
            t.dest_f = (uint32_t)run_of(d.F[c]);
            t.cidx = 0xFFFFFFFFu;
        }
    }

    FlatBlob blob(h.total_bytes);
    {  // zero-fill (padding bytes are part of the image) with every core taking its share of the page faults
        const int64_t pieces = (int64_t)((h.total_bytes + (4u << 20) - 1) / (4u << 20));
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < pieces; i++) {
            const size_t a = (size_t)i * (4u << 20);
            std::memset(blob.data() + a, 0, std::min<size_t>(4u << 20, h.total_bytes - a));
        }
    }
    std::memcpy(blob.data(), &h, sizeof(h));
    SpbRunRec* runs = reinterpret_cast<SpbRunRec*>(blob.data() + h.off_runs);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)R; k++) {
        SpbRunRec r;
        r.x = (uint32_t)start[(size_t)k];
        r.y = (uint32_t)lf[(size_t)k];
        r.z = dest[(size_t)k];
        r.w = (uint32_t)d.heads[(size_t)k] | pack_hi(start[(size_t)k], 8) | pack_hi(lf[(size_t)k], 16);
        runs[k] = r;
    }
    for (uint64_t k = R; k < R + SPB_RUN_PAD; k++) {
        const uint64_t s = k == R ? n : SPB_POS_MAX;
        SpbRunRec r;
        r.x = (uint32_t)s;
        r.y = 0;
        r.z = (uint32_t)R;
        r.w = pack_hi(s, 8);  // head 0: never equals a present character
        runs[k] = r;
    }
    // block headers: the heads of 32 runs + the number of runs of every letter before the block
    {
        auto blk_lo = [&](int c) -> uint32_t { return (uint32_t)slice_lo(num_blocks, c); };
        std::vector<uint32_t> before((size_t)(kChunks + 1) * sigma_pad, 0);
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < kChunks; c++) {
            uint32_t* cnt = &before[(size_t)(c + 1) * sigma_pad];
            for (uint64_t k = (uint64_t)blk_lo(c) * 32, e = std::min<uint64_t>((uint64_t)blk_lo(c + 1) * 32, R); k < e; k++) cnt[cidx[d.heads[(size_t)k]]]++;
        }
        for (int c = 0; c < kChunks; c++)
            for (uint32_t a = 0; a < sigma_pad; a++) before[(size_t)(c + 1) * sigma_pad + a] += before[(size_t)c * sigma_pad + a];
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < kChunks; c++) {
            std::vector<uint32_t> cnt(before.begin() + (size_t)c * sigma_pad, before.begin() + (size_t)(c + 1) * sigma_pad);
            for (uint32_t b = blk_lo(c); b < blk_lo(c + 1); b++) {
                uint8_t* hdr = blob.data() + h.off_blocks + (size_t)b * blk_stride;
                std::memcpy(hdr + 32, cnt.data(), 4 * (size_t)sigma_pad);
                for (uint32_t t = 0; t < 32; t++) {
                    const uint64_t k = (uint64_t)b * 32 + t;
                    if (k < R) {
                        hdr[t] = d.heads[(size_t)k];
                        cnt[cidx[d.heads[(size_t)k]]]++;
                    }
                }
            }
        }
    }
    // per-character records (+ samples)
    {
        SpbCharRec* crec = reinterpret_cast<SpbCharRec*>(blob.data() + h.off_crec);
        SpbCharSamp* csamp = ms ? reinterpret_cast<SpbCharSamp*>(blob.data() + h.off_csamp) : nullptr;
        SpbDocPair* cdoc = use_doc ? reinterpret_cast<SpbDocPair*>(blob.data() + h.off_cdoc) : nullptr;
        // The "last sample / last document of the previous run of the letter" fields of a record belong to the run before it:
        // every run writes them into the NEXT record of its letter (the first record keeps the zero of the image, the
        // sentinel after the last run receives the last run's), which makes the slices independent.
        int bad_thr = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : bad_thr)
        for (int sl = 0; sl < kChunks; sl++) {
            uint32_t o[256];
            std::memcpy(o, &ordBefore[(size_t)sl * 256], sizeof(o));
            for (uint64_t k = chunk_lo(sl), e = chunk_lo(sl + 1); k < e; k++) {
                const uint8_t c = d.heads[(size_t)k];
                const size_t at = (size_t)crec_base[c] + o[c]++;
                const uint64_t thr = d.thr[(size_t)k];
                bad_thr |= thr > n;
                SpbCharRec r;
                r.x = (uint32_t)thr;
                r.y = (uint32_t)lf[(size_t)k];
                r.z = dest[(size_t)k];
                r.w = pack_hi(thr, 0) | pack_hi(lf[(size_t)k], 8);
                crec[at] = r;
                if (ms) {
                    csamp[at].ss = d.samples_start[(size_t)k];
                    csamp[at + 1].sl = d.samples_last[(size_t)k];
                }
                if (use_doc) {
                    cdoc[at].sd = d.doc_start[(size_t)k];
                    cdoc[at + 1].ed = d.doc_end[(size_t)k];
                }
            }
        }
        if (bad_thr) throw FormatError("threshold beyond the BWT");
        for (int c = 0; c < 256; c++) {
            if (!numRunsC[c]) continue;
            const size_t at = (size_t)crec_base[c] + numRunsC[c];
            const uint64_t end = d.F[c] + numC[c];
            SpbCharRec r;
            r.x = (uint32_t)(n + 1);
            r.y = (uint32_t)end;
            r.z = (uint32_t)run_of(end);
            r.w = pack_hi(n + 1, 0) | pack_hi(end, 8);
            crec[at] = r;
            if (ms) csamp[at].ss = 0;    // .sl was written by the last run of the letter
            if (use_doc) cdoc[at].sd = 0;  // .ed likewise
        }
    }
    if (ms) std::memcpy(blob.data() + h.off_text, text.data(), text.size());
    if (num_fast) {
        SpbMoveRow* rows = (rows8 || rows32) ? nullptr : reinterpret_cast<SpbMoveRow*>(blob.data() + h.off_mrows);
        SpbMoveRow32* rows_f = rows32 ? reinterpret_cast<SpbMoveRow32*>(blob.data() + h.off_mrows32) : nullptr;
        // state of the last BWT position of every fast symbol's F-interval: where a jump lands when no later run of the symbol exists
        uint32_t end_k[4] = {0, 0, 0, 0};
        uint64_t end_off[4] = {0, 0, 0, 0};
        for (uint32_t s = 0; s < num_fast; s++) {
            const int c = (int)h.fast_sym[s];
            const uint64_t p = d.F[c] + numC[c] - 1;  // numC[c] >= 1
            const uint64_t kp = run_of(p);
            end_k[s] = (uint32_t)kp;
            end_off[s] = p - start[(size_t)kp];
        }
        uint64_t* rows_c = rows8 ? reinterpret_cast<uint64_t*>(blob.data() + h.off_mrows8) : nullptr;
        SpbCharSamp* msamp = (ms && h.off_msamp && h.msamp_stride == 16) ? reinterpret_cast<SpbCharSamp*>(blob.data() + h.off_msamp) : nullptr;
        uint64_t* msamp32 = (ms && h.off_msamp && h.msamp_stride == 8) ? reinterpret_cast<uint64_t*>(blob.data() + h.off_msamp) : nullptr;
        SpbFlatHeader* hd = reinterpret_cast<SpbFlatHeader*>(blob.data());
        // next run (strictly after k) headed by each fast symbol: the state at every slice boundary comes from the first fast
        // runs of the slices after it, then each slice walks backwards on its own
        std::vector<uint32_t> edge((size_t)kChunks * 4, 0xFFFFFFFFu), first((size_t)kChunks * 4, 0xFFFFFFFFu);
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < kChunks; c++) {  // first run of each fast symbol inside slice c (stops as soon as all are seen)
            uint32_t found = 0;
            for (uint64_t k = chunk_lo(c), e = chunk_lo(c + 1); k < e && found < num_fast; k++) {
                const uint32_t f = fid[d.heads[(size_t)k]];
                if (f < SPB_SYM_SLOW && first[(size_t)c * 4 + f] == 0xFFFFFFFFu) { first[(size_t)c * 4 + f] = (uint32_t)k; found++; }
            }
        }
        for (int c = kChunks - 2; c >= 0; c--)  // edge[c] = next fast runs at or after the end of slice c
            for (int f = 0; f < 4; f++)
                edge[(size_t)c * 4 + f] = first[(size_t)(c + 1) * 4 + f] != 0xFFFFFFFFu ? first[(size_t)(c + 1) * 4 + f] : edge[(size_t)(c + 1) * 4 + f];
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < kChunks; c++) {
        uint32_t nxt[4];  // next runs strictly after k
        for (int s = 0; s < 4; s++) nxt[s] = edge[(size_t)c * 4 + s];
        for (uint64_t k = chunk_lo(c + 1), lo = chunk_lo(c); k-- > lo;) {
            struct Advance {  // after row k is built (any exit path), run k becomes the next run of its own head
                uint32_t* nxt; uint32_t f; uint32_t k;
                ~Advance() { if (f < SPB_SYM_SLOW) nxt[f] = k; }
            } advance{nxt, fid[d.heads[(size_t)k]], (uint32_t)k};
            const uint32_t hk = fid[d.heads[(size_t)k]];
            const uint64_t len_k = d.lens[(size_t)k], st = start[(size_t)k];
            const uint64_t de = dest[(size_t)k], delta = lf[(size_t)k] - start[(size_t)de];
            bool slow = hk >= SPB_SYM_SLOW || len_k > SPB_MV_LEN_MASK || delta > SPB_MV_LEN_MASK;
            uint32_t thr3[3] = {0, 0, 0}, dist3[3] = {0, 0, 0};
            if (hk < SPB_SYM_SLOW) {
                for (uint32_t s = 0; s < num_fast; s++) {
                    if (s == hk) continue;
                    const uint32_t slot = s - (s > hk ? 1u : 0u);
                    const uint32_t k2 = nxt[s];
                    if (k2 == 0xFFFFFFFFu) {  // no later c-run: thr = n+1 (App. C-6), the jump always goes down
                        thr3[slot] = SPB_MV_THR_ALWAYS;
                        dist3[slot] = 0;
                        continue;
                    }
                    const uint64_t thr = d.thr[k2];
                    if (thr <= st) thr3[slot] = 0;
                    else if (thr >= st + len_k) thr3[slot] = SPB_MV_THR_ALWAYS;
                    else thr3[slot] = (thr - st) <= SPB_MV_THR_MAX ? (uint32_t)(thr - st) : SPB_MV_THR_SLOW;
                    const uint64_t dd = (uint64_t)k2 - k;
                    dist3[slot] = dd < SPB_MV_DIST_FAR ? (uint32_t)dd : SPB_MV_DIST_FAR;
                }
            }
            if (rows8) {
                slow = hk >= SPB_SYM_SLOW || len_k > SPB_MV8_LEN_MAX || delta > SPB_MV8_DELTA_MAX;
                uint64_t cls = 0, dist = 0, T = 0;
                for (uint32_t s = 0; s < num_fast && hk < SPB_SYM_SLOW; s++) {
                    if (s == hk) continue;
                    const uint32_t slot = s - (s > hk ? 1u : 0u);
                    const uint32_t v = thr3[slot];  // same derivation as the wide rows: 0 | value | SLOW | ALWAYS
                    uint64_t c3;
                    if (v == 0) c3 = 0;
                    else if (v == SPB_MV_THR_ALWAYS) c3 = 1;
                    else if (v == SPB_MV_THR_SLOW || v > SPB_MV8_T_MAX || (T != 0 && T != v)) c3 = 3;
                    else { c3 = 2; T = v; }
                    cls |= c3 << (2 * slot);
                    const uint32_t k2 = nxt[s];
                    const uint64_t dd = k2 == 0xFFFFFFFFu ? 0 : std::min<uint64_t>((uint64_t)k2 - k, SPB_MV8_DIST_PROBE);
                    dist |= dd << (3 * slot);
                }
                const uint64_t lo = (uint64_t)de | ((uint64_t)std::min<uint64_t>(len_k, SPB_MV8_LEN_MAX) << 24) | ((uint64_t)(slow ? 1 : 0) << 31);
                const uint64_t hi = (uint64_t)std::min<uint64_t>(delta, SPB_MV8_DELTA_MAX) | ((uint64_t)std::min<uint32_t>(hk, SPB_SYM_SLOW) << 8) | (cls << 11) | (T << 17) | (dist << 23);
                rows_c[k * (h.mrows8_stride / 8)] = lo | (hi << 32);
                continue;
            }
            if (rows32) {
                uint32_t dest2[3] = {0, 0, 0}, delta2[3] = {0, 0, 0};
                for (uint32_t s = 0; s < num_fast && hk < SPB_SYM_SLOW; s++) {
                    if (s == hk) continue;
                    const uint32_t slot = s - (s > hk ? 1u : 0u);
                    const uint32_t k2 = nxt[s];
                    uint64_t dk, dd;  // landing state of the "up" jump; "down" is one position before it
                    if (k2 == 0xFFFFFFFFu) { dk = end_k[s]; dd = end_off[s] + 1; }  // always down: lands on mv_end itself
                    else { dk = dest[k2]; dd = lf[k2] - start[(size_t)dest[k2]]; }
                    if (dd > SPB_MV32_D2_MAX) { thr3[slot] = SPB_MV_THR_SLOW; continue; }
                    dest2[slot] = (uint32_t)dk;
                    delta2[slot] = (uint32_t)dd;
                }
                SpbMoveRow32 r;
                r.w[0] = (uint32_t)de;
                r.w[1] = (uint32_t)(len_k > SPB_MV_LEN_MASK ? SPB_MV_LEN_MASK : len_k) | (thr3[0] << 20);
                r.w[2] = (uint32_t)(delta > SPB_MV_LEN_MASK ? SPB_MV_LEN_MASK : delta) | (thr3[1] << 20);
                r.w[3] = thr3[2] | ((hk > SPB_SYM_SLOW ? SPB_SYM_SLOW : hk) << 12) | ((slow ? 1u : 0u) << 15) | (delta2[0] << 16);
                r.w[4] = dest2[0]; r.w[5] = dest2[1]; r.w[6] = dest2[2];
                r.w[7] = delta2[1] | (delta2[2] << 16);
                rows_f[k] = r;
                continue;
            }
            SpbMoveRow r;
            r.x = (uint32_t)de;
            r.y = (uint32_t)(len_k > SPB_MV_LEN_MASK ? SPB_MV_LEN_MASK : len_k) | (thr3[0] << 20);
            r.z = (uint32_t)(delta > SPB_MV_LEN_MASK ? SPB_MV_LEN_MASK : delta) | (thr3[1] << 20);
            r.w = thr3[2] | (dist3[0] << 12) | (dist3[1] << 17) | (dist3[2] << 22) | ((hk > SPB_SYM_SLOW ? SPB_SYM_SLOW : hk) << 27) |
                  ((slow ? 1u : 0u) << 30);
            rows[k] = r;
        }
        }
        if (rows8) {
            for (uint64_t k = R; k < R + SPB_MV8_ROW_PAD; k++)  // row R = "pos == n": generic step only
                rows_c[k * (h.mrows8_stride / 8)] = ((uint64_t)R | (1ull << 24) | (1ull << 31)) | ((uint64_t)(SPB_SYM_SLOW << 8) << 32);
        }
        for (uint64_t k = R; rows32 && k < R + SPB_MV_ROW_PAD; k++) {  // row R = "pos == n": generic step only
            SpbMoveRow32 r;
            std::memset(&r, 0, sizeof(r));
            r.w[0] = (uint32_t)R; r.w[1] = 1;
            r.w[3] = (SPB_SYM_SLOW << 12) | (1u << 15);
            rows_f[k] = r;
        }
        for (uint64_t k = R; !rows8 && !rows32 && k < R + SPB_MV_ROW_PAD; k++) {  // row R = "pos == n": generic step only
            SpbMoveRow r;
            r.x = (uint32_t)R; r.y = 1; r.z = 0;
            r.w = (SPB_SYM_SLOW << 27) | (1u << 30);
            rows[k] = r;
        }
        // per-row "previous run of the same letter" fields (MS samples, documents): the last run of every letter before each
        // slice is found first, then the slices fill their rows independently
        const uint64_t kNone = ~0ull;
        std::vector<uint64_t> last_before((size_t)(kChunks + 1) * 256, kNone);
        if (ms || use_doc) {
#pragma omp parallel for schedule(dynamic, 1)
            for (int c = 0; c < kChunks; c++) {
                uint64_t* last = &last_before[(size_t)(c + 1) * 256];
                for (uint64_t k = chunk_lo(c), e = chunk_lo(c + 1); k < e; k++) last[d.heads[(size_t)k]] = k;
            }
            for (int c = 0; c < kChunks; c++)
                for (int a = 0; a < 256; a++)
                    if (last_before[(size_t)(c + 1) * 256 + a] == kNone) last_before[(size_t)(c + 1) * 256 + a] = last_before[(size_t)c * 256 + a];
        }
        const uint64_t* last_run = &last_before[(size_t)kChunks * 256];  // last run of every letter
        if (ms) {
#pragma omp parallel for schedule(dynamic, 1)
            for (int sl = 0; sl < kChunks; sl++) {
                uint64_t prev_sl[256];
                for (int a = 0; a < 256; a++) {
                    const uint64_t kb = last_before[(size_t)sl * 256 + a];
                    prev_sl[a] = kb == kNone ? 0 : d.samples_last[(size_t)kb];
                }
                for (uint64_t k = chunk_lo(sl), e = chunk_lo(sl + 1); k < e; k++) {
                    const uint8_t c = d.heads[(size_t)k];
                    if (msamp) {
                        msamp[k].ss = d.samples_start[(size_t)k];
                        msamp[k].sl = prev_sl[c];
                    } else if (msamp32)
                        msamp32[k] = (uint64_t)(uint32_t)d.samples_start[(size_t)k] | ((uint64_t)(uint32_t)prev_sl[c] << 32);
                    else  // inline: the second 8 bytes of the 16-byte row slot (text positions < n < 2^32)
                        rows_c[k * 2 + 1] = (uint64_t)(uint32_t)d.samples_start[(size_t)k] | ((uint64_t)(uint32_t)prev_sl[c] << 32);
                    prev_sl[c] = d.samples_last[(size_t)k];
                }
            }
            for (uint32_t s = 0; s < num_fast; s++) hd->mv_end_sl[s] = d.samples_last[(size_t)last_run[h.fast_sym[s]]];
        }
        if (use_doc) {
            SpbDocPair* mdoc = reinterpret_cast<SpbDocPair*>(blob.data() + h.off_mdoc);
#pragma omp parallel for schedule(dynamic, 1)
            for (int sl = 0; sl < kChunks; sl++) {
                uint32_t pe[256];
                for (int a = 0; a < 256; a++) {
                    const uint64_t kb = last_before[(size_t)sl * 256 + a];
                    pe[a] = kb == kNone ? 0 : d.doc_end[(size_t)kb];
                }
                for (uint64_t k = chunk_lo(sl), e = chunk_lo(sl + 1); k < e; k++) {
                    const uint8_t c = d.heads[(size_t)k];
                    mdoc[k].sd = d.doc_start[(size_t)k];
                    mdoc[k].ed = pe[c];
                    pe[c] = d.doc_end[(size_t)k];
                }
            }
            for (uint32_t s = 0; s < num_fast; s++) hd->mv_end_doc[s] = d.doc_end[(size_t)last_run[h.fast_sym[s]]];
        }
        for (uint32_t s = 0; s < num_fast; s++) {
            hd->mv_end_k[s] = end_k[s];
            hd->mv_end_off[s] = (uint32_t)std::min<uint64_t>(end_off[s], 0xFFFFFFFEull);
        }
    }
    return blob;
}

void validate_flat_image(const void* blob, uint64_t bytes) {
    if (bytes < sizeof(SpbFlatHeader)) throw FormatError("flat image smaller than its header");
    SpbFlatHeader h;
    std::memcpy(&h, blob, sizeof(h));
    if (h.magic != SPB_FLAT_MAGIC || h.version != SPB_FLAT_VERSION) throw FormatError("not a spumoni_b200 flat image (magic/version)");
    if (h.total_bytes != bytes) throw FormatError("flat image size mismatch");
    if (h.off_runs + (h.R + SPB_RUN_PAD) * sizeof(SpbRunRec) > bytes || h.off_crec + (h.num_crec + 1) * sizeof(SpbCharRec) > bytes)
        throw FormatError("flat image tables exceed the blob");
}

}  // namespace spb
