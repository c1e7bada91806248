// TEST INFRASTRUCTURE ONLY -- CPU oracle for SPUMONI's MS/PML hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load or execute this file's outputs.  The product (spumoni_b200/) never links or calls it.
//
// What it is: a restatement of the reference's `spumoni run` path on the CPU, operating on the
// reference's own succinct structures as they are laid out on disk (Elias-Fano sd_vector,
// Huffman-shaped wavelet tree over run heads, B=2 run sampling, thr_bv thresholds), with the
// same per-read loops and the same OpenMP batch/critical skeleton, so that it can also serve as
// the timed CPU baseline (kind = "port").
//
// PARITY STATUS: pinned against the reference's OWN code, with its third-party containers stood in for.
// The reference has no tests/fixtures for this path and its build cannot run here (12 repositories fetched at
// CMake time: /root/reference/thirdparty/CMakeLists.txt), but its translation units compile where they lie
// against stand-ins for the un-vendored sdsl-lite / r-index / ShapedSlp / klib / bonsai headers (oracle/shim/,
// recipe oracle/Makefile.ref -> oracle/_ref/spumoni_ref). This file's `spumoni run` output is byte-identical to
// that binary's in every mode (PML/MS x report x document numbers; tests/test_reference_pin.py, committed
// outputs under tests/golden/reference_run/). What that pin does NOT cover: the succinct containers of
// sdsl-lite / r-index themselves (the stand-ins answer rank/select by definition) -- "parity unpinned" still
// holds for SURVEY.md App. A's reconstruction of their byte formats. Further pins: oracle/naive.py, SURVEY.md
// Appendix D, brute-force matching statistics (tests/test_oracle.py).
//
// Reference lines followed (under /root/reference):
//   pml_pointers::load/_query/LF      src/compute_ms_pml.cpp:222-229, :237-286, :180-187
//   ms_pointers::load/_query          src/compute_ms_pml.cpp:551-562, :570-623
//   ms_t::matching_statistics         src/compute_ms_pml.cpp:795-812
//   classify_reads_pml / _ms          src/compute_ms_pml.cpp:845-1034, :1036-1217
//   ms_rle_string ops                 include/ms_rle_string.hpp:95-146
//   thr_bv::operator[] / load         include/thresholds_ds.hpp:478-491, :517-523
//   BatchLoader                       src/batch_loader.cpp:26-131
//   EmpNullDatabase::load             src/emp_null_database.cpp:104-110
//   DocumentArray::load               src/doc_array.cpp:197-201
//   generate_null_*_statistics        src/compute_ms_pml.cpp:1410-1486; EmpNullDatabase ctor src/emp_null_database.cpp:19-80
//   _query with document numbers      src/compute_ms_pml.cpp:288-340 (PML), :625-682 (MS), ms_t :814-828
//   my_load                           include/common.hpp:487-495
//   ri::rle_string / sparse_sd_vector / huff_string, sdsl::sd_vector / wt_huff / int_vector:
//   third party, restated from their published sources.
#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

typedef uint64_t ulint;

// ------------------------------------------------------------------ byte cursor
struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    template <typename T> T get() {
        if (p + sizeof(T) > end) throw std::runtime_error("oracle: truncated index file");
        T v;
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    const uint8_t* take(size_t n) {
        if (p + n > end) throw std::runtime_error("oracle: truncated index file");
        const uint8_t* q = p;
        p += n;
        return q;
    }
};

// ------------------------------------------------------------------ sdsl::int_vector family
struct IntVector {
    ulint bit_size = 0;
    uint8_t width = 64;
    std::vector<uint64_t> w;
    void load(Cursor& c, int fixed_width /*0 = stored*/) {
        bit_size = c.get<uint64_t>();
        width = fixed_width ? (uint8_t)fixed_width : c.get<uint8_t>();
        size_t nw = (size_t)((bit_size + 63) >> 6);
        const uint8_t* d = c.take(nw * 8);
        w.resize(nw + 1);
        std::memcpy(w.data(), d, nw * 8);
        w[nw] = 0;
    }
    inline ulint size() const { return bit_size / width; }
    inline ulint operator[](ulint i) const {
        ulint b = i * width;
        uint32_t off = b & 63;
        uint64_t v = w[b >> 6] >> off;
        if (off + width > 64) v |= w[(b >> 6) + 1] << (64 - off);
        return width == 64 ? v : (v & ((1ULL << width) - 1));
    }
    inline bool bit(ulint i) const { return (w[i >> 6] >> (i & 63)) & 1; }
};

static void skip_select_mcl(Cursor& c) {
    ulint arg_cnt = c.get<uint64_t>();
    if (!arg_cnt) return;
    IntVector tmp;
    tmp.load(c, 0);  // superblock
    tmp.load(c, 1);  // mini_or_long
    ulint sb = (arg_cnt + 4095) >> 12;
    for (ulint i = 0; i < sb; i++) tmp.load(c, 0);
}

static inline uint32_t select_in_word(uint64_t x, uint32_t r /*0-based*/) {
    for (uint32_t k = 0; k < r; k++) x &= x - 1;
    return (uint32_t)__builtin_ctzll(x);
}

// constant-time-ish select over a bit vector (stand-in for sdsl::select_support_mcl)
struct Select {
    const IntVector* bv = nullptr;
    bool ones = true;
    std::vector<ulint> samp;  // position of every 64th argument
    ulint cnt = 0;
    void build(const IntVector* v, bool ones_) {
        bv = v;
        ones = ones_;
        samp.clear();
        cnt = 0;
        for (ulint i = 0; i < v->bit_size; i++)
            if (v->bit(i) == ones) {
                if ((cnt & 63) == 0) samp.push_back(i);
                cnt++;
            }
    }
    inline ulint operator()(ulint j /*1-based*/) const {
        ulint k = j - 1;
        ulint pos = samp[k >> 6];
        uint32_t need = (uint32_t)(k & 63);
        ulint wi = pos >> 6;
        uint64_t word = ones ? bv->w[wi] : ~bv->w[wi];
        word &= ~0ULL << (pos & 63);
        for (;;) {
            uint32_t pc = (uint32_t)__builtin_popcountll(word);
            if (need < pc) return (wi << 6) + select_in_word(word, need);
            need -= pc;
            ++wi;
            word = ones ? bv->w[wi] : ~bv->w[wi];
        }
    }
};

// rank over a bit vector (stand-in for sdsl::rank_support_v)
struct Rank {
    const IntVector* bv = nullptr;
    std::vector<ulint> blk;  // ones before each 512-bit block
    void build(const IntVector* v) {
        bv = v;
        size_t nw = (size_t)((v->bit_size + 63) >> 6);
        blk.assign(nw / 8 + 2, 0);
        ulint s = 0;
        for (size_t i = 0; i < nw; i++) {
            if ((i & 7) == 0) blk[i >> 3] = s;
            s += __builtin_popcountll(v->w[i]);
        }
        for (size_t b = (nw + 7) / 8; b < blk.size(); b++) blk[b] = s;
    }
    inline ulint operator()(ulint i) const {  // ones in [0,i)
        size_t wi = (size_t)(i >> 6);
        ulint r = blk[wi >> 3];
        for (size_t k = wi & ~(size_t)7; k < wi; k++) r += __builtin_popcountll(bv->w[k]);
        if (i & 63) r += __builtin_popcountll(bv->w[wi] & ((1ULL << (i & 63)) - 1));
        return r;
    }
};

// ------------------------------------------------------------------ sdsl::sd_vector<> + ri::sparse_sd_vector
struct SparseSd {
    ulint u = 0, n = 0;  // universe, ones
    ulint m_size = 0;
    uint8_t wl = 0;
    IntVector low, high;
    Select high1, high0;
    void load(Cursor& c) {
        u = c.get<uint64_t>();
        n = c.get<uint64_t>();
        if (u == 0) return;
        m_size = c.get<uint64_t>();
        wl = c.get<uint8_t>();
        low.load(c, 0);
        high.load(c, 1);
        skip_select_mcl(c);
        skip_select_mcl(c);
        high1.build(&high, true);
        high0.build(&high, false);
        if (m_size != u || high1.cnt != n || low.size() != n) throw std::runtime_error("oracle: inconsistent sparse_sd_vector");
    }
    inline ulint size() const { return u; }
    inline ulint number_of_1() const { return n; }
    // sdsl::rank_support_sd<1>::rank : ones in [0,i)
    inline ulint rank(ulint i) const {
        ulint high_val = i >> wl;
        ulint sel_high = high0(high_val + 1);
        ulint rank_low = sel_high - high_val;
        if (rank_low == 0) return 0;
        ulint val_low = i & ((1ULL << wl) - 1);
        do {
            if (!sel_high) return 0;
            --sel_high;
            --rank_low;
        } while (high.bit(sel_high) && low[rank_low] >= val_low);
        return rank_low + 1;
    }
    // ri::sparse_sd_vector::select(i): position of the i-th one, i 0-based
    inline ulint select(ulint i) const { return low[i] + ((high1(i + 1) - i) << wl); }
    inline ulint gapAt(ulint i) const { return i == 0 ? select(0) + 1 : select(i) - select(i - 1); }
};

// ------------------------------------------------------------------ sdsl::wt_huff<> (ri::huff_string)
struct WtHuff {
    static const uint16_t UNDEF = 0xFFFF;
    struct Node { ulint bv_pos, bv_pos_rank; uint16_t parent, child[2]; };
    ulint m_size = 0, m_sigma = 0;
    IntVector bv;
    Rank rk;
    Select sel1, sel0;
    std::vector<Node> nodes;
    uint16_t c_to_leaf[256];
    uint64_t path[256];
    void load(Cursor& c) {
        m_size = c.get<uint64_t>();
        m_sigma = c.get<uint64_t>();
        bv.load(c, 1);
        IntVector skip;
        skip.load(c, 64);  // rank_support_v basic blocks (rebuilt below)
        skip_select_mcl(c);
        skip_select_mcl(c);
        ulint nn = c.get<uint64_t>();
        nodes.resize(nn);
        for (auto& nd : nodes) {
            nd.bv_pos = c.get<uint64_t>();
            nd.bv_pos_rank = c.get<uint64_t>();
            nd.parent = c.get<uint16_t>();
            nd.child[0] = c.get<uint16_t>();
            nd.child[1] = c.get<uint16_t>();
        }
        for (int i = 0; i < 256; i++) c_to_leaf[i] = c.get<uint16_t>();
        for (int i = 0; i < 256; i++) path[i] = c.get<uint64_t>();
        rk.build(&bv);
        sel1.build(&bv, true);
        sel0.build(&bv, false);
    }
    inline ulint size() const { return m_size; }
    inline uint8_t operator[](ulint i) const {
        uint16_t v = 0;
        while (nodes[v].child[0] != UNDEF) {
            const Node& nd = nodes[v];
            ulint ones = rk(nd.bv_pos + i) - nd.bv_pos_rank;
            if (bv.bit(nd.bv_pos + i)) { i = ones; v = nd.child[1]; }
            else { i -= ones; v = nd.child[0]; }
        }
        return (uint8_t)nodes[v].bv_pos_rank;
    }
    inline ulint rank(ulint i, uint8_t c) const {  // occurrences of c in [0,i)
        if (c_to_leaf[c] == UNDEF) return 0;
        uint16_t v = 0;
        uint64_t p = path[c];
        uint32_t pl = (uint32_t)(p >> 56);
        ulint res = i;
        for (uint32_t l = 0; res && l < pl; ++l, p >>= 1) {
            const Node& nd = nodes[v];
            ulint ones = rk(nd.bv_pos + res) - nd.bv_pos_rank;
            res = (p & 1) ? ones : res - ones;
            v = nd.child[p & 1];
        }
        return res;
    }
    inline ulint select(ulint i /*1-based*/, uint8_t c) const {
        uint16_t v = c_to_leaf[c];
        uint64_t p = path[c];
        uint32_t pl = (uint32_t)(p >> 56);
        for (uint32_t l = pl; l > 0; --l) {
            v = nodes[v].parent;
            const Node& nd = nodes[v];
            if ((p >> (l - 1)) & 1) i = sel1(nd.bv_pos_rank + i) - nd.bv_pos + 1;
            else i = sel0(nd.bv_pos - nd.bv_pos_rank + i) - nd.bv_pos + 1;
        }
        return i - 1;
    }
};

// ------------------------------------------------------------------ ri::rle_string + ms_rle_string
struct RleString {
    ulint n = 0, R = 0, B = 2;
    SparseSd runs;
    std::vector<SparseSd> runs_per_letter;
    WtHuff run_heads;
    void load(Cursor& c) {
        n = c.get<uint64_t>();
        R = c.get<uint64_t>();
        B = c.get<uint64_t>();
        if (n == 0) return;
        runs.load(c);
        runs_per_letter.resize(256);
        for (int i = 0; i < 256; i++) runs_per_letter[i].load(c);
        run_heads.load(c);
        if (run_heads.size() != R) throw std::runtime_error("oracle: run_heads.size() != R");
    }
    inline ulint size() const { return n; }
    inline ulint number_of_runs() const { return R; }
    inline ulint number_of_letter(uint8_t c) const { return runs_per_letter[c].size(); }
    inline uint8_t head_of(ulint i) const { return run_heads[i]; }
    inline ulint run_head_rank(ulint i, uint8_t c) const { return run_heads.rank(i, c); }
    inline ulint run_at(ulint i) const {
        uint8_t c = run_heads[i];
        return runs_per_letter[c].gapAt(run_heads.rank(i, c));
    }
    inline ulint run_of_position(ulint i) const {
        ulint last_block = runs.rank(i);
        ulint current_run = last_block * B;
        ulint pos = last_block > 0 ? runs.select(last_block - 1) + 1 : 0;
        while (pos < i) {
            pos += run_at(current_run);
            current_run++;
        }
        if (pos > i) current_run--;
        return current_run;
    }
    inline uint8_t operator[](ulint i) const { return run_heads[run_of_position(i)]; }
    inline ulint rank(ulint i, uint8_t c) const {
        if (runs_per_letter[c].size() == 0) return 0;
        if (i == n) return runs_per_letter[c].size();
        ulint last_block = runs.rank(i);
        ulint current_run = last_block * B;
        ulint pos = last_block > 0 ? runs.select(last_block - 1) + 1 : 0;
        ulint dist = i - pos;
        while (pos < i) {
            pos += run_at(current_run);
            current_run++;
            if (pos <= i) dist = i - pos;
        }
        if (pos > i) current_run--;
        ulint rk = run_heads.rank(current_run, c);
        ulint tail = (run_heads[current_run] == c) * dist;
        if (rk == 0) return tail;
        return runs_per_letter[c].select(rk - 1) + 1 + tail;
    }
    inline ulint select(ulint i, uint8_t c) const {
        ulint j = runs_per_letter[c].rank(i);
        ulint before = (j == 0 ? i : i - (runs_per_letter[c].select(j - 1) + 1));
        ulint r = run_heads.select(j + 1, c);  // huff_string::select(j,c) == wt.select(j+1,c)
        ulint k = (r / B == 0 ? 0 : runs.select(r / B - 1) + 1);
        for (ulint t = (r / B) * B; t < r; ++t) k += run_at(t);
        return k + before;
    }
};

struct ThrBv {
    std::vector<SparseSd> per_letter;
    const RleString* bwt = nullptr;
    void load(Cursor& c, const RleString* b) {
        per_letter.resize(256);
        for (int i = 0; i < 256; i++) per_letter[i].load(c);
        bwt = b;
    }
    inline ulint operator[](ulint i) const {
        uint8_t c = bwt->head_of(i);
        ulint rank = bwt->run_head_rank(i, c);
        if (rank == 0) return 0;
        return per_letter[c].select(rank - 1);
    }
};

static std::vector<uint8_t> slurp(const std::string& path) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) throw std::runtime_error("oracle: cannot open " + path);
    size_t sz = (size_t)in.tellg();
    in.seekg(0);
    std::vector<uint8_t> b(sz);
    in.read((char*)b.data(), (std::streamsize)sz);
    return b;
}

// ------------------------------------------------------------------ pml_pointers / ms_pointers
// DocumentArray (src/doc_array.cpp:184-201): document of the BWT character at the start / end of every run
struct DocArray {
    ulint num_entries = 0;
    IntVector start_runs_doc, end_runs_doc;
    void load(const std::string& path) {
        std::vector<uint8_t> buf = slurp(path);
        Cursor c{buf.data(), buf.data() + buf.size()};
        num_entries = c.get<uint64_t>();
        start_runs_doc.load(c, 0);
        end_runs_doc.load(c, 0);
        if (c.p != c.end) throw std::runtime_error("oracle: trailing bytes in " + path);
    }
};

struct Pointers {
    bool is_ms = false;
    ulint terminator_position = 0;
    std::vector<ulint> F;
    RleString bwt;
    ThrBv thresholds;
    IntVector samples_last, samples_start;

    void load(const std::string& path, bool ms) {
        is_ms = ms;
        std::vector<uint8_t> buf = slurp(path);
        Cursor c{buf.data(), buf.data() + buf.size()};
        terminator_position = c.get<uint64_t>();
        ulint fs = c.get<uint64_t>();
        F.resize(fs);
        for (auto& f : F) f = c.get<uint64_t>();
        bwt.load(c);
        if (ms) samples_last.load(c, 0);
        thresholds.load(c, &bwt);
        if (ms) samples_start.load(c, 0);
        if (c.p != c.end) throw std::runtime_error("oracle: trailing bytes in " + path);
    }
    inline ulint bwt_size() const { return bwt.size(); }
    inline ulint LF(ulint i, uint8_t c) const { return F[c] + bwt.rank(i, c); }
    inline ulint get_last_run_sample() const { return (samples_last[bwt.number_of_runs() - 1] + 1) % bwt.size(); }

    // pml_pointers::_query  (src/compute_ms_pml.cpp:237-286)
    void query_pml(const char* pattern, size_t m, std::vector<size_t>& lengths) const {
        lengths.resize(m);
        auto pos = bwt_size() - 1;
        auto length = 0;
        for (size_t i = 0; i < m; ++i) {
            auto c = pattern[m - i - 1];  // (signed) char, as in the reference
            if (bwt.number_of_letter(c) == 0) { length = 0; }
            else if (pos < bwt.size() && bwt[pos] == c) { length++; }
            else {
                ulint rnk = bwt.rank(pos, c);
                size_t thr = bwt.size() + 1;
                ulint next_pos = pos;
                if (rnk < bwt.number_of_letter(c)) {
                    ulint j = bwt.select(rnk, c);
                    ulint run_of_j = bwt.run_of_position(j);
                    thr = thresholds[run_of_j];
                    length = 0;
                    next_pos = j;
                }
                if (pos < thr) {
                    rnk--;
                    ulint j = bwt.select(rnk, c);
                    length = 0;
                    next_pos = j;
                }
                pos = next_pos;
            }
            lengths[m - i - 1] = length;
            pos = LF(pos, c);
        }
    }
    // pml_pointers::_query with document numbers  (src/compute_ms_pml.cpp:288-340)
    void query_pml(const char* pattern, size_t m, std::vector<size_t>& lengths, std::vector<size_t>& doc_nums, const DocArray& doc_arr) const {
        lengths.resize(m);
        doc_nums.resize(m);
        auto pos = bwt_size() - 1;
        auto length = 0;
        size_t curr_doc_id = doc_arr.end_runs_doc[bwt.number_of_runs() - 1];
        for (size_t i = 0; i < m; ++i) {
            auto c = pattern[m - i - 1];
            if (bwt.number_of_letter(c) == 0) { length = 0; }
            else if (pos < bwt.size() && bwt[pos] == c) { length++; }
            else {
                ulint rnk = bwt.rank(pos, c);
                size_t thr = bwt.size() + 1;
                ulint next_pos = pos;
                if (rnk < bwt.number_of_letter(c)) {
                    ulint j = bwt.select(rnk, c);
                    ulint run_of_j = bwt.run_of_position(j);
                    thr = thresholds[run_of_j];
                    curr_doc_id = doc_arr.start_runs_doc[run_of_j];
                    length = 0;
                    next_pos = j;
                }
                if (pos < thr) {
                    rnk--;
                    ulint j = bwt.select(rnk, c);
                    ulint run_of_j = bwt.run_of_position(j);
                    curr_doc_id = doc_arr.end_runs_doc[run_of_j];
                    length = 0;
                    next_pos = j;
                }
                pos = next_pos;
            }
            lengths[m - i - 1] = length;
            doc_nums[m - i - 1] = curr_doc_id;
            pos = LF(pos, c);
        }
    }
    // ms_pointers::_query with document numbers  (src/compute_ms_pml.cpp:625-682)
    void query_ms(const char* pattern, size_t m, std::vector<size_t>& ptrs, std::vector<size_t>& doc_nums, const DocArray& doc_arr) const {
        ptrs.resize(m);
        doc_nums.resize(m);
        auto pos = bwt_size() - 1;
        auto sample = get_last_run_sample();
        size_t curr_doc_id = doc_arr.end_runs_doc[bwt.number_of_runs() - 1];
        for (size_t i = 0; i < m; ++i) {
            auto c = pattern[m - i - 1];
            if (bwt.number_of_letter(c) == 0) {
                sample = 0;
                ulint run_of_j = bwt.run_of_position(sample);  // a text position used as a BWT position
                curr_doc_id = doc_arr.start_runs_doc[run_of_j];
            }
            else if (pos < bwt.size() && bwt[pos] == c) { sample--; }
            else {
                ulint rnk = bwt.rank(pos, c);
                size_t thr = bwt.size() + 1;
                ulint next_pos = pos;
                if (rnk < bwt.number_of_letter(c)) {
                    ulint j = bwt.select(rnk, c);
                    ulint run_of_j = bwt.run_of_position(j);
                    thr = thresholds[run_of_j];
                    sample = samples_start[run_of_j];
                    curr_doc_id = doc_arr.start_runs_doc[run_of_j];
                    next_pos = j;
                }
                if (pos < thr) {
                    rnk--;
                    ulint j = bwt.select(rnk, c);
                    ulint run_of_j = bwt.run_of_position(j);
                    sample = samples_last[run_of_j];
                    curr_doc_id = doc_arr.end_runs_doc[run_of_j];
                    next_pos = j;
                }
                pos = next_pos;
            }
            ptrs[m - i - 1] = sample;
            doc_nums[m - i - 1] = curr_doc_id;
            pos = LF(pos, c);
        }
    }
    // ms_pointers::_query  (src/compute_ms_pml.cpp:570-623)
    void query_ms(const char* pattern, size_t m, std::vector<size_t>& ptrs) const {
        ptrs.resize(m);
        auto pos = bwt_size() - 1;
        auto sample = get_last_run_sample();
        for (size_t i = 0; i < m; ++i) {
            auto c = pattern[m - i - 1];
            if (bwt.number_of_letter(c) == 0) { sample = 0; }
            else if (pos < bwt.size() && bwt[pos] == c) { sample--; }
            else {
                ulint rnk = bwt.rank(pos, c);
                size_t thr = bwt.size() + 1;
                ulint next_pos = pos;
                if (rnk < bwt.number_of_letter(c)) {
                    ulint j = bwt.select(rnk, c);
                    ulint run_of_j = bwt.run_of_position(j);
                    thr = thresholds[run_of_j];
                    sample = samples_start[run_of_j];
                    next_pos = j;
                }
                if (pos < thr) {
                    rnk--;
                    ulint j = bwt.select(rnk, c);
                    ulint run_of_j = bwt.run_of_position(j);
                    sample = samples_last[run_of_j];
                    next_pos = j;
                }
                pos = next_pos;
            }
            ptrs[m - i - 1] = sample;
            pos = LF(pos, c);
        }
    }
};

// Text of <prefix>.fa: concatenation of the sequence lines (stand-in for the SLP random access,
// ms_t::ra, src/compute_ms_pml.cpp:769-775; ShapedSlp's byte format is not derivable here).
static std::string load_fasta_text(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("oracle: cannot open " + path);
    std::string T, line;
    while (std::getline(in, line)) {
        if (!line.empty() && line[0] == '>') continue;
        while (!line.empty() && std::isspace((unsigned char)line.back())) line.pop_back();
        T += line;
    }
    return T;
}

struct PmlT {
    Pointers ms;
    DocArray doc_arr;
    void matching_statistics(const char* read, size_t len, std::vector<size_t>& lengths) const { ms.query_pml(read, len, lengths); }
    void matching_statistics(const char* read, size_t len, std::vector<size_t>& lengths, std::vector<size_t>& doc_nums) const {
        ms.query_pml(read, len, lengths, doc_nums, doc_arr);
    }
};
struct MsT {
    Pointers ms;
    DocArray doc_arr;
    std::string ra;
    size_t n = 0;
    // ms_t::matching_statistics with document numbers  (src/compute_ms_pml.cpp:814-828)
    void matching_statistics(const char* read, size_t read_length, std::vector<size_t>& lengths, std::vector<size_t>& pointers,
                             std::vector<size_t>& doc_nums) const {
        ms.query_ms(read, read_length, pointers, doc_nums, doc_arr);
        lengths.resize(read_length);
        size_t l = 0;
        for (size_t i = 0; i < pointers.size(); ++i) {
            size_t pos = pointers[i];
            while ((i + l) < read_length && (pos + l) < n && (i < 1 || pos != (pointers[i - 1] + 1)) && read[i + l] == ra[pos + l]) ++l;
            lengths[i] = l;
            l = (l == 0 ? 0 : (l - 1));
        }
    }
    // ms_t::matching_statistics  (src/compute_ms_pml.cpp:795-812)
    void matching_statistics(const char* read, size_t read_length, std::vector<size_t>& lengths, std::vector<size_t>& pointers) const {
        ms.query_ms(read, read_length, pointers);
        lengths.resize(read_length);
        size_t l = 0;
        for (size_t i = 0; i < pointers.size(); ++i) {
            size_t pos = pointers[i];
            while ((i + l) < read_length && (pos + l) < n && (i < 1 || pos != (pointers[i - 1] + 1)) && read[i + l] == ra[pos + l]) ++l;
            lengths[i] = l;
            l = (l == 0 ? 0 : (l - 1));
        }
    }
};

// ------------------------------------------------------------------ EmpNullDatabase::load
struct NullDb {
    uint64_t num_values = 0;
    double ks_stat_threshold = 0, mean_null_stat = 0, percentile_value = 0;
    IntVector null_stats;
    void load(const std::string& path) {
        std::vector<uint8_t> buf = slurp(path);
        Cursor c{buf.data(), buf.data() + buf.size()};
        num_values = c.get<uint64_t>();
        ks_stat_threshold = c.get<double>();
        mean_null_stat = c.get<double>();
        percentile_value = c.get<double>();
        null_stats.load(c, 0);
    }
};

// ------------------------------------------------------------------ BatchLoader (src/batch_loader.cpp)
struct Read { std::string header_line, id, seq; };
struct BatchLoader {
    int fmt = 0;  // 0 unknown, 1 FASTA, 2 FASTQ
    std::string buf;
    std::stringstream batch;
    bool loadBatch(std::ifstream& input, size_t num_bases) {
        if (fmt == 0) {
            if (!input.good()) return false;
            int pk = input.peek();
            if (pk == '>') fmt = 1;
            else if (pk == '@') fmt = 2;
            else if (pk == EOF) return false;
            else throw std::runtime_error("unrecognized input query file type - expects FASTA or FASTQ.");
        }
        batch.clear();
        batch.str("");
        size_t covered = 0, record = 0, lines = 0;
        bool valid = false;
        while (input.good() && covered < num_bases) {
            if (!std::getline(input, buf)) return false;  // FASTQ: the records gathered so far are lost with it (:51)
            lines++;
            record += buf.size();
            valid = true;
            if (fmt == 2) {
                if (lines % 4 == 0) { covered += record / 2; record = 0; }  // half of a record's four lines is its read (:61-66)
            } else if (input.peek() == '>') { covered += record; record = 0; }
            batch << buf << '\n';
        }
        return valid;
    }
    bool grabNextRead(Read& r) {
        if (!std::getline(batch, buf)) return false;
        if (!buf.size()) return false;
        if (fmt == 2 && buf[0] != '@') throw std::runtime_error("Incorrect FASTQ entry, it should start with '@'");
        if (fmt == 1 && buf[0] != '>') throw std::runtime_error("Incorrect FASTA entry, it should start with '>'");
        r.header_line.assign(buf);
        if (buf.size() <= 2) throw std::runtime_error("header line is missing an id. invalid query cannot be processed.");
        auto id_length = buf.find_first_of(" \t\r", 1);
        id_length = (id_length == std::string::npos) ? buf.size() : id_length;
        r.id.assign(buf.substr(1, id_length));
        r.seq.assign("");
        if (fmt == 2) {  // sequence line, '+' line, quality line (:108-120)
            if (!std::getline(batch, buf)) return false;
            while (!buf.empty() && std::isspace((unsigned char)buf.back())) buf.pop_back();
            r.seq.assign(buf);
            if (!std::getline(batch, buf)) return false;
            if (!std::getline(batch, buf)) return false;
            return true;
        }
        while (batch.good() && batch.peek() != '>') {
            if (!std::getline(batch, buf)) return r.seq.size();
            while (!buf.empty() && std::isspace((unsigned char)buf.back())) buf.pop_back();
            r.seq.append(buf);
        }
        return true;
    }
};

struct BinStats { bool found; size_t sum_max, above, below, nbins; };
// bin-max classifier (src/compute_ms_pml.cpp:969-995)
static BinStats classify(const std::vector<size_t>& lengths, size_t bin_width, size_t max_value_thr) {
    BinStats s{false, 0, 0, 0, 0};
    size_t start_pos = 0, end_pos = 0;
    while (start_pos < lengths.size()) {
        end_pos = (start_pos + bin_width < lengths.size()) ? start_pos + bin_width : lengths.size();
        if (lengths.size() - end_pos < bin_width) end_pos = lengths.size();
        auto max_val = *std::max_element(lengths.begin() + start_pos, lengths.begin() + end_pos);
        if (max_val >= max_value_thr) s.above++; else s.below++;
        s.sum_max += max_val;
        s.nbins++;
        start_pos += (end_pos - start_pos);
    }
    s.found = (s.above / (s.above + s.below + 0.0) > 0.50);
    return s;
}

static void report_header(std::ofstream& f, size_t thr) {
    f.precision(4);
    f << std::setw(30) << std::left << "read id:" << std::setw(15) << std::left << "status:"
      << std::setw(19) << std::left << "avg max-value (thr=" << std::setw(2) << std::left << thr
      << std::setw(5) << std::left << "):" << std::setw(12) << std::left << "above thr:"
      << std::setw(12) << std::left << "below thr:" << std::endl;
}
static void report_row(std::ofstream& f, const std::string& id, const BinStats& s) {
    f.precision(3);
    f << std::setw(30) << std::left << id << std::setw(15) << std::left << (s.found ? "FOUND" : "NOT_PRESENT")
      << std::setw(26) << std::left << (s.sum_max + 0.0) / s.nbins << std::setw(12) << std::left << s.above
      << std::setw(12) << std::left << s.below << std::endl;
}

struct RunOpts {
    std::string ref_file, pattern_file;  // ref_file already has ".fa" appended (spumoni.cpp:744-747)
    bool ms = false, write_report = false, use_doc = false;
    size_t threads = 1, bin_width = 150;
};

// classify_reads_pml (src/compute_ms_pml.cpp:845-1034) / classify_reads_ms (:1036-1217), raw-DNA (-n) path
static size_t classify_reads(const PmlT* pml, const MsT* mst, const RunOpts& o, double* seconds) {
    std::ofstream lengths_file(o.pattern_file + (o.ms ? ".lengths" : ".pseudo_lengths"));
    std::ofstream pointers_file, report_file, doc_file;
    if (o.ms) pointers_file.open(o.pattern_file + ".pointers");
    if (o.use_doc) doc_file.open(o.pattern_file + ".doc_numbers");
    std::ostream_iterator<size_t> doc_iter(doc_file, " ");
    std::ostream_iterator<size_t> lengths_iter(lengths_file, " ");
    std::ostream_iterator<size_t> pointers_iter(pointers_file, " ");
    if (o.write_report) report_file.open(o.pattern_file + ".report", std::ofstream::out);
    NullDb null_db;
    null_db.load(o.ref_file + (o.ms ? ".msnulldb" : ".pmlnulldb"));
    size_t max_value_thr = std::max(null_db.percentile_value, 3.0);
    if (!o.ms) max_value_thr += 4;  // raw DNA: !use_dna_letters && !use_promotions (:874-875)
    if (o.write_report) report_header(report_file, max_value_thr);

    std::ifstream input_file(o.pattern_file.c_str());
#ifdef _OPENMP
    omp_set_num_threads((int)o.threads);
#endif
    size_t num_reads = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel
    {
        BatchLoader reader;
        while (true) {
            bool valid_batch = true;
#pragma omp critical
            { valid_batch = reader.loadBatch(input_file, 1000); }
            if (!valid_batch) break;
            Read rs;
            while (true) {
                if (!reader.grabNextRead(rs)) break;
                std::string curr_read(rs.seq);
                std::transform(curr_read.begin(), curr_read.end(), curr_read.begin(), ::toupper);
                if (curr_read.length() == 0) {
                    std::fprintf(stderr, "Warning: %s was empty after digestion\n\n", rs.id.data());
                    std::exit(1);
                }
                std::vector<size_t> lengths, pointers, doc_nums;
                if (o.ms) {
                    if (o.use_doc) mst->matching_statistics(curr_read.c_str(), curr_read.size(), lengths, pointers, doc_nums);
                    else mst->matching_statistics(curr_read.c_str(), curr_read.size(), lengths, pointers);
                } else {
                    if (o.use_doc) pml->matching_statistics(curr_read.c_str(), curr_read.size(), lengths, doc_nums);
                    else pml->matching_statistics(curr_read.c_str(), curr_read.size(), lengths);
                }
                BinStats bs{false, 0, 0, 0, 0};
                if (o.write_report) bs = classify(lengths, o.bin_width, max_value_thr);
#pragma omp atomic
                num_reads++;
#pragma omp critical
                {
                    if (o.use_doc) {
                        doc_file << '>' << rs.id << '\n';
                        std::copy(doc_nums.begin(), doc_nums.end(), doc_iter);
                        doc_file << '\n';
                    }
                    lengths_file << '>' << rs.id << '\n';
                    std::copy(lengths.begin(), lengths.end(), lengths_iter);
                    lengths_file << '\n';
                    if (o.ms) {
                        pointers_file << '>' << rs.id << '\n';
                        std::copy(pointers.begin(), pointers.end(), pointers_iter);
                        pointers_file << '\n';
                    }
                    if (o.write_report) report_row(report_file, rs.id, bs);
                }
            }
        }
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return num_reads;
}

}  // namespace orc

// ====================================================================== C interface (ctypes)
static thread_local std::string g_err;
struct OrcHandle {
    bool ms = false, has_doc = false;
    orc::PmlT pml;
    orc::MsT mst;
};

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ref_file: "<prefix>.fa"; mode 0 = PML (.thrbv.spumoni), 1 = MS (.thrbv.ms + text from <prefix>.fa)
void* orc_open(const char* ref_file, int mode) {
    try {
        OrcHandle* h = new OrcHandle();
        h->ms = mode == 1;
        std::string ref(ref_file);
        if (h->ms) {
            h->mst.ms.load(ref + ".thrbv.ms", true);
            h->mst.ra = orc::load_fasta_text(ref);
            h->mst.n = h->mst.ra.size();
            if (h->mst.n + 1 != h->mst.ms.bwt.size()) throw std::runtime_error("oracle: len(text)+1 != n");
        } else
            h->pml.ms.load(ref + ".thrbv.spumoni", false);
        return h;
    } catch (std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
// pml_t / ms_t constructor with use_doc (src/compute_ms_pml.cpp:712-720, :777-785): loads <ref_file>.doc
int orc_load_doc(void* hv, const char* ref_file) {
    OrcHandle* h = (OrcHandle*)hv;
    try {
        orc::DocArray& d = h->ms ? h->mst.doc_arr : h->pml.doc_arr;
        d.load(std::string(ref_file) + ".doc");
        h->has_doc = true;
        return 0;
    } catch (std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
void orc_close(void* h) { delete (OrcHandle*)h; }
void orc_stats(void* hv, uint64_t* n, uint64_t* r) {
    OrcHandle* h = (OrcHandle*)hv;
    const orc::Pointers& p = h->ms ? h->mst.ms : h->pml.ms;
    *n = p.bwt.size();
    *r = p.bwt.number_of_runs();
}
// one read; out arrays have m entries (u64, as the reference's std::vector<size_t>)
int orc_query(void* hv, const char* read, uint64_t m, uint64_t* lengths, uint64_t* pointers) {
    OrcHandle* h = (OrcHandle*)hv;
    std::vector<size_t> L, P;
    if (h->ms) {
        h->mst.matching_statistics(read, m, L, P);
        if (pointers) std::copy(P.begin(), P.end(), pointers);
    } else
        h->pml.matching_statistics(read, m, L);
    std::copy(L.begin(), L.end(), lengths);
    return 0;
}
// batch over concatenated reads (in-memory timing leg): returns seconds spent in the query loop.
// The per-read std::vector allocation of the reference is kept inside the loop on purpose.
double orc_query_batch_doc(void* hv, const uint8_t* reads, const uint64_t* off, uint64_t n_reads, uint64_t* lengths,
                           uint64_t* pointers, uint64_t* docs, int threads) {
    OrcHandle* h = (OrcHandle*)hv;
    if (!h->has_doc) return -1.0;
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : 1);
#endif
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t r = 0; r < (int64_t)n_reads; r++) {
        std::vector<size_t> L, P, D;
        const char* s = (const char*)reads + off[r];
        size_t m = (size_t)(off[r + 1] - off[r]);
        if (h->ms) h->mst.matching_statistics(s, m, L, P, D);
        else h->pml.matching_statistics(s, m, L, D);
        if (lengths) std::copy(L.begin(), L.end(), lengths + off[r]);
        if (pointers && h->ms) std::copy(P.begin(), P.end(), pointers + off[r]);
        if (docs) std::copy(D.begin(), D.end(), docs + off[r]);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
double orc_query_batch(void* hv, const uint8_t* reads, const uint64_t* off, uint64_t n_reads, uint64_t* lengths,
                       uint64_t* pointers, int threads) {
    OrcHandle* h = (OrcHandle*)hv;
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : 1);
#endif
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t r = 0; r < (int64_t)n_reads; r++) {
        std::vector<size_t> L, P;
        const char* s = (const char*)reads + off[r];
        size_t m = (size_t)(off[r + 1] - off[r]);
        if (h->ms) h->mst.matching_statistics(s, m, L, P);
        else h->pml.matching_statistics(s, m, L);
        if (lengths) std::copy(L.begin(), L.end(), lengths + off[r]);
        if (pointers && h->ms) std::copy(P.begin(), P.end(), pointers + off[r]);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// `spumoni run -r <prefix> -p <pattern> -P|-M -n [-c] [-w bin] [-t threads]` file-to-file
int orc_run(const char* ref_file, const char* pattern_file, int mode, int write_report, uint64_t bin_width,
            int threads, uint64_t* num_reads, double* seconds) {
    const int use_doc = (write_report >> 1) & 1;  // bit 1 of write_report: -d
    write_report &= 1;
    try {
        orc::RunOpts o;
        o.use_doc = use_doc != 0;
        o.ref_file = ref_file;
        o.pattern_file = pattern_file;
        o.ms = mode == 1;
        o.write_report = write_report != 0;
        o.bin_width = bin_width;
        o.threads = threads > 0 ? threads : 1;
        OrcHandle* h = (OrcHandle*)orc_open(ref_file, mode);
        if (!h) return 1;
        if (use_doc && orc_load_doc(h, ref_file)) { delete h; return 1; }
        size_t nr = orc::classify_reads(o.ms ? nullptr : &h->pml, o.ms ? &h->mst : nullptr, o, seconds);
        if (num_reads) *num_reads = nr;
        delete h;
        return 0;
    } catch (std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
// EmpNullDatabase(ref_file, null_reads, ...) for raw-DNA FASTA input: generate_null_pml_statistics /
// generate_null_ms_statistics (src/compute_ms_pml.cpp:1410-1486) + the constructor's summary
// (src/emp_null_database.cpp:42-79).  stats: offsets[n_reads] values; summary = {mean_null_stat, percentile_value}.
int orc_nulldb_generate(void* hv, const uint8_t* reads, const uint64_t* off, uint64_t n_reads, uint64_t* stats, double* summary) {
    OrcHandle* h = (OrcHandle*)hv;
    std::vector<size_t> output_stats;
    for (uint64_t r = 0; r < n_reads; r++) {
        std::string curr_read((const char*)reads + off[r], (size_t)(off[r + 1] - off[r]));
        std::transform(curr_read.begin(), curr_read.end(), curr_read.begin(), ::toupper);
        std::reverse(curr_read.begin(), curr_read.end());  // a reversed read is a null read
        std::vector<size_t> lengths, pointers;
        if (h->ms) h->mst.matching_statistics(curr_read.c_str(), curr_read.length(), lengths, pointers);
        else h->pml.matching_statistics(curr_read.c_str(), curr_read.length(), lengths);
        output_stats.insert(output_stats.end(), lengths.begin(), lengths.end());
    }
    std::copy(output_stats.begin(), output_stats.end(), stats);
    double sum_values = 0.0;
    for (size_t i = 0; i < output_stats.size(); i++) sum_values += output_stats[i];
    summary[0] = sum_values / output_stats.size();
    std::sort(output_stats.begin(), output_stats.end());
    size_t largest_val = 0, curr_val = output_stats.empty() ? 0 : output_stats[0];
    size_t num_occs = 0;
    for (auto x : output_stats) {
        if (x == curr_val) num_occs++;
        else {
            if (num_occs >= 5) largest_val = curr_val;
            curr_val = x;
            num_occs = 1;
        }
    }
    if (num_occs >= 5) largest_val = curr_val;
    summary[1] = (double)largest_val;
    return 0;
}
double orc_nulldb_percentile(const char* path) {
    try {
        orc::NullDb db;
        db.load(path);
        return db.percentile_value;
    } catch (std::exception& e) {
        g_err = e.what();
        return -1.0;
    }
}
int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
}  // extern "C"

#ifdef ORACLE_MAIN
// oracle CLI: spumoni_oracle run -r <prefix> -p <reads.fa> (-P|-M) -n [-c] [-w N] [-t N]
int main(int argc, char** argv) {
    std::string ref, pat;
    int mode = -1, report = 0, threads = 1;
    uint64_t bin = 150;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "run" || a == "-n") continue;
        else if (a == "-r" && i + 1 < argc) ref = argv[++i];
        else if (a == "-p" && i + 1 < argc) pat = argv[++i];
        else if (a == "-P") mode = 0;
        else if (a == "-M") mode = 1;
        else if (a == "-c") report |= 1;
        else if (a == "-d") report |= 2;
        else if (a == "-w" && i + 1 < argc) bin = std::strtoull(argv[++i], nullptr, 10);
        else if (a == "-t" && i + 1 < argc) threads = std::atoi(argv[++i]);
        else { std::fprintf(stderr, "oracle: unknown option %s\n", a.c_str()); return 1; }
    }
    if (ref.empty() || pat.empty() || mode < 0) { std::fprintf(stderr, "usage: spumoni_oracle run -r PREFIX -p READS.fa -P|-M -n [-c] [-d] [-w N] [-t N]\n"); return 1; }
    uint64_t nr = 0;
    double sec = 0;
    if (orc_run((ref + ".fa").c_str(), pat.c_str(), mode, report, bin, threads, &nr, &sec)) {
        std::fprintf(stderr, "oracle: %s\n", orc_last_error());
        return 1;
    }
    std::fprintf(stderr, "[oracle] processed %llu reads in %.3f sec (%d threads)\n", (unsigned long long)nr, sec, threads);
    return 0;
}
#endif
