// TEST INFRASTRUCTURE ONLY (see oracle/README.md): stand-ins for the third-party headers the reference includes but does
// not vendor, so that the reference's OWN translation units (/root/reference/src/compute_ms_pml.cpp, doc_array.cpp,
// emp_null_database.cpp, ks_test.cpp, batch_loader.cpp + its include/*.hpp) compile unmodified, where they lie, into
// oracle/_ref/spumoni_ref (recipe: oracle/Makefile.ref). Nothing here is derived from the reference's sources.
//
// What is stood in for:
//   sdsl-lite   int_vector<w>, bit_vector, structure_tree, util, serialize/load, nullstream
//   r-index     ri::sparse_sd_vector, ri::huff_string, ri::rle_string<>, ri::r_index<> (members the reference touches)
//   ShapedSlp   SelfShapedSlp<> / DirectAccessibleGammaCode<> / SelectSdvec<>: random access to the text; the shim loads
//               the plain text bytes from <ref>.slp instead of a grammar
//   klib        kseq (FASTA/FASTQ records over zlib)
// Each container keeps its values in plain arrays and answers rank/select/access by definition (count of c in [0, i),
// position of the i-th c, ...); the byte streams written/read are the sdsl ones (tools/sdsl_writer.hpp), so index files
// from tools/mkindex load here and files serialised here load in the product. The pin this gives: the reference's own
// query loops, classifier, writers, document-array and null-database code run for real; the succinct containers
// underneath them do not.
#pragma once
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include "../../tools/sdsl_writer.hpp"

// ---------------------------------------------------------------------------------------------- sdsl
namespace sdsl {

struct structure_tree_node {};
struct structure_tree {
    static structure_tree_node* add_child(structure_tree_node*, const std::string&, const std::string&) { return nullptr; }
    static void add_size(structure_tree_node*, uint64_t) {}
};
namespace conf {
const uint64_t SDSL_BLOCK_SIZE = (uint64_t)1 << 22;
}
class nullstream : public std::ostream {
    struct nullbuf : std::streambuf {
        int overflow(int c) override { return traits_type::not_eof(c); }
        std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
    } buf_;
public:
    nullstream() : std::ostream(&buf_) {}
};

namespace shim_detail {
inline uint64_t get_u64(std::istream& in) {
    uint64_t v = 0;
    in.read((char*)&v, 8);
    if (!in) throw std::runtime_error("shim: unexpected end of stream");
    return v;
}
inline uint8_t get_u8(std::istream& in) {
    uint8_t v = 0;
    in.read((char*)&v, 1);
    if (!in) throw std::runtime_error("shim: unexpected end of stream");
    return v;
}
inline uint16_t get_u16(std::istream& in) {
    uint16_t v = 0;
    in.read((char*)&v, 2);
    if (!in) throw std::runtime_error("shim: unexpected end of stream");
    return v;
}
// int_vector<fixed_width or 0> payload
inline sdslw::IntVec read_intvec(std::istream& in, uint8_t fixed_width) {
    sdslw::IntVec v;
    v.bit_size = get_u64(in);
    v.width = fixed_width ? fixed_width : get_u8(in);
    v.w.assign((v.bit_size + 63) / 64, 0);
    in.read((char*)v.w.data(), (std::streamsize)(v.w.size() * 8));
    if (!in) throw std::runtime_error("shim: truncated int_vector");
    v.w.push_back(0);
    return v;
}
inline void skip_select_mcl(std::istream& in) {  // an accelerator, not data
    const uint64_t args = get_u64(in);
    if (!args) return;
    read_intvec(in, 0);
    read_intvec(in, 1);
    for (uint64_t i = 0, sb = (args + 4095) >> 12; i < sb; i++) read_intvec(in, 0);
}
}  // namespace shim_detail

template <uint8_t t_width = 0>
class int_vector {
public:
    typedef uint64_t value_type;
    typedef uint64_t size_type;
    struct reference {
        int_vector* v;
        uint64_t i;
        operator uint64_t() const { return v->vals_[i]; }
        reference& operator=(uint64_t x) {
            v->vals_[i] = x & v->mask();
            return *this;
        }
        reference& operator=(const reference& o) { return *this = (uint64_t)o; }
        reference& operator++() { return *this = (uint64_t)*this + 1; }
        reference& operator+=(uint64_t x) { return *this = (uint64_t)*this + x; }
    };
    // a runtime width outside 1..64 becomes 64, as in sdsl (the reference relies on it: a one-document array asks for width 0)
    int_vector(size_type n = 0, value_type def = 0, uint8_t width = t_width ? t_width : 64)
        : width_(t_width ? t_width : (width >= 1 && width <= 64 ? width : 64)) {
        vals_.assign(n, def & mask());
    }
    size_type size() const { return vals_.size(); }
    bool empty() const { return vals_.empty(); }
    uint8_t width() const { return width_; }
    void width(uint8_t w) {
        if (!t_width) width_ = (w >= 1 && w <= 64) ? w : 64;
    }
    void resize(size_type n) { vals_.resize(n, 0); }
    reference operator[](size_type i) { return reference{this, i}; }
    value_type operator[](size_type i) const { return vals_[i]; }
    typename std::vector<uint64_t>::iterator begin() { return vals_.begin(); }
    typename std::vector<uint64_t>::iterator end() { return vals_.end(); }
    typename std::vector<uint64_t>::const_iterator begin() const { return vals_.begin(); }
    typename std::vector<uint64_t>::const_iterator end() const { return vals_.end(); }
    size_type bit_size() const { return vals_.size() * width_; }
    size_type serialize(std::ostream& out, structure_tree_node* = nullptr, std::string = "") const {
        sdslw::IntVec p(vals_.size(), width_);
        for (size_t i = 0; i < vals_.size(); i++) p.set(i, vals_[i]);
        p.write(out, t_width == 0);
        return 8 + (t_width == 0 ? 1 : 0) + p.w.size() * 8;
    }
    void load(std::istream& in) {
        sdslw::IntVec p = shim_detail::read_intvec(in, t_width);
        width_ = p.width;
        vals_.resize(p.size());
        for (size_t i = 0; i < vals_.size(); i++) vals_[i] = p.get(i);
    }
    // used by util::bit_compress
    std::vector<uint64_t> vals_;
    uint8_t width_;
    uint64_t mask() const { return width_ >= 64 ? ~0ull : ((1ull << width_) - 1); }
};
typedef int_vector<1> bit_vector;

// ones of a bit_vector in [0, i); built once from the vector it is given
template <uint8_t t_b = 1>
class rank_support_v {
public:
    rank_support_v() {}
    explicit rank_support_v(const bit_vector* bv) {
        before_.assign(bv->size() + 1, 0);
        for (uint64_t i = 0; i < bv->size(); i++) before_[i + 1] = before_[i] + (((*bv)[i] != 0) == (t_b != 0));
    }
    uint64_t rank(uint64_t i) const { return before_.at(i); }
    uint64_t operator()(uint64_t i) const { return rank(i); }
private:
    std::vector<uint64_t> before_;
};

namespace util {
template <class T>
std::string class_name(const T&) { return typeid(T).name(); }
// width := bits::hi(max) + 1
template <class V>
void bit_compress(V& v) {
    uint64_t mx = 0;
    for (uint64_t x : v.vals_) mx = std::max(mx, x);
    v.width((uint8_t)(sdslw::hi(mx) + 1));
}
template <class V>
void assign(V& a, V&& b) { a = std::move(b); }
}  // namespace util

// free serialize / load: fundamentals are written raw, classes through their members
template <class T>
typename std::enable_if<std::is_fundamental<T>::value, uint64_t>::type serialize(const T& x, std::ostream& out, structure_tree_node* = nullptr,
                                                                                 std::string = "") {
    out.write((const char*)&x, sizeof(T));
    return sizeof(T);
}
template <class T>
typename std::enable_if<!std::is_fundamental<T>::value, uint64_t>::type serialize(const T& x, std::ostream& out, structure_tree_node* v = nullptr,
                                                                                  std::string name = "") {
    return x.serialize(out, v, name);
}
template <class T>
typename std::enable_if<std::is_fundamental<T>::value, void>::type load(T& x, std::istream& in) { in.read((char*)&x, sizeof(T)); }
template <class T>
typename std::enable_if<!std::is_fundamental<T>::value, void>::type load(T& x, std::istream& in) { x.load(in); }
template <class T>
uint64_t write_member(const T& x, std::ostream& out, structure_tree_node* v = nullptr, std::string name = "") { return serialize(x, out, v, name); }
template <class T>
void read_member(T& x, std::istream& in) { load(x, in); }
template <class T>
uint64_t size_in_bytes(const T& x) {
    nullstream ns;
    return serialize(x, ns);
}
// whole-object file helpers
template <class T>
bool store_to_file(const T& x, const std::string& file) {
    std::ofstream out(file, std::ios::binary);
    if (!out) return false;
    serialize(x, out);
    return (bool)out;
}
template <class T>
bool load_from_file(T& x, const std::string& file) {
    std::ifstream in(file, std::ios::binary);
    if (!in) return false;
    load(x, in);
    return true;
}

}  // namespace sdsl

using namespace std;
using namespace sdsl;

// ---------------------------------------------------------------------------------------------- r-index
namespace ri {

typedef uint64_t ulint;
typedef unsigned char uchar;
typedef std::pair<ulint, ulint> range_t;

const uchar TERMINATOR = 1;  // bytes <= 1 in a BWT stand for the end-of-text symbol

// number of bits needed to write x (1 for x == 0)
inline uint8_t bitsize(uint64_t x) { return x == 0 ? 1 : (uint8_t)(64 - __builtin_clzll(x)); }

// bit vector of `u` bits given by the sorted positions of its ones
class sparse_sd_vector {
public:
    sparse_sd_vector() {}
    explicit sparse_sd_vector(const std::vector<bool>& b) : u_(b.size()) {
        for (size_t i = 0; i < b.size(); i++)
            if (b[i]) ones_.push_back(i);
        index();
    }
    sparse_sd_vector(const std::vector<ulint>& onset, ulint universe) : u_(universe), ones_(onset) { index(); }
    ulint size() const { return u_; }
    ulint number_of_1() const { return ones_.size(); }
    bool operator[](ulint i) const { return std::binary_search(ones_.begin(), ones_.end(), i); }
    bool at(ulint i) const { return (*this)[i]; }
    // ones in [0, i): the bucket of i's high bits bounds the search to the few ones that share them
    ulint rank(ulint i) const {
        if (ones_.empty()) return 0;
        if (i >= u_) return ones_.size();
        const ulint h = i >> shift_;
        return (ulint)(std::lower_bound(ones_.begin() + bucket_[h], ones_.begin() + bucket_[h + 1], i) - ones_.begin());
    }
    // position of the i-th one (0-based)
    ulint select(ulint i) const { return ones_.at((size_t)i); }
    // position of the last one at or before i, as a rank: number of ones in [0, i] minus one (there must be one)
    ulint predecessor_rank(ulint i) const { return rank(i + 1) - 1; }
    ulint predecessor(ulint i) const { return select(predecessor_rank(i)); }
    // distance between the i-th one and the one before it (position + 1 for the first)
    ulint gapAt(ulint i) const { return i == 0 ? select(0) + 1 : select(i) - select(i - 1); }
    ulint serialize(std::ostream& out) const {
        const std::streampos a = out.tellp();
        sdslw::write_sparse_sd(out, ones_, u_);
        const std::streampos b = out.tellp();
        if (a != std::streampos(-1) && b != std::streampos(-1)) return (ulint)(b - a);
        std::ostringstream tmp;  // streams without a position (nullstream): measure on the side
        sdslw::write_sparse_sd(tmp, ones_, u_);
        return (ulint)tmp.str().size();
    }
    void load(std::istream& in) {
        using namespace sdsl::shim_detail;
        ones_.clear();
        bucket_.clear();
        u_ = get_u64(in);
        const uint64_t declared = get_u64(in);
        if (u_ == 0) return;
        const uint64_t size = get_u64(in);
        const uint8_t wl = get_u8(in);
        sdslw::IntVec low = read_intvec(in, 0), high = read_intvec(in, 1);
        skip_select_mcl(in);
        skip_select_mcl(in);
        if (size != u_) throw std::runtime_error("shim: sd_vector size != universe");
        uint64_t seen = 0;
        for (uint64_t b = 0; b < high.bit_size; b++)
            if (high.bit(b)) {
                ones_.push_back(((b - seen) << wl) | (wl ? low.get(seen) : 0));
                seen++;
            }
        if (seen != declared) throw std::runtime_error("shim: sd_vector count mismatch");
        index();
    }
private:
    // bucket_[h] = number of ones whose value >> shift_ is < h (about one bucket per one); built with the vector, read-only afterwards
    void index() {
        shift_ = 0;
        while ((u_ >> shift_) > 2 * ones_.size() + 1) shift_++;
        std::vector<ulint> b((size_t)(u_ >> shift_) + 3, 0);
        for (ulint v : ones_) b[(size_t)(v >> shift_) + 1]++;
        for (size_t k = 1; k < b.size(); k++) b[k] += b[k - 1];
        bucket_.swap(b);
    }
    ulint u_ = 0;
    std::vector<ulint> ones_;
    std::vector<ulint> bucket_;
    uint32_t shift_ = 0;
};
typedef sparse_sd_vector sparse_hyb_vector;  // only named in a typedef of the reference, never instantiated

// byte string with rank / select per symbol (sdsl::wt_huff<> on disk)
class huff_string {
public:
    huff_string() { index(); }
    explicit huff_string(const std::string& s) : s_(s.begin(), s.end()) { index(); }
    uchar operator[](ulint i) const { return s_[(size_t)i]; }
    ulint size() const { return s_.size(); }
    // occurrences of c in [0, i): count sampled every kBlock positions + a scan of the block
    ulint rank(ulint i, uchar c) const {
        const int slot = slot_[c];
        if (slot < 0) return 0;
        const size_t blk = (size_t)(i / kBlock);
        ulint r = sampled_[blk * nslots_ + (size_t)slot];
        for (size_t j = blk * kBlock; j < i; j++) r += s_[j] == c;
        return r;
    }
    // position of the i-th c (0-based)
    ulint select(ulint i, uchar c) const { return occ_[c].at((size_t)i); }
    ulint serialize(std::ostream& out) const {
        std::ostringstream tmp;
        sdslw::write_wt_huff(tmp, s_);
        const std::string bytes = tmp.str();
        out.write(bytes.data(), (std::streamsize)bytes.size());
        return bytes.size();
    }
    void load(std::istream& in) {
        using namespace sdsl::shim_detail;
        const uint64_t size = get_u64(in), sigma = get_u64(in);
        sdslw::IntVec bv = read_intvec(in, 1);
        read_intvec(in, 64);  // rank_support_v
        skip_select_mcl(in);
        skip_select_mcl(in);
        const uint64_t nnodes = get_u64(in);
        struct Node { uint64_t pos, pos_rank; uint16_t parent, child[2]; };
        std::vector<Node> nodes((size_t)nnodes);
        for (auto& nd : nodes) {
            nd.pos = get_u64(in);
            nd.pos_rank = get_u64(in);
            nd.parent = get_u16(in);
            nd.child[0] = get_u16(in);
            nd.child[1] = get_u16(in);
        }
        uint16_t leaf_of[256];
        for (int c = 0; c < 256; c++) leaf_of[c] = get_u16(in);
        for (int c = 0; c < 256; c++) get_u64(in);  // paths
        std::vector<int> symbol((size_t)nnodes, -1);
        uint64_t leaves = 0;
        for (int c = 0; c < 256; c++)
            if (leaf_of[c] != 0xFFFF) { symbol.at(leaf_of[c]) = c; leaves++; }
        if (leaves != sigma) throw std::runtime_error("shim: wt_huff sigma mismatch");
        // level-wise decode: the bits of a node are those of its parent's sequence routed by the parent's bits
        s_.assign((size_t)size, 0);
        if (size) {
            std::vector<std::vector<uint64_t>> members((size_t)nnodes);  // text positions that reach each node, in order
            members[0].resize((size_t)size);
            for (uint64_t i = 0; i < size; i++) members[0][(size_t)i] = i;
            std::vector<size_t> order{0};
            for (size_t q = 0; q < order.size(); q++) {
                const size_t v = order[q];
                if (nodes[v].child[0] == 0xFFFF) {
                    for (uint64_t i : members[v]) s_[(size_t)i] = (uint8_t)symbol[v];
                    continue;
                }
                for (size_t j = 0; j < members[v].size(); j++)
                    members[nodes[v].child[bv.bit(nodes[v].pos + j) ? 1 : 0]].push_back(members[v][j]);
                order.push_back(nodes[v].child[0]);
                order.push_back(nodes[v].child[1]);
                std::vector<uint64_t>().swap(members[v]);
            }
        }
        index();
    }
private:
    static const size_t kBlock = 32;
    void index() {
        for (auto& p : occ_) p.clear();
        for (size_t i = 0; i < s_.size(); i++) occ_[s_[i]].push_back(i);
        nslots_ = 0;
        for (int c = 0; c < 256; c++) slot_[c] = occ_[c].empty() ? -1 : (int)nslots_++;
        sampled_.assign((s_.size() / kBlock + 1) * nslots_, 0);
        std::vector<ulint> cnt(nslots_, 0);
        for (size_t i = 0; i < s_.size(); i++) {
            if (i % kBlock == 0) std::copy(cnt.begin(), cnt.end(), sampled_.begin() + (i / kBlock) * nslots_);
            cnt[(size_t)slot_[s_[i]]]++;
        }
        if (s_.size() % kBlock == 0 && nslots_) std::copy(cnt.begin(), cnt.end(), sampled_.begin() + (s_.size() / kBlock) * nslots_);
    }
    std::vector<uint8_t> s_;
    std::vector<ulint> occ_[256];
    std::vector<ulint> sampled_;
    int slot_[256];
    size_t nslots_ = 0;
};

// run-length encoded string: heads + per-letter run ends + every B-th run end
template <class sparse_bitvector_t = sparse_sd_vector, class string_t = huff_string>
class rle_string {
public:
    rle_string() {}
    rle_string(std::string& input, ulint B = 2) { build(input, B); }
    rle_string(std::ifstream& ifs, ulint B = 2) {
        std::string s((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
        build(s, B);
    }
    uchar operator[](ulint i) { return run_heads[run_of_position(i)]; }
    ulint size() { return n; }
    ulint number_of_runs() { return R; }
    // length of the i-th run
    ulint run_at(ulint i) {
        const uchar c = run_heads[i];
        return runs_per_letter[c].gapAt(run_heads.rank(i, c));
    }
    // [first, last] positions of the j-th run
    range_t run_range(ulint j) {
        const ulint first = run_start(j);
        return {first, first + run_at(j) - 1};
    }
    // index of the run containing position i
    ulint run_of_position(ulint i) {
        ulint block = runs.rank(i);  // sampled run ends strictly before i
        ulint run = block * B;
        ulint pos = block ? runs.select(block - 1) + 1 : 0;  // first position of run `run`
        while (pos + run_at(run) <= i) pos += run_at(run++);
        return run;
    }
    // occurrences of c in [0, i)
    ulint rank(ulint i, uchar c) {
        if (runs_per_letter[c].size() == 0) return 0;
        if (i == n) return runs_per_letter[c].size();
        const ulint run = run_of_position(i);
        const ulint before = run_heads.rank(run, c);  // c-runs strictly before `run`
        const ulint inside = run_heads[run] == c ? i - run_start(run) : 0;
        return (before ? runs_per_letter[c].select(before - 1) + 1 : 0) + inside;
    }
    // position of the i-th c (0-based)
    ulint select(ulint i, uchar c) {
        const ulint j = runs_per_letter[c].rank(i);  // c-runs that end before the i-th c
        const ulint before = j ? runs_per_letter[c].select(j - 1) + 1 : 0;
        return run_start(run_heads.select(j, c)) + (i - before);
    }
    ulint serialize(std::ostream& out) {
        ulint w = 3 * sizeof(ulint);
        out.write((char*)&n, sizeof(n));
        out.write((char*)&R, sizeof(R));
        out.write((char*)&B, sizeof(B));
        if (n == 0) return w;
        w += runs.serialize(out);
        for (ulint i = 0; i < 256; i++) w += runs_per_letter[i].serialize(out);
        w += run_heads.serialize(out);
        return w;
    }
    void load(std::istream& in) {
        in.read((char*)&n, sizeof(n));
        in.read((char*)&R, sizeof(R));
        in.read((char*)&B, sizeof(B));
        if (n == 0) return;
        runs.load(in);
        runs_per_letter = std::vector<sparse_bitvector_t>(256);
        for (ulint i = 0; i < 256; i++) runs_per_letter[i].load(in);
        run_heads.load(in);
    }
protected:
    ulint run_start(ulint j) {
        ulint run = (j / B) * B;
        ulint pos = j / B ? runs.select(j / B - 1) + 1 : 0;
        while (run < j) pos += run_at(run++);
        return pos;
    }
    void build(std::string& input, ulint B_) {
        B = B_;
        n = input.size();
        R = 0;
        std::vector<std::vector<ulint>> per(256);
        std::vector<ulint> per_n(256, 0), onset;
        std::string heads;
        for (ulint i = 0; i < n;) {
            ulint j = i;
            while (j < n && input[j] == input[i]) j++;
            const uchar c = (uchar)input[i];
            if (R % B == B - 1) onset.push_back(j - 1);
            per_n[c] += j - i;
            per[c].push_back(per_n[c] - 1);
            heads.push_back(input[i]);
            R++;
            i = j;
        }
        runs = sparse_bitvector_t(onset, n);
        runs_per_letter = std::vector<sparse_bitvector_t>(256);
        for (ulint c = 0; c < 256; c++) runs_per_letter[c] = sparse_bitvector_t(per[c], per_n[c]);
        run_heads = string_t(heads);
    }
    ulint B = 0;
    sparse_bitvector_t runs;
    std::vector<sparse_bitvector_t> runs_per_letter;
    string_t run_heads;
    ulint n = 0;
    ulint R = 0;
};

// the members of r_index the reference's pml_pointers / ms_pointers inherit
template <class sparse_bv_type = sparse_sd_vector, class rle_string_t = rle_string<>>
class r_index {
public:
    r_index() {}
    ulint bwt_size() { return bwt.size(); }
    // text position following the last sample of the last run
    ulint get_last_run_sample() { return (samples_last[r - 1] + 1) % bwt.size(); }
    // F column from a plain BWT stream; bytes <= 1 count as the terminator
    std::vector<ulint> build_F(std::ifstream& ifs) {
        ifs.clear();
        ifs.seekg(0);
        F = std::vector<ulint>(256, 0);
        int c;
        ulint i = 0;
        while ((c = ifs.get()) != EOF) {
            if (c > 1) F[c]++;
            else { F[1]++; terminator_position = i; }
            i++;
        }
        for (ulint k = 255; k > 0; --k) F[k] = F[k - 1];
        F[0] = 0;
        for (ulint k = 1; k < 256; ++k) F[k] += F[k - 1];
        return F;
    }
    ulint Phi(ulint) { throw std::logic_error("shim: Phi is not provided"); }
protected:
    std::vector<ulint> F;
    rle_string_t bwt;
    ulint terminator_position = 0;
    ulint r = 0;
    sdsl::int_vector<> samples_last;
};

}  // namespace ri

using namespace ri;

// ---------------------------------------------------------------------------------------------- ShapedSlp
template <class A = void>
struct SelectSdvec {};
template <class A = void>
struct DirectAccessibleGammaCode {};
template <class var_t, class B, class C, class D>
class SelfShapedSlp {
public:
    void load(std::istream& in) { text_.assign((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>()); }
    size_t getLen() const { return text_.size(); }
    char charAt(size_t i) const { return text_[i]; }
private:
    std::string text_;
};

// ---------------------------------------------------------------------------------------------- klib kseq
struct kstring_t {
    size_t l = 0, m = 0;
    char* s = nullptr;
};
struct kseq_t {
    kstring_t name, comment, seq, qual;
    gzFile f = nullptr;
    std::string name_, comment_, seq_, qual_, pending_;
    bool have_pending = false, eof = false;
};
inline kseq_t* kseq_init(gzFile f) {
    kseq_t* k = new kseq_t;
    k->f = f;
    return k;
}
inline void kseq_destroy(kseq_t* k) { delete k; }
namespace shim_kseq {
inline bool getline(kseq_t* k, std::string& line) {
    if (k->have_pending) {
        line.swap(k->pending_);
        k->have_pending = false;
        return true;
    }
    line.clear();
    if (k->eof) return false;
    char buf[1 << 16];
    bool any = false;
    while (gzgets(k->f, buf, sizeof(buf))) {
        any = true;
        const size_t len = std::strlen(buf);
        line.append(buf, len);
        if (len && buf[len - 1] == '\n') break;
    }
    if (!any) { k->eof = true; return false; }
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
    return true;
}
}  // namespace shim_kseq
// next FASTA/FASTQ record: >= 0 sequence length, -1 end of file
inline int kseq_read(kseq_t* k) {
    std::string line;
    do {
        if (!shim_kseq::getline(k, line)) return -1;
    } while (line.empty() || (line[0] != '>' && line[0] != '@'));
    const bool fastq = line[0] == '@';
    const size_t sp = line.find_first_of(" \t");
    k->name_ = line.substr(1, sp == std::string::npos ? std::string::npos : sp - 1);
    k->comment_ = sp == std::string::npos ? "" : line.substr(sp + 1);
    k->seq_.clear();
    k->qual_.clear();
    while (shim_kseq::getline(k, line)) {
        if (!line.empty() && (line[0] == '>' || line[0] == '@' || (fastq && line[0] == '+'))) {
            if (fastq && line[0] == '+') {
                while (k->qual_.size() < k->seq_.size() && shim_kseq::getline(k, line)) k->qual_ += line;
            } else {
                k->pending_ = line;
                k->have_pending = true;
            }
            break;
        }
        for (char ch : line)
            if (!std::isspace((unsigned char)ch)) k->seq_.push_back(ch);
    }
    auto bind = [](kstring_t& ks, std::string& s) { ks.l = s.size(); ks.m = s.capacity(); ks.s = &s[0]; };
    bind(k->name, k->name_);
    bind(k->comment, k->comment_);
    bind(k->seq, k->seq_);
    bind(k->qual, k->qual_);
    return (int)k->seq_.size();
}
