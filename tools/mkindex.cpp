// spumoni_mkindex -- FIXTURE TOOL: synthetic references / reads and a stand-in for `spumoni build`.
//
// Not part of the query hot path and never linked into libspumoni_b200.so.  The reference builds
// its index by shelling out to Big-BWT / pfp_thresholds (/root/reference/src/spumoni.cpp:442-545)
// which do not exist in this image, so this tool produces the same artefacts directly:
//   <prefix>.fa                      forward + "_rev_comp" records  (refbuilder.cpp:349-379)
//   <prefix>.fa.bwt.heads/.bwt.len   run heads / 5-byte run lengths (ms_rle_string.hpp:246-247)
//   <prefix>.fa.thr_pos              5-byte thresholds per run      (thresholds_ds.hpp:405-423)
//   <prefix>.fa.ssa/.esa             5+5-byte (pos, SA) samples     (compute_ms_pml.cpp:428-436)
//   <prefix>.fa.thrbv.spumoni        pml_pointers::serialize        (compute_ms_pml.cpp:192-213)
//   <prefix>.fa.thrbv.ms             ms_pointers::serialize         (compute_ms_pml.cpp:517-542)
//   *.pmlnulldb / *.msnulldb         EmpNullDatabase::serialize     (emp_null_database.cpp:82-102)
// Threshold definition (MONI): for consecutive c-runs ending at e and starting at s, the leftmost
// t in (e, s] minimising LCP[t]; 0 for the first run of a character.
#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "sais.hpp"
#include "sdsl_writer.hpp"

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static void die(const std::string& m) {
    std::fprintf(stderr, "spumoni_mkindex: %s\n", m.c_str());
    std::exit(1);
}

struct Rng {  // splitmix64: deterministic across platforms
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    inline uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    inline double unif() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    inline uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
};
static const char ACGT[4] = {'A', 'C', 'G', 'T'};

typedef std::map<std::string, std::string> Args;
static Args parse_args(int argc, char** argv, int from) {
    Args a;
    for (int i = from; i < argc; i++) {
        std::string k = argv[i];
        if (k.rfind("--", 0) != 0) die("unexpected argument " + k);
        k = k.substr(2);
        if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0) a[k] = argv[++i];
        else a[k] = "1";
    }
    return a;
}
static std::string get(const Args& a, const std::string& k, const std::string& d = "") {
    auto it = a.find(k);
    return it == a.end() ? d : it->second;
}
static bool has(const Args& a, const std::string& k) { return a.count(k) > 0; }

// ---------------------------------------------------------------- FASTA
struct Fasta {
    std::vector<std::string> names, seqs;
};
static Fasta read_fasta(const std::string& path) {
    std::ifstream in(path);
    if (!in) die("cannot open " + path);
    Fasta f;
    std::string line;
    while (std::getline(in, line)) {
        while (!line.empty() && std::isspace((unsigned char)line.back())) line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            size_t e = line.find_first_of(" \t", 1);
            f.names.push_back(line.substr(1, e == std::string::npos ? std::string::npos : e - 1));
            f.seqs.emplace_back();
        } else {
            if (f.seqs.empty()) die("FASTA does not start with '>'");
            f.seqs.back() += line;
        }
    }
    return f;
}
static std::string revcomp(const std::string& s) {  // complement table of refbuilder.cpp:22-31 on ACGTN etc.
    static unsigned char tab[256];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 256; i++) tab[i] = (unsigned char)i;
        const char* from = "ABCDGHKMNRSTVWXYabcdghkmnrstvwxy";
        const char* to = "TVGHCDMKNYSABWXRtvghcdmknysabwxr";
        for (int i = 0; from[i]; i++) tab[(unsigned char)from[i]] = (unsigned char)to[i];
        tab['U'] = 'A'; tab['u'] = 'a';
        init = true;
    }
    std::string r(s.size(), 0);
    for (size_t i = 0; i < s.size(); i++) r[i] = (char)tab[(unsigned char)s[s.size() - 1 - i]];
    return r;
}

// ---------------------------------------------------------------- genome
static int cmd_genome(const Args& a) {
    std::string kind = get(a, "kind", "iid"), out = get(a, "out");
    uint64_t len = std::stoull(get(a, "len", "1000000"));
    uint64_t haps = std::stoull(get(a, "haps", "1"));
    double snp = std::stod(get(a, "snp", "0.001")), indel = std::stod(get(a, "indel", "0.0001"));
    Rng rng(std::stoull(get(a, "seed", "1")));
    if (out.empty()) die("genome: --out required");
    std::ofstream o(out);
    if (kind == "iid") {
        std::string s(len, 'A');
        for (auto& c : s) c = ACGT[rng.next() & 3];
        o << ">seq0\n" << s << "\n";
    } else if (kind == "pangenome") {
        uint64_t al = len / haps;
        std::string anc(al, 'A');
        for (auto& c : anc) c = ACGT[rng.next() & 3];
        for (uint64_t h = 0; h < haps; h++) {
            std::string s;
            s.reserve(al + al / 100);
            for (uint64_t i = 0; i < al; i++) {
                double r = rng.unif();
                if (r < snp) {
                    char c;
                    do c = ACGT[rng.next() & 3]; while (c == anc[i]);
                    s.push_back(c);
                } else if (r < snp + indel) {
                    uint64_t l = 1 + rng.below(5);
                    if (rng.next() & 1) { for (uint64_t k = 0; k < l; k++) s.push_back(ACGT[rng.next() & 3]); s.push_back(anc[i]); }
                    else i += l - 1;  // deletion of l bases
                } else
                    s.push_back(anc[i]);
            }
            o << ">hap" << h << "\n" << s << "\n";
        }
    } else
        die("genome: unknown --kind " + kind);
    return 0;
}

// ---------------------------------------------------------------- build
static void put5(std::ofstream& o, uint64_t v) { o.write((const char*)&v, 5); }

template <typename Idx>
static void build_rlbwt(const std::vector<uint8_t>& T /* with terminator 0 at end */, sdslw::RlbwtParts& P, bool verbose) {
    const uint64_t n = T.size();
    double t0 = now_s();
    std::vector<Idx> SA(n);
    suffix_array_sais<Idx>(T.data(), SA.data(), (Idx)n);
    if (verbose) std::fprintf(stderr, "[mkindex] suffix array: %.2f s\n", now_s() - t0);
    t0 = now_s();
    // PLCP via the Phi algorithm (chunk-parallel), then LCP in SA order
    std::vector<Idx> plcp(n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) plcp[(size_t)SA[i]] = i ? SA[i - 1] : (Idx)-1;  // phi
    {
        int nt = 1;
#ifdef _OPENMP
        nt = omp_get_max_threads();
#endif
        int nchunks = nt * 16;
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < nchunks; c++) {
            uint64_t lo = n * (uint64_t)c / nchunks, hi = n * (uint64_t)(c + 1) / nchunks;
            uint64_t h = 0;
            for (uint64_t i = lo; i < hi; i++) {
                Idx j = plcp[i];
                if (j < 0) { plcp[i] = 0; h = 0; continue; }
                while (i + h < n && (uint64_t)j + h < n && T[i + h] == T[(uint64_t)j + h]) h++;
                plcp[i] = (Idx)h;
                if (h) h--;
            }
        }
    }
    std::vector<Idx> lcp(n);
    std::vector<uint8_t> bwt(n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        lcp[i] = plcp[(size_t)SA[i]];
        bwt[i] = SA[i] ? T[(size_t)SA[i] - 1] : T[n - 1];
    }
    std::vector<Idx>().swap(plcp);
    if (verbose) std::fprintf(stderr, "[mkindex] lcp+bwt: %.2f s\n", now_s() - t0);
    t0 = now_s();
    // runs, thresholds, samples
    P.n = n;
    bool present[256] = {false};
    for (uint64_t i = 0; i < n; i++) present[bwt[i]] = true;
    std::vector<int> syms;
    for (int c = 0; c < 256; c++) if (present[c]) syms.push_back(c);
    const uint64_t INF = ~0ULL;
    uint64_t minl[256], minp[256];
    bool seen[256] = {false};
    for (int c = 0; c < 256; c++) { minl[c] = INF; minp[c] = 0; }
    uint64_t i = 0;
    while (i < n) {
        uint8_t c = bwt[i];
        uint64_t j = i;
        // position i: update all symbols with lcp[i], then read off c's threshold
        {
            uint64_t l = (uint64_t)lcp[i];
            for (int s : syms) if (l < minl[s]) { minl[s] = l; minp[s] = i; }
        }
        uint64_t thr = seen[c] ? minp[c] : 0;
        j = i + 1;
        while (j < n && bwt[j] == c) {
            uint64_t l = (uint64_t)lcp[j];
            for (int s : syms) if (l < minl[s]) { minl[s] = l; minp[s] = j; }
            j++;
        }
        P.heads.push_back(c);
        P.lens.push_back(j - i);
        P.thr.push_back(thr);
        P.ssa.push_back((uint64_t)SA[i]);
        P.esa.push_back((uint64_t)SA[j - 1]);
        seen[c] = true;
        minl[c] = INF;  // range for the next c-run starts after position j-1
        i = j;
    }
    if (verbose)
        std::fprintf(stderr, "[mkindex] runs/thresholds: %.2f s  (n=%llu r=%llu n/r=%.2f)\n", now_s() - t0,
                     (unsigned long long)n, (unsigned long long)P.heads.size(), (double)n / P.heads.size());
}

static int cmd_build(const Args& a) {
    std::string input = get(a, "input"), prefix = get(a, "prefix");
    if (input.empty() || prefix.empty()) die("build: --input and --prefix required");
    bool use_rc = !has(a, "no-revcomp"), verbose = has(a, "verbose");
    bool want_pml = has(a, "pml") || !has(a, "ms"), want_ms = has(a, "ms");
    bool keep = has(a, "keep");
    // --docs N: the records are split into N contiguous groups ("documents": the reference groups whole input
    // files, refbuilder.cpp:183-217); writes <prefix>.fa.fdi and <prefix>.fa.doc like `spumoni build -d`
    const uint64_t ndocs = std::stoull(get(a, "docs", "0"));
    Fasta f = read_fasta(input);
    if (ndocs > f.seqs.size()) die("build: --docs exceeds the number of records");
    std::vector<uint64_t> doc_lengths(ndocs, 0);
    std::string ref = prefix + ".fa";
    std::vector<uint8_t> T;
    {
        std::ofstream o(ref);
        if (!o) die("cannot write " + ref);
        for (size_t k = 0; k < f.seqs.size(); k++) {
            if (ndocs) doc_lengths[k * ndocs / f.seqs.size()] += f.seqs[k].size() * (use_rc ? 2 : 1);
            std::string s = f.seqs[k];
            for (auto& c : s) c = (char)std::toupper((unsigned char)c);
            o << '>' << f.names[k] << '\n' << s << '\n';
            T.insert(T.end(), s.begin(), s.end());
            if (use_rc) {
                std::string r = revcomp(s);
                o << '>' << f.names[k] << "_rev_comp" << '\n' << r << '\n';
                T.insert(T.end(), r.begin(), r.end());
            }
        }
    }
    for (auto c : T) if (c <= 1) die("input contains bytes <= 0x01");
    T.push_back(0);
    if (has(a, "dump-text")) {  // the indexed text + terminator as raw bytes (input of the GPU builder, spb_build_parts_device)
        std::ofstream o(get(a, "dump-text"), std::ios::binary);
        o.write((const char*)T.data(), (std::streamsize)T.size());
    }
    if (has(a, "text-only")) {  // <prefix>.fa (+ the dump) only: the run-length BWT comes from somewhere else (mkindex from-parts)
        std::printf("{\"n\": %llu}\n", (unsigned long long)T.size());
        return 0;
    }
    sdslw::RlbwtParts P;
    if (T.size() < (1ULL << 31)) build_rlbwt<int32_t>(T, P, verbose);
    else build_rlbwt<int64_t>(T, P, verbose);
    double t0 = now_s();
    if (keep) {
        std::ofstream h(ref + ".bwt.heads"), l(ref + ".bwt.len"), t(ref + ".thr_pos"), s(ref + ".ssa"), e(ref + ".esa");
        uint64_t pos = 0;
        for (size_t k = 0; k < P.heads.size(); k++) {
            h.put((char)P.heads[k]);
            put5(l, P.lens[k]);
            put5(t, P.thr[k]);
            put5(s, pos); put5(s, P.ssa[k]);
            pos += P.lens[k];
            put5(e, pos - 1); put5(e, P.esa[k]);
        }
    }
    if (want_pml) {
        std::ofstream o(ref + ".thrbv.spumoni", std::ios::binary);
        sdslw::write_pml_index(o, P);
    }
    if (want_ms) {
        std::ofstream o(ref + ".thrbv.ms", std::ios::binary);
        sdslw::write_ms_index(o, P);
    }
    if (ndocs) {
        std::ofstream fdi(ref + ".fdi");
        for (uint64_t g = 0; g < ndocs; g++) fdi << "group_" << (g + 1) << '\t' << doc_lengths[g] << '\n';
        std::ofstream o(ref + ".doc", std::ios::binary);
        sdslw::write_doc_array(o, P, doc_lengths);
    }
    if (verbose) std::fprintf(stderr, "[mkindex] serialise: %.2f s\n", now_s() - t0);
    std::printf("{\"n\": %llu, \"r\": %llu}\n", (unsigned long long)P.n, (unsigned long long)P.heads.size());
    return 0;
}

// ---------------------------------------------------------------- reads
static int cmd_reads(const Args& a) {
    std::string ref = get(a, "ref"), out = get(a, "out"), bin = get(a, "bin");
    if (ref.empty() || (out.empty() && bin.empty())) die("reads: --ref and --out/--bin required");
    uint64_t N = std::stoull(get(a, "n", "1000")), L = std::stoull(get(a, "len", "150"));
    uint64_t Lmax = std::stoull(get(a, "len-max", "0"));  // if set: lengths uniform in [len, len-max]
    double sub = std::stod(get(a, "sub", "0.01")), ins = std::stod(get(a, "ins", "0")), del = std::stod(get(a, "del", "0"));
    double neg = std::stod(get(a, "neg-frac", "0.5"));
    uint64_t seed = std::stoull(get(a, "seed", "2"));
    Fasta f = read_fasta(ref);
    std::string T;
    for (auto& s : f.seqs) T += s;
    if (T.size() < 2) die("reads: reference too short");
    std::vector<uint64_t> off(N + 1, 0);
    {
        for (uint64_t r = 0; r < N; r++) {
            uint64_t len = L;
            if (Lmax > L) { Rng g(seed * 0x100000001B3ULL + r * 2 + 1); len = L + g.below(Lmax - L + 1); }
            off[r + 1] = off[r] + len;
        }
    }
    std::vector<uint8_t> buf(off[N]);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)N; r++) {
        Rng g((seed << 32) ^ (uint64_t)r * 0x9E3779B97F4A7C15ULL);
        uint64_t len = off[r + 1] - off[r];
        uint8_t* o = buf.data() + off[r];
        if (g.unif() < neg) {
            for (uint64_t k = 0; k < len; k++) o[k] = ACGT[g.next() & 3];
        } else {
            uint64_t p = g.below(T.size());
            uint64_t k = 0;
            while (k < len) {
                if (p >= T.size()) p = 0;
                double x = g.unif();
                if (x < ins) o[k++] = ACGT[g.next() & 3];
                else if (x < ins + del) p++;
                else if (x < ins + del + sub) {
                    char c;
                    do c = ACGT[g.next() & 3]; while (c == T[p]);
                    o[k++] = (uint8_t)c; p++;
                } else o[k++] = (uint8_t)T[p++];
            }
        }
    }
    if (!out.empty()) {
        std::ofstream o(out, std::ios::binary);
        const bool fastq = has(a, "fastq");  // four-line records with a constant quality string
        std::string line, qual;
        for (uint64_t r = 0; r < N; r++) {
            line = (fastq ? "@read_" : ">read_") + std::to_string(r) + "\n";
            o.write(line.data(), (std::streamsize)line.size());
            o.write((const char*)buf.data() + off[r], (std::streamsize)(off[r + 1] - off[r]));
            o.put('\n');
            if (fastq) {
                qual.assign((size_t)(off[r + 1] - off[r]), 'I');
                o.write("+\n", 2);
                o.write(qual.data(), (std::streamsize)qual.size());
                o.put('\n');
            }
        }
    }
    if (!bin.empty()) {
        std::ofstream o(bin, std::ios::binary), q(bin + ".off", std::ios::binary);
        o.write((const char*)buf.data(), (std::streamsize)buf.size());
        q.write((const char*)off.data(), (std::streamsize)(off.size() * 8));
    }
    std::printf("{\"reads\": %llu, \"bases\": %llu}\n", (unsigned long long)N, (unsigned long long)off[N]);
    return 0;
}

// ---------------------------------------------------------------- synth
// A PML index of any size WITHOUT a suffix array: a synthetic run-length BWT (heads over ACGT with one terminator run,
// run lengths drawn to a target mean with a heavy tail) and a threshold for every run placed somewhere in the gap to the
// previous run of its letter, serialised exactly like a real one. It is not the BWT of a text, and neither the reference's
// loops nor ours care: every query is rank / select / threshold arithmetic on these arrays. Reads are produced by running
// the backward search forward in time and letting it choose the characters (so they match in stretches, like reads sampled
// with errors from an indexed genome), plus a share of random negatives. This is how the n > 2^32, r ~ 10^8 configurations get a stand-in whose
// tables have the real sizes; generation is O(r) in time and memory.
static int cmd_synth(const Args& a) {
    const std::string prefix = get(a, "prefix"), bin = get(a, "bin"), out = get(a, "out");
    if (prefix.empty()) die("synth: --prefix required");
    const uint64_t R = std::stoull(get(a, "runs", "100000")), seed = std::stoull(get(a, "seed", "1"));
    const double mean_len = std::stod(get(a, "mean-len", "20")), tail = std::stod(get(a, "tail-frac", "0.002"));
    const uint64_t tail_mul = std::stoull(get(a, "tail-mul", "500"));
    const uint64_t N = std::stoull(get(a, "n", "0")), L = std::stoull(get(a, "len", "1000"));
    const double sub = std::stod(get(a, "sub", "0.08")), neg = std::stod(get(a, "neg-frac", "0.1"));
    const uint64_t read_seed = std::stoull(get(a, "read-seed", "0"));  // varies the reads without touching the index
    if (R < 8) die("synth: --runs must be at least 8");
    sdslw::RlbwtParts P;
    P.heads.resize(R);
    P.lens.resize(R);
    P.thr.assign(R, 0);
    Rng g(seed * 0x9E3779B97F4A7C15ULL + 17);
    const uint64_t term_run = 1 + g.below(R - 2);
    uint8_t prev = 0;
    P.n = 0;
    for (uint64_t k = 0; k < R; k++) {
        if (k == term_run) { P.heads[k] = 0; P.lens[k] = 1; prev = 0; P.n += 1; continue; }
        uint8_t c;
        do c = (uint8_t)ACGT[g.next() & 3]; while (c == prev);
        prev = c;
        // geometric with the requested mean, a small share of runs `tail_mul` times longer
        uint64_t len = 1 + (uint64_t)(-std::log(1.0 - g.unif()) * (mean_len - 1.0));
        if (g.unif() < tail) len *= tail_mul;
        P.heads[k] = c;
        P.lens[k] = len;
        P.n += len;
    }
    if (P.n >= (1ULL << 40) - 2) die("synth: n >= 2^40 - 2");
    // thresholds: a position in (end of the previous run of the letter, start of this run]
    std::vector<uint64_t> start(R + 1);
    {
        uint64_t pos = 0, last_end[256];
        bool seen[256] = {false};
        for (uint64_t k = 0; k < R; k++) {
            start[k] = pos;
            const uint8_t c = P.heads[k] <= 1 ? 1 : P.heads[k];
            if (seen[c]) P.thr[k] = last_end[c] + 1 + g.below(pos - last_end[c]);
            seen[c] = true;
            pos += P.lens[k];
            last_end[c] = pos - 1;
        }
        start[R] = pos;
    }
    if (!has(a, "reads-only")) {  // --reads-only: the index of these parameters exists already, only (other) reads are wanted
        std::ofstream fa(prefix + ".fa");
        fa << ">synthetic_rlbwt_no_text\nN\n";
        std::ofstream o(prefix + ".fa.thrbv.spumoni", std::ios::binary);
        sdslw::write_pml_index(o, P);
    }
    std::printf("{\"n\": %llu, \"r\": %llu}\n", (unsigned long long)P.n, (unsigned long long)R);
    if (!N) return 0;
    // LF as a move structure: lf0[k] = LF(start of run k), dest[k] = run holding it (one forward merge per letter)
    std::vector<uint64_t> F;
    sdslw::build_F(P.heads, P.lens, F);
    std::vector<uint64_t> lf0(R);
    std::vector<uint32_t> dest(R);
    {
        uint64_t cnt[256] = {0}, cur[256];
        for (int c = 0; c < 256; c++) cur[c] = ~0ull;
        for (uint64_t k = 0; k < R; k++) {
            const uint8_t c = P.heads[k] <= 1 ? 1 : P.heads[k];
            lf0[k] = F[c] + cnt[c];
            cnt[c] += P.lens[k];
            uint64_t i = cur[c];
            if (i == ~0ull) i = (uint64_t)(std::upper_bound(start.begin(), start.end(), lf0[k]) - start.begin()) - 1;
            else while (start[i + 1] <= lf0[k]) i++;
            cur[c] = i;
            dest[k] = (uint32_t)i;
        }
    }
    std::vector<uint64_t> off(N + 1, 0);
    for (uint64_t r = 0; r < N; r++) off[r + 1] = off[r] + L;
    std::vector<uint8_t> buf(off[N]);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < (int64_t)N; r++) {
        Rng q(((seed + read_seed) << 32) ^ ((uint64_t)r + 1) * 0xD1B54A32D192ED03ULL);
        uint8_t* o = buf.data() + off[r];
        if (q.unif() < neg) {
            for (uint64_t i = 0; i < L; i++) o[i] = (uint8_t)ACGT[q.next() & 3];
            continue;
        }
        // Follow the backward search itself (compute_ms_pml.cpp:237-286) from its initial state pos = n - 1 and let it
        // decide the read: with probability 1 - sub the next character is the one under the cursor (a match), otherwise
        // another letter, and the cursor moves as the search would (threshold of the next run of that letter decides
        // between the run above and the run below). Matches then come in stretches of mean length 1 / sub, as for reads
        // sampled with errors from an indexed text.
        uint64_t k = R - 1, in = P.lens[R - 1] - 1;
        for (uint64_t i = L; i-- > 0;) {
            uint8_t c = P.heads[k];
            if (c <= 1 || q.unif() < sub) {
                uint8_t d;
                do d = (uint8_t)ACGT[q.next() & 3]; while (d == c);
                c = d;
                uint64_t dn = k + 1, up = k;
                while (dn < R && P.heads[dn] != c) dn++;
                bool has_up = false;
                while (up-- > 0)
                    if (P.heads[up] == c) { has_up = true; break; }
                const bool go_up = dn >= R || (has_up && start[k] + in < P.thr[dn]);
                if (go_up && has_up) { k = up; in = P.lens[up] - 1; }
                else if (dn < R) { k = dn; in = 0; }
                else { k = term_run + 1 < R ? term_run + 1 : 0; in = 0; c = P.heads[k]; }  // letter without runs: cannot happen for ACGT at this size
            }
            o[i] = c;
            const uint64_t t = dest[k];
            in += lf0[k] - start[t];  // LF(start[k] + in) = lf0[k] + in, expressed relative to run t
            uint64_t tt = t;
            while (in >= P.lens[tt]) in -= P.lens[tt++];
            k = tt;
        }
    }
    if (!out.empty()) {
        std::ofstream o(out, std::ios::binary);
        for (uint64_t r = 0; r < N; r++) {
            const std::string line = ">read_" + std::to_string(r) + "\n";
            o.write(line.data(), (std::streamsize)line.size());
            o.write((const char*)buf.data() + off[r], (std::streamsize)L);
            o.put('\n');
        }
    }
    if (!bin.empty()) {
        std::ofstream o(bin, std::ios::binary), q(bin + ".off", std::ios::binary);
        o.write((const char*)buf.data(), (std::streamsize)buf.size());
        q.write((const char*)off.data(), (std::streamsize)(off.size() * 8));
    }
    return 0;
}

// ---------------------------------------------------------------- nulldb
static int cmd_nulldb(const Args& a) {
    std::string stats = get(a, "stats"), out = get(a, "out");
    if (stats.empty() || out.empty()) die("nulldb: --stats (u64 LE binary) and --out required");
    std::ifstream in(stats, std::ios::binary | std::ios::ate);
    if (!in) die("cannot open " + stats);
    size_t bytes = (size_t)in.tellg();
    in.seekg(0);
    std::vector<uint64_t> v(bytes / 8);
    in.read((char*)v.data(), (std::streamsize)(v.size() * 8));
    std::ofstream o(out, std::ios::binary);
    sdslw::write_null_db(o, v, std::stod(get(a, "ks-thr", "0.0")));
    return 0;
}

// ---------------------------------------------------------------- from-parts
// What build_spumoni_main / build_spumoni_ms_main do (/root/reference/src/compute_ms_pml.cpp:1384-1408): construct the
// index objects from the PFP intermediates <prefix>.fa.{bwt.heads,bwt.len,thr_pos,ssa,esa} (5-byte integers,
// compute_ms_pml.cpp:44-60, :404-439; thresholds_ds.hpp:393-426) and serialise them.  Lets a test hand-craft run-length
// data (e.g. an index whose positions exceed 2^32) without suffix-sorting a text of that size.
static uint64_t get5(std::ifstream& f) { uint64_t v = 0; f.read((char*)&v, 5); return v; }
static int cmd_from_parts(const Args& a) {
    std::string prefix = get(a, "prefix");
    if (prefix.empty()) die("from-parts: --prefix required");
    const std::string ref = prefix + ".fa";
    bool want_ms = has(a, "ms"), want_pml = has(a, "pml") || !want_ms;
    std::ifstream h(ref + ".bwt.heads", std::ios::binary), l(ref + ".bwt.len", std::ios::binary), t(ref + ".thr_pos", std::ios::binary),
        s(ref + ".ssa", std::ios::binary), e(ref + ".esa", std::ios::binary);
    if (!h || !l || !t) die("from-parts: missing " + ref + ".bwt.heads/.bwt.len/.thr_pos");
    sdslw::RlbwtParts P;
    for (int c; (c = h.get()) != EOF;) P.heads.push_back((uint8_t)c);
    for (size_t k = 0; k < P.heads.size(); k++) {
        P.lens.push_back(get5(l));
        P.thr.push_back(get5(t));
        P.n += P.lens.back();
        if (s && e) { get5(s); P.ssa.push_back(get5(s)); get5(e); P.esa.push_back(get5(e)); }
    }
    if (!l || !t) die("from-parts: .bwt.len / .thr_pos shorter than .bwt.heads");
    if (want_pml) { std::ofstream o(ref + ".thrbv.spumoni", std::ios::binary); sdslw::write_pml_index(o, P); }
    if (want_ms) {
        if (P.ssa.size() != P.heads.size()) die("from-parts: --ms needs .ssa/.esa");
        std::ofstream o(ref + ".thrbv.ms", std::ios::binary);
        sdslw::write_ms_index(o, P);
    }
    std::printf("{\"n\": %llu, \"r\": %llu}\n", (unsigned long long)P.n, (unsigned long long)P.heads.size());
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) die("usage: spumoni_mkindex genome|build|from-parts|reads|nulldb|synth [--key value ...]");
    std::string cmd = argv[1];
    Args a = parse_args(argc, argv, 2);
#ifdef _OPENMP
    if (has(a, "threads")) omp_set_num_threads(std::stoi(get(a, "threads")));
#endif
    if (cmd == "genome") return cmd_genome(a);
    if (cmd == "build") return cmd_build(a);
    if (cmd == "from-parts") return cmd_from_parts(a);
    if (cmd == "reads") return cmd_reads(a);
    if (cmd == "nulldb") return cmd_nulldb(a);
    if (cmd == "synth") return cmd_synth(a);
    die("unknown command " + cmd);
    return 1;
}
