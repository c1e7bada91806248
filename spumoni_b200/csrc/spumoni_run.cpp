// `spumoni run` -- drop-in command line for the query side of SPUMONI, on top of the C ABI.
//
// Keeps the reference's process seam:
//   flags / validation      /root/reference/src/spumoni.cpp:163-206, include/spumoni_main.hpp:233-330
//   run_main                /root/reference/src/spumoni.cpp:733-755
//   run_spumoni_main        /root/reference/src/compute_ms_pml.cpp:1305-1342   (PML)
//   run_spumoni_ms_main     /root/reference/src/compute_ms_pml.cpp:1344-1382   (MS)
//   classify_reads_pml/_ms  /root/reference/src/compute_ms_pml.cpp:845-1034, :1036-1217
//   BatchLoader             /root/reference/src/batch_loader.cpp:26-131 (FASTA path, id quirk :104-106)
// Output files (<reads>.pseudo_lengths | .lengths + .pointers | .report) are byte-identical to the
// reference's single-thread output.  Everything per-base runs on the GPU through
// libspumoni_b200.so; this file only parses FASTA, batches reads, and formats results.
// The document array (-d) writes <reads>.doc_numbers (:859, :1003-1006).
// Not supported (outside SURVEY section 8): minimizer digestion (-m/-a, needs bonsai).
#include <fcntl.h>
#include <getopt.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <sstream>
#include <map>
#include <string>
#include <thread>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/spumoni_b200.h"

#define SPUMONI_VERSION "2.0.9"  // version of the reference this CLI mirrors (spumoni_main.hpp:24)

static void fatal_warning(const std::string& m) { std::fprintf(stderr, "Warning: %s\n\n", m.c_str()); std::exit(1); }
static void fatal_error(const std::string& m) { std::fprintf(stderr, "\n\033[31mError: \033[0m%s\n\n", m.c_str()); std::exit(1); }
static void force_log(const char* func, const std::string& m) { std::fprintf(stderr, "\033[32m[%s] \033[0m%s\n", func, m.c_str()); }
static void status_log(const char* func, const std::string& m) { std::fprintf(stderr, "\033[32m[%s] \033[0m%s ... ", func, m.c_str()); }
static void done_log(double sec) { std::fprintf(stderr, "done.  (%.3f sec)\n", sec); }
static double now_s() { return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); }
static bool is_file(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
static bool ends_with(const std::string& s, const std::string& suf) { return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; }
static void spb_or_die(int rc) { if (rc) fatal_error(spb_last_error()); }

struct RunOptions {  // SpumoniRunOptions, spumoni_main.hpp:233-249
    std::string ref_file, pattern_file;
    bool ms_requested = false, pml_requested = false;
    size_t threads = 1;
    bool use_doc = false, write_report = false, min_digest = true, use_promotions = false, use_dna_letters = false,
         is_general_text = false;
    size_t k = 4, w = 11, bin_size = 150;
};

static int run_usage() {
    std::fprintf(stderr, "spumoni run - Uses a spumoni index to compute MS/PML of patterns w.r.t. a reference.\n");
    std::fprintf(stderr, "Usage: spumoni run [options]\n\n");
    std::fprintf(stderr, "Options:\n\tGeneral options:\n");
    std::fprintf(stderr, "\t%-35sprints this usage message\n", "-h, --help");
    std::fprintf(stderr, "\t%-25s%-10snumber of helper threads (default: 1)\n\n", "-t, --threads", "[INT]");
    std::fprintf(stderr, "\tInput/output options:\n");
    std::fprintf(stderr, "\t%-25s%-10soutput prefix used for index\n", "-r, --ref", "[FILE]");
    std::fprintf(stderr, "\t%-25s%-10spath to patterns file that will be used.\n", "-p, --pattern", "[FILE]");
    std::fprintf(stderr, "\t%-25s%-10suse index to compute MSs\n", "-M, --MS", "");
    std::fprintf(stderr, "\t%-25s%-10suse index to compute PMLs\n", "-P, --PML", "");
    std::fprintf(stderr, "\t%-25s%-10spattern file is general text (default: FASTA)\n", "-g, --general", "");
    std::fprintf(stderr, "\t%-25s%-10suse document array to get assignments\n", "-d, --doc-array", "");
    std::fprintf(stderr, "\t%-25s%-10swrite out the classifications in a report file\n", "-c, --classify", "");
    std::fprintf(stderr, "\t%-25s%-10ssize of region in bp for classification (default: 150)\n\n", "-w, --window", "[INT]");
    std::fprintf(stderr, "\tMinimizer options:\n");
    std::fprintf(stderr, "\t%-25s%-10sturn off minimizer digestion of reads (default: on)\n", "-n, --no-digest", "");
    std::fprintf(stderr, "\t%-25s%-10suse alphabet-promoted minimizers\n", "-m, --minimizer-alphabet", "");
    std::fprintf(stderr, "\t%-25s%-10suse DNA-letter based minimizers\n", "-a, --dna-minimizer", "");
    std::fprintf(stderr, "\t%-25s%-10ssmall window size (k) for finding minimizers (default: 4)\n", "-K, --small-window", "[INT]");
    std::fprintf(stderr, "\t%-25s%-10slarge window size (w) for finding minimizers (default: 11)\n\n", "-W, --large-window", "[INT]");
    return 0;
}

static void parse_run_options(int argc, char** argv, RunOptions* o) {
    static struct option long_options[] = {
        {"help", no_argument, NULL, 'h'},          {"threads", required_argument, NULL, 't'},
        {"ref", required_argument, NULL, 'r'},     {"pattern", required_argument, NULL, 'p'},
        {"MS", no_argument, NULL, 'M'},            {"PML", no_argument, NULL, 'P'},
        {"general-text", no_argument, NULL, 'g'},  {"doc-array", no_argument, NULL, 'd'},
        {"classify", no_argument, NULL, 'c'},      {"window", required_argument, NULL, 'w'},
        {"no-digest", no_argument, NULL, 'n'},     {"minimizer-alphabet", no_argument, NULL, 'm'},
        {"dna-minimizer", no_argument, NULL, 't'}, {"small-window", required_argument, NULL, 'K'},
        {"large-window", required_argument, NULL, 'W'}, {0, 0, 0, 0}};
    int long_index = 0;
    for (int c; (c = getopt_long(argc, argv, "hr:p:MPt:dcnmaK:W:w:g", long_options, &long_index)) >= 0;) {
        switch (c) {
            case 'h': run_usage(); std::exit(1);
            case 'r': o->ref_file.assign(optarg); break;
            case 'p': o->pattern_file.assign(optarg); break;
            case 'M': o->ms_requested = true; break;
            case 'P': o->pml_requested = true; break;
            case 'c': o->write_report = true; break;
            case 'm': o->use_promotions = true; break;
            case 'a': o->use_dna_letters = true; break;
            case 'n': o->min_digest = false; break;
            case 'K': o->k = std::max(std::atoi(optarg), 1); break;
            case 'W': o->w = std::max(std::atoi(optarg), 1); break;
            case 'w': o->bin_size = std::max(std::atoi(optarg), 1); break;
            case 'g': o->is_general_text = true; break;
            case 't': o->threads = std::max(std::atoi(optarg), 1); break;
            case 'd': o->use_doc = true; break;
            default: run_usage(); std::exit(1);
        }
    }
}

static void validate(const RunOptions& o, bool ms) {  // SpumoniRunOptions::validate, spumoni_main.hpp:267-329
    if (o.ref_file == "" || o.pattern_file == "") fatal_warning("Both a reference file (-r) and pattern file (-p) must be provided.");
    if (o.ms_requested == o.pml_requested) fatal_warning("An output type with -M or -P must be specified, only one can be used at a time.");
    const std::string ext = o.use_promotions ? ".bin" : ".fa";
    if (!is_file(o.ref_file + ext)) fatal_error("The following path is not valid: " + o.ref_file + ext + " (remember to only specify output prefix)");
    if (!is_file(o.pattern_file)) fatal_error("The following path is not valid: " + o.pattern_file);
    if (!o.is_general_text && !ends_with(o.pattern_file, ".fa") && !ends_with(o.pattern_file, ".fasta") && !ends_with(o.pattern_file, ".fna"))
        fatal_error("The pattern file provided does not appear to be a FASTA\n       file, please convert to FASTA and re-run.");
    if (o.is_general_text && o.min_digest) fatal_warning("For general-text querying, minimizer digestion must be turned off with -n.");
    if (o.is_general_text && o.threads > 1) fatal_warning("For general-text querying, multi-threading is not available.");
    if (o.is_general_text && o.write_report) fatal_warning("For general-text querying, classification is not available.");
    if (o.use_doc && !is_file(o.ref_file + ext + ".doc"))  // spumoni_main.hpp:301-302
        fatal_warning("document array file (" + o.ref_file + ext + ".doc) is not present, so it cannot be used.");
    if (!is_file(o.ref_file + ext + (ms ? ".thrbv.ms" : ".thrbv.spumoni")))
        fatal_warning("The index required for this computation is not available, please use spumoni build.");
    if (o.k > 4) fatal_warning("small window size (k) cannot be larger than 4 characters.");
    if (o.w < o.k) fatal_warning("large window size (w) should be larger than the small window size (k)");
    if (o.min_digest) {
        if (o.use_promotions && o.use_dna_letters) fatal_error("Only one type of minimizer can be specified from either -m or -a.");
        if (!o.use_promotions && !o.use_dna_letters) fatal_error("A minimizer type must be specified using -m or -a.");
        fatal_error("minimizer digestion (-m/-a) is not supported by spumoni_b200; build and query with -n.");
    } else if (o.use_promotions || o.use_dna_letters)
        fatal_error("A minimizer type should not be specified if intending not to use minimizer digestion.");
    if (o.bin_size < 50 || o.bin_size > 400) fatal_warning("the bin size used is not optimal. Re-run using a value between 50 and 400.");
}

// ---------------------------------------------------------------------------------- input
struct ReadBatch {
    std::vector<std::string> ids;
    std::vector<uint64_t> offsets{0};
    std::vector<uint8_t> bases;
    uint64_t max_len = 0;  // longest read of the batch
    void clear() { ids.clear(); offsets.assign(1, 0); bases.clear(); max_len = 0; }
};

// std::getline semantics (lines split at '\n', a last line without terminator still counts) over 16 MiB reads
class LineReader {
public:
    explicit LineReader(const std::string& path, uint64_t start = 0) : f_(std::fopen(path.c_str(), "rb")), buf_(16u << 20) {
        if (f_ && start) fseeko(f_, (off_t)start, SEEK_SET);
    }
    ~LineReader() { if (f_) std::fclose(f_); }
    bool good() const { return f_ != nullptr; }
    // what std::istream::good() would say after the reads so far: false once a line had to be ended by the end of the file
    bool stream_good() const { return f_ != nullptr && !hit_eof_; }
    int peek() { return fill() ? (unsigned char)buf_[pos_] : EOF; }
    bool getline(std::string& line) {
        line.clear();
        bool any = false;
        for (;;) {
            if (!fill()) { hit_eof_ = true; return any; }
            any = true;
            const char* p = buf_.data() + pos_;
            const char* nl = (const char*)std::memchr(p, '\n', end_ - pos_);
            if (nl) { line.append(p, (size_t)(nl - p)); pos_ = (size_t)(nl - buf_.data()) + 1; return true; }
            line.append(p, end_ - pos_);
            pos_ = end_;
        }
    }
private:
    bool fill() {
        if (pos_ < end_) return true;
        if (!f_ || eof_) return false;
        end_ = std::fread(buf_.data(), 1, buf_.size(), f_);
        pos_ = 0;
        if (end_ == 0) eof_ = true;
        return end_ > 0;
    }
    std::FILE* f_;
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false, hit_eof_ = false;
};

class FastaSource {  // sequential equivalent of BatchLoader::loadBatch + grabNextRead (FASTA, and FASTQ records)
public:
    // start: offset of a record's '>', or of the first line of a FASTQ batch (fast path hand-over; resume_fastq says which --
    // a batch may begin with an empty line, so the first byte cannot tell)
    explicit FastaSource(const std::string& path, uint64_t start = 0, bool resume_fastq = false) : in_(path, start) {
        if (!in_.good()) return;
        int pk = in_.peek();
        if (pk == EOF) { done_ = true; return; }
        if (resume_fastq) { fastq_ = true; return; }
        if (pk == '@') { fastq_ = true; return; }
        if (pk != '>') fatal_error("unrecognized input query file type - expects FASTA or FASTQ.");
        if (!in_.getline(line_)) done_ = true;
    }
    // appends one record; false at end of input
    bool next(ReadBatch& b) {
        if (fastq_) return next_fastq(b);
        if (done_ || line_.empty()) return false;  // an empty header line ends the input (batch_loader.cpp:87)
        if (line_[0] != '>') fatal_error(std::string("Incorrect FASTA entry, it should start with '>' but found ") + line_[0]);
        if (line_.size() <= 2) fatal_error("header line is missing an id. invalid query cannot be processed.");
        size_t idl = line_.find_first_of(" \t\r", 1);
        idl = (idl == std::string::npos) ? line_.size() : idl;
        b.ids.push_back(line_.substr(1, idl));  // keeps the delimiter (batch_loader.cpp:104-106)
        const size_t before = b.bases.size();
        for (;;) {
            if (!in_.getline(line_)) { done_ = true; break; }
            if (!line_.empty() && line_[0] == '>') break;
            size_t e = line_.size();
            while (e > 0 && std::isspace((unsigned char)line_[e - 1])) e--;
            b.bases.insert(b.bases.end(), line_.begin(), line_.begin() + e);
        }
        if (b.bases.size() == before) {  // compute_ms_pml.cpp:926-931
            std::cout << "\n\n";
            fatal_warning(b.ids.back() + " was empty after digestion, commonly due to reads consisting of mostly non-ACGT "
                          "characters. Please remove read or run SPUMONI without minimizer digestion.");
        }
        b.offsets.push_back(b.bases.size());
        if (b.bases.size() - before > b.max_len) b.max_len = b.bases.size() - before;
        return true;
    }
private:
    // FASTQ (batch_loader.cpp:33-34, :61-66, :108-120), kept batch for batch because the batches are observable: the
    // reference gathers lines until the records seen so far hold >= 1000 bases by its estimate (half of the four lines'
    // lengths, :903), and when the file ends inside that loop -- always, if the last line is newline-terminated and the
    // estimate is still below 1000 -- loadBatch returns false and the gathered records are never processed (:51). Inside
    // a batch an empty header line, or a record cut short, ends the batch.
    bool load_fastq_batch() {
        lines_.clear();
        next_line_ = 0;
        size_t covered = 0, record = 0;
        bool valid = false;
        while (in_.stream_good() && covered < 1000) {
            std::string line;
            if (!in_.getline(line)) { lines_.clear(); return false; }
            record += line.size();
            valid = true;
            lines_.push_back(std::move(line));
            if (lines_.size() % 4 == 0) { covered += record / 2; record = 0; }
        }
        return valid;
    }
    bool next_fastq(ReadBatch& b) {
        for (;;) {
            if (done_) return false;
            if (next_line_ >= lines_.size()) {
                if (!load_fastq_batch()) { done_ = true; return false; }
                continue;
            }
            const std::string& hd = lines_[next_line_++];
            if (hd.empty()) { next_line_ = lines_.size(); continue; }  // the rest of the batch is skipped (:87)
            if (hd[0] != '@') fatal_error(std::string("Incorrect FASTQ entry, it should start with '@' but found ") + hd[0]);
            if (hd.size() <= 2) fatal_error("header line is missing an id. invalid query cannot be processed.");
            if (next_line_ + 2 >= lines_.size()) { next_line_ = lines_.size(); continue; }  // record cut short by the end of the file
            size_t idl = hd.find_first_of(" \t\r", 1);
            idl = (idl == std::string::npos) ? hd.size() : idl;
            b.ids.push_back(hd.substr(1, idl));
            const std::string& seq = lines_[next_line_];
            next_line_ += 3;  // sequence, '+' line, qualities
            size_t e = seq.size();
            while (e > 0 && std::isspace((unsigned char)seq[e - 1])) e--;
            if (e == 0) {  // compute_ms_pml.cpp:926-931
                std::cout << "\n\n";
                fatal_warning(b.ids.back() + " was empty after digestion, commonly due to reads consisting of mostly non-ACGT "
                              "characters. Please remove read or run SPUMONI without minimizer digestion.");
            }
            b.bases.insert(b.bases.end(), seq.begin(), seq.begin() + e);
            b.offsets.push_back(b.bases.size());
            if (e > b.max_len) b.max_len = e;
            return true;
        }
    }
    LineReader in_;
    std::string line_;
    bool done_ = false, fastq_ = false;
    std::vector<std::string> lines_;  // FASTQ: the lines of the current batch
    size_t next_line_ = 0;
};

// ---------------------------------------------------------------------------------- output
static inline char* put_uint(char* p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

static inline uint32_t digits10(uint64_t v) {
    uint32_t d = 1;
    while (v >= 10) { v /= 10; d++; }
    return d;
}

// One output file's text for a batch: per read '>' id '\n' then every value followed by one space, then '\n'
// (compute_ms_pml.cpp:1003-1010).  Two passes over the values on -t threads: exact sizes, then the bytes, into one
// contiguous buffer that is reused between batches and written with a single call.
struct TextBuffer {
    std::unique_ptr<char[]> data;
    size_t cap = 0, size = 0;
    std::vector<uint64_t> start;  // per read
    void reserve(size_t n) {
        if (n > cap) { cap = n + n / 8; data.reset(new char[cap]); }
    }
};

template <typename T>
static void format_values(const ReadBatch& b, const T* vals, int threads, TextBuffer& out) {
    const size_t nr = b.ids.size();
    out.start.resize(nr + 1);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t r = 0; r < (int64_t)nr; r++) {
        const uint64_t lo = b.offsets[r], hi = b.offsets[r + 1];
        uint64_t bytes = b.ids[r].size() + 3 + (hi - lo);  // '>' id '\n' ... '\n' + one space per value
        for (uint64_t i = lo; i < hi; i++) bytes += vals[i] < 10 ? 1 : digits10((uint64_t)vals[i]);
        out.start[r + 1] = bytes;
    }
    out.start[0] = 0;
    for (size_t r = 0; r < nr; r++) out.start[r + 1] += out.start[r];
    out.size = out.start[nr];
    out.reserve(out.size);
    char* base = out.data.get();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t r = 0; r < (int64_t)nr; r++) {
        const uint64_t lo = b.offsets[r], hi = b.offsets[r + 1];
        char* p = base + out.start[r];
        *p++ = '>';
        std::memcpy(p, b.ids[r].data(), b.ids[r].size());
        p += b.ids[r].size();
        *p++ = '\n';
        for (uint64_t i = lo; i < hi; i++) {
            const uint64_t v = (uint64_t)vals[i];
            if (v < 10) *p++ = (char)('0' + v); else p = put_uint(p, v);
            *p++ = ' ';
        }
        *p++ = '\n';
    }
}

static void write_all(std::ofstream& f, const TextBuffer& t) { f.write(t.data.get(), (std::streamsize)t.size); }

static void report_header(std::ofstream& f, size_t thr) {  // compute_ms_pml.cpp:877-886
    f.precision(4);
    f << std::setw(30) << std::left << "read id:" << std::setw(15) << std::left << "status:" << std::setw(19) << std::left
      << "avg max-value (thr=" << std::setw(2) << std::left << thr << std::setw(5) << std::left << "):" << std::setw(12)
      << std::left << "above thr:" << std::setw(12) << std::left << "below thr:" << std::endl;
}
static void report_row(std::ostream& f, const std::string& id, const spb_bin_stats& s) {  // :1012-1020
    const size_t nbins = (size_t)s.above + s.below;
    const bool found = (s.above / (s.above + s.below + 0.0) > 0.50);
    f.precision(3);
    f << std::setw(30) << std::left << id << std::setw(15) << std::left << (found ? "FOUND" : "NOT_PRESENT") << std::setw(26)
      << std::left << (s.sum_max + 0.0) / nbins << std::setw(12) << std::left << (size_t)s.above << std::setw(12) << std::left
      << (size_t)s.below << std::endl;
}

// ---------------------------------------------------------------------------------- driver
// Three stages run concurrently on batches of ~64 Mi bases: a parser thread (FASTA -> batch), the calling thread
// (GPU query through the C ABI) and a writer thread (decimal formatting on -t threads + file writes, in input
// order -- the reference's own order for -t 1, App. C-11).
struct Outputs {
    std::ofstream lengths, pointers, report, docs;
};

// The device-formatted text of a batch lands here; allocated once per Work and reused. SPUMONI_B200_PINNED_TEXT=1 page-locks
// it (spb_host_alloc) so that the D2H copies of the sub-batches are real DMA transfers: measured on 10^9 bases, pinning four
// 200 MB arenas costs about what the faster copies save (query stage 0.79-0.86 s against 0.71-0.90 s), so it is opt-in and
// pays only on inputs of many batches.
class TextArena {
public:
    TextArena() = default;
    TextArena(const TextArena&) = delete;
    TextArena& operator=(const TextArena&) = delete;
    ~TextArena() { release(); }
    void reset(uint64_t bytes) {
        release();
        static const bool pin = std::getenv("SPUMONI_B200_PINNED_TEXT") != nullptr;
        void* q = nullptr;
        if (pin && spb_host_alloc(&q, (size_t)bytes) == SPB_OK) { p_ = (char*)q; pinned_ = true; }
        else { p_ = new char[bytes]; pinned_ = false; }
    }
    char* get() const { return p_; }
private:
    void release() {
        if (!p_) return;
        if (pinned_) spb_host_free(p_);
        else delete[] p_;
        p_ = nullptr;
    }
    char* p_ = nullptr;
    bool pinned_ = false;
};

struct Work {
    ReadBatch b;
    std::vector<uint32_t> L;
    std::vector<uint16_t> L16;  // lengths travel as uint16 when every read of the batch is < 65536 bases (lossless)
    std::vector<uint64_t> P;
    std::vector<spb_bin_stats> S;
    std::vector<uint32_t> D;
    bool narrow = false;
    // PML without -d: the .pseudo_lengths bytes come formatted from the device (spb_pml_query_host_text)
    std::vector<uint8_t> idcat;
    std::vector<uint64_t> idoff;
    TextArena text;
    uint64_t text_cap = 0, text_bytes = 0;
    bool device_text = false;
};

class WorkQueue {  // bounded hand-off between two stages; nullptr = end of stream
public:
    void push(Work* w) {
        std::unique_lock<std::mutex> lk(m_);
        q_.push_back(w);
        cv_.notify_all();
    }
    Work* pop() {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return !q_.empty(); });
        Work* w = q_.front();
        q_.pop_front();
        return w;
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Work*> q_;
};

static void query_batch(spb_index* idx, bool ms, const RunOptions& o, spb_query_opts qo, Work& w) {
    ReadBatch& b = w.b;
    const size_t nr = b.ids.size();
    w.device_text = false;
    if (!nr) return;
    static const bool host_format = std::getenv("SPUMONI_B200_HOST_FORMAT") != nullptr;  // keep the decimal formatting on the host
    if (!ms && !o.use_doc && !host_format) {
        w.idcat.clear();
        w.idoff.assign(1, 0);
        for (const auto& id : b.ids) { w.idcat.insert(w.idcat.end(), id.begin(), id.end()); w.idoff.push_back(w.idcat.size()); }
        if (o.write_report) w.S.resize(nr);
        uint64_t need = b.bases.size() * 3 + w.idcat.size() + 3 * nr + 64;  // typical PML text is ~2 bytes per base
        for (;;) {
            if (need > w.text_cap) { w.text.reset(need); w.text_cap = need; }
            const int rc = spb_pml_query_host_text(idx, b.bases.data(), b.offsets.data(), nr, w.idcat.data(), w.idoff.data(), w.text.get(),
                                                   w.text_cap, &w.text_bytes, o.write_report ? w.S.data() : nullptr, &qo);
            if (rc == SPB_ERR_LIMIT && w.text_bytes > w.text_cap) { need = w.text_bytes + w.text_bytes / 2; continue; }  // rare: redo with a larger buffer
            spb_or_die(rc);
            break;
        }
        w.device_text = true;
        return;
    }
    w.narrow = !o.use_doc && b.max_len < 65536;
    void* lengths;
    if (w.narrow) { w.L16.resize(b.bases.size()); lengths = w.L16.data(); qo.length_bytes = 2; }
    else { w.L.resize(b.bases.size()); lengths = w.L.data(); }
    if (ms) w.P.resize(b.bases.size());
    if (o.use_doc) w.D.resize(b.bases.size());
    if (o.write_report) w.S.resize(nr);
    spb_bin_stats* sp = o.write_report ? w.S.data() : nullptr;
    if (o.use_doc) {
        if (ms) spb_or_die(spb_ms_query_host_doc(idx, b.bases.data(), b.offsets.data(), nr, w.L.data(), w.P.data(), w.D.data(), sp, &qo));
        else spb_or_die(spb_pml_query_host_doc(idx, b.bases.data(), b.offsets.data(), nr, w.L.data(), w.D.data(), sp, &qo));
    } else if (ms) spb_or_die(spb_ms_query_host(idx, b.bases.data(), b.offsets.data(), nr, lengths, w.P.data(), sp, &qo));
    else spb_or_die(spb_pml_query_host(idx, b.bases.data(), b.offsets.data(), nr, lengths, sp, &qo));
}

static void write_batch(bool ms, const RunOptions& o, const Work& w, Outputs& out, double& t_format, double& t_write) {
    const ReadBatch& b = w.b;
    const size_t nr = b.ids.size();
    if (!nr) return;
    static thread_local TextBuffer parts;  // one writer thread: the buffer is reused between batches
    double t0 = now_s();
    auto lap = [&](double& acc) { const double t1 = now_s(); acc += t1 - t0; t0 = t1; };
    if (o.use_doc) {  // the reference writes the document numbers first (:1003-1006)
        format_values(b, w.D.data(), (int)o.threads, parts);
        lap(t_format);
        write_all(out.docs, parts);
        lap(t_write);
    }
    if (w.device_text) out.lengths.write(w.text.get(), (std::streamsize)w.text_bytes);
    else {
        if (w.narrow) format_values(b, w.L16.data(), (int)o.threads, parts);
        else format_values(b, w.L.data(), (int)o.threads, parts);
        lap(t_format);
        write_all(out.lengths, parts);
    }
    lap(t_write);
    if (ms) {
        format_values(b, w.P.data(), (int)o.threads, parts);
        lap(t_format);
        write_all(out.pointers, parts);
        lap(t_write);
    }
    if (o.write_report) {  // rows are formatted on -t threads with the reference's manipulators, written in order
        std::vector<std::string> rows((size_t)std::max<size_t>(1, o.threads));
#pragma omp parallel num_threads((int)rows.size())
        {
            const size_t t = (size_t)omp_get_thread_num(), nt = (size_t)omp_get_num_threads();
            std::ostringstream os;
            for (size_t r = nr * t / nt; r < nr * (t + 1) / nt; r++) report_row(os, b.ids[r], w.S[r]);
            rows[t] = os.str();
        }
        lap(t_format);
        for (const auto& r : rows) out.report.write(r.data(), (std::streamsize)r.size());
    }
    lap(t_write);
}

// ---------------------------------------------------------------------------------- fast path (FASTA input)
// The file is memory-mapped and cut into windows at record boundaries; a window is parsed on -t threads straight into
// page-locked buffers (two passes over its records: sizes, then the bytes), queried on one of the GPUs the index is
// resident on by that GPU's query thread (spb_query_host_text_on: lengths, MS pointers and document numbers come back
// formatted by the device), and its texts are written at their final offsets; windows are written in input order, which
// is the reference's own order for -t 1 (App. C-11).  Anything the sequential parser treats specially (a header without an id, an empty
// record) makes the fast path hand the rest of the file, from that window on, to the sequential path below, so that
// messages and partial outputs stay those of BatchLoader.
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) spb_host_free(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 4;
        if (spb_host_alloc(&p, want) != SPB_OK) fatal_error(spb_last_error());
        cap = want;
    }
    ~PinnedBuf() { if (p) spb_host_free(p); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct FastBatch {
    uint64_t seq = 0, file_lo = 0;
    PinnedBuf bases, offsets, idcat, idoff, stats;
    PinnedBuf text[3];             // lengths | MS pointers | document numbers
    uint64_t text_bytes[3] = {0, 0, 0};
    std::vector<uint64_t> starts;  // record starts inside the window (file offsets)
    uint64_t n_reads = 0, n_bases = 0, id_bytes = 0;
    bool anomaly = false, last = false;
};

static inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }  // std::isspace, C locale

// sizes (pass 0) or bytes (pass 1) of the record at [lo, hi): header line, then sequence lines with trailing blanks stripped
static inline void parse_record(const char* d, uint64_t lo, uint64_t hi, bool copy, uint64_t& id_len, uint64_t& n_bases, uint8_t* id_out,
                                uint8_t* base_out, bool& anomaly) {
    const char* eol = (const char*)std::memchr(d + lo, '\n', hi - lo);
    const uint64_t hl_end = eol ? (uint64_t)(eol - d) : hi;  // header line = [lo, hl_end)
    const uint64_t hl = hl_end - lo;
    if (hl <= 2) anomaly = true;  // "header line is missing an id"
    uint64_t idl = hl;            // index of the first delimiter in the line, or its size
    for (uint64_t i = 1; i < hl; i++) { const char c = d[lo + i]; if (c == ' ' || c == '\t' || c == '\r') { idl = i; break; } }
    id_len = std::min(hl > 0 ? hl - 1 : 0, idl);  // substr(1, idl): keeps the delimiter (batch_loader.cpp:104-106)
    if (copy && id_len) std::memcpy(id_out, d + lo + 1, id_len);
    uint64_t nb = 0, p = eol ? hl_end + 1 : hi;
    while (p < hi) {
        const char* e2 = (const char*)std::memchr(d + p, '\n', hi - p);
        const uint64_t le = e2 ? (uint64_t)(e2 - d) : hi;
        uint64_t e = le;
        while (e > p && is_space((unsigned char)d[e - 1])) e--;
        if (copy && e > p) std::memcpy(base_out + nb, d + p, e - p);
        nb += e - p;
        p = e2 ? le + 1 : hi;
    }
    n_bases = nb;
    if (nb == 0) anomaly = true;  // "was empty after digestion"
}

// first record start at or after p ('>' at the beginning of a line), or size
static uint64_t next_record_start(const char* d, uint64_t size, uint64_t p) {
    while (p < size) {
        const char* g = (const char*)std::memchr(d + p, '>', size - p);
        if (!g) return size;
        const uint64_t at = (uint64_t)(g - d);
        if (at == 0 || d[at - 1] == '\n') return at;
        p = at + 1;
    }
    return size;
}
// record starts inside [lo, hi), found on T threads; starts[n] = hi is appended
static void window_record_starts(const char* d, uint64_t lo, uint64_t hi, int T, std::vector<uint64_t>& starts) {
    std::vector<std::vector<uint64_t>> part((size_t)T);
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; t++) {
        const uint64_t a = lo + (hi - lo) * (uint64_t)t / (uint64_t)T, e = lo + (hi - lo) * (uint64_t)(t + 1) / (uint64_t)T;
        uint64_t p = a;
        while (p < e) {
            const char* g = (const char*)std::memchr(d + p, '>', e - p);
            if (!g) break;
            const uint64_t at = (uint64_t)(g - d);
            if (at == 0 || d[at - 1] == '\n') part[(size_t)t].push_back(at);
            p = at + 1;
        }
    }
    starts.clear();
    for (auto& v : part) starts.insert(starts.end(), v.begin(), v.end());
    starts.push_back(hi);
}
// pass 0: off[r + 1] / ido[r + 1] = bases / id bytes of record r, then prefix sums; false if a record needs the sequential parser
static bool window_sizes(const char* d, const std::vector<uint64_t>& starts, int T, uint64_t* off, uint64_t* ido) {
    const uint64_t nr = starts.size() - 1;
    int bad = 0;
#pragma omp parallel for schedule(static) num_threads(T) reduction(| : bad)
    for (int64_t r = 0; r < (int64_t)nr; r++) {
        bool an = false;
        parse_record(d, starts[(size_t)r], starts[(size_t)r + 1], false, ido[r + 1], off[r + 1], nullptr, nullptr, an);
        bad |= an ? 1 : 0;
    }
    if (bad) return false;
    off[0] = 0; ido[0] = 0;
    for (uint64_t r = 0; r < nr; r++) { off[r + 1] += off[r]; ido[r + 1] += ido[r]; }
    return true;
}
// pass 1: the bytes
static void window_fill(const char* d, const std::vector<uint64_t>& starts, int T, const uint64_t* off, const uint64_t* ido, uint8_t* bases, uint8_t* ids) {
    const uint64_t nr = starts.size() - 1;
#pragma omp parallel for schedule(static) num_threads(T)
    for (int64_t r = 0; r < (int64_t)nr; r++) {
        bool an = false;
        uint64_t il, nb;
        parse_record(d, starts[(size_t)r], starts[(size_t)r + 1], true, il, nb, ids + ido[r], bases + off[r], an);
    }
}


// ---- FASTQ on the fast path --------------------------------------------------------------------------------------------
// BatchLoader takes FASTQ four lines at a time, in batches of >= 1000 estimated bases (half of the four lines' lengths,
// compute_ms_pml.cpp:903, batch_loader.cpp:33-66), and LOSES the last batch when the file ends inside its gathering loop
// (:51) -- which records survive depends on the line lengths from the top of the file.  The fast path therefore walks the
// file in segments that begin at a batch boundary (fastq_walk): it (1) finds every line end of the segment on all threads,
// (2) replays the batch formation over the line lengths (a few operations per line), which gives the lines of the complete
// batches -- the batch the segment ends in opens the next segment, the batch the FILE ends in is lost or kept as BatchLoader
// decides -- and (3) checks on all threads that every record in them is ordinary (header starts with '@' and has an id,
// sequence not empty); anything else -- an empty header line skips the rest of its batch, a bad header is fatal -- hands the
// file from that segment on to the sequential parser, which reproduces those cases line by line.
struct FastqLines {
    const char* d = nullptr;     // the segment
    uint64_t size = 0;
    std::vector<uint64_t> nl;    // positions of every '\n' in it
    uint64_t n_lines = 0;        // std::getline lines; at the end of the file a last line without terminator counts
    bool unterminated = false;
    uint64_t start(uint64_t i) const { return i ? nl[(size_t)i - 1] + 1 : 0; }
    uint64_t end(uint64_t i) const { return i < nl.size() ? nl[(size_t)i] : size; }
    uint64_t len(uint64_t i) const { return end(i) - start(i); }
};
static void fastq_index_lines(const char* d, uint64_t size, int T, bool at_eof, FastqLines& L) {
    L.d = d; L.size = size;
    std::vector<std::vector<uint64_t>> part((size_t)T);
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; t++) {
        const uint64_t a = size * (uint64_t)t / (uint64_t)T, e = size * (uint64_t)(t + 1) / (uint64_t)T;
        uint64_t p = a;
        while (p < e) {
            const char* g = (const char*)std::memchr(d + p, '\n', e - p);
            if (!g) break;
            part[(size_t)t].push_back((uint64_t)(g - d));
            p = (uint64_t)(g - d) + 1;
        }
    }
    L.nl.clear();
    for (auto& v : part) L.nl.insert(L.nl.end(), v.begin(), v.end());
    L.unterminated = at_eof && size > 0 && (L.nl.empty() || L.nl.back() + 1 < size);  // (the cut end of an inner segment is no line)
    L.n_lines = L.nl.size() + (L.unterminated ? 1 : 0);
}
// lines [0, result) of the segment are the ones BatchLoader hands to grabNextRead (FastaSource::load_fastq_batch, line for
// line); a batch the segment (not the file) ends in is left to the next segment
static uint64_t fastq_processed_lines(const FastqLines& L, bool at_eof) {
    uint64_t i = 0, processed = 0;
    bool good = true;
    while (good) {
        uint64_t covered = 0, record = 0, cnt = 0;
        bool incomplete = false;
        while (good && covered < 1000) {
            if (i >= L.n_lines) { incomplete = true; good = false; break; }  // at the end of the file: getline fails inside the loop, the batch is dropped
            if (at_eof && L.unterminated && i + 1 == L.n_lines) good = false;  // the line was ended by the end of the file
            record += L.len(i);
            i++; cnt++;
            if (cnt % 4 == 0) { covered += record / 2; record = 0; }
        }
        if (incomplete) break;
        processed = i;
    }
    return processed;
}
// records r (lines 4r .. 4r+3) for r < processed / 4, plus the header of a trailing cut-short record: all ordinary?
static bool fastq_all_ordinary(const FastqLines& L, uint64_t processed, int T) {
    const uint64_t headers = (processed + 3) / 4, records = processed / 4;
    int bad = 0;
#pragma omp parallel for schedule(static) num_threads(T) reduction(| : bad)
    for (int64_t r = 0; r < (int64_t)headers; r++) {
        const uint64_t h = 4 * (uint64_t)r, hs = L.start(h), hl = L.len(h);
        if (hl <= 2 || L.d[hs] != '@') { bad = 1; continue; }  // empty header, not a FASTQ header, or no id
        if ((uint64_t)r < records) {
            const uint64_t ss = L.start(h + 1);
            uint64_t e = L.end(h + 1);
            while (e > ss && is_space((unsigned char)L.d[e - 1])) e--;
            if (e == ss) bad = 1;  // "was empty after digestion"
        }
    }
    return !bad;
}
// a window of about `window` bytes of file starting at record r: the first record boundary at or past it (line starts ascend)
static uint64_t fastq_window_end(const FastqLines& L, uint64_t r, uint64_t records, uint64_t window) {
    const uint64_t lo = L.start(4 * r);
    uint64_t a = r + 1, z = records;
    while (a < z) {
        const uint64_t m = a + (z - a) / 2;
        if (L.start(4 * m) - lo >= window) z = m; else a = m + 1;
    }
    return a;
}
// records [r0, r1): sizes (off[i + 1] / ido[i + 1], then prefix sums), then the bytes
static void fastq_window_sizes(const FastqLines& L, uint64_t r0, uint64_t r1, int T, uint64_t* off, uint64_t* ido) {
#pragma omp parallel for schedule(static) num_threads(T)
    for (int64_t i = 0; i < (int64_t)(r1 - r0); i++) {
        const uint64_t h = 4 * (r0 + (uint64_t)i), hs = L.start(h), hl = L.len(h);
        uint64_t idl = hl;
        for (uint64_t q = 1; q < hl; q++) { const char c = L.d[hs + q]; if (c == ' ' || c == '\t' || c == '\r') { idl = q; break; } }
        ido[i + 1] = std::min(hl - 1, idl);  // substr(1, idl): keeps the delimiter (batch_loader.cpp:104-106)
        const uint64_t ss = L.start(h + 1);
        uint64_t e = L.end(h + 1);
        while (e > ss && is_space((unsigned char)L.d[e - 1])) e--;
        off[i + 1] = e - ss;
    }
    off[0] = 0; ido[0] = 0;
    for (uint64_t i = 0; i < r1 - r0; i++) { off[i + 1] += off[i]; ido[i + 1] += ido[i]; }
}
static void fastq_window_fill(const FastqLines& L, uint64_t r0, uint64_t r1, int T, const uint64_t* off, const uint64_t* ido, uint8_t* bases, uint8_t* ids) {
#pragma omp parallel for schedule(static) num_threads(T)
    for (int64_t i = 0; i < (int64_t)(r1 - r0); i++) {
        const uint64_t h = 4 * (r0 + (uint64_t)i);
        std::memcpy(ids + ido[i], L.d + L.start(h) + 1, ido[i + 1] - ido[i]);
        std::memcpy(bases + off[i], L.d + L.start(h + 1), off[i + 1] - off[i]);
    }
}

// The whole walk: emit(L, r0, r1) gets windows of ordinary records of the current segment.  Returns the file offset where the
// sequential parser has to take over -- the start of a batch -- or `size` when nothing is left for it.
template <typename Emit>
static uint64_t fastq_walk(const char* d, uint64_t size, int T, uint64_t segment, uint64_t window, Emit&& emit) {
    FastqLines L;
    uint64_t lo = 0;
    segment = std::max<uint64_t>(segment, 1);
    while (lo < size) {
        const uint64_t len = std::min(segment, size - lo);
        const bool at_eof = lo + len == size;
        fastq_index_lines(d + lo, len, T, at_eof, L);
        const uint64_t processed = fastq_processed_lines(L, at_eof);
        if (!at_eof && processed == 0) { segment *= 2; continue; }  // not even one batch fits: a longer segment
        if (!fastq_all_ordinary(L, processed, T)) return lo;
        const uint64_t records = processed / 4;
        for (uint64_t r = 0; r < records;) {
            const uint64_t r1 = fastq_window_end(L, r, records, window);
            emit(L, r, r1);
            r = r1;
        }
        if (at_eof) return size;
        lo += L.start(processed);
    }
    return size;
}

class FastRunner {
public:
    FastRunner(const RunOptions& o, bool ms, int n_gpus) : o_(o), ms_(ms), n_gpus_(std::max(1, n_gpus)), threads_((int)std::max<size_t>(1, o.threads)) {
        window_ = ms ? (16ull << 20) : (48ull << 20);  // an MS window carries ~11 bytes of pointer text per base
        if (const char* e = std::getenv("SPUMONI_B200_BATCH_MB")) { long v = std::atol(e); if (v >= 1 && v <= 2048) window_ = (uint64_t)v << 20; }
        batches_.resize((size_t)(2 * n_gpus_ + 3));
    }
    // Page-locking the batch buffers is the expensive part of setting the pipeline up: it is done while the index loads.
    void preallocate() {
        for (auto& b : batches_) {
            b.bases.reserve(window_ + 64);
            b.offsets.reserve((window_ / 64 + 2) * 8);
            b.idoff.reserve((window_ / 64 + 2) * 8);
            b.idcat.reserve(window_ / 8);
            b.text[0].reserve(window_ * 3);
            if (ms_) b.text[1].reserve(window_ * 12);
            if (o_.use_doc) b.text[2].reserve(window_ * 3);
            if (o_.write_report) b.stats.reserve((window_ / 64 + 2) * sizeof(spb_bin_stats));
        }
    }
    // Returns the number of reads processed; *resume_at = file offset where the sequential path must take over (or the file size).
    // text_paths: lengths, pointers (MS), document numbers (-d); empty = not written
    size_t run(spb_index* idx, const spb_query_opts& qo, Outputs& out, const std::string& path, const std::string text_paths[3], uint64_t* resume_at) {
        idx_ = idx; qo_ = qo; out_ = &out;
        const int fd = ::open(path.c_str(), O_RDONLY);
        struct stat st;
        if (fd < 0 || fstat(fd, &st) != 0) fatal_error("cannot open " + path);
        size_ = (uint64_t)st.st_size;
        *resume_at = size_;
        if (size_ == 0) { ::close(fd); return 0; }
        data_ = (const char*)mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
        if (data_ == MAP_FAILED) fatal_error("cannot map " + path);
        madvise((void*)data_, size_, MADV_WILLNEED);
        fastq_ = data_[0] == '@';
        for (int w = 0; w < 3; w++) {
            out_fd_[w] = -1;
            if (text_paths[w].empty()) continue;
            out_fd_[w] = ::open(text_paths[w].c_str(), O_RDWR);
            if (out_fd_[w] < 0) fatal_error("cannot open " + text_paths[w]);
        }
        for (auto& b : batches_) free_q_.push(&b);
        std::thread parser([&] { parse_loop(); });
        std::vector<std::thread> gpus;
        for (int g = 0; g < n_gpus_; g++) gpus.emplace_back([&, g] { query_loop(g); });
        write_loop();
        parser.join();
        for (auto& t : gpus) t.join();
        for (int w = 0; w < 3; w++) if (out_fd_[w] >= 0) ::close(out_fd_[w]);
        munmap((void*)data_, size_);
        ::close(fd);
        *resume_at = resume_;
        return reads_done_;
    }
    uint64_t text_written(int w) const { return text_off_[w]; }
    bool is_fastq() const { return fastq_; }
    double t_parse = 0, t_gpu = 0, t_write = 0, t_report = 0;

private:
    struct Q {  // hand-off queue (FastBatch*, nullptr = end)
        void push(FastBatch* b) { std::unique_lock<std::mutex> lk(m); q.push_back(b); cv.notify_all(); }
        FastBatch* pop() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return !q.empty(); }); FastBatch* b = q.front(); q.pop_front(); return b; }
        std::mutex m; std::condition_variable cv; std::deque<FastBatch*> q;
    };
    void parse_loop() {
        if (fastq_) return parse_loop_fastq();
        uint64_t lo = 0, seq = 0;
        bool stop = false;
        while (lo < size_ && !stop) {
            FastBatch* b = free_q_.pop();
            const double t0 = now_s();
            const uint64_t hi = lo + window_ >= size_ ? size_ : next_record_start(data_, size_, lo + window_);
            b->seq = seq++; b->file_lo = lo; b->anomaly = false; b->last = false;
            const int T = threads_;
            window_record_starts(data_, lo, hi, T, b->starts);
            const uint64_t nr = b->starts.size() - 1;
            bool ok = nr > 0 && b->starts[0] == lo;  // (a file that does not start with '>' was refused before)
            if (ok) {
                b->offsets.reserve((nr + 1) * 8);
                b->idoff.reserve((nr + 1) * 8);
                if (o_.write_report) b->stats.reserve((nr + 1) * sizeof(spb_bin_stats));
                ok = window_sizes(data_, b->starts, T, b->offsets.as<uint64_t>(), b->idoff.as<uint64_t>());
            }
            if (!ok) {  // hand this window and everything after it to the sequential path
                b->anomaly = true;
                b->n_reads = 0;
                resume_ = lo;
                stop = true;
            } else {
                const uint64_t* off = b->offsets.as<uint64_t>();
                const uint64_t* ido = b->idoff.as<uint64_t>();
                b->n_reads = nr; b->n_bases = off[nr]; b->id_bytes = ido[nr];
                b->bases.reserve(b->n_bases + 64);
                b->idcat.reserve(b->id_bytes + 64);
                window_fill(data_, b->starts, T, off, ido, b->bases.as<uint8_t>(), b->idcat.as<uint8_t>());
                lo = hi;
                if (lo >= size_) resume_ = size_;
            }
            t_parse += now_s() - t0;
            parsed_q_.push(b);
        }
        for (int g = 0; g < n_gpus_; g++) parsed_q_.push(nullptr);
    }
    void parse_loop_fastq() {
        const int T = threads_;
        uint64_t seq = 0;
        uint64_t segment = 512ull << 20;  // line index: 8 bytes per line of the segment
        if (const char* e = std::getenv("SPUMONI_B200_FASTQ_SEGMENT_MB")) { long v = std::atol(e); if (v >= 1 && v <= 65536) segment = (uint64_t)v << 20; }
        double t0 = now_s();
        resume_ = fastq_walk(data_, size_, T, segment, window_, [&](const FastqLines& L, uint64_t r, uint64_t r1) {
            t_parse += now_s() - t0;  // (indexing, replay and checks of the segment since the last window)
            FastBatch* b = free_q_.pop();
            t0 = now_s();
            const uint64_t nr = r1 - r;
            b->seq = seq++; b->file_lo = (uint64_t)(L.d - data_) + L.start(4 * r); b->anomaly = false; b->last = false;
            b->offsets.reserve((nr + 1) * 8);
            b->idoff.reserve((nr + 1) * 8);
            if (o_.write_report) b->stats.reserve((nr + 1) * sizeof(spb_bin_stats));
            uint64_t* off = b->offsets.as<uint64_t>();
            uint64_t* ido = b->idoff.as<uint64_t>();
            fastq_window_sizes(L, r, r1, T, off, ido);
            b->n_reads = nr; b->n_bases = off[nr]; b->id_bytes = ido[nr];
            b->bases.reserve(b->n_bases + 64);
            b->idcat.reserve(b->id_bytes + 64);
            fastq_window_fill(L, r, r1, T, off, ido, b->bases.as<uint8_t>(), b->idcat.as<uint8_t>());
            t_parse += now_s() - t0;
            parsed_q_.push(b);
            t0 = now_s();
        });
        t_parse += now_s() - t0;
        for (int g = 0; g < n_gpus_; g++) parsed_q_.push(nullptr);
    }
    void query_loop(int g) {
        for (;;) {
            FastBatch* b = parsed_q_.pop();
            if (!b) break;
            const double t0 = now_s();
            if (b->n_reads) {
                // typical text: ~2 bytes per base of lengths, ~11 of pointers; a buffer that turns out too small is grown and the window redone
                uint64_t need[3] = {b->n_bases * 3 + b->id_bytes + 3 * b->n_reads + 64, b->n_bases * 12 + b->id_bytes + 3 * b->n_reads + 64,
                                    b->n_bases * 3 + b->id_bytes + 3 * b->n_reads + 64};
                for (;;) {
                    spb_text_buf tb[3];
                    for (int w = 0; w < 3; w++) {
                        if (out_fd_[w] >= 0) b->text[w].reserve(need[w]);
                        tb[w] = spb_text_buf{b->text[w].as<char>(), b->text[w].cap, 0};
                    }
                    const int rc = spb_query_host_text_on(idx_, g, b->bases.as<uint8_t>(), b->offsets.as<uint64_t>(), b->n_reads, b->idcat.as<uint8_t>(),
                                                          b->idoff.as<uint64_t>(), &tb[0], ms_ ? &tb[1] : nullptr, out_fd_[2] >= 0 ? &tb[2] : nullptr,
                                                          o_.write_report ? b->stats.as<spb_bin_stats>() : nullptr, &qo_);
                    bool grow = false;
                    for (int w = 0; w < 3; w++) {
                        b->text_bytes[w] = tb[w].bytes;
                        if (rc == SPB_ERR_LIMIT && out_fd_[w] >= 0 && tb[w].bytes > b->text[w].cap) { need[w] = tb[w].bytes + tb[w].bytes / 2; grow = true; }
                    }
                    if (grow) continue;
                    spb_or_die(rc);
                    break;
                }
            } else b->text_bytes[0] = b->text_bytes[1] = b->text_bytes[2] = 0;
            {
                std::unique_lock<std::mutex> lk(gpu_m_);
                t_gpu += now_s() - t0;
            }
            done_q_.push(b);
        }
        done_q_.push(nullptr);
    }
    void write_loop() {
        std::map<uint64_t, FastBatch*> pending;  // windows finish out of order with several GPUs
        uint64_t next_seq = 0;
        int ended = 0;
        while (ended < n_gpus_) {
            FastBatch* b = done_q_.pop();
            if (!b) { ended++; continue; }
            pending[b->seq] = b;
            while (!pending.empty() && pending.begin()->first == next_seq) {
                FastBatch* w = pending.begin()->second;
                pending.erase(pending.begin());
                next_seq++;
                write_batch(*w);
                free_q_.push(w);
            }
        }
    }
    void write_batch(FastBatch& b) {
        if (!b.n_reads) return;
        double t0 = now_s();
        int failed = 0;
        // SPUMONI_B200_WRITE=write (default): one call per window | pwrite: 8 MiB pieces on -t threads | mmap: extend the file,
        // map the window's range and copy on -t threads.  Buffered writes to ONE file serialise on its inode lock whatever
        // the thread count: tools/fs_write_bench measures 5.2 GB/s for one stream, 4.6 for parallel pwrite and 4.7 through a
        // shared mapping on the pool's boxes (profiles/r2d_fs_write_bench.txt) -- and with 2.1 bytes of text per base that
        // write rate, not the GPU, is what bounds the phase.  The (up to three) files of a window are written concurrently.
        static const char* wm = std::getenv("SPUMONI_B200_WRITE");
        const int mode = !wm ? 1 : (std::strcmp(wm, "pwrite") == 0 ? 0 : (std::strcmp(wm, "mmap") == 0 ? 2 : 1));
#pragma omp parallel for schedule(static, 1) num_threads(3) reduction(| : failed) if (ms_ || o_.use_doc)
        for (int w = 0; w < 3; w++) {
            if (out_fd_[w] < 0 || !b.text_bytes[w]) continue;
            const char* text = b.text[w].as<char>();
            const uint64_t base_off = text_off_[w], nbytes = b.text_bytes[w];
            const int fd = out_fd_[w];
            if (mode == 2) {
                const uint64_t pg = (uint64_t)sysconf(_SC_PAGESIZE), lo = base_off / pg * pg, hi = base_off + nbytes;
                if (ftruncate(fd, (off_t)hi) != 0) { failed = 1; continue; }
                char* m = (char*)mmap(nullptr, hi - lo, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)lo);
                if (!m || m == MAP_FAILED) { failed = 1; continue; }
                const uint64_t piece = 2ull << 20, np = (nbytes + piece - 1) / piece;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads_)
                for (int64_t i = 0; i < (int64_t)np; i++) {
                    const uint64_t a = (uint64_t)i * piece, e = std::min(nbytes, a + piece);
                    std::memcpy(m + (base_off - lo) + a, text + a, e - a);
                }
                munmap(m, hi - lo);
            } else {
                const uint64_t piece = mode == 1 ? nbytes : (8ull << 20), np = (nbytes + piece - 1) / piece;
                int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(mode == 1 ? 1 : threads_) reduction(| : bad)
                for (int64_t i = 0; i < (int64_t)np; i++) {
                    uint64_t a = (uint64_t)i * piece;
                    const uint64_t e = std::min(nbytes, a + piece);
                    while (a < e) {
                        const ssize_t wr = pwrite(fd, text + a, e - a, (off_t)(base_off + a));
                        if (wr <= 0) { bad = 1; break; }
                        a += (uint64_t)wr;
                    }
                }
                failed |= bad;
            }
        }
        if (failed) fatal_error("write to the output file failed");
        for (int w = 0; w < 3; w++) text_off_[w] += b.text_bytes[w];
        reads_done_ += b.n_reads;
        t_write += now_s() - t0;
        if (o_.write_report) {  // rows are formatted on -t threads with the reference's manipulators, written in order
            t0 = now_s();
            const uint8_t* ids = b.idcat.as<uint8_t>();
            const uint64_t* ido = b.idoff.as<uint64_t>();
            const spb_bin_stats* S = b.stats.as<spb_bin_stats>();
            std::vector<std::string> rows((size_t)threads_);
            const size_t nr = b.n_reads;
#pragma omp parallel num_threads(threads_)
            {
                const size_t t = (size_t)omp_get_thread_num(), nt = (size_t)omp_get_num_threads();
                std::ostringstream os;
                std::string id;
                for (size_t r = nr * t / nt; r < nr * (t + 1) / nt; r++) {
                    id.assign((const char*)ids + ido[r], ido[r + 1] - ido[r]);
                    report_row(os, id, S[r]);
                }
                if (t < rows.size()) rows[t] = os.str();
            }
            for (const auto& r : rows) out_->report.write(r.data(), (std::streamsize)r.size());
            t_report += now_s() - t0;
        }
    }

    spb_index* idx_ = nullptr;
    const RunOptions& o_;
    bool ms_ = false;
    spb_query_opts qo_;
    Outputs* out_ = nullptr;
    int n_gpus_, threads_;
    uint64_t window_ = 0, size_ = 0, resume_ = 0, text_off_[3] = {0, 0, 0};
    bool fastq_ = false;
    size_t reads_done_ = 0;
    const char* data_ = nullptr;
    int out_fd_[3] = {-1, -1, -1};
    std::vector<FastBatch> batches_;
    Q free_q_, parsed_q_, done_q_;
    std::mutex gpu_m_;
};

static size_t classify_reads(spb_index* idx, bool ms, const RunOptions& o, FastRunner* fast) {
    Outputs out;
    out.lengths.open(o.pattern_file + (ms ? ".lengths" : ".pseudo_lengths"));
    if (ms) out.pointers.open(o.pattern_file + ".pointers");
    if (o.use_doc) out.docs.open(o.pattern_file + ".doc_numbers");
    spb_query_opts qo;
    std::memset(&qo, 0, sizeof(qo));
    qo.flags = o.is_general_text ? 0 : SPB_QUERY_TOUPPER;  // general text is used verbatim (:1219-1297)
    if (!o.is_general_text) {
        // the reference loads the null database unconditionally (:864-869) but only uses it with -c
        spb_nulldb db;
        std::memset(&db, 0, sizeof(db));
        const std::string dbp = o.ref_file + (ms ? ".msnulldb" : ".pmlnulldb");
        if (is_file(dbp)) spb_or_die(spb_nulldb_load(dbp.c_str(), &db));
        else if (o.write_report) fatal_error("The following path is not valid: " + dbp);
        qo.max_value_thr = spb_max_value_threshold(db.percentile_value, ms ? 0 : 1, o.use_dna_letters, o.use_promotions);
        if (o.write_report) {
            out.report.open(o.pattern_file + ".report", std::ofstream::out);
            qo.bin_width = (uint32_t)o.bin_size;
            report_header(out.report, qo.max_value_thr);
        }
    }
    const uint64_t kBatchBases = 64ull << 20;
    constexpr int kWorks = 4;
    const bool timing = std::getenv("SPUMONI_B200_TIMING") != nullptr;  // per-stage busy time on stderr
    double t_parse = 0, t_gpu = 0, t_format = 0, t_write = 0;
    size_t num_reads = 0;
    uint64_t resume_at = 0;
    if (fast) {  // FASTA input: memory-mapped, parsed on all threads, one query thread per GPU, texts formatted on the device
        int first = EOF;
        { std::ifstream probe(o.pattern_file, std::ifstream::binary); first = probe.peek(); }
        if (first == '>' || first == '@') {
            out.lengths.flush();
            const std::string paths[3] = {o.pattern_file + (ms ? ".lengths" : ".pseudo_lengths"), ms ? o.pattern_file + ".pointers" : std::string(),
                                          o.use_doc ? o.pattern_file + ".doc_numbers" : std::string()};
            num_reads = fast->run(idx, qo, out, o.pattern_file, paths, &resume_at);
            out.lengths.seekp((std::streamoff)fast->text_written(0));
            if (ms) out.pointers.seekp((std::streamoff)fast->text_written(1));
            if (o.use_doc) out.docs.seekp((std::streamoff)fast->text_written(2));
            if (timing) std::fprintf(stderr, "\n[timing] busy seconds per stage: parse %.3f | query (H2D + kernels + D2H) %.3f | format 0.000 | write %.3f | report %.3f\n",
                                     fast->t_parse, fast->t_gpu, fast->t_write, fast->t_report);
            struct stat st;
            if (stat(o.pattern_file.c_str(), &st) == 0 && resume_at >= (uint64_t)st.st_size) return num_reads;
        }
    }
    std::vector<Work> works(kWorks);
    WorkQueue free_q, parsed_q, done_q;
    for (auto& w : works) free_q.push(&w);

    std::thread parser([&] {
        Work* w = free_q.pop();
        w->b.clear();
        double t0 = now_s();
        auto flush = [&](bool last) {
            t_parse += now_s() - t0;
            parsed_q.push(w);
            if (last) { parsed_q.push(nullptr); return; }
            w = free_q.pop();
            w->b.clear();
            t0 = now_s();
        };
        if (!o.is_general_text) {
            FastaSource src(o.pattern_file, resume_at, fast && resume_at > 0 && fast->is_fastq());
            while (src.next(w->b)) {
                num_reads++;
                if (w->b.bases.size() >= kBatchBases) flush(false);
            }
        } else {  // classify_general_reads_*: reads are separated by 0x01, ids are read_<i>
            std::ifstream fd(o.pattern_file, std::ifstream::in | std::ifstream::binary);
            char ch = (char)fd.get();
            uint64_t cur = 0;
            while (fd.good()) {
                if (ch == '\x01') {
                    w->b.ids.push_back("read_" + std::to_string(num_reads));
                    w->b.offsets.push_back(w->b.bases.size());
                    if (cur > w->b.max_len) w->b.max_len = cur;
                    cur = 0;
                    num_reads++;
                    if (w->b.bases.size() >= kBatchBases) flush(false);
                } else {
                    w->b.bases.push_back((uint8_t)ch);
                    cur++;
                }
                ch = (char)fd.get();
            }
            w->b.bases.resize(w->b.offsets.back());  // bytes after the last separator are never queried
        }
        flush(true);
    });
    std::thread writer([&] {
        for (;;) {
            Work* w = done_q.pop();
            if (!w) break;
            write_batch(ms, o, *w, out, t_format, t_write);
            free_q.push(w);
        }
    });
    for (;;) {
        Work* w = parsed_q.pop();
        if (!w) break;
        const double t0 = now_s();
        query_batch(idx, ms, o, qo, *w);
        t_gpu += now_s() - t0;
        done_q.push(w);
    }
    done_q.push(nullptr);
    parser.join();
    writer.join();
    if (timing) std::fprintf(stderr, "\n[timing] busy seconds per stage: parse %.3f | query (H2D + kernels + D2H) %.3f | format %.3f | write %.3f\n", t_parse, t_gpu, t_format, t_write);
    return num_reads;
}

static int run_main(int argc, char** argv) {
    if (argc == 1) return run_usage();
    RunOptions o;
    parse_run_options(argc, argv, &o);
    const bool ms = o.ms_requested && !o.pml_requested;
    validate(o, ms);
    o.ref_file += o.use_promotions ? ".bin" : ".fa";
    if (o.is_general_text) o.use_doc = false;  // classify_general_reads_* never look at the document array (compute_ms_pml.cpp:1219-1297)
    const char* tag = ms ? "compute_ms" : "compute_pml";

    // GPUs: SPUMONI_B200_DEVICES="0,2,3" (or the older SPUMONI_B200_DEVICE=<one ordinal>), else as many of the visible GPUs as
    // there are threads (-t, the reference's number of parallel workers, compute_ms_pml.cpp:890-894)
    std::vector<int> devices;
    if (const char* e = std::getenv("SPUMONI_B200_DEVICES")) {
        std::stringstream ss(e);
        std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) devices.push_back(std::atoi(tok.c_str()));
    } else if (const char* e1 = std::getenv("SPUMONI_B200_DEVICE")) devices.push_back(std::atoi(e1));
    else {
        const int visible = spb_device_count();
        for (int d = 0; d < std::max(1, std::min<int>(visible, (int)o.threads)); d++) devices.push_back(d);
    }
    if (devices.empty()) devices.push_back(0);
    const bool fast_ok = !o.is_general_text && !std::getenv("SPUMONI_B200_HOST_FORMAT") && !std::getenv("SPUMONI_B200_SEQUENTIAL");
    std::unique_ptr<FastRunner> fast;
    std::thread prealloc;
    if (fast_ok && spb_device_count() > 0) {  // page-lock the pipeline's buffers while the index loads
        fast.reset(new FastRunner(o, ms, (int)devices.size()));
        prealloc = std::thread([&] { fast->preallocate(); });
    }

    status_log(ms ? "ms_construct" : "pml_construct", ms ? "loading the MS index" : "loading the PML index");
    double t0 = now_s();
    spb_index* idx = nullptr;
    {
        const int rc = spb_index_open_ex(o.ref_file.c_str(), ms ? SPB_MODE_MS : SPB_MODE_PML, o.use_doc ? SPB_OPEN_DOC : 0, &idx);
        if (rc) { if (prealloc.joinable()) prealloc.join(); spb_or_die(rc); }
    }
    if (o.use_doc) { done_log(now_s() - t0); status_log(ms ? "ms_construct" : "pml_construct", "loading the document array"); }
    {
        const int rc = spb_index_replicate(idx, devices.data(), (int)devices.size());
        if (prealloc.joinable()) prealloc.join();
        spb_or_die(rc);
    }
    if (std::getenv("SPUMONI_B200_TIMING")) std::fprintf(stderr, "\n[timing] devices: %d %s\n", (int)devices.size(), spb_replicate_method());
    done_log(now_s() - t0);
    std::cout << std::endl;
    force_log(tag, "input reads will be used directly, no minimizer digestion");

    t0 = now_s();
    status_log(tag, ms ? "processing the reads" : "processing the patterns");
    size_t num_reads = classify_reads(idx, ms, o, fast.get());
    done_log(now_s() - t0);
    force_log(tag, "finished processing " + std::to_string(num_reads) + " reads. results are saved in *." + (ms ? "lengths" : "pseudo_lengths") + " file.");
    std::cout << std::endl;
    spb_index_close(idx);
    return 0;
}

int main(int argc, char** argv) {
    std::fprintf(stderr, "\n\033[1m\033[31mSPUMONI version: %s \033[0m(spumoni_b200 query engine)\n\n", SPUMONI_VERSION);
    if (argc > 1) {
        if (std::strcmp(argv[1], "run") == 0) return run_main(argc - 1, argv + 1);
        if (std::strcmp(argv[1], "build") == 0) {
            std::fprintf(stderr, "spumoni_b200 provides the `run` sub-command only; build indexes with the reference's `spumoni build`.\n");
            return 1;
        }
    }
    std::fprintf(stderr, "SPUMONI has different sub-commands to run which can used as follows:\nUsage: spumoni <command> [options]\n\n");
    std::fprintf(stderr, "Commands:\n\trun\tcomputes MSs or PMLs for patterns against already built SPUMONI index.\n\n");
    return 1;
}
