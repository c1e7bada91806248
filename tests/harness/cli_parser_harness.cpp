// Test harness (CPU): runs the product CLI's own input parser (FastaSource / LineReader of spumoni_run.cpp, included as is)
// over a file and prints every record as ">id\n<sequence>\n", so that the parser can be compared with the reference's
// BatchLoader without a GPU. Exit code 1 on the inputs the CLI rejects.
#define main spumoni_cli_main
#include "../../spumoni_b200/csrc/spumoni_run.cpp"
#undef main

// argv[2] = window size in bytes, argv[3] = threads, argv[4] = FASTQ segment size in bytes (default: the whole file): the fast path's window parser (window_record_starts / window_sizes /
// window_fill, the functions FastRunner::parse_loop calls) over the windows of the file, handing over to the sequential
// parser at the first window that needs it -- exactly what `spumoni run` does, minus the GPU.
static void fast_parse(const char* path, uint64_t window, int T, uint64_t segment, ReadBatch& b) {
    std::ifstream f(path, std::ifstream::binary);
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const uint64_t size = data.size();
    uint64_t lo = 0, resume = size;
    if (size && data[0] == '@') {  // FASTQ: segments, line index, replay of the batch formation, validation, windows of records
        const uint64_t got = fastq_walk(data.data(), size, T, segment, window, [&](const FastqLines& L, uint64_t r, uint64_t r1) {
            std::vector<uint64_t> off(r1 - r + 1), ido(r1 - r + 1);
            fastq_window_sizes(L, r, r1, T, off.data(), ido.data());
            std::vector<uint8_t> bases(off[r1 - r] + 1), ids(ido[r1 - r] + 1);
            fastq_window_fill(L, r, r1, T, off.data(), ido.data(), bases.data(), ids.data());
            for (uint64_t i = 0; i < r1 - r; i++) {
                b.ids.emplace_back((const char*)ids.data() + ido[i], ido[i + 1] - ido[i]);
                b.bases.insert(b.bases.end(), bases.begin() + off[i], bases.begin() + off[i + 1]);
                b.offsets.push_back(b.bases.size());
            }
        });
        std::fprintf(stderr, got >= size ? "fastq:fast\n" : "fastq:sequential from %llu\n", (unsigned long long)got);
        if (got < size) {
            FastaSource src(path, got, got > 0);
            while (src.next(b)) {}
        }
        return;
    }
    if (size && data[0] != '>') { resume = 0; lo = size; }
    while (lo < size) {
        const uint64_t hi = lo + window >= size ? size : next_record_start(data.data(), size, lo + window);
        std::vector<uint64_t> starts;
        window_record_starts(data.data(), lo, hi, T, starts);
        const uint64_t nr = starts.size() - 1;
        std::vector<uint64_t> off(nr + 1), ido(nr + 1);
        if (!(nr > 0 && starts[0] == lo && window_sizes(data.data(), starts, T, off.data(), ido.data()))) { resume = lo; break; }
        std::vector<uint8_t> bases(off[nr] + 1), ids(ido[nr] + 1);
        window_fill(data.data(), starts, T, off.data(), ido.data(), bases.data(), ids.data());
        for (uint64_t r = 0; r < nr; r++) {
            b.ids.emplace_back((const char*)ids.data() + ido[r], ido[r + 1] - ido[r]);
            b.bases.insert(b.bases.end(), bases.begin() + off[r], bases.begin() + off[r + 1]);
            b.offsets.push_back(b.bases.size());
        }
        lo = hi;
    }
    if (resume < size) {
        FastaSource src(path, resume);
        while (src.next(b)) {}
    }
}

int main(int argc, char** argv) {
    if (argc != 2 && argc != 4 && argc != 5) return 2;
    ReadBatch b;
    if (argc >= 4) fast_parse(argv[1], (uint64_t)std::atoll(argv[2]), std::atoi(argv[3]), argc == 5 ? (uint64_t)std::atoll(argv[4]) : ~0ull >> 1, b);
    else {
        FastaSource src(argv[1]);
        while (src.next(b)) {}
    }
    for (size_t r = 0; r < b.ids.size(); r++) {
        std::fputc('>', stdout);
        std::fwrite(b.ids[r].data(), 1, b.ids[r].size(), stdout);
        std::fputc('\n', stdout);
        std::fwrite(b.bases.data() + b.offsets[r], 1, (size_t)(b.offsets[r + 1] - b.offsets[r]), stdout);
        std::fputc('\n', stdout);
    }
    return 0;
}
