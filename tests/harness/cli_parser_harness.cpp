// Test harness (CPU): runs the product CLI's own input parser (FastaSource / LineReader of spumoni_run.cpp, included as is)
// over a file and prints every record as ">id\n<sequence>\n", so that the parser can be compared with the reference's
// BatchLoader without a GPU. Exit code 1 on the inputs the CLI rejects.
#define main spumoni_cli_main
#include "../../spumoni_b200/csrc/spumoni_run.cpp"
#undef main

int main(int argc, char** argv) {
    if (argc != 2) return 2;
    FastaSource src(argv[1]);
    ReadBatch b;
    while (src.next(b)) {}
    for (size_t r = 0; r < b.ids.size(); r++) {
        std::fputc('>', stdout);
        std::fwrite(b.ids[r].data(), 1, b.ids[r].size(), stdout);
        std::fputc('\n', stdout);
        std::fwrite(b.bases.data() + b.offsets[r], 1, (size_t)(b.offsets[r + 1] - b.offsets[r]), stdout);
        std::fputc('\n', stdout);
    }
    return 0;
}
