// Development tool: how fast can this box write one large file?  (the floor of the `spumoni run` phase: its output is 2-3 bytes per base)
// usage: fs_write_bench <path> <MiB> <threads>   -- one write() stream | parallel pwrite | ftruncate / fallocate + shared mapping + parallel memcpy
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <omp.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    const char* path = argv[1];
    const size_t bytes = (size_t)atoll(argv[2]) << 20;
    const int T = atoi(argv[3]);
    std::vector<char> buf(bytes);
#pragma omp parallel for
    for (long i = 0; i < (long)bytes; i += 4096) memset(&buf[i], 'a' + (i >> 12) % 26, std::min<size_t>(4096, bytes - i));
    for (int mode = 0; mode < 4; mode++) {
        unlink(path);
        int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
        double t0 = now();
        if (mode == 0) { size_t a = 0; while (a < bytes) { ssize_t w = write(fd, buf.data() + a, std::min<size_t>(bytes - a, 64u << 20)); a += w; } }
        else if (mode == 1) {
            const size_t piece = 8u << 20; const long np = (bytes + piece - 1) / piece;
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
            for (long i = 0; i < np; i++) { size_t a = i * piece, e = std::min(bytes, a + piece); while (a < e) { ssize_t w = pwrite(fd, buf.data() + a, e - a, a); a += w; } }
        } else if (mode == 2 || mode == 3) {
            if (mode == 3) { if (posix_fallocate(fd, 0, bytes) != 0) perror("fallocate"); } else if (ftruncate(fd, bytes) != 0) perror("ftruncate");
            char* m = (char*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            if (m == MAP_FAILED) { perror("mmap"); return 1; }
            const size_t piece = 2u << 20; const long np = (bytes + piece - 1) / piece;
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
            for (long i = 0; i < np; i++) { size_t a = i * piece, e = std::min(bytes, a + piece); memcpy(m + a, buf.data() + a, e - a); }
            munmap(m, bytes);
        }
        double t1 = now();
        close(fd);
        printf("mode %d (%s): %.3f s  %.2f GB/s\n", mode, mode == 0 ? "write" : mode == 1 ? "parallel pwrite" : mode == 2 ? "ftruncate+mmap+memcpy" : "fallocate+mmap+memcpy", t1 - t0, bytes / 1e9 / (t1 - t0));
    }
    unlink(path);
}
