// wbench.cpp -- how fast can one process put N MiB into ONE file?  (sizing the file driver's writer)
//   wbench MODE MiB THREADS PATH     MODE 0: pwrite   1: memcpy into a MAP_SHARED mapping   2: fallocate + pwrite
//                                         3: fallocate + O_DIRECT pwrite
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
using namespace std;
static double now() { return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    if (argc < 5) return 1;
    const int mode = atoi(argv[1]);
    const size_t N = (size_t)atol(argv[2]) << 20;
    const int nt = atoi(argv[3]);
    const char *path = argv[4];
    char *src = nullptr;
    if (posix_memalign((void **)&src, 4096, N)) return 2;
    memset(src, 1, N);
    int fd = open(path, O_RDWR | O_CREAT | O_TRUNC | (mode == 3 ? O_DIRECT : 0), 0644);
    if (fd < 0) { perror("open"); return 3; }
    if (ftruncate(fd, N)) return 4;
    double t = now();
    if (mode >= 2 && posix_fallocate(fd, 0, N)) perror("fallocate");
    double tf = now() - t;
    vector<thread> th;
    const size_t c = N / nt;
    if (mode != 1) {
        for (int i = 0; i < nt; i++)
            th.emplace_back([&, i] {
                size_t off = i * c, len = c;
                while (len) {
                    ssize_t w = pwrite(fd, src + off, min(len, (size_t)64 << 20), off);
                    if (w <= 0) { perror("pwrite"); return; }
                    off += w;
                    len -= w;
                }
            });
        for (auto &x : th) x.join();
    } else {
        char *m = (char *)mmap(0, N, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        for (int i = 0; i < nt; i++) th.emplace_back([&, i] { memcpy(m + i * c, src + i * c, c); });
        for (auto &x : th) x.join();
        munmap(m, N);
    }
    double dt = now() - t;
    printf("mode %d threads %d %s: %.2f GB/s (fallocate %.3f s)\n", mode, nt, path, N / 1e9 / dt, tf);
    close(fd);
    unlink(path);
}
