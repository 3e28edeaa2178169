// hostbw.cpp -- host memory bandwidth ceilings for the host-pointer (e2e) path: what T threads can write with
// non-temporal stores, write with plain stores (read-for-ownership), read, and copy.
//   g++ -O3 -march=native -pthread -o hostbw hostbw.cpp ; ./hostbw [MiB per thread] [threads ...]
#include <immintrin.h>
#include <sched.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

enum Mode { NT_WRITE, WRITE, READ, COPY_NT };
static const char* NAMES[] = {"nt_write", "write", "read", "copy_nt"};

static double run(Mode m, int T, size_t bytes_per_thread, const std::vector<int>& cpus)
{
    std::vector<std::thread> th;
    std::vector<double*> bufs(T), srcs(T);
    const size_t n = bytes_per_thread / 8;
    for (int t = 0; t < T; ++t) {
        bufs[t] = (double*)aligned_alloc(64, n * 8);
        srcs[t] = (double*)aligned_alloc(64, n * 8);
    }
    std::atomic<int> ready{0}, go{0};
    std::vector<double> t0(T), t1(T);
    std::atomic<long long> sink{0};
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            if (!cpus.empty()) {
                cpu_set_t s; CPU_ZERO(&s); CPU_SET(cpus[t % cpus.size()], &s);
                sched_setaffinity(0, sizeof s, &s);
            }
            double* b = bufs[t]; double* s = srcs[t];
            memset(b, 1, n * 8); memset(s, 2, n * 8);      // first touch on this thread's node
            ready++;
            while (!go.load()) {}
            t0[t] = now();
            for (int rep = 0; rep < 4; ++rep) {
                if (m == NT_WRITE) {
                    const __m128d v = _mm_set1_pd((double)rep);
                    for (size_t i = 0; i < n; i += 8) {
                        _mm_stream_pd(b + i, v); _mm_stream_pd(b + i + 2, v);
                        _mm_stream_pd(b + i + 4, v); _mm_stream_pd(b + i + 6, v);
                    }
                    _mm_sfence();
                } else if (m == WRITE) {
                    memset(b, rep, n * 8);
                } else if (m == READ) {
                    __m128d a = _mm_setzero_pd();
                    for (size_t i = 0; i < n; i += 2) a = _mm_add_pd(a, _mm_load_pd(b + i));
                    sink += (long long)_mm_cvtsd_f64(a);
                } else {
                    for (size_t i = 0; i < n; i += 2) _mm_stream_pd(b + i, _mm_load_pd(s + i));
                    _mm_sfence();
                }
            }
            t1[t] = now();
        });
    while (ready.load() < T) {}
    go = 1;
    for (auto& x : th) x.join();
    double a = t0[0], z = t1[0];
    for (int t = 0; t < T; ++t) { a = std::min(a, t0[t]); z = std::max(z, t1[t]); }
    for (int t = 0; t < T; ++t) { free(bufs[t]); free(srcs[t]); }
    return 4.0 * T * n * 8 / (z - a) / 1e9;
}

int main(int argc, char** argv)
{
    size_t mib = argc > 1 ? atoi(argv[1]) : 256;
    cpu_set_t set; sched_getaffinity(0, sizeof set, &set);
    std::vector<int> cpus;
    for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &set)) cpus.push_back(c);
    printf("affinity: %zu cpus:", cpus.size());
    for (int c : cpus) printf(" %d", c);
    printf("\n");
    std::vector<int> Ts;
    for (int i = 2; i < argc; ++i) Ts.push_back(atoi(argv[i]));
    if (Ts.empty()) { for (int t = 1; t < (int)cpus.size(); t *= 2) Ts.push_back(t); Ts.push_back((int)cpus.size()); }
    for (int T : Ts)
        for (int m = 0; m < 4; ++m)
            printf("threads %3d %-9s %7.1f GB/s%s\n", T, NAMES[m], run((Mode)m, T, mib << 20, cpus),
                   m == COPY_NT ? " (bytes written; as many again are read)" : "");
    return 0;
}
