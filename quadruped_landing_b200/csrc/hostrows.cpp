// hostrows.cpp -- see hostrows.h.  Built with g++ (not nvcc) so that the AVX-512 row writer can live behind a
// function-level target attribute and be selected at run time.
#include "hostrows.h"

#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#if defined(__x86_64__)
#include <immintrin.h>
#define QLH_X86 1
#else
#define QLH_X86 0
#endif

namespace qlhost {

static inline void cpu_relax()
{
#if QLH_X86
    _mm_pause();
#endif
}

// ================================================================================================ Pool
Pool::Pool(int nthreads, const std::vector<int>& cpus, bool pin)
{
    const int nw = std::max(0, nthreads - 1);
    workers_.reserve(nw);
    for (int i = 0; i < nw; ++i) {
        // the caller's thread is the pool's first member and is left unpinned: workers take cpus[1..]
        const int cpu = (pin && !cpus.empty()) ? cpus[(i + 1) % cpus.size()] : -1;
        workers_.emplace_back(&Pool::worker_main, this, i, cpu);
    }
}

Pool::~Pool()
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
        epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
}

void Pool::run_blocks()
{
    const int64_t n = n_, blk = block_;
    for (;;) {
        const int64_t i = next_.fetch_add(blk, std::memory_order_relaxed);
        if (i >= n) break;
        (*fn_)(i, std::min(n, i + blk));
    }
}

void Pool::worker_main(int, int cpu)
{
    if (cpu >= 0) {
        cpu_set_t s;
        CPU_ZERO(&s);
        CPU_SET(cpu, &s);
        sched_setaffinity(0, sizeof s, &s);       // best effort
    }
    uint64_t seen = 0;
    for (;;) {
        // spin for a while (the next pipeline chunk usually arrives within microseconds), then sleep
        int spins = 0;
        while (epoch_.load(std::memory_order_acquire) == seen) {
            if (++spins < 20000) { cpu_relax(); continue; }
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
        }
        seen = epoch_.load(std::memory_order_acquire);
        if (stop_) return;
        run_blocks();
        pending_.fetch_sub(1, std::memory_order_acq_rel);
    }
}

void Pool::parallel_for(int64_t n, int64_t block, const std::function<void(int64_t, int64_t)>& fn)
{
    if (n <= 0) return;
    if (workers_.empty() || n <= block) { fn(0, n); return; }
    {
        std::lock_guard<std::mutex> lk(mu_);
        fn_ = &fn;
        n_ = n;
        block_ = std::max<int64_t>(1, block);
        next_.store(0, std::memory_order_relaxed);
        pending_.store((int)workers_.size(), std::memory_order_relaxed);
        epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    run_blocks();
    while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
}

void Pool::start(int64_t n, int64_t block, std::function<void(int64_t, int64_t)> fn)
{
    if (n <= 0) return;
    if (workers_.empty()) { fn(0, n); return; }
    finish();
    {
        std::lock_guard<std::mutex> lk(mu_);
        owned_ = std::move(fn);
        fn_ = &owned_;
        n_ = n;
        block_ = std::max<int64_t>(1, block);
        next_.store(0, std::memory_order_relaxed);
        pending_.store((int)workers_.size(), std::memory_order_relaxed);
        epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
}

void Pool::finish()
{
    while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
}

// ================================================================================================ Worker
Worker::Worker()
{
    th_ = std::thread([this] {
        for (;;) {
            std::function<void()> fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                fn = std::move(q_.front());
                q_.erase(q_.begin());
                busy_ = true;
            }
            fn();
            {
                std::lock_guard<std::mutex> lk(mu_);
                busy_ = false;
            }
            cv_.notify_all();
        }
    });
}

Worker::~Worker()
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    th_.join();
}

void Worker::submit(std::function<void()> fn)
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        q_.push_back(std::move(fn));
    }
    cv_.notify_all();
}

void Worker::wait()
{
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return q_.empty() && !busy_; });
}

// ================================================================================================ CPU sets
std::vector<int> affinity_cpus()
{
    std::vector<int> out;
    cpu_set_t s;
    CPU_ZERO(&s);
    if (sched_getaffinity(0, sizeof s, &s) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &s)) out.push_back(c);
    if (out.empty()) {
        unsigned hw = std::thread::hardware_concurrency();
        for (unsigned c = 0; c < (hw ? hw : 1u); ++c) out.push_back((int)c);
    }
    return out;
}

std::vector<int> parse_cpulist(const char* text)
{
    std::vector<int> out;
    const char* p = text;
    while (p && *p) {
        while (*p && !std::isdigit((unsigned char)*p)) ++p;
        if (!*p) break;
        char* e = nullptr;
        long a = std::strtol(p, &e, 10), b = a;
        p = e;
        if (*p == '-') { b = std::strtol(p + 1, &e, 10); p = e; }
        if (a < 0 || b < a || b - a > 4096) return {};
        for (long c = a; c <= b; ++c) out.push_back((int)c);
    }
    return out;
}

std::vector<int> cpus_near_pci_device(const char* bdf)
{
    const std::vector<int> all = affinity_cpus();
    if (!bdf || !*bdf) return all;
    std::string id(bdf);
    for (auto& ch : id) ch = (char)std::tolower((unsigned char)ch);
    if (id.size() > 12) id = id.substr(id.size() - 12);        // "00000000:53:00.0" -> "0000:53:00.0"
    const std::string path = "/sys/bus/pci/devices/" + id + "/local_cpulist";
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return all;
    char buf[4096];
    const size_t n = std::fread(buf, 1, sizeof buf - 1, f);
    std::fclose(f);
    buf[n] = 0;
    const std::vector<int> local = parse_cpulist(buf);
    std::vector<int> out;
    for (int c : all)
        if (std::find(local.begin(), local.end(), c) != local.end()) out.push_back(c);
    return out.empty() ? all : out;
}

std::vector<int> cpu_slice(const std::vector<int>& cpus, int part, int nparts)
{
    if (cpus.empty() || nparts <= 1) return cpus;
    part = ((part % nparts) + nparts) % nparts;
    const size_t n = cpus.size();
    if ((size_t)nparts >= n) return {cpus[(size_t)part % n]};
    const size_t lo = n * (size_t)part / (size_t)nparts, hi = n * (size_t)(part + 1) / (size_t)nparts;
    return std::vector<int>(cpus.begin() + (long)lo, cpus.begin() + (long)hi);
}

// ================================================================================================ RowPlan
bool RowPlan::have_avx512()
{
#if QLH_X86
    if (const char* e = std::getenv("QLNLP_NO_AVX512"))
        if (*e && *e != '0') return false;
    return __builtin_cpu_supports("avx512f");
#else
    return false;
#endif
}

RowPlan::RowPlan(int64_t nnz, const double* tmpl, int64_t nvals, const int32_t* pos)
    : nnz_(nnz), nvals_(nvals), tmpl_(tmpl, tmpl + nnz), vd_((size_t)nnz, 0), zeros_((size_t)nvals + 8, 0.0),
      avx512_(have_avx512())
{
    tmpl_.resize((size_t)nnz + 8, 0.0);            // the line loads may run up to 7 elements past the row image
    for (int64_t i = 0; i < nvals; ++i) vd_[(size_t)pos[i]] = 1;
    for (int a = 0; a < 8; ++a) {
        Aligned& A = al_[a];
        A.head = (int)std::min<int64_t>((8 - a) & 7, nnz);
        const int64_t rest = nnz - A.head;
        A.nlines = rest / 8;
        A.tail = (int)(rest % 8);
        A.mask.assign((size_t)A.nlines, 0);
        A.konst.assign((size_t)A.nlines, 0);
        int64_t vi = 0;
        for (int e = 0; e < A.head; ++e) vi += vd_[(size_t)e];
        for (int64_t l = 0; l < A.nlines; ++l) {
            const int64_t base = A.head + 8 * l;
            unsigned m = 0;
            for (int e = 0; e < 8; ++e) m |= (unsigned)vd_[(size_t)(base + e)] << e;
            A.mask[(size_t)l] = (uint8_t)m;
            uint8_t k = 0;
            for (int e = 0; e < 8; ++e) k |= (tmpl_[(size_t)(base + e)] != 0.0) ? 1 : 0;
            A.konst[(size_t)l] = k;
            if (m) {
                A.touched.push_back({(uint32_t)l, (uint32_t)vi, (uint8_t)m, k});
                vi += __builtin_popcount(m);
            }
        }
        A.tail_src = vi;
    }
}

int64_t RowPlan::touched_lines(int a) const
{
    const Aligned& A = al_[a & 7];
    return (int64_t)A.touched.size() + (A.head ? 1 : 0) + (A.tail ? 1 : 0);
}

void RowPlan::row_generic(const Aligned& A, const double* in, double* out, bool touched_only) const
{
    const double* tm = tmpl_.data();
    const uint8_t* vd = vd_.data();
    int64_t vi = 0;
    for (int e = 0; e < A.head; ++e) {
        if (vd[e]) out[e] = in[vi++];
        else if (!touched_only) out[e] = tm[e];
    }
    auto line = [&](int64_t l, int64_t v0) {
        const int64_t base = A.head + 8 * l;
        const unsigned m = A.mask[(size_t)l];
        double tmp[8];
        for (int e = 0; e < 8; ++e) tmp[e] = ((m >> e) & 1u) ? in[v0++] : tm[base + e];
#if QLH_X86
        for (int j = 0; j < 8; j += 2) _mm_stream_pd(out + base + j, _mm_loadu_pd(tmp + j));
#else
        std::memcpy(out + base, tmp, sizeof tmp);
#endif
        return v0;
    };
    if (touched_only) {
        for (size_t t = 0; t < A.touched.size(); ++t) line(A.touched[t].line, A.touched[t].src);
    } else {
        for (int64_t l = 0; l < A.nlines; ++l) vi = line(l, vi);
    }
    vi = A.tail_src;
    for (int64_t e = nnz_ - A.tail; e < nnz_; ++e) {
        if (vd[e]) out[e] = in[vi++];
        else if (!touched_only) out[e] = tm[e];
    }
}

#if QLH_X86
__attribute__((target("avx512f"))) void RowPlan::row_avx512(const Aligned& A, const double* in, double* out,
                                                            bool touched_only) const
{
    const double* tm = tmpl_.data();
    const uint8_t* vd = vd_.data();
    int64_t vi = 0;
    for (int e = 0; e < A.head; ++e) {
        if (vd[e]) out[e] = in[vi++];
        else if (!touched_only) out[e] = tm[e];
    }
    const double* tl = tm + A.head;
    double* ol = out + A.head;                      // 64-byte aligned by construction
    if (touched_only) {
        const Touched* tt = A.touched.data();
        const size_t nt = A.touched.size();
        for (size_t t = 0; t < nt; ++t) {
            const Touched e = tt[t];
            const int64_t l = e.line;
            // most lines hold no constant besides zeros: their image needs no load
            const __m512d v = e.konst ? _mm512_mask_expandloadu_pd(_mm512_loadu_pd(tl + 8 * l), (__mmask8)e.mask, in + e.src)
                                      : _mm512_maskz_expandloadu_pd((__mmask8)e.mask, in + e.src);
            _mm512_stream_pd(ol + 8 * l, v);
        }
    } else {
        const uint8_t* mask = A.mask.data();
        const uint8_t* konst = A.konst.data();
        for (int64_t l = 0; l < A.nlines; ++l) {
            const __mmask8 m = (__mmask8)mask[l];
            const __m512d v = konst[l] ? _mm512_mask_expandloadu_pd(_mm512_loadu_pd(tl + 8 * l), m, in + vi)
                                       : _mm512_maskz_expandloadu_pd(m, in + vi);
            vi += __builtin_popcount((unsigned)m);
            _mm512_stream_pd(ol + 8 * l, v);
        }
    }
    vi = A.tail_src;
    for (int64_t e = nnz_ - A.tail; e < nnz_; ++e) {
        if (vd[e]) out[e] = in[vi++];
        else if (!touched_only) out[e] = tm[e];
    }
}
#else
void RowPlan::row_avx512(const Aligned& A, const double* in, double* out, bool touched_only) const
{
    row_generic(A, in, out, touched_only);
}
#endif

void RowPlan::build_rows(const double* vals, int64_t ldv, double* out, int64_t ldout, int64_t r0, int64_t r1,
                         bool touched_only) const
{
    for (int64_t r = r0; r < r1; ++r) {
        double* o = out + r * ldout;
        const double* in = vals ? vals + r * ldv : zeros_.data();
        const Aligned& A = al_[(reinterpret_cast<uintptr_t>(o) >> 3) & 7];
        if (avx512_) row_avx512(A, in, o, touched_only);
        else row_generic(A, in, o, touched_only);
    }
#if QLH_X86
    _mm_sfence();
#endif
}

void RowPlan::build(Pool* pool, const double* vals, int64_t ldv, double* out, int64_t ldout, int64_t rows,
                    bool touched_only) const
{
    if (!pool) { build_rows(vals, ldv, out, ldout, 0, rows, touched_only); return; }
    pool->parallel_for(rows, 4, [&](int64_t a, int64_t b) { build_rows(vals, ldv, out, ldout, a, b, touched_only); });
}

void RowPlan::build_async(Pool* pool, const double* vals, int64_t ldv, double* out, int64_t ldout, int64_t rows,
                          bool touched_only) const
{
    if (!pool) { build_rows(vals, ldv, out, ldout, 0, rows, touched_only); return; }
    pool->start(rows, 2, [=](int64_t a, int64_t b) { build_rows(vals, ldv, out, ldout, a, b, touched_only); });
}

}  // namespace qlhost
