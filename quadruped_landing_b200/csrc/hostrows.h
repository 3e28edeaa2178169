// hostrows.h -- host side of the host-pointer (e2e) path: rebuilds the caller's Jacobian rows from the compact
// VALS stream the device ships over PCIe, with a persistent worker pool.  Plain C++ (g++), no CUDA: the CPU tests
// exercise it without a GPU.  Nothing here does arithmetic: rows are assembled from the constant image of the
// pattern (zeros, +-1) and the value-dependent entries computed on the device.
#pragma once

#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace qlhost {

// ---------------------------------------------------------------------------------------------------------
// Persistent worker pool.  parallel_for hands out [0, n) in blocks through an atomic counter; the calling thread
// works too.  Workers spin briefly before they sleep, so back-to-back jobs (one per pipeline chunk) start within
// microseconds.  Workers are pinned to the CPUs given at construction (one each) when pinning is enabled.
class Pool {
public:
    Pool(int nthreads, const std::vector<int>& cpus, bool pin);
    ~Pool();
    int size() const { return (int)workers_.size() + 1; }
    void parallel_for(int64_t n, int64_t block, const std::function<void(int64_t, int64_t)>& fn);
    // The same job on the WORKERS only: returns at once, so the caller can keep feeding the GPU; busy() / finish()
    // tell when it is done.  One job at a time.  A pool without workers runs the job inside start().
    void start(int64_t n, int64_t block, std::function<void(int64_t, int64_t)> fn);
    bool busy() const { return pending_.load(std::memory_order_acquire) != 0; }
    void finish();

private:
    void worker_main(int idx, int cpu);
    void run_blocks();
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::atomic<uint64_t> epoch_{0};          // bumped once per job
    std::atomic<int> pending_{0};             // workers that have not finished the current job
    std::atomic<int64_t> next_{0};
    int64_t n_ = 0, block_ = 1;
    const std::function<void(int64_t, int64_t)>* fn_ = nullptr;
    std::function<void(int64_t, int64_t)> owned_;           // the job of start()
    bool stop_ = false;
};

// One persistent thread that runs submitted closures in order (the per-device driver of a multi-device handle).
class Worker {
public:
    Worker();
    ~Worker();
    void submit(std::function<void()> fn);
    void wait();

private:
    std::thread th_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<std::function<void()>> q_;
    bool busy_ = false, stop_ = false;
};

// CPUs this process may run on (sched_getaffinity), ascending.
std::vector<int> affinity_cpus();
// Parse a sysfs cpulist ("0-15,32-47"); empty on failure.
std::vector<int> parse_cpulist(const char* text);
// The CPUs close to the PCI device `bdf` ("0000:53:00.0", any case) that this process may use; the whole affinity
// set when sysfs has no answer.
std::vector<int> cpus_near_pci_device(const char* bdf);
// Slice `part` of `nparts` of a CPU list (contiguous, balanced; never empty when the list is not).
std::vector<int> cpu_slice(const std::vector<int>& cpus, int part, int nparts);

// ---------------------------------------------------------------------------------------------------------
// Row plan: how to assemble one row of a target pattern (SPARSE_BLOCK or SPARSE_TRUE values of one evaluation)
// from the constant image of the pattern and the VALS stream.
//   tmpl[nnz]      the constant image: what the row holds where it does not depend on Z (0, +1, -1)
//   pos[nvals]     target position of every VALS element, strictly ascending
// Rows are written a 64-byte line at a time with non-temporal stores (AVX-512 expand-loads when the CPU has them),
// so a row is written once and never read.  In TOUCHED mode only the lines that contain a value-dependent entry are
// rewritten (42 % of a SPARSE_BLOCK row at the default instance): the caller's row must already hold the constant
// image (qlnlp_host_output_register).
class RowPlan {
public:
    RowPlan(int64_t nnz, const double* tmpl, int64_t nvals, const int32_t* pos);
    // rows on the pool's workers, asynchronously (Pool::start); the arrays must stay valid until pool->finish()
    void build_async(Pool* pool, const double* vals, int64_t ldv, double* out, int64_t ldout, int64_t rows, bool touched_only) const;
    int64_t nnz() const { return nnz_; }
    int64_t nvals() const { return nvals_; }
    // lines (64 B) rewritten per row in TOUCHED mode for a row that starts `a` doubles into a line
    int64_t touched_lines(int a) const;
    // vals == nullptr: write the constant image only (value-dependent entries as 0)
    void build_rows(const double* vals, int64_t ldv, double* out, int64_t ldout, int64_t r0, int64_t r1, bool touched_only) const;
    // rows through the pool
    void build(Pool* pool, const double* vals, int64_t ldv, double* out, int64_t ldout, int64_t rows, bool touched_only) const;
    static bool have_avx512();

private:
    struct Touched {                      // one 64-byte line that holds value-dependent entries
        uint32_t line;                    // line index inside the row
        uint32_t src;                     // VALS index of its first value
        uint8_t mask;                     // which of its 8 elements are value-dependent
        uint8_t konst;                    // 1: its constant image is not all zeros
    };
    struct Aligned {                      // per row alignment a = (address / 8) % 8
        int head = 0, tail = 0;           // scalar elements before the first / after the last full line
        int64_t nlines = 0;
        std::vector<uint8_t> mask;        // [nlines] bit e set: element e of the line is value-dependent
        std::vector<uint8_t> konst;       // [nlines] 1: the constant image of the line is not all zeros
        std::vector<Touched> touched;     // the lines with mask != 0, in order (what a registered row rewrites)
        int64_t tail_src = 0;             // VALS index of the first value-dependent tail element
    };
    const Aligned& aligned(int a) const { return al_[a]; }
    void row_generic(const Aligned& A, const double* in, double* out, bool touched_only) const;
    void row_avx512(const Aligned& A, const double* in, double* out, bool touched_only) const;
    int64_t nnz_, nvals_;
    std::vector<double> tmpl_;
    std::vector<uint8_t> vd_;             // [nnz] 1: value-dependent
    std::vector<double> zeros_;           // stand-in VALS stream for constant-only builds
    Aligned al_[8];
    bool avx512_;
};

}  // namespace qlhost
