// qlnlp.cu -- C ABI (include/qlnlp.h) over the fused evaluator kernel, plus the host-side planner
// (segment plan, Jacobian structure, bounds).  Built with nvcc for sm_100a only; no CPU fallback.
#include "../../include/qlnlp.h"

#include <cuda_runtime.h>
#if defined(__x86_64__)
#include <emmintrin.h>      // _mm_stream_pd: rebuild host rows without read-for-ownership traffic
#endif

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "layout.h"
#include "qlnlp_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(e_ == cudaErrorMemoryAllocation ? QLNLP_ENOMEM : QLNLP_ECUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                \
    } while (0)

constexpr int TICKET_POOL = 64;

// per-stream scratch of the host-pointer pipeline
struct HostLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr; // the chunk's last copy has landed
    int64_t cap = 0;            // evaluations the buffers hold
    double *Z = nullptr, *x0 = nullptr, *xf = nullptr, *f = nullptr, *grad = nullptr, *g = nullptr, *jac = nullptr;
    double* stage = nullptr;    // pinned host rows of SPARSE_TRUE values awaiting expansion (compact transfer)
};

}  // namespace

struct qlnlp_handle_s {
    QlClass cls;
    int device = 0;
    int jac_mode = QLNLP_JAC_SPARSE_BLOCK;
    double x0[QL_NX], xf[QL_NX];
    std::vector<double> cost;          // field-major [QL_NCOST][npad]
    int npad = 0;
    std::vector<QlSeg> segs;
    std::vector<int> seg_begin;

    // device state, created at the first evaluation
    bool dev_ready = false;
    double* d_cost = nullptr;
    double* d_x0xf = nullptr;          // x0[15] | xf[15]
    QlSeg* d_segs = nullptr;
    int* d_seg_begin = nullptr;
    long long* d_dense_lin = nullptr;  // DENSE mode: linear index of every SPARSE_BLOCK value
    double* d_dense = nullptr;         // DENSE mode: m x n grid
    int sm_count = 0;
    int blocks_per_sm[3] = {0, 0, 0};  // [JM_NONE, JM_BLOCK, JM_TRUE]
    size_t smem[3] = {0, 0, 0};
    double rmb = 0, rmf = 0, rIb = 0;  // reciprocals of the divisors
    bool fastdiv = false;              // reciprocal-FMA division verified exact for this model
    int64_t last_launch[5] = {0, 0, 0, 0, 0};
    HostLane lanes[2];
    std::map<cudaStream_t, unsigned*> tickets;   // work counters, one pair per stream the handle has launched on
    unsigned* ticket_pool = nullptr;             // pre-zeroed counters (128 B apart) so that a launch needs no
    int ticket_pool_used = 0;                    // allocation: launches stay legal inside CUDA-graph capture
    std::vector<int32_t> true2block;   // position of every SPARSE_TRUE value inside a SPARSE_BLOCK row
    int64_t ldz_e = 0, ldgrad_e = 0, ldg_e = 0, ldjac_e = 0;   // even leading dimensions of the scratch
};

namespace {

// ---------------------------------------------------------------------------- host planner
// SPARSE_BLOCK structure in value order, straight from the closed-form layout (1-based output).
void sparse_block_structure(const QlClass& c, int64_t* rows, int64_t* cols)
{
    int64_t n = 0;
    for (int k = 1; k <= c.N; ++k) {
        const int ncol = (k < c.N) ? QL_NZK : QL_NX;
        for (int j = 0; j < ncol; ++j) {
            const int64_t col = (int64_t)QL_NZK * (k - 1) + j + 1;
            auto put = [&](int64_t row0) { rows[n] = row0 + 1; cols[n] = col; ++n; };
            if (j < QL_NX) {
                if (k == 1) for (int i = 0; i < QL_NX; ++i) put(i);                              // init rows
                if (k == c.N) for (int i = 0; i < QL_NX - 1; ++i) put(c.c_term + i);             // term rows
                if (k >= 2) for (int i = 0; i < QL_NX; ++i) put(c.c_dyn + QL_NX * (k - 2) + i);  // -I block
                if (k < c.N) for (int i = 0; i < QL_NX; ++i) put(c.c_dyn + QL_NX * (k - 1) + i); // RK4 block
                const int jf = (c.init_mode == 1) ? 4 : 6, jo = (c.init_mode == 1) ? 6 : 4;
                if (j == jf) put(c.c_cfirst + (k - 1));
                if (j == jo && k >= c.k_trans) put(c.c_cother + (k - c.k_trans));
                if (j == 1 || j == 2) put(c.c_body + (k - 1));
            } else {
                for (int i = 0; i < QL_NX; ++i) put(c.c_dyn + QL_NX * (k - 1) + i);
                if (k == c.N - 1 && (j == 16 || j == 18)) put(c.c_fctrl);
            }
        }
    }
}

// Segment plan: per pass, consecutive knots paired; template ids from the constant-image signature.
void plan_segments(const QlClass& c, std::vector<QlSeg>& segs, std::vector<int>& seg_begin)
{
    std::map<std::vector<int>, int> ids;
    segs.clear();
    seg_begin.assign(1, 0);
    for (int p = 0; p < c.npass; ++p) {
        const int ka = p * QL_LANES + 1, kb = std::min(c.N, ka + QL_LANES - 1);
        int idx = 0;
        for (int k = ka; k <= kb; k += 2, ++idx) {
            QlSeg s;
            std::memset(&s, 0, sizeof s);
            const int nk = (k + 1 <= kb) ? 2 : 1;
            s.k0 = (short)k;
            s.nk = (signed char)nk;
            s.start = ql_run_off(c, k);
            s.end = (k + nk > c.N) ? c.nnz : ql_run_off(c, k + nk);
            // everything that determines the constant image of the segment
            std::vector<int> sig{s.start & 1, nk};
            for (int q = k; q < k + nk; ++q) {
                sig.push_back(q == 1);
                sig.push_back(q == c.N - 1);
                sig.push_back(q == c.N);
                sig.push_back(ql_e4(c, q));
                sig.push_back(ql_e6(c, q));
                sig.push_back(q >= c.k_trans);          // RK4 block constants of mode 3 vs the initial mode
                sig.push_back(q == c.k_trans - 1);      // jump knot: masked rows hold 0 instead of 1
            }
            auto it = ids.find(sig);
            if (it == ids.end()) it = ids.emplace(sig, (int)ids.size()).first;
            s.tmpl = (short)it->second;
            s.buf = (signed char)(segs.size() & 1);      // alternate over the whole evaluation
            segs.push_back(s);
        }
        seg_begin.push_back((int)segs.size());
    }
}

// q = a*r; q' = fma(fma(-q, b, a), r, q) is the correctly rounded a/b for r = RN(1/b) (Markstein) unless b's
// significand is all ones.  Verify per divisor on pseudo-random numerators; any miss disables the fast path.
bool fastdiv_is_exact(double b)
{
    if (!(std::isfinite(b)) || b == 0.0) return false;
    const double r = 1.0 / b;
    unsigned long long s = 88172645463325252ULL;
    for (int i = 0; i < 200000; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const double m = 1.0 + (double)(s >> 12) * (1.0 / 4503599627370496.0);
        const double a = std::ldexp((s & 1) ? -m : m, (int)((s >> 3) % 81) - 40);
        const double q = a * r;
        if (std::fma(std::fma(-q, b, a), r, q) != a / b) return false;
    }
    return true;
}

template <int JM>
const void* kernel_fn_jm(bool fast, bool ragged)
{
    if (fast) return ragged ? (const void*)ql::eval_kernel<JM, true, true> : (const void*)ql::eval_kernel<JM, true, false>;
    return ragged ? (const void*)ql::eval_kernel<JM, false, true> : (const void*)ql::eval_kernel<JM, false, false>;
}
// ragged: rows addressed through offset tables (qlnlp_eval_ragged_device) instead of leading dimensions
const void* kernel_fn(int jm, bool fast, bool ragged)
{
    if (jm == ql::JM_BLOCK) return kernel_fn_jm<ql::JM_BLOCK>(fast, ragged);
    if (jm == ql::JM_TRUE) return kernel_fn_jm<ql::JM_TRUE>(fast, ragged);
    return kernel_fn_jm<ql::JM_NONE>(fast, ragged);
}

// the sparse pattern batched evaluations of this handle produce (DENSE handles batch in SPARSE_BLOCK)
int batch_jm(qlnlp_handle h);
int batch_nnz(qlnlp_handle h);

// SPARSE_TRUE structure (1-based, value order): like sparse_block_structure restricted to structural non-zeros
void sparse_true_structure(const QlClass& c, int64_t* rows, int64_t* cols)
{
    int64_t n = 0;
    for (int k = 1; k <= c.N; ++k) {
        const int ncol = (k < c.N) ? QL_NZK : QL_NX;
        const int mode = (k >= c.k_trans) ? 3 : c.init_mode;
        const QlTruePattern pat = ql_true_pattern(mode, k < c.N && k == c.k_trans - 1);
        for (int j = 0; j < ncol; ++j) {
            const int64_t col = (int64_t)QL_NZK * (k - 1) + j + 1;
            auto put = [&](int64_t row0) { rows[n] = row0 + 1; cols[n] = col; ++n; };
            if (j < QL_NX) {
                if (k == 1) put(j);                                              // init diagonal
                if (k == c.N && j < QL_NX - 1) put(c.c_term + j);                // term diagonal
                if (k >= 2) put(c.c_dyn + QL_NX * (k - 2) + j);                  // -I diagonal
            }
            if (k < c.N)
                for (int e = 0; e < pat.n; ++e)
                    if (pat.J[e] == j) put(c.c_dyn + QL_NX * (k - 1) + pat.I[e]);
            if (j < QL_NX) {
                const int jf = (c.init_mode == 1) ? 4 : 6, jo = (c.init_mode == 1) ? 6 : 4;
                if (j == jf) put(c.c_cfirst + (k - 1));
                if (j == jo && k >= c.k_trans) put(c.c_cother + (k - c.k_trans));
                if (j == 1 || j == 2) put(c.c_body + (k - 1));
            } else if (k == c.N - 1 && (j == 16 || j == 18)) {
                put(c.c_fctrl);
            }
        }
    }
}

int check_handle(qlnlp_handle h)
{
    if (!h) return fail(QLNLP_EINVAL, "null handle");
    return QLNLP_OK;
}

int ensure_device(qlnlp_handle h)
{
    if (h->dev_ready) {
        CUDA_TRY(cudaSetDevice(h->device));
        return QLNLP_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(QLNLP_ENODEVICE, "no CUDA device available (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (h->device < 0 || h->device >= ndev) return fail(QLNLP_EINVAL, "device %d out of range (0..%d)", h->device, ndev - 1);
    CUDA_TRY(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, h->device));
    if (prop.major != 10)
        return fail(QLNLP_ENODEVICE, "device %d is sm_%d%d; this build targets sm_100a (B200) only", h->device, prop.major,
                    prop.minor);
    h->sm_count = prop.multiProcessorCount;

    CUDA_TRY(cudaMalloc(&h->d_cost, h->cost.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(h->d_cost, h->cost.data(), h->cost.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&h->d_x0xf, 2 * QL_NX * sizeof(double)));
    CUDA_TRY(cudaMemcpy(h->d_x0xf, h->x0, QL_NX * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d_x0xf + QL_NX, h->xf, QL_NX * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&h->d_segs, h->segs.size() * sizeof(QlSeg)));
    CUDA_TRY(cudaMemcpy(h->d_segs, h->segs.data(), h->segs.size() * sizeof(QlSeg), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&h->d_seg_begin, h->seg_begin.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpy(h->d_seg_begin, h->seg_begin.data(), h->seg_begin.size() * sizeof(int), cudaMemcpyHostToDevice));

    for (int wj = 0; wj < 3; ++wj) {
        h->smem[wj] = ql::smem_bytes(h->cls.N, wj);
        if (h->smem[wj] > (size_t)prop.sharedMemPerBlockOptin)
            return fail(QLNLP_EINVAL, "N=%d needs %zu B of shared memory per warp (> %zu)", h->cls.N, h->smem[wj],
                        (size_t)prop.sharedMemPerBlockOptin);
      for (int rg = 0; rg < 2; ++rg) {
        const void* fn = kernel_fn(wj, h->fastdiv, rg != 0);
        // the attribute is per FUNCTION, shared by every handle of the process: always raise it to the device
        // limit, never to this handle's own (possibly smaller) requirement
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, QL_LANES, h->smem[wj]));
        if (nb < 1) return fail(QLNLP_ECUDA, "kernel does not fit on an SM");
        if (const char* e = std::getenv("QLNLP_BLOCKS_PER_SM")) {       // tuning knob: fewer resident warps per SM
            const int cap = std::atoi(e);
            if (cap >= 1 && cap < nb) nb = cap;
        }
        h->blocks_per_sm[wj] = rg ? std::min(h->blocks_per_sm[wj], nb) : nb;
      }
    }
    for (auto& ln : h->lanes) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaMalloc(&h->ticket_pool, 128 * TICKET_POOL));
    CUDA_TRY(cudaMemset(h->ticket_pool, 0, 128 * TICKET_POOL));
    // the set-up copies above ran on the legacy default stream and may still be in flight when cudaMemcpy returns
    // (pageable source); launches go to arbitrary, possibly non-blocking streams, so finish the set-up first
    CUDA_TRY(cudaDeviceSynchronize());
    h->dev_ready = true;
    return QLNLP_OK;
}

int launch(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io, cudaStream_t stream, int jm_force = -1,
           const qlnlp_ragged_io* rg = nullptr)
{
    const QlClass& c = h->cls;
    if (B < 0) return fail(QLNLP_EINVAL, "negative batch");
    if (!io || !io->Z) return fail(QLNLP_EINVAL, "io->Z is required");
    if (rg) {
        if (!rg->z_off || (io->g && !rg->g_off) || (io->jac && !rg->jac_off))
            return fail(QLNLP_EINVAL, "ragged launch: offset tables are required for every requested array");
    } else {
    if (io->ldz < c.n_nlp) return fail(QLNLP_EINVAL, "ldz %lld < n_nlp %d", (long long)io->ldz, c.n_nlp);
    if (io->grad && io->ldgrad < c.n_nlp) return fail(QLNLP_EINVAL, "ldgrad %lld < n_nlp %d", (long long)io->ldgrad, c.n_nlp);
    if (io->g && io->ldg < c.m_nlp) return fail(QLNLP_EINVAL, "ldg %lld < m_nlp %d", (long long)io->ldg, c.m_nlp);
    }
    const int jm_jac = jm_force >= 0 ? jm_force : batch_jm(h);
    const int nnz_jac = jm_jac == ql::JM_TRUE ? c.nnz_true : c.nnz;
    if (!rg && io->jac && io->ldjac < nnz_jac) return fail(QLNLP_EINVAL, "ldjac %lld < nnz %d", (long long)io->ldjac, nnz_jac);
    if ((reinterpret_cast<uintptr_t>(io->Z) & 7) != 0) return fail(QLNLP_EINVAL, "Z must be 8-byte aligned");
    if (B == 0) return QLNLP_OK;

    ql::Launch P;
    P.c = c;
    P.rmb = h->rmb; P.rmf = h->rmf; P.rIb = h->rIb;
    P.cost = h->d_cost;
    P.npad = h->npad;
    P.nseg = (int)h->segs.size();
    P.x0_def = h->d_x0xf;
    P.xf_def = h->d_x0xf + QL_NX;
    P.segs = h->d_segs;
    P.seg_begin = h->d_seg_begin;
    P.Z = io->Z; P.ldz = io->ldz;
    P.x0 = io->x0; P.xf = io->xf;
    P.f = io->f;
    P.grad = io->grad; P.ldgrad = io->ldgrad;
    P.g = io->g; P.ldg = io->ldg;
    P.jac = io->jac; P.ldjac = io->ldjac;
    P.B = B;
    P.index = rg ? reinterpret_cast<const long long*>(rg->index) : nullptr;
    P.z_off = rg ? reinterpret_cast<const long long*>(rg->z_off) : nullptr;
    P.g_off = rg ? reinterpret_cast<const long long*>(rg->g_off) : nullptr;
    P.j_off = rg ? reinterpret_cast<const long long*>(rg->jac_off) : nullptr;
    P.bulk = (io->jac && (reinterpret_cast<uintptr_t>(io->jac) & 15) == 0 && (rg || (io->ldjac & 1) == 0)) ? 1 : 0;
    if (rg) P.zbulk = ((reinterpret_cast<uintptr_t>(io->Z) & 15) == 0 && (rg->flags & QLNLP_RAGGED_Z_PADDED)) ? 1 : 0;
    else P.zbulk = ((reinterpret_cast<uintptr_t>(io->Z) & 15) == 0 && (io->ldz & 1) == 0) ? 1 : 0;   // ldz even > n_nlp (odd)

    const int wj = io->jac ? jm_jac : ql::JM_NONE;
    // Resident warps per SM.  The SPARSE_BLOCK stream is store-bound and the memory system takes the output of a
    // few fast warps better than that of all 8 that fit (occupancy sweeps in profiles/r01_ablation.md, section 7):
    // 5 per SM, 6 for short batches that also want the cost/gradient.
    int per_sm = h->blocks_per_sm[wj];
    if (wj == ql::JM_BLOCK && !std::getenv("QLNLP_BLOCKS_PER_SM")) {
        const bool want_cost = io->f || io->grad;
        const bool short_batch = B < 8 * (int64_t)h->sm_count * 6;
        per_sm = std::min(per_sm, (want_cost && short_batch) ? 6 : 5);
    }
    const int64_t resident = (int64_t)h->sm_count * per_sm;
    const int grid = (int)std::min<int64_t>(B, resident);
    // work counter of this stream (launches on one stream are ordered, so they can share it; the kernel's last CTA
    // re-arms it).  Concurrent launches of the handle on different streams get different counters.
    auto it = h->tickets.find(stream);
    if (it == h->tickets.end()) {
        unsigned* d = nullptr;
        if (h->ticket_pool_used < TICKET_POOL) {
            d = h->ticket_pool + 32 * h->ticket_pool_used++;      // zeroed (and synchronised) at set-up
        } else {
            CUDA_TRY(cudaMalloc(&d, 128));
            // zero it ON THIS STREAM: cudaMemset on device memory is asynchronous (legacy default stream) and would
            // not be ordered before a launch on a non-blocking stream
            CUDA_TRY(cudaMemsetAsync(d, 0, 128, stream));
        }
        it = h->tickets.emplace(stream, d).first;
    }
    P.ticket = it->second;
    void* args[] = {&P};
    CUDA_TRY(cudaLaunchKernel(kernel_fn(wj, h->fastdiv, rg != nullptr), dim3(grid), dim3(QL_LANES), args, h->smem[wj], stream));
    h->last_launch[0] = grid;
    h->last_launch[1] = QL_LANES;
    h->last_launch[2] = (int64_t)h->smem[wj];
    h->last_launch[3] = per_sm;
    h->last_launch[4] = h->sm_count;
    return QLNLP_OK;
}

void free_lane(HostLane& ln)
{
    if (ln.stage) cudaFreeHost(ln.stage);
    ln.stage = nullptr;
    cudaFree(ln.Z); cudaFree(ln.x0); cudaFree(ln.xf); cudaFree(ln.f); cudaFree(ln.grad); cudaFree(ln.g); cudaFree(ln.jac);
    ln.Z = ln.x0 = ln.xf = ln.f = ln.grad = ln.g = ln.jac = nullptr;
    ln.cap = 0;
}

int reserve_lane(qlnlp_handle h, HostLane& ln, int64_t cap)
{
    if (ln.cap >= cap) return QLNLP_OK;
    free_lane(ln);
    const QlClass& c = h->cls;
    h->ldz_e = (c.n_nlp + 1) & ~1;
    h->ldgrad_e = h->ldz_e;
    h->ldg_e = (c.m_nlp + 1) & ~1;
    h->ldjac_e = (batch_nnz(h) + 1) & ~1;
    CUDA_TRY(cudaMalloc(&ln.Z, sizeof(double) * cap * h->ldz_e));
    CUDA_TRY(cudaMalloc(&ln.x0, sizeof(double) * cap * QL_NX));
    CUDA_TRY(cudaMalloc(&ln.xf, sizeof(double) * cap * QL_NX));
    CUDA_TRY(cudaMalloc(&ln.f, sizeof(double) * cap));
    CUDA_TRY(cudaMalloc(&ln.grad, sizeof(double) * cap * h->ldgrad_e));
    CUDA_TRY(cudaMalloc(&ln.g, sizeof(double) * cap * h->ldg_e));
    CUDA_TRY(cudaMalloc(&ln.jac, sizeof(double) * cap * h->ldjac_e));
    if (batch_jm(h) == ql::JM_BLOCK)
        CUDA_TRY(cudaHostAlloc(&ln.stage, sizeof(double) * cap * ((c.nnz_true + 1) & ~1), cudaHostAllocDefault));
    ln.cap = cap;
    return QLNLP_OK;
}

// ---- compact transfer of SPARSE_BLOCK rows to the host -----------------------------------------------------
// 85 % of a SPARSE_BLOCK row are structural zeros.  For host-pointer batches the device therefore produces the
// SPARSE_TRUE values (38.7 KB instead of 257 KB per evaluation cross PCIe) and host threads rebuild the rows the
// caller asked for: zero-fill + scatter through `true2block`.  No arithmetic happens on the host.
void build_true2block(qlnlp_handle h)
{
    const QlClass& c = h->cls;
    std::vector<int64_t> rb(c.nnz), cb(c.nnz), rt(c.nnz_true), ct(c.nnz_true);
    sparse_block_structure(c, rb.data(), cb.data());
    sparse_true_structure(c, rt.data(), ct.data());
    h->true2block.resize(c.nnz_true);
    int64_t j = 0;
    for (int64_t i = 0; i < c.nnz_true; ++i) {          // both lists are column-major sorted; TRUE is a sub-sequence
        while (rb[j] != rt[i] || cb[j] != ct[i]) ++j;
        h->true2block[i] = (int32_t)j;
    }
}

int host_threads()
{
    if (const char* e = std::getenv("QLNLP_HOST_THREADS")) {
        const int n = std::atoi(e);
        if (n > 0) return n;
    }
    unsigned hw = std::thread::hardware_concurrency();
    if (!hw) hw = 4;
    // one process per GPU (torchrun): share the host cores between the ranks of this node
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) {
        const int n = std::atoi(e);
        if (n > 1) hw = std::max(1u, hw / (unsigned)n);
    }
    return (int)std::min<unsigned>(hw, 64);
}

void expand_rows(const qlnlp_handle h, const double* stage, int64_t ldt, double* dst, int64_t lddst, int64_t rows)
{
    const int nnz_t = h->cls.nnz_true, nnz_b = h->cls.nnz;
    const int32_t* map = h->true2block.data();
    const int T = (int)std::min<int64_t>(host_threads(), rows);
    // Rows are produced front to back, a 16-byte pair at a time, merging the (sorted) non-zero positions into
    // a stream of zeros; on x86-64 the pairs go out as non-temporal stores, so the row is written once and
    // never read (a memset + scatter would first pull every line into the cache).
    auto work = [&](int t) {
        for (int64_t r = t; r < rows; r += T) {
            double* out = dst + r * lddst;
            const double* in = stage + r * ldt;
            int i = 0, pos = 0;
            if ((reinterpret_cast<uintptr_t>(out) & 15) != 0 && nnz_b > 0) {     // 8-byte aligned row: peel one element
                out[0] = (i < nnz_t && map[i] == 0) ? in[i++] : 0.0;
                pos = 1;
            }
            for (; pos + 1 < nnz_b; pos += 2) {
                double a = 0.0, b = 0.0;
                if (i < nnz_t && map[i] == pos) a = in[i++];
                if (i < nnz_t && map[i] == pos + 1) b = in[i++];
#if defined(__x86_64__)
                _mm_stream_pd(out + pos, _mm_set_pd(b, a));
#else
                out[pos] = a; out[pos + 1] = b;
#endif
            }
            if (pos < nnz_b) out[pos] = (i < nnz_t && map[i] == pos) ? in[i++] : 0.0;
        }
#if defined(__x86_64__)
        _mm_sfence();
#endif
    };
    if (T <= 1) { work(0); return; }
    std::vector<std::thread> pool;
    pool.reserve(T - 1);
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
}

// rows of `width` doubles: host (ld_h) <-> device (ld_d)
cudaError_t copy_rows(void* dst, int64_t ld_dst, const void* src, int64_t ld_src, int64_t width, int64_t rows,
                      cudaMemcpyKind kind, cudaStream_t s)
{
    if (ld_dst == width && ld_src == width)
        return cudaMemcpyAsync(dst, src, sizeof(double) * width * rows, kind, s);
    return cudaMemcpy2DAsync(dst, sizeof(double) * ld_dst, src, sizeof(double) * ld_src, sizeof(double) * width, rows, kind, s);
}

constexpr int64_t HOST_CHUNK = 512;   // evaluations per pipeline stage
constexpr int64_t COMPACT_MIN_B = 64; // below this the rows are copied as they are

int eval_host(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io)
{
    const QlClass& c = h->cls;
    if (!io || !io->Z) return fail(QLNLP_EINVAL, "io->Z is required");
    if (io->ldz < c.n_nlp) return fail(QLNLP_EINVAL, "ldz < n_nlp");
    if (io->grad && io->ldgrad < c.n_nlp) return fail(QLNLP_EINVAL, "ldgrad < n_nlp");
    if (io->g && io->ldg < c.m_nlp) return fail(QLNLP_EINVAL, "ldg < m_nlp");
    if (io->jac && io->ldjac < batch_nnz(h)) return fail(QLNLP_EINVAL, "ldjac < nnz");
    if (B <= 0) return B == 0 ? QLNLP_OK : fail(QLNLP_EINVAL, "negative batch");
    const int64_t chunk = std::min<int64_t>(B, HOST_CHUNK);
    const int nlanes = (B > chunk) ? 2 : 1;
    for (int l = 0; l < nlanes; ++l) {
        int rc = reserve_lane(h, h->lanes[l], chunk);
        if (rc) return rc;
    }
    // SPARSE_BLOCK rows for a host caller: ship the structural non-zeros, rebuild the rows with host threads
    const bool compact = io->jac && batch_jm(h) == ql::JM_BLOCK && B >= COMPACT_MIN_B;
    const int64_t ldt = (c.nnz_true + 1) & ~1;
    if (compact && h->true2block.empty()) build_true2block(h);

    struct Pending { HostLane* ln; int64_t b0, nb; } prev = {nullptr, 0, 0};
    auto finish = [&](const Pending& pd) -> int {
        if (!pd.ln) return QLNLP_OK;
        CUDA_TRY(cudaEventSynchronize(pd.ln->done));
        expand_rows(h, pd.ln->stage, ldt, io->jac + pd.b0 * io->ldjac, io->ldjac, pd.nb);
        return QLNLP_OK;
    };
    int li = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, li ^= 1) {
        HostLane& ln = h->lanes[nlanes == 2 ? li : 0];
        const int64_t nb = std::min(chunk, B - b0);
        cudaStream_t s = ln.stream;
        CUDA_TRY(copy_rows(ln.Z, h->ldz_e, io->Z + b0 * io->ldz, io->ldz, c.n_nlp, nb, cudaMemcpyHostToDevice, s));
        if (io->x0) CUDA_TRY(cudaMemcpyAsync(ln.x0, io->x0 + b0 * QL_NX, sizeof(double) * nb * QL_NX, cudaMemcpyHostToDevice, s));
        if (io->xf) CUDA_TRY(cudaMemcpyAsync(ln.xf, io->xf + b0 * QL_NX, sizeof(double) * nb * QL_NX, cudaMemcpyHostToDevice, s));
        qlnlp_batch_io d;
        std::memset(&d, 0, sizeof d);
        d.Z = ln.Z; d.ldz = h->ldz_e;
        d.x0 = io->x0 ? ln.x0 : nullptr;
        d.xf = io->xf ? ln.xf : nullptr;
        d.f = io->f ? ln.f : nullptr;
        d.grad = io->grad ? ln.grad : nullptr; d.ldgrad = h->ldgrad_e;
        d.g = io->g ? ln.g : nullptr; d.ldg = h->ldg_e;
        d.jac = io->jac ? ln.jac : nullptr; d.ldjac = compact ? ldt : h->ldjac_e;
        int rc = launch(h, nb, &d, s, compact ? ql::JM_TRUE : -1);
        if (rc) return rc;
        if (io->f) CUDA_TRY(cudaMemcpyAsync(io->f + b0, ln.f, sizeof(double) * nb, cudaMemcpyDeviceToHost, s));
        if (io->grad) CUDA_TRY(copy_rows(io->grad + b0 * io->ldgrad, io->ldgrad, ln.grad, h->ldgrad_e, c.n_nlp, nb, cudaMemcpyDeviceToHost, s));
        if (io->g) CUDA_TRY(copy_rows(io->g + b0 * io->ldg, io->ldg, ln.g, h->ldg_e, c.m_nlp, nb, cudaMemcpyDeviceToHost, s));
        if (compact) {
            CUDA_TRY(cudaMemcpyAsync(ln.stage, ln.jac, sizeof(double) * nb * ldt, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaEventRecord(ln.done, s));
            if (int rc2 = finish(prev)) return rc2;          // expand the previous chunk while this one is in flight
            prev = {&ln, b0, nb};
        } else if (io->jac) {
            CUDA_TRY(copy_rows(io->jac + b0 * io->ldjac, io->ldjac, ln.jac, h->ldjac_e, batch_nnz(h), nb, cudaMemcpyDeviceToHost, s));
        }
    }
    if (int rc = finish(prev)) return rc;
    for (int l = 0; l < nlanes; ++l) CUDA_TRY(cudaStreamSynchronize(h->lanes[l].stream));
    return QLNLP_OK;
}

int batch_jm(qlnlp_handle h) { return h->jac_mode == QLNLP_JAC_SPARSE_TRUE ? ql::JM_TRUE : ql::JM_BLOCK; }
int batch_nnz(qlnlp_handle h) { return h->jac_mode == QLNLP_JAC_SPARSE_TRUE ? h->cls.nnz_true : h->cls.nnz; }

}  // namespace

// =============================================================================== C ABI
extern "C" {

int qlnlp_version(void) { return QLNLP_VERSION; }

const char* qlnlp_last_error(void) { return g_err.c_str(); }

int qlnlp_create(const qlnlp_problem_desc* d, int device, int jac_mode, qlnlp_handle* out)
{
    if (!d || !out) return fail(QLNLP_EINVAL, "null argument");
    *out = nullptr;
    if (d->N < 2 || d->N > QL_MAX_N) return fail(QLNLP_EINVAL, "N=%lld outside [2, 1024]", (long long)d->N);
    if (d->k_trans < 1 || d->k_trans > d->N) return fail(QLNLP_EINVAL, "k_trans=%lld outside [1, N]", (long long)d->k_trans);
    if (d->init_mode != 1 && d->init_mode != 2) return fail(QLNLP_EINVAL, "init_mode must be 1 or 2");
    if (jac_mode != QLNLP_JAC_SPARSE_BLOCK && jac_mode != QLNLP_JAC_DENSE && jac_mode != QLNLP_JAC_SPARSE_TRUE) return fail(QLNLP_EINVAL, "unknown jac_mode %d", jac_mode);
    if (!d->Q || !d->R || !d->q || !d->r || !d->c) return fail(QLNLP_EINVAL, "cost tables are required");
    qlnlp_handle h = new (std::nothrow) qlnlp_handle_s();
    if (!h) return fail(QLNLP_ENOMEM, "out of host memory");
    ql_class_init(&h->cls, (int)d->N, (int)d->k_trans, (int)d->init_mode, d->model.g, d->model.mb, d->model.mf, d->model.lb);
    ql_class_finish(&h->cls);
    h->device = device;
    h->jac_mode = jac_mode;
    std::memcpy(h->x0, d->x0, sizeof h->x0);
    std::memcpy(h->xf, d->xf, sizeof h->xf);
    const int N = h->cls.N;
    h->npad = (N + 31) & ~31;
    h->cost.assign((size_t)QL_NCOST * h->npad, 0.0);
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < QL_NX; ++i) {
            h->cost[(size_t)i * h->npad + k] = d->Q[k * QL_NX + i];
            h->cost[(size_t)(15 + i) * h->npad + k] = d->q[k * QL_NX + i];
        }
        for (int i = 0; i < QL_NU; ++i) {
            h->cost[(size_t)(30 + i) * h->npad + k] = d->R[k * QL_NU + i];
            h->cost[(size_t)(35 + i) * h->npad + k] = d->r[k * QL_NU + i];
        }
        h->cost[(size_t)40 * h->npad + k] = d->c[k];
    }
    plan_segments(h->cls, h->segs, h->seg_begin);
    h->rmb = 1.0 / h->cls.mb;
    h->rmf = 1.0 / h->cls.mf;
    h->rIb = 1.0 / h->cls.Ib;
    h->fastdiv = fastdiv_is_exact(h->cls.mb) && fastdiv_is_exact(h->cls.mf) && fastdiv_is_exact(h->cls.Ib) &&
                 fastdiv_is_exact(6.0) && !std::getenv("QLNLP_IEEE_DIV");   // env: force the IEEE-division kernels
    *out = h;
    return QLNLP_OK;
}

int qlnlp_destroy(qlnlp_handle h)
{
    if (!h) return QLNLP_OK;
    if (h->dev_ready) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();     // launches on caller streams may still use the handle's tables and counters
        for (auto& ln : h->lanes) {
            if (ln.stream) cudaStreamSynchronize(ln.stream);
            free_lane(ln);
            if (ln.stream) cudaStreamDestroy(ln.stream);
            if (ln.done) cudaEventDestroy(ln.done);
        }
        for (auto& kv : h->tickets)
            if (kv.second < h->ticket_pool || kv.second >= h->ticket_pool + 32 * TICKET_POOL) cudaFree(kv.second);
        cudaFree(h->ticket_pool);
        cudaFree(h->d_cost); cudaFree(h->d_x0xf); cudaFree(h->d_segs); cudaFree(h->d_seg_begin);
        cudaFree(h->d_dense_lin); cudaFree(h->d_dense);
    }
    delete h;
    return QLNLP_OK;
}

int qlnlp_dims(qlnlp_handle h, int64_t* n_nlp, int64_t* m_nlp, int64_t* nnz, int64_t* nnz_block)
{
    if (int rc = check_handle(h)) return rc;
    if (n_nlp) *n_nlp = h->cls.n_nlp;
    if (m_nlp) *m_nlp = h->cls.m_nlp;
    if (nnz) *nnz = (h->jac_mode == QLNLP_JAC_DENSE) ? (int64_t)h->cls.m_nlp * h->cls.n_nlp : batch_nnz(h);
    if (nnz_block) *nnz_block = h->cls.nnz;
    return QLNLP_OK;
}

int qlnlp_jacobian_structure(qlnlp_handle h, int64_t* rows, int64_t* cols)
{
    if (int rc = check_handle(h)) return rc;
    if (!rows || !cols) return fail(QLNLP_EINVAL, "null output");
    if (h->jac_mode == QLNLP_JAC_DENSE) {
        // vec(Tuple.(CartesianIndices(zeros(m, n)))): column-major, row fastest (moi.jl:31-33)
        int64_t n = 0;
        for (int64_t col = 1; col <= h->cls.n_nlp; ++col)
            for (int64_t row = 1; row <= h->cls.m_nlp; ++row) { rows[n] = row; cols[n] = col; ++n; }
    } else if (h->jac_mode == QLNLP_JAC_SPARSE_TRUE) {
        sparse_true_structure(h->cls, rows, cols);
    } else {
        sparse_block_structure(h->cls, rows, cols);
    }
    return QLNLP_OK;
}

int qlnlp_constraint_bounds(qlnlp_handle h, double* lb, double* ub)
{
    if (int rc = check_handle(h)) return rc;
    if (!lb || !ub) return fail(QLNLP_EINVAL, "null output");
    for (int i = 0; i < h->cls.m_nlp; ++i) { lb[i] = 0.0; ub[i] = 0.0; }         // nlp.jl:66-67
    for (int i = 0; i < h->cls.N; ++i) ub[h->cls.c_body + i] = INFINITY;          // nlp.jl:69
    return QLNLP_OK;
}

int qlnlp_variable_bounds(qlnlp_handle h, double* xl, double* xu)
{
    if (int rc = check_handle(h)) return rc;
    if (!xl || !xu) return fail(QLNLP_EINVAL, "null output");
    const int N = h->cls.N;
    const double half_pi = 3.14159265358979323846 / 2;
    for (int i = 0; i < h->cls.n_nlp; ++i) { xl[i] = -INFINITY; xu[i] = INFINITY; }
    for (int k = 1; k <= N; ++k) {                                                // moi.jl:53-67, 1-based as written
        xl[3 + 20 * (k - 1) - 1] = -half_pi;
        xu[3 + 20 * (k - 1) - 1] = half_pi;
        if (k < N) {
            xl[20 + 20 * (k - 1) - 1] = 0.001;
            xu[20 + 20 * (k - 1) - 1] = 0.02;
            xl[22 + 20 * (k - 1) - 1] = 0.0;      // "lower bound of F": these hit yb / x1 of knot k+1, as in the reference
            xl[24 + 20 * (k - 1) - 1] = 0.0;
        }
    }
    return QLNLP_OK;
}

int qlnlp_eval_batch_device(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io, void* stream)
{
    if (int rc = check_handle(h)) return rc;
    if (int rc = ensure_device(h)) return rc;
    return launch(h, B, io, static_cast<cudaStream_t>(stream));
}

int qlnlp_eval_ragged_device(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io, const qlnlp_ragged_io* rg, void* stream)
{
    if (int rc = check_handle(h)) return rc;
    if (!rg) return fail(QLNLP_EINVAL, "null ragged descriptor");
    if (int rc = ensure_device(h)) return rc;
    return launch(h, B, io, static_cast<cudaStream_t>(stream), -1, rg);
}

int qlnlp_eval_batch_host(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io)
{
    if (int rc = check_handle(h)) return rc;
    if (int rc = ensure_device(h)) return rc;
    return eval_host(h, B, io);
}

static int eval_one(qlnlp_handle h, const double* x, double* f, double* grad, double* g, double* jac)
{
    if (int rc = check_handle(h)) return rc;
    if (!x) return fail(QLNLP_EINVAL, "null x");
    if (int rc = ensure_device(h)) return rc;
    qlnlp_batch_io io;
    std::memset(&io, 0, sizeof io);
    io.Z = x; io.ldz = h->cls.n_nlp;
    io.f = f;
    io.grad = grad; io.ldgrad = h->cls.n_nlp;
    io.g = g; io.ldg = h->cls.m_nlp;
    io.jac = jac; io.ldjac = batch_nnz(h);
    return eval_host(h, 1, &io);
}

int qlnlp_eval_objective(qlnlp_handle h, const double* x, double* f)
{
    if (!f) return fail(QLNLP_EINVAL, "null output");
    return eval_one(h, x, f, nullptr, nullptr, nullptr);
}

int qlnlp_eval_objective_gradient(qlnlp_handle h, const double* x, double* grad)
{
    if (!grad) return fail(QLNLP_EINVAL, "null output");
    return eval_one(h, x, nullptr, grad, nullptr, nullptr);
}

int qlnlp_eval_constraint(qlnlp_handle h, const double* x, double* g)
{
    if (!g) return fail(QLNLP_EINVAL, "null output");
    return eval_one(h, x, nullptr, nullptr, g, nullptr);
}

int qlnlp_eval_constraint_jacobian(qlnlp_handle h, const double* x, double* vals)
{
    if (int rc = check_handle(h)) return rc;
    if (!vals) return fail(QLNLP_EINVAL, "null output");
    if (h->jac_mode != QLNLP_JAC_DENSE) return eval_one(h, x, nullptr, nullptr, nullptr, vals);

    // DENSE: evaluate SPARSE_BLOCK on the device, scatter into the zeroed m x n grid, copy back
    if (!x) return fail(QLNLP_EINVAL, "null x");
    if (int rc = ensure_device(h)) return rc;
    const QlClass& c = h->cls;
    const size_t dense_n = (size_t)c.m_nlp * c.n_nlp;
    if (!h->d_dense_lin) {
        std::vector<int64_t> rows(c.nnz), cols(c.nnz);
        sparse_block_structure(c, rows.data(), cols.data());
        std::vector<long long> lin(c.nnz);
        for (int i = 0; i < c.nnz; ++i) lin[i] = (rows[i] - 1) + (long long)c.m_nlp * (cols[i] - 1);
        CUDA_TRY(cudaMalloc(&h->d_dense_lin, sizeof(long long) * c.nnz));
        CUDA_TRY(cudaMemcpy(h->d_dense_lin, lin.data(), sizeof(long long) * c.nnz, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc(&h->d_dense, sizeof(double) * dense_n));
        CUDA_TRY(cudaDeviceSynchronize());     // set-up copy on the default stream before work on the lane's stream
    }
    HostLane& ln = h->lanes[0];
    if (int rc = reserve_lane(h, ln, 1)) return rc;
    cudaStream_t s = ln.stream;
    CUDA_TRY(cudaMemcpyAsync(ln.Z, x, sizeof(double) * c.n_nlp, cudaMemcpyHostToDevice, s));
    qlnlp_batch_io d;
    std::memset(&d, 0, sizeof d);
    d.Z = ln.Z; d.ldz = h->ldz_e;
    d.jac = ln.jac; d.ldjac = h->ldjac_e;
    if (int rc = launch(h, 1, &d, s)) return rc;
    CUDA_TRY(cudaMemsetAsync(h->d_dense, 0, sizeof(double) * dense_n, s));
    ql::scatter_dense_kernel<<<(c.nnz + 255) / 256, 256, 0, s>>>(ln.jac, h->d_dense_lin, h->d_dense, c.nnz);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(vals, h->d_dense, sizeof(double) * dense_n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return QLNLP_OK;
}

int qlnlp_launch_info(qlnlp_handle h, int64_t info[5])
{
    if (int rc = check_handle(h)) return rc;
    if (!info) return fail(QLNLP_EINVAL, "null output");
    for (int i = 0; i < 5; ++i) info[i] = h->last_launch[i];
    return QLNLP_OK;
}

/* Test hook (not part of the public header): the segment plan, so CPU tests can check it
 * against the structure without a GPU.  out[i*6 + {0..5}] = k0 nk start end tmpl buf. */
int qlnlp_debug_segments(qlnlp_handle h, int64_t* out, int64_t cap, int64_t* nseg)
{
    if (int rc = check_handle(h)) return rc;
    if (nseg) *nseg = (int64_t)h->segs.size();
    if (out) {
        for (size_t i = 0; i < h->segs.size() && (int64_t)i < cap; ++i) {
            const QlSeg& s = h->segs[i];
            int64_t* o = out + 6 * i;
            o[0] = s.k0; o[1] = s.nk; o[2] = s.start; o[3] = s.end; o[4] = s.tmpl; o[5] = s.buf;
        }
    }
    return QLNLP_OK;
}

}  // extern "C"
