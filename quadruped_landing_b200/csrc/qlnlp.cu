// qlnlp.cu -- C ABI (include/qlnlp.h) over the fused evaluator kernel, plus the host-side planner
// (segment plan, Jacobian structure, bounds) and the host-pointer pipeline.  Built with nvcc for sm_100a only;
// no CPU fallback: every evaluation entry point runs the CUDA kernels or fails.
#include "../../include/qlnlp.h"

#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "hostrows.h"
#include "layout.h"
#include "qlnlp_hess.cuh"
#include "qlnlp_kin.cuh"

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
constexpr int NJM = 4;                 // kernel value streams: JM_NONE, JM_BLOCK, JM_TRUE, JM_VALS
constexpr int MAX_LANES = 4;           // pipeline depth of the host-pointer path
constexpr int64_t COMPACT_MIN_B = 64;  // below this, host-pointer batches copy the pattern rows as they are

// Every entry point that touches the device makes the handle's device current and restores the caller's on exit
// (a multi-GPU host -- torch, CUDA.jl -- must not find its current device changed by a call into this library).
struct DeviceGuard {
    int prev = -1, dev;
    explicit DeviceGuard(int d) : dev(d)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};

// per-stream scratch of the host-pointer pipeline
struct HostLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr; // the chunk's last copy has landed
    int64_t cap = 0;            // evaluations the buffers hold
    double *Z = nullptr, *x0 = nullptr, *xf = nullptr, *f = nullptr, *grad = nullptr, *g = nullptr, *jac = nullptr;
    double* stage = nullptr;    // pinned host rows of VALS values awaiting the row builder
};

// single evaluations (the four MOI callbacks): one launch per decision vector, results cached
struct OneEval {
    cudaStream_t stream = nullptr;
    double* hx = nullptr;       // pinned: x | f grad g vals (packed, the layout of `dout`)
    double* hout = nullptr;
    double* dx = nullptr;       // device: x (padded row)
    double* dout = nullptr;     // device: f(2) | grad(ldz_e) | g(ldg_e) | vals(ldv_e)
    bool zero_copy = false;     // the kernel reads x from / writes the results to the pinned block itself (mapped memory)
    double* mx = nullptr;       // device-side addresses of hx / hout when zero_copy
    double* mout = nullptr;
    int64_t o_grad = 0, o_g = 0, o_vals = 0, total = 0;
    bool valid = false;         // hout holds the results for the x stored in hx
};

struct Registration {
    double* ptr;
    int64_t ld, rows;
};

// class table of a ragged launch (device copy of one QlRagClass per handle) + the launch geometry for its largest horizon
struct RagTable {
    ql::QlRagClass* d_classes = nullptr;
    int ncls = 0;
    QlClass layout;                    // the class with the largest N: sizes the shared-memory carve-up
    bool fastdiv = false;
    size_t smem[NJM] = {0, 0, 0, 0};
    int blocks_per_sm[NJM] = {0, 0, 0, 0};
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

    // multi-device parent: one complete handle per device; the parent itself only answers the integer queries
    std::vector<qlnlp_handle_s*> subs;
    std::vector<std::unique_ptr<qlhost::Worker>> drivers;
    int part = 0, nparts = 1;          // this handle's share of the host cores (multi-device handles split them)

    // device state, created at the first evaluation
    bool dev_ready = false;
    double* d_cost = nullptr;
    double* d_x0xf = nullptr;          // x0[15] | xf[15]
    QlSeg* d_segs = nullptr;
    int* d_seg_begin = nullptr;
    long long* d_dense_lin = nullptr;  // DENSE mode: linear index of every SPARSE_BLOCK value
    double* d_dense = nullptr;         // DENSE mode: m x n grid
    int sm_count = 0;
    int blocks_per_sm[NJM] = {0, 0, 0, 0};
    size_t smem[NJM] = {0, 0, 0, 0};
    double rmb = 0, rmf = 0, rIb = 0;  // reciprocals of the divisors
    bool fastdiv = false;              // reciprocal-FMA division verified exact for this model
    int64_t last_launch[5] = {0, 0, 0, 0, 0};
    HostLane lanes[MAX_LANES];
    OneEval one;
    std::map<cudaStream_t, unsigned*> tickets;   // work counters, one pair per stream the handle has launched on
    unsigned* ticket_pool = nullptr;             // pre-zeroed counters (128 B apart) so that a launch needs no
    int ticket_pool_used = 0;                    // allocation: launches stay legal inside CUDA-graph capture
    std::string pci_bus_id;
    // Lagrangian Hessian (single evaluations on host pointers): device scratch x | lambda | values
    double* d_hess_in = nullptr;
    double* d_hess_out = nullptr;
    int hess_blocks_per_sm = 0;
    size_t smem_per_sm = 0, smem_optin = 0;
    // opt-in kinematic rows (qlnlp_kin.cuh): 2N more constraints, 8N more entries in either sparse pattern
    bool kin = false;
    double leg_reach = 0;              // l1 + l2 + lb/2
    std::vector<int32_t> kin_map;      // caller-visible value order of the batch pattern -> scratch position or -(entry + 1)
    int* d_kin_map = nullptr;
    std::map<cudaStream_t, std::pair<double*, int64_t>> kin_scratch;   // per stream: reference-order rows (device), capacity
    uint64_t serial = 0;               // unique per handle ever created (addresses get reused)
    std::map<std::vector<uint64_t>, RagTable> rag_tables;       // ragged launches led by this handle (key: the classes' serials)

    // host-pointer path: row plan of the handle's batch pattern, worker pool, registered output buffers
    std::unique_ptr<qlhost::RowPlan> plan;
    std::unique_ptr<qlhost::Pool> pool;
    std::vector<Registration> regs;
    int64_t opt_host_chunk = 512;
    int64_t opt_host_threads = 0;      // 0: this handle's share of the process's CPUs
    int64_t opt_pin_threads = 1;
    int64_t opt_x_cache = 1;
    int64_t ldz_e = 0, ldgrad_e = 0, ldg_e = 0, ldjac_e = 0, ldv_e = 0;   // even leading dimensions of the scratch
    int64_t stat_host_rows = 0, stat_host_lines = 0;                     // rows / 64-byte lines the row builder wrote
    double stat_t_enqueue = 0, stat_t_wait = 0, stat_t_build = 0, stat_t_total = 0;   // seconds, host-pointer batches
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
        for (int k = ka; k <= kb; k += 2) {
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
            s.buf = 0;      // unused: the kernel alternates the staging buffers with a running counter (ql_seg_buffer)
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
    if (jm == ql::JM_VALS)      // host-pointer path only: never ragged
        return fast ? (const void*)ql::eval_kernel<ql::JM_VALS, true, false> : (const void*)ql::eval_kernel<ql::JM_VALS, false, false>;
    return kernel_fn_jm<ql::JM_NONE>(fast, ragged);
}

// the sparse pattern batched evaluations of this handle produce (DENSE handles batch in SPARSE_BLOCK)
int batch_jm(const qlnlp_handle_s* h) { return h->jac_mode == QLNLP_JAC_SPARSE_TRUE ? ql::JM_TRUE : ql::JM_BLOCK; }
int std_nnz(const qlnlp_handle_s* h) { return h->jac_mode == QLNLP_JAC_SPARSE_TRUE ? h->cls.nnz_true : h->cls.nnz; }
// what the caller sees: the kinematic rows add 8 entries per knot to either pattern and 2 rows per knot to g
int batch_nnz(const qlnlp_handle_s* h) { return std_nnz(h) + (h->kin ? 8 * h->cls.N : 0); }
int m_rows(const qlnlp_handle_s* h) { return h->cls.m_nlp + (h->kin ? 2 * h->cls.N : 0); }
int jm_nnz(const QlClass& c, int jm) { return jm == ql::JM_TRUE ? c.nnz_true : jm == ql::JM_VALS ? c.nnz_vals : c.nnz; }

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

// ---- what a host-side row is made of ----------------------------------------------------------------------
// Position of every SPARSE_TRUE value inside a SPARSE_BLOCK row (both column-major sorted; TRUE is a sub-sequence).
std::vector<int32_t> true_to_block(const QlClass& c)
{
    std::vector<int64_t> rb(c.nnz), cb(c.nnz), rt(c.nnz_true), ct(c.nnz_true);
    sparse_block_structure(c, rb.data(), cb.data());
    sparse_true_structure(c, rt.data(), ct.data());
    std::vector<int32_t> map(c.nnz_true);
    int64_t j = 0;
    for (int64_t i = 0; i < c.nnz_true; ++i) {
        while (rb[j] != rt[i] || cb[j] != ct[i]) ++j;
        map[i] = (int32_t)j;
    }
    return map;
}

struct ModePattern { int nvar; const unsigned char *vi, *vj; int ncon; const unsigned char *ci, *cj; const double* cv; };
ModePattern mode_pattern(int mode)
{
    if (mode == 1) return {QL_NJ_MODE1, QL_PAT_I_MODE1, QL_PAT_J_MODE1, QL_NJC_MODE1, QL_CPAT_I_MODE1, QL_CPAT_J_MODE1, QL_CPAT_V_MODE1};
    if (mode == 2) return {QL_NJ_MODE2, QL_PAT_I_MODE2, QL_PAT_J_MODE2, QL_NJC_MODE2, QL_CPAT_I_MODE2, QL_CPAT_J_MODE2, QL_CPAT_V_MODE2};
    return {QL_NJ_MODE3, QL_PAT_I_MODE3, QL_PAT_J_MODE3, QL_NJC_MODE3, QL_CPAT_I_MODE3, QL_CPAT_J_MODE3, QL_CPAT_V_MODE3};
}
// rows the jump Jacobian keeps (planar_quadruped.jl:262-263)
const int JUMP_KEEP[QL_NX] = {1, 1, 1, 1, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 0};

// The constant image of a SPARSE_BLOCK row (everything jac_c! assigns that does not depend on Z; 0 elsewhere) and
// the position of every VALS element inside the row, in VALS order (layout.h: ql_vals_run_off; the kernel's writer
// is ql_vals_write_run in true_run.h).
void block_image_and_vals_map(const QlClass& c, std::vector<double>& image, std::vector<int32_t>& vals_pos)
{
    image.assign((size_t)c.nnz, 0.0);
    vals_pos.clear();
    vals_pos.reserve((size_t)c.nnz_vals);
    for (int k = 1; k <= c.N; ++k) {
        const int base = ql_run_off(c, k);
        ql_write_run_constants(c, k, image.data() + base);
        const int theta = base + ql_theta_pos(c, k);
        if (k == c.N) { vals_pos.push_back(theta); continue; }
        const bool jump = (k == c.k_trans - 1);
        const ModePattern mp = mode_pattern(k >= c.k_trans ? 3 : c.init_mode);
        for (int e = 0; e < mp.ncon; ++e)
            image[(size_t)(base + ql_rk4_pos(c, k, mp.ci[e], mp.cj[e]))] = (jump && !JUMP_KEEP[mp.ci[e]]) ? 0.0 : mp.cv[e];
        bool theta_done = false;
        for (int e = 0; e < mp.nvar; ++e) {
            if (jump && !JUMP_KEEP[mp.vi[e]]) continue;          // constant zero at the jump knot: stays in the image
            if (mp.vj[e] >= 3 && !theta_done) { vals_pos.push_back(theta); theta_done = true; }
            vals_pos.push_back(base + ql_rk4_pos(c, k, mp.vi[e], mp.vj[e]));
        }
        if (!theta_done) vals_pos.push_back(theta);
    }
}

// Row plan of the handle's batch pattern (SPARSE_BLOCK, or SPARSE_TRUE as a sub-sequence of it).
int ensure_plan(qlnlp_handle h)
{
    if (h->plan) return QLNLP_OK;
    const QlClass& c = h->cls;
    std::vector<double> image;
    std::vector<int32_t> pos;
    block_image_and_vals_map(c, image, pos);
    if ((int)pos.size() != c.nnz_vals) return fail(QLNLP_ECUDA, "internal: VALS map has %zu entries, layout says %d", pos.size(), c.nnz_vals);
    for (size_t i = 1; i < pos.size(); ++i)
        if (pos[i] <= pos[i - 1]) return fail(QLNLP_ECUDA, "internal: VALS map is not ascending at %zu", i);
    if (batch_jm(h) == ql::JM_TRUE) {
        const std::vector<int32_t> t2b = true_to_block(c);
        std::vector<int32_t> b2t((size_t)c.nnz, -1);
        for (size_t t = 0; t < t2b.size(); ++t) b2t[(size_t)t2b[t]] = (int32_t)t;
        std::vector<double> timage(t2b.size());
        for (size_t t = 0; t < t2b.size(); ++t) timage[t] = image[(size_t)t2b[t]];
        for (auto& p : pos) {
            if (b2t[(size_t)p] < 0) return fail(QLNLP_ECUDA, "internal: value-dependent entry outside SPARSE_TRUE");
            p = b2t[(size_t)p];
        }
        image.swap(timage);
    }
    h->plan.reset(new (std::nothrow) qlhost::RowPlan((int64_t)image.size(), image.data(), (int64_t)pos.size(), pos.data()));
    if (!h->plan) return fail(QLNLP_ENOMEM, "out of host memory");
    return QLNLP_OK;
}

int env_int(const char* name, int dflt)
{
    if (const char* e = std::getenv(name)) {
        const int n = std::atoi(e);
        if (n > 0) return n;
    }
    return dflt;
}

bool env_flag(const char* name, bool dflt)
{
    const char* e = std::getenv(name);
    return (e && *e) ? std::atoi(e) != 0 : dflt;
}

// The worker pool of this handle: its share of the CPUs the process may use.  With one process per GPU (torchrun)
// the ranks of a node split the cores (LOCAL_RANK of LOCAL_WORLD_SIZE); a multi-device handle splits its share
// again between its devices; on a multi-socket host the share is taken from the CPUs next to the GPU.
int ensure_pool(qlnlp_handle h)
{
    if (h->pool) return QLNLP_OK;
    const int lws = env_int("LOCAL_WORLD_SIZE", 1);
    const int lr = std::getenv("LOCAL_RANK") ? std::atoi(std::getenv("LOCAL_RANK")) : 0;
    const std::vector<int> all = qlhost::affinity_cpus();
    std::vector<int> near = qlhost::cpus_near_pci_device(h->pci_bus_id.c_str());
    const int nparts = lws * h->nparts, part = lr * h->nparts + h->part;
    std::vector<int> mine;
    if (near.size() == all.size() || near.empty()) {
        mine = qlhost::cpu_slice(all, part, nparts);
    } else {
        // several CPU groups (sockets): the parts are spread evenly over them
        const int groups = std::max<int>(1, (int)((all.size() + near.size() / 2) / near.size()));
        const int per_group = (nparts + groups - 1) / groups;
        mine = qlhost::cpu_slice(near, part % per_group, per_group);
    }
    int T = (int)mine.size();
    if (h->opt_host_threads > 0) T = (int)h->opt_host_threads;
    T = env_int("QLNLP_HOST_THREADS", T);
    T = std::max(1, std::min(T, 256));
    const bool pin = h->opt_pin_threads != 0 && env_flag("QLNLP_PIN_THREADS", true);
    h->pool.reset(new (std::nothrow) qlhost::Pool(T, mine, pin));
    if (!h->pool) return fail(QLNLP_ENOMEM, "out of host memory");
    return QLNLP_OK;
}

// Dynamic shared memory to ask for so that `per_sm` CTAs fit on an SM but `per_sm + 1` never do, whatever carve-out the
// kernel runs on (see launch(): programmatic dependent launch and spare CTA slots).
size_t padded_smem(size_t smem_per_sm, size_t smem_optin, int per_sm, size_t smem)
{
    if (!env_flag("QLNLP_PAD_SMEM", true)) return smem;
    const size_t pad = (smem_per_sm / (size_t)(per_sm + 1) - 1024 + 128) & ~(size_t)127;
    if (pad > smem && (size_t)per_sm * (pad + 1024) <= smem_per_sm && pad <= smem_optin) return pad;
    return smem;
}

// cudaFuncAttributePreferredSharedMemoryCarveout value for `need` bytes of shared memory per SM (see set_carveout)
int carveout_pct(size_t need, size_t smem_per_sm)
{
    static const size_t kCarveKB[] = {0, 8, 16, 32, 64, 100, 132, 164, 196, 228};
    const size_t unified = (smem_per_sm + 65535) / 65536 * 65536;      // 228 KB -> 256 KB
    size_t cfg = smem_per_sm;
    for (size_t kb : kCarveKB)
        if (kb * 1024 >= need) { cfg = std::min(cfg, kb * 1024); break; }
    return (int)std::min<size_t>(100, cfg * 100 / std::max<size_t>(1, unified));
}

int set_carveout(const void* fn, int resident_blocks, size_t smem_per_block, size_t smem_per_sm)
{
    static std::map<const void*, size_t> g_need;
    static std::mutex g_mu;                    // the devices of a multi-device handle are set up from their driver threads
    std::lock_guard<std::mutex> lk(g_mu);
    size_t& need = g_need[fn];
    need = std::max(need, (size_t)resident_blocks * (smem_per_block + 1024));
    // The attribute is a percentage of the SM's UNIFIED L1 + shared memory (256 KB on sm_100; measured: 50 -> 132 KB,
    // 58..65 -> 164 KB, 72 -> 196 KB, >= 79 -> 228 KB), rounded up to a carve-out the SM supports (0, 8, 16, 32, 64, 100,
    // 132, 164, 196, 228 KB).  Ask for the smallest of those that holds `need`.  (Taking the percentage of the 228 KB
    // shared-memory maximum instead left the kernel without a Jacobian on 228 KB of shared memory and 28 KB of L1:
    // ncu launch__shared_mem_config_size; 46 % of its cost-table loads missed L1.  profiles/r02_kernel_ab.md)
    int pct = carveout_pct(need, smem_per_sm);
    pct = env_int("QLNLP_CARVEOUT_PCT", pct);          // tuning knob
    if (!std::getenv("QLNLP_MAX_CARVEOUT"))
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    return QLNLP_OK;
}

int check_handle(qlnlp_handle h)
{
    if (!h) return fail(QLNLP_EINVAL, "null handle");
    return QLNLP_OK;
}
// the handle that owns device state: a multi-device parent forwards single-device work to its first device
qlnlp_handle first(qlnlp_handle h) { return h->subs.empty() ? h : h->subs[0]; }

// Binds the handle to its device (first call: tables, occupancy, streams).  The caller holds a DeviceGuard.
int ensure_device(qlnlp_handle h)
{
    if (h->dev_ready) return QLNLP_OK;
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
    h->smem_per_sm = prop.sharedMemPerMultiprocessor;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, h->device) == cudaSuccess) h->pci_bus_id = bus;

    CUDA_TRY(cudaMalloc(&h->d_cost, h->cost.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(h->d_cost, h->cost.data(), h->cost.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&h->d_x0xf, 2 * QL_NX * sizeof(double)));
    CUDA_TRY(cudaMemcpy(h->d_x0xf, h->x0, QL_NX * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d_x0xf + QL_NX, h->xf, QL_NX * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&h->d_segs, h->segs.size() * sizeof(QlSeg)));
    CUDA_TRY(cudaMemcpy(h->d_segs, h->segs.data(), h->segs.size() * sizeof(QlSeg), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&h->d_seg_begin, h->seg_begin.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpy(h->d_seg_begin, h->seg_begin.data(), h->seg_begin.size() * sizeof(int), cudaMemcpyHostToDevice));

    for (int wj = 0; wj < NJM; ++wj) {
        h->smem[wj] = ql::smem_bytes(h->cls.N, wj);
        if (h->smem[wj] > (size_t)prop.sharedMemPerBlockOptin)
            return fail(QLNLP_EINVAL, "N=%d needs %zu B of shared memory per warp (> %zu)", h->cls.N, h->smem[wj],
                        (size_t)prop.sharedMemPerBlockOptin);
        for (int rg = 0; rg < 2; ++rg) {
            if (rg && wj == ql::JM_VALS) continue;
            const void* fn = kernel_fn(wj, h->fastdiv, rg != 0);
            // the attribute is per FUNCTION, shared by every handle of the process: always raise it to the device
            // limit, never to this handle's own (possibly smaller) requirement
            CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
            CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            int nb = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, QL_LANES, h->smem[wj]));
            if (nb < 1) return fail(QLNLP_ECUDA, "kernel does not fit on an SM");
            // never more resident warps than the kernel was compiled for (its __launch_bounds__)
            if (wj == ql::JM_NONE) nb = std::min(nb, QL_NONE_WARPS);
            else if (wj != ql::JM_BLOCK) nb = std::min(nb, QL_TRUE_WARPS);
            const char* env = std::getenv("QLNLP_BLOCKS_PER_SM");              // tuning knob: fewer resident warps per SM
            if (env) {
                const int cap = std::atoi(env);
                if (cap >= 1 && cap < nb) nb = cap;
            }
            h->blocks_per_sm[wj] = rg ? std::min(h->blocks_per_sm[wj], nb) : nb;
            // Shared memory the resident warps really need: whatever the SM has beyond that serves as L1 (the cost
            // table and the boundary states are re-read by every warp).  The attribute belongs to the FUNCTION, so it
            // only ever grows (handles of other horizons share it).
            {
                const int res = (wj == ql::JM_BLOCK && !env) ? std::min(nb, 6) : nb;
                const size_t per_block = (wj == ql::JM_BLOCK) ? padded_smem(h->smem_per_sm, (size_t)prop.sharedMemPerBlockOptin, res, h->smem[wj]) : h->smem[wj];
                if (int rc = set_carveout(fn, res, per_block, h->smem_per_sm)) return rc;
            }
        }
    }
    for (auto& ln : h->lanes) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&h->one.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaMalloc(&h->ticket_pool, 128 * TICKET_POOL));
    CUDA_TRY(cudaMemset(h->ticket_pool, 0, 128 * TICKET_POOL));
    const QlClass& c = h->cls;
    h->ldz_e = (c.n_nlp + 1) & ~1;
    h->ldgrad_e = h->ldz_e;
    h->ldg_e = (m_rows(h) + 1) & ~1;
    h->ldjac_e = (batch_nnz(h) + 1) & ~1;
    h->ldv_e = (c.nnz_vals + 1) & ~1;
    // the set-up copies above ran on the legacy default stream and may still be in flight when cudaMemcpy returns
    // (pageable source); launches go to arbitrary, possibly non-blocking streams, so finish the set-up first
    CUDA_TRY(cudaDeviceSynchronize());
    h->dev_ready = true;
    return QLNLP_OK;
}

// Work counter of a stream (launches on one stream are ordered, so they can share it; the kernel's last CTA re-arms
// it).  Concurrent launches of the handle on different streams get different counters.
int stream_ticket(qlnlp_handle h, cudaStream_t stream, unsigned** out)
{
    auto it = h->tickets.find(stream);
    if (it == h->tickets.end()) {
        if (h->ticket_pool_used >= TICKET_POOL) {
            // Every counter of the pre-zeroed pool belongs to a stream.  Streams come and go (torch pools), so wait
            // for the handle's outstanding launches and start the pool over rather than allocating (an allocation
            // here would make the launch illegal inside a CUDA-graph capture, and a faulted kernel could leave a
            // counter un-re-armed: the reset also heals that).
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(stream, &cap);
            if (cap != cudaStreamCaptureStatusNone)
                return fail(QLNLP_EINVAL, "more than %d streams used with this handle: cannot recycle work counters during a graph capture", TICKET_POOL);
            CUDA_TRY(cudaDeviceSynchronize());
            CUDA_TRY(cudaMemset(h->ticket_pool, 0, 128 * TICKET_POOL));
            CUDA_TRY(cudaDeviceSynchronize());
            h->tickets.clear();
            h->ticket_pool_used = 0;
        }
        unsigned* d = h->ticket_pool + 32 * h->ticket_pool_used++;      // zeroed (and synchronised) at set-up
        it = h->tickets.emplace(stream, d).first;
    }
    *out = it->second;
    return QLNLP_OK;
}

int launch(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io, cudaStream_t stream, int jm_force = -1,
           const qlnlp_ragged_io* rg = nullptr, const RagTable* table = nullptr, const int32_t* class_of = nullptr)
{
    const QlClass& c = h->cls;
    if (rg && !table) return fail(QLNLP_EINVAL, "internal: ragged launch without a class table");
    if (B < 0) return fail(QLNLP_EINVAL, "negative batch");
    if (B == 0) return QLNLP_OK;          // an empty shard is not an error (its pointers may be NULL)
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
    const int nnz_jac = jm_nnz(c, jm_jac);
    if (!rg && io->jac && io->ldjac < nnz_jac) return fail(QLNLP_EINVAL, "ldjac %lld < nnz %d", (long long)io->ldjac, nnz_jac);
    if ((reinterpret_cast<uintptr_t>(io->Z) & 7) != 0) return fail(QLNLP_EINVAL, "Z must be 8-byte aligned");

    ql::Launch P;
    P.c = rg ? table->layout : c;
    P.classes = rg ? table->d_classes : nullptr;
    P.cls_of = rg ? reinterpret_cast<const int*>(class_of) : nullptr;
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
    int per_sm = rg ? table->blocks_per_sm[wj] : h->blocks_per_sm[wj];
    const size_t smem = rg ? table->smem[wj] : h->smem[wj];
    const bool fast = rg ? table->fastdiv : h->fastdiv;
    if (wj == ql::JM_BLOCK && !std::getenv("QLNLP_BLOCKS_PER_SM")) {
        const bool want_cost = io->f || io->grad;
        const bool short_batch = B < 8 * (int64_t)h->sm_count * 6;
        per_sm = std::min(per_sm, (want_cost && short_batch) ? 6 : 5);
    }
    const int64_t resident = (int64_t)h->sm_count * per_sm;
    int grid = (int)std::min<int64_t>(B, resident);
    if (const char* e = std::getenv("QLNLP_GRID")) {                 // experiment knob: explicit grid size
        const int gsz = std::atoi(e);
        if (gsz >= 1 && gsz <= resident) grid = (int)std::min<int64_t>(B, gsz);
    }
    if (int rc = stream_ticket(h, stream, &P.ticket)) return rc;
    // With programmatic dependent launch the next grid's CTAs are placed as soon as an SM has room for one.  The
    // SPARSE_BLOCK kernel runs on fewer CTAs per SM than would fit (5 or 6 of up to 7 by shared memory): a spare slot
    // would take a CTA of the next grid while this grid is still running at full strength, and the next grid would
    // start out unevenly spread over the SMs (measured: -5 % on the headline once another handle had raised the
    // kernel's carve-out to 228 KB).  Ask for enough shared memory per CTA that per_sm + 1 of them never fit.
    // (Only that kernel: the others are capped by their registers, and padding them would cost L1.)
    const size_t smem_launch = (wj == ql::JM_BLOCK) ? padded_smem(h->smem_per_sm, h->smem_optin, per_sm, smem) : smem;
    void* args[] = {&P};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(QL_LANES);
    cfg.dynamicSmemBytes = smem_launch;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // programmatic dependent launch: back-to-back evaluations on a stream overlap the next grid's start-up with this
    // grid's drain (B = 4,096: +1 % full evaluation, +4..6 % for the compact-output kernels; the kernel waits for its
    // predecessors before it touches any data, see griddepcontrol.wait in eval_kernel).  QLNLP_PDL=0 turns it off.
    cfg.numAttrs = env_flag("QLNLP_PDL", true) ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelExC(&cfg, kernel_fn(wj, fast, rg != nullptr), args));
    h->last_launch[0] = grid;
    h->last_launch[1] = QL_LANES;
    h->last_launch[2] = (int64_t)smem_launch;
    h->last_launch[3] = per_sm;
    h->last_launch[4] = h->sm_count;
    return QLNLP_OK;
}

// The class table of a ragged launch over the handles `hs` (all on one device, already bound to it), cached in hs[0].
int ragged_table(const qlnlp_handle* hs, int ncls, const RagTable** out)
{
    qlnlp_handle lead = hs[0];
    std::vector<uint64_t> key((size_t)ncls);
    for (int i = 0; i < ncls; ++i) key[(size_t)i] = hs[i]->serial;
    auto it = lead->rag_tables.find(key);
    if (it != lead->rag_tables.end()) { *out = &it->second; return QLNLP_OK; }
    RagTable t;
    t.ncls = ncls;
    t.fastdiv = true;
    t.layout = hs[0]->cls;
    std::vector<ql::QlRagClass> host((size_t)ncls);
    for (int i = 0; i < ncls; ++i) {
        qlnlp_handle h = hs[i];
        ql::QlRagClass& r = host[(size_t)i];
        std::memset(&r, 0, sizeof r);
        r.c = h->cls;
        r.rmb = h->rmb; r.rmf = h->rmf; r.rIb = h->rIb;
        r.cost = h->d_cost;
        r.x0_def = h->d_x0xf;
        r.xf_def = h->d_x0xf + QL_NX;
        r.segs = h->d_segs;
        r.seg_begin = h->d_seg_begin;
        r.npad = h->npad;
        r.nseg = (int)h->segs.size();
        t.fastdiv = t.fastdiv && h->fastdiv;
        if (h->cls.N > t.layout.N) t.layout = h->cls;
    }
    // device records are QL_RAGCLASS_BYTES apart (the kernel copies whole slots)
    std::vector<unsigned char> packed((size_t)ncls * QL_RAGCLASS_BYTES, 0);
    for (int i = 0; i < ncls; ++i) std::memcpy(packed.data() + (size_t)i * QL_RAGCLASS_BYTES, &host[(size_t)i], sizeof(ql::QlRagClass));
    void* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, packed.size()));
    CUDA_TRY(cudaMemcpy(d, packed.data(), packed.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaDeviceSynchronize());
    t.d_classes = static_cast<ql::QlRagClass*>(d);
    for (int wj = 0; wj < NJM; ++wj) {
        if (wj == ql::JM_VALS) continue;
        t.smem[wj] = ql::smem_bytes(t.layout.N, wj);
        if (t.smem[wj] > lead->smem_optin) return fail(QLNLP_EINVAL, "N=%d needs %zu B of shared memory per warp", t.layout.N, t.smem[wj]);
        const void* fn = kernel_fn(wj, t.fastdiv, true);
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, QL_LANES, t.smem[wj]));
        if (nb < 1) return fail(QLNLP_ECUDA, "kernel does not fit on an SM");
        if (wj == ql::JM_NONE) nb = std::min(nb, QL_NONE_WARPS);
        else if (wj != ql::JM_BLOCK) nb = std::min(nb, QL_TRUE_WARPS);
        const char* env = std::getenv("QLNLP_BLOCKS_PER_SM");
        if (env) {
            const int cap = std::atoi(env);
            if (cap >= 1 && cap < nb) nb = cap;
        }
        t.blocks_per_sm[wj] = nb;
        {
            const int res = (wj == ql::JM_BLOCK && !env) ? std::min(nb, 6) : nb;
            const size_t per_block = (wj == ql::JM_BLOCK) ? padded_smem(lead->smem_per_sm, lead->smem_optin, res, t.smem[wj]) : t.smem[wj];
            if (int rc = set_carveout(fn, res, per_block, lead->smem_per_sm)) return rc;
        }
    }
    *out = &lead->rag_tables.emplace(key, t).first->second;
    return QLNLP_OK;
}

// ---- opt-in kinematic rows (qlnlp_kin.cuh) ---------------------------------------------------------------------
// The caller-visible structure of a handle with kinematic rows: the reference's entries + 8 per knot, column-major.
// map[o] = position in the reference-order row, or -(8 (k-1) + e + 1) for a kinematic entry.
void kin_structure(const qlnlp_handle_s* h, std::vector<int64_t>& rows, std::vector<int64_t>& cols, std::vector<int32_t>& map)
{
    const QlClass& c = h->cls;
    const int n0 = std_nnz(h);
    std::vector<int64_t> r0((size_t)n0), c0((size_t)n0);
    if (h->jac_mode == QLNLP_JAC_SPARSE_TRUE) sparse_true_structure(c, r0.data(), c0.data());
    else sparse_block_structure(c, r0.data(), c0.data());
    struct Ent { int64_t col, row; int32_t src; };
    std::vector<Ent> all;
    all.reserve((size_t)n0 + 8 * (size_t)c.N);
    for (int i = 0; i < n0; ++i) all.push_back({c0[(size_t)i], r0[(size_t)i], i});
    static const int COL[8] = {0, 1, 3, 4, 0, 1, 5, 6};
    for (int k = 1; k <= c.N; ++k)
        for (int e = 0; e < 8; ++e)
            all.push_back({(int64_t)QL_NZK * (k - 1) + COL[e] + 1, (int64_t)c.m_nlp + 2 * (k - 1) + (e >> 2) + 1, -(8 * (k - 1) + e + 1)});
    std::sort(all.begin(), all.end(), [](const Ent& a, const Ent& b) { return a.col != b.col ? a.col < b.col : a.row < b.row; });
    rows.resize(all.size()); cols.resize(all.size()); map.resize(all.size());
    for (size_t i = 0; i < all.size(); ++i) { rows[i] = all[i].row; cols[i] = all[i].col; map[i] = all[i].src; }
}

int ensure_kin(qlnlp_handle h)
{
    if (h->d_kin_map) return QLNLP_OK;
    std::vector<int64_t> rows, cols;
    kin_structure(h, rows, cols, h->kin_map);
    CUDA_TRY(cudaMalloc(&h->d_kin_map, sizeof(int) * h->kin_map.size()));
    CUDA_TRY(cudaMemcpy(h->d_kin_map, h->kin_map.data(), sizeof(int) * h->kin_map.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaDeviceSynchronize());
    return QLNLP_OK;
}

// launch() for a handle with kinematic rows: io describes the CALLER's arrays (g rows of m_nlp + 2N, jac rows of
// nnz + 8N).  The fused kernel writes the reference's g rows in place and its Jacobian values into a scratch row per
// evaluation; two small kernels append / interleave the kinematic part.
int launch_kin(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io, cudaStream_t stream)
{
    const QlClass& c = h->cls;
    if (B == 0) return QLNLP_OK;
    if (!io || !io->Z) return fail(QLNLP_EINVAL, "io->Z is required");
    if (io->g && io->ldg < m_rows(h)) return fail(QLNLP_EINVAL, "ldg %lld < m_nlp %d", (long long)io->ldg, m_rows(h));
    if (io->jac && io->ldjac < batch_nnz(h)) return fail(QLNLP_EINVAL, "ldjac %lld < nnz %d", (long long)io->ldjac, batch_nnz(h));
    if (int rc = ensure_kin(h)) return rc;
    qlnlp_batch_io d = *io;
    const int64_t lds = (std_nnz(h) + 1) & ~1;
    if (io->jac) {
        auto& sc = h->kin_scratch[stream];
        if (sc.second < B * lds) {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(stream, &cap);
            if (cap != cudaStreamCaptureStatusNone) return fail(QLNLP_EINVAL, "kinematic rows: the scratch would have to grow during a graph capture");
            if (sc.first) { CUDA_TRY(cudaDeviceSynchronize()); cudaFree(sc.first); sc.first = nullptr; sc.second = 0; }
            CUDA_TRY(cudaMalloc(&sc.first, sizeof(double) * (size_t)(B * lds)));
            sc.second = B * lds;
        }
        d.jac = sc.first;
        d.ldjac = lds;
    }
    if (int rc = launch(h, B, &d, stream)) return rc;
    const int threads = 256;
    if (io->g) {
        const long long total = (long long)B * 2 * c.N;
        ql::kin_g_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(io->Z, io->ldz, io->g, io->ldg, B, c.N, c.m_nlp);
        CUDA_TRY(cudaGetLastError());
    }
    if (io->jac) {
        const int nnz_out = batch_nnz(h);
        const long long total = (long long)B * nnz_out;
        ql::kin_expand_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(
            h->d_kin_map, nnz_out, d.jac, lds, io->Z, io->ldz, io->jac, io->ldjac, B);
        CUDA_TRY(cudaGetLastError());
    }
    return QLNLP_OK;
}

// ---- Lagrangian Hessian ----------------------------------------------------------------------------------------
void hessian_structure(const QlClass& c, int64_t* rows, int64_t* cols)
{
    int64_t n = 0;
    for (int k = 1; k <= c.N; ++k) {
        const int64_t base = (int64_t)QL_NZK * (k - 1);
        if (k == c.N) {
            for (int i = 0; i < QL_NX; ++i) { rows[n] = base + i + 1; cols[n] = base + i + 1; ++n; }
            continue;
        }
        const int mode = (k >= c.k_trans) ? 3 : c.init_mode;
        const int len = mode == 3 ? QL_HESS_LEN_MODE3 : QL_HESS_LEN_MODE1;
        const unsigned char* R = mode == 1 ? QL_HESS_R_MODE1 : mode == 2 ? QL_HESS_R_MODE2 : QL_HESS_R_MODE3;
        const unsigned char* C = mode == 1 ? QL_HESS_C_MODE1 : mode == 2 ? QL_HESS_C_MODE2 : QL_HESS_C_MODE3;
        for (int e = 0; e < len; ++e) { rows[n] = base + R[e] + 1; cols[n] = base + C[e] + 1; ++n; }
    }
}

int launch_hessian(qlnlp_handle h, int64_t B, const double* Z, int64_t ldz, const double* sigma, double sigma0,
                   const double* lam, int64_t ldlam, double* H, int64_t ldh, cudaStream_t stream)
{
    const QlClass& c = h->cls;
    if (B < 0) return fail(QLNLP_EINVAL, "negative batch");
    if (B == 0) return QLNLP_OK;
    if (!Z || !lam || !H) return fail(QLNLP_EINVAL, "Z, lambda and H are required");
    if (ldz < c.n_nlp || ldlam < c.m_nlp || ldh < ql_hess_nnz(c)) return fail(QLNLP_EINVAL, "leading dimension too small");
    const void* fn = h->fastdiv ? (const void*)ql::hess_kernel<true> : (const void*)ql::hess_kernel<false>;
    const size_t smem = ql::hess_smem_bytes(c.N);
    if (!h->hess_blocks_per_sm) {
        if (smem > h->smem_optin) return fail(QLNLP_EINVAL, "N=%d needs %zu B of shared memory per warp", c.N, smem);
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, QL_LANES, smem));
        if (nb < 1) return fail(QLNLP_ECUDA, "Hessian kernel does not fit on an SM");
        h->hess_blocks_per_sm = nb;
        if (int rc = set_carveout(fn, nb, smem, h->smem_per_sm)) return rc;      // the rest of the SM's memory is L1 (cost table)
    }
    ql::HessLaunch P;
    P.c = c;
    P.rmb = h->rmb; P.rmf = h->rmf; P.rIb = h->rIb;
    P.cost = h->d_cost; P.npad = h->npad;
    P.Z = Z; P.ldz = ldz;
    P.lam = lam; P.ldlam = ldlam;
    P.sigma = sigma; P.sigma0 = sigma0;
    P.H = H; P.ldh = ldh;
    P.B = B;
    P.bulk = ((reinterpret_cast<uintptr_t>(H) & 15) == 0 && (ldh & 1) == 0) ? 1 : 0;
    P.zbulk = ((reinterpret_cast<uintptr_t>(Z) & 15) == 0 && (ldz & 1) == 0) ? 1 : 0;
    if (int rc = stream_ticket(h, stream, &P.ticket)) return rc;
    const int grid = (int)std::min<int64_t>(B, (int64_t)h->sm_count * h->hess_blocks_per_sm);
    void* args[] = {&P};
    CUDA_TRY(cudaLaunchKernel(fn, dim3(grid), dim3(QL_LANES), args, smem, stream));
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
    CUDA_TRY(cudaMalloc(&ln.Z, sizeof(double) * cap * h->ldz_e));
    CUDA_TRY(cudaMalloc(&ln.x0, sizeof(double) * cap * QL_NX));
    CUDA_TRY(cudaMalloc(&ln.xf, sizeof(double) * cap * QL_NX));
    CUDA_TRY(cudaMalloc(&ln.f, sizeof(double) * cap));
    CUDA_TRY(cudaMalloc(&ln.grad, sizeof(double) * cap * h->ldgrad_e));
    CUDA_TRY(cudaMalloc(&ln.g, sizeof(double) * cap * h->ldg_e));
    CUDA_TRY(cudaMalloc(&ln.jac, sizeof(double) * cap * std::max(h->ldjac_e, h->ldv_e)));
    CUDA_TRY(cudaHostAlloc(&ln.stage, sizeof(double) * cap * h->ldv_e, cudaHostAllocDefault));
    ln.cap = cap;
    return QLNLP_OK;
}

// rows of `width` doubles: host (ld_h) <-> device (ld_d)
cudaError_t copy_rows(void* dst, int64_t ld_dst, const void* src, int64_t ld_src, int64_t width, int64_t rows,
                      cudaMemcpyKind kind, cudaStream_t s)
{
    if (ld_dst == width && ld_src == width)
        return cudaMemcpyAsync(dst, src, sizeof(double) * width * rows, kind, s);
    return cudaMemcpy2DAsync(dst, sizeof(double) * ld_dst, src, sizeof(double) * ld_src, sizeof(double) * width, rows, kind, s);
}

// is [jac, jac + B rows) inside a registered output buffer, on a row boundary, with the registered row stride?
bool is_registered(const qlnlp_handle_s* h, const double* jac, int64_t ldjac, int64_t B)
{
    for (const Registration& r : h->regs) {
        if (ldjac != r.ld || jac < r.ptr) continue;
        const int64_t off = jac - r.ptr;
        if (off % r.ld == 0 && off / r.ld + B <= r.rows) return true;
    }
    return false;
}

// ---- host-pointer batches ------------------------------------------------------------------------------------
// Pipeline over up to MAX_LANES streams, `chunk` evaluations per stage: H2D of the decision vectors, one fused launch,
// D2H of f / grad / g straight into the caller's arrays.  The Jacobian rows cross PCIe as the VALS stream (only the
// value-dependent entries: 2,794 of the 32,161 SPARSE_BLOCK values at the default instance) into a pinned staging
// buffer; while the next chunks are in flight the handle's worker pool assembles the caller's rows from the constant
// image of the pattern and those values (hostrows.cpp) -- every 64-byte line of a row, or, when the caller has
// registered the output buffer (qlnlp_host_output_register), only the lines that hold a value-dependent entry.
int eval_host_impl(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io)
{
    const QlClass& c = h->cls;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, env_int("QLNLP_HOST_CHUNK", (int)h->opt_host_chunk)));
    // Chunk schedule: uniform chunks.  QLNLP_HOST_RAMP=1 puts a quarter and a half chunk at either end (the first rows
    // reach the row builder sooner, the last build is shorter); measured: no consistent gain (875 / 840 vs 856 / 855 k
    // evals/s, profiles/r02_host_path.md) -- the D2H copies, not the fill, bound the pipeline -- so it is off.
    std::vector<int64_t> c_start, c_size;
    {
        std::vector<int64_t> front, back;
        if (B >= 3 * chunk && chunk >= 64 && env_flag("QLNLP_HOST_RAMP", false)) {
            front = {chunk / 4, chunk / 2};
            back = {chunk / 2, chunk / 4};
        }
        int64_t pos = 0, tail = 0;
        for (int64_t v : back) tail += v;
        for (int64_t v : front) { c_start.push_back(pos); c_size.push_back(v); pos += v; }
        while (B - tail - pos > 0) {
            const int64_t v = std::min(chunk, B - tail - pos);
            c_start.push_back(pos); c_size.push_back(v); pos += v;
        }
        for (int64_t v : back) { c_start.push_back(pos); c_size.push_back(v); pos += v; }
    }
    const int64_t nchunks = (int64_t)c_start.size();
    const int nlanes = (int)std::min<int64_t>(MAX_LANES, nchunks);
    for (int l = 0; l < nlanes; ++l)
        if (int rc = reserve_lane(h, h->lanes[l], chunk)) return rc;
    const bool compact = io->jac && B >= COMPACT_MIN_B && !h->kin;      // kinematic rows: the rows come back as they are
    bool touched_only = false;
    if (compact) {
        if (int rc = ensure_plan(h)) return rc;
        if (int rc = ensure_pool(h)) return rc;
        touched_only = is_registered(h, io->jac, io->ldjac, B);
    }
    const int nnz_b = batch_nnz(h);
    // Device-side leading dimensions.  Padded (even) rows give the kernel its TMA / 16-byte paths, but when the caller's
    // rows are tightly packed a padded device copy needs a PITCHED transfer (one DMA descriptor per 9.7 KB row: 36 GB/s
    // instead of 50+).  PCIe is what bounds this path, not the kernel: mirror the caller's packing.
    const int64_t ldz_d = (io->ldz == c.n_nlp) ? c.n_nlp : h->ldz_e;
    const int64_t ldgrad_d = (io->grad && io->ldgrad == c.n_nlp) ? c.n_nlp : h->ldgrad_e;
    const int m_out = m_rows(h);
    const int64_t ldg_d = (io->g && io->ldg == m_out) ? m_out : h->ldg_e;
    const int64_t ldjac_d = (io->jac && io->ldjac == nnz_b) ? nnz_b : h->ldjac_e;      // uncompacted rows only
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();

    const bool vals_first = env_flag("QLNLP_HOST_VALS_FIRST", true);
    auto enqueue = [&](int64_t i) -> int {
        const double t0 = now();
        HostLane& ln = h->lanes[i % nlanes];
        const int64_t b0 = c_start[(size_t)i], nb = c_size[(size_t)i];
        cudaStream_t s = ln.stream;
        CUDA_TRY(copy_rows(ln.Z, ldz_d, io->Z + b0 * io->ldz, io->ldz, c.n_nlp, nb, cudaMemcpyHostToDevice, s));
        if (io->x0) CUDA_TRY(cudaMemcpyAsync(ln.x0, io->x0 + b0 * QL_NX, sizeof(double) * nb * QL_NX, cudaMemcpyHostToDevice, s));
        if (io->xf) CUDA_TRY(cudaMemcpyAsync(ln.xf, io->xf + b0 * QL_NX, sizeof(double) * nb * QL_NX, cudaMemcpyHostToDevice, s));
        qlnlp_batch_io d;
        std::memset(&d, 0, sizeof d);
        d.Z = ln.Z; d.ldz = ldz_d;
        d.x0 = io->x0 ? ln.x0 : nullptr;
        d.xf = io->xf ? ln.xf : nullptr;
        d.f = io->f ? ln.f : nullptr;
        d.grad = io->grad ? ln.grad : nullptr; d.ldgrad = ldgrad_d;
        d.g = io->g ? ln.g : nullptr; d.ldg = ldg_d;
        d.jac = io->jac ? ln.jac : nullptr; d.ldjac = compact ? h->ldv_e : ldjac_d;
        if (int rc = h->kin ? launch_kin(h, nb, &d, s) : launch(h, nb, &d, s, compact ? ql::JM_VALS : -1)) return rc;
        // the staged values go first: the row builder only waits for them, f / grad / g land in the caller's arrays
        // by DMA while it works (the lanes are synchronised before the call returns)
        if (compact && vals_first) {
            CUDA_TRY(cudaMemcpyAsync(ln.stage, ln.jac, sizeof(double) * nb * h->ldv_e, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaEventRecord(ln.done, s));
        }
        if (io->f) CUDA_TRY(cudaMemcpyAsync(io->f + b0, ln.f, sizeof(double) * nb, cudaMemcpyDeviceToHost, s));
        if (io->grad) CUDA_TRY(copy_rows(io->grad + b0 * io->ldgrad, io->ldgrad, ln.grad, ldgrad_d, c.n_nlp, nb, cudaMemcpyDeviceToHost, s));
        if (io->g) CUDA_TRY(copy_rows(io->g + b0 * io->ldg, io->ldg, ln.g, ldg_d, m_out, nb, cudaMemcpyDeviceToHost, s));
        if (compact && !vals_first) {
            CUDA_TRY(cudaMemcpyAsync(ln.stage, ln.jac, sizeof(double) * nb * h->ldv_e, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaEventRecord(ln.done, s));
        } else if (!compact && io->jac) {
            CUDA_TRY(copy_rows(io->jac + b0 * io->ldjac, io->ldjac, ln.jac, ldjac_d, nnz_b, nb, cudaMemcpyDeviceToHost, s));
        }
        h->stat_t_enqueue += now() - t0;
        return QLNLP_OK;
    };
    // Row assembly runs on the pool's workers while this thread keeps the device fed: chunk j is handed to the pool
    // as soon as its copies have landed and the pool is free; chunk i may be enqueued once chunk i - nlanes has been
    // assembled (its pinned stage is reused).
    int64_t enq = 0, handed = 0, built = 0;
    double t_build0 = 0;
    auto progress = [&](bool may_block) -> int {
        if (!compact) { handed = built = enq; return QLNLP_OK; }
        if (handed > built) {                                   // a build is running
            if (h->pool->busy()) {
                if (!may_block) return QLNLP_OK;
                h->pool->finish();
            }
            ++built;
            h->stat_t_build += now() - t_build0;
        }
        if (handed == built && handed < enq) {
            HostLane& ln = h->lanes[handed % nlanes];
            const double t0 = now();
            cudaError_t q = may_block ? cudaEventSynchronize(ln.done) : cudaEventQuery(ln.done);
            h->stat_t_wait += now() - t0;
            if (q == cudaErrorNotReady) return QLNLP_OK;
            if (q != cudaSuccess) return fail(QLNLP_ECUDA, "host pipeline: %s", cudaGetErrorString(q));
            const int64_t b0 = c_start[(size_t)handed], nb = c_size[(size_t)handed];
            double* rows = io->jac + b0 * io->ldjac;
            t_build0 = now();
            h->plan->build_async(h->pool.get(), ln.stage, h->ldv_e, rows, io->ldjac, nb, touched_only);
            ++handed;
            h->stat_host_rows += nb;
            h->stat_host_lines += nb * (touched_only ? h->plan->touched_lines((int)((reinterpret_cast<uintptr_t>(rows) >> 3) & 7))
                                                     : (h->plan->nnz() + 7) / 8);
        }
        return QLNLP_OK;
    };
    for (int64_t i = 0; i < nchunks; ++i) {
        while (compact && i - built >= nlanes)
            if (int rc = progress(true)) return rc;
        if (int rc = enqueue(i)) return rc;
        ++enq;
        if (int rc = progress(false)) return rc;
    }
    while (built < enq)
        if (int rc = progress(true)) return rc;
    const double t_sync = now();
    for (int l = 0; l < nlanes; ++l) CUDA_TRY(cudaStreamSynchronize(h->lanes[l].stream));
    h->stat_t_wait += now() - t_sync;
    h->stat_t_total += now() - t_begin;
    return QLNLP_OK;
}

int eval_host(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io)
{
    const QlClass& c = h->cls;
    if (B == 0) return QLNLP_OK;
    if (B < 0) return fail(QLNLP_EINVAL, "negative batch");
    if (!io || !io->Z) return fail(QLNLP_EINVAL, "io->Z is required");
    if (io->ldz < c.n_nlp) return fail(QLNLP_EINVAL, "ldz < n_nlp");
    if (io->grad && io->ldgrad < c.n_nlp) return fail(QLNLP_EINVAL, "ldgrad < n_nlp");
    if (io->g && io->ldg < m_rows(h)) return fail(QLNLP_EINVAL, "ldg < m_nlp");
    if (io->jac && io->ldjac < batch_nnz(h)) return fail(QLNLP_EINVAL, "ldjac < nnz");
    if (io->jac && (reinterpret_cast<uintptr_t>(io->jac) & 7) != 0) return fail(QLNLP_EINVAL, "jac must be 8-byte aligned");
    DeviceGuard guard(h->device);
    if (int rc = ensure_device(h)) return rc;
    const int rc = eval_host_impl(h, B, io);
    if (rc != QLNLP_OK) {
        // copies into caller-owned host memory (and out of the pinned stage) may still be in flight: finish them
        // before the caller gets its buffers back
        const std::string msg = g_err;
        if (h->pool) h->pool->finish();
        for (auto& ln : h->lanes)
            if (ln.stream) cudaStreamSynchronize(ln.stream);
        cudaGetLastError();
        g_err = msg;
    }
    return rc;
}

// contiguous, balanced split of range(B): the first B % n shards get one extra (sharding.shard_bounds)
void shard_bounds(int64_t B, int n, int s, int64_t* lo, int64_t* hi)
{
    const int64_t q = B / n, r = B % n;
    *lo = s * q + std::min<int64_t>(s, r);
    *hi = *lo + q + (s < r ? 1 : 0);
}

qlnlp_batch_io shard_io(const QlClass&, const qlnlp_batch_io& io, int64_t lo)
{
    qlnlp_batch_io d = io;
    d.Z = io.Z + lo * io.ldz;
    if (io.x0) d.x0 = io.x0 + lo * QL_NX;
    if (io.xf) d.xf = io.xf + lo * QL_NX;
    if (io.f) d.f = io.f + lo;
    if (io.grad) d.grad = io.grad + lo * io.ldgrad;
    if (io.g) d.g = io.g + lo * io.ldg;
    if (io.jac) d.jac = io.jac + lo * io.ldjac;
    return d;
}

// ---- single evaluations -------------------------------------------------------------------------------------
enum { HAVE_F = 1, HAVE_GRAD = 2, HAVE_G = 4, HAVE_J = 8 };

int reserve_one(qlnlp_handle h)
{
    OneEval& o = h->one;
    if (o.hx) return QLNLP_OK;
    o.o_grad = 2;
    o.o_g = o.o_grad + h->ldgrad_e;
    o.o_vals = o.o_g + h->ldg_e;
    o.total = o.o_vals + h->ldv_e;
    // One decision vector is latency-bound: with mapped pinned memory the kernel fetches x and stores its 41 KB of
    // results over PCIe itself, which saves the two copy calls (QLNLP_ONE_ZEROCOPY=0 goes back to explicit copies).
    o.zero_copy = env_flag("QLNLP_ONE_ZEROCOPY", true);
    CUDA_TRY(cudaHostAlloc(&o.hx, sizeof(double) * (h->ldz_e + o.total), o.zero_copy ? cudaHostAllocMapped : cudaHostAllocDefault));
    o.hout = o.hx + h->ldz_e;
    if (o.zero_copy) {
        void* dp = nullptr;
        if (cudaHostGetDevicePointer(&dp, o.hx, 0) == cudaSuccess && dp) {
            o.mx = static_cast<double*>(dp);
            o.mout = o.mx + h->ldz_e;
        } else {
            cudaGetLastError();
            o.zero_copy = false;
        }
    }
    CUDA_TRY(cudaMalloc(&o.dx, sizeof(double) * h->ldz_e));
    CUDA_TRY(cudaMalloc(&o.dout, sizeof(double) * o.total));
    o.valid = false;
    return QLNLP_OK;
}

// One decision vector on host pointers.  Ipopt asks for f, grad f, g and the Jacobian values of an iterate in four
// separate callbacks (moi.jl:1-24): the first callback at a new x evaluates EVERYTHING with one launch (f, grad, g and
// the value-dependent Jacobian entries: one H2D copy, one kernel, one D2H copy of 41 KB) and keeps the results; the
// other callbacks at the same x (compared bit for bit, 9.7 KB) are served from that cache.  The Jacobian values are
// assembled in the caller's array from the constant image of the pattern + the cached entries.
int eval_one(qlnlp_handle hh, const double* x, double* f, double* grad, double* g, double* vals)
{
    if (int rc = check_handle(hh)) return rc;
    if (!x) return fail(QLNLP_EINVAL, "null x");
    qlnlp_handle h = first(hh);
    const QlClass& c = h->cls;
    DeviceGuard guard(h->device);
    if (int rc = ensure_device(h)) return rc;
    if (h->kin) {
        // kinematic rows: the plain route -- copy x in, evaluate what was asked for, copy it out (no x cache)
        HostLane& ln = h->lanes[0];
        if (int rc = reserve_lane(h, ln, 1)) return rc;
        cudaStream_t s = ln.stream;
        CUDA_TRY(cudaMemcpyAsync(ln.Z, x, sizeof(double) * c.n_nlp, cudaMemcpyHostToDevice, s));
        qlnlp_batch_io d;
        std::memset(&d, 0, sizeof d);
        d.Z = ln.Z; d.ldz = h->ldz_e;
        d.f = f ? ln.f : nullptr;
        d.grad = grad ? ln.grad : nullptr; d.ldgrad = h->ldgrad_e;
        d.g = g ? ln.g : nullptr; d.ldg = h->ldg_e;
        d.jac = vals ? ln.jac : nullptr; d.ldjac = h->ldjac_e;
        int rc = launch_kin(h, 1, &d, s);
        cudaError_t e = cudaSuccess;
        if (rc == QLNLP_OK && f) e = cudaMemcpyAsync(f, ln.f, sizeof(double), cudaMemcpyDeviceToHost, s);
        if (rc == QLNLP_OK && e == cudaSuccess && grad) e = cudaMemcpyAsync(grad, ln.grad, sizeof(double) * c.n_nlp, cudaMemcpyDeviceToHost, s);
        if (rc == QLNLP_OK && e == cudaSuccess && g) e = cudaMemcpyAsync(g, ln.g, sizeof(double) * m_rows(h), cudaMemcpyDeviceToHost, s);
        if (rc == QLNLP_OK && e == cudaSuccess && vals) e = cudaMemcpyAsync(vals, ln.jac, sizeof(double) * batch_nnz(h), cudaMemcpyDeviceToHost, s);
        if (rc == QLNLP_OK && e != cudaSuccess) rc = fail(QLNLP_ECUDA, "single evaluation: %s", cudaGetErrorString(e));
        const std::string msg = g_err;
        const cudaError_t es = cudaStreamSynchronize(s);
        if (rc == QLNLP_OK && es != cudaSuccess) return fail(QLNLP_ECUDA, "single evaluation: %s", cudaGetErrorString(es));
        g_err = msg;
        return rc;
    }
    if (int rc = reserve_one(h)) return rc;
    OneEval& o = h->one;
    const bool hit = h->opt_x_cache && o.valid && std::memcmp(o.hx, x, sizeof(double) * c.n_nlp) == 0;
    if (!hit) {
        o.valid = false;
        std::memcpy(o.hx, x, sizeof(double) * c.n_nlp);
        cudaStream_t s = o.stream;
        double* const zin = o.zero_copy ? o.mx : o.dx;
        double* const res = o.zero_copy ? o.mout : o.dout;
        if (!o.zero_copy) CUDA_TRY(cudaMemcpyAsync(o.dx, o.hx, sizeof(double) * c.n_nlp, cudaMemcpyHostToDevice, s));
        qlnlp_batch_io d;
        std::memset(&d, 0, sizeof d);
        d.Z = zin; d.ldz = o.zero_copy ? c.n_nlp : h->ldz_e;        // odd ld: plain (non-TMA) loads from host memory
        d.f = res;
        d.grad = res + o.o_grad; d.ldgrad = h->ldgrad_e;
        d.g = res + o.o_g; d.ldg = h->ldg_e;
        d.jac = res + o.o_vals; d.ldjac = o.zero_copy ? h->ldv_e + 1 : h->ldv_e;   // odd ld: plain stores into host memory
        if (int rc = launch(h, 1, &d, s, ql::JM_VALS)) return rc;
        if (!o.zero_copy) CUDA_TRY(cudaMemcpyAsync(o.hout, o.dout, sizeof(double) * o.total, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        o.valid = true;
    }
    if (f) *f = o.hout[0];
    if (grad) std::memcpy(grad, o.hout + o.o_grad, sizeof(double) * c.n_nlp);
    if (g) std::memcpy(g, o.hout + o.o_g, sizeof(double) * c.m_nlp);
    if (vals) {
        if (int rc = ensure_plan(h)) return rc;
        h->plan->build_rows(o.hout + o.o_vals, h->ldv_e, vals, h->plan->nnz(), 0, 1, false);
    }
    return QLNLP_OK;
}

// DENSE: evaluate SPARSE_BLOCK on the device, scatter into the zeroed m x n grid, copy back
int eval_dense_jacobian(qlnlp_handle hh, const double* x, double* vals)
{
    if (!x) return fail(QLNLP_EINVAL, "null x");
    qlnlp_handle h = first(hh);
    DeviceGuard guard(h->device);
    if (int rc = ensure_device(h)) return rc;
    const QlClass& c = h->cls;
    const int m = m_rows(h), nz = batch_nnz(h);          // DENSE handles batch in SPARSE_BLOCK (+ the kinematic entries)
    const size_t dense_n = (size_t)m * c.n_nlp;
    if (!h->d_dense_lin) {
        std::vector<int64_t> rows((size_t)nz), cols((size_t)nz);
        if (h->kin) { std::vector<int32_t> map; kin_structure(h, rows, cols, map); }
        else sparse_block_structure(c, rows.data(), cols.data());
        std::vector<long long> lin((size_t)nz);
        for (int i = 0; i < nz; ++i) lin[(size_t)i] = (rows[(size_t)i] - 1) + (long long)m * (cols[(size_t)i] - 1);
        CUDA_TRY(cudaMalloc(&h->d_dense_lin, sizeof(long long) * nz));
        CUDA_TRY(cudaMemcpy(h->d_dense_lin, lin.data(), sizeof(long long) * nz, cudaMemcpyHostToDevice));
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
    int rc = h->kin ? launch_kin(h, 1, &d, s) : launch(h, 1, &d, s);
    if (rc == QLNLP_OK) {
        cudaError_t e = cudaMemsetAsync(h->d_dense, 0, sizeof(double) * dense_n, s);
        if (e == cudaSuccess) {
            ql::scatter_dense_kernel<<<(nz + 255) / 256, 256, 0, s>>>(ln.jac, h->d_dense_lin, h->d_dense, nz);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(vals, h->d_dense, sizeof(double) * dense_n, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) rc = fail(QLNLP_ECUDA, "dense Jacobian: %s", cudaGetErrorString(e));
    }
    const std::string msg = g_err;
    const cudaError_t es = cudaStreamSynchronize(s);      // also on the error path: `vals` is caller-owned
    if (rc == QLNLP_OK && es != cudaSuccess) return fail(QLNLP_ECUDA, "dense Jacobian: %s", cudaGetErrorString(es));
    g_err = msg;
    return rc;
}

void destroy_device_state(qlnlp_handle h)
{
    if (!h->dev_ready) return;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();     // launches on caller streams may still use the handle's tables and counters
    for (auto& ln : h->lanes) {
        if (ln.stream) cudaStreamSynchronize(ln.stream);
        free_lane(ln);
        if (ln.stream) cudaStreamDestroy(ln.stream);
        if (ln.done) cudaEventDestroy(ln.done);
    }
    if (h->one.stream) cudaStreamDestroy(h->one.stream);
    if (h->one.hx) cudaFreeHost(h->one.hx);
    cudaFree(h->one.dx); cudaFree(h->one.dout);
    cudaFree(h->d_hess_in); cudaFree(h->d_hess_out);
    cudaFree(h->d_kin_map);
    for (auto& kv : h->kin_scratch) cudaFree(kv.second.first);
    for (auto& kv : h->rag_tables) cudaFree(kv.second.d_classes);
    cudaFree(h->ticket_pool);
    cudaFree(h->d_cost); cudaFree(h->d_x0xf); cudaFree(h->d_segs); cudaFree(h->d_seg_begin);
    cudaFree(h->d_dense_lin); cudaFree(h->d_dense);
}

int create_one(const qlnlp_problem_desc* d, int device, int jac_mode, qlnlp_handle* out)
{
    qlnlp_handle h = new (std::nothrow) qlnlp_handle_s();
    if (!h) return fail(QLNLP_ENOMEM, "out of host memory");
    static std::atomic<uint64_t> g_serial{0};
    h->serial = ++g_serial;
    ql_class_init(&h->cls, (int)d->N, (int)d->k_trans, (int)d->init_mode, d->model.g, d->model.mb, d->model.mf, d->model.lb);
    ql_class_finish(&h->cls);
    h->device = device;
    h->jac_mode = jac_mode & ~QLNLP_WITH_KINEMATICS;
    h->kin = (jac_mode & QLNLP_WITH_KINEMATICS) != 0;
    h->leg_reach = d->model.l1 + d->model.l2 + d->model.lb / 2;
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

int check_desc(const qlnlp_problem_desc* d, int jac_mode_flags)
{
    const int jac_mode = jac_mode_flags & ~QLNLP_WITH_KINEMATICS;
    if (d->N < 2 || d->N > QL_MAX_N) return fail(QLNLP_EINVAL, "N=%lld outside [2, 1024]", (long long)d->N);
    if (d->k_trans < 1 || d->k_trans > d->N) return fail(QLNLP_EINVAL, "k_trans=%lld outside [1, N]", (long long)d->k_trans);
    if (d->init_mode != 1 && d->init_mode != 2) return fail(QLNLP_EINVAL, "init_mode must be 1 or 2");
    if (jac_mode != QLNLP_JAC_SPARSE_BLOCK && jac_mode != QLNLP_JAC_DENSE && jac_mode != QLNLP_JAC_SPARSE_TRUE) return fail(QLNLP_EINVAL, "unknown jac_mode %d", jac_mode);
    if (!d->Q || !d->R || !d->q || !d->r || !d->c) return fail(QLNLP_EINVAL, "cost tables are required");
    return QLNLP_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int qlnlp_version(void) { return QLNLP_VERSION; }

const char* qlnlp_last_error(void) { return g_err.c_str(); }

int qlnlp_create(const qlnlp_problem_desc* d, int device, int jac_mode, qlnlp_handle* out)
{
    if (!d || !out) return fail(QLNLP_EINVAL, "null argument");
    *out = nullptr;
    if (int rc = check_desc(d, jac_mode)) return rc;
    return create_one(d, device, jac_mode, out);
}

int qlnlp_create_multi(const qlnlp_problem_desc* d, const int* devices, int ndev, int jac_mode, qlnlp_handle* out)
{
    if (!d || !out || !devices) return fail(QLNLP_EINVAL, "null argument");
    *out = nullptr;
    if (ndev < 1 || ndev > 64) return fail(QLNLP_EINVAL, "ndev=%d outside [1, 64]", ndev);
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(QLNLP_EINVAL, "device %d listed twice", devices[i]);
    if (int rc = check_desc(d, jac_mode)) return rc;
    qlnlp_handle parent = nullptr;
    if (int rc = create_one(d, devices[0], jac_mode, &parent)) return rc;
    for (int i = 0; i < ndev; ++i) {
        qlnlp_handle sub = nullptr;
        if (int rc = create_one(d, devices[i], jac_mode, &sub)) { qlnlp_destroy(parent); return rc; }
        sub->part = i;
        sub->nparts = ndev;
        parent->subs.push_back(sub);
        parent->drivers.emplace_back(new qlhost::Worker());
    }
    *out = parent;
    return QLNLP_OK;
}

int qlnlp_destroy(qlnlp_handle h)
{
    if (!h) return QLNLP_OK;
    for (auto& w : h->drivers) w->wait();
    h->drivers.clear();
    for (qlnlp_handle s : h->subs) qlnlp_destroy(s);
    destroy_device_state(h);
    delete h;
    return QLNLP_OK;
}

int qlnlp_dims(qlnlp_handle h, int64_t* n_nlp, int64_t* m_nlp, int64_t* nnz, int64_t* nnz_block)
{
    if (int rc = check_handle(h)) return rc;
    if (n_nlp) *n_nlp = h->cls.n_nlp;
    if (m_nlp) *m_nlp = m_rows(h);
    if (nnz) *nnz = (h->jac_mode == QLNLP_JAC_DENSE) ? (int64_t)m_rows(h) * h->cls.n_nlp : batch_nnz(h);
    if (nnz_block) *nnz_block = h->cls.nnz + (h->kin ? 8 * h->cls.N : 0);
    return QLNLP_OK;
}

int qlnlp_devices(qlnlp_handle h, int* devices, int cap, int* ndev)
{
    if (int rc = check_handle(h)) return rc;
    const int n = h->subs.empty() ? 1 : (int)h->subs.size();
    if (ndev) *ndev = n;
    if (devices)
        for (int i = 0; i < n && i < cap; ++i) devices[i] = h->subs.empty() ? h->device : h->subs[i]->device;
    return QLNLP_OK;
}

int qlnlp_shard_bounds(int64_t B, int nshards, int shard, int64_t* lo, int64_t* hi)
{
    if (B < 0 || nshards < 1 || shard < 0 || shard >= nshards || !lo || !hi) return fail(QLNLP_EINVAL, "bad shard arguments");
    shard_bounds(B, nshards, shard, lo, hi);
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
            for (int64_t row = 1; row <= m_rows(h); ++row) { rows[n] = row; cols[n] = col; ++n; }
    } else if (h->kin) {
        std::vector<int64_t> r, c;
        std::vector<int32_t> map;
        kin_structure(h, r, c, map);
        std::copy(r.begin(), r.end(), rows);
        std::copy(c.begin(), c.end(), cols);
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
    for (int i = 0; i < m_rows(h); ++i) { lb[i] = 0.0; ub[i] = 0.0; }            // nlp.jl:66-67
    for (int i = 0; i < h->cls.N; ++i) ub[h->cls.c_body + i] = INFINITY;          // nlp.jl:69
    if (h->kin)                                                                   // nlp.jl:70 (commented out upstream)
        for (int i = 0; i < 2 * h->cls.N; ++i) ub[h->cls.m_nlp + i] = h->leg_reach;
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

int qlnlp_set_option(qlnlp_handle h, const char* name, int64_t value)
{
    if (int rc = check_handle(h)) return rc;
    if (!name) return fail(QLNLP_EINVAL, "null option name");
    auto apply = [&](qlnlp_handle t) -> int {
        const std::string n(name);
        if (n == "host_chunk") { if (value < 1) return fail(QLNLP_EINVAL, "host_chunk must be >= 1"); t->opt_host_chunk = value; }
        else if (n == "host_threads") { if (value < 0) return fail(QLNLP_EINVAL, "host_threads must be >= 0"); t->opt_host_threads = value; t->pool.reset(); }
        else if (n == "pin_threads") { t->opt_pin_threads = value != 0; t->pool.reset(); }
        else if (n == "x_cache") { t->opt_x_cache = value != 0; t->one.valid = false; }
        else return fail(QLNLP_EINVAL, "unknown option '%s'", name);
        return QLNLP_OK;
    };
    if (int rc = apply(h)) return rc;
    for (qlnlp_handle s : h->subs)
        if (int rc = apply(s)) return rc;
    return QLNLP_OK;
}

int qlnlp_eval_batch_device(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io, void* stream)
{
    if (int rc = check_handle(h)) return rc;
    if (!h->subs.empty()) return fail(QLNLP_EINVAL, "multi-device handle: use qlnlp_eval_batch_device_multi");
    if (B == 0) return QLNLP_OK;
    DeviceGuard guard(h->device);
    if (int rc = ensure_device(h)) return rc;
    if (h->kin) return launch_kin(h, B, io, static_cast<cudaStream_t>(stream));
    return launch(h, B, io, static_cast<cudaStream_t>(stream));
}

int qlnlp_eval_batch_device_multi(qlnlp_handle h, const int64_t* B, const qlnlp_batch_io* io, void* const* streams)
{
    if (int rc = check_handle(h)) return rc;
    if (!B || !io) return fail(QLNLP_EINVAL, "null argument");
    const int n = h->subs.empty() ? 1 : (int)h->subs.size();
    for (int i = 0; i < n; ++i) {
        qlnlp_handle s = h->subs.empty() ? h : h->subs[i];
        if (B[i] == 0) continue;
        DeviceGuard guard(s->device);
        if (int rc = ensure_device(s)) return rc;
        cudaStream_t st = streams ? static_cast<cudaStream_t>(streams[i]) : nullptr;
        if (int rc = s->kin ? launch_kin(s, B[i], io + i, st) : launch(s, B[i], io + i, st)) return rc;
    }
    return QLNLP_OK;
}

int qlnlp_synchronize(qlnlp_handle h)
{
    if (int rc = check_handle(h)) return rc;
    const int n = h->subs.empty() ? 1 : (int)h->subs.size();
    for (int i = 0; i < n; ++i) {
        qlnlp_handle s = h->subs.empty() ? h : h->subs[i];
        if (!s->dev_ready) continue;
        DeviceGuard guard(s->device);
        CUDA_TRY(cudaDeviceSynchronize());
    }
    return QLNLP_OK;
}

int qlnlp_eval_ragged_device(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io, const qlnlp_ragged_io* rg, void* stream)
{
    if (int rc = check_handle(h)) return rc;
    if (!h->subs.empty()) return fail(QLNLP_EINVAL, "multi-device handle: ragged launches take a single-device handle");
    if (!rg) return fail(QLNLP_EINVAL, "null ragged descriptor");
    if (h->kin) return fail(QLNLP_EINVAL, "ragged launches do not carry the kinematic rows");
    if (B == 0) return QLNLP_OK;
    DeviceGuard guard(h->device);
    if (int rc = ensure_device(h)) return rc;
    const RagTable* table = nullptr;
    if (int rc = ragged_table(&h, 1, &table)) return rc;
    return launch(h, B, io, static_cast<cudaStream_t>(stream), -1, rg, table, nullptr);
}

int qlnlp_eval_ragged_classes(const qlnlp_handle* hs, int ncls, int64_t B, const qlnlp_batch_io* io,
                              const qlnlp_ragged_io* rg, const int32_t* class_of, void* stream)
{
    if (!hs || ncls < 1 || ncls > 4096) return fail(QLNLP_EINVAL, "bad class list");
    if (!rg) return fail(QLNLP_EINVAL, "null ragged descriptor");
    if (ncls > 1 && !class_of) return fail(QLNLP_EINVAL, "class_of is required with more than one class");
    for (int i = 0; i < ncls; ++i) {
        if (int rc = check_handle(hs[i])) return rc;
        if (!hs[i]->subs.empty()) return fail(QLNLP_EINVAL, "multi-device handle: ragged launches take single-device handles");
        if (hs[i]->device != hs[0]->device) return fail(QLNLP_EINVAL, "the classes of a ragged launch must share one device");
        if (batch_jm(hs[i]) != batch_jm(hs[0])) return fail(QLNLP_EINVAL, "the classes of a ragged launch must share one Jacobian pattern");
        if (hs[i]->kin) return fail(QLNLP_EINVAL, "ragged launches do not carry the kinematic rows");
    }
    if (B == 0) return QLNLP_OK;
    DeviceGuard guard(hs[0]->device);
    for (int i = 0; i < ncls; ++i)
        if (int rc = ensure_device(hs[i])) return rc;
    const RagTable* table = nullptr;
    if (int rc = ragged_table(hs, ncls, &table)) return rc;
    return launch(hs[0], B, io, static_cast<cudaStream_t>(stream), -1, rg, table, class_of);
}

int qlnlp_eval_batch_host(qlnlp_handle h, int64_t B, const qlnlp_batch_io* io)
{
    if (int rc = check_handle(h)) return rc;
    if (h->subs.empty()) return eval_host(h, B, io);
    if (B == 0) return QLNLP_OK;
    if (B < 0 || !io) return fail(QLNLP_EINVAL, "bad batch");
    // contiguous shards, one driver thread per device, no exchange between the devices
    const int n = (int)h->subs.size();
    std::vector<int> rcs(n, QLNLP_OK);
    std::vector<std::string> msgs(n);
    for (int i = 0; i < n; ++i) {
        int64_t lo, hi;
        shard_bounds(B, n, i, &lo, &hi);
        if (hi == lo) continue;
        const qlnlp_batch_io d = shard_io(h->cls, *io, lo);
        qlnlp_handle s = h->subs[i];
        h->drivers[i]->submit([s, d, lo, hi, i, &rcs, &msgs] {
            rcs[i] = eval_host(s, hi - lo, &d);
            if (rcs[i]) msgs[i] = g_err;        // g_err is per thread
        });
    }
    for (auto& w : h->drivers) w->wait();
    for (int i = 0; i < n; ++i)
        if (rcs[i]) return fail(rcs[i], "device %d: %s", h->subs[i]->device, msgs[i].c_str());
    return QLNLP_OK;
}

int qlnlp_host_output_register(qlnlp_handle h, double* jac, int64_t ldjac, int64_t B)
{
    if (int rc = check_handle(h)) return rc;
    if (!jac || B < 1) return fail(QLNLP_EINVAL, "bad buffer");
    if ((reinterpret_cast<uintptr_t>(jac) & 7) != 0) return fail(QLNLP_EINVAL, "jac must be 8-byte aligned");
    if (ldjac < batch_nnz(h)) return fail(QLNLP_EINVAL, "ldjac < nnz");
    if (h->kin) return fail(QLNLP_EINVAL, "registered output rows are not available with the kinematic rows");
    const int n = h->subs.empty() ? 1 : (int)h->subs.size();
    // every device's handle may be asked for any slice of the buffer: all of them learn the registration; the
    // constant image is written once, by the first handle's pool
    for (int i = 0; i < n; ++i) {
        qlnlp_handle s = h->subs.empty() ? h : h->subs[i];
        if (int rc = ensure_plan(s)) return rc;
        for (auto it = s->regs.begin(); it != s->regs.end();)
            it = (it->ptr == jac) ? s->regs.erase(it) : it + 1;
        s->regs.push_back({jac, ldjac, B});
    }
    qlnlp_handle s0 = first(h);
    if (int rc = ensure_pool(s0)) return rc;
    s0->plan->build(s0->pool.get(), nullptr, 0, jac, ldjac, B, false);
    return QLNLP_OK;
}

int qlnlp_host_output_unregister(qlnlp_handle h, double* jac)
{
    if (int rc = check_handle(h)) return rc;
    const int n = h->subs.empty() ? 1 : (int)h->subs.size();
    for (int i = 0; i < n; ++i) {
        qlnlp_handle s = h->subs.empty() ? h : h->subs[i];
        for (auto it = s->regs.begin(); it != s->regs.end();)
            it = (it->ptr == jac) ? s->regs.erase(it) : it + 1;
    }
    return QLNLP_OK;
}

int qlnlp_host_pin(void* ptr, int64_t bytes)
{
    if (!ptr || bytes <= 0) return fail(QLNLP_EINVAL, "bad buffer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(QLNLP_ENODEVICE, "no CUDA device available");
    CUDA_TRY(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return QLNLP_OK;
}

int qlnlp_host_unpin(void* ptr)
{
    if (!ptr) return fail(QLNLP_EINVAL, "bad buffer");
    CUDA_TRY(cudaHostUnregister(ptr));
    return QLNLP_OK;
}

/* Page-locked host memory on 2 MB pages where the kernel grants them (madvise MADV_HUGEPAGE): rows of 257 KB each
 * touch 63 small pages, and on a virtualised host the page walks of a row builder that hops from line to line cost as
 * much as the stores.  The allocation is 2 MB aligned, zero-filled and registered with CUDA. */
int qlnlp_host_alloc(int64_t bytes, void** out)
{
    if (!out || bytes <= 0) return fail(QLNLP_EINVAL, "bad allocation request");
    *out = nullptr;
    const size_t huge = (size_t)2 << 20;
    const size_t len = ((size_t)bytes + huge - 1) / huge * huge;
    void* raw = mmap(nullptr, len + huge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (raw == MAP_FAILED) return fail(QLNLP_ENOMEM, "mmap of %zu bytes failed", len + huge);
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(raw) + huge - 1) / huge * huge);
    // give back the unaligned head and tail so that the mapping is exactly [base, base + len)
    if (base > (char*)raw) munmap(raw, (size_t)(base - (char*)raw));
    char* end = (char*)raw + len + huge;
    if (end > base + len) munmap(base + len, (size_t)(end - (base + len)));
#ifdef MADV_HUGEPAGE
    madvise(base, len, MADV_HUGEPAGE);        // best effort
#endif
    std::memset(base, 0, len);                // fault the pages in (as huge pages where possible) before pinning them
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
        cudaError_t e = cudaHostRegister(base, len, cudaHostRegisterPortable);
        if (e != cudaSuccess) {
            munmap(base, len);
            return fail(QLNLP_ECUDA, "cudaHostRegister of %zu bytes: %s", len, cudaGetErrorString(e));
        }
    } else {
        cudaGetLastError();                   // no device: plain (unpinned) memory, still usable by the CPU-side tests
    }
    *out = base;
    return QLNLP_OK;
}

int qlnlp_host_free(void* ptr, int64_t bytes)
{
    if (!ptr) return QLNLP_OK;
    const size_t huge = (size_t)2 << 20;
    const size_t len = ((size_t)bytes + huge - 1) / huge * huge;
    if (cudaHostUnregister(ptr) != cudaSuccess) cudaGetLastError();
    munmap(ptr, len);
    return QLNLP_OK;
}

/* ---- Lagrangian Hessian.  No reference counterpart: src/moi.jl:26-28 offers [:Grad, :Jac]. ---- */
int qlnlp_hessian_nnz(qlnlp_handle h, int64_t* nnz)
{
    if (int rc = check_handle(h)) return rc;
    if (!nnz) return fail(QLNLP_EINVAL, "null output");
    *nnz = ql_hess_nnz(h->cls);
    return QLNLP_OK;
}

int qlnlp_hessian_structure(qlnlp_handle h, int64_t* rows, int64_t* cols)
{
    if (int rc = check_handle(h)) return rc;
    if (!rows || !cols) return fail(QLNLP_EINVAL, "null output");
    hessian_structure(h->cls, rows, cols);
    return QLNLP_OK;
}

int qlnlp_eval_hessian_batch_device(qlnlp_handle h, int64_t B, const double* Z, int64_t ldz, const double* sigma,
                                    const double* lambda, int64_t ldlambda, double* H, int64_t ldh, void* stream)
{
    if (int rc = check_handle(h)) return rc;
    if (!h->subs.empty()) return fail(QLNLP_EINVAL, "multi-device handle: the Hessian takes a single-device handle");
    if (h->kin) return fail(QLNLP_EINVAL, "the Hessian does not cover the kinematic rows");
    if (B == 0) return QLNLP_OK;
    DeviceGuard guard(h->device);
    if (int rc = ensure_device(h)) return rc;
    return launch_hessian(h, B, Z, ldz, sigma, 1.0, lambda, ldlambda, H, ldh, static_cast<cudaStream_t>(stream));
}

int qlnlp_eval_hessian_lagrangian(qlnlp_handle hh, const double* x, double sigma, const double* lambda, double* vals)
{
    if (int rc = check_handle(hh)) return rc;
    if (!x || !lambda || !vals) return fail(QLNLP_EINVAL, "null argument");
    if (hh->kin) return fail(QLNLP_EINVAL, "the Hessian does not cover the kinematic rows");
    qlnlp_handle h = first(hh);
    const QlClass& c = h->cls;
    DeviceGuard guard(h->device);
    if (int rc = ensure_device(h)) return rc;
    const int64_t ldl = (c.m_nlp + 1) & ~1, ldh = (ql_hess_nnz(c) + 1) & ~1;
    if (!h->d_hess_in) {
        CUDA_TRY(cudaMalloc(&h->d_hess_in, sizeof(double) * (h->ldz_e + ldl)));
        CUDA_TRY(cudaMalloc(&h->d_hess_out, sizeof(double) * ldh));
    }
    cudaStream_t s = h->one.stream;
    CUDA_TRY(cudaMemcpyAsync(h->d_hess_in, x, sizeof(double) * c.n_nlp, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->d_hess_in + h->ldz_e, lambda, sizeof(double) * c.m_nlp, cudaMemcpyHostToDevice, s));
    int rc = launch_hessian(h, 1, h->d_hess_in, h->ldz_e, nullptr, sigma, h->d_hess_in + h->ldz_e, ldl, h->d_hess_out, ldh, s);
    if (rc == QLNLP_OK) {
        cudaError_t e = cudaMemcpyAsync(vals, h->d_hess_out, sizeof(double) * ql_hess_nnz(c), cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) rc = fail(QLNLP_ECUDA, "Hessian: %s", cudaGetErrorString(e));
    }
    const std::string msg = g_err;
    const cudaError_t es = cudaStreamSynchronize(s);      // also on the error path: x / lambda / vals are caller-owned
    if (rc == QLNLP_OK && es != cudaSuccess) return fail(QLNLP_ECUDA, "Hessian: %s", cudaGetErrorString(es));
    g_err = msg;
    return rc;
}

/* Batched initial guesses on the device (main.ipynb:181-196): Z[b] = the class guess `base` with the
 * first 14 states of knots 1..k_trans interpolated from x0[b] to the handle's terminal state.  All device pointers. */
int qlnlp_initial_guess_batch_device(qlnlp_handle hh, int64_t B, const double* base, const double* x0, double* Z,
                                     int64_t ldz, void* stream)
{
    if (int rc = check_handle(hh)) return rc;
    if (B == 0) return QLNLP_OK;
    if (B < 0 || !base || !x0 || !Z) return fail(QLNLP_EINVAL, "bad arguments");
    qlnlp_handle h = first(hh);
    if (ldz < h->cls.n_nlp) return fail(QLNLP_EINVAL, "ldz < n_nlp");
    DeviceGuard guard(h->device);
    if (int rc = ensure_device(h)) return rc;
    const long long total = (long long)B * h->cls.n_nlp;
    const int threads = 256;
    ql::initial_guess_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        base, x0, h->d_x0xf + QL_NX, Z, ldz, B, h->cls.n_nlp, h->cls.k_trans);
    CUDA_TRY(cudaGetLastError());
    return QLNLP_OK;
}

int qlnlp_eval_all(qlnlp_handle h, const double* x, double* f, double* grad, double* g, double* vals)
{
    if (int rc = check_handle(h)) return rc;
    if (vals && h->jac_mode == QLNLP_JAC_DENSE) {
        if (int rc = eval_one(h, x, f, grad, g, nullptr)) return rc;
        return eval_dense_jacobian(h, x, vals);
    }
    return eval_one(h, x, f, grad, g, vals);
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
    return eval_dense_jacobian(h, x, vals);
}

int qlnlp_launch_info(qlnlp_handle h, int64_t info[5])
{
    if (int rc = check_handle(h)) return rc;
    if (!info) return fail(QLNLP_EINVAL, "null output");
    for (int i = 0; i < 5; ++i) info[i] = first(h)->last_launch[i];
    return QLNLP_OK;
}

int qlnlp_host_path_info(qlnlp_handle h, int64_t info[8])
{
    if (int rc = check_handle(h)) return rc;
    if (!info) return fail(QLNLP_EINVAL, "null output");
    qlnlp_handle s = first(h);
    if (int rc = ensure_plan(s)) return rc;
    int64_t rows = 0, lines = 0;
    const int n = h->subs.empty() ? 1 : (int)h->subs.size();
    for (int i = 0; i < n; ++i) {
        const qlnlp_handle t = h->subs.empty() ? h : h->subs[i];
        rows += t->stat_host_rows;
        lines += t->stat_host_lines;
    }
    info[0] = s->pool ? s->pool->size() : 0;          // worker threads per device (0: pool not started yet)
    info[1] = s->cls.nnz_vals;                        // doubles per evaluation that cross PCIe for the Jacobian
    info[2] = s->plan->nnz();                         // doubles per row of the batch pattern
    info[3] = s->plan->touched_lines(0);              // 64-byte lines rewritten per registered row (aligned row)
    info[4] = (s->plan->nnz() + 7) / 8;               // 64-byte lines per row
    info[5] = qlhost::RowPlan::have_avx512() ? 1 : 0;
    info[6] = rows;                                   // rows assembled so far
    info[7] = lines;                                  // lines written so far
    return QLNLP_OK;
}

/* seconds spent so far by host-pointer batches of the first device: [0] total, [1] enqueueing copies and launches,
 * [2] waiting for the device / PCIe, [3] assembling rows (not part of the public header: tools/e2e_probe.py) */
int qlnlp_debug_host_times(qlnlp_handle h, double out[4])
{
    if (int rc = check_handle(h)) return rc;
    qlnlp_handle s = first(h);
    out[0] = s->stat_t_total; out[1] = s->stat_t_enqueue; out[2] = s->stat_t_wait; out[3] = s->stat_t_build;
    return QLNLP_OK;
}

/* Launch geometry arithmetic, checkable without a device: out[0] = dynamic shared memory a SPARSE_BLOCK launch asks for
 * so that per_sm CTAs fit on an SM but per_sm + 1 never do, out[1] = the carve-out percentage requested for per_sm such
 * CTAs (not part of the public header: tests/test_cabi_cpu.py) */
int qlnlp_debug_launch_geometry(int64_t smem_per_sm, int64_t smem_optin, int per_sm, int64_t smem, int64_t out[2])
{
    if (!out || smem_per_sm <= 0 || per_sm < 1 || smem < 0) return fail(QLNLP_EINVAL, "bad arguments");
    const size_t padded = padded_smem((size_t)smem_per_sm, (size_t)smem_optin, per_sm, (size_t)smem);
    out[0] = (int64_t)padded;
    out[1] = carveout_pct((size_t)per_sm * (padded + 1024), (size_t)smem_per_sm);
    return QLNLP_OK;
}

/* Test hooks (not part of the public header): let CPU tests check the integer logic without a GPU.
 * qlnlp_debug_segments: out[i*6 + {0..5}] = k0 nk start end tmpl buf(of the first evaluation). */
int qlnlp_debug_segments(qlnlp_handle h, int64_t* out, int64_t cap, int64_t* nseg)
{
    if (int rc = check_handle(h)) return rc;
    if (nseg) *nseg = (int64_t)h->segs.size();
    if (out) {
        for (size_t i = 0; i < h->segs.size() && (int64_t)i < cap; ++i) {
            const QlSeg& s = h->segs[i];
            int64_t* o = out + 6 * i;
            o[0] = s.k0; o[1] = s.nk; o[2] = s.start; o[3] = s.end; o[4] = s.tmpl; o[5] = ql_seg_buffer((unsigned)i);
        }
    }
    return QLNLP_OK;
}

/* position of every VALS element inside a row of the handle's batch pattern; returns the count through *n */
int qlnlp_debug_vals_map(qlnlp_handle h, int32_t* pos, int64_t cap, int64_t* n)
{
    if (int rc = check_handle(h)) return rc;
    const QlClass& c = h->cls;
    std::vector<double> image;
    std::vector<int32_t> p;
    block_image_and_vals_map(c, image, p);
    if (batch_jm(h) == ql::JM_TRUE) {
        const std::vector<int32_t> t2b = true_to_block(c);
        std::vector<int32_t> b2t((size_t)c.nnz, -1);
        for (size_t t = 0; t < t2b.size(); ++t) b2t[(size_t)t2b[t]] = (int32_t)t;
        for (auto& q : p) q = b2t[(size_t)q];
    }
    if (n) *n = (int64_t)p.size();
    if (pos)
        for (size_t i = 0; i < p.size() && (int64_t)i < cap; ++i) pos[i] = p[i];
    return QLNLP_OK;
}

/* the host row builder alone (no device): rows of the batch pattern from VALS rows; threads <= 0: the handle's pool */
int qlnlp_debug_build_rows(qlnlp_handle h, const double* vals, int64_t ldv, double* jac, int64_t ldjac, int64_t rows,
                           int touched_only, int threads)
{
    if (int rc = check_handle(h)) return rc;
    qlnlp_handle s = first(h);
    if (int rc = ensure_plan(s)) return rc;
    if (threads == 1) { s->plan->build_rows(vals, ldv, jac, ldjac, 0, rows, touched_only != 0); return QLNLP_OK; }
    const bool async = threads < -1;           // negative: the workers-only path the host pipeline uses (Pool::start)
    if (async) threads = -threads;
    if (threads > 1) { s->opt_host_threads = threads; s->pool.reset(); }
    if (int rc = ensure_pool(s)) return rc;
    if (async) {
        s->plan->build_async(s->pool.get(), vals, ldv, jac, ldjac, rows, touched_only != 0);
        s->pool->finish();
    } else {
        s->plan->build(s->pool.get(), vals, ldv, jac, ldjac, rows, touched_only != 0);
    }
    return QLNLP_OK;
}

}  // extern "C"
