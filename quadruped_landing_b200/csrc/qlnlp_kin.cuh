// qlnlp_kin.cuh -- the OPT-IN kinematic (leg-length) rows.
//
// The reference carries them commented out (src/nlp.jl:60,70; src/constraints.jl:115-138, 276-288): cinds[8] = 2 rows per
// knot, d[2k-1] = norm(pb - p1), d[2k] = norm(pb - p2), bounds [0, l1 + l2 + lb/2].  A handle created with
// QLNLP_WITH_KINEMATICS switches them on: m_nlp grows by 2N and every Jacobian pattern by 8N entries (rows 2k-1 / 2k
// touch xb, yb and the foot's x, y), merged into the column-major value order.  The default path is untouched: the
// fused kernel evaluates the reference's rows as always (into a scratch row for the Jacobian values) and the two small
// kernels here append the kinematic rows of g and interleave their Jacobian entries.  This costs the opt-in path a
// second pass over the value stream; it is a correctness feature, not a tuned one.
#pragma once

#include <cuda_runtime.h>

#include "layout.h"

namespace ql {

// distance of the body to foot `foot` (0: p1 = x[3:4], 1: p2 = x[5:6]) and its direction
__device__ __forceinline__ double kin_norm(const double* x, int foot, double* dx, double* dy)
{
    *dx = __dsub_rn(x[0], x[foot ? 5 : 3]);
    *dy = __dsub_rn(x[1], x[foot ? 6 : 4]);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(*dx, *dx), __dmul_rn(*dy, *dy)));     // norm(d) = sqrt(dx*dx + dy*dy)
}

// g[b][c_kin + 2(k-1) + foot] for every evaluation b and knot k
__global__ void kin_g_kernel(const double* __restrict__ Z, long long ldz, double* __restrict__ g, long long ldg,
                             long long B, int N, int c_kin)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 2 * N) return;
    const long long b = t / (2 * N);
    const int r = (int)(t - b * 2 * N);
    const double* x = Z + b * ldz + (long long)QL_NZK * (r >> 1);
    double dx, dy;
    g[b * ldg + c_kin + r] = kin_norm(x, r & 1, &dx, &dy);
}

// Final Jacobian rows: out[b][o] = scratch[b][map[o]] for the reference's entries (map[o] >= 0) or the kinematic entry
// number -(map[o] + 1) = 8 (k-1) + e:  e = 4 foot + which,  which 0: d/dxb = dx/n, 1: d/dyb = dy/n, 2: d/dx_foot = -dx/n,
// 3: d/dy_foot = -dy/n   (constraints.jl:276-288 with the state indices the constraint itself uses)
__global__ void kin_expand_kernel(const int* __restrict__ map, int nnz_out, const double* __restrict__ scratch, long long lds,
                                  const double* __restrict__ Z, long long ldz, double* __restrict__ out, long long ldo, long long B)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * nnz_out) return;
    const long long b = t / nnz_out;
    const int o = (int)(t - b * nnz_out);
    const int m = __ldg(map + o);
    double v;
    if (m >= 0) {
        v = scratch[b * lds + m];
    } else {
        const int code = -(m + 1), k0 = code >> 3, e = code & 7;
        double dx, dy;
        const double n = kin_norm(Z + b * ldz + (long long)QL_NZK * k0, e >> 2, &dx, &dy);
        const double d = (e & 1) ? dy : dx;
        v = __ddiv_rn((e & 2) ? -d : d, n);
    }
    out[b * ldo + o] = v;
}

}  // namespace ql
