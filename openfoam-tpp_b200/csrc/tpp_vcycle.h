// The multigrid V-cycle used as the PCG preconditioner, templated on the value type R.
//
// The cycle is pure HBM traffic (Jacobi sweeps, residuals, transfers) and is applied ~40 times
// per time step, so it runs in FP32 by default: the outer PCG (operator, dot products,
// residual norms, convergence test) stays FP64 and reaches the same tolerance; the
// preconditioner only has to be a good approximate inverse.  R = double is kept (TPP_FP32=0).
// Sums inside a row are accumulated in R, global reductions (the scaling factor) in double.
#pragma once
#include "tpp_linsolve.h"

namespace tpp {

template <class R>
struct VL {
    int n, nf, nCp, W, ell, nOwn;
    const int *cn, *rs;  // adjacency: other row (ELL slot-major / CSR), CSR row starts
    const R *diag, *ev;  // matrix values (ev: off-diagonal magnitude per adjacency entry, 0 towards ghosts)
    // transfer (set on the level being restricted to / prolonged from)
    const int *agg, *aggStart, *aggRows;
    // kernel arguments
    const R *in, *b, *r, *c, *Ac, *xc;
    R* out;
    const double* sf;  // device scalars [num, den] of the correction scaling
    R omega;
};

// ---- row operators -------------------------------------------------------------------------
template <class R, int W>
HD R vl_ell_off(const VL<R>& L, int c, const R* x) {
    int o[W];
    R v[W], xv[W];
#pragma unroll
    for (int k = 0; k < W; k++) { o[k] = L.cn[(size_t)k * L.nCp + c]; v[k] = L.ev[(size_t)k * L.nCp + c]; }
#pragma unroll
    for (int k = 0; k < W; k++) xv[k] = (o[k] >= 0 && o[k] < L.nOwn) ? x[o[k]] : R(0);
    R s = 0;
#pragma unroll
    for (int k = 0; k < W; k++) s += v[k] * xv[k];
    return s;
}
template <class R>
HD R vl_off(const VL<R>& L, int c, const R* x) {
    if (L.ell) {
        if (L.W == 4) return vl_ell_off<R, 4>(L, c, x);
        if (L.W == 5) return vl_ell_off<R, 5>(L, c, x);
        if (L.W == 6) return vl_ell_off<R, 6>(L, c, x);
        R s = 0;
        for (int k = 0; k < L.W; k++) { int o = L.cn[(size_t)k * L.nCp + c]; if (o >= 0 && o < L.nOwn) s += L.ev[(size_t)k * L.nCp + c] * x[o]; }
        return s;
    }
    R s = 0;
    for (int k = L.rs[c]; k < L.rs[c + 1]; k++) { int o = L.cn[k]; if (o >= 0 && o < L.nOwn) s += L.ev[k] * x[o]; }
    return s;
}
template <class R> HD R vl_Ax(const VL<R>& L, int c, const R* x) { return L.diag[c] * x[c] - vl_off(L, c, x); }

template <class R> HD void vb_jacobi0(const VL<R>& L, int c) { L.out[c] = L.omega * L.b[c] / L.diag[c]; }
template <class R> HD void vb_jacobi(const VL<R>& L, int c) { L.out[c] = L.in[c] + L.omega * (L.b[c] - vl_Ax(L, c, L.in)) / L.diag[c]; }
template <class R> HD void vb_residual(const VL<R>& L, int c) { L.out[c] = L.b[c] - vl_Ax(L, c, L.in); }
// restriction: out[I] = sum of the fine residual r over the members of coarse row I (fixed order)
template <class R> HD void vb_restrict(const VL<R>& L, int I) {
    R s = 0;
    for (int k = L.aggStart[I]; k < L.aggStart[I + 1]; k++) s += L.r[L.aggRows[k]];
    L.out[I] = s;
}
// prolongation: out[i] = xc[agg[i]]  (fine rows)
template <class R> HD void vb_prolong(const VL<R>& L, int i) { L.out[i] = L.xc[L.agg[i]]; }
// GAMGSolver::scale: x += sf c + omega (r - sf A c)/diag, sf = (r.c)/(c.Ac) from the device scalars
template <class R> HD void vb_scale_apply(const VL<R>& L, int i) {
    double den = L.sf[1];
    R sf = (R)(L.sf[0] / (fabs(den) < VSMALL ? (den >= 0 ? VSMALL : -VSMALL) : den));
    L.out[i] += sf * L.c[i] + L.omega * (L.r[i] - sf * L.Ac[i]) / L.diag[i];
}
template <class R> struct CastArgs { const double* src; R* dst; const R* rsrc; double* ddst; };
template <class R> HD void vb_cast_in(const CastArgs<R>& a, int i) { a.dst[i] = (R)a.src[i]; }
template <class R> HD void vb_cast_out(const CastArgs<R>& a, int i) { a.ddst[i] = (double)a.rsrc[i]; }

#ifdef TPP_EMU
#define DEF_VKERNEL(name, VIEW) \
    template <class R> inline void vk_##name(const VIEW<R>& L, int n) { for (int i = 0; i < n; i++) vb_##name(L, i); }
#define VLAUNCH(ctx, name, view, n) do { if ((n) > 0) { prof_begin(ctx, "v_" #name); vk_##name(view, n); prof_end(ctx); (ctx).launches++; } } while (0)
#else
#define DEF_VKERNEL(name, VIEW)                                                          \
    template <class R> __global__ void __launch_bounds__(256) vk_##name(const VIEW<R> L, int n) { \
        int i = blockIdx.x * blockDim.x + threadIdx.x;                                   \
        if (i < n) vb_##name(L, i);                                                      \
    }
#define VLAUNCH(ctx, name, view, n) do { if ((n) > 0) { prof_begin(ctx, "v_" #name); vk_##name<<<((n) + 255) / 256, 256, 0, (ctx).stream>>>(view, n); prof_end(ctx); (ctx).launches++; } } while (0)
#endif

DEF_VKERNEL(jacobi0, VL)
DEF_VKERNEL(jacobi, VL)
DEF_VKERNEL(residual, VL)
DEF_VKERNEL(restrict, VL)
DEF_VKERNEL(prolong, VL)
DEF_VKERNEL(scale_apply, VL)
DEF_VKERNEL(cast_in, CastArgs)
DEF_VKERNEL(cast_out, CastArgs)

#ifndef TPP_EMU
// ---- CSR levels: COOP lanes per row ---------------------------------------------------------
template <class R, int COOP>
DEV R vl_coop_off(const VL<R>& L, int c, const R* x, int lane) {
    R s = 0;
    const int b = L.rs[c], e = L.rs[c + 1];
    for (int k = b + lane; k < e; k += COOP) {
        int o = L.cn[k];
        if (o >= 0 && o < L.nOwn) s += L.ev[k] * x[o];
    }
#pragma unroll
    for (int off = COOP / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    return s;
}
// mode 0: Jacobi sweep ; 1: residual
template <class R, int COOP>
__global__ void __launch_bounds__(256) vk_csr_row_op(const VL<R> L, int mode) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int c = gid / COOP, lane = gid % COOP;
    bool live = c < L.n;
    int cc = live ? c : L.n - 1;
    R off = vl_coop_off<R, COOP>(L, cc, L.in, lane);
    if (live && lane == 0) {
        R ax = L.diag[c] * L.in[c] - off;
        if (mode == 0) L.out[c] = L.in[c] + L.omega * (L.b[c] - ax) / L.diag[c];
        else L.out[c] = L.b[c] - ax;
    }
}
// out = A in fused with the partial sums of r.in and in.out (double accumulation)
template <class R>
__global__ void __launch_bounds__(256) vk_spmv_dot2(const VL<R> L, double* partialNum, double* partialDen) {
    double v = 0, w = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < L.n; c += gridDim.x * blockDim.x) {
        R x = L.in[c];
        R y = vl_Ax(L, c, L.in);
        L.out[c] = y;
        v += (double)L.r[c] * (double)x;
        w += (double)y * (double)x;
    }
    v = block_sum(v);
    __syncthreads();
    w = block_sum(w);
    if (threadIdx.x == 0) { partialNum[blockIdx.x] = v; partialDen[blockIdx.x] = w; }
}
template <class R, int COOP>
__global__ void __launch_bounds__(256) vk_csr_spmv_dot2(const VL<R> L, double* partialNum, double* partialDen) {
    double v = 0, w = 0;
    const int lane = threadIdx.x % COOP;
    const int sub = (threadIdx.x % 32) / COOP;
    const int rowsPerWarp = 32 / COOP;
    const int warpId = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const int nWarps = gridDim.x * blockDim.x / 32;
    for (int base = warpId * rowsPerWarp; base < L.n; base += nWarps * rowsPerWarp) {  // warp-uniform trips
        int i0 = base + sub;
        bool live = i0 < L.n;
        int i = live ? i0 : L.n - 1;
        R off = vl_coop_off<R, COOP>(L, i, L.in, lane);
        if (live && lane == 0) {
            R x = L.in[i];
            R y = L.diag[i] * x - off;
            L.out[i] = y;
            v += (double)L.r[i] * (double)x;
            w += (double)y * (double)x;
        }
    }
    v = block_sum(v);
    __syncthreads();
    w = block_sum(w);
    if (threadIdx.x == 0) { partialNum[blockIdx.x] = v; partialDen[blockIdx.x] = w; }
}
// Jacobi-preconditioned CG on the coarsest level, one CTA; vectors in R, reductions in double.
// out = x, b = rhs; scratch: r -> (R*)L.r, p -> (R*)L.c, Ap -> (R*)L.Ac (writable aliases)
template <class R>
__global__ void __launch_bounds__(1024) vk_coarse_cg(const VL<R> L, R* r, R* p, R* Ap, int maxIter, double relTol) {
    __shared__ double red[32];
    __shared__ double s_rz, s_rz0;
    const int n = L.n, t = threadIdx.x, T = blockDim.x;
    R* x = L.out;
    const R* b = L.b;
    auto bsum = [&](double v) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((t & 31) == 0) red[t >> 5] = v;
        __syncthreads();
        double s = 0;
        for (int k = 0; k < T / 32; k++) s += red[k];  // every thread sums the same 32 values in the same order
        return s;
    };
    double loc = 0;
    for (int i = t; i < n; i += T) { x[i] = 0; r[i] = b[i]; R z = b[i] / L.diag[i]; p[i] = z; loc += (double)b[i] * (double)z; }
    double rz = bsum(loc);
    if (t == 0) { s_rz = rz; s_rz0 = rz; }
    __syncthreads();
    if (!(rz > 0)) return;
    for (int it = 0; it < maxIter; it++) {
        loc = 0;
        for (int i = t; i < n; i += T) { R y = vl_Ax(L, i, (const R*)p); Ap[i] = y; loc += (double)y * (double)p[i]; }
        double pAp = bsum(loc);
        R alpha = (R)(s_rz / pAp);
        loc = 0;
        for (int i = t; i < n; i += T) { x[i] += alpha * p[i]; R rr = r[i] - alpha * Ap[i]; r[i] = rr; loc += (double)rr * (double)rr / (double)L.diag[i]; }
        double rzn = bsum(loc);
        if (rzn <= relTol * relTol * s_rz0) break;
        R beta = (R)(rzn / s_rz);
        __syncthreads();
        if (t == 0) s_rz = rzn;
        for (int i = t; i < n; i += T) p[i] = r[i] / L.diag[i] + beta * p[i];
        __syncthreads();
    }
}
#endif

}  // namespace tpp
