// The multigrid V-cycle used as the PCG preconditioner, templated on the value type R.
//
// The cycle is pure HBM traffic (Jacobi sweeps, residuals, transfers) and is applied ~30 times
// per time step, so it runs in FP32 by default: the outer PCG (operator, dot products,
// residual norms, convergence test) stays FP64 and reaches the same tolerance; the
// preconditioner only has to be a good approximate inverse.  R = double is kept (TPP_FP32=0).
// Sums inside a row are accumulated in R, global reductions (the scaling factor) in double.
//
// Two regimes:
//   * large levels (bandwidth-bound): one kernel per operation (vk_*), rows distributed over
//     the ranks with one halo exchange per operator application;
//   * the tail (every level below TPP_TAIL_ROWS rows, gathered onto every rank): these levels
//     live in L2 and are launch/latency-bound, so the whole sub-cycle - smoothing, residual,
//     restriction, coarsest-level CG, prolongation, correction scaling - is ONE persistent
//     cooperative kernel (vk_tail) with grid barriers between the phases.
#pragma once
#include "tpp_linsolve.h"

namespace tpp {

template <class R>
struct VL {
    int n, nf, nCp, W, ell, nOwn;
    const int *cn, *rs;  // adjacency: other row (ELL slot-major / CSR), CSR row starts
    const R *diag, *ev;  // matrix values (ev: off-diagonal magnitude per adjacency entry)
    // coarse levels on the device: ELL part (the first ellW entries of every row, slot-major with
    // stride nPad: coalesced, no row-start indirection, all gathers of a row in flight at once)
    // plus the overflow entries of the few longer rows in CSR form
    int ellW, nPad;
    const int *ecn, *ors, *ocn;
    const R *eev, *oev;
    // transfer (set on the level being restricted to / prolonged from)
    const int *agg, *aggStart, *aggRows;
    // kernel arguments
    const R *in, *b, *r, *c, *Ac, *xc;
    R* out;
    const double* sf;  // device scalars [num, den] of the correction scaling
    R omega;
    // fused sweeps (vcol_*): first iterate om0 b/diag formed per column (mode 2); corrected iterate
    // in + oc xc[aggF] formed per column (mode 3)
    R om0, oc;
    const int* aggF;  // [n] coarse row of every row of THIS level
};

// column functors: the value of column j of the vector a row operator is applied to
template <class R> struct ColPlain {
    const R* x; int nOwn;
    HD R operator()(int j) const { return x[j]; }
};
template <class R> struct ColFirst {  // om0 b/diag, never stored; rows of other ranks (ghosts) count as zero
    const R *b, *d; R om0; int n;
    HD R operator()(int j) const { return j < n ? om0 * b[j] / d[j] : R(0); }
};
template <class R> struct ColCorr {  // in + oc xc[aggF] (levels without ghost rows only)
    const R *x, *xc; const int* aggF; R oc;
    HD R operator()(int j) const { return x[j] + oc * xc[aggF[j]]; }
};

// ---- row operators -------------------------------------------------------------------------
template <class R, int W, class G>
HD R vl_ell_off_g(const VL<R>& L, int c, const G& g) {
    int o[W];
    R v[W], xv[W];
#pragma unroll
    for (int k = 0; k < W; k++) { o[k] = L.cn[(size_t)k * L.nCp + c]; v[k] = L.ev[(size_t)k * L.nCp + c]; }
#pragma unroll
    for (int k = 0; k < W; k++) xv[k] = (o[k] >= 0 && o[k] < L.nOwn) ? g(o[k]) : R(0);
    R s = 0;
#pragma unroll
    for (int k = 0; k < W; k++) s += v[k] * xv[k];
    return s;
}
template <class R, class G>
HD R vl_off_g(const VL<R>& L, int c, const G& g) {
    if (L.ell) {
        if (L.W == 4) return vl_ell_off_g<R, 4>(L, c, g);
        if (L.W == 5) return vl_ell_off_g<R, 5>(L, c, g);
        if (L.W == 6) return vl_ell_off_g<R, 6>(L, c, g);
        R s = 0;
        for (int k = 0; k < L.W; k++) { int o = L.cn[(size_t)k * L.nCp + c]; if (o >= 0 && o < L.nOwn) s += L.ev[(size_t)k * L.nCp + c] * g(o); }
        return s;
    }
    R s = 0;
    for (int k = L.rs[c]; k < L.rs[c + 1]; k++) { int o = L.cn[k]; if (o >= 0 && o < L.nOwn) s += L.ev[k] * g(o); }
    return s;
}
template <class R> HD R vl_off(const VL<R>& L, int c, const R* x) { return vl_off_g(L, c, ColPlain<R>{x, L.nOwn}); }
template <class R> HD R vl_Ax(const VL<R>& L, int c, const R* x) { return L.diag[c] * x[c] - vl_off(L, c, x); }
// one Jacobi sweep on the vector g (value of row c: xc_), relaxation L.omega
template <class R, class G> HD R vl_sweep_g(const VL<R>& L, int c, const G& g) {
    const R xi = g(c);
    return xi + L.omega * (L.b[c] - (L.diag[c] * xi - vl_off_g(L, c, g))) / L.diag[c];
}

template <class R> HD void vb_jacobi0(const VL<R>& L, int c) { L.out[c] = L.omega * L.b[c] / L.diag[c]; }
template <class R> HD void vb_jacobi(const VL<R>& L, int c) { L.out[c] = L.in[c] + L.omega * (L.b[c] - vl_Ax(L, c, L.in)) / L.diag[c]; }
template <class R> HD void vb_residual(const VL<R>& L, int c) { L.out[c] = L.b[c] - vl_Ax(L, c, L.in); }
// the first two sweeps from a zero guess in one pass (the iterate om0 b/diag is never stored)
template <class R> HD void vb_jacobi_first(const VL<R>& L, int c) { L.out[c] = vl_sweep_g(L, c, ColFirst<R>{L.b, L.diag, L.om0, L.n}); }
// prolongation + over-correction + the first post-sweep in one pass
template <class R> HD void vb_jacobi_corr(const VL<R>& L, int c) { L.out[c] = vl_sweep_g(L, c, ColCorr<R>{L.in, L.xc, L.aggF, L.oc}); }
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
// correction with a fixed over-correction factor: out[i] += omega xc[agg[i]]
template <class R> HD void vb_prolong_add(const VL<R>& L, int i) { L.out[i] += L.omega * L.xc[L.agg[i]]; }
template <class R> struct CastArgs { const double* src; R* dst; const R* rsrc; double* ddst; };
template <class R> HD void vb_cast_in(const CastArgs<R>& a, int i) { a.dst[i] = (R)a.src[i]; }
template <class R> HD void vb_cast_out(const CastArgs<R>& a, int i) { a.ddst[i] = (double)a.rsrc[i]; }
// halo send buffer of a level: the owned row behind every processor face, in ghost order
template <class R> struct PackArgs { const int* owner; const R* src; R* dst; };
template <class R> HD void vb_pack(const PackArgs<R>& a, int j) { a.dst[j] = a.src[a.owner[j]]; }

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
#define VLAUNCH(ctx, name, view, n) do { if ((n) > 0) { prof_begin(ctx, "v_" #name); vk_##name<<<((n) + 255) / 256, 256, 0, (ctx).stream>>>(view, n); LAUNCH_CHECK("v_" #name); prof_end(ctx); (ctx).launches++; } } while (0)
#endif

DEF_VKERNEL(jacobi0, VL)
DEF_VKERNEL(jacobi, VL)
DEF_VKERNEL(residual, VL)
DEF_VKERNEL(jacobi_first, VL)
DEF_VKERNEL(jacobi_corr, VL)
DEF_VKERNEL(restrict, VL)
DEF_VKERNEL(prolong, VL)
DEF_VKERNEL(scale_apply, VL)
DEF_VKERNEL(prolong_add, VL)
DEF_VKERNEL(cast_in, CastArgs)
DEF_VKERNEL(cast_out, CastArgs)
DEF_VKERNEL(pack, PackArgs)

// ---- the tail: all small levels in one persistent kernel ---------------------------------------
constexpr int TAIL_MAXLV = 12;
constexpr int TAIL_MAXSW = 8;  // most sweeps per smoothing group in the tail
constexpr int TAIL_THREADS = 512;  // one CTA per SM: 128 registers per thread for the 16-wide ELL rows
template <class R>
struct TLv {
    int n, coop;                    // rows; lanes that walk one row together (4, 8 or 16)
    const int *rs, *cn;             // CSR (global numbering: the tail has no ghost rows)
    const R *ev, *diag;
    int ellW, nPad;                 // ELL + overflow form of the same rows (0: CSR only)
    const int *ecn, *ors, *ocn;
    const R *eev, *oev;
    const int* agg;                 // [n] row of the next (coarser) tail level; unused on the last
    const int *aggStart, *aggRows;  // members (rows of the previous tail level) of each row; unused on level 0
    R *x, *y, *b, *r;               // work vectors [n]; level 0: b = input, x = output
};
template <class R>
struct TailArgs {
    int T;  // tail levels
    TLv<R> lv[TAIL_MAXLV];
    unsigned* bar;    // grid-barrier counter, zeroed before every launch
    int* err;         // set if a barrier timed out (never in a healthy launch)
    R overcorr;                               // fixed over-correction factor of the prolonged correction
    R omPre[TAIL_MAXSW], omPost[TAIL_MAXSW];  // relaxation factor of every sweep (smootherOmega)
    int nPre, nPost, cgIter;
    double cgTol;
    int cgDeflate;        // remove the constant mode before the CG (closed domains: see tail_coarse_cg)
    R *cgR, *cgP, *cgAp;  // coarsest-level CG scratch (global memory; unused when the level is staged in shared memory)
    int cgSmem;           // > 0: the coarsest level is staged in shared memory in ELL form of this width
};
#ifndef TPP_EMU
struct GridBar {
    unsigned* ctr;
    unsigned nb, gen;
    int* err;
};
// all CTAs of the (cooperatively launched, hence co-resident) grid; bounded spin
DEV void gsync(GridBar& g) {
    __syncthreads();
    if (threadIdx.x == 0) {  // release on arrival, acquire on departure; bar.sync extends both to the CTA
        g.gen += g.nb;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(g.ctr) : "memory");
        unsigned v;
        long spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(g.ctr) : "memory");
            if (v >= g.gen) break;
            if ((++spins & 1023) == 0 && (*(volatile int*)g.err != 0 || spins > (1L << 24))) { *(volatile int*)g.err = 1; break; }
        } while (true);
    }
    __syncthreads();
}
// Row sweep of a tail level.  These levels sit in L2 and every phase is a chain of dependent
// loads (row start -> column/coefficient -> x[column]); with one CTA per SM the only way to hide
// that latency is memory-level parallelism, so every thread keeps TAIL_U rows in flight: `coop`
// lanes walk one row, rows base + u*rpp (u < TAIL_U) are handled together.  f(row, offdiag sum)
// runs on lane 0 of the row.  col = cn[k] or, through the aggregate map, map[cn[k]] (a
// prolonged coarse vector that is never stored).
constexpr int TAIL_U = 4;
template <class R, class G, class F>
DEV void tail_rows(const TLv<R>& L, G colval, int tid, int nth, F f) {
    // colval(j): the value of column j of the vector the operator is applied to - a stored vector,
    // or one that is never stored (the first Jacobi iterate omega0 b/diag; the iterate plus the
    // prolonged coarse correction): fusing those into the sweep that consumes them saves a phase
    // (a grid barrier) each
    if (L.ellW > 0) {
        // ELL + overflow: one thread per row, no row-start indirection; the (up to 16) column
        // indices and coefficients of the row are loaded together, then all x values
        const int W = L.ellW;
        for (int row = tid; row < L.n; row += nth) {
            int o[16];
            R v[16], xv[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const bool in = k < W;
                o[k] = in ? L.ecn[(size_t)k * L.nPad + row] : -1;
                v[k] = in ? L.eev[(size_t)k * L.nPad + row] : R(0);
            }
            const int ob = L.ors[row], oe = L.ors[row + 1];
#pragma unroll
            for (int k = 0; k < 16; k++) xv[k] = o[k] >= 0 ? colval(o[k]) : R(0);
            R s = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) s += v[k] * xv[k];
            for (int k = ob; k < oe; k++) s += L.oev[k] * colval(L.ocn[k]);
            f(row, s);
        }
        return;
    }
    const int coop = L.coop, lane = tid % coop, sub = tid / coop, rpp = nth / coop, n = L.n;
    for (int base = 0; base < n; base += TAIL_U * rpp) {
        int row[TAIL_U], k[TAIL_U], e[TAIL_U];
        R s[TAIL_U];
#pragma unroll
        for (int u = 0; u < TAIL_U; u++) {
            int r = base + u * rpp + sub;
            row[u] = r < n ? r : -1;
            int q = r < n ? r : n - 1;
            k[u] = L.rs[q] + lane;
            e[u] = r < n ? L.rs[q + 1] : 0;
            s[u] = 0;
        }
        bool any = true;
        while (any) {
            int o[TAIL_U];
            R v[TAIL_U], xv[TAIL_U];
#pragma unroll
            for (int u = 0; u < TAIL_U; u++) {
                bool in = k[u] < e[u];
                o[u] = in ? L.cn[k[u]] : -1;
                v[u] = in ? L.ev[k[u]] : R(0);
            }
#pragma unroll
            for (int u = 0; u < TAIL_U; u++) xv[u] = o[u] >= 0 ? colval(o[u]) : R(0);
            any = false;
#pragma unroll
            for (int u = 0; u < TAIL_U; u++) {
                s[u] += v[u] * xv[u];
                k[u] += coop;
                any = any || k[u] < e[u];
            }
            any = __any_sync(0xffffffffu, any);  // warp-uniform trip count (shuffles below)
        }
#pragma unroll
        for (int u = 0; u < TAIL_U; u++) {
            R t = s[u];
            for (int off = coop >> 1; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (row[u] >= 0 && lane == 0) f(row[u], t);
        }
    }
}
// result in every thread of the CTA; sh holds one double per warp
DEV double tail_block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = (threadIdx.x & 31) < (blockDim.x >> 5) ? sh[threadIdx.x & 31] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);  // same tree in every warp
    return s;
}
// Jacobi-preconditioned CG on the coarsest tail level by ONE CTA (zero initial guess).  The
// whole level - CSR with 16-bit columns, coefficients, diagonal, the CG vectors - is staged in
// shared memory when it fits (`useSmem`, decided on the host), so an iteration costs a few
// hundred cycles instead of a chain of L2 round trips.
template <class R>
DEV void tail_coarse_cg(const TLv<R>& L, R* gr, R* gp, R* gAp, int maxIter, double relTol, double* sh, unsigned char* smem, int ellW, int deflate) {
    // ellW > 0: the level is staged in shared memory in ELL form - the first ellW entries of every row
    // slot-major (entry s of row i at [s n + i]: the lanes of a warp, which hold consecutive rows, hit
    // consecutive banks; a CSR copy costs 3-4-way bank conflicts on every load because neighbouring rows
    // start ~15 words apart), 16-bit columns, padding = (0, the row itself); the few longer rows finish
    // from the CSR arrays in L2.  The CG vectors live in shared memory as well.  ellW == 0: CSR from L2.
    const int n = L.n, t = threadIdx.x, T = blockDim.x;
    const int* rs = L.rs;
    const R* dg = L.diag;
    R *x = L.x, *r = gr, *p = gp, *Ap = gAp;
    const R* sev = nullptr;
    const unsigned short* scn = nullptr;
    if (ellW > 0) {
        // layout: ev[ellW n] diag[n] x[n] r[n] p[n] Ap[n] (R) | cn[ellW n] (u16)
        R* wev = (R*)smem;
        R* sdg = wev + (size_t)ellW * n;
        R* sx = sdg + n; R* sr = sx + n; R* sp = sr + n; R* sAp = sp + n;
        unsigned short* wcn = (unsigned short*)(sAp + n);
        for (int i = t; i < n; i += T) {
            const int b0 = rs[i], e0 = rs[i + 1];
            for (int q = 0; q < ellW; q++) {
                const int k = b0 + q;
                const bool in = k < e0;
                wev[(size_t)q * n + i] = in ? L.ev[k] : R(0);
                wcn[(size_t)q * n + i] = (unsigned short)(in ? L.cn[k] : i);
            }
            sdg[i] = L.diag[i];
        }
        sev = wev; scn = wcn; dg = sdg; x = sx; r = sr; p = sp; Ap = sAp;
        __syncthreads();
    }
    // off-diagonal sum of row i applied to the vector v (shared or global)
    auto offsum = [&](int i, const R* v) {
        R s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int k = rs[i];
        const int e = rs[i + 1];
        if (sev) {
#pragma unroll 2
            for (int q = 0; q < ellW; q += 4) {  // ellW is a multiple of 4
                s0 += sev[(size_t)q * n + i] * v[scn[(size_t)q * n + i]];
                s1 += sev[(size_t)(q + 1) * n + i] * v[scn[(size_t)(q + 1) * n + i]];
                s2 += sev[(size_t)(q + 2) * n + i] * v[scn[(size_t)(q + 2) * n + i]];
                s3 += sev[(size_t)(q + 3) * n + i] * v[scn[(size_t)(q + 3) * n + i]];
            }
            for (k += ellW; k < e; k++) s0 += L.ev[k] * v[L.cn[k]];
        } else {
            for (; k + 3 < e; k += 4) { s0 += L.ev[k] * v[L.cn[k]]; s1 += L.ev[k + 1] * v[L.cn[k + 1]]; s2 += L.ev[k + 2] * v[L.cn[k + 2]]; s3 += L.ev[k + 3] * v[L.cn[k + 3]]; }
            for (; k < e; k++) s0 += L.ev[k] * v[L.cn[k]];
        }
        return (s0 + s1) + (s2 + s3);
    };
    const R* b = L.b;
    // Constant-mode deflation.  A closed tank (pressure level fixed by a reference cell only,
    // fvSolution:85-86) leaves the operator nearly singular: the constant vector has the energy of the
    // one doubled diagonal entry, and a handful of Jacobi-CG iterations does not see it - the outer PCG
    // then stalls (24 k-cell tutorial tank: 20 capped iterations at 1e-7 instead of 9-13).  The best
    // constant c = (1.b)/(1.A 1) is taken out first; with a Dirichlet boundary (the open tanks) 1.A 1 is
    // large and the step is skipped.
    R c0 = 0;
    if (deflate) {
        double sb = 0, sq = 0, sd = 0;
        for (int i = t; i < n; i += T) p[i] = R(1);
        __syncthreads();
        for (int i = t; i < n; i += T) {
            const R rsum = dg[i] - offsum(i, p);
            sb += (double)b[i]; sq += (double)rsum; sd += (double)dg[i];
            Ap[i] = rsum;  // row sums, used once below
        }
        sb = tail_block_sum(sb, sh); sq = tail_block_sum(sq, sh); sd = tail_block_sum(sd, sh);
        c0 = sq > 1e-12 * sd ? (R)(sb / sq) : R(0);
    }
    double loc = 0;
    for (int i = t; i < n; i += T) { x[i] = c0; R bi = deflate ? b[i] - c0 * Ap[i] : b[i]; r[i] = bi; R z = bi / dg[i]; p[i] = z; loc += (double)bi * (double)z; }
    double rz = tail_block_sum(loc, sh);
    const double rz0 = rz;
    if (rz > 0)
        for (int it = 0; it < maxIter; it++) {
            loc = 0;
            __syncthreads();
            for (int i = t; i < n; i += T) {
                R y = dg[i] * p[i] - offsum(i, p);
                Ap[i] = y;
                loc += (double)y * (double)p[i];
            }
            double pAp = tail_block_sum(loc, sh);
            R alpha = (R)(rz / pAp);
            loc = 0;
            for (int i = t; i < n; i += T) { x[i] += alpha * p[i]; R rr = r[i] - alpha * Ap[i]; r[i] = rr; loc += (double)(rr * rr / dg[i]); }
            double rzn = tail_block_sum(loc, sh);
            if (rzn <= relTol * relTol * rz0) break;
            R beta = (R)(rzn / rz);
            rz = rzn;
            for (int i = t; i < n; i += T) p[i] = r[i] / dg[i] + beta * p[i];
        }
    if (ellW > 0) {
        __syncthreads();
        for (int i = t; i < n; i += T) L.x[i] = x[i];
    }
}
// The V-cycle over the tail levels lv[0..T-1]: lv[0].b -> lv[0].x.  Same cycle as the per-kernel
// levels: nPre Chebyshev-Jacobi sweeps from a zero guess, residual, restriction; CG on the coarsest
// level; prolongation with the fixed over-correction factor, nPost sweeps.  A phase costs a grid
// barrier (the levels sit in L2: a phase is a few dependent L2 round trips plus the barrier), so the
// first iterate omega0 b/diag is never stored (the second sweep, or the residual, recomputes it per
// column) and the corrected iterate x + oc P xc is formed inside the first post-sweep: 3 + nPost
// phases per level for nPre <= 2.
template <class R>
__global__ void __launch_bounds__(TAIL_THREADS, 1) vk_tail(const TailArgs<R> A) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double sh[TAIL_THREADS / 32];
    GridBar gb{A.bar, gridDim.x, 0u, A.err};
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int nPre = A.nPre > 1 ? A.nPre : 1, nPost = A.nPost > 1 ? A.nPost : 1;
    for (int t = 0; t < A.T - 1; t++) {
        const TLv<R>& L = A.lv[t];
        // iterate after the pre-sweeps ends in L.x: sweeps 1..nPre-1 alternate buffers, the last writes x
        const R om0 = A.omPre[0];
        auto first = [&](int j) { return om0 * L.b[j] / L.diag[j]; };  // sweep 0, never stored
        const R* cur = nullptr;  // nullptr: the iterate is `first`
        for (int s = 1; s < nPre; s++) {
            R* out = ((nPre - 1 - s) & 1) ? L.y : L.x;
            const R om = A.omPre[s];
            if (cur == nullptr)
                tail_rows(L, first, tid, nth, [&](int i, R off) { R xi = first(i); out[i] = xi + om * (L.b[i] - (L.diag[i] * xi - off)) / L.diag[i]; });
            else {
                const R* in = cur;
                tail_rows(L, [&](int j) { return in[j]; }, tid, nth, [&](int i, R off) { out[i] = in[i] + om * (L.b[i] - (L.diag[i] * in[i] - off)) / L.diag[i]; });
            }
            cur = out;
            gsync(gb);
        }
        if (cur == nullptr) {  // nPre == 1: x = first, residual from the on-the-fly iterate
            tail_rows(L, first, tid, nth, [&](int i, R off) { R xi = first(i); L.x[i] = xi; L.r[i] = L.b[i] - (L.diag[i] * xi - off); });
        } else {
            const R* in = cur;
            tail_rows(L, [&](int j) { return in[j]; }, tid, nth, [&](int i, R off) { L.r[i] = L.b[i] - (L.diag[i] * in[i] - off); });
        }
        gsync(gb);
        const TLv<R>& C = A.lv[t + 1];
        for (int I = tid; I < C.n; I += nth) {
            R s = 0;
            for (int k = C.aggStart[I]; k < C.aggStart[I + 1]; k++) s += L.r[C.aggRows[k]];
            C.b[I] = s;
        }
        gsync(gb);
    }
    if (blockIdx.x == 0) tail_coarse_cg(A.lv[A.T - 1], A.cgR, A.cgP, A.cgAp, A.cgIter, A.cgTol, sh, smem, A.cgSmem, A.cgDeflate);
    gsync(gb);
    for (int t = A.T - 2; t >= 0; t--) {
        const TLv<R>& L = A.lv[t];
        const R* xc = A.lv[t + 1].x;
        const R oc = A.overcorr;
        // first post-sweep on the corrected iterate x + oc xc[agg] (formed per column), then the
        // remaining sweeps; the last one writes x
        const R* in = L.x;
        auto corrected = [&](int j) { return in[j] + oc * xc[L.agg[j]]; };
        for (int s = 0; s < nPost; s++) {
            R* out = ((nPost - 1 - s) & 1) ? L.y : L.x;
            const R om = A.omPost[s];
            if (s == 0) {
                // (in == L.x; out may alias it only when nPost is odd: then write through y and fix up below)
                R* o2 = out == L.x ? L.y : out;
                tail_rows(L, corrected, tid, nth, [&](int i, R off) { R xi = corrected(i); o2[i] = xi + om * (L.b[i] - (L.diag[i] * xi - off)) / L.diag[i]; });
                if (out == L.x) {  // nPost odd: copy back after the phase
                    gsync(gb);
                    for (int i = tid; i < L.n; i += nth) L.x[i] = L.y[i];
                }
                in = out;
            } else {
                const R* ii = in;
                tail_rows(L, [&](int j) { return ii[j]; }, tid, nth, [&](int i, R off) { out[i] = ii[i] + om * (L.b[i] - (L.diag[i] * ii[i] - off)) / L.diag[i]; });
                in = out;
            }
            if (s + 1 < nPost || t > 0) gsync(gb);
        }
    }
}

// ---- CSR levels: COOP lanes per row ---------------------------------------------------------
template <class R, int COOP, class G>
DEV R vl_coop_off_g(const VL<R>& L, int c, const G& g, int lane) {
    R s = 0;
    const int b = L.rs[c], e = L.rs[c + 1];
    for (int k = b + lane; k < e; k += COOP) {
        int o = L.cn[k];
        if (o >= 0 && o < L.nOwn) s += L.ev[k] * g(o);
    }
#pragma unroll
    for (int off = COOP / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    return s;
}
template <class R, int COOP> DEV R vl_coop_off(const VL<R>& L, int c, const R* x, int lane) { return vl_coop_off_g<R, COOP>(L, c, ColPlain<R>{x, L.nOwn}, lane); }
// mode 0: Jacobi sweep ; 1: residual ; 2: first two sweeps fused ; 3: correction + sweep fused
template <class R, int COOP>
__global__ void __launch_bounds__(256) vk_csr_row_op(const VL<R> L, int mode) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int c = gid / COOP, lane = gid % COOP;
    bool live = c < L.n;
    int cc = live ? c : L.n - 1;
    R off, xi;
    if (mode == 2) { ColFirst<R> g{L.b, L.diag, L.om0, L.n}; off = vl_coop_off_g<R, COOP>(L, cc, g, lane); xi = g(cc); }
    else if (mode == 3) { ColCorr<R> g{L.in, L.xc, L.aggF, L.oc}; off = vl_coop_off_g<R, COOP>(L, cc, g, lane); xi = g(cc); }
    else { off = vl_coop_off<R, COOP>(L, cc, L.in, lane); xi = L.in[cc]; }
    if (live && lane == 0) {
        R ax = L.diag[c] * xi - off;
        if (mode == 1) L.out[c] = L.b[c] - ax;
        else L.out[c] = xi + L.omega * (L.b[c] - ax) / L.diag[c];
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
// ---- mesh level (ELL), two adjacent rows per thread ------------------------------------------------
// The slot-major ELL arrays make rows c, c+1 adjacent in memory: one 8-byte (FP32) load serves
// both, so a thread has twice the bytes in flight with half the load instructions - the kernels
// are bound by how much the SM keeps outstanding, not by arithmetic.  Sums stay in slot order.
template <class R> struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<double> { typedef double2 type; };
template <class R, int W, class G>
DEV void vl_ell2_Ax_g(const VL<R>& L, int c, const G& x, R xi0, R xi1, R& ax0, R& ax1) {
    typedef typename Vec2<R>::type R2;
    int2 o[W];
    R2 v[W];
    const R2 dg = *reinterpret_cast<const R2*>(L.diag + c);
    R2 xi; xi.x = xi0; xi.y = xi1;
#pragma unroll
    for (int k = 0; k < W; k++) o[k] = *reinterpret_cast<const int2*>(L.cn + (size_t)k * L.nCp + c);
#pragma unroll
    for (int k = 0; k < W; k++) v[k] = *reinterpret_cast<const R2*>(L.ev + (size_t)k * L.nCp + c);
    R xa[W], xb[W];
#pragma unroll
    for (int k = 0; k < W; k++) {
        xa[k] = (o[k].x >= 0 && o[k].x < L.nOwn) ? x(o[k].x) : R(0);
        xb[k] = (o[k].y >= 0 && o[k].y < L.nOwn) ? x(o[k].y) : R(0);
    }
    R s0 = 0, s1 = 0;
#pragma unroll
    for (int k = 0; k < W; k++) { s0 += v[k].x * xa[k]; s1 += v[k].y * xb[k]; }
    ax0 = dg.x * xi.x - s0;
    ax1 = dg.y * xi.y - s1;
}
template <class R, int W>
DEV void vl_ell2_Ax(const VL<R>& L, int c, const R* x, R& ax0, R& ax1) {
    typedef typename Vec2<R>::type R2;
    const R2 xi = *reinterpret_cast<const R2*>(x + c);
    vl_ell2_Ax_g<R, W>(L, c, ColPlain<R>{x, L.nOwn}, xi.x, xi.y, ax0, ax1);
}
// MODE 0: Jacobi sweep ; 1: residual ; 2: first two sweeps from a zero guess fused (the iterate
// om0 b/diag is formed per column) ; 3: prolongation + over-correction + sweep fused
template <class R, int W, int MODE>
__global__ void __launch_bounds__(256) vk_ell2_row_op(const VL<R> L) {
    typedef typename Vec2<R>::type R2;
    const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= L.n) return;
    if (c + 1 < L.n) {
        R ax0, ax1;
        const R2 bb = *reinterpret_cast<const R2*>(L.b + c);
        const R2 dg = *reinterpret_cast<const R2*>(L.diag + c);
        R2 xi;
        if (MODE == 2) {
            xi.x = L.om0 * bb.x / dg.x; xi.y = L.om0 * bb.y / dg.y;
            vl_ell2_Ax_g<R, W>(L, c, ColFirst<R>{L.b, L.diag, L.om0, L.n}, xi.x, xi.y, ax0, ax1);
        } else if (MODE == 3) {
            const R2 xin = *reinterpret_cast<const R2*>(L.in + c);
            const int2 ag = *reinterpret_cast<const int2*>(L.aggF + c);
            xi.x = xin.x + L.oc * L.xc[ag.x]; xi.y = xin.y + L.oc * L.xc[ag.y];
            vl_ell2_Ax_g<R, W>(L, c, ColCorr<R>{L.in, L.xc, L.aggF, L.oc}, xi.x, xi.y, ax0, ax1);
        } else {
            xi = *reinterpret_cast<const R2*>(L.in + c);
            vl_ell2_Ax_g<R, W>(L, c, ColPlain<R>{L.in, L.nOwn}, xi.x, xi.y, ax0, ax1);
        }
        R2 res;
        if (MODE == 1) { res.x = bb.x - ax0; res.y = bb.y - ax1; }
        else {
            res.x = xi.x + L.omega * (bb.x - ax0) / dg.x;
            res.y = xi.y + L.omega * (bb.y - ax1) / dg.y;
        }
        *reinterpret_cast<R2*>(L.out + c) = res;
    } else {  // odd row count: the last row alone
        if (MODE == 1) L.out[c] = L.b[c] - vl_Ax(L, c, L.in);
        else if (MODE == 2) vb_jacobi_first(L, c);
        else if (MODE == 3) vb_jacobi_corr(L, c);
        else vb_jacobi(L, c);
    }
}
template <class R, int W>
__global__ void __launch_bounds__(256) vk_ell2_spmv_dot2(const VL<R> L, double* partialNum, double* partialDen) {
    typedef typename Vec2<R>::type R2;
    double v = 0, w = 0;
    for (int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x); c < L.n; c += 2 * gridDim.x * blockDim.x) {
        if (c + 1 < L.n) {
            R ax0, ax1;
            const R2 xi = *reinterpret_cast<const R2*>(L.in + c), rr = *reinterpret_cast<const R2*>(L.r + c);
            vl_ell2_Ax<R, W>(L, c, L.in, ax0, ax1);
            R2 y; y.x = ax0; y.y = ax1;
            *reinterpret_cast<R2*>(L.out + c) = y;
            v += (double)rr.x * (double)xi.x;
            v += (double)rr.y * (double)xi.y;
            w += (double)ax0 * (double)xi.x;
            w += (double)ax1 * (double)xi.y;
        } else {
            R x = L.in[c];
            R y = vl_Ax(L, c, L.in);
            L.out[c] = y;
            v += (double)L.r[c] * (double)x;
            w += (double)y * (double)x;
        }
    }
    v = block_sum(v);
    __syncthreads();
    w = block_sum(w);
    if (threadIdx.x == 0) { partialNum[blockIdx.x] = v; partialDen[blockIdx.x] = w; }
}

// ---- coarse levels in ELL + overflow form: one thread per row ------------------------------------
template <class R, int W, class G>
DEV R vl_ellc_off_g(const VL<R>& L, int c, const G& x) {
    int o[W];
    R v[W], xv[W];
#pragma unroll
    for (int k = 0; k < W; k++) { o[k] = L.ecn[(size_t)k * L.nPad + c]; v[k] = L.eev[(size_t)k * L.nPad + c]; }
#pragma unroll
    for (int k = 0; k < W; k++) xv[k] = o[k] >= 0 ? x(o[k]) : R(0);
    R s = 0;
#pragma unroll
    for (int k = 0; k < W; k++) s += v[k] * xv[k];
    for (int k = L.ors[c]; k < L.ors[c + 1]; k++) s += L.oev[k] * x(L.ocn[k]);
    return s;
}
template <class R, int W> DEV R vl_ellc_off(const VL<R>& L, int c, const R* x) { return vl_ellc_off_g<R, W>(L, c, ColPlain<R>{x, L.nOwn}); }
// mode 0: Jacobi sweep ; 1: residual ; 2: first two sweeps fused ; 3: correction + sweep fused
template <class R, int W>
__global__ void __launch_bounds__(256) vk_ellc_row_op(const VL<R> L, int mode) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= L.n) return;
    R xi, off;
    if (mode == 2) { ColFirst<R> g{L.b, L.diag, L.om0, L.n}; xi = g(c); off = vl_ellc_off_g<R, W>(L, c, g); }
    else if (mode == 3) { ColCorr<R> g{L.in, L.xc, L.aggF, L.oc}; xi = g(c); off = vl_ellc_off_g<R, W>(L, c, g); }
    else { xi = L.in[c]; off = vl_ellc_off<R, W>(L, c, L.in); }
    R ax = L.diag[c] * xi - off;
    if (mode == 1) L.out[c] = L.b[c] - ax;
    else L.out[c] = xi + L.omega * (L.b[c] - ax) / L.diag[c];
}
template <class R, int W>
__global__ void __launch_bounds__(256) vk_ellc_spmv_dot2(const VL<R> L, double* partialNum, double* partialDen) {
    double v = 0, w = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < L.n; c += gridDim.x * blockDim.x) {
        R x = L.in[c];
        R y = L.diag[c] * x - vl_ellc_off<R, W>(L, c, L.in);
        L.out[c] = y;
        v += (double)L.r[c] * (double)x;
        w += (double)y * (double)x;
    }
    v = block_sum(v);
    __syncthreads();
    w = block_sum(w);
    if (threadIdx.x == 0) { partialNum[blockIdx.x] = v; partialDen[blockIdx.x] = w; }
}
// ELL values from the CSR ones: dst[i] = src[idx[i]] (0 for padding)
template <class R> struct GatherArgs { const double* src; const int* idx; R* dst; };
template <class R>
__global__ void __launch_bounds__(256) vk_cast_gather(const GatherArgs<R> a, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { int k = a.idx[i]; a.dst[i] = k >= 0 ? (R)a.src[k] : R(0); }
}
// Jacobi-preconditioned CG on a whole (tiny) mesh without any coarse level, one CTA; vectors
// in R, reductions in double.  out = x, b = rhs
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
#else
// ---- host emulation of the tail (tests of the launch logic only) --------------------------------
template <class R> inline void tail_host_once(const TailArgs<R>& A);
// experiment knob (host emulation only): TPP_TAIL_CYCLES > 1 iterates the tail cycle on the
// residual of its first level, i.e. a nearly exact solve of that level
template <class R>
inline void tail_host(const TailArgs<R>& A) {
    const char* e = getenv("TPP_TAIL_CYCLES");
    const int nc = e ? atoi(e) : 1;
    if (nc <= 1 || A.T < 2) { tail_host_once(A); return; }
    const TLv<R>& L = A.lv[0];
    std::vector<R> b0(L.b, L.b + L.n), xs(L.n, R(0));
    for (int c = 0; c < nc; c++) {
        tail_host_once(A);
        for (int i = 0; i < L.n; i++) xs[i] += L.x[i];
        for (int i = 0; i < L.n; i++) {
            R s = 0;
            for (int k = L.rs[i]; k < L.rs[i + 1]; k++) s += L.ev[k] * xs[L.cn[k]];
            L.b[i] = b0[i] - (L.diag[i] * xs[i] - s);
        }
    }
    for (int i = 0; i < L.n; i++) { L.x[i] = xs[i]; L.b[i] = b0[i]; }
}
template <class R>
inline void tail_host_once(const TailArgs<R>& A) {
    // the same cycle as vk_tail, phase by phase, with the vectors it never stores materialised
    auto off = [](const TLv<R>& L, int row, const std::vector<R>& x) {
        R s = 0;
        for (int k = L.rs[row]; k < L.rs[row + 1]; k++) s += L.ev[k] * x[L.cn[k]];
        return s;
    };
    const int nPre = A.nPre > 1 ? A.nPre : 1, nPost = A.nPost > 1 ? A.nPost : 1;
    for (int t = 0; t < A.T - 1; t++) {
        const TLv<R>& L = A.lv[t];
        std::vector<R> cur(L.n), nxt(L.n);
        for (int i = 0; i < L.n; i++) cur[i] = A.omPre[0] * L.b[i] / L.diag[i];
        for (int s = 1; s < nPre; s++) {
            for (int i = 0; i < L.n; i++) nxt[i] = cur[i] + A.omPre[s] * (L.b[i] - (L.diag[i] * cur[i] - off(L, i, cur))) / L.diag[i];
            cur.swap(nxt);
        }
        for (int i = 0; i < L.n; i++) { L.x[i] = cur[i]; L.r[i] = L.b[i] - (L.diag[i] * cur[i] - off(L, i, cur)); }
        const TLv<R>& C = A.lv[t + 1];
        for (int I = 0; I < C.n; I++) {
            R s = 0;
            for (int k = C.aggStart[I]; k < C.aggStart[I + 1]; k++) s += L.r[C.aggRows[k]];
            C.b[I] = s;
        }
    }
    {
        const TLv<R>& L = A.lv[A.T - 1];
        const int n = L.n;
        R *x = L.x, *r = A.cgR, *p = A.cgP, *Ap = A.cgAp;
        std::vector<R> pv(n);
        // constant-mode deflation (see tail_coarse_cg)
        double sb = 0, sq = 0, sd = 0;
        std::vector<R> rsum(n);
        for (int i = 0; i < n; i++) {
            R o = 0;
            for (int k = L.rs[i]; k < L.rs[i + 1]; k++) o += L.ev[k];
            rsum[i] = L.diag[i] - o;
            sb += (double)L.b[i]; sq += (double)rsum[i]; sd += (double)L.diag[i];
        }
        const R c0 = (A.cgDeflate && sq > 1e-12 * sd) ? (R)(sb / sq) : R(0);
        double rz = 0;
        for (int i = 0; i < n; i++) { x[i] = c0; R ri = L.b[i] - c0 * rsum[i]; r[i] = ri; p[i] = ri / L.diag[i]; rz += (double)ri * (double)p[i]; }
        const double rz0 = rz;
        if (rz > 0)
            for (int it = 0; it < A.cgIter; it++) {
                double pAp = 0;
                pv.assign(p, p + n);
                for (int i = 0; i < n; i++) { Ap[i] = L.diag[i] * p[i] - off(L, i, pv); pAp += (double)Ap[i] * (double)p[i]; }
                R alpha = (R)(rz / pAp);
                double rzn = 0;
                for (int i = 0; i < n; i++) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; rzn += (double)r[i] * (double)r[i] / (double)L.diag[i]; }
                if (rzn <= A.cgTol * A.cgTol * rz0) break;
                R beta = (R)(rzn / rz);
                rz = rzn;
                for (int i = 0; i < n; i++) p[i] = r[i] / L.diag[i] + beta * p[i];
            }
    }
    for (int t = A.T - 2; t >= 0; t--) {
        const TLv<R>& L = A.lv[t];
        const R* xc = A.lv[t + 1].x;
        std::vector<R> cur(L.n), nxt(L.n);
        for (int i = 0; i < L.n; i++) cur[i] = L.x[i] + A.overcorr * xc[L.agg[i]];
        for (int s = 0; s < nPost; s++) {
            for (int i = 0; i < L.n; i++) nxt[i] = cur[i] + A.omPost[s] * (L.b[i] - (L.diag[i] * cur[i] - off(L, i, cur))) / L.diag[i];
            cur.swap(nxt);
        }
        for (int i = 0; i < L.n; i++) L.x[i] = cur[i];
    }
}
#endif

}  // namespace tpp
