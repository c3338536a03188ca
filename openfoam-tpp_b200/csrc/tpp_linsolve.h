// p_rgh linear solver on the GPU (SURVEY.md §8a rows a11/a12;
// ref: circularSloshingTank/system/fvSolution:42-66).
//
// OpenFOAM's GAMG smoothers (DIC, DICGaussSeidel) are sequential face/cell sweeps with no
// parallel form.  The GPU path keeps OpenFOAM's *convergence contract* — the same L1
// residual normalised by lduMatrix::normFactor, the same tolerance / relTol / maxIter tests
// — and reaches it with a conjugate-gradient iteration preconditioned by an aggregation
// multigrid built on the device:
//   * agglomeration: pairwise handshake matching on the faceAreaPair weights (the weights
//     OpenFOAM's faceAreaPair agglomerator uses), two matching passes merged per level
//     (the equivalent of mergeLevels 2), built once and cached (the mesh moves rigidly);
//     aggregates never cross a rank boundary, processor faces are agglomerated into coarse
//     processor faces (OpenFOAM's GAMG processor-interface agglomeration), so every level
//     keeps its inter-rank couplings and its own halo exchange;
//   * Galerkin coarse operators re-summed every solve by segment gathers (no atomics);
//   * the V-cycle itself (tpp_vcycle.h): damped Jacobi, scaled corrections; the large levels
//     run one kernel per operation, the small ones ("tail": gathered onto every rank) run
//     inside ONE persistent cooperative kernel with grid barriers;
//   * SpMV is cell-gathered over the ELL table (fine level) / CSR rows (coarse levels);
//     dot products are two-stage with a fixed grid, so every sum has a fixed order.
#pragma once
#include "tpp_common.h"

#include <algorithm>
#include <numeric>

namespace tpp {

// One multigrid level in FP64 (Galerkin assembly, matching, the outer Krylov operator); the
// finest level aliases the solver's arrays.  Rows [0,n) are owned, columns [n,nOwn) are ghost
// rows behind processor faces (filled by the level's halo exchange).
struct LV {
    int n, nf, nCp, W, ell;       // rows, faces, ELL stride/width, ell=1: ELL, 0: CSR
    const int *cf, *cn, *rs;      // adjacency: (face<<1|side), other row; CSR row starts
    const int *own, *nei;         // [nf]
    double *diag, *upper, *rsum;  // matrix: diag, positive off-diagonal magnitude, row sums
    double *ev;                   // off-diagonal magnitudes per adjacency entry (ELL / CSR order)
    int nOwn;                     // valid columns: n + ghosts (n alone while matching: ghosts never pair)
    double nGlob;                 // global row count (normFactor's mean)
    // transfer from the next finer level
    const int *aggStart, *aggRows;   // CSR: members of each row of this level
    const int *segStart, *segFaces;  // CSR: fine faces summed into each face of this level
    const double *fupper, *frsum;    // the finer level's coefficients
    // kernel arguments
    const double *in, *b;
    double* out;
    double omega;
    // matching work
    int *match, *prop, *root;
    const double* fw;  // face weights
};

#define FOR_ROW(L, c)                                                     \
    {                                                                     \
        const int cnt_ = (L).ell ? (L).W : (L).rs[(c) + 1] - (L).rs[(c)]; \
        const size_t base_ = (L).ell ? (size_t)(c) : (size_t)(L).rs[(c)]; \
        const size_t str_ = (L).ell ? (size_t)(L).nCp : 1;                \
        for (int s_ = 0; s_ < cnt_; s_++) {                               \
            const int e_ = (L).cf[base_ + s_ * str_];                     \
            if (e_ < 0) break;                                            \
            const int f = e_ >> 1;                                        \
            const int o = (L).cn[base_ + s_ * str_];                      \
            if (o < 0 || o >= (L).nOwn) continue;
#define END_ROW }}

// y = A x on a level: out = diag*in - sum upper[f]*in[o]
// ELL row with a compile-time width: all index/coefficient loads are issued before the gathers
// (no early exit), which is what a latency-bound gather kernel needs; padded / boundary slots
// carry o < 0.  The sum runs in slot order, so it equals the generic loop bit for bit.
template <int W>
HD double ell_offdiag(const int* cn, const double* ev, size_t nCp, int c, const double* x) {
    int o[W];
    double v[W], xv[W];
#pragma unroll
    for (int k = 0; k < W; k++) { o[k] = cn[(size_t)k * nCp + c]; v[k] = ev[(size_t)k * nCp + c]; }
#pragma unroll
    for (int k = 0; k < W; k++) xv[k] = o[k] >= 0 ? x[o[k]] : 0.0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < W; k++) if (o[k] >= 0) s += v[k] * xv[k];
    return s;
}
// sum of offdiag magnitude * x[other] over the row (per-entry coefficients ev)
HD double row_offdiag(const int* cn, const double* ev, const int* rs, int ell, int W, size_t nCp, int c, const double* x) {
    if (ell) {
        if (W == 4) return ell_offdiag<4>(cn, ev, nCp, c, x);
        if (W == 5) return ell_offdiag<5>(cn, ev, nCp, c, x);
        if (W == 6) return ell_offdiag<6>(cn, ev, nCp, c, x);
        double s = 0;
        for (int k = 0; k < W; k++) { int o = cn[(size_t)k * nCp + c]; if (o >= 0) s += ev[(size_t)k * nCp + c] * x[o]; }
        return s;
    }
    double s = 0;
    for (int k = rs[c]; k < rs[c + 1]; k++) { int o = cn[k]; if (o >= 0) s += ev[k] * x[o]; }
    return s;
}
HD double row_Ax(const LV& L, int c, const double* x) {
    return L.diag[c] * x[c] - row_offdiag(L.cn, L.ev, L.rs, L.ell, L.W, (size_t)L.nCp, c, x);
}
// per-entry coefficients from the per-face ones (once per solve and level)
HD void b_fill_ev(const LV& L, int c) {
    const int cnt = L.ell ? L.W : L.rs[c + 1] - L.rs[c];
    const size_t base = L.ell ? (size_t)c : (size_t)L.rs[c];
    const size_t str = L.ell ? (size_t)L.nCp : 1;
    for (int k = 0; k < cnt; k++) {
        int e = L.cf[base + k * str], o = L.cn[base + k * str];
        L.ev[base + k * str] = (e >= 0 && o >= 0 && o < L.nOwn) ? L.upper[e >> 1] : 0.0;
    }
}
// first sweep from a zero guess: out = omega*b/diag  (also the diagonal preconditioner)
HD void b_jacobi0(const LV& L, int c) { L.out[c] = L.omega * L.b[c] / L.diag[c]; }
// rsum = diag - sum upper  (what is left of the row after the Laplacian part: boundary terms);
// couplings to ghost rows are part of the row
HD void b_rowsum(const LV& L, int c) {
    double s = 0;
    FOR_ROW(L, c) s += L.upper[f]; END_ROW
    L.rsum[c] = L.diag[c] - s;
}
// Galerkin: coarse face coefficient = sum of the fine faces in its segment (fixed order)
HD void b_coarse_upper(const LV& L, int F) {
    double s = 0;
    for (int k = L.segStart[F]; k < L.segStart[F + 1]; k++) s += L.fupper[L.segFaces[k]];
    L.upper[F] = s;
}
// coarse row sum = sum of member row sums ; coarse diag = row sum + sum of coarse off-diagonals
HD void b_coarse_diag(const LV& L, int I) {
    double r = 0;
    for (int k = L.aggStart[I]; k < L.aggStart[I + 1]; k++) r += L.frsum[L.aggRows[k]];
    L.rsum[I] = r;
    double s = 0;
    FOR_ROW(L, I) s += L.upper[f]; END_ROW
    L.diag[I] = r + s;
}
// ---- pairwise matching (handshake) on face weights --------------------------------------------
// Exact weight ties are the rule on extruded meshes; breaking them by cell index makes the
// handshake degenerate into chains (one pair per round).  A symmetric hash of the edge gives
// both ends the same pseudo-random order, so every locally dominant edge matches each round.
HD unsigned long long edge_hash(int a, int b) {
    unsigned long long x = ((unsigned long long)(unsigned)(a < b ? a : b) << 32) | (unsigned)(a < b ? b : a);
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
HD void b_match_propose(const LV& L, int c) {
    if (L.match[c] >= 0) { L.prop[c] = -1; return; }
    int best = -1;
    double bw = -1.0;
    unsigned long long bh = 0;
    FOR_ROW(L, c)
        if (L.match[o] < 0) {
            double w = L.fw[f];
            unsigned long long h = edge_hash(c, o);
            if (w > bw || (w == bw && h > bh)) { bw = w; best = o; bh = h; }
        }
    END_ROW
    L.prop[c] = best;
}
HD void b_match_accept(const LV& L, int c) {
    int p = L.prop[c];
    if (p >= 0 && L.prop[p] == c) L.match[c] = p;
}
// roots: pair -> min(c, partner); unmatched -> joins the aggregate of its heaviest matched
// neighbour, or stays single
HD void b_match_root(const LV& L, int c) {
    int m = L.match[c];
    if (m >= 0) { L.root[c] = c < m ? c : m; return; }
    int best = -1;
    double bw = -1.0;
    unsigned long long bh = 0;
    FOR_ROW(L, c)
        if (L.match[o] >= 0) {
            double w = L.fw[f];
            unsigned long long h = edge_hash(c, o);
            if (w > bw || (w == bw && h > bh)) { bw = w; best = o; bh = h; }
        }
    END_ROW
    if (best < 0) L.root[c] = c;
    else { int bm = L.match[best]; L.root[c] = best < bm ? best : bm; }
}

DEF_KERNEL(fill_ev, LV)
DEF_KERNEL(jacobi0, LV)
DEF_KERNEL(rowsum, LV)
DEF_KERNEL(coarse_upper, LV)
DEF_KERNEL(coarse_diag, LV)
DEF_KERNEL(match_propose, LV)
DEF_KERNEL(match_accept, LV)
DEF_KERNEL(match_root, LV)

// ---------------------------------------------------------------------------------------
// reductions and fused Krylov kernels (fixed grid -> fixed summation order)
// ---------------------------------------------------------------------------------------
// scal[] layout on the device
enum { S_WARA = 0, S_WARA_OLD, S_WAPA, S_RES, S_NORM, S_XSUM, S_TMP0, S_TMP1, S_MAX0, S_MAX1, S_DONE, S_ITERS, S_TOLA, S_TOLR, S_MAXIT, S_RESF, S_COUNT = 16 };

#ifndef TPP_EMU
DEV double block_sum(double v) {
    __shared__ double sh[BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wp] = v;
    __syncthreads();
    if (wp == 0) {
        v = lane < BLOCK / 32 ? sh[lane] : 0.0;
        for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    return v;  // valid in thread 0
}
DEV double block_max(double v) {
    __shared__ double shm[BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) shm[wp] = v;
    __syncthreads();
    if (wp == 0) {
        v = lane < BLOCK / 32 ? shm[lane] : 0.0;
        for (int o = 4; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    }
    return v;
}
// mode 0: sum a*b ; 1: sum |a| ; 2: sum a ; 3: max a (a >= 0)
__global__ void __launch_bounds__(256) k_reduce(const double* a, const double* b, int n, int mode, double* partial) {
    double v = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (mode == 0) v += a[i] * b[i];
        else if (mode == 1) v += fabs(a[i]);
        else if (mode == 2) v += a[i];
        else v = fmax(v, a[i]);
    }
    v = mode == 3 ? block_max(v) : block_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
}
__global__ void __launch_bounds__(256) k_reduce_final(const double* partial, int np, int mode, double* out) {
    double v = 0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) v = mode == 3 ? fmax(v, partial[i]) : v + partial[i];
    v = mode == 3 ? block_max(v) : block_sum(v);
    if (threadIdx.x == 0) *out = v;
}
// two final reductions (sums) in one launch: CTA 0 -> out0, CTA 1 -> out1
__global__ void __launch_bounds__(256) k_reduce_final2(const double* p0, const double* p1, int np, double* out0, double* out1) {
    const double* partial = blockIdx.x == 0 ? p0 : p1;
    double v = 0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) v += partial[i];
    v = block_sum(v);
    if (threadIdx.x == 0) *(blockIdx.x == 0 ? out0 : out1) = v;
}
// wA = A pA fused with partial sums of wA.pA
// (grid-stride over a fixed grid of RED_BLOCKS CTAs = every warp the SMs can hold; a partial per
// row-CTA instead was measured slower: the single-CTA final sum over 24 k partials costs more
// than the main kernel gains)
__global__ void __launch_bounds__(256) k_spmv_dot(LV L, double* partial) {
    double v = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < L.n; c += gridDim.x * blockDim.x) {
        double y = row_Ax(L, c, L.in);
        L.out[c] = y;
        v += y * L.in[c];
    }
    v = block_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
}
// the same with two adjacent rows per thread on the ELL mesh level (16-byte loads of the slot-major
// arrays: twice the bytes in flight per thread; every row's sum keeps its slot order)
template <int W>
DEV void ell2_Ax64(const LV& L, int c, const double* x, double& a0, double& a1) {
    int2 o[W];
    double2 v[W];
    const double2 dg = *reinterpret_cast<const double2*>(L.diag + c), xi = *reinterpret_cast<const double2*>(x + c);
#pragma unroll
    for (int k = 0; k < W; k++) o[k] = *reinterpret_cast<const int2*>(L.cn + (size_t)k * L.nCp + c);
#pragma unroll
    for (int k = 0; k < W; k++) v[k] = *reinterpret_cast<const double2*>(L.ev + (size_t)k * L.nCp + c);
    double xa[W], xb[W];
#pragma unroll
    for (int k = 0; k < W; k++) { xa[k] = o[k].x >= 0 ? x[o[k].x] : 0.0; xb[k] = o[k].y >= 0 ? x[o[k].y] : 0.0; }
    double s0 = 0, s1 = 0;
#pragma unroll
    for (int k = 0; k < W; k++) { if (o[k].x >= 0) s0 += v[k].x * xa[k]; if (o[k].y >= 0) s1 += v[k].y * xb[k]; }
    a0 = dg.x * xi.x - s0;
    a1 = dg.y * xi.y - s1;
}
template <int W>
__global__ void __launch_bounds__(256) k_spmv_dot_ell2(LV L, double* partial) {
    double v = 0;
    for (int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x); c < L.n; c += 2 * gridDim.x * blockDim.x) {
        if (c + 1 < L.n) {
            double a0, a1;
            ell2_Ax64<W>(L, c, L.in, a0, a1);
            double2 y; y.x = a0; y.y = a1;
            *reinterpret_cast<double2*>(L.out + c) = y;
            v += a0 * L.in[c];
            v += a1 * L.in[c + 1];
        } else {
            double y = row_Ax(L, c, L.in);
            L.out[c] = y;
            v += y * L.in[c];
        }
    }
    v = block_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
}
// x += alpha pA ; r -= alpha wA ; partial sums of |r|   (alpha = scal[WARA]/scal[WAPA])
__global__ void __launch_bounds__(256) k_update_xr(int n, double* x, double* r, const double* pA, const double* wA, const double* scal, double* partial) {
    if (scal[S_DONE] != 0.0) return;  // converged earlier in this chunk of iterations: x and r stay (the final residual is in S_RESF)
    double alpha = scal[S_WARA] / scal[S_WAPA];
    double v = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        x[c] += alpha * pA[c];
        double rr = r[c] - alpha * wA[c];
        r[c] = rr;
        v += fabs(rr);
    }
    v = block_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
}
// pA = z + beta pA  (beta = WARA/WARA_OLD; first iteration: pA = z)
__global__ void __launch_bounds__(256) k_update_p(int n, double* pA, const double* z, const double* scal) {
    const bool first = scal[S_WARA_OLD] == 0.0;
    double beta = first ? 0.0 : scal[S_WARA] / scal[S_WARA_OLD];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) pA[c] = first ? z[c] : z[c] + beta * pA[c];
}
__global__ void k_scal_set(double* scal, int dst, double v) { scal[dst] = v; }
// r = b - Ax ; normFactor pieces: |Ax - xbar*sumA| + |b - xbar*sumA| ; sumA = rsum
__global__ void __launch_bounds__(256) k_init_residual(LV L, const double* x, const double* b, double* r, const double* scal, double* partialRes, double* partialNorm) {
    double xbar = scal[S_XSUM] / L.nGlob;
    double v = 0, w = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < L.n; c += gridDim.x * blockDim.x) {
        double ax = row_Ax(L, c, x);
        double rr = b[c] - ax;
        r[c] = rr;
        v += fabs(rr);
        double p = L.rsum[c] * xbar;
        w += fabs(ax - p) + fabs(b[c] - p);
    }
    v = block_sum(v);
    __syncthreads();
    w = block_sum(w);
    if (threadIdx.x == 0) { partialRes[blockIdx.x] = v; partialNorm[blockIdx.x] = w; }
}
__global__ void k_scal_copy(double* scal, int dst, int src) { scal[dst] = scal[src]; }
__global__ void __launch_bounds__(256) k_add(int n, double* x, const double* z) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) x[c] += z[c];
}
// one row of the probes log: p at the probe cells (OpenFOAM's "not found" value outside the mesh)
__global__ void k_probe_row(const double* p, const int* cells, int n, double* row) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) row[k] = cells[k] >= 0 ? p[cells[k]] : -1.79769e+307;
}
// End of a PCG iteration, on the device: count it and apply OpenFOAM's stopping rule (residual
// below tolerance or relTol * initial, singular search direction, maxIter), so that a short solve
// can run several iterations per host round trip and still stop at exactly the same iteration.
// Thresholds are pre-multiplied by the normFactor: S_RES is the raw L1 norm.
__global__ void k_pcg_check(double* scal) {
    if (scal[S_DONE] != 0.0) return;
    const double it = scal[S_ITERS] + 1.0, res = scal[S_RES];
    scal[S_ITERS] = it;
    if (res < scal[S_TOLA] || res < scal[S_TOLR] || !(fabs(scal[S_WAPA]) >= scal[S_NORM] * VSMALL) || it >= scal[S_MAXIT] || !(res == res)) {
        scal[S_DONE] = 1.0;
        scal[S_RESF] = res;  // the residual the solve ended with: S_RES itself is scratch for the rest of the chunk
    }
}
__global__ void k_pcg_begin(double* scal, double tolA, double tolR, double maxIt) {
    scal[S_DONE] = 0.0; scal[S_ITERS] = 0.0; scal[S_TOLA] = tolA; scal[S_TOLR] = tolR; scal[S_MAXIT] = maxIt;
    scal[S_WARA] = 0.0;  // WARA_OLD == 0 marks the first iteration for k_update_p
}


// ---- halo exchange over peer memory (NVLink / NVSwitch), one kernel per exchange ---------------
// Every rank owns a window (cudaMalloc, exported with cudaIpc): per processor patch two data
// slots (sequence parity) and a flag.  The kernel
//   1. packs the owned values behind its processor faces straight into the NEIGHBOURS' windows
//      (peer stores), system fence, the last CTA then publishes the exchange's sequence number in
//      the neighbours' flags;
//   2. waits (bounded) until its own flags show that the neighbours' data of this exchange have
//      arrived, and copies them into the ghost rows.
// Step 1 never waits, so two ranks can never block each other; the parity slots are safe because
// a rank can only be one exchange ahead of its neighbour (it needs the neighbour's flag of
// exchange k to leave exchange k).  Replaces pack kernel + grouped ncclSend/ncclRecv (~16 us)
// by one ~4 us kernel; the sequence counter lives on the device so the kernel is graph-replayable.
constexpr int P2P_MAXPATCH = 8;
struct P2PArgs {
    int nG, nc, nPatch;
    int off[P2P_MAXPATCH];                     // ghost offset of every patch (ascending)
    char* peerData[P2P_MAXPATCH];              // my slot pair inside the neighbour's window
    unsigned long long* peerFlag[P2P_MAXPATCH];
    const char* myData[P2P_MAXPATCH];          // the neighbour's slot pair inside my window
    unsigned long long* myFlag[P2P_MAXPATCH];
    size_t slot[P2P_MAXPATCH];                 // bytes of one slot
    const int* owner;                          // owned row behind every processor face
    const void* src;
    void* ghost;
    unsigned long long* seq;                   // exchanges completed so far (device counter)
    unsigned* putDone;
    int* err;
};
DEV unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
DEV void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
DEV unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <class T>
__global__ void __launch_bounds__(256) k_halo_p2p(const P2PArgs a) {
    __shared__ int s_last;
    const unsigned long long seq = *(volatile unsigned long long*)a.seq + 1;
    const size_t par = (size_t)(seq & 1);
    const long total = (long)a.nG * a.nc;
    const T* src = (const T*)a.src;
    for (long e = blockIdx.x * 256L + threadIdx.x; e < total; e += gridDim.x * 256L) {
        int j = (int)(e / a.nc), k = (int)(e % a.nc), p = 0;
        while (p + 1 < a.nPatch && j >= a.off[p + 1]) p++;
        ((T*)(a.peerData[p] + par * a.slot[p]))[(size_t)(j - a.off[p]) * a.nc + k] = src[(size_t)a.owner[j] * a.nc + k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(a.putDone, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {  // every CTA has finished its stores (and has read *seq)
        if (threadIdx.x < a.nPatch) {
            __threadfence_system();
            st_release_sys(a.peerFlag[threadIdx.x], seq);
        }
        if (threadIdx.x == 0) { *a.putDone = 0; *(volatile unsigned long long*)a.seq = seq; }
    }
    if (threadIdx.x < a.nPatch) {
        const unsigned long long t0 = global_ns();
        long spins = 0;
        while (ld_acquire_sys(a.myFlag[threadIdx.x]) < seq) {
            if ((++spins & 255) == 0 && (*(volatile int*)a.err != 0 || global_ns() - t0 > 30000000000ull)) { *(volatile int*)a.err = 2; break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    T* ghost = (T*)a.ghost;
    for (long e = blockIdx.x * 256L + threadIdx.x; e < total; e += gridDim.x * 256L) {
        int j = (int)(e / a.nc), k = (int)(e % a.nc), p = 0;
        while (p + 1 < a.nPatch && j >= a.off[p + 1]) p++;
        ghost[e] = __ldcg(&((const T*)(a.myData[p] + par * a.slot[p]))[(size_t)(j - a.off[p]) * a.nc + k]);
    }
}

// ---- the same exchange with data and flag in one word ("LL" protocol) -----------------------------
// Every 32 data bits travel in an 8-byte store together with the exchange's sequence number, and an
// aligned 8-byte store is single-copy atomic: the receiver simply polls each word until the flag
// matches, so there is no system fence, no "last CTA publishes a flag" round and no ordering
// between words at all - the cost of an exchange is one kernel launch plus one NVLink traversal.
// (Twice the bytes on the wire; the messages are a few hundred KB, latency is what matters.)
DEV void st_ll(uint2* p, unsigned data, unsigned flag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(flag) : "memory");
}
DEV uint2 ld_ll(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
template <class T>
__global__ void __launch_bounds__(256) k_halo_ll(const P2PArgs a) {
    constexpr int WPV = sizeof(T) / 4;  // 32-bit words per value
    __shared__ int s_last;
    const unsigned long long seq = *(volatile unsigned long long*)a.seq + 1;
    const unsigned flag = (unsigned)seq;
    const size_t par = (size_t)(seq & 1);
    const long total = (long)a.nG * a.nc;
    const T* src = (const T*)a.src;
    for (long e = blockIdx.x * 256L + threadIdx.x; e < total; e += gridDim.x * 256L) {
        int j = (int)(e / a.nc), k = (int)(e % a.nc), p = 0;
        while (p + 1 < a.nPatch && j >= a.off[p + 1]) p++;
        union { T v; unsigned w[WPV]; } u;
        u.v = src[(size_t)a.owner[j] * a.nc + k];
        uint2* dst = reinterpret_cast<uint2*>(a.peerData[p] + par * a.slot[p]) + ((size_t)(j - a.off[p]) * a.nc + k) * WPV;
#pragma unroll
        for (int q = 0; q < WPV; q++) st_ll(dst + q, u.w[q], flag);
    }
    T* ghost = (T*)a.ghost;
    const unsigned long long t0 = global_ns();
    for (long e = blockIdx.x * 256L + threadIdx.x; e < total; e += gridDim.x * 256L) {
        int j = (int)(e / a.nc), k = (int)(e % a.nc), p = 0;
        while (p + 1 < a.nPatch && j >= a.off[p + 1]) p++;
        const uint2* from = reinterpret_cast<const uint2*>(a.myData[p] + par * a.slot[p]) + ((size_t)(j - a.off[p]) * a.nc + k) * WPV;
        union { T v; unsigned w[WPV]; } u;
#pragma unroll
        for (int q = 0; q < WPV; q++) {
            uint2 r = ld_ll(from + q);
            long spins = 0;
            while (r.y != flag) {
                if ((++spins & 1023) == 0 && (*(volatile int*)a.err != 0 || global_ns() - t0 > 30000000000ull)) { *(volatile int*)a.err = 2; break; }
                r = ld_ll(from + q);
            }
            u.w[q] = r.x;
        }
        ghost[e] = u.v;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.putDone, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) { *a.putDone = 0; *(volatile unsigned long long*)a.seq = seq; }
}

// ---- all-reduce of a few device scalars over the same peer windows (LL words, all-to-all) ---------
// The Krylov scalars and the multigrid scaling dots are 1-2 doubles, ~300 times per step: pure
// latency.  Every rank stores its values (LL words) into every rank's window and then sums what
// arrived in rank order - one small kernel, no fence, and bit-identical results on all ranks (the
// ranks must agree on every convergence decision).  Slots: [parity][source rank][value][word].
constexpr int AR_MAXR = 8, AR_MAXV = 4;
struct ARArgs {
    int rank, size, n, op;            // op 0: sum, 1: max
    uint2* win[AR_MAXR];              // the all-reduce region of every rank's window (win[rank] = mine)
    double* vals;                     // in/out: n device scalars
    unsigned long long* seq;
    int* err;
};
__global__ void __launch_bounds__(64) k_allreduce_ll(const ARArgs a) {
    __shared__ unsigned sw[AR_MAXR][AR_MAXV][2];
    const unsigned long long seq = *(volatile unsigned long long*)a.seq + 1;
    const unsigned flag = (unsigned)seq;
    const int par = (int)(seq & 1);
    const int t = threadIdx.x, r = t / (AR_MAXV * 2), i = (t / 2) % AR_MAXV, q = t % 2;
    const bool live = r < a.size && i < a.n;
    if (live) {
        union { double v; unsigned w[2]; } u;
        u.v = a.vals[i];
        st_ll(a.win[r] + ((par * AR_MAXR + a.rank) * AR_MAXV + i) * 2 + q, u.w[q], flag);
    }
    if (live) {
        const uint2* from = a.win[a.rank] + ((par * AR_MAXR + r) * AR_MAXV + i) * 2 + q;
        const unsigned long long t0 = global_ns();
        uint2 v = ld_ll(from);
        long spins = 0;
        while (v.y != flag) {
            if ((++spins & 1023) == 0 && (*(volatile int*)a.err != 0 || global_ns() - t0 > 30000000000ull)) { *(volatile int*)a.err = 3; break; }
            v = ld_ll(from);
        }
        sw[r][i][q] = v.x;
    }
    __syncthreads();
    if (t < a.n) {
        double acc = 0;
        for (int rr = 0; rr < a.size; rr++) {
            union { double v; unsigned w[2]; } u;
            u.w[0] = sw[rr][t][0]; u.w[1] = sw[rr][t][1];
            acc = rr == 0 ? u.v : (a.op == 0 ? acc + u.v : fmax(acc, u.v));
        }
        a.vals[t] = acc;
    }
    if (t == 0) *(volatile unsigned long long*)a.seq = seq;
}
#endif

// ---- alpha = iso surface statistics on the device (the reference's interface_summary.csv) -------
// main.py:761-780: cell data averaged to the points, contour at 0.5, max / min / mean z and the
// number of the contour's points - one point per mesh edge whose end values straddle the level.
struct IsoArgs {
    int nP, nE;
    const int *pcStart, *pcCells;  // point -> cells (CSR)
    const int *edgeA, *edgeB;      // mesh edges (point pairs)
    const double *alpha, *points0;
    double* ptAlpha;
    double iso;
    double R[9], T[3], cofg[3];    // rigid transform to the lab frame
    double* partial;               // [4][RED_BLOCKS]: sum z, count, max z, max(-z)
    double* out;                   // [4]: max z, min z, mean z, count
};
HD void b_iso_point(const IsoArgs& a, int p) {
    double s = 0;
    const int b = a.pcStart[p], e = a.pcStart[p + 1];
    for (int k = b; k < e; k++) s += a.alpha[a.pcCells[k]];
    a.ptAlpha[p] = e > b ? s / (double)(e - b) : 0.0;
}
HD bool iso_edge_z(const IsoArgs& a, int e, double& z) {
    const int pa = a.edgeA[e], pb = a.edgeB[e];
    const double va = a.ptAlpha[pa], vb = a.ptAlpha[pb];
    if ((va >= a.iso) == (vb >= a.iso)) return false;
    const double t = (a.iso - va) / (vb - va);
    double q[3];
    for (int k = 0; k < 3; k++) q[k] = (a.points0[3 * pa + k] + t * (a.points0[3 * pb + k] - a.points0[3 * pa + k])) - a.cofg[k];
    z = (a.R[6] * q[0] + a.R[7] * q[1] + a.R[8] * q[2]) + a.cofg[2] + a.T[2];
    return true;
}
#ifndef TPP_EMU
__global__ void __launch_bounds__(256) k_iso_point(const IsoArgs a) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < a.nP) b_iso_point(a, p);
}
__global__ void __launch_bounds__(256) k_iso_edges(const IsoArgs a) {
    double s = 0, n = 0, mx = -1e300, mn = -1e300;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < a.nE; e += gridDim.x * blockDim.x) {
        double z;
        if (iso_edge_z(a, e, z)) { s += z; n += 1.0; mx = fmax(mx, z); mn = fmax(mn, -z); }
    }
    // block_max starts from the values given (no clamp at 0): shift by using sums of indicator-free maxima
    __shared__ double sh[4][BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o); n += __shfl_down_sync(0xffffffffu, n, o);
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o)); mn = fmax(mn, __shfl_down_sync(0xffffffffu, mn, o));
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wp] = s; sh[1][wp] = n; sh[2][wp] = mx; sh[3][wp] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < BLOCK / 32; k++) { s += sh[0][k]; n += sh[1][k]; mx = fmax(mx, sh[2][k]); mn = fmax(mn, sh[3][k]); }
        a.partial[blockIdx.x] = s; a.partial[gridDim.x + blockIdx.x] = n; a.partial[2 * gridDim.x + blockIdx.x] = mx; a.partial[3 * gridDim.x + blockIdx.x] = mn;
    }
}
__global__ void k_iso_final(const IsoArgs a, int nb) {
    if (threadIdx.x != 0) return;
    double s = 0, n = 0, mx = -1e300, mn = -1e300;
    for (int k = 0; k < nb; k++) { s += a.partial[k]; n += a.partial[nb + k]; mx = fmax(mx, a.partial[2 * nb + k]); mn = fmax(mn, a.partial[3 * nb + k]); }
    a.out[0] = n > 0 ? mx : 0.0; a.out[1] = n > 0 ? -mn : 0.0; a.out[2] = n > 0 ? s / n : 0.0; a.out[3] = n;
}
#endif

#ifndef TPP_EMU
// ---- all-gather of the tail right-hand side over the peer windows ---------------------------------
// Once per V-cycle every rank contributes its slice of the gathered level's restricted residual
// (a few thousand values) and needs everybody else's: with NCCL an all-reduce of zero-padded vectors
// (~40 us at 8 ranks), here LL words stored straight into every rank's window and polled from the
// own one (slices are disjoint, so nothing is summed and every rank ends with identical bits).
// Slots: [parity][global row][word].
struct GatherArgs2 {
    int rank, size;
    int rowOff[AR_MAXR + 1];    // slice of every rank in the gathered vector
    uint2* win[AR_MAXR];        // the gather window of every rank (win[rank] = mine)
    void* vec;                  // in: my slice filled; out: all slices
    unsigned long long* seq;
    unsigned* done;
    int* err;
};
template <class T>
__global__ void __launch_bounds__(256) k_gather_ll(const GatherArgs2 a) {
    constexpr int WPV = sizeof(T) / 4;
    __shared__ int s_last;
    const unsigned long long seq = *(volatile unsigned long long*)a.seq + 1;
    const unsigned flag = (unsigned)seq;
    const size_t n = (size_t)a.rowOff[a.size];
    const size_t par = (size_t)(seq & 1) * n * WPV;
    T* v = (T*)a.vec;
    const int lo = a.rowOff[a.rank], hi = a.rowOff[a.rank + 1];
    // put: my rows into every other rank's window (one peer after the other: consecutive threads write
    // consecutive words of one destination)
    for (int r = 0; r < a.size; r++) {
        if (r == a.rank) continue;
        uint2* dst = a.win[r] + par;
        for (int e = lo + blockIdx.x * 256 + threadIdx.x; e < hi; e += gridDim.x * 256) {
            union { T x; unsigned w[WPV]; } u;
            u.x = v[e];
#pragma unroll
            for (int q = 0; q < WPV; q++) st_ll(dst + (size_t)e * WPV + q, u.w[q], flag);
        }
    }
    // get: everybody else's rows from my window
    const uint2* src = a.win[a.rank] + par;
    const unsigned long long t0 = global_ns();
    for (int e = blockIdx.x * 256 + threadIdx.x; e < (int)n; e += gridDim.x * 256) {
        if (e >= lo && e < hi) continue;
        union { T x; unsigned w[WPV]; } u;
#pragma unroll
        for (int q = 0; q < WPV; q++) {
            uint2 w = ld_ll(src + (size_t)e * WPV + q);
            long spins = 0;
            while (w.y != flag) {
                if ((++spins & 1023) == 0 && (*(volatile int*)a.err != 0 || global_ns() - t0 > 30000000000ull)) { *(volatile int*)a.err = 4; break; }
                w = ld_ll(src + (size_t)e * WPV + q);
            }
            u.w[q] = w.x;
        }
        v[e] = u.x;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) { *a.done = 0; *(volatile unsigned long long*)a.seq = seq; }
}
#endif

struct Reducer {
    double *partial = nullptr, *partial2 = nullptr;
    int cap = RED_BLOCKS;  // partials: at least one per 256 rows of the mesh (full-grid kernels)
    void init(int rows = 0) { cap = std::max(4 * RED_BLOCKS, (rows + BLOCK - 1) / BLOCK); partial = dalloc<double>(cap); partial2 = dalloc<double>(cap); }
    void free() { dev_free(partial); dev_free(partial2); }
    // mode as k_reduce
    void reduce(Ctx& ctx, const double* a, const double* b, int n, int mode, double* out) {
#ifdef TPP_EMU
        double v = 0;
        for (int i = 0; i < n; i++) {
            if (mode == 0) v += a[i] * b[i];
            else if (mode == 1) v += std::fabs(a[i]);
            else if (mode == 2) v += a[i];
            else v = std::max(v, a[i]);
        }
        *out = v;
        ctx.launches += 2;
#else
        prof_begin(ctx, "reduce");
        k_reduce<<<RED_BLOCKS, BLOCK, 0, ctx.stream>>>(a, b, n, mode, partial);
        k_reduce_final<<<1, BLOCK, 0, ctx.stream>>>(partial, RED_BLOCKS, mode, out);
        prof_end(ctx);
        ctx.launches += 2;
#endif
    }
};

}  // namespace tpp
