// Common definitions for the sm_100a solver: device view of the solver state, kernel
// definition / launch macros, deterministic reductions.
//
// Every kernel is a named __global__ function `k_<name>(DV d, int n)` whose body is the
// inline function `b_<name>(d, i)` — one thread per cell / face / boundary face, 256
// threads per CTA.  Nothing here is a dense contraction: the kernels are HBM-bound gathers
// over the owner/neighbour (LDU) addressing, so the design rules that matter are coalesced
// streaming of the per-face / per-cell arrays, cell-gathered (atomic-free, fixed-order)
// sums, and enough CTAs in flight to cover the 148 SMs.
//
// TPP_EMU (tests only): compiles the same bodies as plain host loops so the launch logic can
// be unit-tested on a machine without a GPU.  The product library never defines it.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef TPP_EMU
#define HD inline
#define DEV inline
#else
#include <cuda_runtime.h>
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#endif

namespace tpp {

constexpr double SMALL = 1e-15, VSMALL = 1e-300, ROOTVSMALL = 1e-150;
constexpr int BLOCK = 256;
constexpr int RED_BLOCKS = 148 * 8;  // fixed grid of the reducing kernels (2048 threads per SM), one partial per CTA -> deterministic sums

// Scalars that change every time step.  Kernels launched one by one get them by value inside DV; the
// graph-captured step (small meshes: launch overhead is the cost) reads them from this block, which
// the first node of the step's graph copies in from pinned host memory, so the same captured
// kernels serve every step.  Translation-only motion (the reference's orbital shaker).
struct StepScal {
    double dt, rdt[2];  // rdt[0] = 1/deltaT, rdt[1] = 1/(deltaT / nAlphaSubCycles)
    double dT[3], wallU[3], Tn[3];
};

// Device view: raw pointers + per-launch scalars, passed by value to every kernel.
struct DV {
    const StepScal* ss;  // nullptr: the by-value scalars below are current
    int rdtSel;          // which of ss->rdt this launch means by rDeltaT
    // sizes
    int nC, nF, nI, nB, nCp;  // nCp: padded cell count (ELL stride)
    int W;                    // ELL width (max faces per cell)
    // topology
    const int *own, *nei;      // [nF], [nI]
    const int *cf, *cn;        // ELL [W*nCp]: (face<<1)|isNeighbour or -1 ; other cell or -1
    const signed char *bcU, *bcA, *bcP;  // per boundary face
    const double *bInletAlpha, *bP0;     // per boundary face
    // geometry (current orientation)
    double *Sf, *magSf, *w, *dc, *corrVec, *dPN, *V, *gh, *ghf, *meshPhi;
    const double *Sf0, *dPN0, *corrVec0, *C0, *Cf0;  // body-frame references
    const double* points0;                            // undisplaced points (rotating motion only)
    const int *fOff, *fLab;                           // face -> point labels (rotating motion only)
    // fields
    double *alpha, *alpha0, *alpha_b, *U, *U_b, *U0, *U0_b, *p_rgh, *p_rgh_b, *pGrad_b, *p;
    double *rho, *rho_b, *rho0, *phi, *Uf, *Uf0, *alphaPhi, *rhoPhi;
    // alpha work
    double *grad, *phiBD, *phiCorr, *lambda, *alphaPhiUn, *sumPhip, *mSumPhim, *psiMaxn, *psiMinn, *lambdap, *lambdam;
    // momentum
    double *gradU, *mLower, *mUpper, *mExpl, *mDiag, *mSource, *mBIC, *mBBC;
    // pressure
    double *rAU, *HbyA, *HbyA_b, *rAUf, *phiHbyA, *phig, *pUpper, *pCorrFlux, *pDiag, *pSource, *rec;
    double *cellTmp;  // [2*nC] Courant work
    // halo packing (domain decomposition)
    const int* procOwner;
    const double* xsrc;
    double* xbuf;
    int xnc;
    // generic scalar-gradient arguments
    const double *gs, *gsb;
    double* gout;
    // per-launch scalars
    double dt, rDeltaT, subW, deltaN, cAlpha, rho1, rho2, nu1, nu2;
    double g[3];
    double dT[3];   // translation increment of this step
    double wallU[3];
    int moving, refCell, needRef;
    double pRefShift;
    double R[9], Rold[9], Tn[3], To[3], cofg[3];  // rigid transforms (new / old)
    int rotating;
    // surface tension (allocated only when sigma != 0; stf == nullptr otherwise)
    double *gradA, *nHatf, *sigmaK, *stf;
    double sigma;
};

HD double s_dt(const DV& d) { return d.ss ? d.ss->dt : d.dt; }
HD double s_rdt(const DV& d) { return d.ss ? d.ss->rdt[d.rdtSel] : d.rDeltaT; }
HD double s_dT(const DV& d, int k) { return d.ss ? d.ss->dT[k] : d.dT[k]; }
HD double s_wallU(const DV& d, int k) { return d.ss ? d.ss->wallU[k] : d.wallU[k]; }
HD double s_Tn(const DV& d, int k) { return d.ss ? d.ss->Tn[k] : d.Tn[k]; }
HD double sign_(double x) { return x >= 0 ? 1.0 : -1.0; }
HD double pos0_(double x) { return x >= 0 ? 1.0 : 0.0; }
HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
HD double mag3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
HD double dmin(double a, double b) { return a < b ? a : b; }
HD double dmax(double a, double b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------
// memory + launch plumbing
// ---------------------------------------------------------------------------------------
struct ProfRec {
    const char* name;
#ifndef TPP_EMU
    cudaEvent_t e0, e1;
#endif
};
struct Ctx {
    long launches = 0;
    bool prof = false;
    std::vector<ProfRec> recs;
#ifndef TPP_EMU
    cudaStream_t stream = nullptr;
    bool ownStream = false;
#endif
    std::string err;
};
// per-launch CUDA-event timing on the launching stream (bench.py's roofline numbers)
inline void prof_begin(Ctx& c, const char* name) {
    if (!c.prof) return;
    ProfRec r;
    r.name = name;
#ifndef TPP_EMU
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, c.stream);
#endif
    c.recs.push_back(r);
}
inline void prof_end(Ctx& c) {
    if (!c.prof) return;
#ifndef TPP_EMU
    cudaEventRecord(c.recs.back().e1, c.stream);
#endif
}

struct CudaFailure {  // thrown by CUDA_CHECK / LAUNCH_CHECK, caught at the C-ABI (API_CATCH)
    std::string what;
};
#ifdef TPP_EMU
inline void* dev_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 1); }
inline void dev_free(void* p) { free(p); }
inline void h2d(Ctx&, void* d, const void* h, size_t n) { memcpy(d, h, n); }
inline void d2h(Ctx&, void* h, const void* d, size_t n) { memcpy(h, d, n); }
inline void d2d(Ctx&, void* dst, const void* src, size_t n) { memcpy(dst, src, n); }
inline void dev_zero(Ctx&, void* d, size_t n) { memset(d, 0, n); }
inline void dev_sync(Ctx&) {}
#define DEF_KERNEL(name, VIEW) \
    inline void k_##name(const VIEW& d, int n) { for (int i = 0; i < n; i++) b_##name(d, i); }
#define LAUNCH(ctx, name, view, n) do { if ((n) > 0) { prof_begin(ctx, #name); k_##name(view, n); prof_end(ctx); (ctx).launches++; } } while (0)
// cell kernels templated on the ELL width WT (see FOR_CELL_FACES)
#define DEF_KERNEL_WB(name, minb) DEF_KERNEL_W(name)
#define DEF_KERNEL_W(name) \
    template <int WT> inline void k_##name(const DV& d, int n) { for (int i = 0; i < n; i++) b_##name<WT>(d, i); }
#define LAUNCH_W(ctx, name, view, n) do { if ((n) > 0) { prof_begin(ctx, #name); \
    switch ((view).W) { case 4: k_##name<4>(view, n); break; case 5: k_##name<5>(view, n); break; case 6: k_##name<6>(view, n); break; default: k_##name<0>(view, n); } \
    prof_end(ctx); (ctx).launches++; } } while (0)
#else
// A failed CUDA call becomes a C++ exception that every C-ABI entry point catches (API_GUARD in
// tppvof.cu) and turns into a negative return code + tpp_last_error(): no abort(), no exception
// across the ABI (include/tppvof.h).
[[noreturn]] inline void cuda_fail(cudaError_t e, const char* expr, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error '%s' in %s at %s:%d", cudaGetErrorString(e), expr, file, line);
    throw CudaFailure{buf};
}
#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) ::tpp::cuda_fail(e_, #x, __FILE__, __LINE__); } while (0)
// after a kernel launch: configuration errors (too many resources, bad grid) surface here and not
// at an unrelated later call.  Not a synchronisation.
#define LAUNCH_CHECK(name) do { cudaError_t e_ = cudaPeekAtLastError(); if (e_ != cudaSuccess) ::tpp::cuda_fail(e_, "launch of " name, __FILE__, __LINE__); } while (0)
// Zero-filled device memory.  The fill runs on a per-thread non-blocking utility stream and is
// waited for here: nothing touches the legacy default stream, which would order against every
// blocking stream of the process - illegal while another host thread (another handle of a sweep)
// is capturing a CUDA graph.
inline void* dev_alloc(size_t bytes) {
    void* p = nullptr;
    thread_local cudaStream_t util = nullptr;
    thread_local int utilDev = -1;
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    if (util == nullptr || utilDev != dev) { CUDA_CHECK(cudaStreamCreateWithFlags(&util, cudaStreamNonBlocking)); utilDev = dev; }
    CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 8));
    CUDA_CHECK(cudaMemsetAsync(p, 0, bytes ? bytes : 8, util));
    CUDA_CHECK(cudaStreamSynchronize(util));
    return p;
}
inline void dev_free(void* p) { if (p) cudaFree(p); }
inline void h2d(Ctx& c, void* d, const void* h, size_t n) { CUDA_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, c.stream)); CUDA_CHECK(cudaStreamSynchronize(c.stream)); }
inline void d2h(Ctx& c, void* h, const void* d, size_t n) { CUDA_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, c.stream)); CUDA_CHECK(cudaStreamSynchronize(c.stream)); }
inline void d2d(Ctx& c, void* dst, const void* src, size_t n) { CUDA_CHECK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, c.stream)); }
inline void dev_zero(Ctx& c, void* d, size_t n) { CUDA_CHECK(cudaMemsetAsync(d, 0, n, c.stream)); }
inline void dev_sync(Ctx& c) { CUDA_CHECK(cudaStreamSynchronize(c.stream)); }
#define DEF_KERNEL(name, VIEW)                                             \
    __global__ void __launch_bounds__(256) k_##name(const VIEW d, int n) { \
        int i = blockIdx.x * blockDim.x + threadIdx.x;                     \
        if (i < n) b_##name(d, i);                                         \
    }
#define LAUNCH(ctx, name, view, n) do { if ((n) > 0) { prof_begin(ctx, #name); k_##name<<<((n) + 255) / 256, 256, 0, (ctx).stream>>>(view, n); LAUNCH_CHECK(#name); prof_end(ctx); (ctx).launches++; } } while (0)
// cell kernels templated on the ELL width WT (see FOR_CELL_FACES)
#define DEF_KERNEL_W(name) DEF_KERNEL_WB(name, 4)
#define DEF_KERNEL_WB(name, minb)                                                       \
    template <int WT> __global__ void __launch_bounds__(256, (WT > 4 && (minb) > 2) ? 2 : (minb)) k_##name(const DV d, int n) { \
        int i = blockIdx.x * blockDim.x + threadIdx.x;                                  \
        if (i < n) b_##name<WT>(d, i);                                                  \
    }
#define LAUNCH_W(ctx, name, view, n) do { if ((n) > 0) { prof_begin(ctx, #name); const int g_ = ((n) + 255) / 256; \
    switch ((view).W) { case 4: k_##name<4><<<g_, 256, 0, (ctx).stream>>>(view, n); break; case 5: k_##name<5><<<g_, 256, 0, (ctx).stream>>>(view, n); break; \
                        case 6: k_##name<6><<<g_, 256, 0, (ctx).stream>>>(view, n); break; default: k_##name<0><<<g_, 256, 0, (ctx).stream>>>(view, n); } \
    LAUNCH_CHECK(#name); prof_end(ctx); (ctx).launches++; } } while (0)
#endif

template <class T>
T* dalloc(size_t n) { return (T*)dev_alloc(n * sizeof(T)); }

}  // namespace tpp
