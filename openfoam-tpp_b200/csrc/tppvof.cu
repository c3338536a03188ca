// libtppvof.so — host driver + C-ABI (include/tppvof.h) of the sm_100a incompressibleVoF step.
//
// Replaces, for one case, what `foamRun` does between reading the case and writing time
// directories (/root/reference/circularSloshingTank/Makefile:85,98; main.py:333-348).
// Layout in HBM: SoA arrays in OpenFOAM cell/face order (internal faces first, then the
// patches), vectors interleaved xyz; the cell->face ELL table drives every cell-gathered sum.
#include "../../include/tppvof.h"
#include "tpp_kernels.h"
#include "tpp_linsolve.h"
#include "tpp_vcycle.h"
#include "tpp_caseio.h"

#include <array>
#include <chrono>
#include <map>
#include <memory>
#include <tuple>
#ifndef TPP_EMU
#include <dlfcn.h>
#include <nccl.h>
#endif

using namespace tpp;

namespace {

// last error of the calling thread (handles may be driven from different host threads)
thread_local std::string g_err;
// every C-ABI entry point is a function-try-block ending in API_CATCH: nothing escapes as an
// exception or an abort(); the message is kept for tpp_last_error()
#ifdef TPP_EMU
#define API_DEVICE(s) do { if ((s) == nullptr) { g_err = "null handle"; return -1; } } while (0)
#else
// a handle may be driven from any host thread (one thread per concurrent sweep case): make its
// device current for the calling thread
#define API_DEVICE(s) do { if ((s) == nullptr) { g_err = "null handle"; return -1; } cudaSetDevice((s)->device); } while (0)
#endif
#ifdef TPP_EMU
#define CUDA_CHECK_OR_EMU(x) (void)0
#else
#define CUDA_CHECK_OR_EMU(x) CUDA_CHECK(x)
#endif
#define API_CATCH(code)                                                                         \
    catch (const tpp::CudaFailure& e) { g_err = e.what; fprintf(stderr, "tppvof: %s\n", g_err.c_str()); return (code); } \
    catch (const std::exception& e) { g_err = std::string("internal error: ") + e.what(); return (code); }               \
    catch (...) { g_err = "internal error"; return (code); }

struct Level {
    int n = 0, nf = 0, nfLoc = 0, nG = 0, nnz = 0;  // rows, faces (local + processor), ghost rows, CSR entries
    double nGlob = 0;
    // halo patches of this level (offset / count in ghost order, neighbour rank)
    std::vector<int> poff, pcnt, ppeer;
    // device
    int *cf = nullptr, *cn = nullptr, *rs = nullptr, *own = nullptr, *nei = nullptr, *dOwner = nullptr;
    int *agg = nullptr, *aggStart = nullptr, *aggRows = nullptr, *segStart = nullptr, *segFaces = nullptr;
    double *ev = nullptr, *diag = nullptr, *upper = nullptr, *rsum = nullptr;
    // ELL + overflow form of the CSR rows (distributed levels smoothed kernel by kernel)
    int ellW = 0, nPad = 0, nOv = 0;
    int *ecn = nullptr, *esrc = nullptr, *ors = nullptr, *ocn = nullptr, *osrc = nullptr;
    void free() {
        for (void* p : {(void*)ecn, (void*)esrc, (void*)ors, (void*)ocn, (void*)osrc}) dev_free(p);
        for (void* p : {(void*)cf, (void*)cn, (void*)rs, (void*)own, (void*)nei, (void*)dOwner, (void*)agg, (void*)aggStart, (void*)aggRows, (void*)segStart, (void*)segFaces,
                        (void*)ev, (void*)diag, (void*)upper, (void*)rsum})
            dev_free(p);
    }
};

struct SolveStats { int iters = 0; double r0 = 0, r = 0; };

// tuning knobs (environment overrides of the multigrid defaults; used by the tuning scripts)
inline int knob(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
inline double knobd(const char* name, double dflt) { const char* v = getenv(name); return v ? atof(v) : dflt; }
// Relaxation factor of sweep k of a group of m Jacobi sweeps.  TPP_CHEB=1 (default): the m sweeps
// together apply the degree-m Chebyshev polynomial of D^-1 A on [lmax/ratio, lmax] (Richardson
// form: omega_k = 1/root_k, no extra vector; fine for the m <= 4 used here).  lmax = 2 is the
// Gershgorin bound of these diagonally dominant M-matrices (every level: piecewise-constant
// Galerkin sums keep the sign pattern).  TPP_CHEB=0: plain damped Jacobi, TPP_OMEGA.
inline double smootherOmega(int k, int m) {
    static const int cheb = knob("TPP_CHEB", 1);
    static const double om = knobd("TPP_OMEGA", 0.8), lmax = knobd("TPP_CHEB_MAX", 2.0), ratio = knobd("TPP_CHEB_RATIO", 4.0);
    if (!cheb || m < 1) return om;
    const double a = lmax / ratio, theta = 0.5 * (lmax + a), delta = 0.5 * (lmax - a);
    return 1.0 / (theta - delta * cos(M_PI * (2.0 * k + 1.0) / (2.0 * m)));
}

// Inter-rank transport.  Product: NCCL send/recv + all-reduce on the solver's stream, resolved
// from the NCCL library the host process (torch) already loaded.  Tests / host emulation: two
// host callbacks (gloo in the CPU tests).
typedef int (*exchange_cb_t)(void* user, const double* send, double* recv, int ncomp);
typedef int (*allreduce_cb_t)(void* user, double* vals, int n, int op);
struct Comm {
    int rank = 0, size = 1;
    bool active = false;
    exchange_cb_t xcb = nullptr;
    allreduce_cb_t rcb = nullptr;
    void* user = nullptr;
    std::vector<double> hsend, hrecv;
#ifndef TPP_EMU
    void* lib = nullptr;
    ncclComm_t nccl = nullptr;
    ncclResult_t (*pGetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*pCommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*pCommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*pSend)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*pRecv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*pAllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*pGroupStart)() = nullptr;
    ncclResult_t (*pGroupEnd)() = nullptr;
    // halo exchange over peer memory (k_halo_p2p): my window, the neighbours' windows
    bool p2p = false, ll = false;  // ll: data + flag in one 8-byte word (k_halo_ll) instead of slots + flags (k_halo_p2p)
    char* window = nullptr;
    std::vector<void*> opened;               // cudaIpcOpenMemHandle results (closed at destroy)
    std::vector<char*> peerBase;             // per patch: base of the neighbour's window
    std::vector<size_t> myOff, peerOff, slot;  // per patch: data offset in my / the neighbour's window, slot bytes
    std::vector<int> peerFlagIdx;            // per patch: my flag index inside the neighbour's window
    unsigned long long* seq = nullptr;
    unsigned long long* arSeq = nullptr;     // all-reduce sequence counter
    std::vector<char*> allBase;              // every rank's window (all-reduce over peer memory), empty: NCCL
    unsigned* putDone = nullptr;
    int* p2pErr = nullptr;
    // all-gather of the tail right-hand side (k_gather_ll): a second window, sized when the tail is known
    char* gwin = nullptr;
    std::vector<char*> gBase;                // every rank's gather window, empty: NCCL all-reduce
    std::vector<int> gRowOff;                // slice of every rank in the gathered level
    unsigned long long* gSeq = nullptr;
    unsigned* gDone = nullptr;
#endif
};

}  // namespace

struct tpp_solver {
    Ctx ctx;
    int device = 0;
    caseio::Case* cs = nullptr;  // the case directory behind a handle made by tpp_open (owned)
    // host mesh
    int nP = 0, nF = 0, nI = 0, nC = 0, nB = 0, nPatch = 0, W = 0, nCp = 0;
    std::vector<double> points0;
    std::vector<int> fOff, fLab, own, nei;
    std::vector<int> pStart, pSize, bcU, bcA, bcP;
    std::vector<double> C0, Cf0, V, Sf0, magSf;
    tpp_config_t cfg;
    std::vector<double> motion;
    bool hasRotation = false;
    // device
    DV d;
    std::vector<void*> allocs;
    std::map<std::string, std::pair<double*, long>> reg;
    double* scal = nullptr;       // device scalars
    double* hscal = nullptr;      // pinned host mirror
    Reducer red;
    // multigrid
    std::vector<Level> levels;  // distributed coarse levels (levels[0] = first coarse); the last one is gathered
    std::vector<Level> tail;    // tail[0] = levels[gatherLevel] over all ranks, then the replicated coarser levels
    int gatherLevel = -1, tailRowOff = 0, tailFaceOff = 0, tailGrid = 1, tailCgW = 0;
    size_t tailSmem = 0;  // dynamic shared memory of vk_tail (coarsest level staged on chip), 0: not staged
    std::vector<std::array<int, 3>> tailCopy;  // processor-face coefficient ranges (src face, count, dst face) of the gather
    unsigned* tailBar = nullptr;
    int* tailErr = nullptr;
    double *kr = nullptr, *kz = nullptr, *kp = nullptr, *kw = nullptr, *fineEv = nullptr, *fineRsum = nullptr;
    int *match = nullptr, *prop = nullptr, *root = nullptr;
    bool amgBuilt = false;
    struct GraphKey {
        const void *x, *diag; int type, precond, nv;
        bool operator<(const GraphKey& o) const { return std::tie(x, diag, type, precond, nv) < std::tie(o.x, o.diag, o.type, o.precond, o.nv); }
    };
#ifndef TPP_EMU
    struct GraphRec { cudaGraphExec_t exec; long nodes; };
    std::map<GraphKey, GraphRec> graphs;
#endif
    // domain decomposition: processor-patch faces are renumbered as internal faces whose
    // neighbour is a ghost cell nC + j; ghost values arrive by halo exchange
    int nG = 0, nIloc = 0;
    long nGlobal = 0;
    std::vector<int> permDev2File;           // face renumbering (identity when nG == 0)
    std::vector<int> procOwner, procOff, procCnt, procPeer;
    int* dProcOwner = nullptr;
    double* sendbuf = nullptr;
    Comm comm;
    double* permBuf = nullptr;
    int* dPerm = nullptr;
    // Internal renumbering (single-rank meshes): cells in Morton order of their centres, internal faces
    // by (lower cell, higher cell).  Files, tpp_get / tpp_set / tpp_solve stay in OpenFOAM order - the
    // reference never renumbers its gmsh meshes (circularSloshingTank/Makefile:71-86) - and so does every
    // floating-point sum: face orientation is kept and a cell's ELL slots follow the FILE face index.
    bool renumbered = false;
    std::vector<int> cellFileOf, cellNewOf, faceFileOf, faceNewOf;  // new -> file and file -> new (faces: all nF, identity beyond nI)
    int *dCellFileOf = nullptr, *dFaceFileOf = nullptr;
    // (kind, components) of a registered array from its length: 'C' cells, 'F' all faces, 'I' internal faces
    std::pair<char, int> kindOf(long n) const {
        std::pair<char, int> k{0, 0};
        int hits = 0;
        auto tryK = [&](char c, long base) { for (int nc : {1, 3, 9}) if (base > 0 && n == nc * base && !(c != 'C' && nc == 9)) { k = {c, nc}; hits++; } };
        tryK('C', nC); tryK('F', nF); tryK('I', nI);
        if (hits != 1) k = {0, hits > 1 ? -1 : 0};  // ambiguous lengths: (0, -1)
        return k;
    }
    void ensurePermBuf() {
        if (permBuf) return;
        permBuf = A<double>(std::max(3 * (size_t)nF, 9 * (size_t)nC));
        dPerm = upload(permDev2File);
    }
    // device (internal) order <-> OpenFOAM file order of a registered array, on the device.
    // toFile: dst[file] = src[internal] ; else dst[internal] = src[file].  Returns false when the array
    // needs no permutation (boundary arrays, unknown lengths, no renumbering).
    bool permute(const double* src, double* dst, long n, bool toFile) {
        if (!renumbered) return false;
        auto k = kindOf(n);
        if (k.first == 0 || (k.first == 'F' && nG > 0)) return false;  // (all-face arrays of a processor mesh: permDev2File, tpp_get / tpp_set)
        const int* perm = k.first == 'C' ? dCellFileOf : dFaceFileOf;
        const int rows = k.first == 'C' ? nC : k.first == 'F' ? nF : nI;
        d.xsrc = src; d.xbuf = dst; d.xnc = k.second; d.procOwner = perm;
        if (toFile) LAUNCH(ctx, face_to_file, d, rows); else LAUNCH(ctx, file_to_face, d, rows);
        d.procOwner = dProcOwner;
        return true;
    }
    std::vector<double> hW, hDc, hCorr, hDPN;  // kept for the processor-face geometry pass
    // time
    double t = 0, dt = 0, dt0 = 0, startTime = 0, Co = 0, alphaCo = 0;
    long step = 0;
    int writeTimeIndex = 0;
    double Rn[9], Ro[9], Tn[3], To[3];
    SolveStats lastSolve[2];
    // probes
    std::vector<int> probeCells;
    std::vector<double> probeLog;
    double* probeDev = nullptr;
    int* probeIdx = nullptr;

    template <class T> T* A(size_t n) { T* p = dalloc<T>(n + 9 * (size_t)nG); allocs.push_back(p); return p; }  // room for ghost cells
    double* AD(const char* name, size_t n) { double* p = A<double>(n); reg[name] = {p, (long)n}; return p; }

    // ---- rigid motion (Function1s::Table linear + clamp; sixDoFMotion XYZ quaternion) --------
    void motionAt(double time, double R[9], double T[3]) const {
        double v[6] = {0, 0, 0, 0, 0, 0};
        int n = cfg.n_motion;
        if (n > 0) {
            const double* m = motion.data();
            if (time <= m[0]) for (int k = 0; k < 6; k++) v[k] = m[1 + k];
            else if (time >= m[7 * (n - 1)]) for (int k = 0; k < 6; k++) v[k] = m[7 * (n - 1) + 1 + k];
            else {
                int lo = 0, hi = n - 1;
                while (hi - lo > 1) { int mid = (lo + hi) / 2; if (m[7 * mid] <= time) lo = mid; else hi = mid; }
                double s = (time - m[7 * lo]) / (m[7 * hi] - m[7 * lo]);
                for (int k = 0; k < 6; k++) v[k] = m[7 * lo + 1 + k] + s * (m[7 * hi + 1 + k] - m[7 * lo + 1 + k]);
            }
        }
        const double d2r = M_PI / 180.0;
        double ax = v[3] * d2r, ay = v[4] * d2r, az = v[5] * d2r;
        double cx = cos(ax), sx = sin(ax), cy = cos(ay), sy = sin(ay), cz = cos(az), sz = sin(az);
        R[0] = cy * cz;                 R[1] = -cy * sz;                R[2] = sy;
        R[3] = sx * sy * cz + cx * sz;  R[4] = -sx * sy * sz + cx * cz; R[5] = -sx * cy;
        R[6] = -cx * sy * cz + sx * sz; R[7] = cx * sy * sz + sx * cz;  R[8] = cx * cy;
        T[0] = v[0]; T[1] = v[1]; T[2] = v[2];
    }

    // ---- host geometry at the undisplaced points (primitiveMesh conventions) -------------------
    void hostGeometry(std::vector<double>& w, std::vector<double>& dc, std::vector<double>& corr, std::vector<double>& dPN) {
        const double* P = points0.data();
        Cf0.assign(3 * nF, 0); Sf0.assign(3 * nF, 0); magSf.assign(nF, 0);
        for (int f = 0; f < nF; f++) {
            int s = fOff[f], n = fOff[f + 1] - s;
            const int* l = &fLab[s];
            double* cfp = &Cf0[3 * f]; double* sfp = &Sf0[3 * f];
            if (n == 3) {
                const double *a = P + 3 * l[0], *b = P + 3 * l[1], *c = P + 3 * l[2];
                double e1[3], e2[3];
                for (int k = 0; k < 3; k++) { cfp[k] = (1.0 / 3.0) * (a[k] + b[k] + c[k]); e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; }
                sfp[0] = 0.5 * (e1[1] * e2[2] - e1[2] * e2[1]);
                sfp[1] = 0.5 * (e1[2] * e2[0] - e1[0] * e2[2]);
                sfp[2] = 0.5 * (e1[0] * e2[1] - e1[1] * e2[0]);
            } else {
                double fc[3] = {0, 0, 0}, sumN[3] = {0, 0, 0}, sumA = 0, sumAc[3] = {0, 0, 0};
                for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) fc[k] += P[3 * l[i] + k];
                for (int k = 0; k < 3; k++) fc[k] /= n;
                for (int i = 0; i < n; i++) {
                    const double *a = P + 3 * l[i], *b = P + 3 * l[(i + 1) % n];
                    double c[3], e1[3], e2[3], nn[3];
                    for (int k = 0; k < 3; k++) { c[k] = a[k] + b[k] + fc[k]; e1[k] = b[k] - a[k]; e2[k] = fc[k] - a[k]; }
                    nn[0] = e1[1] * e2[2] - e1[2] * e2[1]; nn[1] = e1[2] * e2[0] - e1[0] * e2[2]; nn[2] = e1[0] * e2[1] - e1[1] * e2[0];
                    double an = sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
                    sumA += an;
                    for (int k = 0; k < 3; k++) { sumN[k] += nn[k]; sumAc[k] += an * c[k]; }
                }
                for (int k = 0; k < 3; k++) { cfp[k] = sumA < ROOTVSMALL ? fc[k] : (1.0 / 3.0) * sumAc[k] / sumA; sfp[k] = 0.5 * sumN[k]; }
            }
            magSf[f] = sqrt(sfp[0] * sfp[0] + sfp[1] * sfp[1] + sfp[2] * sfp[2]);
        }
        std::vector<double> cEst(3 * nC, 0.0);
        std::vector<int> nCF(nC, 0);
        for (int f = 0; f < nF; f++) { for (int k = 0; k < 3; k++) cEst[3 * own[f] + k] += Cf0[3 * f + k]; nCF[own[f]]++; }
        for (int f = 0; f < nI; f++) { if (nei[f] >= nC) continue; for (int k = 0; k < 3; k++) cEst[3 * nei[f] + k] += Cf0[3 * f + k]; nCF[nei[f]]++; }
        for (int c = 0; c < nC; c++) for (int k = 0; k < 3; k++) cEst[3 * c + k] /= nCF[c];
        C0.assign(3 * nC, 0.0); V.assign(nC, 0.0);
        for (int f = 0; f < nF; f++) {
            int o = own[f];
            double dd[3];
            for (int k = 0; k < 3; k++) dd[k] = Cf0[3 * f + k] - cEst[3 * o + k];
            double pyr = Sf0[3 * f] * dd[0] + Sf0[3 * f + 1] * dd[1] + Sf0[3 * f + 2] * dd[2];
            for (int k = 0; k < 3; k++) C0[3 * o + k] += pyr * (0.75 * Cf0[3 * f + k] + 0.25 * cEst[3 * o + k]);
            V[o] += pyr;
        }
        for (int f = 0; f < nI; f++) {
            int n = nei[f];
            if (n >= nC) continue;
            double dd[3];
            for (int k = 0; k < 3; k++) dd[k] = cEst[3 * n + k] - Cf0[3 * f + k];
            double pyr = Sf0[3 * f] * dd[0] + Sf0[3 * f + 1] * dd[1] + Sf0[3 * f + 2] * dd[2];
            for (int k = 0; k < 3; k++) C0[3 * n + k] += pyr * (0.75 * Cf0[3 * f + k] + 0.25 * cEst[3 * n + k]);
            V[n] += pyr;
        }
        for (int c = 0; c < nC; c++) {
            for (int k = 0; k < 3; k++) C0[3 * c + k] = fabs(V[c]) > VSMALL ? C0[3 * c + k] / V[c] : cEst[3 * c + k];
            V[c] *= (1.0 / 3.0);
        }
        w.assign(nF, 1.0); dc.assign(nF, 0.0); corr.assign(3 * (size_t)nI, 0.0); dPN.assign(3 * (size_t)nI, 0.0);
        for (int f = 0; f < nF; f++) {
            const double* S = &Sf0[3 * f];
            double nf[3] = {S[0] / magSf[f], S[1] / magSf[f], S[2] / magSf[f]};
            if (f < nI && nei[f] >= nC) {
                w[f] = 0.5; dc[f] = 0.0;  // processor face: finalised once the ghost centres are known
            } else if (f < nI) {
                double dO[3], dN[3], dd[3];
                for (int k = 0; k < 3; k++) {
                    dO[k] = Cf0[3 * f + k] - C0[3 * own[f] + k];
                    dN[k] = C0[3 * nei[f] + k] - Cf0[3 * f + k];
                    dd[k] = C0[3 * nei[f] + k] - C0[3 * own[f] + k];
                    dPN[3 * f + k] = dd[k];
                }
                double so = fabs(dot3(S, dO)), sn = fabs(dot3(S, dN));
                w[f] = sn / (so + sn);
                dc[f] = 1.0 / dmax(dot3(nf, dd), 0.05 * mag3(dd));
                for (int k = 0; k < 3; k++) corr[3 * f + k] = nf[k] - dd[k] * dc[f];
            } else {
                double dd[3];
                for (int k = 0; k < 3; k++) dd[k] = Cf0[3 * f + k] - C0[3 * own[f] + k];
                double nd = dot3(nf, dd);
                double dl[3] = {nf[0] * nd, nf[1] * nd, nf[2] * nd};
                dc[f] = 1.0 / dmax(dot3(nf, dl), 0.05 * mag3(dl));
            }
        }
    }

    template <class T> T* upload(const std::vector<T>& v) {
        T* p = A<T>(v.size());
        if (!v.empty()) h2d(ctx, p, v.data(), v.size() * sizeof(T));
        return p;
    }
    double* uploadD(const char* name, const std::vector<double>& v) {
        double* p = upload(v);
        reg[name] = {p, (long)v.size()};
        return p;
    }

    // Morton (Z-order) key of a point in the bounding box, 21 bits per axis
    static unsigned long long mortonKey(const double* x, const double* lo, const double* inv) {
        unsigned long long key = 0;
        unsigned q[3];
        for (int k = 0; k < 3; k++) { double t = (x[k] - lo[k]) * inv[k]; q[k] = (unsigned)std::min(2097151.0, std::max(0.0, t * 2097151.0)); }
        for (int b = 20; b >= 0; b--) for (int k = 0; k < 3; k++) key = (key << 1) | ((q[k] >> b) & 1u);
        return key;
    }
    // TPP_RENUMBER: 0 never, 1 always, default -1 = when it pays: the mean label distance |owner -
    // neighbour| over the internal faces (what decides whether a gather hits the same cache lines as
    // its neighbours) drops at least threefold.  A mesh written layer by layer keeps its order; a gmsh
    // mesh (no spatial coherence in its labels) is renumbered: 2.2 x faster steps at 6.2 M cells.
    void decideRenumbering(std::vector<double>& w, std::vector<double>& dc, std::vector<double>& corr, std::vector<double>& dPN) {
        const int mode = knob("TPP_RENUMBER", -1);
        const int nIl = nG > 0 ? nIloc : nI;  // faces between two cells of this rank (processor faces keep their place)
        if (mode == 0 || nC < 2 || nIl < 1) return;
        for (long n : {(long)nC, (long)nF, (long)nI})  // lengths must tell the array kinds apart (tpp_get / tpp_set)
            for (long m2 : {(long)nC, (long)nF, (long)nI, (long)nB})
                for (int a : {1, 3, 9}) for (int b : {1, 3, 9}) if (!(n == m2 && a == b) && (long)a * n == (long)b * m2 && n != m2) return;
        if (nC == nI || nC == nF || nB == nC || nB == nI) return;
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, inv[3];
        for (int c = 0; c < nC; c++) for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], C0[3 * c + k]); hi[k] = std::max(hi[k], C0[3 * c + k]); }
        double ext = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
        for (int k = 0; k < 3; k++) inv[k] = ext > 0 ? 1.0 / ext : 0.0;  // one scale for all axes: cubic Morton cells
        std::vector<std::pair<unsigned long long, int>> keyed(nC);
        for (int c = 0; c < nC; c++) keyed[c] = {mortonKey(&C0[3 * c], lo, inv), c};
        std::sort(keyed.begin(), keyed.end());
        std::vector<int> newOf(nC), fileOf(nC);
        for (int i = 0; i < nC; i++) { fileOf[i] = keyed[i].second; newOf[keyed[i].second] = i; }
        double before = 0, after = 0;
        for (int f = 0; f < nIl; f++) { before += std::abs(own[f] - nei[f]); after += std::abs(newOf[own[f]] - newOf[nei[f]]); }
        // every rank decides for itself: the labels are private to the rank (ghost rows and processor faces stay)
        if (mode < 0 && !(after * 3.0 <= before)) return;
        renumbered = true;
        cellFileOf = fileOf; cellNewOf = newOf;
        // local internal faces by (lower new cell, higher new cell); orientation (owner / neighbour roles) kept
        std::vector<std::pair<unsigned long long, int>> fk(nIl);
        for (int f = 0; f < nIl; f++) {
            unsigned a = (unsigned)newOf[own[f]], b = (unsigned)newOf[nei[f]];
            fk[f] = {((unsigned long long)std::min(a, b) << 32) | std::max(a, b), f};
        }
        std::sort(fk.begin(), fk.end());
        faceFileOf.resize(nF);
        for (int f = 0; f < nIl; f++) { faceFileOf[f] = fk[f].second; faceNewOf[fk[f].second] = f; }
        for (int f = nIl; f < nF; f++) faceFileOf[f] = f;
        auto permF = [&](std::vector<double>& a, int nc, int n) {  // face arrays: entry f <- old entry faceFileOf[f]
            std::vector<double> o(a.size());
            for (int f = 0; f < n; f++) for (int k = 0; k < nc; k++) o[(size_t)f * nc + k] = a[(size_t)faceFileOf[f] * nc + k];
            for (size_t i = (size_t)n * nc; i < a.size(); i++) o[i] = a[i];
            a.swap(o);
        };
        auto permC = [&](std::vector<double>& a, int nc) {
            std::vector<double> o(a.size());
            for (int c = 0; c < nC; c++) for (int k = 0; k < nc; k++) o[(size_t)c * nc + k] = a[(size_t)fileOf[c] * nc + k];
            a.swap(o);
        };
        permF(Cf0, 3, nIl); permF(Sf0, 3, nIl); permF(magSf, 1, nIl); permF(w, 1, nIl); permF(dc, 1, nIl); permF(corr, 3, nIl); permF(dPN, 3, nIl);
        permC(C0, 3); permC(V, 1);
        std::vector<int> own2(nF), nei2(nI), off2(nF + 1, 0), lab2;
        lab2.reserve(fLab.size());
        for (int f = 0; f < nF; f++) {
            const int ff = faceFileOf[f];
            own2[f] = newOf[own[ff]];
            if (f < nI) nei2[f] = nei[ff] < nC ? newOf[nei[ff]] : nei[ff];  // ghost rows keep their labels
            for (int k = fOff[ff]; k < fOff[ff + 1]; k++) lab2.push_back(fLab[k]);
            off2[f + 1] = (int)lab2.size();
        }
        own.swap(own2); nei.swap(nei2); fOff.swap(off2); fLab.swap(lab2);
        for (int& c : procOwner) c = newOf[c];
        if (nG > 0) {
            // face arrays cross the ABI through permDev2File (device face -> file face of the rank's
            // processorN mesh): compose it with the renumbering, and point faceFileOf at the file too
            std::vector<int> comp(nF);
            for (int f = 0; f < nF; f++) comp[f] = permDev2File[faceFileOf[f]];
            permDev2File.swap(comp);
        }
        if (knob("TPP_VERBOSE", 0)) fprintf(stderr, "tppvof: cells renumbered in Morton order (mean |owner - neighbour| %.0f -> %.0f)\n", before / nIl, after / nIl);
    }

    // What gmshToFoam / decomposePar guarantee about constant/polyMesh (Makefile:73,77) and the kernels rely
    // on: checked once, so a damaged mesh or configuration is an error code (include/tppvof.h), not a
    // stray index on the device.
    static bool validate(const tpp_mesh_t* m, const tpp_config_t* c) {
        char buf[256];
        auto bad = [&](const char* fmt, long a = 0, long b = 0, long c2 = 0) { snprintf(buf, sizeof buf, fmt, a, b, c2); g_err = std::string("invalid mesh / configuration: ") + buf; return false; };
        if (!m || !c) return bad("null mesh or configuration");
        if (m->n_points < 4 || m->n_cells < 1 || m->n_faces < 4 || m->n_internal < 0 || m->n_internal > m->n_faces || m->n_patches < 0)
            return bad("sizes (points %ld, faces %ld, cells %ld)", m->n_points, m->n_faces, m->n_cells);
        if (!m->points || !m->face_offsets || !m->face_labels || !m->owner || (m->n_internal > 0 && !m->neighbour)) return bad("null mesh array");
        if (m->n_patches > 0 && (!m->patch_start || !m->patch_size || !m->patch_bc_u || !m->patch_bc_alpha || !m->patch_bc_p || !m->patch_inlet_alpha || !m->patch_p0)) return bad("null patch array");
        for (long i = 0; i < 3L * m->n_points; i++) if (!std::isfinite(m->points[i])) return bad("point %ld is not finite", i / 3);
        if (m->face_offsets[0] != 0) return bad("face_offsets[0] != 0");
        for (int f = 0; f < m->n_faces; f++) {
            const int a = m->face_offsets[f], b = m->face_offsets[f + 1];
            if (b - a < 3) return bad("face %ld has %ld points", f, b - a);
            for (int k = a; k < b; k++) if (m->face_labels[k] < 0 || m->face_labels[k] >= m->n_points) return bad("face %ld: point label %ld out of range", f, m->face_labels[k]);
            if (m->owner[f] < 0 || m->owner[f] >= m->n_cells) return bad("face %ld: owner %ld out of range", f, m->owner[f]);
            if (f < m->n_internal && (m->neighbour[f] <= m->owner[f] || m->neighbour[f] >= m->n_cells))
                return bad("face %ld: neighbour %ld must lie in (owner %ld, n_cells) - upper-triangular order", f, m->neighbour[f], m->owner[f]);
        }
        {
            std::vector<int> nf(m->n_cells, 0);
            for (int f = 0; f < m->n_faces; f++) { nf[m->owner[f]]++; if (f < m->n_internal) nf[m->neighbour[f]]++; }
            for (int c2 = 0; c2 < m->n_cells; c2++) if (nf[c2] < 4) return bad("cell %ld has %ld faces (a closed cell needs 4)", c2, nf[c2]);
        }
        int next = m->n_internal;  // the patches tile [n_internal, n_faces) in order
        for (int p = 0; p < m->n_patches; p++) {
            if (m->patch_size[p] < 0 || m->patch_start[p] != next) return bad("patch %ld: start %ld, expected %ld", p, m->patch_start[p], next);
            next += m->patch_size[p];
        }
        if (next != m->n_faces) return bad("the patches cover faces up to %ld of %ld", next, m->n_faces);
        if (!(c->delta_t > 0) || !(c->end_time >= c->start_time) || !(c->max_co > 0) || !(c->max_alpha_co > 0) || !(c->max_delta_t > 0) || !(c->write_interval > 0))
            return bad("time control (deltaT, endTime, maxCo, maxAlphaCo, maxDeltaT and writeInterval must be positive)");
        if (c->n_alpha_subcycles < 1 || c->n_alpha_corr < 1 || c->n_limiter_iter < 0 || c->n_correctors < 1) return bad("nAlphaSubCycles / nAlphaCorr / nCorrectors must be >= 1");
        if (!(c->rho1 > 0) || !(c->rho2 > 0) || !(c->nu1 >= 0) || !(c->nu2 >= 0) || !(c->sigma >= 0)) return bad("rho must be positive, nu and sigma non-negative");
        if (c->n_motion < 0 || (c->n_motion > 0 && !c->motion)) return bad("motion table");
        for (const tpp_solver_t* sc : {&c->p_rgh, &c->p_rgh_final})
            if (!(sc->tolerance >= 0) || !(sc->rel_tol >= 0) || sc->max_iter < 1) return bad("p_rgh solver controls (tolerance, relTol >= 0, maxIter >= 1)");
        return true;
    }
    bool build(const tpp_mesh_t* m, const tpp_config_t* c) {
        if (!validate(m, c)) return false;
        nP = m->n_points; nF = m->n_faces; nI = m->n_internal; nC = m->n_cells; nB = nF - nI; nPatch = m->n_patches;
        points0.assign(m->points, m->points + 3 * (size_t)nP);
        fOff.assign(m->face_offsets, m->face_offsets + nF + 1);
        fLab.assign(m->face_labels, m->face_labels + fOff[nF]);
        own.assign(m->owner, m->owner + nF);
        nei.assign(m->neighbour, m->neighbour + nI);
        pStart.assign(m->patch_start, m->patch_start + nPatch);
        pSize.assign(m->patch_size, m->patch_size + nPatch);
        bcU.assign(m->patch_bc_u, m->patch_bc_u + nPatch);
        bcA.assign(m->patch_bc_alpha, m->patch_bc_alpha + nPatch);
        bcP.assign(m->patch_bc_p, m->patch_bc_p + nPatch);
        cfg = *c;
        if (c->n_motion > 0) motion.assign(c->motion, c->motion + 7 * (size_t)c->n_motion);
        cfg.motion = nullptr;
        for (int i = 0; i < cfg.n_motion; i++)
            for (int k = 4; k < 7; k++) if (motion[7 * i + k] != 0.0) hasRotation = true;
        // processor patches (BC codes -1) must follow the physical ones, as decomposePar writes them
        nIloc = nI;
        {
            bool seenProc = false;
            for (int p = 0; p < nPatch; p++) {
                bool isP = bcU[p] < 0 || bcA[p] < 0 || bcP[p] < 0;
                if (isP) {
                    seenProc = true;
                    if (pSize[p] == 0) continue;
                    procOff.push_back(nG); procCnt.push_back(pSize[p]);
                    procPeer.push_back(m->patch_neighb_proc ? m->patch_neighb_proc[p] : -1);
                    nG += pSize[p];
                } else if (seenProc && pSize[p] > 0) { g_err = "processor patches must come after the physical patches"; return false; }
            }
        }
        if (nG > 0) {
            const int nBphys = nB - nG;
            permDev2File.resize(nF);
            for (int f = 0; f < nI; f++) permDev2File[f] = f;
            for (int j = 0; j < nG; j++) permDev2File[nI + j] = nI + nBphys + j;
            for (int b = 0; b < nBphys; b++) permDev2File[nI + nG + b] = nI + b;
            std::vector<int> own2(nF), off2(nF + 1, 0), lab2;
            lab2.reserve(fLab.size());
            for (int fd = 0; fd < nF; fd++) {
                int ff = permDev2File[fd];
                own2[fd] = own[ff];
                for (int k = fOff[ff]; k < fOff[ff + 1]; k++) lab2.push_back(fLab[k]);
                off2[fd + 1] = (int)lab2.size();
            }
            own.swap(own2); fOff.swap(off2); fLab.swap(lab2);
            nei.resize(nI + nG);
            procOwner.resize(nG);
            for (int j = 0; j < nG; j++) { nei[nI + j] = nC + j; procOwner[j] = own[nI + j]; }
            nI += nG;
            nB -= nG;
        }
        // per boundary face BC tables
        std::vector<signed char> fU(nB, -1), fA(nB, -1), fP(nB, -1);
        std::vector<double> fInlet(nB, 0.0), fP0(nB, 0.0);
        for (int p = 0; p < nPatch; p++) {
            if (bcU[p] < 0 || bcA[p] < 0 || bcP[p] < 0) continue;
            for (int i = 0; i < pSize[p]; i++) {
                int b = pStart[p] - nIloc + i;
                if (b < 0 || b >= nB) { g_err = "patch range outside the boundary faces"; return false; }
                fU[b] = (signed char)bcU[p]; fA[b] = (signed char)bcA[p]; fP[b] = (signed char)bcP[p];
                fInlet[b] = m->patch_inlet_alpha[p]; fP0[b] = m->patch_p0[p];
            }
        }
        for (int b = 0; b < nB; b++) if (fU[b] < 0) { g_err = "boundary face without a patch"; return false; }
        // geometry in FILE order first: OpenFOAM's face loops accumulate the cell centres and volumes in
        // that order, and the oracle does the same (bit-exact V, C)
        std::vector<double> w, dc, corr, dPN;
        hostGeometry(w, dc, corr, dPN);
        faceNewOf.resize(nF);
        for (int f = 0; f < nF; f++) faceNewOf[f] = f;
        decideRenumbering(w, dc, corr, dPN);
        // ELL cell->face table.  Slots follow the FILE face index (a cell's neighbour-side and owner-side
        // faces interleave in face order), which is the order OpenFOAM's face loops scatter in; the entries
        // are internal face labels.
        std::vector<int> cnt(nC, 0);
        for (int f = 0; f < nF; f++) cnt[own[f]]++;
        for (int f = 0; f < nI; f++) if (nei[f] < nC) cnt[nei[f]]++;
        W = 0;
        for (int c = 0; c < nC; c++) W = std::max(W, cnt[c]);
        nCp = (nC + 31) / 32 * 32;
        std::vector<int> cf((size_t)W * nCp, -1), cn((size_t)W * nCp, -1), fill(nC, 0);
        for (int ff = 0; ff < nF; ff++) {
            const int f = faceNewOf[ff];
            int o = own[f];
            cf[(size_t)fill[o] * nCp + o] = f << 1;
            cn[(size_t)fill[o] * nCp + o] = f < nI ? nei[f] : -1;
            fill[o]++;
            if (f < nI && nei[f] < nC) {
                int n = nei[f];
                cf[(size_t)fill[n] * nCp + n] = (f << 1) | 1;
                cn[(size_t)fill[n] * nCp + n] = o;
                fill[n]++;
            }
        }
        memset(&d, 0, sizeof(d));
        d.nC = nC; d.nF = nF; d.nI = nI; d.nB = nB; d.nCp = nCp; d.W = W;
        d.own = upload(own); d.nei = upload(nei); d.cf = upload(cf); d.cn = upload(cn);
        if (renumbered) { dCellFileOf = upload(cellFileOf); dFaceFileOf = upload(faceFileOf); }
        d.bcU = upload(fU); d.bcA = upload(fA); d.bcP = upload(fP);
        d.bInletAlpha = upload(fInlet); d.bP0 = upload(fP0);
        d.Sf = uploadD("Sf", Sf0); d.magSf = uploadD("magSf", magSf); d.w = uploadD("w", w); d.dc = uploadD("dc", dc);
        d.corrVec = uploadD("corrVec", corr); d.dPN = uploadD("dPN", dPN); d.V = uploadD("V", V);
        d.C0 = uploadD("C0", C0); d.Cf0 = uploadD("Cf0", Cf0);
        if (hasRotation) {
            d.Sf0 = upload(Sf0); d.dPN0 = upload(dPN); d.corrVec0 = upload(corr);
            d.points0 = upload(points0); d.fOff = upload(fOff); d.fLab = upload(fLab);
        }
        d.gh = AD("gh", nC); d.ghf = AD("ghf", nF); d.meshPhi = AD("meshPhi", nF);
        d.alpha = AD("alpha", nC); d.alpha0 = AD("alpha0", nC); d.alpha_b = AD("alpha_b", nB);
        d.U = AD("U", 3 * (size_t)nC); d.U_b = AD("U_b", 3 * (size_t)nB); d.U0 = AD("U0", 3 * (size_t)nC); d.U0_b = AD("U0_b", 3 * (size_t)nB);
        d.p_rgh = AD("p_rgh", nC); d.p_rgh_b = AD("p_rgh_b", nB); d.pGrad_b = AD("pGrad_b", nB); d.p = AD("p", nC);
        d.rho = AD("rho", nC); d.rho_b = AD("rho_b", nB); d.rho0 = AD("rho0", nC);
        d.phi = AD("phi", nF); d.Uf = AD("Uf", 3 * (size_t)nF); d.Uf0 = AD("Uf0", 3 * (size_t)nF);
        d.alphaPhi = AD("alphaPhi", nF); d.rhoPhi = AD("rhoPhi", nF);
        d.grad = AD("grad", 3 * (size_t)nC); d.phiBD = AD("phiBD", nF); d.phiCorr = AD("phiCorr", nI); d.lambda = AD("lambda", nI);
        d.alphaPhiUn = AD("alphaPhiUn", nF); d.sumPhip = AD("sumPhip", nC); d.mSumPhim = AD("mSumPhim", nC);
        d.psiMaxn = AD("psiMaxn", nC); d.psiMinn = AD("psiMinn", nC); d.lambdap = AD("lambdap", nC); d.lambdam = AD("lambdam", nC);
        d.gradU = AD("gradU", 9 * (size_t)nC); d.mLower = AD("mLower", nI); d.mUpper = AD("mUpper", nI); d.mExpl = AD("mExpl", 3 * (size_t)nF);
        d.mDiag = AD("mDiag", nC); d.mSource = AD("mSource", 3 * (size_t)nC); d.mBIC = AD("mBIC", 3 * (size_t)nB); d.mBBC = AD("mBBC", 3 * (size_t)nB);
        d.rAU = AD("rAU", nC); d.HbyA = AD("HbyA", 3 * (size_t)nC); d.HbyA_b = AD("HbyA_b", 3 * (size_t)nB); d.rAUf = AD("rAUf", nF);
        d.phiHbyA = AD("phiHbyA", nF); d.phig = AD("phig", nF); d.pUpper = AD("pUpper", nI); d.pCorrFlux = AD("pCorrFlux", nI);
        d.pDiag = AD("pDiag", nC); d.pSource = AD("pSource", nC); d.rec = AD("rec", nF);
        if (cfg.sigma != 0.0) {  // surface tension work arrays (the reference's sigma 0 cases never pay for them)
            d.gradA = AD("gradA", 3 * (size_t)nC); d.nHatf = AD("nHatf", nF); d.sigmaK = AD("sigmaK", nC); d.stf = AD("stf", nF);
        }
        d.cellTmp = A<double>(2 * (size_t)nC);
        kr = A<double>(nC); kz = A<double>(nC); kp = A<double>(nC); kw = A<double>(nC); fineEv = A<double>((size_t)W * nCp);
        sendbuf = A<double>(9 * (size_t)std::max(nG, 1)); dProcOwner = upload(procOwner); d.procOwner = dProcOwner; nGlobal = nC; fineRsum = A<double>(nC);
        scal = A<double>(S_COUNT);
#ifdef TPP_EMU
        hscal = (double*)calloc(S_COUNT, sizeof(double));
#else
        CUDA_CHECK(cudaMallocHost(&hscal, S_COUNT * sizeof(double)));
#endif
        red.init(nC);
        d.sigma = cfg.sigma;
        d.cAlpha = cfg.c_alpha; d.rho1 = cfg.rho1; d.rho2 = cfg.rho2; d.nu1 = cfg.nu1; d.nu2 = cfg.nu2;
        for (int k = 0; k < 3; k++) { d.g[k] = cfg.g[k]; d.cofg[k] = cfg.cofg[k]; }
        d.moving = cfg.n_motion > 0; d.rotating = hasRotation;
        double vs = 0;
        for (double v : V) vs += v;
        d.deltaN = 1e-8 / cbrt(vs / nC);
        t = startTime = cfg.start_time; dt = dt0 = cfg.delta_t;
        motionAt(t, Rn, Tn);
        memcpy(Ro, Rn, sizeof(Rn)); memcpy(To, Tn, sizeof(Tn));
        setTransform();
        orientGeometry();
        // rho = rho2, alpha = 0 ...
        std::vector<double> r2(nC, cfg.rho2), r2b(nB, cfg.rho2);
        h2d(ctx, d.rho, r2.data(), nC * sizeof(double)); h2d(ctx, d.rho0, r2.data(), nC * sizeof(double));
        if (nB) h2d(ctx, d.rho_b, r2b.data(), nB * sizeof(double));
        // pressure reference
        d.needRef = 1;
        for (int p = 0; p < nPatch; p++) if (bcP[p] == TPP_P_TOTAL_PRESSURE && pSize[p] > 0) d.needRef = 0;
        d.refCell = -1;
        if (d.needRef && nG > 0) d.needRef = 0;  // decided over all ranks in finalizeParallel()
        else if (d.needRef) {
            d.refCell = findCell(cfg.p_ref_point);
            if (d.refCell < 0) { g_err = "pRefPoint is outside the mesh and p_rgh needs a reference"; return false; }
        }
        return true;
    }

    void setTransform() {
        memcpy(d.R, Rn, sizeof(Rn)); memcpy(d.Rold, Ro, sizeof(Ro));
        for (int k = 0; k < 3; k++) { d.Tn[k] = Tn[k]; d.To[k] = To[k]; d.dT[k] = Tn[k] - To[k]; d.wallU[k] = (Tn[k] - To[k]) / dt; }
    }
    void orientGeometry() {
        if (hasRotation) LAUNCH(ctx, rotate_face, d, nF);
        LAUNCH(ctx, gh_face, d, nF);
        LAUNCH(ctx, gh_cell, d, nC);
    }

    int findCell(const double* x) const {
        // on the mesh at its current rigid position: map the point back to the body frame
        double q[3] = {x[0] - cfg.cofg[0] - Tn[0], x[1] - cfg.cofg[1] - Tn[1], x[2] - cfg.cofg[2] - Tn[2]}, y[3];
        for (int k = 0; k < 3; k++) y[k] = Rn[k] * q[0] + Rn[3 + k] * q[1] + Rn[6 + k] * q[2] + cfg.cofg[k];  // R^T q
        std::vector<int> nbad(nC, 0);
        for (int f = 0; f < nF; f++) {
            double dd[3] = {y[0] - Cf0[3 * f], y[1] - Cf0[3 * f + 1], y[2] - Cf0[3 * f + 2]};
            double s = dot3(dd, &Sf0[3 * f]);
            double tol = 1e-12 * magSf[f] * sqrt(magSf[f]);
            if (s > tol) nbad[own[f]]++;
            if (f < nI && nei[f] < nC && s < -tol) nbad[nei[f]]++;  // (a processor face's neighbour is a ghost row)
        }
        int best = -1;
        double bd = 1e300;
        for (int c = 0; c < nC; c++)
            if (nbad[c] == 0) {
                double dd[3] = {y[0] - C0[3 * c], y[1] - C0[3 * c + 1], y[2] - C0[3 * c + 2]};
                double q2 = dot3(dd, dd);
                if (q2 < bd) { bd = q2; best = c; }
            }
        return best;
    }

    // ---- halo exchange: owner-side values of my processor faces -> the neighbour's ghost cells
    void X(double* field, int nc) {
        if (!comm.active || nG == 0) return;
#ifndef TPP_EMU
        if (comm.p2p) { p2pExchange<double>(dProcOwner, procOff, nG, field, field + (size_t)nC * nc, nc); return; }
#endif
        d.xsrc = field; d.xbuf = sendbuf; d.xnc = nc;
        LAUNCH(ctx, pack_halo, d, nG);
        double* ghost = field + (size_t)nC * nc;
#ifndef TPP_EMU
        if (comm.nccl) {
            prof_begin(ctx, "halo_sendrecv");
            comm.pGroupStart();
            for (size_t i = 0; i < procCnt.size(); i++) {
                comm.pSend(sendbuf + (size_t)procOff[i] * nc, (size_t)procCnt[i] * nc, ncclDouble, procPeer[i], comm.nccl, ctx.stream);
                comm.pRecv(ghost + (size_t)procOff[i] * nc, (size_t)procCnt[i] * nc, ncclDouble, procPeer[i], comm.nccl, ctx.stream);
            }
            comm.pGroupEnd();
            prof_end(ctx);
            ctx.launches++;
            return;
        }
#endif
        comm.hsend.resize((size_t)nG * nc); comm.hrecv.resize((size_t)nG * nc);
        d2h(ctx, comm.hsend.data(), sendbuf, (size_t)nG * nc * sizeof(double));
        comm.xcb(comm.user, comm.hsend.data(), comm.hrecv.data(), nc);
        h2d(ctx, ghost, comm.hrecv.data(), (size_t)nG * nc * sizeof(double));
    }
#ifndef TPP_EMU
    // one halo exchange over peer memory; `off` = the level's patch offsets in ghost order
    template <class T> void p2pExchange(const int* owner, const std::vector<int>& off, int ng, const T* src, T* ghost, int nc) {
        P2PArgs a;
        memset(&a, 0, sizeof(a));
        a.nG = ng; a.nc = nc; a.nPatch = (int)off.size();
        for (int p = 0; p < a.nPatch; p++) {
            a.off[p] = off[p];
            a.peerData[p] = comm.peerBase[p] + comm.peerOff[p];
            a.peerFlag[p] = (unsigned long long*)(comm.peerBase[p] + 128 * (size_t)comm.peerFlagIdx[p]);
            a.myData[p] = comm.window + comm.myOff[p];
            a.myFlag[p] = (unsigned long long*)(comm.window + 128 * (size_t)p);
            a.slot[p] = comm.slot[p];
        }
        a.owner = owner; a.src = src; a.ghost = ghost; a.seq = comm.seq; a.putDone = comm.putDone; a.err = comm.p2pErr;
        const long total = (long)ng * nc;
        const int grid = (int)std::max(1L, std::min(296L, (total + 1023) / 1024));
        prof_begin(ctx, "halo_p2p");
        if (comm.ll) k_halo_ll<T><<<grid, 256, 0, ctx.stream>>>(a);
        else k_halo_p2p<T><<<grid, 256, 0, ctx.stream>>>(a);
        prof_end(ctx);
        ctx.launches++;
    }
    // windows, IPC handles (gathered through the NCCL communicator), the neighbours' layout
    void setupP2P() {
        if (!comm.active || !comm.nccl || !knob("TPP_P2P", 2)) return;
        const int np = (int)procCnt.size();
        double fail = np > P2P_MAXPATCH ? 1.0 : 0.0;
        size_t total = 4096;  // flags: 128 B apart
        comm.myOff.assign(np, 0); comm.slot.assign(np, 0);
        for (int p = 0; p < np; p++) {
            comm.slot[p] = ((size_t)procCnt[p] * 9 * 2 * sizeof(double) + 255) / 256 * 256;  // up to 9 components (grad U); LL words carry 4 data bytes in 8
            comm.myOff[p] = total;
            total += 2 * comm.slot[p];
        }
        comm.window = (char*)dev_alloc(total);
        comm.seq = (unsigned long long*)dev_alloc(64); comm.arSeq = (unsigned long long*)dev_alloc(64); comm.putDone = (unsigned*)dev_alloc(64); comm.p2pErr = (int*)dev_alloc(64);
        cudaIpcMemHandle_t mine;
        if (cudaIpcGetMemHandle(&mine, comm.window) != cudaSuccess) { cudaGetLastError(); fail = 1.0; }
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        std::vector<double> hv(64 * (size_t)comm.size, 0.0);  // one double per byte: summing with zeros is exact
        for (int k = 0; k < 64; k++) hv[64 * (size_t)comm.rank + k] = (double)((unsigned char*)&mine)[k];
        hostAllreduce(hv, 0);
        std::map<int, char*> base;
        // every rank's window for the all-reduce (<= 8 ranks), the neighbours' for the halos
        std::vector<int> toOpen;
        const bool llar = comm.size <= AR_MAXR && knob("TPP_LLAR", 1);
        if (llar) { for (int q = 0; q < comm.size; q++) if (q != comm.rank) toOpen.push_back(q); }
        else for (int p = 0; p < np; p++) toOpen.push_back(procPeer[p]);
        for (size_t p = 0; p < toOpen.size() && fail == 0.0; p++) {
            int q = toOpen[p];
            if (base.count(q)) continue;
            cudaIpcMemHandle_t h;
            for (int k = 0; k < 64; k++) ((unsigned char*)&h)[k] = (unsigned char)(hv[64 * (size_t)q + k] + 0.5);
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); fail = 1.0; break; }
            comm.opened.push_back(ptr);
            base[q] = (char*)ptr;
        }
        // what the neighbour calls the patch facing me: its data offset and flag index
        std::vector<double> so((size_t)nG), sf((size_t)nG), ro, rf;
        for (int p = 0; p < np; p++) for (int k = 0; k < procCnt[p]; k++) { so[procOff[p] + k] = (double)comm.myOff[p]; sf[procOff[p] + k] = (double)p; }
        hostExchange(procOff, procCnt, procPeer, so, ro);
        hostExchange(procOff, procCnt, procPeer, sf, rf);
        std::vector<double> fv(1, fail);
        hostAllreduce(fv, 1);
        if (fv[0] != 0.0) {
            if (comm.rank == 0) fprintf(stderr, "tppvof: peer-memory halo exchange unavailable (cudaIpc / peer access), using ncclSend/ncclRecv\n");
            return;
        }
        comm.peerBase.assign(np, nullptr); comm.peerOff.assign(np, 0); comm.peerFlagIdx.assign(np, 0);
        for (int p = 0; p < np; p++) {
            comm.peerBase[p] = base[procPeer[p]];
            comm.peerOff[p] = (size_t)(ro[procOff[p]] + 0.5);
            comm.peerFlagIdx[p] = (int)(rf[procOff[p]] + 0.5);
        }
        // nobody may store into a window before everybody has opened and zeroed theirs
        std::vector<double> bar(1, 0.0);
        dev_sync(ctx);
        hostAllreduce(bar, 0);
        comm.p2p = true;
        comm.ll = knob("TPP_P2P", 2) >= 2;
        if (llar) {
            comm.allBase.assign(comm.size, nullptr);
            for (int q = 0; q < comm.size; q++) comm.allBase[q] = q == comm.rank ? comm.window : base[q];
        }
    }
#endif
    // all-reduce of n device scalars scal[idx..idx+n): op 0 sum, 1 max
    void allreduce(int idx, int n, int op) {
        if (!comm.active) return;
#ifndef TPP_EMU
        if (!comm.allBase.empty() && n <= AR_MAXV) {
            ARArgs a;
            memset(&a, 0, sizeof(a));
            a.rank = comm.rank; a.size = comm.size; a.n = n; a.op = op;
            for (int r = 0; r < comm.size; r++) a.win[r] = reinterpret_cast<uint2*>(comm.allBase[r] + 2048);
            a.vals = scal + idx; a.seq = comm.arSeq; a.err = comm.p2pErr;
            prof_begin(ctx, "allreduce_ll");
            k_allreduce_ll<<<1, 64, 0, ctx.stream>>>(a);
            prof_end(ctx);
            ctx.launches++;
            return;
        }
        if (comm.nccl) {
            prof_begin(ctx, "allreduce_nccl");
            comm.pAllReduce(scal + idx, scal + idx, n, ncclDouble, op == 0 ? ncclSum : ncclMax, comm.nccl, ctx.stream);
            prof_end(ctx);
            ctx.launches++;
            return;
        }
#endif
        double v[8];
        d2h(ctx, v, scal + idx, n * sizeof(double));
        comm.rcb(comm.user, v, n, op);
        h2d(ctx, scal + idx, v, n * sizeof(double));
    }
    // processor-face geometry once the neighbours' cell centres are here
    bool finalizeParallel() {
        {   // does any rank hold a fixed-value p_rgh patch?
            double fixes = 0;
            for (int p = 0; p < nPatch; p++) if (bcP[p] == TPP_P_TOTAL_PRESSURE && pSize[p] > 0) fixes = 1;
            h2d(ctx, scal + S_TMP0, &fixes, sizeof(double));
            allreduce(S_TMP0, 1, 1);
            d2h(ctx, &fixes, scal + S_TMP0, sizeof(double));
            if (fixes == 0 && comm.size > 1) {
                // a closed domain (sloshingTank3D6DoF: one wall patch, hierarchical (4 2 2),
                // system/decomposeParDict:17-29): p_rgh needs the reference of fvSolution:85-86 on ONE
                // rank - the lowest one whose share of the mesh contains pRefPoint
                d.needRef = 1;
                const int c = findCell(cfg.p_ref_point);
                double who = c >= 0 ? (double)(comm.size - comm.rank) : 0.0;  // max picks the lowest rank with a hit
                h2d(ctx, scal + S_TMP0, &who, sizeof(double));
                allreduce(S_TMP0, 1, 1);
                d2h(ctx, &who, scal + S_TMP0, sizeof(double));
                if (who == 0.0) { g_err = "pRefPoint is outside the mesh and p_rgh needs a reference"; return false; }
                d.refCell = (comm.size - (int)(who + 0.5)) == comm.rank ? c : -1;
            }
        }
        double ng = (double)nC;
        h2d(ctx, scal + S_TMP0, &ng, sizeof(double));
        allreduce(S_TMP0, 1, 0);
        d2h(ctx, &ng, scal + S_TMP0, sizeof(double));
        nGlobal = (long)(ng + 0.5);
        if (nG == 0) return true;
        X(const_cast<double*>(d.C0), 3);
        std::vector<double> gc(3 * (size_t)nG);
        d2h(ctx, gc.data(), d.C0 + 3 * (size_t)nC, gc.size() * sizeof(double));
        std::vector<double> w(nG), dcv(nG), corr(3 * (size_t)nG), dpn(3 * (size_t)nG);
        for (int j = 0; j < nG; j++) {
            int f = nIloc + j, o = own[f];
            const double* S = &Sf0[3 * f];
            double nf[3] = {S[0] / magSf[f], S[1] / magSf[f], S[2] / magSf[f]};
            double dO[3], dN[3], dd[3];
            for (int k = 0; k < 3; k++) {
                dO[k] = Cf0[3 * f + k] - C0[3 * o + k];
                dN[k] = gc[3 * j + k] - Cf0[3 * f + k];
                dd[k] = gc[3 * j + k] - C0[3 * o + k];
                dpn[3 * j + k] = dd[k];
            }
            double so = fabs(dot3(S, dO)), sn = fabs(dot3(S, dN));
            w[j] = sn / (so + sn);
            dcv[j] = 1.0 / dmax(dot3(nf, dd), 0.05 * mag3(dd));
            for (int k = 0; k < 3; k++) corr[3 * j + k] = nf[k] - dd[k] * dcv[j];
        }
        h2d(ctx, d.w + nIloc, w.data(), nG * sizeof(double));
        h2d(ctx, d.dc + nIloc, dcv.data(), nG * sizeof(double));
        h2d(ctx, d.corrVec + 3 * (size_t)nIloc, corr.data(), corr.size() * sizeof(double));
        h2d(ctx, d.dPN + 3 * (size_t)nIloc, dpn.data(), dpn.size() * sizeof(double));
        if (hasRotation) {
            h2d(ctx, const_cast<double*>(d.corrVec0) + 3 * (size_t)nIloc, corr.data(), corr.size() * sizeof(double));
            h2d(ctx, const_cast<double*>(d.dPN0) + 3 * (size_t)nIloc, dpn.data(), dpn.size() * sizeof(double));
            orientGeometry();
        }
        return true;
    }

    // ---- stages -------------------------------------------------------------------------------
    void readScal() { d2h(ctx, hscal, scal, S_COUNT * sizeof(double)); }

    void courant() {
        LAUNCH_W(ctx, courant, d, nC);
        red.reduce(ctx, d.cellTmp, nullptr, nC, 3, scal + S_MAX0);
        red.reduce(ctx, d.cellTmp + nC, nullptr, nC, 3, scal + S_MAX1);
        allreduce(S_MAX0, 2, 1);
        readScal();
        Co = 0.5 * hscal[S_MAX0] * dt;
        alphaCo = 0.5 * hscal[S_MAX1] * dt;
    }
    void adjustDeltaT() {
        if (!cfg.adjust_time_step) return;
        double dd = cfg.max_delta_t;
        if (Co > SMALL) dd = std::min(dd, cfg.max_co / Co * dt);
        if (alphaCo > SMALL) dd = std::min(dd, cfg.max_alpha_co / alphaCo * dt);
        dt = std::min(1.2 * dt, dd);
        double timeToNextWrite = std::max(0.0, (writeTimeIndex + 1) * cfg.write_interval - (t - startTime));
        double nSteps = timeToNextWrite / dt - SMALL;
        if (nSteps < 2147483647.0) {
            int n = (int)nSteps + 1;
            double nd = timeToNextWrite / n;
            if (nd >= dt) dt = std::min(nd, 2.0 * dt); else dt = std::max(nd, 0.2 * dt);
        }
    }
    // ---- the per-step scalar block (StepScal) and the graph-captured explicit part of a step -------
    // Small meshes (the reference's own 8 k - 42 k-cell cases, and every case of a sweep) spend a step
    // on launch overhead: ~85 explicit kernels between the two pressure solves.  With translation-only
    // motion, an open tank and one rank their arguments never change except for a handful of scalars,
    // so the three explicit segments of a step (up to the first solve, between the solves, after the
    // last) are captured once into CUDA graphs whose kernels read those scalars from `ssDev`.  A step is
    // then ~3 graph launches + the solver's iteration graphs: 8 concurrent sweep cases no longer queue
    // behind the context's launch path.
    StepScal* ssDev = nullptr;
    StepScal* ssHost = nullptr;  // pinned (CUDA) / the block itself (host emulation)
    struct SegGraph { bool have = false;
#ifndef TPP_EMU
        cudaGraphExec_t exec = nullptr;
#endif
        long nodes = 0; };
    SegGraph seg[3];
    bool stepScalMode() const { return !comm.active && !hasRotation && cfg.n_motion >= 0 && knob("TPP_STEP_SS", 1) != 0; }
    bool stepGraphMode() const {
#ifdef TPP_EMU
        return false;
#else
        return stepScalMode() && !d.needRef && !ctx.prof && cfg.n_non_orth == 0 && ctx.stream != nullptr && ctx.stream != cudaStreamLegacy && knob("TPP_STEP_GRAPH", nGlobal <= 1000000 ? 1 : 0) != 0;
#endif
    }
    void ensureStepScal() {
        if (ssHost) return;
#ifdef TPP_EMU
        ssHost = (StepScal*)calloc(1, sizeof(StepScal));
        ssDev = ssHost;
#else
        CUDA_CHECK(cudaMallocHost(&ssHost, sizeof(StepScal)));
        memset(ssHost, 0, sizeof(StepScal));
        ssDev = (StepScal*)dev_alloc(sizeof(StepScal));
#endif
    }
    // the current values -> the pinned block (and, outside a captured segment, on to the device)
    void fillStepScal(bool upload) {
        if (!stepScalMode()) { d.ss = nullptr; return; }
        ensureStepScal();
        ssHost->dt = dt;
        ssHost->rdt[0] = 1.0 / dt;
        ssHost->rdt[1] = 1.0 / (dt / std::max(cfg.n_alpha_subcycles, 1));
        for (int k = 0; k < 3; k++) { ssHost->dT[k] = Tn[k] - To[k]; ssHost->wallU[k] = (Tn[k] - To[k]) / dt; ssHost->Tn[k] = Tn[k]; }
        d.ss = ssDev;
#ifndef TPP_EMU
        if (upload) CUDA_CHECK(cudaMemcpyAsync(ssDev, ssHost, sizeof(StepScal), cudaMemcpyHostToDevice, ctx.stream));
#else
        (void)upload;
#endif
    }
    // run `body` (kernel launches with fixed arguments) as segment k: captured on first use, replayed after
    template <class F> void segment(int k, F body) {
#ifndef TPP_EMU
        if (stepGraphMode()) {
            SegGraph& g = seg[k];
            if (!g.have) {
                const long l0 = ctx.launches;
                cudaGraph_t gr;
                CUDA_CHECK(cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal));
                body();
                CUDA_CHECK(cudaStreamEndCapture(ctx.stream, &gr));
                CUDA_CHECK(cudaGraphInstantiate(&g.exec, gr, 0));
                cudaGraphDestroy(gr);
                g.nodes = ctx.launches - l0;
                ctx.launches = l0;
                g.have = true;
            }
            CUDA_CHECK(cudaGraphLaunch(g.exec, ctx.stream));
            ctx.launches += g.nodes;
            return;
        }
#endif
        (void)k;
        body();
    }
    void dropStepGraphs() {
#ifndef TPP_EMU
        for (auto& g : seg) { if (g.have) cudaGraphExecDestroy(g.exec); g = SegGraph(); }
#endif
    }
    bool advanceTimeHost() {
        dt0 = dt;
        t += dt;
        step++;
        int wi = (int)(((t - startTime) + 0.5 * dt) / cfg.write_interval);
        if (wi > writeTimeIndex) { writeTimeIndex = wi; return true; }
        return false;
    }
    void advanceTimeDevice() {
        X(d.U, 3);  // ghosts may be stale after a tpp_set
        d2d(ctx, d.U0, d.U, 3 * (size_t)(nC + nG) * sizeof(double));
        d2d(ctx, d.U0_b, d.U_b, 3 * (size_t)nB * sizeof(double));
        d2d(ctx, d.rho0, d.rho, (size_t)(nC + nG) * sizeof(double));
        d2d(ctx, d.Uf0, d.Uf, 3 * (size_t)nF * sizeof(double));
    }
    bool advanceTime() {
        bool wr = advanceTimeHost();
        advanceTimeDevice();
        return wr;
    }
    void moveMeshHost() {
        d.dt = dt;
        if (cfg.n_motion <= 0) return;
        memcpy(Ro, Rn, sizeof(Rn)); memcpy(To, Tn, sizeof(Tn));
        motionAt(t, Rn, Tn);
        setTransform();
    }
    void moveMeshDevice() {
        if (cfg.n_motion <= 0) return;
        orientGeometry();
        if (hasRotation) LAUNCH(ctx, meshphi_rot, d, nF);
        else LAUNCH(ctx, meshphi_trans, d, nF);
    }
    void moveMesh() {
        moveMeshHost();
        fillStepScal(true);
        moveMeshDevice();
    }
    void alphaBCs() { LAUNCH(ctx, alpha_bc, d, nB); }
    void UBCs() { d.dt = dt; LAUNCH(ctx, U_bc, d, nB); }
    void mixture() { LAUNCH(ctx, mixture_cell, d, nC); LAUNCH(ctx, mixture_bnd, d, nB); }
    void gradScalar(const double* s, const double* sb, double* out) {
        d.gs = s; d.gsb = sb; d.gout = out;
        LAUNCH_W(ctx, grad_scalar, d, nC);
    }
    // accumulate: fold this sub-cycle's share of the step's alpha flux (alphaphi_acc) and the limited
    // flux into the last limiter iteration's face pass
    void alphaSubCycle(double dts, bool accumulate = false) {
        d.rDeltaT = 1.0 / dts; d.rdtSel = cfg.n_alpha_subcycles > 1 ? 1 : 0;
        d2d(ctx, d.alpha0, d.alpha, nC * sizeof(double));
        alphaBCs();
        X(d.alpha, 1);
        gradScalar(d.alpha, d.alpha_b, d.grad);
        X(d.grad, 3);
        LAUNCH(ctx, alpha_flux, d, nF);
        LAUNCH_W(ctx, mules_setup, d, nC);
        const bool fuse = accumulate && cfg.n_limiter_iter > 0 && knob("TPP_MULES_FUSE", 1);
        for (int j = 0; j < cfg.n_limiter_iter; j++) {
            LAUNCH_W(ctx, mules_cell, d, nC);
            X(d.lambdap, 1);
            X(d.lambdam, 1);
            if (fuse && j == cfg.n_limiter_iter - 1) LAUNCH(ctx, mules_face_final, d, nF);
            else LAUNCH(ctx, mules_face, d, nI);
        }
        if (!fuse) {
            LAUNCH(ctx, mules_phipsi, d, nF);
            if (accumulate) LAUNCH(ctx, alphaphi_acc, d, nF);
        }
        LAUNCH_W(ctx, mules_update, d, nC);
        alphaBCs();
    }
    void alphaPredictor() {
        int n = cfg.n_alpha_subcycles;
        if (n > 1) {
            double total = dt, dts = dt / n;
            dev_zero(ctx, d.alphaPhi, nF * sizeof(double));
            d.subW = dts / total;
            for (int s = 0; s < n; s++)
                for (int a = 0; a < cfg.n_alpha_corr; a++) alphaSubCycle(dts, a == cfg.n_alpha_corr - 1);
        } else {
            for (int a = 0; a < cfg.n_alpha_corr; a++) alphaSubCycle(dt);
            d2d(ctx, d.alphaPhi, d.alphaPhiUn, nF * sizeof(double));
        }
        mixture();
        X(d.alpha, 1);
        X(d.rho, 1);
        LAUNCH(ctx, rhophi, d, nF);
        interfaceCorrect();
    }
    // interfaceProperties::correct + surfaceTensionForce for sigma != 0 (tpp_kernels.h); alpha, alpha_b and
    // the alpha ghosts are the predictor's final ones
    void interfaceCorrect() {
        if (cfg.sigma == 0.0) return;
        gradScalar(d.alpha, d.alpha_b, d.gradA);
        X(d.gradA, 3);
        LAUNCH(ctx, nhat_face, d, nF);
        LAUNCH_W(ctx, curvature, d, nC);
        X(d.sigmaK, 1);
        LAUNCH(ctx, stf_face, d, nF);
    }
    void momentum() {
        UBCs();
        X(d.U, 3);
        d.rDeltaT = 1.0 / dt; d.rdtSel = 0;
        LAUNCH_W(ctx, grad_U, d, nC);
        X(d.gradU, 9);
        LAUNCH(ctx, mom_face, d, nI);
        LAUNCH(ctx, mom_bnd, d, nB);
        LAUNCH_W(ctx, mom_cell, d, nC);
    }
    void computeHbyA() {
        LAUNCH_W(ctx, HbyA, d, nC);
        LAUNCH(ctx, HbyA_bnd, d, nB);
        X(d.rAU, 1);
        X(d.HbyA, 3);
    }
    void pcPrepare() {
        d.dt = dt; d.rDeltaT = 1.0 / dt; d.rdtSel = 0;
        computeHbyA();
        gradScalar(d.rho, d.rho_b, d.grad);
        X(d.grad, 3);
        LAUNCH(ctx, phiHbyA, d, nF);
    }
    void pcAssemble() {
        LAUNCH(ctx, p_total, d, nB);
        X(d.p_rgh, 1);
        gradScalar(d.p_rgh, d.p_rgh_b, d.grad);
        X(d.grad, 3);
        LAUNCH(ctx, p_face, d, nI);
        LAUNCH_W(ctx, p_cell, d, nC);
    }
    void pcFinish() {
        X(d.p_rgh, 1);
        LAUNCH(ctx, flux, d, nF);
        LAUNCH_W(ctx, U_recon, d, nC);
        UBCs();
        X(d.U, 3);
    }
    void pcEnd() {
        if (cfg.n_motion > 0) LAUNCH(ctx, Uf, d, nF);
        LAUNCH(ctx, p, d, nC);
        if (d.needRef) {
            double pc;
            if (comm.active) {  // the rank that owns the reference cell tells the others
                if (d.refCell >= 0) d2d(ctx, scal + S_TMP0, d.p + d.refCell, sizeof(double));
                else dev_zero(ctx, scal + S_TMP0, sizeof(double));
                allreduce(S_TMP0, 1, 0);
                d2h(ctx, &pc, scal + S_TMP0, sizeof(double));
            } else
                d2h(ctx, &pc, d.p + d.refCell, sizeof(double));
            d.pRefShift = cfg.p_ref_value - pc;
            LAUNCH(ctx, p_shift, d, nC);
            LAUNCH(ctx, p_evaluate, d, nB);
        }
    }
    void pressureCorrector(bool finalIter) {
        pcPrepare();
        for (int nonOrth = 0; nonOrth <= cfg.n_non_orth; nonOrth++) {
            bool finalNonOrth = nonOrth == cfg.n_non_orth;
            pcAssemble();
            int which = (finalIter && finalNonOrth) ? 1 : 0;
            const tpp_solver_t& ctl = which ? cfg.p_rgh_final : cfg.p_rgh;
            lastSolve[which] = solve(ctl, d.pDiag, d.pUpper, d.pSource, d.p_rgh);
            LAUNCH(ctx, p_evaluate, d, nB);
            if (finalNonOrth) pcFinish();
        }
        pcEnd();
    }
    // ---- run statistics (tpp_stats): iteration counts over the steps since the last reset and the
    // alpha-volume balance sum(alpha V)(t) - sum(alpha V)(t_reset) + int dt sum_b alphaPhi_b = 0
    bool statsOn = false;
    double stSteps = 0, stIt[2] = {0, 0}, stItMax[2] = {0, 0}, stCap = 0, stVol0 = 0, stBndInt = 0;
    double alphaVolume() {
        red.reduce(ctx, d.alpha, d.V, nC, 0, scal + S_TMP0);
        double v;
        d2h(ctx, &v, scal + S_TMP0, sizeof(double));
        return v;
    }
    void statsReset() {
        stSteps = stIt[0] = stIt[1] = stItMax[0] = stItMax[1] = stCap = stBndInt = 0;
        stVol0 = alphaVolume();
    }
    void statsStep() {
        stSteps += 1;
        for (int k = 0; k < 2; k++) { stIt[k] += lastSolve[k].iters; stItMax[k] = std::max(stItMax[k], (double)lastSolve[k].iters); }
        if (lastSolve[1].iters >= cfg.p_rgh_final.max_iter) stCap += 1;
        if (nB > 0) {
            red.reduce(ctx, d.alphaPhi + nI, nullptr, nB, 2, scal + S_TMP0);  // nI counts the processor faces; nB the physical boundary
            double v;
            d2h(ctx, &v, scal + S_TMP0, sizeof(double));
            stBndInt += dt * v;
        }
    }
    // ---- asynchronous read-back (tpp_get_async / tpp_sync): a snapshot of the array (file order) is
    // taken on the solver's stream into a staging buffer of its own, the copy to the host runs on a
    // second stream and overlaps whatever the solver does next (the next step, the next tpp_set)
    std::map<std::string, double*> stage;
#ifndef TPP_EMU
    cudaStream_t d2hStream = nullptr;
    cudaEvent_t snapEvent = nullptr;
    std::map<std::string, cudaEvent_t> copied;  // per array: its last copy to the host has left the staging buffer
#endif
    // ---- interface statistics on the device (tpp_interface) ---------------------------
    bool isoBuilt = false;
    IsoArgs iso;
    void buildIso() {
        isoBuilt = true;
        memset(&iso, 0, sizeof(iso));
        // distinct (point, cell) incidences and distinct edges from the face loops
        std::vector<unsigned long long> pc, ed;
        pc.reserve(fLab.size() * 2); ed.reserve(fLab.size());
        for (int f = 0; f < nF; f++) {
            const int b = fOff[f], n = fOff[f + 1] - b;
            for (int i = 0; i < n; i++) {
                const unsigned p0 = (unsigned)fLab[b + i], p1 = (unsigned)fLab[b + (i + 1) % n];
                pc.push_back(((unsigned long long)p0 << 32) | (unsigned)own[f]);
                if (f < nI && nei[f] < nC) pc.push_back(((unsigned long long)p0 << 32) | (unsigned)nei[f]);
                ed.push_back(((unsigned long long)std::min(p0, p1) << 32) | std::max(p0, p1));
            }
        }
        std::sort(pc.begin(), pc.end()); pc.erase(std::unique(pc.begin(), pc.end()), pc.end());
        std::sort(ed.begin(), ed.end()); ed.erase(std::unique(ed.begin(), ed.end()), ed.end());
        std::vector<int> st(nP + 1, 0), cells(pc.size()), ea(ed.size()), eb(ed.size());
        for (size_t k = 0; k < pc.size(); k++) { st[(pc[k] >> 32) + 1]++; cells[k] = (int)(pc[k] & 0xffffffffu); }
        for (int p = 0; p < nP; p++) st[p + 1] += st[p];
        for (size_t k = 0; k < ed.size(); k++) { ea[k] = (int)(ed[k] >> 32); eb[k] = (int)(ed[k] & 0xffffffffu); }
        iso.nP = nP; iso.nE = (int)ed.size();
        iso.pcStart = upload(st); iso.pcCells = upload(cells); iso.edgeA = upload(ea); iso.edgeB = upload(eb);
        iso.points0 = d.points0 ? d.points0 : upload(points0);
        iso.ptAlpha = A<double>(nP); iso.partial = A<double>(4 * (size_t)RED_BLOCKS); iso.out = A<double>(4);
    }
    void interfaceSummary(double level, double* out4) {
        if (!isoBuilt) buildIso();
        iso.alpha = d.alpha; iso.iso = level;
        for (int k = 0; k < 9; k++) iso.R[k] = Rn[k];
        for (int k = 0; k < 3; k++) { iso.T[k] = Tn[k]; iso.cofg[k] = cfg.cofg[k]; }
#ifdef TPP_EMU
        for (int p = 0; p < nP; p++) b_iso_point(iso, p);
        double s = 0, n = 0, mx = -1e300, mn = 1e300;
        for (int e = 0; e < iso.nE; e++) { double z; if (iso_edge_z(iso, e, z)) { s += z; n += 1; mx = std::max(mx, z); mn = std::min(mn, z); } }
        out4[0] = n > 0 ? mx : 0; out4[1] = n > 0 ? mn : 0; out4[2] = n > 0 ? s / n : 0; out4[3] = n;
        ctx.launches += 3;
#else
        prof_begin(ctx, "iso_surface");
        k_iso_point<<<(nP + 255) / 256, 256, 0, ctx.stream>>>(iso);
        const int nb = std::max(1, std::min(RED_BLOCKS, (iso.nE + 255) / 256));
        k_iso_edges<<<nb, 256, 0, ctx.stream>>>(iso);
        k_iso_final<<<1, 32, 0, ctx.stream>>>(iso, nb);
        LAUNCH_CHECK("k_iso");
        prof_end(ctx);
        ctx.launches += 3;
        d2h(ctx, out4, iso.out, 4 * sizeof(double));
#endif
    }
    void fail(const std::string& msg) {
        if (ctx.err.empty()) { ctx.err = msg; fprintf(stderr, "tppvof: %s\n", msg.c_str()); }
    }
    // device-side error flags (bounded waits that gave up), read once per step: a run that lost a
    // halo or a grid barrier must stop with an error, not carry on with stale data
    void checkDeviceFlags() {
#ifndef TPP_EMU
        if (comm.p2p) {
            int e = 0;
            d2h(ctx, &e, comm.p2pErr, sizeof(int));
            if (e) fail("peer-memory halo exchange / all-reduce: a neighbour's data did not arrive within 30 s");
        }
#endif
        if (!tail.empty()) {
            int e = 0;
            d2h(ctx, &e, tailErr, sizeof(int));
            if (e) fail("vk_tail: grid barrier timed out (the cooperative grid was not co-resident?)");
        }
    }
    bool oneStep() {
        courant();
        if (!std::isfinite(Co) || !std::isfinite(alphaCo)) fail("Courant number is not finite (the solution diverged)");
        adjustDeltaT();
        if (!std::isfinite(dt) || !(dt > 0)) fail("deltaT is not finite / positive");
        bool wr = advanceTimeHost();
        moveMeshHost();
        if (stepGraphMode()) {
            // segment 0: everything up to the first pressure solve; 1: between two solves; 2: after the last
            fillStepScal(false);  // pinned block only: the segment's first node copies it to the device
            segment(0, [&] {
                CUDA_CHECK_OR_EMU(cudaMemcpyAsync(ssDev, ssHost, sizeof(StepScal), cudaMemcpyHostToDevice, ctx.stream));
                ctx.launches++;
                advanceTimeDevice();
                moveMeshDevice();
                alphaPredictor();
                momentum();
                pcPrepare();
                pcAssemble();
            });
            for (int corr = 0; corr < cfg.n_correctors; corr++) {
                const bool last = corr == cfg.n_correctors - 1;
                const int which = last ? 1 : 0;
                lastSolve[which] = solve(which ? cfg.p_rgh_final : cfg.p_rgh, d.pDiag, d.pUpper, d.pSource, d.p_rgh);
                segment(last ? 2 : 1, [&] {
                    LAUNCH(ctx, p_evaluate, d, nB);
                    pcFinish();
                    pcEnd();
                    if (!last) { pcPrepare(); pcAssemble(); }
                });
            }
        } else {
            fillStepScal(true);
            advanceTimeDevice();
            moveMeshDevice();
            alphaPredictor();
            momentum();
            for (int corr = 0; corr < cfg.n_correctors; corr++) pressureCorrector(corr == cfg.n_correctors - 1);
        }
        if (!probeCells.empty()) sampleProbes();
        if (statsOn) statsStep();
        checkDeviceFlags();
        return wr;
    }
    // probes (system/functions:17-33): one small gather kernel per step appends a row to a device
    // ring; the rows come to the host in one copy when the ring is full, at a write time
    // (tpp_run_to_write returns) or when tpp_probe_log asks - not one blocking copy per probe and step
    static constexpr int PROBE_RING = 1024;
    int* dProbeCells = nullptr;
    double* dProbeRing = nullptr;
    int probePending = 0, probeUploaded = -1;
    std::vector<double> probeTimes;
    void sampleProbes() {
        const int np = (int)probeCells.size();
        if (probeUploaded != np || dProbeCells == nullptr) {
            flushProbes();
            dProbeCells = upload(probeCells);
            dProbeRing = A<double>((size_t)PROBE_RING * np);
            probeUploaded = np;
        }
#ifdef TPP_EMU
        for (int k = 0; k < np; k++) dProbeRing[(size_t)probePending * np + k] = probeCells[k] >= 0 ? d.p[probeCells[k]] : -1.79769e+307;
#else
        k_probe_row<<<1, 64, 0, ctx.stream>>>(d.p, dProbeCells, np, dProbeRing + (size_t)probePending * np);
#endif
        ctx.launches++;
        probeTimes.push_back(t);
        if (++probePending == PROBE_RING) flushProbes();
    }
    void flushProbes() {
        const int np = probeUploaded;
        if (probePending == 0 || np <= 0) return;
        std::vector<double> rows((size_t)probePending * np);
        d2h(ctx, rows.data(), dProbeRing, rows.size() * sizeof(double));
        for (int r = 0; r < probePending; r++) {
            probeLog.push_back(probeTimes[r]);
            for (int k = 0; k < np; k++) probeLog.push_back(rows[(size_t)r * np + k]);
        }
        probeTimes.clear();
        probePending = 0;
    }

    // ---- multigrid hierarchy (cached: the mesh only moves rigidly) ------------------------------
    // Levels: 0 = the mesh (ELL), levels[l] = hierarchy level l + 1 (CSR).  levels[0..gatherLevel)
    // are smoothed where they live (rows distributed over the ranks, halo exchange per operator
    // application); levels[gatherLevel] is assembled distributed, then gathered onto every rank
    // as tail[0]; tail[1..] continue the coarsening on the gathered graph (replicated).
    LV fineView(double* diag, double* upper) {
        LV L;
        memset(&L, 0, sizeof(L));
        L.n = nC; L.nf = nI; L.nCp = nCp; L.W = W; L.ell = 1;
        L.cf = d.cf; L.cn = d.cn; L.own = d.own; L.nei = d.nei;
        L.diag = diag; L.upper = upper; L.rsum = fineRsum; L.ev = fineEv;
        L.nOwn = nC + nG; L.nGlob = (double)nGlobal;
        return L;
    }
    static LV viewOf(Level& v) {
        LV L;
        memset(&L, 0, sizeof(L));
        L.n = v.n; L.nf = v.nf; L.ell = 0; L.cf = v.cf; L.cn = v.cn; L.rs = v.rs; L.own = v.own; L.nei = v.nei;
        L.diag = v.diag; L.upper = v.upper; L.rsum = v.rsum; L.ev = v.ev; L.nOwn = v.n + v.nG; L.nGlob = v.nGlob;
        L.aggStart = v.aggStart; L.aggRows = v.aggRows; L.segStart = v.segStart; L.segFaces = v.segFaces;
        return L;
    }

    // ---- host-side collectives used while the hierarchy is built ------------------------------
    // neighbour exchange of one double per processor face of a (coarse) level's halo
    void hostExchange(const std::vector<int>& poff, const std::vector<int>& pcnt, const std::vector<int>& ppeer, const std::vector<double>& send, std::vector<double>& recv) {
        recv.assign(send.size(), 0.0);
        if (!comm.active || send.empty()) return;
#ifndef TPP_EMU
        if (comm.nccl) {
            double* ds = dalloc<double>(send.size());
            double* dr = dalloc<double>(send.size());
            h2d(ctx, ds, send.data(), send.size() * sizeof(double));
            comm.pGroupStart();
            for (size_t i = 0; i < pcnt.size(); i++) {
                if (pcnt[i] == 0) continue;
                comm.pSend(ds + poff[i], (size_t)pcnt[i], ncclDouble, ppeer[i], comm.nccl, ctx.stream);
                comm.pRecv(dr + poff[i], (size_t)pcnt[i], ncclDouble, ppeer[i], comm.nccl, ctx.stream);
            }
            comm.pGroupEnd();
            d2h(ctx, recv.data(), dr, send.size() * sizeof(double));
            dev_free(ds); dev_free(dr);
            return;
        }
#endif
        // callback transport: it knows the mesh-level ghost layout only; a coarse patch never has
        // more faces than the mesh-level patch it agglomerates, so it rides in that patch's slots
        std::vector<double> fs((size_t)nG, 0.0), fr((size_t)nG, 0.0);
        for (size_t i = 0; i < pcnt.size(); i++) for (int k = 0; k < pcnt[i]; k++) fs[procOff[i] + k] = send[poff[i] + k];
        comm.xcb(comm.user, fs.data(), fr.data(), 1);
        for (size_t i = 0; i < pcnt.size(); i++) for (int k = 0; k < pcnt[i]; k++) recv[poff[i] + k] = fr[procOff[i] + k];
    }
    void hostAllreduce(std::vector<double>& v, int op) {
        if (!comm.active || v.empty()) return;
#ifndef TPP_EMU
        if (comm.nccl) {
            double* dv = dalloc<double>(v.size());
            h2d(ctx, dv, v.data(), v.size() * sizeof(double));
            comm.pAllReduce(dv, dv, v.size(), ncclDouble, op == 0 ? ncclSum : ncclMax, comm.nccl, ctx.stream);
            d2h(ctx, v.data(), dv, v.size() * sizeof(double));
            dev_free(dv);
            return;
        }
#endif
        comm.rcb(comm.user, v.data(), (int)v.size(), op);
    }
    double globalCount(double n) {
        std::vector<double> v(1, n);
        hostAllreduce(v, 0);
        return v[0];
    }
    // device vectors: sum over the ranks, in place
    template <class R> void allreduceDev(R* p, size_t n) {
        if (!comm.active || n == 0) return;
#ifndef TPP_EMU
        if (comm.nccl) {
            prof_begin(ctx, "tail_allreduce");
            comm.pAllReduce(p, p, n, sizeof(R) == 4 ? ncclFloat : ncclDouble, ncclSum, comm.nccl, ctx.stream);
            prof_end(ctx);
            ctx.launches++;
            return;
        }
#endif
        std::vector<R> h(n);
        d2h(ctx, h.data(), p, n * sizeof(R));
        std::vector<double> v(h.begin(), h.end());
        comm.rcb(comm.user, v.data(), (int)n, 0);
        for (size_t i = 0; i < n; i++) h[i] = (R)v[i];
        h2d(ctx, p, h.data(), n * sizeof(R));
    }

    // every rank's slice of the tail right-hand side -> every rank (the slices are disjoint)
    template <class R> void gatherTail(R* vec) {
        if (!comm.active) return;
#ifndef TPP_EMU
        if (!comm.gBase.empty()) {
            GatherArgs2 a;
            memset(&a, 0, sizeof(a));
            a.rank = comm.rank; a.size = comm.size;
            for (int r = 0; r <= comm.size; r++) a.rowOff[r] = comm.gRowOff[r];
            for (int r = 0; r < comm.size; r++) a.win[r] = reinterpret_cast<uint2*>(comm.gBase[r]);
            a.vec = vec; a.seq = comm.gSeq; a.done = comm.gDone; a.err = comm.p2pErr;
            const int nb = std::max(1, std::min(32, (tail[0].n + 255) / 256));
            prof_begin(ctx, "tail_gather_ll");
            k_gather_ll<R><<<nb, 256, 0, ctx.stream>>>(a);
            LAUNCH_CHECK("k_gather_ll");
            prof_end(ctx);
            ctx.launches++;
            return;
        }
#endif
        allreduceDev(vec, (size_t)tail[0].n);
    }
    // the gather windows: allocated once the tail is known, handles exchanged like the halo windows'
    void setupGatherWindow() {
#ifndef TPP_EMU
        if (!comm.active || comm.allBase.empty() || tail.empty() || comm.size > AR_MAXR || !knob("TPP_LLGATHER", 1)) return;
        const size_t bytes = 2 * (size_t)tail[0].n * 2 * sizeof(uint2) + 256;  // two parities, up to two words per value
        comm.gwin = (char*)dev_alloc(bytes);
        comm.gSeq = (unsigned long long*)dev_alloc(64); comm.gDone = (unsigned*)dev_alloc(64);
        cudaIpcMemHandle_t mine;
        double fail = cudaIpcGetMemHandle(&mine, comm.gwin) != cudaSuccess ? 1.0 : 0.0;
        if (fail != 0.0) cudaGetLastError();
        std::vector<double> hv(64 * (size_t)comm.size, 0.0);
        for (int k = 0; k < 64; k++) hv[64 * (size_t)comm.rank + k] = (double)((unsigned char*)&mine)[k];
        hostAllreduce(hv, 0);
        std::vector<char*> base(comm.size, nullptr);
        for (int q = 0; q < comm.size && fail == 0.0; q++) {
            if (q == comm.rank) { base[q] = comm.gwin; continue; }
            cudaIpcMemHandle_t h;
            for (int k = 0; k < 64; k++) ((unsigned char*)&h)[k] = (unsigned char)(hv[64 * (size_t)q + k] + 0.5);
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); fail = 1.0; break; }
            comm.opened.push_back(ptr);
            base[q] = (char*)ptr;
        }
        std::vector<double> fv(1, fail);
        dev_sync(ctx);
        hostAllreduce(fv, 1);  // also the barrier: nobody stores before everybody has opened and zeroed
        if (fv[0] == 0.0) comm.gBase = base;
#endif
    }

    // one pairwise matching pass on the device; returns host `root`
    int matchCap = 0;
    void matchPass(LV G, int n, const double* fwDev, std::vector<int>& rootH) {
        if (n > matchCap) {
            dev_free(match); dev_free(prop); dev_free(root);
            matchCap = n;
            match = dalloc<int>(n); prop = dalloc<int>(n); root = dalloc<int>(n);
        }
        std::vector<int> m1(n, -1);
        h2d(ctx, match, m1.data(), n * sizeof(int));
        G.match = match; G.prop = prop; G.root = root; G.fw = fwDev; G.nOwn = n;  // ghost rows never pair
        for (int r = 0; r < knob("TPP_ROUNDS", 8); r++) {
            LAUNCH(ctx, match_propose, G, n);
            LAUNCH(ctx, match_accept, G, n);
        }
        LAUNCH(ctx, match_root, G, n);
        rootH.resize(n);
        d2h(ctx, rootH.data(), root, n * sizeof(int));
    }

    // a level's graph on the host: faces [0,nfLoc) join two owned rows, faces [nfLoc,nf) are
    // processor faces whose neighbour is the ghost row n + j (j in halo order)
    struct HostGraph {
        int n = 0, nf = 0, nfLoc = 0;
        double nGlob = 0;
        std::vector<int> own, nei;
        std::vector<double> fw;
        std::vector<int> poff, pcnt, ppeer;  // halo patches
        std::vector<int> ghostRow;           // the peer's row behind each processor face
        int nG() const { return nf - nfLoc; }
    };
    // CSR rows (neighbour-side faces then owner-side faces: ascending face index)
    static void csrOf(const HostGraph& g, std::vector<int>& rs, std::vector<int>& cf, std::vector<int>& cn) {
        rs.assign(g.n + 1, 0);
        for (int f = 0; f < g.nf; f++) { rs[g.own[f] + 1]++; if (f < g.nfLoc) rs[g.nei[f] + 1]++; }
        for (int c = 0; c < g.n; c++) rs[c + 1] += rs[c];
        cf.assign((size_t)rs[g.n], -1); cn.assign((size_t)rs[g.n], -1);
        std::vector<int> cur(rs.begin(), rs.end() - 1);
        for (int f = 0; f < g.nf; f++) {
            int o = g.own[f], n = g.nei[f];
            cf[cur[o]] = f << 1; cn[cur[o]] = n; cur[o]++;
            if (f < g.nfLoc) { cf[cur[n]] = (f << 1) | 1; cn[cur[n]] = o; cur[n]++; }
        }
    }
    // aggregate `g` by the map agg (nc aggregates; ghostAgg = the peers' aggregates behind the
    // processor faces): coarse graph and, per coarse face, the fine faces summed into it.
    // Coarse processor faces are the distinct (my aggregate, peer aggregate) pairs of a patch in
    // the order of (aggregate on the lower rank, aggregate on the higher rank): both sides of an
    // interface enumerate them identically, as OpenFOAM's GAMG interface agglomeration does.
    void coarsen(const HostGraph& g, const std::vector<int>& agg, int nc, const std::vector<int>& ghostAgg, HostGraph& c, std::vector<int>& segStart, std::vector<int>& segFaces) {
        typedef std::pair<unsigned long long, int> KF;
        std::vector<KF> keys;
        keys.reserve(g.nfLoc);
        for (int f = 0; f < g.nfLoc; f++) {
            int a = agg[g.own[f]], b = agg[g.nei[f]];
            if (a != b) keys.push_back({((unsigned long long)std::min(a, b) << 32) | (unsigned)std::max(a, b), f});
        }
        std::sort(keys.begin(), keys.end());
        c = HostGraph();
        c.n = nc;
        segStart.clear(); segFaces.clear(); segFaces.reserve(keys.size() + g.nG());
        unsigned long long last = ~0ull;
        for (size_t k = 0; k < keys.size(); k++) {
            if (keys[k].first != last) {
                c.own.push_back((int)(keys[k].first >> 32));
                c.nei.push_back((int)(keys[k].first & 0xffffffffu));
                c.fw.push_back(0.0);
                segStart.push_back((int)segFaces.size());
                last = keys[k].first;
            }
            segFaces.push_back(keys[k].second);
            c.fw.back() += g.fw[keys[k].second];
        }
        c.nfLoc = (int)c.own.size();
        for (size_t p = 0; p < g.pcnt.size(); p++) {
            const bool low = comm.rank < g.ppeer[p];
            keys.clear();
            for (int k = 0; k < g.pcnt[p]; k++) {
                int j = g.poff[p] + k, f = g.nfLoc + j;
                unsigned a = (unsigned)agg[g.own[f]], b = (unsigned)ghostAgg[j];
                keys.push_back({low ? ((unsigned long long)a << 32) | b : ((unsigned long long)b << 32) | a, f});
            }
            std::sort(keys.begin(), keys.end());
            c.poff.push_back((int)c.own.size() - c.nfLoc);
            c.ppeer.push_back(g.ppeer[p]);
            last = ~0ull;
            int cnt = 0;
            for (size_t k = 0; k < keys.size(); k++) {
                if (keys[k].first != last) {
                    unsigned hi = (unsigned)(keys[k].first >> 32), lo = (unsigned)(keys[k].first & 0xffffffffu);
                    c.own.push_back((int)(low ? hi : lo));
                    c.ghostRow.push_back((int)(low ? lo : hi));
                    c.nei.push_back(nc + (int)c.own.size() - 1 - c.nfLoc);
                    c.fw.push_back(0.0);
                    segStart.push_back((int)segFaces.size());
                    last = keys[k].first;
                    cnt++;
                }
                segFaces.push_back(keys[k].second);
                c.fw.back() += g.fw[keys[k].second];
            }
            c.pcnt.push_back(cnt);
        }
        segStart.push_back((int)segFaces.size());
        c.nf = (int)c.own.size();
    }
    template <class T> static T* upNew(Ctx& ctx, const std::vector<T>& h) {
        T* p = dalloc<T>(std::max<size_t>(h.size(), 1));
        if (!h.empty()) h2d(ctx, p, h.data(), h.size() * sizeof(T));
        return p;
    }
    // ELL + overflow form of a level's CSR rows: the width is the smallest of 6/8/12/16 that leaves
    // at most 4 % of the entries in the overflow lists
    void buildEllc(Level& v, const std::vector<int>& rs, const std::vector<int>& cn, int ovInv = 25) {
        const int n = v.n;
        if (n == 0) return;
        int Wl = 16;
        for (int w : {6, 8, 12, 16}) {
            long ov = 0;
            for (int i = 0; i < n; i++) ov += std::max(0, rs[i + 1] - rs[i] - w);
            if (ov * ovInv <= (long)rs[n]) { Wl = w; break; }
        }
        const int nPad = (n + 31) / 32 * 32;
        std::vector<int> ecn((size_t)Wl * nPad, -1), esrc((size_t)Wl * nPad, -1), ors(n + 1, 0), ocn, osrc;
        for (int i = 0; i < n; i++) {
            for (int k = rs[i]; k < rs[i + 1]; k++) {
                int s = k - rs[i];
                if (s < Wl) { ecn[(size_t)s * nPad + i] = cn[k]; esrc[(size_t)s * nPad + i] = k; }
                else { ocn.push_back(cn[k]); osrc.push_back(k); }
            }
            ors[i + 1] = (int)ocn.size();
        }
        v.ellW = Wl; v.nPad = nPad; v.nOv = (int)ocn.size();
        v.ecn = upNew(ctx, ecn); v.esrc = upNew(ctx, esrc); v.ors = upNew(ctx, ors); v.ocn = upNew(ctx, ocn); v.osrc = upNew(ctx, osrc);
    }
    // one hierarchy level below `g`: `passes` matching passes merged (mergeLevels); g becomes the
    // coarse graph.  dist: rows are distributed (halo patches, global decisions).  Returns false
    // when the coarsening has stalled.
    bool makeLevel(HostGraph& g, bool fineEll, bool dist, int stopRows, Level& v) {
        std::vector<int> aggTot, segS, segF;
        HostGraph cur = g, nxt;
        bool first = true;
        double* fwDev = dalloc<double>(std::max(g.nf, 1));
        for (int pass = 0; pass < knob("TPP_PASSES", 2); pass++) {
            h2d(ctx, fwDev, cur.fw.data(), cur.nf * sizeof(double));
            LV M;
            std::vector<int> rs, cf, cn;
            int *drs = nullptr, *dcf = nullptr, *dcn = nullptr;
            if (first && fineEll) M = fineView(nullptr, nullptr);
            else {
                csrOf(cur, rs, cf, cn);
                drs = upNew(ctx, rs); dcf = upNew(ctx, cf); dcn = upNew(ctx, cn);
                memset(&M, 0, sizeof(M));
                M.n = cur.n; M.nf = cur.nf; M.ell = 0; M.rs = drs; M.cf = dcf; M.cn = dcn;
            }
            std::vector<int> rootH;
            matchPass(M, cur.n, fwDev, rootH);
            dev_sync(ctx);
            dev_free(drs); dev_free(dcf); dev_free(dcn);
            std::vector<int> rank(cur.n, -1), agg(cur.n);
            int nc = 0;
            for (int i = 0; i < cur.n; i++) if (rootH[i] == i) rank[i] = nc++;
            for (int i = 0; i < cur.n; i++) agg[i] = rank[rootH[i]];
            std::vector<int> ghostAgg(cur.nG(), 0);
            if (dist && comm.active) {
                std::vector<double> snd(cur.nG()), rcv;
                for (int j = 0; j < cur.nG(); j++) snd[j] = (double)agg[cur.own[cur.nfLoc + j]];
                hostExchange(cur.poff, cur.pcnt, cur.ppeer, snd, rcv);
                for (int j = 0; j < cur.nG(); j++) ghostAgg[j] = (int)(rcv[j] + 0.5);
            }
            std::vector<int> sS, sF;
            coarsen(cur, agg, nc, ghostAgg, nxt, sS, sF);
            nxt.nGlob = dist ? globalCount((double)nxt.n) : (double)nxt.n;
            if (knob("TPP_VERBOSE", 0)) fprintf(stderr, "amg pass%s: n %d nf %d (+%d proc) -> n %d nf %d (+%d proc), avg degree %.1f, global rows %.0f\n", dist ? "" : " (tail)", cur.n, cur.nfLoc, cur.nG(), nxt.n, nxt.nfLoc, nxt.nG(), 2.0 * nxt.nfLoc / std::max(nxt.n, 1), nxt.nGlob);
            if (first) { aggTot = agg; segS = sS; segF = sF; first = false; }
            else {
                for (auto& a : aggTot) a = agg[a];
                // flatten: coarse face -> intermediate faces -> fine faces
                std::vector<int> nS(1, 0), nF2;
                for (int F = 0; F < nxt.nf; F++) {
                    for (int k = sS[F]; k < sS[F + 1]; k++) {
                        int mid = sF[k];
                        for (int q = segS[mid]; q < segS[mid + 1]; q++) nF2.push_back(segF[q]);
                    }
                    nS.push_back((int)nF2.size());
                }
                segS.swap(nS); segF.swap(nF2);
            }
            cur = nxt;
            if (cur.nGlob <= stopRows) break;
        }
        dev_free(fwDev);
        if (cur.nGlob >= 0.9 * g.nGlob) return false;  // stalled
        v = Level();
        v.n = cur.n; v.nf = cur.nf; v.nfLoc = cur.nfLoc; v.nG = cur.nG(); v.nGlob = cur.nGlob;
        v.poff = cur.poff; v.pcnt = cur.pcnt; v.ppeer = cur.ppeer;
        std::vector<int> rs, cf, cn;
        csrOf(cur, rs, cf, cn);
        v.nnz = rs[cur.n];
        v.rs = upNew(ctx, rs); v.cf = upNew(ctx, cf); v.cn = upNew(ctx, cn); v.own = upNew(ctx, cur.own); v.nei = upNew(ctx, cur.nei);
        if (knob("TPP_ELLC", 1)) buildEllc(v, rs, cn, knob("TPP_ELLC_OV", dist ? 25 : 12));
        v.agg = upNew(ctx, aggTot); v.segStart = upNew(ctx, segS); v.segFaces = upNew(ctx, segF);
        std::vector<int> howner(cur.own.begin() + cur.nfLoc, cur.own.end());
        v.dOwner = upNew(ctx, howner);
        // members of each aggregate, ascending fine index
        std::vector<int> aS(cur.n + 1, 0), aR(g.n);
        for (int i = 0; i < g.n; i++) aS[aggTot[i] + 1]++;
        for (int c = 0; c < cur.n; c++) aS[c + 1] += aS[c];
        std::vector<int> pos(aS.begin(), aS.end() - 1);
        for (int i = 0; i < g.n; i++) aR[pos[aggTot[i]]++] = i;
        v.aggStart = upNew(ctx, aS); v.aggRows = upNew(ctx, aR);
        v.ev = dalloc<double>(std::max(v.nnz, 1));
        v.diag = dalloc<double>(v.n); v.upper = dalloc<double>(std::max(v.nf, 1)); v.rsum = dalloc<double>(v.n);
        g = cur;
        return true;
    }
    // the gather level's graph over all ranks: rows and faces rank by rank (a rank's local
    // faces, then its processor faces towards higher ranks)
    void gatherGraph(const HostGraph& g, HostGraph& G) {
        const int me = comm.active ? comm.rank : 0, world = comm.active ? comm.size : 1;
        std::vector<double> cnt(3 * (size_t)world, 0.0);
        int nUp = 0;
        for (size_t p = 0; p < g.pcnt.size(); p++) if (g.ppeer[p] > me) nUp += g.pcnt[p];
        cnt[3 * me] = g.n; cnt[3 * me + 1] = g.nfLoc; cnt[3 * me + 2] = nUp;
        hostAllreduce(cnt, 0);
        std::vector<int> rowOff(world + 1, 0), faceOff(world + 1, 0);
        for (int r = 0; r < world; r++) {
            rowOff[r + 1] = rowOff[r] + (int)(cnt[3 * r] + 0.5);
            faceOff[r + 1] = faceOff[r] + (int)(cnt[3 * r + 1] + 0.5) + (int)(cnt[3 * r + 2] + 0.5);
        }
        G = HostGraph();
        G.n = rowOff[world]; G.nf = G.nfLoc = faceOff[world]; G.nGlob = G.n;
        std::vector<double> own(G.nf, 0.0), nei(G.nf, 0.0), fw(G.nf, 0.0);
        tailRowOff = rowOff[me]; tailFaceOff = faceOff[me];
#ifndef TPP_EMU
        comm.gRowOff = rowOff;
#endif
        tailCopy.clear();
        int at = faceOff[me];
        for (int f = 0; f < g.nfLoc; f++, at++) { own[at] = g.own[f] + rowOff[me]; nei[at] = g.nei[f] + rowOff[me]; fw[at] = g.fw[f]; }
        for (size_t p = 0; p < g.pcnt.size(); p++) {
            if (g.ppeer[p] <= me || g.pcnt[p] == 0) continue;
            tailCopy.push_back({g.nfLoc + g.poff[p], g.pcnt[p], at});
            for (int k = 0; k < g.pcnt[p]; k++, at++) {
                int j = g.poff[p] + k, f = g.nfLoc + j;
                own[at] = g.own[f] + rowOff[me];
                nei[at] = g.ghostRow[j] + rowOff[g.ppeer[p]];
                fw[at] = g.fw[f];
            }
        }
        hostAllreduce(own, 0); hostAllreduce(nei, 0); hostAllreduce(fw, 0);
        G.own.resize(G.nf); G.nei.resize(G.nf); G.fw = fw;
        for (int f = 0; f < G.nf; f++) { G.own[f] = (int)(own[f] + 0.5); G.nei[f] = (int)(nei[f] + 0.5); }
    }

    void buildAMG() {
        amgBuilt = true;
        const int coarsestTarget = knob("TPP_COARSEST", 1500), maxLevels = 24;
        const int tailRows = std::max(knob("TPP_TAIL_ROWS", 60000), coarsestTarget);
        if (nGlobal <= coarsestTarget) return;
        // faceAreaPair weights |Sf/sqrt(|Sf|) * (1, 1.01, 1.02)|
        HostGraph g;
        g.n = nC; g.nf = nI; g.nfLoc = nIloc; g.nGlob = (double)nGlobal;
        g.own.assign(own.begin(), own.begin() + nI); g.nei.assign(nei.begin(), nei.begin() + nI); g.fw.resize(nI);
        g.poff = procOff; g.pcnt = procCnt; g.ppeer = procPeer;
        for (int f = 0; f < nI; f++) {
            double s = sqrt(magSf[f]);
            double v[3] = {Sf0[3 * f] / s * 1.0, Sf0[3 * f + 1] / s * 1.01, Sf0[3 * f + 2] / s * 1.02};
            g.fw[f] = mag3(v);
        }
        // distributed levels, down to the first one small enough to be gathered onto every rank
        while ((int)levels.size() < maxLevels) {
            Level v;
            if (!makeLevel(g, levels.empty(), true, coarsestTarget, v)) break;
            levels.push_back(v);
            if (g.nGlob <= tailRows) break;
        }
        if (levels.empty()) return;
        gatherLevel = (int)levels.size() - 1;
        HostGraph G;
        gatherGraph(g, G);
        {
            Level t;
            t.n = G.n; t.nf = t.nfLoc = G.nf; t.nG = 0; t.nGlob = G.n;
            std::vector<int> rs, cf, cn;
            csrOf(G, rs, cf, cn);
            t.nnz = rs[G.n];
            t.rs = upNew(ctx, rs); t.cf = upNew(ctx, cf); t.cn = upNew(ctx, cn); t.own = upNew(ctx, G.own); t.nei = upNew(ctx, G.nei);
            t.ev = dalloc<double>(std::max(t.nnz, 1));
            t.diag = dalloc<double>(t.n); t.upper = dalloc<double>(std::max(t.nf, 1)); t.rsum = dalloc<double>(t.n);
            if (knob("TPP_ELLC", 1)) buildEllc(t, rs, cn, 12);
            tail.push_back(t);
        }
        while ((int)tail.size() < TAIL_MAXLV && G.n > coarsestTarget) {
            Level v;
            if (!makeLevel(G, false, false, coarsestTarget, v)) break;
            tail.push_back(v);
        }
        setupGatherWindow();
        tailBar = (unsigned*)dev_alloc(64);
        tailErr = (int*)dev_alloc(64);
#ifndef TPP_EMU
        {
            int dev = 0, sms = 0, perSm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            // the coarsest level staged in the CTA's shared memory in ELL form (tail_coarse_cg): the
            // smallest width (multiple of 4) that leaves at most 2 % of the entries to the CSR overflow
            // and fits 200 KB together with the diagonal and the four CG vectors
            const size_t rb = knob("TPP_FP32", 1) ? 4 : 8, cn_ = (size_t)tail.back().n;
            tailSmem = 0; tailCgW = 0;
            if (cn_ < 65536 && knob("TPP_CG_SMEM", 1)) {
                std::vector<int> rsH(cn_ + 1);
                d2h(ctx, rsH.data(), tail.back().rs, (cn_ + 1) * sizeof(int));
                for (int w = 4; w <= 64; w += 4) {
                    long ov = 0;
                    for (size_t i = 0; i < cn_; i++) ov += std::max(0, rsH[i + 1] - rsH[i] - w);
                    const size_t need_ = cn_ * w * (rb + 2) + 5 * cn_ * rb + 64;
                    if (need_ > 200 * 1024) break;
                    tailCgW = w; tailSmem = need_;
                    if (ov * 50 <= (long)rsH[cn_]) break;
                }
            }
            if (knob("TPP_FP32", 1)) {
                if (tailSmem) CUDA_CHECK(cudaFuncSetAttribute(vk_tail<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tailSmem));
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, vk_tail<float>, TAIL_THREADS, tailSmem);
            } else {
                if (tailSmem) CUDA_CHECK(cudaFuncSetAttribute(vk_tail<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tailSmem));
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, vk_tail<double>, TAIL_THREADS, tailSmem);
            }
            tailGrid = sms;  // one CTA per SM
            if (perSm < 1) throw tpp::CudaFailure{"vk_tail does not fit an SM"};
            // small tails do not need every SM: fewer CTAs make the grid barrier cheaper
            int need = (tail[0].n * (tail[0].ellW > 0 ? 1 : 4) + TAIL_THREADS - 1) / TAIL_THREADS;
            tailGrid = std::max(1, std::min(tailGrid, need));
            if (knob("TPP_TAIL_CTAS", 0) > 0) tailGrid = std::min(tailGrid, knob("TPP_TAIL_CTAS", 0));
        }
#endif
    }

    // Galerkin coefficients of every level from the fine matrix
    void galerkin(LV& F0) {
        LAUNCH(ctx, rowsum, F0, F0.n);
        LAUNCH(ctx, fill_ev, F0, F0.n);
        LV F = F0;
        for (int l = 0; l <= gatherLevel; l++) {
            LV L = viewOf(levels[l]);
            L.fupper = F.upper; L.frsum = F.rsum;
            LAUNCH(ctx, coarse_upper, L, L.nf);
            LAUNCH(ctx, coarse_diag, L, L.n);
            if (l < gatherLevel) LAUNCH(ctx, fill_ev, L, L.n);
            F = L;
        }
        if (tail.empty()) return;
        // gather: every rank's slice of the gather level's diagonal and face coefficients
        Level& S = levels[gatherLevel];
        Level& T0 = tail[0];
        if (comm.active) { dev_zero(ctx, T0.diag, T0.n * sizeof(double)); dev_zero(ctx, T0.upper, T0.nf * sizeof(double)); }
        d2d(ctx, T0.diag + tailRowOff, S.diag, S.n * sizeof(double));
        if (S.nfLoc) d2d(ctx, T0.upper + tailFaceOff, S.upper, S.nfLoc * sizeof(double));
        for (auto& c : tailCopy) d2d(ctx, T0.upper + c[2], S.upper + c[0], c[1] * sizeof(double));
        allreduceDev(T0.diag, (size_t)T0.n);
        allreduceDev(T0.upper, (size_t)T0.nf);
        LV L0 = viewOf(T0);
        LAUNCH(ctx, rowsum, L0, L0.n);
        LAUNCH(ctx, fill_ev, L0, L0.n);
        F = L0;
        for (size_t t = 1; t < tail.size(); t++) {
            LV L = viewOf(tail[t]);
            L.fupper = F.upper; L.frsum = F.rsum;
            LAUNCH(ctx, coarse_upper, L, L.nf);
            LAUNCH(ctx, coarse_diag, L, L.n);
            LAUNCH(ctx, fill_ev, L, L.n);
            F = L;
        }
    }

    // ---- V-cycle in precision R (tpp_vcycle.h): per-level storage, conversion, cycle, preconditioner
    template <class R> struct VStore {
        std::vector<R*> diag, ev, x, b, t0, r, Ac, send;  // index 0 = fine level, 1.. = distributed coarse levels
        std::vector<R*> eev, oev;                         // ELL + overflow values of the coarse levels
        std::vector<R*> tdiag, tev, tx, ty, tb, tr, teev, toev;  // tail levels
        R *cgR = nullptr, *cgP = nullptr, *cgAp = nullptr;
        bool ready = false;
    };
    VStore<float> vsF;
    VStore<double> vsD;
    template <class R> VStore<R>& vstore();
    int vLevels() const { return 1 + std::max(gatherLevel, 0); }  // levels smoothed kernel by kernel
    int vRows(int lv) const { return lv == 0 ? nC : levels[lv - 1].n; }
    int vGhosts(int lv) const { return lv == 0 ? nG : levels[lv - 1].nG; }
    size_t vEntries(int lv) const { return lv == 0 ? (size_t)W * nCp : (size_t)std::max(levels[lv - 1].nnz, 1); }
    template <class R> void ensureVStore() {
        VStore<R>& v = vstore<R>();
        if (v.ready) return;
        for (int lv = 0; lv < vLevels(); lv++) {
            size_t n = (size_t)vRows(lv) + (size_t)vGhosts(lv);
            v.diag.push_back(dalloc<R>(n)); v.ev.push_back(dalloc<R>(vEntries(lv)));
            v.x.push_back(dalloc<R>(n)); v.b.push_back(dalloc<R>(n)); v.t0.push_back(dalloc<R>(n)); v.r.push_back(dalloc<R>(n)); v.Ac.push_back(dalloc<R>(n));
            v.send.push_back(dalloc<R>(std::max(vGhosts(lv), 1)));
            const bool ellc = lv > 0 && levels[lv - 1].ellW > 0;
            v.eev.push_back(dalloc<R>(ellc ? (size_t)levels[lv - 1].ellW * levels[lv - 1].nPad : 1));
            v.oev.push_back(dalloc<R>(ellc ? (size_t)std::max(levels[lv - 1].nOv, 1) : 1));
        }
        for (auto* vec : {&v.eev, &v.oev}) for (R* p : *vec) allocs.push_back(p);
        for (auto* vec : {&v.diag, &v.ev, &v.x, &v.b, &v.t0, &v.r, &v.Ac, &v.send}) for (R* p : *vec) allocs.push_back(p);
        for (size_t t = 0; t < tail.size(); t++) {
            size_t n = (size_t)tail[t].n;
            v.tdiag.push_back(dalloc<R>(n)); v.tev.push_back(dalloc<R>(std::max(tail[t].nnz, 1)));
            v.tx.push_back(dalloc<R>(n)); v.ty.push_back(dalloc<R>(n)); v.tb.push_back(dalloc<R>(n)); v.tr.push_back(dalloc<R>(n));
            v.teev.push_back(dalloc<R>(tail[t].ellW > 0 ? (size_t)tail[t].ellW * tail[t].nPad : 1));
            v.toev.push_back(dalloc<R>(std::max(tail[t].nOv, 1)));
        }
        for (auto* vec : {&v.tdiag, &v.tev, &v.tx, &v.ty, &v.tb, &v.tr, &v.teev, &v.toev}) for (R* p : *vec) allocs.push_back(p);
        if (!tail.empty()) {
            size_t n = (size_t)tail.back().n;
            v.cgR = dalloc<R>(n); v.cgP = dalloc<R>(n); v.cgAp = dalloc<R>(n);
            allocs.push_back(v.cgR); allocs.push_back(v.cgP); allocs.push_back(v.cgAp);
        }
        v.ready = true;
    }
    template <class R> VL<R> vview(int lv) {
        VStore<R>& v = vstore<R>();
        VL<R> L;
        memset(&L, 0, sizeof(L));
        if (lv == 0) { L.n = nC; L.nf = nI; L.nCp = nCp; L.W = W; L.ell = 1; L.cn = d.cn; L.nOwn = levels.empty() ? nC : nC + nG; }
        else { Level& c = levels[lv - 1]; L.n = c.n; L.nf = c.nf; L.ell = 0; L.cn = c.cn; L.rs = c.rs; L.nOwn = c.n + c.nG; L.agg = c.agg; L.aggStart = c.aggStart; L.aggRows = c.aggRows; }
        if (lv < vLevels()) { L.diag = v.diag[lv]; L.ev = v.ev[lv]; }
        if (lv > 0 && lv < vLevels() && levels[lv - 1].ellW > 0) {
            Level& c = levels[lv - 1];
            L.ellW = c.ellW; L.nPad = c.nPad; L.ecn = c.ecn; L.ors = c.ors; L.ocn = c.ocn; L.eev = v.eev[lv]; L.oev = v.oev[lv];
        }
        return L;
    }
    // matrix values of every level in precision R (after galerkin(), once per solve)
    template <class R> void convertLevels(LV& F0) {
        VStore<R>& v = vstore<R>();
        CastArgs<R> a;
        memset(&a, 0, sizeof(a));
        for (int lv = 0; lv < vLevels(); lv++) {
            a.src = lv == 0 ? F0.diag : levels[lv - 1].diag; a.dst = v.diag[lv];
            VLAUNCH(ctx, cast_in, a, vRows(lv));
            a.src = lv == 0 ? F0.ev : levels[lv - 1].ev; a.dst = v.ev[lv];
            VLAUNCH(ctx, cast_in, a, (int)vEntries(lv));
#ifndef TPP_EMU
            if (lv > 0 && levels[lv - 1].ellW > 0) {
                Level& c = levels[lv - 1];
                GatherArgs<R> ga{c.ev, c.esrc, v.eev[lv]};
                const int ne = c.ellW * c.nPad;
                vk_cast_gather<R><<<(ne + 255) / 256, 256, 0, ctx.stream>>>(ga, ne);
                if (c.nOv > 0) { GatherArgs<R> go{c.ev, c.osrc, v.oev[lv]}; vk_cast_gather<R><<<(c.nOv + 255) / 256, 256, 0, ctx.stream>>>(go, c.nOv); }
                ctx.launches += 2;
            }
#endif
        }
        for (size_t t = 0; t < tail.size(); t++) {
            a.src = tail[t].diag; a.dst = v.tdiag[t];
            VLAUNCH(ctx, cast_in, a, tail[t].n);
            a.src = tail[t].ev; a.dst = v.tev[t];
            VLAUNCH(ctx, cast_in, a, std::max(tail[t].nnz, 1));
#ifndef TPP_EMU
            if (tail[t].ellW > 0) {
                Level& c = tail[t];
                GatherArgs<R> ga{c.ev, c.esrc, v.teev[t]};
                const int ne = c.ellW * c.nPad;
                vk_cast_gather<R><<<(ne + 255) / 256, 256, 0, ctx.stream>>>(ga, ne);
                if (c.nOv > 0) { GatherArgs<R> go{c.ev, c.osrc, v.toev[t]}; vk_cast_gather<R><<<(c.nOv + 255) / 256, 256, 0, ctx.stream>>>(go, c.nOv); }
                ctx.launches += 2;
            }
#endif
        }
    }
    // halo exchange of a V-cycle vector on level lv (0 = mesh): owned rows behind my processor
    // faces -> the neighbours' ghost rows [n, n + nG)
    template <class R> void XL(int lv, R* vec) {
        const int ng = vGhosts(lv);
        if (!comm.active || ng == 0 || levels.empty()) return;
        VStore<R>& v = vstore<R>();
        const std::vector<int>& off = lv == 0 ? procOff : levels[lv - 1].poff;
        const std::vector<int>& cnt = lv == 0 ? procCnt : levels[lv - 1].pcnt;
        const std::vector<int>& peer = lv == 0 ? procPeer : levels[lv - 1].ppeer;
#ifndef TPP_EMU
        if (comm.p2p) { p2pExchange<R>(lv == 0 ? dProcOwner : levels[lv - 1].dOwner, off, ng, vec, vec + vRows(lv), 1); return; }
#endif
        PackArgs<R> a;
        a.owner = lv == 0 ? dProcOwner : levels[lv - 1].dOwner; a.src = vec; a.dst = v.send[lv];
        VLAUNCH(ctx, pack, a, ng);
        R* ghost = vec + vRows(lv);
#ifndef TPP_EMU
        if (comm.nccl) {
            prof_begin(ctx, "v_halo_sendrecv");
            comm.pGroupStart();
            for (size_t i = 0; i < cnt.size(); i++) {
                if (cnt[i] == 0) continue;
                comm.pSend(v.send[lv] + off[i], (size_t)cnt[i], sizeof(R) == 4 ? ncclFloat : ncclDouble, peer[i], comm.nccl, ctx.stream);
                comm.pRecv(ghost + off[i], (size_t)cnt[i], sizeof(R) == 4 ? ncclFloat : ncclDouble, peer[i], comm.nccl, ctx.stream);
            }
            comm.pGroupEnd();
            prof_end(ctx);
            ctx.launches++;
            return;
        }
#endif
        std::vector<R> hs(ng), hr(ng);
        d2h(ctx, hs.data(), v.send[lv], ng * sizeof(R));
        std::vector<double> s(hs.begin(), hs.end()), r;
        hostExchange(off, cnt, peer, s, r);
        for (int j = 0; j < ng; j++) hr[j] = (R)r[j];
        h2d(ctx, ghost, hr.data(), ng * sizeof(R));
    }
    // ghost rows for a smoothing sweep that is not the first of its group.  Exact (default): a halo
    // exchange.  TPP_LAG bit 0: the second pre-sweep from a zero guess sees zero ghosts; bit 1: a
    // later post-sweep reuses the ghost values of the sweep before it (`prev`) - the smoother
    // becomes block-Jacobi-like across rank interfaces for those sweeps only; residuals, the
    // correction's A c and the first post-sweep always see exchanged values.
    template <class R> void XLsmooth(int lv, R* vec, int kind, const R* prev = nullptr) {
        const int ng = vGhosts(lv), lag = knob("TPP_LAG", 3);
        if (!comm.active || ng == 0 || levels.empty()) return;
        if (kind == 1 && (lag & 1)) { dev_zero(ctx, vec + vRows(lv), ng * sizeof(R)); ctx.launches++; return; }
        if (kind == 2 && (lag & 2) && prev) { d2d(ctx, vec + vRows(lv), prev + vRows(lv), ng * sizeof(R)); ctx.launches++; return; }
        XL<R>(lv, vec);
    }
    // 0 Jacobi sweep, 1 residual, 2 the first two sweeps from a zero guess in one pass, 3 prolongation
    // + over-correction + first post-sweep in one pass
    template <class R> void vRowOp(VL<R>& L, int mode) {
        static const char* ellName[4] = {"v_jacobi", "v_residual", "v_jacobi_first", "v_jacobi_corr"};
        static const char* csrName[4] = {"v_jacobi_csr", "v_residual_csr", "v_jacobi_first_csr", "v_jacobi_corr_csr"};
#ifndef TPP_EMU
        if (L.ell && (L.W == 4 || L.W == 6) && knob("TPP_ELL2", 1)) {
            prof_begin(ctx, ellName[mode]);
            const int g_ = ((L.n + 1) / 2 + 255) / 256;
#define ELL2_CASE(WW) switch (mode) { case 0: vk_ell2_row_op<R, WW, 0><<<g_, 256, 0, ctx.stream>>>(L); break; case 1: vk_ell2_row_op<R, WW, 1><<<g_, 256, 0, ctx.stream>>>(L); break; \
                                      case 2: vk_ell2_row_op<R, WW, 2><<<g_, 256, 0, ctx.stream>>>(L); break; default: vk_ell2_row_op<R, WW, 3><<<g_, 256, 0, ctx.stream>>>(L); }
            if (L.W == 4) { ELL2_CASE(4) } else { ELL2_CASE(6) }
#undef ELL2_CASE
            LAUNCH_CHECK("vk_ell2_row_op");
            prof_end(ctx);
            ctx.launches++;
            return;
        }
        if (!L.ell && L.ellW > 0) {
            prof_begin(ctx, csrName[mode]);
            const int g_ = (L.n + 255) / 256;
            switch (L.ellW) {
                case 6: vk_ellc_row_op<R, 6><<<g_, 256, 0, ctx.stream>>>(L, mode); break;
                case 8: vk_ellc_row_op<R, 8><<<g_, 256, 0, ctx.stream>>>(L, mode); break;
                case 12: vk_ellc_row_op<R, 12><<<g_, 256, 0, ctx.stream>>>(L, mode); break;
                default: vk_ellc_row_op<R, 16><<<g_, 256, 0, ctx.stream>>>(L, mode); break;
            }
            LAUNCH_CHECK("vk_ellc_row_op");
            prof_end(ctx);
            ctx.launches++;
            return;
        }
        if (!L.ell) {
            prof_begin(ctx, csrName[mode]);
            if (2 * (long)L.nf <= 10 * (long)L.n) vk_csr_row_op<R, 4><<<(L.n * 4 + 255) / 256, 256, 0, ctx.stream>>>(L, mode);
            else vk_csr_row_op<R, 8><<<(L.n * 8 + 255) / 256, 256, 0, ctx.stream>>>(L, mode);
            LAUNCH_CHECK("vk_csr_row_op");
            prof_end(ctx);
            ctx.launches++;
            return;
        }
#endif
        if (mode == 0) VLAUNCH(ctx, jacobi, L, L.n);
        else if (mode == 1) VLAUNCH(ctx, residual, L, L.n);
        else if (mode == 2) VLAUNCH(ctx, jacobi_first, L, L.n);
        else VLAUNCH(ctx, jacobi_corr, L, L.n);
    }
    // out = A in ; scal[S_TMP0] = r.in ; scal[S_TMP1] = in.out  (summed over the ranks)
    template <class R> void vSpmvDot2(VL<R>& L) {
#ifdef TPP_EMU
        double v = 0, w = 0;
        for (int c = 0; c < L.n; c++) { R y = vl_Ax(L, c, L.in); L.out[c] = y; v += (double)L.r[c] * (double)L.in[c]; w += (double)y * (double)L.in[c]; }
        scal[S_TMP0] = v; scal[S_TMP1] = w;
        if (knob("TPP_VERBOSE", 0) >= 2) fprintf(stderr, "sf n=%d %.4f\n", L.n, v / w);
#else
        prof_begin(ctx, L.ell ? "v_spmv_dot2" : "v_spmv_dot2_csr");
        int nb = std::min(RED_BLOCKS, (L.n + BLOCK - 1) / BLOCK);
        if (L.ell && (L.W == 4 || L.W == 6) && knob("TPP_ELL2", 1)) {
            nb = std::min(4 * RED_BLOCKS, ((L.n + 1) / 2 + BLOCK - 1) / BLOCK);  // row pairs: more CTAs than the streaming kernels need
            if (L.W == 4) vk_ell2_spmv_dot2<R, 4><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial, red.partial2);
            else vk_ell2_spmv_dot2<R, 6><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial, red.partial2);
        } else if (L.ell) vk_spmv_dot2<R><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial, red.partial2);
        else if (L.ellW > 0) {
            switch (L.ellW) {
                case 6: vk_ellc_spmv_dot2<R, 6><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial, red.partial2); break;
                case 8: vk_ellc_spmv_dot2<R, 8><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial, red.partial2); break;
                case 12: vk_ellc_spmv_dot2<R, 12><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial, red.partial2); break;
                default: vk_ellc_spmv_dot2<R, 16><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial, red.partial2); break;
            }
        } else {
            const bool shortRows = 2 * (long)L.nf <= 10 * (long)L.n;
            nb = std::min(RED_BLOCKS, (L.n * (shortRows ? 4 : 8) + 255) / 256);
            if (shortRows) vk_csr_spmv_dot2<R, 4><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial, red.partial2);
            else vk_csr_spmv_dot2<R, 8><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial, red.partial2);
        }
        k_reduce_final2<<<2, BLOCK, 0, ctx.stream>>>(red.partial, red.partial2, nb, scal + S_TMP0, scal + S_TMP1);
        prof_end(ctx);
#endif
        ctx.launches += 2;
        allreduce(S_TMP0, 2, 0);
    }
    template <class R> void vCoarseSolve(VL<R>& L, R* r, R* p, R* Ap) {  // L.b -> L.out (a mesh without coarse levels)
        const int maxIt = knob("TPP_CITER", 8);
        const double tol = knobd("TPP_CTOL", 0.05);
#ifdef TPP_EMU
        int n = L.n;
        R* x = L.out;
        double rz = 0;
        for (int i = 0; i < n; i++) { x[i] = 0; r[i] = L.b[i]; p[i] = L.b[i] / L.diag[i]; rz += (double)L.b[i] * (double)p[i]; }
        double rz0 = rz;
        if (rz > 0)
            for (int it = 0; it < maxIt; it++) {
                double pAp = 0;
                for (int i = 0; i < n; i++) { Ap[i] = vl_Ax(L, i, (const R*)p); pAp += (double)Ap[i] * (double)p[i]; }
                R alpha = (R)(rz / pAp);
                double rzn = 0;
                for (int i = 0; i < n; i++) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; rzn += (double)r[i] * (double)r[i] / (double)L.diag[i]; }
                if (rzn <= tol * tol * rz0) break;
                R beta = (R)(rzn / rz);
                rz = rzn;
                for (int i = 0; i < n; i++) p[i] = r[i] / L.diag[i] + beta * p[i];
            }
#else
        prof_begin(ctx, "v_coarse_cg");
        vk_coarse_cg<R><<<1, 1024, 0, ctx.stream>>>(L, r, p, Ap, maxIt, tol);
        prof_end(ctx);
#endif
        ctx.launches++;
    }
    // the tail sub-cycle: tb[0] (restricted residual, gathered) -> tx[0]
    template <class R> void runTail(int nPre, int nPost) {
        VStore<R>& v = vstore<R>();
        TailArgs<R> A;
        memset(&A, 0, sizeof(A));
        A.T = (int)tail.size();
        for (int t = 0; t < A.T; t++) {
            Level& c = tail[t];
            TLv<R>& L = A.lv[t];
            L.n = c.n;
            double deg = (double)c.nnz / std::max(c.n, 1);
            L.coop = deg <= 5 ? 4 : deg <= 10 ? 8 : 16;
            L.rs = c.rs; L.cn = c.cn; L.ev = v.tev[t]; L.diag = v.tdiag[t];
            if (c.ellW > 0 && t + 1 < A.T && knob("TPP_TAIL_ELL", 1)) { L.ellW = c.ellW; L.nPad = c.nPad; L.ecn = c.ecn; L.ors = c.ors; L.ocn = c.ocn; L.eev = v.teev[t]; L.oev = v.toev[t]; }
            if (t + 1 < A.T) L.agg = tail[t + 1].agg;
            if (t > 0) { L.aggStart = c.aggStart; L.aggRows = c.aggRows; }
            L.x = v.tx[t]; L.y = v.ty[t]; L.b = v.tb[t]; L.r = v.tr[t];
        }
        A.bar = tailBar; A.err = tailErr;
        A.overcorr = (R)knobd("TPP_TAIL_OVERCORR", 1.8);
        // 8 CG iterations on the coarsest level give the same PCG counts as 16; fewer sweeps on the
        // small levels cost 1-2 PCG iterations at 6 M cells (tools/knob_sweep.py), so they keep nPre/nPost
        A.nPre = std::min(knob("TPP_TAIL_NPRE", nPre), TAIL_MAXSW); A.nPost = std::min(knob("TPP_TAIL_NPOST", nPost), TAIL_MAXSW);
        for (int k = 0; k < TAIL_MAXSW; k++) { A.omPre[k] = (R)smootherOmega(k, std::max(A.nPre, 1)); A.omPost[k] = (R)smootherOmega(k, std::max(A.nPost, 1)); }
        A.cgIter = knob("TPP_CITER", 8); A.cgTol = knobd("TPP_CTOL", 0.05);
        A.cgR = v.cgR; A.cgP = v.cgP; A.cgAp = v.cgAp; A.cgSmem = tailSmem > 0 ? tailCgW : 0;
        A.cgDeflate = knob("TPP_DEFLATE", d.needRef ? 1 : 0);  // closed domains only: an open boundary pins the level
        prof_begin(ctx, "v_tail");
#ifdef TPP_EMU
        tail_host(A);
#else
        CUDA_CHECK(cudaMemsetAsync(tailBar, 0, sizeof(unsigned), ctx.stream));
        void* args[] = {&A};
        if (knob("TPP_TAIL_COOP", 1)) CUDA_CHECK(cudaLaunchCooperativeKernel((void*)vk_tail<R>, dim3(tailGrid), dim3(TAIL_THREADS), args, tailSmem, ctx.stream));
        else vk_tail<R><<<tailGrid, TAIL_THREADS, tailSmem, ctx.stream>>>(A);
#endif
        prof_end(ctx);
        ctx.launches += 2;
    }
    // x ~= A^-1 b on level lv (0 = fine); x, b are the level's own buffers unless given
    template <class R> void vcycleT(int lv, const R* b, R* x, bool zeroGuess, int nPre, int nPost) {
        VStore<R>& v = vstore<R>();
        VL<R> L = vview<R>(lv);
        const R omega = (R)knobd("TPP_OMEGA", 0.8);
        if (levels.empty()) {  // a mesh too small for a hierarchy: one-CTA CG on the rank's own rows
            L.b = b; L.out = x;
            vCoarseSolve(L, v.r[lv], v.t0[lv], v.Ac[lv]);
            return;
        }
        L.omega = omega; L.b = b;
        R *cur = x, *oth = v.t0[lv];
        const bool ghosts = comm.active && vGhosts(lv) > 0;
        const int fuse = knob("TPP_FUSE", 3);
        // the first iterate om0 b/diag is not stored when a second sweep follows: that sweep forms it per
        // column (one pass and one launch less); rows of other ranks count as zero there, which is what
        // TPP_LAG bit 0 does to the second sweep anyway
        const bool fuseFirst = zeroGuess && nPre >= 2 && (fuse & 1) && (!ghosts || (knob("TPP_LAG", 3) & 1));
        for (int s = 0; s < std::max(nPre, 1); s++) {
            L.omega = (R)smootherOmega(s, std::max(nPre, 1));
            if (s == 0 && zeroGuess) { if (fuseFirst) continue; L.out = cur; VLAUNCH(ctx, jacobi0, L, L.n); }
            else if (s == 1 && fuseFirst) { L.om0 = (R)smootherOmega(0, std::max(nPre, 1)); L.in = nullptr; L.out = cur; vRowOp(L, 2); }
            else { XLsmooth<R>(lv, cur, s == 1 && zeroGuess ? 1 : 0); L.in = cur; L.out = oth; vRowOp(L, 0); std::swap(cur, oth); }
        }
        XL<R>(lv, cur);
        L.in = cur; L.out = v.r[lv];
        vRowOp(L, 1);
        const bool toTail = lv + 1 == vLevels();
        VL<R> Cn = vview<R>(lv + 1);
        Cn.r = v.r[lv];
        if (toTail) {
            if (comm.active) dev_zero(ctx, v.tb[0], tail[0].n * sizeof(R));
            Cn.out = v.tb[0] + tailRowOff;
        } else Cn.out = v.b[lv + 1];
        VLAUNCH(ctx, restrict, Cn, Cn.n);
        if (toTail) {
            gatherTail<R>(v.tb[0]);
            runTail<R>(nPre, nPost);
        } else vcycleT<R>(lv + 1, v.b[lv + 1], v.x[lv + 1], true, nPre, nPost);
        // prolonged correction c = P x_c in `oth`, A c, scaling, x += ...
        VL<R> Pn = vview<R>(lv + 1);
        Pn.xc = toTail ? v.tx[0] + tailRowOff : v.x[lv + 1];
        // Coarse correction with a FIXED over-correction factor (Braess' remedy for the constant
        // interpolation of plain aggregation) instead of GAMG's energy-minimising scaleCorrection:
        // no A c product, no dot products and - on several GPUs - no halo exchange / all-reduce for them
        // on every level of every cycle.  Measured on the 6.2 M-cell tank (B200, 12 steps, p_rghFinal
        // iterations at the end): scaled everywhere 11-12; 1.6 on the kernel levels + 1.8 in the tail
        // 10-12 (also with five kernel levels, TPP_TAIL_ROWS=5000: 11-12 vs 12-14 scaled); 1.8 everywhere
        // does not converge (nested over-corrections compound), nor does 2.2 on a single level.  A solve
        // that nevertheless stops at maxIter switches the handle back to the scaled correction
        // (`scaledFallback`).  TPP_NOSCALE_FROM=99 selects the scaled correction outright.
        bool corrFused = false;
        const int noScaleFrom = scaledFallback ? 99 : knob("TPP_NOSCALE_FROM", 0);
        if (lv >= noScaleFrom && !ghosts && (fuse & 2)) {
            // prolongation, over-correction and the first post-sweep in one pass over the matrix
            corrFused = true;
        } else if (lv >= noScaleFrom) {
            Pn.out = cur; Pn.omega = (R)(lv == 0 ? knobd("TPP_OVERCORR0", overcorrKernel()) : lv == 1 ? knobd("TPP_OVERCORR1", overcorrKernel()) : overcorrKernel());
            VLAUNCH(ctx, prolong_add, Pn, L.n);
        } else {
            Pn.out = oth;
            VLAUNCH(ctx, prolong, Pn, L.n);
            XL<R>(lv, oth);
            L.in = oth; L.out = v.Ac[lv]; L.r = v.r[lv];
            vSpmvDot2(L);
            L.c = oth; L.Ac = v.Ac[lv]; L.r = v.r[lv]; L.out = cur; L.sf = scal + S_TMP0; L.omega = (R)knobd("TPP_SCALEJ", 1.0);
            VLAUNCH(ctx, scale_apply, L, L.n);
            L.omega = omega;
        }
        for (int s = 0; s < std::max(nPost, 1); s++) {
            L.omega = (R)smootherOmega(s, std::max(nPost, 1));
            L.in = cur; L.out = oth;
            if (s == 0 && corrFused) {
                L.xc = Pn.xc; L.aggF = Pn.agg;
                L.oc = (R)(lv == 0 ? knobd("TPP_OVERCORR0", overcorrKernel()) : lv == 1 ? knobd("TPP_OVERCORR1", overcorrKernel()) : overcorrKernel());
                vRowOp(L, 3);
            } else {
                if (s == 0) XL<R>(lv, cur);
                else XLsmooth<R>(lv, cur, 2, oth);
                vRowOp(L, 0);
            }
            std::swap(cur, oth);
        }
        if (cur != x) d2d(ctx, x, cur, L.n * sizeof(R));
    }
    template <class R> void preconditionT(const tpp_solver_t& ctl, const double* r, double* z) {
        VStore<R>& v = vstore<R>();
        CastArgs<R> a;
        memset(&a, 0, sizeof(a));
        a.src = r; a.dst = v.b[0];
        VLAUNCH(ctx, cast_in, a, nC);
        int nv = ctl.type == 1 ? 1 : std::max(ctl.n_vcycles, 1);
        int nPre = knob("TPP_NPRE", 2), nPost = knob("TPP_NPOST", 2);
        for (int cyc = 0; cyc < nv; cyc++) vcycleT<R>(0, v.b[0], v.x[0], cyc == 0, nPre, nPost);
        a.rsrc = v.x[0]; a.ddst = z;
        VLAUNCH(ctx, cast_out, a, nC);
    }
    bool useFp32() const { return knob("TPP_FP32", 1) != 0; }
    bool scaledFallback = false;  // set when a fixed-factor solve hit maxIter: scaled corrections from then on
    static double overcorrKernel() { static const double v = knobd("TPP_OVERCORR", 1.6); return v; }
    void precondition(LV& F0, const tpp_solver_t& ctl, const double* r, double* z) {
        if (ctl.type == 0 && ctl.precond == 0) {  // PCG + DIC requested: diagonal preconditioning
            LV L = F0;
            L.omega = 1.0; L.b = const_cast<double*>(r); L.out = z;
            LAUNCH(ctx, jacobi0, L, L.n);
            return;
        }
        if (useFp32()) preconditionT<float>(ctl, r, z);
        else preconditionT<double>(ctl, r, z);
    }


    SolveStats solve(const tpp_solver_t& ctl, double* diag, double* upper, const double* b, double* x) {
        SolveStats st;
        if (!amgBuilt) buildAMG();
        LV F0 = fineView(diag, upper);
        bool useAMG = !levels.empty() && !(ctl.type == 0 && ctl.precond == 0);
        const bool jacobiOnly = ctl.type == 0 && ctl.precond == 0;
        if (useAMG) galerkin(F0); else { LAUNCH(ctx, rowsum, F0, F0.n); LAUNCH(ctx, fill_ev, F0, F0.n); }
        if (!jacobiOnly) {
            if (useFp32()) { ensureVStore<float>(); convertLevels<float>(F0); }
            else { ensureVStore<double>(); convertLevels<double>(F0); }
        }
        LV FG = F0;  // the global operator: full rows, ghost columns filled by halo exchange
        red.reduce(ctx, x, nullptr, nC, 2, scal + S_XSUM);
        allreduce(S_XSUM, 1, 0);
        X(x, 1);
        initResidual(FG, x, b);
        allreduce(S_RES, 2, 0);
        readScal();
        double nf = hscal[S_NORM] + 1e-20;
        st.r0 = st.r = hscal[S_RES] / nf;
        auto conv = [&](double r) { return r < ctl.tolerance || (ctl.rel_tol > 0 && r < ctl.rel_tol * st.r0); };
        if (conv(st.r)) return st;
        if (ctl.type == 1 && useAMG && knob("TPP_GAMG_STATIONARY", 0)) {
            // `solver GAMG` taken literally (fvSolution:42-48): stationary V-cycle iterations
            // x += V(b - A x) until the residual criterion is met.  Off by default: the same V-cycle as the
            // preconditioner of PCG (the default for both solver entries) reaches relTol 0.01 in about
            // half the cycles, and OpenFOAM's own criterion (tolerance, relTol) is what the caller asked for.
            const int maxIt = ctl.max_iter > 0 ? ctl.max_iter : 1000;
            do {
                precondition(F0, ctl, kr, kz);
#ifdef TPP_EMU
                for (int c = 0; c < nC; c++) x[c] += kz[c];
#else
                k_add<<<RED_BLOCKS, BLOCK, 0, ctx.stream>>>(nC, x, kz);
#endif
                ctx.launches++;
                X(x, 1);
                initResidual(FG, x, b);
                allreduce(S_RES, 2, 0);
                readScal();
                st.r = hscal[S_RES] / nf;
            } while (++st.iters < maxIt && !conv(st.r) && std::isfinite(st.r));
            if (!std::isfinite(st.r)) fail("the p_rgh solver residual is not finite (diverged)");
            return st;
        }
        // The stopping rule runs on the device (k_pcg_check after every iteration).  Large meshes read the
        // scalars back after every iteration (an iteration is ~1 ms, the round trip nothing); small ones
        // (the reference's own 8 k - 42 k-cell cases) launch as many iterations as the previous solve of
        // this kind took, less one, before the first read-back: a converged solve ignores the rest of its
        // chunk (k_update_xr returns at once), so the result is the same as with a check per iteration.
        pcgBegin(ctl.tolerance * nf, ctl.rel_tol > 0 ? ctl.rel_tol * st.r0 * nf : -1.0, (double)ctl.max_iter);
        dev_zero(ctx, kp, nC * sizeof(double));
        const int which = &ctl == &cfg.p_rgh_final ? 1 : 0;
        const bool chunked = knob("TPP_CHUNK", nGlobal < 500000 ? 1 : 0) != 0 && !ctx.prof;
        int chunk = chunked ? std::max(1, std::min(lastIters[which] - 1, ctl.max_iter)) : 1;
        int launched = 0;
        while (true) {
            for (int k = 0; k < chunk; k++) iteration(F0, FG, ctl, x);
            launched += chunk;
            readScal();
            if (hscal[S_DONE] != 0.0 || launched >= ctl.max_iter) break;
            chunk = 1;
        }
        st.iters = (int)(hscal[S_ITERS] + 0.5);
        // (S_RESF, not S_RES: iterations launched after the one that converged leave other reductions' sums there)
        st.r = (hscal[S_DONE] != 0.0 ? hscal[S_RESF] : hscal[S_RES]) / nf;
        lastIters[which] = st.iters;
        if (!std::isfinite(st.r)) fail("the p_rgh solver residual is not finite (diverged)");
        if (useAMG && !scaledFallback && knob("TPP_NOSCALE_FROM", 0) < 99 && st.iters >= ctl.max_iter && ctl.max_iter >= 20 && !conv(st.r) && st.r > 10 * ctl.tolerance) {
            // the fixed over-correction factors did not suit this hierarchy: use the adaptive scaling
            scaledFallback = true;
#ifndef TPP_EMU
            for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);
            graphs.clear();
#endif
            fprintf(stderr, "tppvof: p_rgh solve stopped at maxIter %d (residual %.3g): switching the multigrid to the scaled coarse correction\n", ctl.max_iter, st.r);
        }
        return st;
    }

    // one PCG iteration: z = M r ; wArA ; pA ; wA = A pA ; x, r update ; |r|
    void iterationBody(LV& F0, LV& FG, const tpp_solver_t& ctl, double* x) {
        precondition(F0, ctl, kr, kz);
        scalCopy(S_WARA_OLD, S_WARA);
        red.reduce(ctx, kz, kr, nC, 0, scal + S_WARA);
        allreduce(S_WARA, 1, 0);
        updateP();
        X(kp, 1);
        spmvDot(FG);
        allreduce(S_WAPA, 1, 0);
        updateXR(x);
        allreduce(S_RES, 1, 0);
        pcgCheck();
    }
    // (The FP64 V-cycle, TPP_FP32=0, once produced NaNs under graph replay; that was the
    // allocation-time memset racing a non-blocking caller stream, fixed in dev_alloc - it is
    // graphed like the FP32 one now and gives the same PCG counts.)
    // The iteration is ~110 small launches with fixed arguments: captured once per
    // (solver entry, solution vector) into a CUDA graph and replayed (launch-bound otherwise).
    void iteration(LV& F0, LV& FG, const tpp_solver_t& ctl, double* x) {
#ifndef TPP_EMU
        // (the legacy default stream cannot be captured)
        if (!ctx.prof && ctx.stream != nullptr && ctx.stream != cudaStreamLegacy && knob("TPP_GRAPH", 1) &&  (!comm.active || (comm.nccl && knob("TPP_GRAPH_PAR", 1)))) {
            GraphKey key{x, F0.diag, ctl.type, ctl.precond, ctl.n_vcycles};
            auto it = graphs.find(key);
            if (it == graphs.end()) {
                long l0 = ctx.launches;
                cudaGraph_t gr;
                CUDA_CHECK(cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal));
                iterationBody(F0, FG, ctl, x);
                CUDA_CHECK(cudaStreamEndCapture(ctx.stream, &gr));
                GraphRec rec;
                CUDA_CHECK(cudaGraphInstantiate(&rec.exec, gr, 0));
                cudaGraphDestroy(gr);
                rec.nodes = ctx.launches - l0;
                ctx.launches = l0;
                it = graphs.emplace(key, rec).first;
            }
            CUDA_CHECK(cudaGraphLaunch(it->second.exec, ctx.stream));
            ctx.launches += it->second.nodes;
            return;
        }
#endif
        iterationBody(F0, FG, ctl, x);
    }
    int lastIters[2] = {1, 1};
    void pcgCheck() {
#ifdef TPP_EMU
        if (scal[S_DONE] == 0.0) {
            scal[S_ITERS] += 1.0;
            const double res = scal[S_RES];
            if (res < scal[S_TOLA] || res < scal[S_TOLR] || !(fabs(scal[S_WAPA]) >= scal[S_NORM] * VSMALL) || scal[S_ITERS] >= scal[S_MAXIT] || !(res == res)) { scal[S_DONE] = 1.0; scal[S_RESF] = res; }
        }
#else
        k_pcg_check<<<1, 1, 0, ctx.stream>>>(scal);
#endif
        ctx.launches++;
    }
    void pcgBegin(double tolA, double tolR, double maxIt) {
#ifdef TPP_EMU
        scal[S_DONE] = 0.0; scal[S_ITERS] = 0.0; scal[S_TOLA] = tolA; scal[S_TOLR] = tolR; scal[S_MAXIT] = maxIt; scal[S_WARA] = 0.0;
#else
        k_pcg_begin<<<1, 1, 0, ctx.stream>>>(scal, tolA, tolR, maxIt);
#endif
        ctx.launches++;
    }
    void scalSet(int dst, double v) {
#ifdef TPP_EMU
        scal[dst] = v;
#else
        k_scal_set<<<1, 1, 0, ctx.stream>>>(scal, dst, v);
#endif
        ctx.launches++;
    }
    void scalCopy(int dst, int src) {
#ifdef TPP_EMU
        scal[dst] = scal[src];
#else
        k_scal_copy<<<1, 1, 0, ctx.stream>>>(scal, dst, src);
#endif
        ctx.launches++;
    }
    void initResidual(LV& F0, const double* x, const double* b) {
#ifdef TPP_EMU
        double xbar = scal[S_XSUM] / F0.nGlob, v = 0, w = 0;
        for (int c = 0; c < nC; c++) {
            double ax = row_Ax(F0, c, x), rr = b[c] - ax;
            kr[c] = rr; v += fabs(rr);
            double p = F0.rsum[c] * xbar;
            w += fabs(ax - p) + fabs(b[c] - p);
        }
        scal[S_RES] = v; scal[S_NORM] = w;
        ctx.launches += 3;
#else
        prof_begin(ctx, "init_residual");
        const int nb = std::min(RED_BLOCKS, (nC + BLOCK - 1) / BLOCK);
        k_init_residual<<<nb, BLOCK, 0, ctx.stream>>>(F0, x, b, kr, scal, red.partial, red.partial2);
        k_reduce_final2<<<2, BLOCK, 0, ctx.stream>>>(red.partial, red.partial2, nb, scal + S_RES, scal + S_NORM);
        prof_end(ctx);
        ctx.launches += 2;
#endif
    }
    void updateP() {
#ifdef TPP_EMU
        const bool first = scal[S_WARA_OLD] == 0.0;
        double beta = first ? 0.0 : scal[S_WARA] / scal[S_WARA_OLD];
        for (int c = 0; c < nC; c++) kp[c] = first ? kz[c] : kz[c] + beta * kp[c];
#else
        prof_begin(ctx, "update_p");
        k_update_p<<<RED_BLOCKS, BLOCK, 0, ctx.stream>>>(nC, kp, kz, scal);
        prof_end(ctx);
#endif
        ctx.launches++;
    }
    void spmvDot(LV& F0) {
#ifdef TPP_EMU
        double v = 0;
        for (int c = 0; c < nC; c++) { kw[c] = row_Ax(F0, c, kp); v += kw[c] * kp[c]; }
        scal[S_WAPA] = v;
#else
        LV L = F0;
        L.in = kp; L.out = kw;
        prof_begin(ctx, "spmv_dot");
        int nb = std::min(RED_BLOCKS, (nC + BLOCK - 1) / BLOCK);
        if (L.ell && (L.W == 4 || L.W == 6) && knob("TPP_ELL2", 1)) nb = std::min(4 * RED_BLOCKS, ((nC + 1) / 2 + BLOCK - 1) / BLOCK);
        if (L.ell && L.W == 4 && knob("TPP_ELL2", 1)) k_spmv_dot_ell2<4><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial);
        else if (L.ell && L.W == 6 && knob("TPP_ELL2", 1)) k_spmv_dot_ell2<6><<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial);
        else k_spmv_dot<<<nb, BLOCK, 0, ctx.stream>>>(L, red.partial);
        k_reduce_final<<<1, BLOCK, 0, ctx.stream>>>(red.partial, nb, 2, scal + S_WAPA);
        prof_end(ctx);
#endif
        ctx.launches += 2;
    }
    void updateXR(double* x) {
#ifdef TPP_EMU
        if (scal[S_DONE] == 0.0) {
            double alpha = scal[S_WARA] / scal[S_WAPA], v = 0;
            for (int c = 0; c < nC; c++) { x[c] += alpha * kp[c]; kr[c] -= alpha * kw[c]; v += fabs(kr[c]); }
            scal[S_RES] = v;
        }
#else
        prof_begin(ctx, "update_xr");
        const int nb = std::min(RED_BLOCKS, (nC + BLOCK - 1) / BLOCK);
        k_update_xr<<<nb, BLOCK, 0, ctx.stream>>>(nC, x, kr, kp, kw, scal, red.partial);
        k_reduce_final<<<1, BLOCK, 0, ctx.stream>>>(red.partial, nb, 2, scal + S_RES);
        prof_end(ctx);
#endif
        ctx.launches += 2;
    }

    void destroy() {
        for (auto& kv : stage) dev_free(kv.second);
        for (void* p : allocs) dev_free(p);
        for (auto& l : levels) l.free();
        for (auto& l : tail) l.free();
        dev_free(match); dev_free(prop); dev_free(root);
        dev_free(tailBar); dev_free(tailErr);
        red.free();
#ifdef TPP_EMU
        free(hscal);
#else
        for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);
        dropStepGraphs();
        if (ssHost) cudaFreeHost(ssHost);
        dev_free(ssDev);
        if (d2hStream) { cudaStreamSynchronize(d2hStream); cudaStreamDestroy(d2hStream); cudaEventDestroy(snapEvent); for (auto& kv : copied) cudaEventDestroy(kv.second); }
        for (void* p : comm.opened) cudaIpcCloseMemHandle(p);
        dev_free(comm.gwin); dev_free(comm.gSeq); dev_free(comm.gDone);
        dev_free(comm.window); dev_free(comm.seq); dev_free(comm.arSeq); dev_free(comm.putDone); dev_free(comm.p2pErr);
        if (hscal) cudaFreeHost(hscal);
        if (ctx.stream && ctx.ownStream) cudaStreamDestroy(ctx.stream);
#endif
    }
};

template <> tpp_solver::VStore<float>& tpp_solver::vstore<float>() { return vsF; }
template <> tpp_solver::VStore<double>& tpp_solver::vstore<double>() { return vsD; }

// ------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------
extern "C" {

const char* tpp_last_error(void) { return g_err.c_str(); }
const char* tpp_version(void) {
#ifdef TPP_EMU
    return "tppvof 0.1 (HOST EMULATION - tests only)";
#else
    return "tppvof 0.1 (sm_100a)";
#endif
}

int tpp_create(const tpp_mesh_t* mesh, const tpp_config_t* cfg, int device, tpp_handle* out) try {
    if (!out) { g_err = "tpp_create: null handle pointer"; return -1; }
    *out = nullptr;
#ifndef TPP_EMU
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_err = std::string("no usable CUDA device (") + cudaGetErrorString(e) + "): libtppvof has no CPU path";
        return -2;
    }
    if (device < 0 || device >= ndev) { g_err = "device index out of range"; return -3; }
    CUDA_CHECK(cudaSetDevice(device));
#endif
    tpp_solver* s = new tpp_solver();
    s->device = device;
#ifndef TPP_EMU
    CUDA_CHECK(cudaStreamCreateWithFlags(&s->ctx.stream, cudaStreamNonBlocking));  // no implicit ordering against the legacy stream or other handles
    s->ctx.ownStream = true;
#endif
    if (!s->build(mesh, cfg)) { s->destroy(); delete s; return -1; }
    dev_sync(s->ctx);
    *out = s;
    return 0;
} API_CATCH(-100)

int tpp_destroy(tpp_handle s) try {
    if (!s) return 0;
#ifndef TPP_EMU
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->ctx.stream);
#endif
    s->destroy();
    delete s->cs;
    delete s;
    return 0;
} API_CATCH(-100)

static void pointsNow(tpp_solver* s, std::vector<double>& out) {
    out.resize(3 * (size_t)s->nP);
    for (int i = 0; i < s->nP; i++) {
        double q[3] = {s->points0[3 * i] - s->cfg.cofg[0], s->points0[3 * i + 1] - s->cfg.cofg[1], s->points0[3 * i + 2] - s->cfg.cofg[2]};
        for (int k = 0; k < 3; k++) out[3 * i + k] = (s->Rn[3 * k] * q[0] + s->Rn[3 * k + 1] * q[1] + s->Rn[3 * k + 2] * q[2]) + s->cfg.cofg[k] + s->Tn[k];
    }
}

long tpp_size(tpp_handle s, const char* name) try {
    API_DEVICE(s);
    if (!strcmp(name, "points")) return 3L * s->nP;
    auto it = s->reg.find(name);
    return it == s->reg.end() ? -1 : it->second.second;
} API_CATCH(-100)
// face-sized arrays are kept in device face order (processor faces right after the internal
// ones); callers see OpenFOAM's file order
static int faceComp(tpp_solver* s, const char* name) {
    static const char* f1[] = {"phi", "meshPhi", "alphaPhi", "rhoPhi", "phiBD", "alphaPhiUn", "rAUf", "phiHbyA", "phig", "rec", "ghf", "magSf", "w", "dc", "nHatf", "stf"};
    static const char* f3[] = {"Uf", "Uf0", "mExpl", "Sf", "Cf0"};
    if (s->nG == 0) return 0;
    for (auto n : f1) if (!strcmp(n, name)) return 1;
    for (auto n : f3) if (!strcmp(n, name)) return 3;
    return 0;
}
long tpp_get(tpp_handle s, const char* name, double* out, long cap) try {
    API_DEVICE(s);
    if (!strcmp(name, "points")) {
        std::vector<double> p;
        pointsNow(s, p);
        memcpy(out, p.data(), std::min<long>(cap, (long)p.size()) * sizeof(double));
        return (long)p.size();
    }
    auto it = s->reg.find(name);
    if (it == s->reg.end()) { g_err = std::string("unknown array ") + name; return -1; }
    long n = std::min<long>(cap, it->second.second);
    if (int nc = faceComp(s, name)) {  // device order -> OpenFOAM file order, on the device
        s->ensurePermBuf();
        s->d.xsrc = it->second.first; s->d.xbuf = s->permBuf; s->d.xnc = nc; s->d.procOwner = s->dPerm;
        LAUNCH(s->ctx, face_to_file, s->d, s->nF);
        s->d.procOwner = s->dProcOwner;
        d2h(s->ctx, out, s->permBuf, n * sizeof(double));
        return it->second.second;
    }
    if (s->renumbered) {
        s->ensurePermBuf();
        if (s->permute(it->second.first, s->permBuf, it->second.second, true)) {
            d2h(s->ctx, out, s->permBuf, n * sizeof(double));
            return it->second.second;
        }
    }
    d2h(s->ctx, out, it->second.first, n * sizeof(double));
    return it->second.second;
} API_CATCH(-100)
long tpp_get_async(tpp_handle s, const char* name, double* out, long cap) try {
    API_DEVICE(s);
    auto it = s->reg.find(name);
    if (it == s->reg.end()) { g_err = std::string("unknown array ") + name; return -1; }
    const long len = it->second.second, n = std::min<long>(cap, len);
    double*& st = s->stage[name];
    if (!st) st = dalloc<double>((size_t)len);
#ifndef TPP_EMU
    if (!s->d2hStream) { CUDA_CHECK(cudaStreamCreateWithFlags(&s->d2hStream, cudaStreamNonBlocking)); CUDA_CHECK(cudaEventCreateWithFlags(&s->snapEvent, cudaEventDisableTiming)); }
    cudaEvent_t& cp = s->copied[name];
    if (!cp) CUDA_CHECK(cudaEventCreateWithFlags(&cp, cudaEventDisableTiming));
    else CUDA_CHECK(cudaStreamWaitEvent(s->ctx.stream, cp, 0));  // the previous snapshot of this array is on its way out: do not overwrite it yet
#endif
    // snapshot in file order on the solver's stream
    bool done = false;
    if (int nc = faceComp(s, name)) {
        s->ensurePermBuf();
        s->d.xsrc = it->second.first; s->d.xbuf = st; s->d.xnc = nc; s->d.procOwner = s->dPerm;
        LAUNCH(s->ctx, face_to_file, s->d, s->nF);
        s->d.procOwner = s->dProcOwner;
        done = true;
    } else if (s->renumbered) {
        s->ensurePermBuf();
        done = s->permute(it->second.first, st, len, true);
    }
    if (!done) d2d(s->ctx, st, it->second.first, (size_t)len * sizeof(double));
#ifdef TPP_EMU
    memcpy(out, st, (size_t)n * sizeof(double));
#else
    CUDA_CHECK(cudaEventRecord(s->snapEvent, s->ctx.stream));
    CUDA_CHECK(cudaStreamWaitEvent(s->d2hStream, s->snapEvent, 0));
    CUDA_CHECK(cudaMemcpyAsync(out, st, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->d2hStream));
    CUDA_CHECK(cudaEventRecord(cp, s->d2hStream));
#endif
    return len;
} API_CATCH(-100)
int tpp_sync(tpp_handle s) try {
    API_DEVICE(s);
#ifndef TPP_EMU
    if (s->d2hStream) {
        // the next snapshot of an array must not overtake the copy of its previous one
        CUDA_CHECK(cudaStreamSynchronize(s->d2hStream));
    }
#endif
    dev_sync(s->ctx);
    return 0;
} API_CATCH(-100)
long tpp_set(tpp_handle s, const char* name, const double* in, long n) try {
    API_DEVICE(s);
    auto it = s->reg.find(name);
    if (it == s->reg.end()) { g_err = std::string("unknown array ") + name; return -1; }
    if (n != it->second.second) { g_err = std::string("size mismatch for ") + name; return -2; }
    if (int nc = faceComp(s, name)) {
        s->ensurePermBuf();
        h2d(s->ctx, s->permBuf, in, n * sizeof(double));
        s->d.xsrc = s->permBuf; s->d.xbuf = it->second.first; s->d.xnc = nc; s->d.procOwner = s->dPerm;
        LAUNCH(s->ctx, file_to_face, s->d, s->nF);
        s->d.procOwner = s->dProcOwner;
        dev_sync(s->ctx);
        return n;
    }
    if (s->renumbered && s->kindOf(n).first != 0 && !(s->kindOf(n).first == 'F' && s->nG > 0)) {
        s->ensurePermBuf();
        h2d(s->ctx, s->permBuf, in, n * sizeof(double));
        s->permute(s->permBuf, it->second.first, n, false);
        dev_sync(s->ctx);
        return n;
    }
    h2d(s->ctx, it->second.first, in, n * sizeof(double));
    return n;
} API_CATCH(-100)
long tpp_get_int(tpp_handle s, const char* name, int* out, long cap) try {
    API_DEVICE(s);
    const int* src = nullptr;
    long n = 0;
    if (!strcmp(name, "cf")) { src = s->d.cf; n = (long)s->W * s->nCp; }
    else if (!strcmp(name, "cn")) { src = s->d.cn; n = (long)s->W * s->nCp; }
    else if (!strcmp(name, "owner")) { src = s->d.own; n = s->nF; }
    else if (!strcmp(name, "neighbour")) { src = s->d.nei; n = s->nI; }
    else if (!strcmp(name, "cellFileOf") || !strcmp(name, "faceFileOf")) {  // internal label -> file label (identity: not renumbered)
        const bool cells = name[0] == 'c';
        const long m = cells ? s->nC : s->nF;
        for (long i = 0; out && i < std::min(cap, m); i++) out[i] = s->renumbered ? (cells ? s->cellFileOf[i] : s->faceFileOf[i]) : (int)i;
        return m;
    }
    else if (!strcmp(name, "layout")) {
        int v[6] = {s->nC, s->nCp, s->W, s->nI, s->nB, s->nG};
        memcpy(out, v, std::min<long>(cap, 6) * sizeof(int));
        return 6;
    } else { g_err = std::string("unknown integer array ") + name; return -1; }
    if (out) d2h(s->ctx, out, src, std::min(cap, n) * sizeof(int));
    return n;
} API_CATCH(-100)
int tpp_device_ptr(tpp_handle s, const char* name, void** ptr, long* n) try {
    API_DEVICE(s);
    auto it = s->reg.find(name);
    if (it == s->reg.end()) { g_err = std::string("unknown array ") + name; return -1; }
    *ptr = it->second.first;
    *n = it->second.second;
    return 0;
} API_CATCH(-100)
int tpp_init_fields(tpp_handle s) try {
    API_DEVICE(s);
    s->alphaBCs();
    s->mixture();
    s->X(s->d.alpha, 1); s->X(s->d.rho, 1); s->X(s->d.U, 3); s->X(s->d.p_rgh, 1);
    d2d(s->ctx, s->d.rho0, s->d.rho, (size_t)(s->nC + s->nG) * sizeof(double));
    dev_sync(s->ctx);
    return 0;
} API_CATCH(-100)
int tpp_set_delta_t(tpp_handle s, double dt) {
    if (!s || !(dt > 0)) { g_err = "tpp_set_delta_t: null handle or non-positive deltaT"; return -1; }
    s->dt = s->dt0 = dt;
    return 0;
}
int tpp_set_time(tpp_handle s, double t, double dt) try {
    API_DEVICE(s);
    s->t = t; s->dt = s->dt0 = dt;
    s->motionAt(t, s->Rn, s->Tn);
    memcpy(s->Ro, s->Rn, sizeof(s->Rn)); memcpy(s->To, s->Tn, sizeof(s->Tn));
    s->setTransform();
    s->fillStepScal(true);
    s->orientGeometry();
    dev_sync(s->ctx);
    return 0;
} API_CATCH(-100)

int tpp_step(tpp_handle s, int n) try {
    API_DEVICE(s);
    for (int i = 0; i < n && s->ctx.err.empty(); i++) s->oneStep();
    dev_sync(s->ctx);
    if (!s->ctx.err.empty()) { g_err = s->ctx.err; return -1; }
    return 0;
} API_CATCH(-100)
int tpp_run_to_write(tpp_handle s, long max_steps) try {
    API_DEVICE(s);
    for (long i = 0; i < max_steps; i++) {
        if (!(s->t < s->cfg.end_time - 0.5 * s->dt)) { dev_sync(s->ctx); return 0; }
        const bool wr = s->oneStep();
        if (!s->ctx.err.empty()) { dev_sync(s->ctx); g_err = s->ctx.err; return -1; }
        if (wr) { dev_sync(s->ctx); return 1; }
    }
    dev_sync(s->ctx);
    return 2;
} API_CATCH(-100)
int tpp_stage(tpp_handle s, const char* name) try {
    API_DEVICE(s);
    std::string n(name);
    s->d.dt = s->dt;
    s->fillStepScal(true);
    if (n == "courant") s->courant();
    else if (n == "adjustDeltaT") s->adjustDeltaT();
    else if (n == "advanceTime") s->advanceTime();
    else if (n == "moveMesh") s->moveMesh();
    else if (n == "alphaBCs") s->alphaBCs();
    else if (n == "UBCs") s->UBCs();
    else if (n == "mixture") s->mixture();
    else if (n == "alphaSubCycle") s->alphaSubCycle(s->dt / s->cfg.n_alpha_subcycles);
    else if (n == "alphaPredictor") s->alphaPredictor();
    else if (n == "momentum") s->momentum();
    else if (n == "HbyA") s->computeHbyA();
    else if (n == "pcPrepare") s->pcPrepare();
    else if (n == "pcAssemble") s->pcAssemble();
    else if (n == "pcFinish") { LAUNCH(s->ctx, p_evaluate, s->d, s->nB); s->pcFinish(); }
    else if (n == "pcEnd") s->pcEnd();
    else if (n == "pressureCorrector:0") s->pressureCorrector(false);
    else if (n == "pressureCorrector:1") s->pressureCorrector(true);
    else { g_err = "unknown stage " + n; return -1; }
    dev_sync(s->ctx);
    return 0;
} API_CATCH(-100)
int tpp_info(tpp_handle s, double* o) try {
    API_DEVICE(s);
    o[0] = s->t; o[1] = s->dt; o[2] = (double)s->step; o[3] = s->Co; o[4] = s->alphaCo;
    o[5] = s->lastSolve[0].iters; o[6] = s->lastSolve[0].r0; o[7] = s->lastSolve[0].r;
    o[8] = s->lastSolve[1].iters; o[9] = s->lastSolve[1].r0; o[10] = s->lastSolve[1].r;
    o[11] = s->d.needRef ? (s->renumbered && s->d.refCell >= 0 ? s->cellFileOf[s->d.refCell] : s->d.refCell) : -1; o[12] = s->d.deltaN; o[13] = s->writeTimeIndex;
    o[14] = (double)s->levels.size(); o[15] = (double)s->ctx.launches;
    return 0;
} API_CATCH(-100)
int tpp_interface(tpp_handle s, double iso, double* out5) try {
    API_DEVICE(s);
    if (s->comm.active) { g_err = "tpp_interface: a decomposed mesh is not supported (points on processor patches see only this rank's cells); reconstruct first"; return -1; }
    s->interfaceSummary(iso, out5);
    out5[4] = s->t;
    return 0;
} API_CATCH(-100)
int tpp_stats(tpp_handle s, int reset, double* o) try {
    API_DEVICE(s);
    if (o) {
        o[0] = s->stSteps; o[1] = s->stIt[0]; o[2] = s->stIt[1]; o[3] = s->stItMax[0]; o[4] = s->stItMax[1]; o[5] = s->stCap;
        o[6] = s->stVol0; o[7] = s->statsOn ? s->alphaVolume() : 0.0; o[8] = s->stBndInt;
        dev_sync(s->ctx);
    }
    if (reset >= 0) { s->statsOn = reset != 0; if (s->statsOn) s->statsReset(); }
    return 0;
} API_CATCH(-100)
int tpp_solve(tpp_handle s, const tpp_solver_t* ctl, const double* diag, const double* upper, const double* b, double* x, double* r0, double* r) try {
    API_DEVICE(s);
    double* xd = s->d.cellTmp;
    auto put = [&](double* dst, const double* src, long n) {  // file order in
        if (s->renumbered) { s->ensurePermBuf(); h2d(s->ctx, s->permBuf, src, n * sizeof(double)); s->permute(s->permBuf, dst, n, false); dev_sync(s->ctx); }
        else h2d(s->ctx, dst, src, n * sizeof(double));
    };
    put(s->d.pDiag, diag, s->nC); put(s->d.pUpper, upper, s->nI); put(s->d.pSource, b, s->nC); put(xd, x, s->nC);
    SolveStats st = s->solve(*ctl, s->d.pDiag, s->d.pUpper, s->d.pSource, xd);
    if (s->renumbered) { s->permute(xd, s->permBuf, s->nC, true); d2h(s->ctx, x, s->permBuf, s->nC * sizeof(double)); }
    else d2h(s->ctx, x, xd, s->nC * sizeof(double));
    *r0 = st.r0; *r = st.r;
    return st.iters;
} API_CATCH(-100)
int tpp_set_probes(tpp_handle s, int n, const int* cells) try {  // cell labels of the mesh FILE (as tpp_find_cell returns them)
    API_DEVICE(s);
    if (n < 0 || (n > 0 && !cells)) { g_err = "tpp_set_probes: bad probe list"; return -1; }
    for (int i = 0; i < n; i++) if (cells[i] >= s->nC) { g_err = "tpp_set_probes: cell label " + std::to_string(cells[i]) + " out of range"; return -2; }
    s->flushProbes();
    s->probeCells.assign(cells, cells + n);
    s->probeUploaded = -1;
    if (s->renumbered) for (int& c : s->probeCells) if (c >= 0 && c < s->nC) c = s->cellNewOf[c];
    return 0;
} API_CATCH(-100)
long tpp_probe_log(tpp_handle s, double* out, long cap_rows) try {
    API_DEVICE(s);
    s->flushProbes();
    long w = 1 + (long)s->probeCells.size();
    long rows = (long)s->probeLog.size() / w;
    long n = std::min(rows, cap_rows);
    memcpy(out, s->probeLog.data(), n * w * sizeof(double));
    s->probeLog.erase(s->probeLog.begin(), s->probeLog.begin() + n * w);
    return n;
} API_CATCH(-100)
int tpp_find_cell(tpp_handle s, const double* xyz) try {  // -1: the point is in no cell (also: null handle)
    if (!s || !xyz) { g_err = "tpp_find_cell: null handle or point"; return -1; }
    int c = s->findCell(xyz);
    return (c >= 0 && s->renumbered) ? s->cellFileOf[c] : c;
} API_CATCH(-1)
int tpp_use_stream(tpp_handle s, void* stream) try {
    API_DEVICE(s);
    s->dropStepGraphs();
#ifndef TPP_EMU
    cudaStreamSynchronize(s->ctx.stream);
    if (s->ctx.ownStream) cudaStreamDestroy(s->ctx.stream);
    s->ctx.stream = (cudaStream_t)stream;
    s->ctx.ownStream = false;
#else
    (void)s; (void)stream;
#endif
    return 0;
} API_CATCH(-100)
int tpp_profile(tpp_handle s, int on) try {
    API_DEVICE(s);
    s->dropStepGraphs();
    s->ctx.prof = on != 0;
    return 0;
} API_CATCH(-100)
long tpp_profile_report(tpp_handle s, char* buf, long cap) try {
    API_DEVICE(s);
    dev_sync(s->ctx);
    std::map<std::string, std::pair<long, double>> agg;
    for (auto& r : s->ctx.recs) {
        float ms = 0;
#ifndef TPP_EMU
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
#endif
        auto& a = agg[r.name];
        a.first++;
        a.second += ms;
    }
    s->ctx.recs.clear();
    std::string out;
    char line[160];
    for (auto& kv : agg) {
        snprintf(line, sizeof(line), "%s %ld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if ((long)out.size() + 1 > cap) return -(long)out.size() - 1;
    memcpy(buf, out.c_str(), out.size() + 1);
    return (long)out.size();
} API_CATCH(-100)
int tpp_amg_levels(tpp_handle s, int* n_rows, int* n_faces, int cap) try {
    API_DEVICE(s);
    int k = 0;
    if (k < cap) { n_rows[k] = s->nC; n_faces[k] = s->nIloc; }
    k++;
    for (auto& l : s->levels) { if (k < cap) { n_rows[k] = l.n; n_faces[k] = l.nfLoc; } k++; }
    for (size_t t = 1; t < s->tail.size(); t++) { if (k < cap) { n_rows[k] = s->tail[t].n; n_faces[k] = s->tail[t].nf; } k++; }
    return k;
} API_CATCH(-100)
int tpp_amg_layout(tpp_handle s, int* out4) try {
    API_DEVICE(s);
    out4[0] = s->levels.empty() ? 0 : s->vLevels();
    out4[1] = (int)s->tail.size();
    out4[2] = s->tail.empty() ? 0 : s->tail[0].n;
    out4[3] = s->tailGrid;
    return 0;
} API_CATCH(-100)
int tpp_ghost_layout(tpp_handle s, int* n_ghost, int* n_patches, int* off, int* cnt, int* peer, int cap) try {
    API_DEVICE(s);
    *n_ghost = s->nG;
    *n_patches = (int)s->procCnt.size();
    for (int i = 0; i < (int)s->procCnt.size() && i < cap; i++) { off[i] = s->procOff[i]; cnt[i] = s->procCnt[i]; peer[i] = s->procPeer[i]; }
    return 0;
} API_CATCH(-100)
int tpp_comm_callbacks(tpp_handle s, int rank, int n_ranks, exchange_cb_t xcb, allreduce_cb_t rcb, void* user) try {
    API_DEVICE(s);
    s->comm.rank = rank; s->comm.size = n_ranks; s->comm.xcb = xcb; s->comm.rcb = rcb; s->comm.user = user;
    s->comm.active = n_ranks > 1;
    if (!s->finalizeParallel()) return -1;
    dev_sync(s->ctx);
    return 0;
} API_CATCH(-100)
#ifndef TPP_EMU
static bool loadNccl(Comm& c, const char* path) {
    if (c.lib) return true;
    c.lib = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!c.lib) { g_err = std::string("cannot open NCCL: ") + dlerror(); return false; }
#define SYM(field, name) *(void**)(&c.field) = dlsym(c.lib, name); if (!c.field) { g_err = std::string("NCCL symbol missing: ") + name; return false; }
    SYM(pGetUniqueId, "ncclGetUniqueId") SYM(pCommInitRank, "ncclCommInitRank") SYM(pCommDestroy, "ncclCommDestroy")
    SYM(pSend, "ncclSend") SYM(pRecv, "ncclRecv") SYM(pAllReduce, "ncclAllReduce") SYM(pGroupStart, "ncclGroupStart") SYM(pGroupEnd, "ncclGroupEnd")
#undef SYM
    return true;
}
#endif
int tpp_nccl_unique_id(const char* nccl_path, char* out128) try {
#ifndef TPP_EMU
    Comm c;
    if (!loadNccl(c, nccl_path)) return -1;
    ncclUniqueId id;
    if (c.pGetUniqueId(&id) != ncclSuccess) { g_err = "ncclGetUniqueId failed"; return -2; }
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return 0;
#else
    (void)nccl_path; (void)out128;
    g_err = "host emulation has no NCCL transport (use tpp_comm_callbacks)";
    return -1;
#endif
} API_CATCH(-100)
int tpp_comm_init(tpp_handle s, int rank, int n_ranks, const char* id128, const char* nccl_path) try {
    API_DEVICE(s);
#ifndef TPP_EMU
    if (!loadNccl(s->comm, nccl_path)) return -1;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    CUDA_CHECK(cudaSetDevice(s->device));
    if (s->comm.pCommInitRank(&s->comm.nccl, n_ranks, id, rank) != ncclSuccess) { g_err = "ncclCommInitRank failed"; return -2; }
    s->comm.rank = rank; s->comm.size = n_ranks; s->comm.active = n_ranks > 1;
    s->setupP2P();
    if (!s->finalizeParallel()) return -3;
    dev_sync(s->ctx);
    return 0;
#else
    (void)s; (void)rank; (void)n_ranks; (void)id128; (void)nccl_path;
    g_err = "host emulation has no NCCL transport (use tpp_comm_callbacks)";
    return -1;
#endif
} API_CATCH(-100)
}

// ------------------------------------------------------------------------------------
// case directories (tpp_caseio.h): what `foamRun` does around the time loop - read the case
// (Makefile:85 `foamRun`, started in the case directory), write the time directories and the
// probes log.  Host code only; everything below goes through the array ABI above, so the state
// a handle made by tpp_open holds is bit for bit the one the Python host sets up (tests/test_caseio.py).
// ------------------------------------------------------------------------------------
namespace {
#define CASE_CATCH(code) catch (const caseio::Error& e) { g_err = e.what(); return (code); } API_CATCH(-100)

std::vector<double> getArray(tpp_handle s, const char* name) {
    long n = tpp_size(s, name);
    if (n < 0) caseio::fail(std::string("internal: the solver has no array '") + name + "'");
    std::vector<double> a((size_t)n);
    if (tpp_get(s, name, a.data(), n) != n) caseio::fail(std::string("tpp_get(") + name + "): " + g_err);
    return a;
}
void setArray(tpp_handle s, const char* name, const std::vector<double>& a) {
    if (tpp_size(s, name) != (long)a.size()) caseio::fail(std::string("internal: size of '") + name + "' does not match the case");
    if (tpp_set(s, name, a.data(), (long)a.size()) != (long)a.size()) caseio::fail(std::string("tpp_set(") + name + "): " + g_err);
}
void check(int rc, const char* what) {
    if (rc != 0) caseio::fail(std::string(what) + ": " + g_err);
}

// start fields of the opened case (Solver.load_case_fields + the restart part of foamrun.run_case)
void caseStart(tpp_handle s) {
    caseio::Case& cs = *s->cs;
    const int nC = cs.mesh.nCells, nI = cs.mesh.nInternal(), nF = cs.mesh.nFaces();
    const std::string td = cs.dir + "/" + cs.startName;
    setArray(s, "alpha", cs.alpha.internal.expand(nC, 1, td + "/alpha.water:internalField"));
    setArray(s, "U", cs.U.internal.expand(nC, 3, td + "/U:internalField"));
    setArray(s, "p_rgh", cs.p_rgh.internal.expand(nC, 1, td + "/p_rgh:internalField"));
    check(tpp_init_fields(s), "tpp_init_fields");
    // boundary values the next step reads before it re-evaluates them (old-time wall / atmosphere
    // velocity in ddtCorr, p_rgh on the fixedFluxPressure walls): a restart takes them from the files
    const caseio::Field* flds[2] = {&cs.U, &cs.p_rgh};
    const char* arrs[2] = {"U_b", "p_rgh_b"};
    for (int k = 0; k < 2; k++) {
        const int nc = flds[k]->nc;
        std::vector<double> cur = getArray(s, arrs[k]);
        bool changed = false;
        for (auto& q : cs.mesh.patches) {
            if (q.type == "processor" || q.nFaces == 0) continue;
            const caseio::BoundaryEntry* e = flds[k]->patch(q.name);
            if (!e || !e->value.present) continue;
            std::vector<double> v = e->value.expand(q.nFaces, nc, td + ":" + q.name + ".value");
            std::copy(v.begin(), v.end(), cur.begin() + (size_t)(q.startFace - nI) * nc);
            changed = true;
        }
        if (changed) setArray(s, arrs[k], cur);
    }
    if (cs.hasRestartDt) check(tpp_set_delta_t(s, cs.restartDt), "tpp_set_delta_t");
    if (cs.hasFlux && cs.startValue > 0) {  // restart: internal + boundary values of the face fields
        const caseio::Field* ff[2] = {&cs.phi, &cs.Uf};
        const char* names[2] = {"phi", "Uf"};
        for (int k = 0; k < 2; k++) {
            const int nc = ff[k]->nc;
            std::vector<double> a((size_t)nF * nc, 0.0), in = ff[k]->internal.expand(nI, nc, td + "/" + names[k] + ":internalField");
            std::copy(in.begin(), in.end(), a.begin());
            for (auto& q : cs.mesh.patches) {
                const caseio::BoundaryEntry* e = ff[k]->patch(q.name);
                if (!e || !e->value.present) continue;
                std::vector<double> v = e->value.expand(q.nFaces, nc, td + "/" + names[k] + ":" + q.name + ".value");
                std::copy(v.begin(), v.end(), a.begin() + (size_t)q.startFace * nc);
            }
            setArray(s, names[k], a);
        }
        check(tpp_set_time(s, cs.startValue, cs.hasRestartDt && cs.restartDt != 0.0 ? cs.restartDt : cs.cfg.c.delta_t), "tpp_set_time");
    }
    cs.started = true;
}

// one time directory from the device state (foamrun.write_time)
void caseWrite(tpp_handle s) {
    caseio::Case& cs = *s->cs;
    const caseio::Mesh& m = cs.mesh;
    const int nC = m.nCells, nI = m.nInternal();
    const bool bin = cs.cfg.writeBinary;
    const int prec = cs.cfg.writePrecision;
    double info[16];
    check(tpp_info(s, info), "tpp_info");
    const std::string name = caseio::timeName(info[0], cs.cfg.timePrecision), tdir = cs.dir + "/" + name;
    caseio::makeDirs(tdir);
    long nBphys = 0;
    for (auto& q : m.patches) if (q.type != "processor") nBphys += q.nFaces;

    std::vector<std::vector<double>> keep;  // per-patch values stay alive until the file is written
    // vol field: physical patches from the solver's boundary array, processor patches from the adjacent cells
    auto volPatches = [&](const caseio::Field* src, const std::vector<double>& cells, const std::vector<double>& bnd, int nc) {
        std::vector<caseio::PatchOut> out;
        for (auto& q : m.patches) {
            caseio::PatchOut po;
            po.name = q.name;
            const caseio::BoundaryEntry* e = src ? src->patch(q.name) : nullptr;
            if (src && e) po.entries = e->entries;
            else po.entries = {{"type", src ? "calculated" : (q.type == "processor" ? "processor" : "calculated")}};
            po.n = q.nFaces;
            if (q.type == "processor") {
                keep.emplace_back((size_t)q.nFaces * nc);
                for (int f = 0; f < q.nFaces; f++)
                    for (int c = 0; c < nc; c++) keep.back()[(size_t)f * nc + c] = cells[(size_t)m.owner[q.startFace + f] * nc + c];
                po.value = keep.back().data();
            } else po.value = bnd.data() + (size_t)(q.startFace - nI) * nc;
            out.push_back(po);
        }
        return out;
    };
    // surface field: the slice of the boundary part of the array (file face order)
    auto facePatches = [&](const std::vector<double>& arr, int nc) {
        std::vector<caseio::PatchOut> out;
        for (auto& q : m.patches) {
            caseio::PatchOut po;
            po.name = q.name;
            po.entries = {{"type", q.type == "processor" ? "processor" : "calculated"}};
            po.n = q.nFaces;
            po.value = arr.data() + (size_t)q.startFace * nc;
            out.push_back(po);
        }
        return out;
    };
    // one field at a time: the host copy of a field is dropped before the next one is fetched (a 50 M-cell
    // tank's Uf alone is 2.4 GB)
    std::vector<double> p_rgh_b = getArray(s, "p_rgh_b"), rho_b = getArray(s, "rho_b");
    if ((long)rho_b.size() < nBphys || (long)p_rgh_b.size() < nBphys) caseio::fail("internal: boundary arrays are smaller than the case's physical patches");
    {
        std::vector<double> alpha = getArray(s, "alpha"), alpha_b = getArray(s, "alpha_b");
        caseio::writeField(tdir + "/alpha.water", "volScalarField", "alpha.water", name, "[0 0 0 0 0 0 0]", alpha.data(), nC, 1, volPatches(&cs.alpha, alpha, alpha_b, 1), bin, prec);
    }
    {
        std::vector<double> U = getArray(s, "U"), U_b = getArray(s, "U_b");
        caseio::writeField(tdir + "/U", "volVectorField", "U", name, "[0 1 -1 0 0 0 0]", U.data(), nC, 3, volPatches(&cs.U, U, U_b, 3), bin, prec);
    }
    {
        std::vector<double> p_rgh = getArray(s, "p_rgh");
        caseio::writeField(tdir + "/p_rgh", "volScalarField", "p_rgh", name, "[1 -1 -2 0 0 0 0]", p_rgh.data(), nC, 1, volPatches(&cs.p_rgh, p_rgh, p_rgh_b, 1), bin, prec);
    }
    {
        std::vector<double> pb((size_t)nBphys);
        {
            std::vector<double> ghf = getArray(s, "ghf");
            if ((long)ghf.size() < nI + nBphys) caseio::fail("internal: ghf is smaller than the case's faces");
            for (long k = 0; k < nBphys; k++) pb[k] = p_rgh_b[k] + rho_b[k] * ghf[nI + k];
        }
        std::vector<double> p = getArray(s, "p");
        caseio::writeField(tdir + "/p", "volScalarField", "p", name, "[1 -1 -2 0 0 0 0]", p.data(), nC, 1, volPatches(nullptr, p, pb, 1), bin, prec);
    }
    {
        std::vector<double> rho = getArray(s, "rho");
        caseio::writeField(tdir + "/rho", "volScalarField", "rho", name, "[1 -3 0 0 0 0 0]", rho.data(), nC, 1, volPatches(nullptr, rho, rho_b, 1), bin, prec);
    }
    {
        std::vector<double> phi = getArray(s, "phi");
        caseio::writeField(tdir + "/phi", "surfaceScalarField", "phi", name, "[0 3 -1 0 0 0 0]", phi.data(), nI, 1, facePatches(phi, 1), bin, prec);
    }
    {
        std::vector<double> Uf = getArray(s, "Uf");
        caseio::writeField(tdir + "/Uf", "surfaceVectorField", "Uf", name, "[0 1 -1 0 0 0 0]", Uf.data(), nI, 3, facePatches(Uf, 3), bin, prec);
    }
    if (cs.cfg.c.n_motion > 0) {
        caseio::makeDirs(tdir + "/polyMesh");
        std::vector<double> pts = getArray(s, "points");
        caseio::writePoints(tdir + "/polyMesh/points", name + "/polyMesh", pts.data(), (long)pts.size() / 3, bin);
    }
    // uniform/time last: its presence marks the directory as complete (caseio::latestTime)
    caseio::makeDirs(tdir + "/uniform");
    caseio::Out o(tdir + "/uniform/time");
    o.str(caseio::fileHeader("dictionary", "time", name + "/uniform", false));
    char b[512];
    snprintf(b, sizeof b, "value           %.17g;\n\nname            \"%s\";\n\nindex           %ld;\n\ndeltaT          %.17g;\n\ndeltaT0         %.17g;\n", info[0], name.c_str(), (long)info[2], info[1], info[1]);
    o.str(b);
    o.str(caseio::FILE_END);
    o.close();
}

std::string g6(double v) { return caseio::fmtNum(v, 6); }
std::string shortest(double v) {  // the fewest digits that read back as v
    for (int p = 1; p < 17; p++) {
        std::string t = caseio::fmtNum(v, p);
        if (strtod(t.c_str(), nullptr) == v) return t;
    }
    return caseio::fmtNum(v, 17);
}
void probeRows(caseio::Case& cs, const double* rows, long n, long w) {
    if (!cs.probesFile) return;
    for (long r = 0; r < n; r++) {
        char b[64];
        snprintf(b, sizeof b, "%-13s ", g6(rows[r * w]).c_str());
        std::string line = b;
        for (long k = 1; k < w; k++) {
            snprintf(b, sizeof b, "%-13s", g6(rows[r * w + k]).c_str());
            line += std::string(k > 1 ? " " : "") + b;
        }
        while (!line.empty() && line.back() == ' ') line.pop_back();
        fprintf(cs.probesFile, "%s\n", line.c_str());
    }
    fflush(cs.probesFile);
}
}  // namespace

extern "C" {

int tpp_open(const char* case_dir, int processor, int device, tpp_handle* out) try {
    if (!out) { g_err = "tpp_open: null handle pointer"; return -1; }
    *out = nullptr;
    if (!case_dir) { g_err = "tpp_open: null case directory"; return -1; }
    std::unique_ptr<caseio::Case> cs(new caseio::Case());
    caseio::load(case_dir, processor, *cs);
    tpp_mesh_t m = caseio::meshView(*cs);
    tpp_handle s = nullptr;
    int rc = tpp_create(&m, &cs->cfg.c, device, &s);
    if (rc != 0) return rc;
    s->cs = cs.release();
    if (processor < 0) {
        try {
            caseStart(s);
        } catch (...) {
            std::string keep;
            try { throw; } catch (const std::exception& e) { keep = e.what(); } catch (const tpp::CudaFailure& e) { keep = e.what; } catch (...) { keep = "internal error"; }
            tpp_destroy(s);
            g_err = keep;
            return -1;
        }
    }
    *out = s;
    return 0;
} CASE_CATCH(-4)

int tpp_case_start(tpp_handle s) try {
    API_DEVICE(s);
    if (!s->cs) { g_err = "tpp_case_start: the handle was not made by tpp_open"; return -1; }
    caseStart(s);
    return 0;
} CASE_CATCH(-4)

int tpp_write_time(tpp_handle s) try {
    API_DEVICE(s);
    if (!s->cs) { g_err = "tpp_write_time: the handle was not made by tpp_open"; return -1; }
    caseWrite(s);
    return 0;
} CASE_CATCH(-4)

long tpp_case_query(tpp_handle s, const char* what, char* text, long cap) try {
    API_DEVICE(s);
    if (!s->cs || !what) { g_err = "tpp_case_query: the handle was not made by tpp_open"; return -1; }
    caseio::Case& cs = *s->cs;
    std::string w(what), t;
    long v = 0;
    if (w == "n_cells") v = cs.mesh.nCells;
    else if (w == "n_faces") v = cs.mesh.nFaces();
    else if (w == "n_internal") v = cs.mesh.nInternal();
    else if (w == "n_points") v = (long)cs.mesh.points.size() / 3;
    else if (w == "n_patches") v = (long)cs.mesh.patches.size();
    else if (w == "n_probes") v = (long)cs.cfg.probes.size() / 3;
    else if (w == "write_binary") v = cs.cfg.writeBinary;
    else if (w == "start_time") { t = cs.startName; v = (long)t.size(); }
    else if (w == "time") { double info[16]; tpp_info(s, info); t = caseio::timeName(info[0], cs.cfg.timePrecision); v = (long)t.size(); }
    else if (w == "dir") { t = cs.dir; v = (long)t.size(); }
    else { g_err = "tpp_case_query: unknown item '" + w + "'"; return -2; }
    if (text && cap > 0) { strncpy(text, t.c_str(), (size_t)cap - 1); text[cap - 1] = 0; }
    return v;
} CASE_CATCH(-4)

long tpp_run_case(tpp_handle s, long max_steps, int flags) try {
    const bool verbose = flags & TPP_RUN_LOG;
    API_DEVICE(s);
    if (!s->cs) { g_err = "tpp_run_case: the handle was not made by tpp_open"; return -1; }
    caseio::Case& cs = *s->cs;
    if (!cs.started) { g_err = "tpp_run_case: call tpp_case_start first (a processor share starts after tpp_comm_init)"; return -1; }
    const int np = (int)(cs.cfg.probes.size() / 3);
    double info[16];
    check(tpp_info(s, info), "tpp_info");
    if (cs.cfg.hasProbes && np > 0 && cs.probeCells.empty()) {
        // `probes` function object (system/functions:17-33): postProcessing/probes/<start time>/p
        for (int k = 0; k < np; k++) cs.probeCells.push_back(tpp_find_cell(s, &cs.cfg.probes[3 * k]));
        check(tpp_set_probes(s, np, cs.probeCells.data()), "tpp_set_probes");
        if (s->nG == 0) {  // a processor share leaves the file to its host, which merges the ranks' rows (tpp_probe_log)
            std::string d = cs.dir + "/postProcessing/probes/" + cs.startName;
            caseio::makeDirs(d);
            cs.probesFile = fopen((d + "/p").c_str(), "w");
            if (!cs.probesFile) caseio::fail(d + "/p: cannot open for writing");
            for (int k = 0; k < np; k++) fprintf(cs.probesFile, "# Probe %d (%s %s %s)\n", k, g6(cs.cfg.probes[3 * k]).c_str(), g6(cs.cfg.probes[3 * k + 1]).c_str(), g6(cs.cfg.probes[3 * k + 2]).c_str());
            std::string hdr = "# Time        ";
            for (int k = 0; k < np; k++) {
                char b[32];
                snprintf(b, sizeof b, "%-13d", k);
                hdr += std::string(k ? " " : "") + b;
            }
            fprintf(cs.probesFile, "%s\n", hdr.c_str());
            std::vector<double> pnow = getArray(s, "p"), row(1 + np);
            row[0] = cs.startValue;
            for (int k = 0; k < np; k++) row[1 + k] = cs.probeCells[k] >= 0 ? pnow[cs.probeCells[k]] : -1.79769e307;
            probeRows(cs, row.data(), 1, 1 + np);
        }
    }
    // in-situ interface statistics (`foamRun -interface`): the rows the reference's extract_interface derives
    // from the time directories afterwards (main.py:751-780), computed on the device at every write time
    struct Closer { FILE* f = nullptr; ~Closer() { if (f) fclose(f); } } iface;
    auto ifaceRow = [&]() {
        double o[5];
        check(tpp_interface(s, 0.5, o), "tpp_interface");
        fprintf(iface.f, "\n%s,%s,%s,%s,%ld", shortest(o[4]).c_str(), shortest(o[0]).c_str(), shortest(o[1]).c_str(), shortest(o[2]).c_str(), (long)o[3]);
        fflush(iface.f);
    };
    if ((flags & TPP_RUN_INTERFACE) && s->nG == 0) {
        std::string d = cs.dir + "/postProcessing/interface";
        caseio::makeDirs(d);
        const bool fresh = cs.startValue == 0 && !cs.interfaceStarted;
        iface.f = fopen((d + "/interface_summary.csv").c_str(), fresh ? "w" : "a");
        if (!iface.f) caseio::fail(d + "/interface_summary.csv: cannot open for writing");
        if (fresh) {
            fprintf(iface.f, "time,max_z,min_z,mean_z,num_points");
            ifaceRow();
        }
        cs.interfaceStarted = true;
    }
    const long step0 = (long)info[2];
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<double> rows;
    for (;;) {
        check(tpp_info(s, info), "tpp_info");
        long budget = max_steps < 0 ? 1000000000L : std::max(0L, max_steps - ((long)info[2] - step0));
        if (budget == 0) break;
        int rc = tpp_run_to_write(s, budget);
        if (rc < 0) return rc;
        check(tpp_info(s, info), "tpp_info");
        if (np > 0 && !cs.probeCells.empty()) {
            rows.resize((size_t)(1 + np) * 4096);
            for (long n; (n = tpp_probe_log(s, rows.data(), 4096)) > 0;) probeRows(cs, rows.data(), n, 1 + np);
        }
        if (rc != 1) break;
        caseWrite(s);
        if (iface.f) ifaceRow();
        if (verbose) {
            double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("Time = %s  step %ld  deltaT = %.6g  Co = %.3g  p_rghFinal iters %d res %.2e  ExecutionTime = %.2f s\n", caseio::timeName(info[0], cs.cfg.timePrecision).c_str(), (long)info[2], info[1],
                   info[3], (int)info[8], info[10], el);
            fflush(stdout);
        }
    }
    check(tpp_info(s, info), "tpp_info");
    return (long)info[2] - step0;
} CASE_CATCH(-4)
}

extern "C" {
// internalField of an OpenFOAM vol/surface field file (ascii or binary), for hosts that post-process
// time directories without a FoamFile reader (the reference reads alpha.water for its interface
// metric, main.py:727-806).  Returns the number of doubles the field holds (values x components;
// a `uniform` internalField holds one value) and copies up to cap of them; needs no handle.
long tpp_read_field(const char* path, double* out, long cap, int* n_comp, int* uniform) try {
    if (!path) { g_err = "tpp_read_field: null path"; return -1; }
    caseio::Field f = caseio::readField(path);
    if (n_comp) *n_comp = f.nc;
    if (uniform) *uniform = f.internal.uniform;
    const double* src = f.internal.uniform ? f.internal.u : f.internal.a.data();
    long n = f.internal.uniform ? f.nc : (long)f.internal.a.size();
    if (out && cap > 0) memcpy(out, src, (size_t)std::min(n, cap) * sizeof(double));
    return n;
} CASE_CATCH(-4)
}
