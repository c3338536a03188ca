// TEST INFRASTRUCTURE — CPU restatement of the OpenFOAM-13 incompressibleVoF step.
// See vof_oracle.h for scope, provenance and the "parity unpinned" statement.
//
// Citations: `ref:` = file:line under /root/reference that selects/configures the step;
// [OF13-MEM] = OpenFOAM-13 upstream source restated from memory (not available here).
//
// Serial, FP64, face loops in OpenFOAM order (internal faces ascending, then patches), so
// that a cell-gathered GPU kernel that adds its faces in ascending face index reproduces
// every sum bit for bit.  Build with -ffp-contract=off (see Makefile).
#include "vof_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

typedef std::vector<double> dvec;
typedef std::vector<int> ivec;

const double SMALL = 1e-15, VSMALL = 1e-300, ROOTVSMALL = 1e-150, GREAT = 1e15;
std::string g_err;

inline double sign(double x) { return x >= 0 ? 1.0 : -1.0; }
inline double pos0(double x) { return x >= 0 ? 1.0 : 0.0; }
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double mag3(const double* a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
inline void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// ------------------------------------------------------------------------------------
// linear algebra on LDU addressing  [OF13-MEM: lduMatrix, PCG.C, GAMGSolver*.C,
// DICSmoother.C, DICGaussSeidelSmoother.C, GaussSeidelSmoother.C, pairGAMGAgglomerate.C]
// configured at ref: circularSloshingTank/system/fvSolution:42-66
// ------------------------------------------------------------------------------------
struct Ldu {
    int n = 0, nf = 0;
    ivec l, u;  // lower (owner) / upper (neighbour) addressing, upper-triangular order
    ivec ownerStart;
    dvec diag, upper;  // symmetric: lower == upper
    void finalize() {
        ownerStart.assign(n + 1, 0);
        for (int f = 0; f < nf; f++) ownerStart[l[f] + 1]++;
        for (int c = 0; c < n; c++) ownerStart[c + 1] += ownerStart[c];
    }
    void Amul(dvec& y, const dvec& x) const {
        for (int c = 0; c < n; c++) y[c] = diag[c] * x[c];
        for (int f = 0; f < nf; f++) {
            y[u[f]] += upper[f] * x[l[f]];
            y[l[f]] += upper[f] * x[u[f]];
        }
    }
    void sumA(dvec& s) const {
        for (int c = 0; c < n; c++) s[c] = diag[c];
        for (int f = 0; f < nf; f++) {
            s[u[f]] += upper[f];
            s[l[f]] += upper[f];
        }
    }
};

double sumMag(const dvec& a) {
    double s = 0;
    for (double v : a) s += std::fabs(v);
    return s;
}
double sumProd(const dvec& a, const dvec& b) {
    double s = 0;
    for (size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
    return s;
}

// lduMatrix::solver::normFactor
double normFactor(const Ldu& A, const dvec& x, const dvec& b, const dvec& Ax) {
    dvec t(A.n);
    A.sumA(t);
    double xa = 0;
    for (double v : x) xa += v;
    xa /= std::max(A.n, 1);
    double s = 0;
    for (int c = 0; c < A.n; c++) {
        double p = t[c] * xa;
        s += std::fabs(Ax[c] - p) + std::fabs(b[c] - p);
    }
    return s + 1e-20;
}

struct DIC {
    dvec rD;
    void init(const Ldu& A) {
        rD = A.diag;
        for (int f = 0; f < A.nf; f++) rD[A.u[f]] -= A.upper[f] * A.upper[f] / rD[A.l[f]];
        for (int c = 0; c < A.n; c++) rD[c] = 1.0 / rD[c];
    }
    // DICPreconditioner::precondition
    void precondition(const Ldu& A, dvec& w, const dvec& r) const {
        for (int c = 0; c < A.n; c++) w[c] = rD[c] * r[c];
        for (int f = 0; f < A.nf; f++) w[A.u[f]] -= rD[A.u[f]] * A.upper[f] * w[A.l[f]];
        for (int f = A.nf - 1; f >= 0; f--) w[A.l[f]] -= rD[A.l[f]] * A.upper[f] * w[A.u[f]];
    }
    // DICSmoother::smooth (one sweep)
    void sweep(const Ldu& A, dvec& x, const dvec& b, dvec& rA) const {
        A.Amul(rA, x);
        for (int c = 0; c < A.n; c++) rA[c] = (b[c] - rA[c]) * rD[c];
        for (int f = 0; f < A.nf; f++) rA[A.u[f]] -= rD[A.u[f]] * A.upper[f] * rA[A.l[f]];
        for (int f = A.nf - 1; f >= 0; f--) rA[A.l[f]] -= rD[A.l[f]] * A.upper[f] * rA[A.u[f]];
        for (int c = 0; c < A.n; c++) x[c] += rA[c];
    }
};

// GaussSeidelSmoother::smooth (one sweep), owner-face form
void gaussSeidelSweep(const Ldu& A, dvec& x, const dvec& b, dvec& bPrime) {
    bPrime = b;
    for (int c = 0; c < A.n; c++) {
        double xi = bPrime[c];
        int s = A.ownerStart[c], e = A.ownerStart[c + 1];
        for (int f = s; f < e; f++) xi -= A.upper[f] * x[A.u[f]];
        xi /= A.diag[c];
        for (int f = s; f < e; f++) bPrime[A.u[f]] -= A.upper[f] * xi;
        x[c] = xi;
    }
}

struct SolveStats {
    int iters = 0;
    double r0 = 0, r = 0;
};

bool converged(double r, double r0, double tol, double relTol) { return r < tol || (relTol > 0 && r < relTol * r0); }

SolveStats pcgDIC(const Ldu& A, dvec& x, const dvec& b, double tol, double relTol, int maxIter) {
    SolveStats st;
    int n = A.n;
    dvec wA(n), rA(n), pA(n);
    A.Amul(wA, x);
    for (int c = 0; c < n; c++) rA[c] = b[c] - wA[c];
    double nf = normFactor(A, x, b, wA);
    st.r0 = st.r = sumMag(rA) / nf;
    if (converged(st.r, st.r0, tol, relTol)) return st;
    DIC pre;
    pre.init(A);
    double wArA = 0, wArAold;
    do {
        pre.precondition(A, wA, rA);
        wArAold = wArA;
        wArA = sumProd(wA, rA);
        if (st.iters == 0)
            pA = wA;
        else {
            double beta = wArA / wArAold;
            for (int c = 0; c < n; c++) pA[c] = wA[c] + beta * pA[c];
        }
        A.Amul(wA, pA);
        double wApA = sumProd(wA, pA);
        if (std::fabs(wApA) / nf < VSMALL) break;
        double alpha = wArA / wApA;
        for (int c = 0; c < n; c++) {
            x[c] += alpha * pA[c];
            rA[c] -= alpha * wA[c];
        }
        st.r = sumMag(rA) / nf;
    } while (++st.iters < maxIter && !converged(st.r, st.r0, tol, relTol));
    return st;
}

struct Gamg {
    struct Level {
        Ldu A;
        ivec restrictAddr;      // fine cell -> coarse cell (from the previous level)
        ivec faceRestrictAddr;  // fine face -> coarse face, or -1-coarseCell if absorbed
        DIC dic;
    };
    std::vector<Level> lv;  // lv[0] = first coarse level
    const Ldu* fine = nullptr;
    DIC fineDic;
    orc_solver_t ctl;
    dvec faceW0;
    bool haveAgglom = false;

    // pairGAMGAgglomeration::agglomerate  [OF13-MEM]
    static int agglomeratePairs(const Ldu& A, const dvec& w, bool forward, ivec& map) {
        int n = A.n;
        ivec off(n + 1, 0), cf(2 * A.nf);
        for (int f = 0; f < A.nf; f++) {
            off[A.l[f] + 1]++;
            off[A.u[f] + 1]++;
        }
        for (int c = 0; c < n; c++) off[c + 1] += off[c];
        ivec cur(off.begin(), off.end() - 1);
        for (int f = 0; f < A.nf; f++) {
            cf[cur[A.l[f]]++] = f;
            cf[cur[A.u[f]]++] = f;
        }
        map.assign(n, -1);
        int nc = 0;
        for (int ci = 0; ci < n; ci++) {
            int c = forward ? ci : n - ci - 1;
            if (map[c] >= 0) continue;
            int match = -1;
            double wmax = -GREAT;
            for (int k = off[c]; k < off[c + 1]; k++) {
                int f = cf[k];
                if (map[A.u[f]] < 0 && map[A.l[f]] < 0 && w[f] > wmax) {
                    match = f;
                    wmax = w[f];
                }
            }
            if (match >= 0) {
                map[A.u[match]] = nc;
                map[A.l[match]] = nc;
                nc++;
            } else {
                int cm = -1;
                double cw = -GREAT;
                for (int k = off[c]; k < off[c + 1]; k++) {
                    int f = cf[k];
                    if (w[f] > cw) {
                        cm = f;
                        cw = w[f];
                    }
                }
                if (cm >= 0) map[c] = std::max(map[A.u[cm]], map[A.l[cm]]);
            }
        }
        for (int c = 0; c < n; c++)
            if (map[c] < 0) map[c] = nc++;
        if (!forward)
            for (int c = 0; c < n; c++) map[c] = nc - map[c] - 1;
        return nc;
    }

    void buildAgglomeration(const Ldu& A, const dvec& faceWeights) {
        fine = &A;
        lv.clear();
        lv.reserve(64);
        const Ldu* cur = &A;
        dvec w = faceWeights;
        bool forward = true;
        int nCoarsest = ctl.n_cells_coarsest > 0 ? ctl.n_cells_coarsest : 10;
        while ((int)lv.size() < 49) {
            ivec map;
            int nc = agglomeratePairs(*cur, w, forward, map);
            forward = !forward;
            if (nc < nCoarsest || nc >= cur->n) break;
            Level L;
            L.restrictAddr = map;
            // coarse addressing: unique (min,max) pairs in upper-triangular order
            std::vector<std::pair<long long, int>> keys;
            keys.reserve(cur->nf);
            for (int f = 0; f < cur->nf; f++) {
                int a = map[cur->l[f]], b = map[cur->u[f]];
                if (a != b) keys.push_back({(long long)std::min(a, b) * nc + std::max(a, b), f});
            }
            std::sort(keys.begin(), keys.end());
            L.faceRestrictAddr.assign(cur->nf, 0);
            for (int f = 0; f < cur->nf; f++) {
                int a = map[cur->l[f]], b = map[cur->u[f]];
                if (a == b) L.faceRestrictAddr[f] = -1 - a;
            }
            L.A.n = nc;
            long long last = -1;
            for (auto& kf : keys) {
                if (kf.first != last) {
                    L.A.l.push_back((int)(kf.first / nc));
                    L.A.u.push_back((int)(kf.first % nc));
                    last = kf.first;
                }
                L.faceRestrictAddr[kf.second] = (int)L.A.l.size() - 1;
            }
            L.A.nf = (int)L.A.l.size();
            L.A.finalize();
            dvec wc(L.A.nf, 0.0);
            for (int f = 0; f < cur->nf; f++)
                if (L.faceRestrictAddr[f] >= 0) wc[L.faceRestrictAddr[f]] += w[f];
            w.swap(wc);
            lv.push_back(std::move(L));
            cur = &lv.back().A;
            // vector may reallocate: re-point after push
        }
        // fix dangling `cur` pointers is unnecessary: levels are addressed by index below
        haveAgglom = true;
    }

    // GAMGSolver::agglomerateMatrix
    void agglomerateMatrix() {
        const Ldu* f = fine;
        for (size_t i = 0; i < lv.size(); i++) {
            Level& L = lv[i];
            L.A.diag.assign(L.A.n, 0.0);
            L.A.upper.assign(L.A.nf, 0.0);
            for (int c = 0; c < f->n; c++) L.A.diag[L.restrictAddr[c]] += f->diag[c];
            for (int k = 0; k < f->nf; k++) {
                int a = L.faceRestrictAddr[k];
                if (a >= 0)
                    L.A.upper[a] += f->upper[k];
                else
                    L.A.diag[-1 - a] += 2 * f->upper[k];
            }
            f = &L.A;
        }
        fineDic.init(*fine);
        for (auto& L : lv) L.dic.init(L.A);
    }

    void smooth(const Ldu& A, const DIC& dic, dvec& x, const dvec& b, int sweeps) const {
        dvec tmp(A.n);
        for (int s = 0; s < sweeps; s++) {
            if (ctl.smoother == 0)
                dic.sweep(A, x, b, tmp);
            else if (ctl.smoother == 1) {
                dic.sweep(A, x, b, tmp);
                gaussSeidelSweep(A, x, b, tmp);
            } else
                gaussSeidelSweep(A, x, b, tmp);
        }
    }

    // GAMGSolver::scale
    static void scale(const Ldu& A, dvec& field, dvec& Acf, const dvec& source) {
        A.Amul(Acf, field);
        double num = 0, den = 0;
        for (int c = 0; c < A.n; c++) {
            num += source[c] * field[c];
            den += Acf[c] * field[c];
        }
        double d = std::fabs(den) < VSMALL ? (den >= 0 ? VSMALL : -VSMALL) : den;
        double sf = num / d;
        for (int c = 0; c < A.n; c++) field[c] = sf * field[c] + (source[c] - sf * Acf[c]) / A.diag[c];
    }

    // GAMGSolver::Vcycle
    void Vcycle(dvec& psi, const dvec& source, dvec& finestResidual) const {
        int nl = (int)lv.size();
        if (nl == 0) {
            smooth(*fine, fineDic, psi, source, ctl.n_finest_sweeps);
            return;
        }
        int coarsest = nl - 1;
        std::vector<dvec> corr(nl), src(nl);
        for (int i = 0; i < nl; i++) {
            corr[i].assign(lv[i].A.n, 0.0);
            src[i].assign(lv[i].A.n, 0.0);
        }
        for (int c = 0; c < fine->n; c++) src[0][lv[0].restrictAddr[c]] += finestResidual[c];
        std::vector<dvec> pre(nl);
        for (int l = 0; l < coarsest; l++) {
            if (ctl.n_pre_sweeps) {
                std::fill(corr[l].begin(), corr[l].end(), 0.0);
                smooth(lv[l].A, lv[l].dic, corr[l], src[l], ctl.n_pre_sweeps);
                dvec ACf(lv[l].A.n);
                if (l < coarsest - 1) scale(lv[l].A, corr[l], ACf, src[l]);
                lv[l].A.Amul(ACf, corr[l]);
                for (int c = 0; c < lv[l].A.n; c++) src[l][c] -= ACf[c];
            }
            std::fill(src[l + 1].begin(), src[l + 1].end(), 0.0);
            for (int c = 0; c < lv[l].A.n; c++) src[l + 1][lv[l + 1].restrictAddr[c]] += src[l][c];
        }
        // coarsest level: PCG/DIC to the solver's own tolerance
        std::fill(corr[coarsest].begin(), corr[coarsest].end(), 0.0);
        pcgDIC(lv[coarsest].A, corr[coarsest], src[coarsest], ctl.tolerance, ctl.rel_tol, 1000);
        for (int l = coarsest - 1; l >= 0; l--) {
            dvec preSm;
            if (ctl.n_pre_sweeps) preSm = corr[l];
            for (int c = 0; c < lv[l].A.n; c++) corr[l][c] = corr[l + 1][lv[l + 1].restrictAddr[c]];
            dvec ACf(lv[l].A.n);
            if (l < coarsest - 1) scale(lv[l].A, corr[l], ACf, src[l]);
            if (ctl.n_pre_sweeps)
                for (int c = 0; c < lv[l].A.n; c++) corr[l][c] += preSm[c];
            smooth(lv[l].A, lv[l].dic, corr[l], src[l], ctl.n_post_sweeps);
        }
        dvec fc(fine->n), Apsi(fine->n);
        for (int c = 0; c < fine->n; c++) fc[c] = corr[0][lv[0].restrictAddr[c]];
        scale(*fine, fc, Apsi, finestResidual);
        for (int c = 0; c < fine->n; c++) psi[c] += fc[c];
        smooth(*fine, fineDic, psi, source, ctl.n_finest_sweeps);
    }

    SolveStats solve(dvec& psi, const dvec& source) const {
        SolveStats st;
        const Ldu& A = *fine;
        dvec Apsi(A.n), res(A.n);
        A.Amul(Apsi, psi);
        double nf = normFactor(A, psi, source, Apsi);
        for (int c = 0; c < A.n; c++) res[c] = source[c] - Apsi[c];
        st.r0 = st.r = sumMag(res) / nf;
        if (converged(st.r, st.r0, ctl.tolerance, ctl.rel_tol)) return st;
        do {
            Vcycle(psi, source, res);
            A.Amul(Apsi, psi);
            for (int c = 0; c < A.n; c++) res[c] = source[c] - Apsi[c];
            st.r = sumMag(res) / nf;
        } while (++st.iters < ctl.max_iter && !converged(st.r, st.r0, ctl.tolerance, ctl.rel_tol));
        return st;
    }

    // GAMGPreconditioner::precondition
    void precondition(dvec& wA, const dvec& rA) const {
        const Ldu& A = *fine;
        std::fill(wA.begin(), wA.end(), 0.0);
        dvec res = rA, AwA(A.n);
        for (int cyc = 0; cyc < ctl.n_vcycles; cyc++) {
            Vcycle(wA, rA, res);
            if (cyc < ctl.n_vcycles - 1) {
                A.Amul(AwA, wA);
                for (int c = 0; c < A.n; c++) res[c] = rA[c] - AwA[c];
            }
        }
    }
};

SolveStats pcgGamg(const Ldu& A, const Gamg& G, dvec& x, const dvec& b, double tol, double relTol, int maxIter) {
    SolveStats st;
    int n = A.n;
    dvec wA(n), rA(n), pA(n);
    A.Amul(wA, x);
    for (int c = 0; c < n; c++) rA[c] = b[c] - wA[c];
    double nf = normFactor(A, x, b, wA);
    st.r0 = st.r = sumMag(rA) / nf;
    if (converged(st.r, st.r0, tol, relTol)) return st;
    double wArA = 0, wArAold;
    do {
        G.precondition(wA, rA);
        wArAold = wArA;
        wArA = sumProd(wA, rA);
        if (st.iters == 0)
            pA = wA;
        else {
            double beta = wArA / wArAold;
            for (int c = 0; c < n; c++) pA[c] = wA[c] + beta * pA[c];
        }
        A.Amul(wA, pA);
        double wApA = sumProd(wA, pA);
        if (std::fabs(wApA) / nf < VSMALL) break;
        double alpha = wArA / wApA;
        for (int c = 0; c < n; c++) {
            x[c] += alpha * pA[c];
            rA[c] -= alpha * wA[c];
        }
        st.r = sumMag(rA) / nf;
    } while (++st.iters < maxIter && !converged(st.r, st.r0, tol, relTol));
    return st;
}

}  // namespace

// ------------------------------------------------------------------------------------
// solver state
// ------------------------------------------------------------------------------------
struct orc_state {
    // mesh
    int nP, nF, nI, nC, nB, nPatch;
    dvec points0, points, pointsOld;
    ivec fOff, fLab, own, nei;
    ivec pStart, pSize, bcU, bcA, bcP;
    dvec pInletAlpha, pP0;
    ivec facePatch;  // boundary face (index - nI) -> patch
    orc_config_t cfg;
    dvec motion;
    // geometry
    dvec Cf, Sf, magSf, C, V, V0, w, dc, corrVec, CfOld;
    dvec meshPhi;
    double deltaN = 0;
    // experiment switches for the [OF13-MEM] choices that cannot be checked against upstream source
    // (environment switches read at create time, used with tools/validate_run.py; results in profiles/r2_physics/README.md;
    // all default to the restatement documented in DESIGN.md §2)
    int xClip = 0;       // ORC_X_CLIP=1: clip the compressed face value to [0,1]
    int xOwn = 0;        // ORC_X_OWN=1: MULES local extrema include the cell's own value
    int xBndExt = 0;     // ORC_X_BND=1: boundary face values enter the MULES extrema
    double xDdt = -1;    // ORC_X_DDT=c: fixed ddtCorr coupling coefficient c instead of 1 - min(|corr|/|phi|, 1)
    int xNoRelax = 0;    // ORC_X_NORELAX=1: skip the diagonal-dominance step of UEqn.relax(1)
    // fields
    dvec alpha, alpha_b, U, U_b, p_rgh, p_rgh_b, p, rho, rho_b, phi, Uf;
    dvec U0, U0_b, rho0, Uf0;
    dvec alphaPhi, rhoPhi;
    dvec pGrad_b;  // fixedFluxPressure gradient
    // alpha predictor intermediates (last sub-cycle)
    dvec gradAlpha, alphaPhiUn, phiBD, phiCorr, lambda;
    // surface tension (sigma > 0 only; an extension: the reference runs sigma 0, constant/phaseProperties:19)
    dvec gradA, nHatf, sigmaK, stf;
    // momentum
    dvec gradU, mLower, mUpper, mDiag, mSource, mBIC, mBBC;
    // pressure
    dvec rAU, HbyA, HbyA_b, rAUf, phiHbyA, phig, pUpper, pDiag, pSource, gradRho, gradP, pCorrFlux;
    // time
    double t = 0, dt = 0, dt0 = 0, startTime = 0;
    long step = 0;
    int writeTimeIndex = 0;
    double Co = 0, alphaCo = 0;
    bool needRef = false;
    int refCell = -1;
    SolveStats lastSolve[2];
    // solver
    Ldu A;
    Gamg gamg[2];
    // probes
    ivec probeCells;
    dvec probeLog;
    std::map<std::string, dvec*> reg;

    // --------------------------------------------------------------------------------
    void registerFields() {
#define R(x) reg[#x] = &x
        R(points); R(Cf); R(Sf); R(magSf); R(C); R(V); R(V0); R(w); R(dc); R(corrVec); R(meshPhi);
        R(alpha); R(alpha_b); R(U); R(U_b); R(p_rgh); R(p_rgh_b); R(p); R(rho); R(rho_b); R(phi); R(Uf);
        R(U0); R(U0_b); R(rho0); R(Uf0); R(alphaPhi); R(rhoPhi); R(pGrad_b);
        R(gradAlpha); R(alphaPhiUn); R(phiBD); R(phiCorr); R(lambda);
        R(gradA); R(nHatf); R(sigmaK); R(stf);
        R(gradU); R(mLower); R(mUpper); R(mDiag); R(mSource); R(mBIC); R(mBBC);
        R(rAU); R(HbyA); R(HbyA_b); R(rAUf); R(phiHbyA); R(phig); R(pUpper); R(pDiag); R(pSource);
        R(gradRho); R(gradP); R(pCorrFlux);
#undef R
    }

    // ---- motion: Function1s::Table (linear, clamped) + sixDoFMotion [OF13-MEM] ----------
    // ref: circularSloshingTank/constant/dynamicMeshDict:17-44, generate_motion.py:13-42
    void motionAt(double time, double R[9], double T[3]) const {
        double v[6] = {0, 0, 0, 0, 0, 0};
        int n = cfg.n_motion;
        if (n > 0) {
            const double* m = motion.data();
            if (time <= m[0])
                for (int k = 0; k < 6; k++) v[k] = m[1 + k];
            else if (time >= m[7 * (n - 1)])
                for (int k = 0; k < 6; k++) v[k] = m[7 * (n - 1) + 1 + k];
            else {
                int lo = 0, hi = n - 1;
                while (hi - lo > 1) {
                    int mid = (lo + hi) / 2;
                    if (m[7 * mid] <= time) lo = mid; else hi = mid;
                }
                double s = (time - m[7 * lo]) / (m[7 * hi] - m[7 * lo]);
                for (int k = 0; k < 6; k++) v[k] = m[7 * lo + 1 + k] + s * (m[7 * hi + 1 + k] - m[7 * lo + 1 + k]);
            }
        }
        const double d2r = M_PI / 180.0;
        double ax = v[3] * d2r, ay = v[4] * d2r, az = v[5] * d2r;
        double cx = cos(ax), sx = sin(ax), cy = cos(ay), sy = sin(ay), cz = cos(az), sz = sin(az);
        // R = Rx * Ry * Rz
        R[0] = cy * cz;               R[1] = -cy * sz;              R[2] = sy;
        R[3] = sx * sy * cz + cx * sz; R[4] = -sx * sy * sz + cx * cz; R[5] = -sx * cy;
        R[6] = -cx * sy * cz + sx * sz; R[7] = cx * sy * sz + sx * cz;  R[8] = cx * cy;
        T[0] = v[0]; T[1] = v[1]; T[2] = v[2];
    }
    void transformPoints(double time, dvec& out) const {
        double R[9], T[3];
        motionAt(time, R, T);
        const double* c = cfg.cofg;
        out.resize(3 * nP);
        for (int i = 0; i < nP; i++) {
            double q[3] = {points0[3 * i] - c[0], points0[3 * i + 1] - c[1], points0[3 * i + 2] - c[2]};
            for (int k = 0; k < 3; k++) out[3 * i + k] = (R[3 * k] * q[0] + R[3 * k + 1] * q[1] + R[3 * k + 2] * q[2]) + c[k] + T[k];
        }
    }

    // ---- geometry [OF13-MEM: primitiveMeshFaceCentresAndAreas.C, ...CellCentresAndVols.C,
    //      surfaceInterpolation.C (weights, nonOrthDeltaCoeffs, nonOrthCorrectionVectors)] ----
    void calcGeometry() {
        Cf.assign(3 * nF, 0); Sf.assign(3 * nF, 0); magSf.assign(nF, 0);
        const double* P = points.data();
        for (int f = 0; f < nF; f++) {
            int s = fOff[f], n = fOff[f + 1] - s;
            const int* l = &fLab[s];
            if (n == 3) {
                const double *a = P + 3 * l[0], *b = P + 3 * l[1], *c = P + 3 * l[2];
                double e1[3], e2[3], nn[3];
                for (int k = 0; k < 3; k++) {
                    Cf[3 * f + k] = (1.0 / 3.0) * (a[k] + b[k] + c[k]);
                    e1[k] = b[k] - a[k];
                    e2[k] = c[k] - a[k];
                }
                cross3(e1, e2, nn);
                for (int k = 0; k < 3; k++) Sf[3 * f + k] = 0.5 * nn[k];
            } else {
                double fc[3] = {0, 0, 0}, sumN[3] = {0, 0, 0}, sumA = 0, sumAc[3] = {0, 0, 0};
                for (int i = 0; i < n; i++)
                    for (int k = 0; k < 3; k++) fc[k] += P[3 * l[i] + k];
                for (int k = 0; k < 3; k++) fc[k] /= n;
                for (int i = 0; i < n; i++) {
                    const double *a = P + 3 * l[i], *b = P + 3 * l[(i + 1) % n];
                    double c[3], e1[3], e2[3], nn[3];
                    for (int k = 0; k < 3; k++) {
                        c[k] = a[k] + b[k] + fc[k];
                        e1[k] = b[k] - a[k];
                        e2[k] = fc[k] - a[k];
                    }
                    cross3(e1, e2, nn);
                    double an = mag3(nn);
                    sumA += an;
                    for (int k = 0; k < 3; k++) {
                        sumN[k] += nn[k];
                        sumAc[k] += an * c[k];
                    }
                }
                for (int k = 0; k < 3; k++) {
                    Cf[3 * f + k] = sumA < ROOTVSMALL ? fc[k] : (1.0 / 3.0) * sumAc[k] / sumA;
                    Sf[3 * f + k] = 0.5 * sumN[k];
                }
            }
            magSf[f] = mag3(&Sf[3 * f]);
        }
        dvec cEst(3 * nC, 0.0);
        ivec nCF(nC, 0);
        for (int f = 0; f < nF; f++) {
            for (int k = 0; k < 3; k++) cEst[3 * own[f] + k] += Cf[3 * f + k];
            nCF[own[f]]++;
        }
        for (int f = 0; f < nI; f++) {
            for (int k = 0; k < 3; k++) cEst[3 * nei[f] + k] += Cf[3 * f + k];
            nCF[nei[f]]++;
        }
        for (int c = 0; c < nC; c++)
            for (int k = 0; k < 3; k++) cEst[3 * c + k] /= nCF[c];
        C.assign(3 * nC, 0.0);
        V.assign(nC, 0.0);
        for (int f = 0; f < nF; f++) {
            int o = own[f];
            double d[3];
            for (int k = 0; k < 3; k++) d[k] = Cf[3 * f + k] - cEst[3 * o + k];
            double pyr = dot3(&Sf[3 * f], d);
            for (int k = 0; k < 3; k++) C[3 * o + k] += pyr * (0.75 * Cf[3 * f + k] + 0.25 * cEst[3 * o + k]);
            V[o] += pyr;
        }
        for (int f = 0; f < nI; f++) {
            int n = nei[f];
            double d[3];
            for (int k = 0; k < 3; k++) d[k] = cEst[3 * n + k] - Cf[3 * f + k];
            double pyr = dot3(&Sf[3 * f], d);
            for (int k = 0; k < 3; k++) C[3 * n + k] += pyr * (0.75 * Cf[3 * f + k] + 0.25 * cEst[3 * n + k]);
            V[n] += pyr;
        }
        for (int c = 0; c < nC; c++) {
            for (int k = 0; k < 3; k++) C[3 * c + k] = std::fabs(V[c]) > VSMALL ? C[3 * c + k] / V[c] : cEst[3 * c + k];
            V[c] *= (1.0 / 3.0);
        }
        w.assign(nF, 1.0);
        dc.assign(nF, 0.0);
        corrVec.assign(3 * nF, 0.0);
        for (int f = 0; f < nF; f++) {
            const double* S = &Sf[3 * f];
            double nf[3] = {S[0] / magSf[f], S[1] / magSf[f], S[2] / magSf[f]};
            if (f < nI) {
                double dO[3], dN[3], d[3];
                for (int k = 0; k < 3; k++) {
                    dO[k] = Cf[3 * f + k] - C[3 * own[f] + k];
                    dN[k] = C[3 * nei[f] + k] - Cf[3 * f + k];
                    d[k] = C[3 * nei[f] + k] - C[3 * own[f] + k];
                }
                double so = std::fabs(dot3(S, dO)), sn = std::fabs(dot3(S, dN));
                w[f] = sn / (so + sn);
                dc[f] = 1.0 / std::max(dot3(nf, d), 0.05 * mag3(d));
                for (int k = 0; k < 3; k++) corrVec[3 * f + k] = nf[k] - d[k] * dc[f];
            } else {
                // fvPatch::delta(): patch-normal delta for non-coupled patches
                double d[3];
                for (int k = 0; k < 3; k++) d[k] = Cf[3 * f + k] - C[3 * own[f] + k];
                double nd = dot3(nf, d);
                double dl[3] = {nf[0] * nd, nf[1] * nd, nf[2] * nd};
                dc[f] = 1.0 / std::max(dot3(nf, dl), 0.05 * mag3(dl));
            }
        }
    }

    // swept volume of a linearly moving triangle (exact for the ruled prism)
    static double triSwept(const double* a, const double* b, const double* c, const double* a1, const double* b1, const double* c1) {
        double e1[3], e2[3], g1[3], g2[3], dm[3];
        for (int k = 0; k < 3; k++) {
            e1[k] = b[k] - a[k];
            e2[k] = c[k] - a[k];
            double da = a1[k] - a[k], db = b1[k] - b[k], dcc = c1[k] - c[k];
            g1[k] = db - da;
            g2[k] = dcc - da;
            dm[k] = (1.0 / 3.0) * (da + db + dcc);
        }
        double A0[3], A1a[3], A1b[3], A2[3];
        cross3(e1, e2, A0);
        cross3(e1, g2, A1a);
        cross3(g1, e2, A1b);
        cross3(g1, g2, A2);
        double s = 0;
        for (int k = 0; k < 3; k++) s += dm[k] * (0.5 * A0[k] + 0.25 * (A1a[k] + A1b[k]) + (1.0 / 6.0) * A2[k]);
        return s;
    }

    // fvMesh::movePoints -> meshPhi = sweptVol/deltaT [OF13-MEM: face::sweptVol]
    void moveMesh() {
        if (cfg.n_motion <= 0) {
            meshPhi.assign(nF, 0.0);
            V0 = V;
            CfOld = Cf;
            return;
        }
        pointsOld = points;
        CfOld = Cf;
        V0 = V;
        transformPoints(t, points);
        dvec CfO = Cf;
        calcGeometry();
        meshPhi.assign(nF, 0.0);
        const double *P0 = pointsOld.data(), *P1 = points.data();
        for (int f = 0; f < nF; f++) {
            int s = fOff[f], n = fOff[f + 1] - s;
            const int* l = &fLab[s];
            double sv = 0;
            if (n == 3)
                sv = triSwept(P0 + 3 * l[0], P0 + 3 * l[1], P0 + 3 * l[2], P1 + 3 * l[0], P1 + 3 * l[1], P1 + 3 * l[2]);
            else
                for (int i = 0; i < n; i++) {
                    int j = (i + 1) % n;
                    sv += triSwept(&CfO[3 * f], P0 + 3 * l[i], P0 + 3 * l[j], &Cf[3 * f], P1 + 3 * l[i], P1 + 3 * l[j]);
                }
            meshPhi[f] = sv / dt;
        }
    }

    // ---- boundary conditions (ref: circularSloshingTank/0/{U,alpha.water,p_rgh}:22-31) ----
    // inletOutlet / zeroGradient  [OF13-MEM: inletOutletFvPatchField]
    void alphaBCs() {
        for (int b = 0; b < nB; b++) {
            int f = nI + b, pt = facePatch[b];
            if (bcA[pt] == ORC_A_INLET_OUTLET)
                alpha_b[b] = phi[f] >= 0 ? alpha[own[f]] : pInletAlpha[pt];  // valueFraction = 1 - pos0(phi)
            else
                alpha_b[b] = alpha[own[f]];
        }
    }
    // movingWallVelocity / pressureInletOutletVelocity [OF13-MEM]
    void UBCs() {
        for (int b = 0; b < nB; b++) {
            int f = nI + b, pt = facePatch[b], c = own[f];
            double n[3] = {Sf[3 * f] / magSf[f], Sf[3 * f + 1] / magSf[f], Sf[3 * f + 2] / magSf[f]};
            if (bcU[pt] == ORC_U_MOVING_WALL) {
                if (cfg.n_motion > 0) {
                    double Up[3];
                    for (int k = 0; k < 3; k++) Up[k] = (Cf[3 * f + k] - CfOld[3 * f + k]) / dt;
                    double Un = meshPhi[f] / (magSf[f] + VSMALL);
                    double nUp = dot3(n, Up);
                    for (int k = 0; k < 3; k++) U_b[3 * b + k] = Up[k] + n[k] * (Un - nUp);
                }
            } else {
                // directionMixed: refValue 0, refGrad 0, valueFraction = neg(phi)*(I - nn)
                const double* Uc = &U[3 * c];
                if (phi[f] < 0) {
                    double nu = dot3(n, Uc);
                    for (int k = 0; k < 3; k++) U_b[3 * b + k] = n[k] * nu;  // (I - vf) & Uc = nn & Uc
                } else
                    for (int k = 0; k < 3; k++) U_b[3 * b + k] = Uc[k];
            }
        }
    }
    // totalPressure: p = p0 - 0.5 rho (1 - pos0(phi)) |U|^2  [OF13-MEM]
    void pTotalPressure() {
        for (int b = 0; b < nB; b++) {
            int f = nI + b, pt = facePatch[b];
            if (bcP[pt] == ORC_P_TOTAL_PRESSURE) {
                const double* u = &U_b[3 * b];
                p_rgh_b[b] = pP0[pt] - 0.5 * rho_b[b] * (1.0 - pos0(phi[f])) * dot3(u, u);
            }
        }
    }
    void pEvaluate() {
        for (int b = 0; b < nB; b++) {
            int f = nI + b, pt = facePatch[b];
            if (bcP[pt] == ORC_P_FIXED_FLUX) p_rgh_b[b] = p_rgh[own[f]] + pGrad_b[b] / dc[f];
        }
    }

    // ---- Gauss linear gradient of a scalar with boundary values ------------------------
    void gradScalar(const dvec& s, const dvec& sb, dvec& g) const {
        g.assign(3 * nC, 0.0);
        for (int f = 0; f < nI; f++) {
            double sf = w[f] * s[own[f]] + (1.0 - w[f]) * s[nei[f]];
            for (int k = 0; k < 3; k++) {
                double v = Sf[3 * f + k] * sf;
                g[3 * own[f] + k] += v;
                g[3 * nei[f] + k] -= v;
            }
        }
        for (int f = nI; f < nF; f++)
            for (int k = 0; k < 3; k++) g[3 * own[f] + k] += Sf[3 * f + k] * sb[f - nI];
        for (int c = 0; c < nC; c++)
            for (int k = 0; k < 3; k++) g[3 * c + k] /= V[c];
    }

    // ---- S0: Courant numbers [OF13-MEM: fluidSolver::correctCoNum, twoPhaseSolver] -------
    void courant() {
        dvec sumPhi(nC, 0.0);
        for (int f = 0; f < nI; f++) {
            double m = std::fabs(phi[f]);
            sumPhi[own[f]] += m;
            sumPhi[nei[f]] += m;
        }
        for (int f = nI; f < nF; f++) sumPhi[own[f]] += std::fabs(phi[f]);
        double mx = 0, mxa = 0;
        for (int c = 0; c < nC; c++) {
            double v = sumPhi[c] / V[c];
            mx = std::max(mx, v);
            double near = pos0(alpha[c] - 0.01) * pos0(0.99 - alpha[c]);
            mxa = std::max(mxa, near * sumPhi[c] / V[c]);
        }
        Co = 0.5 * mx * dt;
        alphaCo = 0.5 * mxa * dt;
    }

    // ---- S1: adjustDeltaT + Time::adjustDeltaT (ref: system/controlDict:27-31,47-51) ----
    void adjustDeltaT() {
        if (!cfg.adjust_time_step) return;
        double d = cfg.max_delta_t;
        if (Co > SMALL) d = std::min(d, cfg.max_co / Co * dt);
        if (alphaCo > SMALL) d = std::min(d, cfg.max_alpha_co / alphaCo * dt);
        dt = std::min(1.2 * dt, d);
        timeAdjustDeltaT();
    }
    void timeAdjustDeltaT() {
        double timeToNextWrite = std::max(0.0, (writeTimeIndex + 1) * cfg.write_interval - (t - startTime));
        double nSteps = timeToNextWrite / dt - SMALL;
        if (nSteps < 2147483647.0) {
            int n = (int)nSteps + 1;
            double nd = timeToNextWrite / n;
            if (nd >= dt) dt = std::min(nd, 2.0 * dt);
            else dt = std::max(nd, 0.2 * dt);
        }
    }
    // Time::operator++ ; returns true when the new time is a write time
    bool advanceTime() {
        dt0 = dt;
        t += dt;
        step++;
        U0 = U; U0_b = U_b; rho0 = rho; Uf0 = Uf;
        int wi = (int)(((t - startTime) + 0.5 * dt) / cfg.write_interval);
        if (wi > writeTimeIndex) {
            writeTimeIndex = wi;
            return true;
        }
        return false;
    }

    // ---- S3: alphaPredictor (ref: system/fvSolution:19-23, system/fvSchemes:30) -----------
    static double vanLeerLimiter(double flux, double pP, double pN, const double* gP, const double* gN, const double* d) {
        double gradf = pN - pP;
        double gradcf = flux > 0 ? dot3(d, gP) : dot3(d, gN);
        double r;
        if (std::fabs(gradcf) >= 1000 * std::fabs(gradf)) r = 2 * 1000 * sign(gradcf) * sign(gradf) - 1;
        else r = 2 * (gradcf / gradf) - 1;
        return (r + std::fabs(r)) / (1 + std::fabs(r));
    }

    void alphaSubCycle(double dts) {
        const double rDeltaT = 1.0 / dts;
        dvec alpha0 = alpha;
        alphaBCs();
        gradScalar(alpha, alpha_b, gradAlpha);
        alphaPhiUn.assign(nF, 0.0); phiBD.assign(nF, 0.0); phiCorr.assign(nF, 0.0);
        // interfaceCompression(vanLeer, cAlpha) face value and flux [OF13-MEM: interfaceCompression.C]
        for (int f = 0; f < nI; f++) {
            int P = own[f], N = nei[f];
            double d[3] = {C[3 * N] - C[3 * P], C[3 * N + 1] - C[3 * P + 1], C[3 * N + 2] - C[3 * P + 2]};
            double lim = vanLeerLimiter(phi[f], alpha[P], alpha[N], &gradAlpha[3 * P], &gradAlpha[3 * N], d);
            double wf = lim * w[f] + (1.0 - lim) * pos0(phi[f]);
            double vf = wf * alpha[P] + (1.0 - wf) * alpha[N];
            double gf[3];
            for (int k = 0; k < 3; k++) gf[k] = w[f] * gradAlpha[3 * P + k] + (1.0 - w[f]) * gradAlpha[3 * N + k];
            double mg = mag3(gf) + deltaN;
            double nHatf = (gf[0] / mg) * Sf[3 * f] + (gf[1] / mg) * Sf[3 * f + 1] + (gf[2] / mg) * Sf[3 * f + 2];
            vf += cfg.c_alpha * sign(phi[f]) * vf * (1.0 - vf) * nHatf / magSf[f];
            if (xClip) vf = std::min(std::max(vf, 0.0), 1.0);
            alphaPhiUn[f] = phi[f] * vf;
            phiBD[f] = phi[f] * (phi[f] >= 0 ? alpha[P] : alpha[N]);
            phiCorr[f] = alphaPhiUn[f] - phiBD[f];
        }
        for (int f = nI; f < nF; f++) {
            alphaPhiUn[f] = phi[f] * alpha_b[f - nI];
            phiBD[f] = alphaPhiUn[f];
            phiCorr[f] = 0.0;
        }
        // MULES::limiter [OF13-MEM: MULESTemplates.C]
        const double psiMax = 1.0, psiMin = 0.0;
        dvec psiMaxn(nC, psiMin), psiMinn(nC, psiMax), sumPhiBD(nC, 0.0), sumPhip(nC, 0.0), mSumPhim(nC, 0.0);
        for (int f = 0; f < nI; f++) {
            int P = own[f], N = nei[f];
            psiMaxn[P] = std::max(psiMaxn[P], alpha[N]);
            psiMinn[P] = std::min(psiMinn[P], alpha[N]);
            psiMaxn[N] = std::max(psiMaxn[N], alpha[P]);
            psiMinn[N] = std::min(psiMinn[N], alpha[P]);
            sumPhiBD[P] += phiBD[f];
            sumPhiBD[N] -= phiBD[f];
            double pc = phiCorr[f];
            if (pc > 0) { sumPhip[P] += pc; mSumPhim[N] += pc; }
            else { mSumPhim[P] -= pc; sumPhip[N] -= pc; }
        }
        if (xOwn)
            for (int c = 0; c < nC; c++) {
                psiMaxn[c] = std::max(psiMaxn[c], alpha[c]);
                psiMinn[c] = std::min(psiMinn[c], alpha[c]);
            }
        for (int f = nI; f < nF; f++) {
            // neither zeroGradient nor inletOutlet fixesValue(): no boundary extrema are added
            int P = own[f];
            if (xBndExt) {
                psiMaxn[P] = std::max(psiMaxn[P], alpha_b[f - nI]);
                psiMinn[P] = std::min(psiMinn[P], alpha_b[f - nI]);
            }
            sumPhiBD[P] += phiBD[f];
            double pc = phiCorr[f];
            if (pc > 0) sumPhip[P] += pc; else mSumPhim[P] -= pc;
        }
        for (int c = 0; c < nC; c++) {
            psiMaxn[c] = std::min(psiMaxn[c], psiMax);
            psiMinn[c] = std::max(psiMinn[c], psiMin);
            // mesh.moving() branch (Vsc0 = V0, Vsc = V for a rigid body)
            psiMaxn[c] = V[c] * (rDeltaT * psiMaxn[c]) - (V0[c] * rDeltaT) * alpha0[c] + sumPhiBD[c];
            psiMinn[c] = V[c] * (0.0 - rDeltaT * psiMinn[c]) + (V0[c] * rDeltaT) * alpha0[c] - sumPhiBD[c];
        }
        lambda.assign(nF, 1.0);
        dvec sumlPhip(nC), mSumlPhim(nC);
        for (int j = 0; j < cfg.n_limiter_iter; j++) {
            std::fill(sumlPhip.begin(), sumlPhip.end(), 0.0);
            std::fill(mSumlPhim.begin(), mSumlPhim.end(), 0.0);
            for (int f = 0; f < nI; f++) {
                int P = own[f], N = nei[f];
                double lp = lambda[f] * phiCorr[f];
                if (lp > 0) { sumlPhip[P] += lp; mSumlPhim[N] += lp; }
                else { mSumlPhim[P] -= lp; sumlPhip[N] -= lp; }
            }
            for (int f = nI; f < nF; f++) {
                int P = own[f];
                double lp = lambda[f] * phiCorr[f];
                if (lp > 0) sumlPhip[P] += lp; else mSumlPhim[P] -= lp;
            }
            for (int c = 0; c < nC; c++) {
                sumlPhip[c] = std::max(std::min((sumlPhip[c] + psiMaxn[c]) / (mSumPhim[c] + ROOTVSMALL), 1.0), 0.0);
                mSumlPhim[c] = std::max(std::min((mSumlPhim[c] + psiMinn[c]) / (sumPhip[c] + ROOTVSMALL), 1.0), 0.0);
            }
            const dvec &lambdam = sumlPhip, &lambdap = mSumlPhim;
            for (int f = 0; f < nI; f++) {
                if (phiCorr[f] > 0) lambda[f] = std::min(lambda[f], std::min(lambdap[own[f]], lambdam[nei[f]]));
                else lambda[f] = std::min(lambda[f], std::min(lambdam[own[f]], lambdap[nei[f]]));
            }
            // non-coupled boundary faces carry phiCorr = 0: their lambda never matters
        }
        // phiPsi = phiBD + lambda*phiCorr ; MULES::explicitSolve
        dvec div(nC, 0.0);
        for (int f = 0; f < nI; f++) {
            alphaPhiUn[f] = phiBD[f] + lambda[f] * phiCorr[f];
            div[own[f]] += alphaPhiUn[f];
            div[nei[f]] -= alphaPhiUn[f];
        }
        for (int f = nI; f < nF; f++) {
            alphaPhiUn[f] = phiBD[f] + lambda[f] * phiCorr[f];
            div[own[f]] += alphaPhiUn[f];
        }
        for (int c = 0; c < nC; c++) {
            double psiIf = div[c] / V[c];
            alpha[c] = (V0[c] * alpha0[c] * rDeltaT / V[c] - psiIf) / rDeltaT;
        }
        alphaBCs();
    }

    void mixtureCorrect() {
        for (int c = 0; c < nC; c++) rho[c] = alpha[c] * cfg.rho1 + (1.0 - alpha[c]) * cfg.rho2;
        for (int b = 0; b < nB; b++) rho_b[b] = alpha_b[b] * cfg.rho1 + (1.0 - alpha_b[b]) * cfg.rho2;
    }
    // rho*nu of incompressibleTwoPhaseVoFMixture [OF13-MEM]
    double muOf(double a, double r) const {
        double la = std::min(std::max(a, 0.0), 1.0);
        double mu = la * cfg.rho1 * cfg.nu1 + (1.0 - la) * cfg.rho2 * cfg.nu2;
        double nu = mu / (la * cfg.rho1 + (1.0 - la) * cfg.rho2);
        return r * nu;
    }

    // ---- interfaceProperties::correct + surfaceTensionForce (continuum surface force) ------
    // [OF13-MEM: interfaceProperties::calculateK, surfaceTensionForce]; only when sigma != 0
    // (constant/phaseProperties:19 is 0 in every reference case; 0/alpha.water:22-25 is
    // zeroGradient on the walls, so there is no contact-angle correction of nHat):
    //   gradAlpha = fvc::grad(alpha1)  (Gauss linear; boundary value = cell value with the
    //               normal component replaced by the patch's snGrad, gaussGrad::correctBoundaryConditions)
    //   nHatf = (interpolate(gradAlpha) / (|interpolate(gradAlpha)| + deltaN)) & Sf
    //   K = -fvc::div(nHatf)           (boundary value = the cell's, extrapolatedCalculated)
    //   stf = interpolate(sigma K) * snGrad(alpha1)   (snGradSchemes corrected, system/fvSchemes:47)
    double alphaSnGradBnd(int b) const { return dc[nI + b] * (alpha_b[b] - alpha[own[nI + b]]); }
    void interfaceCorrect() {
        if (cfg.sigma == 0.0) return;
        gradScalar(alpha, alpha_b, gradA);
        nHatf.assign(nF, 0.0); stf.assign(nF, 0.0);
        dvec div(nC, 0.0);
        for (int f = 0; f < nF; f++) {
            int P = own[f];
            double gf[3];
            if (f < nI) {
                int N = nei[f];
                for (int k = 0; k < 3; k++) gf[k] = w[f] * gradA[3 * P + k] + (1.0 - w[f]) * gradA[3 * N + k];
            } else {
                double m = magSf[f];
                double n[3] = {Sf[3 * f] / m, Sf[3 * f + 1] / m, Sf[3 * f + 2] / m};
                double corr = alphaSnGradBnd(f - nI) - dot3(n, &gradA[3 * P]);
                for (int k = 0; k < 3; k++) gf[k] = gradA[3 * P + k] + n[k] * corr;
            }
            double mg = mag3(gf) + deltaN;
            nHatf[f] = (gf[0] / mg) * Sf[3 * f] + (gf[1] / mg) * Sf[3 * f + 1] + (gf[2] / mg) * Sf[3 * f + 2];
            div[P] += nHatf[f];
            if (f < nI) div[nei[f]] -= nHatf[f];
        }
        sigmaK.resize(nC);
        for (int c = 0; c < nC; c++) sigmaK[c] = cfg.sigma * (0.0 - div[c] / V[c]);
        for (int f = 0; f < nF; f++) {
            int P = own[f];
            if (f < nI) {
                int N = nei[f];
                double gf[3];
                for (int k = 0; k < 3; k++) gf[k] = w[f] * gradA[3 * P + k] + (1.0 - w[f]) * gradA[3 * N + k];
                double sn = dc[f] * (alpha[N] - alpha[P]) + dot3(&corrVec[3 * f], gf);
                stf[f] = (w[f] * sigmaK[P] + (1.0 - w[f]) * sigmaK[N]) * sn;
            } else
                stf[f] = sigmaK[P] * alphaSnGradBnd(f - nI);
        }
    }

    void alphaPredictor() {
        int n = cfg.n_alpha_subcycles;
        if (n > 1) {
            double total = dt, dts = dt / n;
            dvec acc(nF, 0.0);
            for (int s = 0; s < n; s++) {
                for (int a = 0; a < cfg.n_alpha_corr; a++) alphaSubCycle(dts);
                for (int f = 0; f < nF; f++) acc[f] += (dts / total) * alphaPhiUn[f];
            }
            alphaPhi = acc;
        } else {
            for (int a = 0; a < cfg.n_alpha_corr; a++) alphaSubCycle(dt);
            alphaPhi = alphaPhiUn;
        }
        mixtureCorrect();
        rhoPhi.resize(nF);
        for (int f = 0; f < nF; f++) rhoPhi[f] = alphaPhi[f] * (cfg.rho1 - cfg.rho2) + phi[f] * cfg.rho2;
        interfaceCorrect();
    }

    // ---- S4: momentum matrix (assembled, never solved: fvSolution:80 momentumPredictor no)
    // ref: system/fvSchemes:19,29,32,37  [OF13-MEM: EulerDdtScheme, gaussConvectionScheme,
    // LimitedScheme<vanLeerV>, gaussLaplacianScheme, linearViscousStress::divDevTau]
    static double vanLeerVLimiter(double flux, const double* uP, const double* uN, const double* gP, const double* gN, const double* d) {
        double gv[3] = {uN[0] - uP[0], uN[1] - uP[1], uN[2] - uP[2]};
        double gradf = dot3(gv, gv);
        const double* g = flux > 0 ? gP : gN;
        double dg[3];  // d & grad: dg_j = d_i g_ij
        for (int j = 0; j < 3; j++) dg[j] = d[0] * g[j] + d[1] * g[3 + j] + d[2] * g[6 + j];
        double gradcf = dot3(gv, dg);
        double r;
        if (std::fabs(gradcf) >= 1000 * std::fabs(gradf)) r = 2 * 1000 * sign(gradcf) * sign(gradf) - 1;
        else r = 2 * (gradcf / gradf) - 1;
        return (r + std::fabs(r)) / (1 + std::fabs(r));
    }

    void gradVector() {
        gradU.assign(9 * nC, 0.0);
        for (int f = 0; f < nI; f++) {
            int P = own[f], N = nei[f];
            double uf[3];
            for (int j = 0; j < 3; j++) uf[j] = w[f] * U[3 * P + j] + (1.0 - w[f]) * U[3 * N + j];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    double v = Sf[3 * f + i] * uf[j];
                    gradU[9 * P + 3 * i + j] += v;
                    gradU[9 * N + 3 * i + j] -= v;
                }
        }
        for (int f = nI; f < nF; f++)
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) gradU[9 * own[f] + 3 * i + j] += Sf[3 * f + i] * U_b[3 * (f - nI) + j];
        for (int c = 0; c < nC; c++)
            for (int k = 0; k < 9; k++) gradU[9 * c + k] /= V[c];
    }

    // mu * dev2(T(gradU)) as a row-major tensor
    static void devTensor(double mu, const double* g, double* T) {
        double tr = g[0] + g[4] + g[8];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) T[3 * i + j] = mu * (g[3 * j + i] - (i == j ? (2.0 / 3.0) * tr : 0.0));
    }

    void momentum() {
        UBCs();
        gradVector();
        const double rDeltaT = 1.0 / dt;
        mLower.assign(nI, 0.0); mUpper.assign(nI, 0.0); mDiag.assign(nC, 0.0); mSource.assign(3 * nC, 0.0);
        mBIC.assign(3 * nB, 0.0); mBBC.assign(3 * nB, 0.0);
        dvec diagSum(nC, 0.0), src(3 * nC, 0.0);
        for (int f = 0; f < nI; f++) {
            int P = own[f], N = nei[f];
            double d[3] = {C[3 * N] - C[3 * P], C[3 * N + 1] - C[3 * P + 1], C[3 * N + 2] - C[3 * P + 2]};
            double F = rhoPhi[f];
            double lim = vanLeerVLimiter(F, &U[3 * P], &U[3 * N], &gradU[9 * P], &gradU[9 * N], d);
            double wf = lim * w[f] + (1.0 - lim) * pos0(F);
            double cl = -wf * F, cu = cl + F;
            double muP = muOf(alpha[P], rho[P]), muN = muOf(alpha[N], rho[N]);
            double muf = w[f] * muP + (1.0 - w[f]) * muN;
            double lc = muf * magSf[f] * dc[f];
            mLower[f] = cl - lc;
            mUpper[f] = cu - lc;
            diagSum[P] -= mLower[f];  // negSumDiag of the summed matrix
            diagSum[N] -= mUpper[f];
            // explicit: non-orthogonal correction + div(mu dev2(T(gradU)))
            double TP[9], TN[9];
            devTensor(muP, &gradU[9 * P], TP);
            devTensor(muN, &gradU[9 * N], TN);
            for (int j = 0; j < 3; j++) {
                double gcorr = 0, dev = 0;
                for (int i = 0; i < 3; i++) {
                    double gf = w[f] * gradU[9 * P + 3 * i + j] + (1.0 - w[f]) * gradU[9 * N + 3 * i + j];
                    gcorr += corrVec[3 * f + i] * gf;
                    dev += Sf[3 * f + i] * (w[f] * TP[3 * i + j] + (1.0 - w[f]) * TN[3 * i + j]);
                }
                double e = muf * magSf[f] * gcorr + dev;
                src[3 * P + j] += e;
                src[3 * N + j] -= e;
            }
        }
        for (int b = 0; b < nB; b++) {
            int f = nI + b, P = own[f], pt = facePatch[b];
            double n[3] = {Sf[3 * f] / magSf[f], Sf[3 * f + 1] / magSf[f], Sf[3 * f + 2] / magSf[f]};
            double F = rhoPhi[f];
            double mub = muOf(alpha_b[b], rho_b[b]);
            double lc = mub * magSf[f];
            double vIC[3], vBC[3], gIC[3], gBC[3];
            const double *Ub = &U_b[3 * b], *Uc = &U[3 * P];
            if (bcU[pt] == ORC_U_MOVING_WALL) {
                for (int k = 0; k < 3; k++) {
                    vIC[k] = 0.0; vBC[k] = Ub[k];
                    gIC[k] = -dc[f]; gBC[k] = dc[f] * Ub[k];
                }
            } else {
                // directionMixed coefficients via snGradTransformDiag
                for (int k = 0; k < 3; k++) {
                    double vfkk = phi[f] < 0 ? 1.0 - n[k] * n[k] : 0.0;
                    double sTD = std::sqrt(std::fabs(vfkk));
                    vIC[k] = 1.0 - sTD;
                    vBC[k] = Ub[k] - vIC[k] * Uc[k];
                    gIC[k] = -dc[f] * sTD;
                    double sn = (Ub[k] - Uc[k]) * dc[f];
                    gBC[k] = sn - gIC[k] * Uc[k];
                }
            }
            for (int k = 0; k < 3; k++) {
                mBIC[3 * b + k] = F * vIC[k] - lc * gIC[k];
                mBBC[3 * b + k] = -F * vBC[k] + lc * gBC[k];
            }
            // boundary face of div(mu dev2(T(gradU))): gradU_b = gradU_P + n (snGrad - n & gradU_P)
            double gb[9], T[9];
            for (int j = 0; j < 3; j++) {
                double ng = n[0] * gradU[9 * P + j] + n[1] * gradU[9 * P + 3 + j] + n[2] * gradU[9 * P + 6 + j];
                double sn = (Ub[j] - Uc[j]) * dc[f];
                for (int i = 0; i < 3; i++) gb[3 * i + j] = gradU[9 * P + 3 * i + j] + n[i] * (sn - ng);
            }
            devTensor(mub, gb, T);
            for (int j = 0; j < 3; j++) {
                double dev = 0;
                for (int i = 0; i < 3; i++) dev += Sf[3 * f + i] * T[3 * i + j];
                src[3 * P + j] += dev;
            }
        }
        // UEqn.relax() with the reference's relaxation factor 1 (fvSolution:89-95) is not a no-op:
        // fvMatrix::relax enforces diagonal dominance, D = max(|D + sum_b max|iC||, sum|offdiag|)
        // - sum_b min(iC), and moves the difference to the source with the current U [OF13-MEM].
        dvec sumOff(nC, 0.0), bMax(nC, 0.0), bMin(nC, 0.0);
        for (int f = 0; f < nI; f++) {
            sumOff[own[f]] += std::fabs(mUpper[f]);
            sumOff[nei[f]] += std::fabs(mLower[f]);
        }
        for (int b = 0; b < nB; b++) {
            int P = own[nI + b];
            const double* iC = &mBIC[3 * b];
            bMax[P] += std::max(std::fabs(iC[0]), std::max(std::fabs(iC[1]), std::fabs(iC[2])));
            bMin[P] += std::min(iC[0], std::min(iC[1], iC[2]));
        }
        for (int c = 0; c < nC; c++) {
            double D0 = rDeltaT * rho[c] * V[c] + diagSum[c];
            double D = D0 + bMax[c];
            D = std::max(std::fabs(D), sumOff[c]);
            D = D - bMin[c];
            if (xNoRelax) D = D0;
            mDiag[c] = D;
            for (int k = 0; k < 3; k++)
                mSource[3 * c + k] = (rDeltaT * rho0[c] * U0[3 * c + k] * V0[c] + src[3 * c + k]) + (D - D0) * U[3 * c + k];
        }
    }

    // ---- S5: pressure corrector (ref: system/fvSolution:42-66,78-87) ----------------------
    void computeHbyA() {
        // A = (diag + cmptAv(boundary diag))/V ; H = (b - sum_N a_N U_N)/V  [OF13-MEM: fvMatrix::A/H]
        dvec D = mDiag;
        dvec H(3 * nC, 0.0);
        for (int b = 0; b < nB; b++) {
            int P = own[nI + b];
            double av = (mBIC[3 * b] + mBIC[3 * b + 1] + mBIC[3 * b + 2]) / 3.0;
            D[P] += av;
            for (int k = 0; k < 3; k++) H[3 * P + k] += (av - mBIC[3 * b + k]) * U[3 * P + k];
        }
        dvec lduH(3 * nC, 0.0);
        for (int f = 0; f < nI; f++) {
            int P = own[f], N = nei[f];
            for (int k = 0; k < 3; k++) {
                lduH[3 * N + k] -= mLower[f] * U[3 * P + k];
                lduH[3 * P + k] -= mUpper[f] * U[3 * N + k];
            }
        }
        for (int c = 0; c < 3 * nC; c++) H[c] += lduH[c] + mSource[c];
        dvec bbc(3 * nC, 0.0);
        for (int b = 0; b < nB; b++)
            for (int k = 0; k < 3; k++) bbc[3 * own[nI + b] + k] += mBBC[3 * b + k];
        for (int c = 0; c < 3 * nC; c++) H[c] += bbc[c];
        rAU.resize(nC); HbyA.resize(3 * nC); HbyA_b.resize(3 * nB);
        for (int c = 0; c < nC; c++) {
            double A = D[c] / V[c];
            rAU[c] = 1.0 / A;
            for (int k = 0; k < 3; k++) HbyA[3 * c + k] = rAU[c] * (H[3 * c + k] / V[c]);
        }
        // constrainHbyA: fixed-value patches take U_b, assignable ones keep the extrapolation
        for (int b = 0; b < nB; b++) {
            int pt = facePatch[b], P = own[nI + b];
            for (int k = 0; k < 3; k++) HbyA_b[3 * b + k] = bcU[pt] == ORC_U_MOVING_WALL ? U_b[3 * b + k] : HbyA[3 * P + k];
        }
    }

    void pcPrepare() {
        computeHbyA();
        const double rDeltaT = 1.0 / dt;
        rAUf.resize(nF); phiHbyA.resize(nF); phig.resize(nF);
        gradScalar(rho, rho_b, gradRho);
        for (int f = 0; f < nF; f++) {
            int P = own[f];
            const double* S = &Sf[3 * f];
            double flux, ddtCorr, rhorAUf, snGradRho;
            if (f < nI) {
                int N = nei[f];
                double wf = w[f];
                rAUf[f] = wf * rAU[P] + (1.0 - wf) * rAU[N];
                double hf[3], u0f[3], gr[3];
                for (int k = 0; k < 3; k++) {
                    hf[k] = wf * HbyA[3 * P + k] + (1.0 - wf) * HbyA[3 * N + k];
                    u0f[k] = wf * U0[3 * P + k] + (1.0 - wf) * U0[3 * N + k];
                    gr[k] = wf * gradRho[3 * P + k] + (1.0 - wf) * gradRho[3 * N + k];
                }
                flux = dot3(S, hf);
                double phiUf0 = dot3(S, &Uf0[3 * f]);
                double pc = phiUf0 - dot3(S, u0f);
                double coeff = 1.0 - std::min(std::fabs(pc) / (std::fabs(phiUf0) + SMALL), 1.0);
                if (xDdt >= 0) coeff = xDdt;
                ddtCorr = coeff * rDeltaT * pc;
                rhorAUf = wf * (rho[P] * rAU[P]) + (1.0 - wf) * (rho[N] * rAU[N]);
                snGradRho = dc[f] * (rho[N] - rho[P]) + dot3(&corrVec[3 * f], gr);
            } else {
                int b = f - nI, pt = facePatch[b];
                rAUf[f] = rAU[P];
                flux = dot3(S, &HbyA_b[3 * b]);
                double phiUf0 = dot3(S, &Uf0[3 * f]);
                double pc = phiUf0 - dot3(S, &U0_b[3 * b]);
                double coeff = bcU[pt] == ORC_U_MOVING_WALL ? 0.0 : 1.0 - std::min(std::fabs(pc) / (std::fabs(phiUf0) + SMALL), 1.0);
                ddtCorr = coeff * rDeltaT * pc;
                rhorAUf = rho_b[b] * rAU[P];
                snGradRho = dc[f] * (rho_b[b] - rho[P]);
            }
            double ghf = dot3(cfg.g, &Cf[3 * f]);
            double st = cfg.sigma != 0.0 && (int)stf.size() == nF ? stf[f] : 0.0;
            phig[f] = (st - ghf * snGradRho) * rAUf[f] * magSf[f];
            phiHbyA[f] = (flux + rhorAUf * ddtCorr) + phig[f];
        }
        // constrainPressure: fixedFluxPressure gradient
        for (int b = 0; b < nB; b++) {
            int f = nI + b;
            if (bcP[facePatch[b]] == ORC_P_FIXED_FLUX)
                pGrad_b[b] = (phiHbyA[f] - dot3(&Sf[3 * f], &U_b[3 * b])) / (magSf[f] * rAUf[f]);
        }
    }
    void pcAssemble() {
            pTotalPressure();
            gradScalar(p_rgh, p_rgh_b, gradP);
            pUpper.assign(nI, 0.0); pDiag.assign(nC, 0.0); pSource.assign(nC, 0.0); pCorrFlux.assign(nI, 0.0);
            dvec divPhi(nC, 0.0), divCorr(nC, 0.0);
            for (int f = 0; f < nI; f++) {
                int P = own[f], N = nei[f];
                double c = rAUf[f] * magSf[f];
                pUpper[f] = c * dc[f];
                double g[3];
                for (int k = 0; k < 3; k++) g[k] = w[f] * gradP[3 * P + k] + (1.0 - w[f]) * gradP[3 * N + k];
                pCorrFlux[f] = c * dot3(&corrVec[3 * f], g);
                pDiag[P] += pUpper[f];
                pDiag[N] += pUpper[f];
                divPhi[P] += phiHbyA[f];
                divPhi[N] -= phiHbyA[f];
                divCorr[P] += pCorrFlux[f];
                divCorr[N] -= pCorrFlux[f];
            }
            dvec bDiag(nC, 0.0), bSrc(nC, 0.0);
            for (int b = 0; b < nB; b++) {
                int f = nI + b, P = own[f];
                divPhi[P] += phiHbyA[f];
                double c = rAUf[f] * magSf[f];
                if (bcP[facePatch[b]] == ORC_P_TOTAL_PRESSURE) {
                    bDiag[P] += c * dc[f];
                    bSrc[P] += c * dc[f] * p_rgh_b[b];
                } else
                    bSrc[P] += c * pGrad_b[b];
            }
            for (int c = 0; c < nC; c++) {
                pDiag[c] += bDiag[c];
                pSource[c] = (divCorr[c] - divPhi[c]) + bSrc[c];
            }
            if (needRef) {
                pSource[refCell] += pDiag[refCell] * p_rgh[refCell];
                pDiag[refCell] += pDiag[refCell];
            }
    }
    void pcFinish() {
                // phi = phiHbyA - flux(p) ; U = HbyA + rAU*reconstruct((phig - flux)/rAUf)
                dvec T(9 * nC, 0.0), rv(3 * nC, 0.0);
                for (int f = 0; f < nF; f++) {
                    int P = own[f];
                    double fl;
                    if (f < nI) fl = pUpper[f] * (p_rgh[nei[f]] - p_rgh[P]) + pCorrFlux[f];
                    else {
                        int b = f - nI;
                        double c = rAUf[f] * magSf[f];
                        fl = bcP[facePatch[b]] == ORC_P_TOTAL_PRESSURE ? c * dc[f] * (p_rgh_b[b] - p_rgh[P]) : c * pGrad_b[b];
                    }
                    phi[f] = phiHbyA[f] - fl;
                    double ssf = (phig[f] - fl) / rAUf[f];
                    double sh[3] = {Sf[3 * f] / magSf[f], Sf[3 * f + 1] / magSf[f], Sf[3 * f + 2] / magSf[f]};
                    for (int i = 0; i < 3; i++) {
                        for (int j = 0; j < 3; j++) {
                            double v = sh[i] * Sf[3 * f + j];
                            T[9 * P + 3 * i + j] += v;
                            if (f < nI) T[9 * nei[f] + 3 * i + j] += v;
                        }
                        rv[3 * P + i] += sh[i] * ssf;
                        if (f < nI) rv[3 * nei[f] + i] += sh[i] * ssf;
                    }
                }
                for (int c = 0; c < nC; c++) {
                    const double* t = &T[9 * c];
                    double xx = t[0], xy = t[1], xz = t[2], yx = t[3], yy = t[4], yz = t[5], zx = t[6], zy = t[7], zz = t[8];
                    double det = xx * (yy * zz - yz * zy) - xy * (yx * zz - yz * zx) + xz * (yx * zy - yy * zx);
                    double inv[9] = {yy * zz - zy * yz, xz * zy - xy * zz, xy * yz - xz * yy,
                                     zx * yz - yx * zz, xx * zz - xz * zx, yx * xz - xx * yz,
                                     yx * zy - yy * zx, xy * zx - xx * zy, xx * yy - yx * xy};
                    for (int i = 0; i < 3; i++) {
                        double r = (inv[3 * i] / det) * rv[3 * c] + (inv[3 * i + 1] / det) * rv[3 * c + 1] + (inv[3 * i + 2] / det) * rv[3 * c + 2];
                        U[3 * c + i] = HbyA[3 * c + i] + rAU[c] * r;
                    }
                }
                UBCs();
    }
    void pcEnd() {
        // fvc::correctUf ; fvc::makeRelative ; p = p_rgh + rho*gh
        if (cfg.n_motion > 0) {
            for (int f = 0; f < nF; f++) {
                double uf[3];
                if (f < nI)
                    for (int k = 0; k < 3; k++) uf[k] = w[f] * U[3 * own[f] + k] + (1.0 - w[f]) * U[3 * nei[f] + k];
                else
                    for (int k = 0; k < 3; k++) uf[k] = U_b[3 * (f - nI) + k];
                double n[3] = {Sf[3 * f] / magSf[f], Sf[3 * f + 1] / magSf[f], Sf[3 * f + 2] / magSf[f]};
                double a = phi[f] / magSf[f] - dot3(n, uf);
                for (int k = 0; k < 3; k++) Uf[3 * f + k] = uf[k] + n[k] * a;
                phi[f] -= meshPhi[f];
            }
        }
        for (int c = 0; c < nC; c++) p[c] = p_rgh[c] + rho[c] * dot3(cfg.g, &C[3 * c]);
        if (needRef) {
            double shift = cfg.p_ref_value - p[refCell];
            for (int c = 0; c < nC; c++) {
                p[c] += shift;
                p_rgh[c] = p[c] - rho[c] * dot3(cfg.g, &C[3 * c]);
            }
            pEvaluateAfterShift(shift);
        }
    }
    void pressureCorrector(bool finalIter) {
        pcPrepare();
        for (int nonOrth = 0; nonOrth <= cfg.n_non_orth; nonOrth++) {
            bool finalNonOrth = nonOrth == cfg.n_non_orth;
            pcAssemble();
            const orc_solver_t& ctl = (finalIter && finalNonOrth) ? cfg.p_rgh_final : cfg.p_rgh;
            int which = (finalIter && finalNonOrth) ? 1 : 0;
            lastSolve[which] = solveP(ctl, which, pDiag, pUpper, pSource, p_rgh);
            pEvaluate();
            if (finalNonOrth) pcFinish();
        }
        pcEnd();
    }
    void pEvaluateAfterShift(double shift) {
        for (int b = 0; b < nB; b++)
            if (bcP[facePatch[b]] == ORC_P_FIXED_FLUX) p_rgh_b[b] = p_rgh[own[nI + b]] + pGrad_b[b] / dc[nI + b];
        (void)shift;
    }

    // `upper` holds the positive Laplacian coefficients; the matrix off-diagonals are -upper
    SolveStats solveP(const orc_solver_t& ctl, int which, const dvec& diag, const dvec& upper, const dvec& b, dvec& x) {
        A.diag = diag;
        A.upper.resize(nI);
        for (int f = 0; f < nI; f++) A.upper[f] = -upper[f];
        if (ctl.type == 0 && ctl.precond == 0) return pcgDIC(A, x, b, ctl.tolerance, ctl.rel_tol, ctl.max_iter);
        Gamg& G = gamg[which];
        G.ctl = ctl;
        if (!G.haveAgglom) {
            // faceAreaPair weights: |Sf/sqrt(|Sf|) * (1, 1.01, 1.02)|  [OF13-MEM]
            dvec fw(nI);
            for (int f = 0; f < nI; f++) {
                double s = std::sqrt(magSf[f]);
                double v[3] = {Sf[3 * f] / s * 1.0, Sf[3 * f + 1] / s * 1.01, Sf[3 * f + 2] / s * 1.02};
                fw[f] = mag3(v);
            }
            G.buildAgglomeration(A, fw);
        }
        G.fine = &A;
        G.agglomerateMatrix();
        if (ctl.type == 1) return G.solve(x, b);
        return pcgGamg(A, G, x, b, ctl.tolerance, ctl.rel_tol, ctl.max_iter);
    }

    int findCell(const double* x) const {
        // cell whose faces all see the point on the inner side; nearest centre as tie-break
        ivec nbad(nC, 0);
        for (int f = 0; f < nF; f++) {
            double d[3] = {x[0] - Cf[3 * f], x[1] - Cf[3 * f + 1], x[2] - Cf[3 * f + 2]};
            double s = dot3(d, &Sf[3 * f]);
            double tol = 1e-12 * magSf[f] * std::sqrt(magSf[f]);
            if (s > tol) nbad[own[f]]++;
            if (f < nI && s < -tol) nbad[nei[f]]++;
        }
        int best = -1;
        double bd = 1e300;
        for (int c = 0; c < nC; c++)
            if (nbad[c] == 0) {
                double d[3] = {x[0] - C[3 * c], x[1] - C[3 * c + 1], x[2] - C[3 * c + 2]};
                double dd = dot3(d, d);
                if (dd < bd) { bd = dd; best = c; }
            }
        return best;
    }

    bool oneStep() {
        courant();
        adjustDeltaT();
        bool wr = advanceTime();
        moveMesh();
        alphaPredictor();
        momentum();
        for (int corr = 0; corr < cfg.n_correctors; corr++) pressureCorrector(corr == cfg.n_correctors - 1);
        if (!probeCells.empty()) {
            probeLog.push_back(t);
            for (int c : probeCells) probeLog.push_back(c >= 0 ? p[c] : -1.79769e+307);
        }
        return wr;
    }
};

// ------------------------------------------------------------------------------------
// C API
// ------------------------------------------------------------------------------------
extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

orc_state* orc_create(const orc_mesh_t* m, const orc_config_t* c) {
    orc_state* s = new orc_state();
    s->nP = m->n_points; s->nF = m->n_faces; s->nI = m->n_internal; s->nC = m->n_cells;
    s->nB = s->nF - s->nI; s->nPatch = m->n_patches;
    s->points0.assign(m->points, m->points + 3 * s->nP);
    s->fOff.assign(m->face_offsets, m->face_offsets + s->nF + 1);
    s->fLab.assign(m->face_labels, m->face_labels + s->fOff[s->nF]);
    s->own.assign(m->owner, m->owner + s->nF);
    s->nei.assign(m->neighbour, m->neighbour + s->nI);
    s->pStart.assign(m->patch_start, m->patch_start + s->nPatch);
    s->pSize.assign(m->patch_size, m->patch_size + s->nPatch);
    s->bcU.assign(m->patch_bc_u, m->patch_bc_u + s->nPatch);
    s->bcA.assign(m->patch_bc_alpha, m->patch_bc_alpha + s->nPatch);
    s->bcP.assign(m->patch_bc_p, m->patch_bc_p + s->nPatch);
    s->pInletAlpha.assign(m->patch_inlet_alpha, m->patch_inlet_alpha + s->nPatch);
    s->pP0.assign(m->patch_p0, m->patch_p0 + s->nPatch);
    s->facePatch.assign(s->nB, -1);
    for (int p = 0; p < s->nPatch; p++)
        for (int i = 0; i < s->pSize[p]; i++) s->facePatch[s->pStart[p] - s->nI + i] = p;
    for (int b = 0; b < s->nB; b++)
        if (s->facePatch[b] < 0) {
            g_err = "boundary face without a patch";
            delete s;
            return nullptr;
        }
    s->cfg = *c;
    if (c->n_motion > 0) s->motion.assign(c->motion, c->motion + 7 * c->n_motion);
    s->cfg.motion = nullptr;
    s->t = s->startTime = c->start_time;
    s->dt = s->dt0 = c->delta_t;
    s->transformPoints(s->t, s->points);
    s->calcGeometry();
    s->V0 = s->V;
    s->CfOld = s->Cf;
    s->pointsOld = s->points;
    double vs = 0;
    for (double v : s->V) vs += v;
    s->deltaN = 1e-8 / std::cbrt(vs / s->nC);  // interfaceProperties deltaN [OF13-MEM]
    if (const char* e = getenv("ORC_X_CLIP")) s->xClip = atoi(e);
    if (const char* e = getenv("ORC_X_OWN")) s->xOwn = atoi(e);
    if (const char* e = getenv("ORC_X_BND")) s->xBndExt = atoi(e);
    if (const char* e = getenv("ORC_X_DDT")) s->xDdt = atof(e);
    if (const char* e = getenv("ORC_X_NORELAX")) s->xNoRelax = atoi(e);
    int nC = s->nC, nF = s->nF, nB = s->nB;
    s->meshPhi.assign(nF, 0); s->alpha.assign(nC, 0); s->alpha_b.assign(nB, 0); s->U.assign(3 * nC, 0);
    s->U_b.assign(3 * nB, 0); s->p_rgh.assign(nC, 0); s->p_rgh_b.assign(nB, 0); s->p.assign(nC, 0);
    s->rho.assign(nC, c->rho2); s->rho_b.assign(nB, c->rho2); s->phi.assign(nF, 0); s->Uf.assign(3 * nF, 0);
    s->U0 = s->U; s->U0_b = s->U_b; s->rho0 = s->rho; s->Uf0 = s->Uf;
    s->alphaPhi.assign(nF, 0); s->rhoPhi.assign(nF, 0); s->pGrad_b.assign(nB, 0);
    s->needRef = true;
    for (int p = 0; p < s->nPatch; p++)
        if (s->bcP[p] == ORC_P_TOTAL_PRESSURE && s->pSize[p] > 0) s->needRef = false;
    if (s->needRef) {
        s->refCell = s->findCell(c->p_ref_point);
        if (s->refCell < 0) {
            g_err = "pRefPoint is outside the mesh and p_rgh needs a reference";
            delete s;
            return nullptr;
        }
    }
    s->A.n = nC; s->A.nf = s->nI; s->A.l.assign(s->own.begin(), s->own.begin() + s->nI); s->A.u = s->nei;
    s->A.finalize();
    s->registerFields();
    return s;
}

void orc_destroy(orc_state* s) { delete s; }

long orc_size(orc_state* s, const char* name) {
    auto it = s->reg.find(name);
    return it == s->reg.end() ? -1 : (long)it->second->size();
}
long orc_get(orc_state* s, const char* name, double* out, long cap) {
    auto it = s->reg.find(name);
    if (it == s->reg.end()) return -1;
    long n = std::min<long>(cap, it->second->size());
    std::memcpy(out, it->second->data(), n * sizeof(double));
    return (long)it->second->size();
}
long orc_set(orc_state* s, const char* name, const double* in, long n) {
    if (!strcmp(name, "time")) { s->t = in[0]; return 1; }
    if (!strcmp(name, "deltaT")) { s->dt = in[0]; return 1; }
    auto it = s->reg.find(name);
    if (it == s->reg.end()) return -1;
    it->second->assign(in, in + n);
    return n;
}
long orc_get_int(orc_state* s, const char* name, int* out, long cap) {
    const ivec* v = nullptr;
    if (!strcmp(name, "owner")) v = &s->own;
    else if (!strcmp(name, "neighbour")) v = &s->nei;
    else if (!strcmp(name, "facePatch")) v = &s->facePatch;
    else if (!strncmp(name, "gamgRestrict", 12)) {
        int which = name[12] - '0', lvl = atoi(name + 14);
        if (which < 0 || which > 1 || lvl >= (int)s->gamg[which].lv.size()) return -1;
        v = &s->gamg[which].lv[lvl].restrictAddr;
    }
    if (!v) return -1;
    long n = std::min<long>(cap, v->size());
    std::memcpy(out, v->data(), n * sizeof(int));
    return (long)v->size();
}

int orc_stage(orc_state* s, const char* name) {
    std::string n(name);
    if (n == "courant") s->courant();
    else if (n == "adjustDeltaT") s->adjustDeltaT();
    else if (n == "advanceTime") s->advanceTime();
    else if (n == "moveMesh") s->moveMesh();
    else if (n == "alphaBCs") s->alphaBCs();
    else if (n == "UBCs") s->UBCs();
    else if (n == "mixture") s->mixtureCorrect();
    else if (n == "alphaSubCycle") s->alphaSubCycle(s->dt / s->cfg.n_alpha_subcycles);
    else if (n == "alphaPredictor") s->alphaPredictor();
    else if (n == "momentum") s->momentum();
    else if (n == "HbyA") s->computeHbyA();
    else if (n == "pcPrepare") s->pcPrepare();
    else if (n == "pcAssemble") s->pcAssemble();
    else if (n == "pcFinish") { s->pEvaluate(); s->pcFinish(); }
    else if (n == "pcEnd") s->pcEnd();
    else if (n == "pressureCorrector:0") s->pressureCorrector(false);
    else if (n == "pressureCorrector:1") s->pressureCorrector(true);
    else {
        g_err = "unknown stage " + n;
        return -1;
    }
    return 0;
}

int orc_step(orc_state* s, int n) {
    for (int i = 0; i < n; i++) s->oneStep();
    return 0;
}

int orc_run_to_write(orc_state* s, long max_steps) {
    for (long i = 0; i < max_steps; i++) {
        if (!(s->t < s->cfg.end_time - 0.5 * s->dt)) return 0;
        if (s->oneStep()) return 1;
    }
    return 2;
}

void orc_info(orc_state* s, double* o) {
    o[0] = s->t; o[1] = s->dt; o[2] = (double)s->step; o[3] = s->Co; o[4] = s->alphaCo;
    o[5] = s->lastSolve[0].iters; o[6] = s->lastSolve[0].r0; o[7] = s->lastSolve[0].r;
    o[8] = s->lastSolve[1].iters; o[9] = s->lastSolve[1].r0; o[10] = s->lastSolve[1].r;
    o[11] = s->needRef ? s->refCell : -1; o[12] = s->deltaN; o[13] = s->writeTimeIndex;
    o[14] = (double)s->gamg[0].lv.size(); o[15] = (double)s->gamg[1].lv.size();
}

int orc_solve(orc_state* s, const orc_solver_t* ctl, const double* diag, const double* upper, const double* b,
              double* x, double* r0, double* r) {
    dvec d(diag, diag + s->nC), u(upper, upper + s->nI), bb(b, b + s->nC), xx(x, x + s->nC);
    // a private agglomeration slot so test solves do not disturb the step's caches
    orc_state* t = s;
    Gamg saved = t->gamg[1];
    t->gamg[1] = Gamg();
    SolveStats st = t->solveP(*ctl, 1, d, u, bb, xx);
    t->gamg[1] = saved;
    std::memcpy(x, xx.data(), s->nC * sizeof(double));
    *r0 = st.r0;
    *r = st.r;
    return st.iters;
}

void orc_set_probes(orc_state* s, int n, const int* cells) { s->probeCells.assign(cells, cells + n); }
long orc_probe_log(orc_state* s, double* out, long cap_rows) {
    long w = 1 + (long)s->probeCells.size();
    long rows = (long)s->probeLog.size() / w;
    long n = std::min(rows, cap_rows);
    std::memcpy(out, s->probeLog.data(), n * w * sizeof(double));
    s->probeLog.erase(s->probeLog.begin(), s->probeLog.begin() + n * w);
    return n;
}
int orc_find_cell(orc_state* s, const double* xyz) { return s->findCell(xyz); }
}
