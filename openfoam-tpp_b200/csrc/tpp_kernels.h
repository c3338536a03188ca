// Per-step kernels of the incompressibleVoF PIMPLE step.
//
// Each b_<name>(d, i) is the body of the CUDA kernel k_<name>; see tpp_common.h.
// Cell kernels gather over the cell's faces through the ELL table (cf/cn) whose slots are
// sorted by ascending face index, i.e. exactly the order in which OpenFOAM's face loops
// (internal faces, then patches) would scatter into that cell: every sum below reproduces
// the serial face-loop sum bit for bit without atomics (built with -fmad=false).
//
// What is computed follows the reference's dictionaries:
//   alpha:    system/fvSchemes:30 (interfaceCompression vanLeer 1), fvSolution:19-23
//   momentum: system/fvSchemes:19,29,32,37 ; fvSolution:80 (assembled, never solved)
//   pressure: system/fvSchemes:37,47 ; fvSolution:81-86 ; constant/g:18
//   BCs:      0/U:22-31, 0/alpha.water:22-31, 0/p_rgh:22-31
// and OpenFOAM-13's algorithm for each ([OF13-MEM]).
#pragma once
#include "tpp_common.h"

namespace tpp {

// WT = compile-time ELL width (tets 4, prisms 5, hexes 6; 0 = run-time width): the slot loop is
// fully unrolled and has no early exit, so the index loads of all slots are issued together and
// the gathers of the slots overlap - these kernels are latency-bound otherwise.  Padded slots
// (e < 0) only ever follow the real ones, so skipping them keeps the summation order.
#define FOR_CELL_FACES(d, c)                                  \
    _Pragma("unroll")                                         \
    for (int s_ = 0; s_ < (WT > 0 ? WT : (d).W); s_++) {      \
        const int e_ = (d).cf[(size_t)s_ * (d).nCp + (c)];    \
        if (e_ < 0) { if (WT > 0) continue; else break; }     \
        const int f = e_ >> 1;                                \
        const int isN = e_ & 1;                               \
        const int o = (d).cn[(size_t)s_ * (d).nCp + (c)];     \
        (void)o; (void)isN;
#define END_CELL_FACES }
// Load-first form of the hottest cell kernels (WT > 0): the (face, neighbour) indices of all
// slots, then every gathered operand of all slots, are loaded before the first dependent use,
// so one thread keeps 4-6 x (operands per slot) loads in flight; the arithmetic that follows is
// the run-time-width loop's, slot by slot in the same order (bit-identical sums).  A padded slot
// reads face 0 / the cell itself (valid addresses) and is skipped in the arithmetic.
template <int WT> HD void load_slots(const DV& d, int c, int* e, int* o) {
#pragma unroll
    for (int k = 0; k < WT; k++) { e[k] = d.cf[(size_t)k * d.nCp + c]; o[k] = d.cn[(size_t)k * d.nCp + c]; }
}

// ---- S0 Courant numbers ------------------------------------------------------------------
template <int WT> HD void b_courant(const DV& d, int c) {
    double s = 0;
    FOR_CELL_FACES(d, c) s += fabs(d.phi[f]); END_CELL_FACES
    double v = s / d.V[c];
    d.cellTmp[c] = v;
    double near = pos0_(d.alpha[c] - 0.01) * pos0_(0.99 - d.alpha[c]);
    d.cellTmp[d.nC + c] = near * s / d.V[c];
}

// ---- boundary conditions -----------------------------------------------------------------
HD void b_alpha_bc(const DV& d, int b) {
    int f = d.nI + b;
    if (d.bcA[b] == 1) d.alpha_b[b] = d.phi[f] >= 0 ? d.alpha[d.own[f]] : d.bInletAlpha[b];
    else d.alpha_b[b] = d.alpha[d.own[f]];
}

HD void b_U_bc(const DV& d, int b) {
    int f = d.nI + b, c = d.own[f];
    double m = d.magSf[f];
    double n[3] = {d.Sf[3 * f] / m, d.Sf[3 * f + 1] / m, d.Sf[3 * f + 2] / m};
    if (d.bcU[b] == 0) {
        if (d.moving) {
            double Up[3];
            if (d.rotating) {
                double q[3] = {d.Cf0[3 * f] - d.cofg[0], d.Cf0[3 * f + 1] - d.cofg[1], d.Cf0[3 * f + 2] - d.cofg[2]};
                for (int k = 0; k < 3; k++) {
                    double cn_ = (d.R[3 * k] * q[0] + d.R[3 * k + 1] * q[1] + d.R[3 * k + 2] * q[2]) + d.cofg[k] + d.Tn[k];
                    double co_ = (d.Rold[3 * k] * q[0] + d.Rold[3 * k + 1] * q[1] + d.Rold[3 * k + 2] * q[2]) + d.cofg[k] + d.To[k];
                    Up[k] = (cn_ - co_) / d.dt;
                }
            } else
                for (int k = 0; k < 3; k++) Up[k] = s_wallU(d, k);
            double Un = d.meshPhi[f] / (m + VSMALL);
            double nUp = dot3(n, Up);
            for (int k = 0; k < 3; k++) d.U_b[3 * b + k] = Up[k] + n[k] * (Un - nUp);
        }
    } else {
        const double* Uc = &d.U[3 * c];
        if (d.phi[f] < 0) {
            double nu = dot3(n, Uc);
            for (int k = 0; k < 3; k++) d.U_b[3 * b + k] = n[k] * nu;
        } else
            for (int k = 0; k < 3; k++) d.U_b[3 * b + k] = Uc[k];
    }
}

HD void b_p_total(const DV& d, int b) {
    if (d.bcP[b] == 1) {
        int f = d.nI + b;
        const double* u = &d.U_b[3 * b];
        d.p_rgh_b[b] = d.bP0[b] - 0.5 * d.rho_b[b] * (1.0 - pos0_(d.phi[f])) * dot3(u, u);
    }
}

HD void b_p_evaluate(const DV& d, int b) {
    if (d.bcP[b] == 0) {
        int f = d.nI + b;
        d.p_rgh_b[b] = d.p_rgh[d.own[f]] + d.pGrad_b[b] / d.dc[f];
    }
}

// ---- Gauss linear gradient of a scalar (gs, gsb -> gout) -----------------------------------
template <int WT> HD void b_grad_scalar(const DV& d, int c) {
    double g[3] = {0, 0, 0};
    const double Vc = d.V[c];  // independent of the gathers: issued with the first batch of loads
    if constexpr (WT > 0) {
        int e[WT], o[WT];
        load_slots<WT>(d, c, e, o);
        const double own = d.gs[c];
        double wv[WT], S[WT][3], ov[WT];
#pragma unroll
        for (int k = 0; k < WT; k++) {
            const bool live = e[k] >= 0;
            const int f = live ? e[k] >> 1 : 0;
            wv[k] = d.w[f];
            S[k][0] = d.Sf[3 * f]; S[k][1] = d.Sf[3 * f + 1]; S[k][2] = d.Sf[3 * f + 2];
            const double* q = f >= d.nI ? &d.gsb[f - d.nI] : &d.gs[live ? o[k] : c];
            ov[k] = *q;
        }
#pragma unroll
        for (int k = 0; k < WT; k++) {
            if (e[k] < 0) continue;
            const int f = e[k] >> 1, isN = e[k] & 1;
            if (f < d.nI) {
                double gP = isN ? ov[k] : own, gN = isN ? own : ov[k];
                double sf = wv[k] * gP + (1.0 - wv[k]) * gN;
                for (int j = 0; j < 3; j++) {
                    double v = S[k][j] * sf;
                    if (isN) g[j] -= v; else g[j] += v;
                }
            } else
                for (int j = 0; j < 3; j++) g[j] += S[k][j] * ov[k];
        }
    } else {
        FOR_CELL_FACES(d, c)
            if (f < d.nI) {
                int P = isN ? o : c, N = isN ? c : o;
                double sf = d.w[f] * d.gs[P] + (1.0 - d.w[f]) * d.gs[N];
                for (int k = 0; k < 3; k++) {
                    double v = d.Sf[3 * f + k] * sf;
                    if (isN) g[k] -= v; else g[k] += v;
                }
            } else
                for (int k = 0; k < 3; k++) g[k] += d.Sf[3 * f + k] * d.gsb[f - d.nI];
        END_CELL_FACES
    }
    for (int k = 0; k < 3; k++) d.gout[3 * c + k] = g[k] / Vc;
}

// ---- surface tension, sigma != 0 only (an extension: constant/phaseProperties:19 is sigma 0 in the
// reference).  interfaceProperties::calculateK + surfaceTensionForce [OF13-MEM]:
//   nHatf = (interpolate(grad alpha) / (|.| + deltaN)) & Sf ; K = -div(nHatf) ;
//   stf = interpolate(sigma K) * snGrad(alpha)   (corrected snGrad, system/fvSchemes:47)
// Boundary faces take the cell gradient with its normal part replaced by the patch's snGrad
// (zeroGradient walls, 0/alpha.water:22-25: no contact-angle correction) and the cell's curvature.
HD void face_grad_alpha(const DV& d, int f, double* gf) {
    const int P = d.own[f];
    if (f < d.nI) {
        const int N = d.nei[f];
        const double wl = d.w[f];
        for (int k = 0; k < 3; k++) gf[k] = wl * d.gradA[3 * P + k] + (1.0 - wl) * d.gradA[3 * N + k];
    } else {
        const double m = d.magSf[f];
        const double n[3] = {d.Sf[3 * f] / m, d.Sf[3 * f + 1] / m, d.Sf[3 * f + 2] / m};
        const double corr = d.dc[f] * (d.alpha_b[f - d.nI] - d.alpha[P]) - dot3(n, &d.gradA[3 * P]);
        for (int k = 0; k < 3; k++) gf[k] = d.gradA[3 * P + k] + n[k] * corr;
    }
}
HD void b_nhat_face(const DV& d, int f) {
    double gf[3];
    face_grad_alpha(d, f, gf);
    const double mg = mag3(gf) + d.deltaN;
    d.nHatf[f] = (gf[0] / mg) * d.Sf[3 * f] + (gf[1] / mg) * d.Sf[3 * f + 1] + (gf[2] / mg) * d.Sf[3 * f + 2];
}
template <int WT> HD void b_curvature(const DV& d, int c) {
    double s = 0;
    FOR_CELL_FACES(d, c)
        if (isN) s -= d.nHatf[f]; else s += d.nHatf[f];
    END_CELL_FACES
    d.sigmaK[c] = d.sigma * (0.0 - s / d.V[c]);
}
HD void b_stf_face(const DV& d, int f) {
    const int P = d.own[f];
    if (f < d.nI) {
        const int N = d.nei[f];
        const double wl = d.w[f];
        double gf[3];
        face_grad_alpha(d, f, gf);
        const double sn = d.dc[f] * (d.alpha[N] - d.alpha[P]) + dot3(&d.corrVec[3 * f], gf);
        d.stf[f] = (wl * d.sigmaK[P] + (1.0 - wl) * d.sigmaK[N]) * sn;
    } else
        d.stf[f] = d.sigmaK[P] * (d.dc[f] * (d.alpha_b[f - d.nI] - d.alpha[P]));
}

// ---- S3 alpha: interfaceCompression(vanLeer) flux, upwind flux, MULES -------------------------
HD double vanLeer_limiter(double flux, double pP, double pN, const double* gP, const double* gN, const double* dd) {
    double gradf = pN - pP;
    double gradcf = flux > 0 ? dot3(dd, gP) : dot3(dd, gN);
    double r;
    if (fabs(gradcf) >= 1000 * fabs(gradf)) r = 2 * 1000 * sign_(gradcf) * sign_(gradf) - 1;
    else r = 2 * (gradcf / gradf) - 1;
    return (r + fabs(r)) / (1 + fabs(r));
}

HD void b_alpha_flux(const DV& d, int f) {
    double ph = d.phi[f];
    if (f < d.nI) {
        int P = d.own[f], N = d.nei[f];
        double aP = d.alpha[P], aN = d.alpha[N];
        const double *gP = &d.grad[3 * P], *gN = &d.grad[3 * N];
        double lim = vanLeer_limiter(ph, aP, aN, gP, gN, &d.dPN[3 * f]);
        double wl = d.w[f];
        double wf = lim * wl + (1.0 - lim) * pos0_(ph);
        double vf = wf * aP + (1.0 - wf) * aN;
        double gf[3];
        for (int k = 0; k < 3; k++) gf[k] = wl * gP[k] + (1.0 - wl) * gN[k];
        double mg = mag3(gf) + d.deltaN;
        double nHatf = (gf[0] / mg) * d.Sf[3 * f] + (gf[1] / mg) * d.Sf[3 * f + 1] + (gf[2] / mg) * d.Sf[3 * f + 2];
        vf += d.cAlpha * sign_(ph) * vf * (1.0 - vf) * nHatf / d.magSf[f];
        double un = ph * vf;
        double bd = ph * (ph >= 0 ? aP : aN);
        d.phiBD[f] = bd;
        d.phiCorr[f] = un - bd;
        d.lambda[f] = 1.0;
    } else {
        d.phiBD[f] = ph * d.alpha_b[f - d.nI];
    }
}

template <int WT> HD void b_mules_setup(const DV& d, int c) {
    double mx = 0.0, mn = 1.0, sBD = 0, sP = 0, mM = 0;  // psiMin = 0, psiMax = 1
    const double V = d.V[c], a0 = d.alpha0[c];
    if constexpr (WT > 0) {
        int e[WT], o[WT];
        load_slots<WT>(d, c, e, o);
        double av[WT], bdv[WT], pcv[WT];
#pragma unroll
        for (int k = 0; k < WT; k++) {
            const bool live = e[k] >= 0;
            const int f = live ? e[k] >> 1 : 0;
            const bool in = f < d.nI;
            av[k] = d.alpha[(live && in) ? o[k] : c];
            bdv[k] = d.phiBD[f];
            pcv[k] = d.phiCorr[in ? f : 0];
        }
#pragma unroll
        for (int k = 0; k < WT; k++) {
            if (e[k] < 0) continue;
            const int f = e[k] >> 1, isN = e[k] & 1;
            if (f < d.nI) {
                double a = av[k];
                mx = dmax(mx, a);
                mn = dmin(mn, a);
                double bd = bdv[k], pc = pcv[k];
                if (isN) {
                    sBD -= bd;
                    if (pc > 0) mM += pc; else sP -= pc;
                } else {
                    sBD += bd;
                    if (pc > 0) sP += pc; else mM -= pc;
                }
            } else {
                sBD += bdv[k];
                mM -= 0.0;
            }
        }
    } else {
        FOR_CELL_FACES(d, c)
            if (f < d.nI) {
                double a = d.alpha[o];
                mx = dmax(mx, a);
                mn = dmin(mn, a);
                double bd = d.phiBD[f], pc = d.phiCorr[f];
                if (isN) {
                    sBD -= bd;
                    if (pc > 0) mM += pc; else sP -= pc;
                } else {
                    sBD += bd;
                    if (pc > 0) sP += pc; else mM -= pc;
                }
            } else {
                sBD += d.phiBD[f];
                mM -= 0.0;  // boundary phiCorr is identically 0 (non-coupled patches)
            }
        END_CELL_FACES
    }
    mx = dmin(mx, 1.0);
    mn = dmax(mn, 0.0);
    const double rdt = s_rdt(d);
    d.psiMaxn[c] = V * (rdt * mx) - (V * rdt) * a0 + sBD;
    d.psiMinn[c] = V * (0.0 - rdt * mn) + (V * rdt) * a0 - sBD;
    d.sumPhip[c] = sP;
    d.mSumPhim[c] = mM;
}

template <int WT> HD void b_mules_cell(const DV& d, int c) {
    double sl = 0, ml = 0;
    const double pMax = d.psiMaxn[c], pMin = d.psiMinn[c], mSm = d.mSumPhim[c], sPp = d.sumPhip[c];
    if constexpr (WT > 0) {
        int e[WT], o[WT];
        load_slots<WT>(d, c, e, o);
        double lv[WT], pcv[WT];
#pragma unroll
        for (int k = 0; k < WT; k++) {
            const int f = e[k] >= 0 ? e[k] >> 1 : 0;
            const int fi = f < d.nI ? f : 0;
            lv[k] = d.lambda[fi];
            pcv[k] = d.phiCorr[fi];
        }
#pragma unroll
        for (int k = 0; k < WT; k++) {
            if (e[k] < 0) continue;
            const int f = e[k] >> 1, isN = e[k] & 1;
            if (f < d.nI) {
                double lp = lv[k] * pcv[k];
                if (isN) {
                    if (lp > 0) ml += lp; else sl -= lp;
                } else {
                    if (lp > 0) sl += lp; else ml -= lp;
                }
            } else
                ml -= 0.0;
        }
    } else {
        FOR_CELL_FACES(d, c)
            if (f < d.nI) {
                double lp = d.lambda[f] * d.phiCorr[f];
                if (isN) {
                    if (lp > 0) ml += lp; else sl -= lp;
                } else {
                    if (lp > 0) sl += lp; else ml -= lp;
                }
            } else
                ml -= 0.0;
        END_CELL_FACES
    }
    d.lambdam[c] = dmax(dmin((sl + pMax) / (mSm + ROOTVSMALL), 1.0), 0.0);
    d.lambdap[c] = dmax(dmin((ml + pMin) / (sPp + ROOTVSMALL), 1.0), 0.0);
}

HD void b_mules_face(const DV& d, int f) {
    int P = d.own[f], N = d.nei[f];
    double l = d.lambda[f];
    if (d.phiCorr[f] > 0) l = dmin(l, dmin(d.lambdap[P], d.lambdam[N]));
    else l = dmin(l, dmin(d.lambdam[P], d.lambdap[N]));
    d.lambda[f] = l;
}

HD void b_mules_phipsi(const DV& d, int f) {
    d.alphaPhiUn[f] = f < d.nI ? d.phiBD[f] + d.lambda[f] * d.phiCorr[f] : d.phiBD[f] + 1.0 * 0.0;
}

HD void b_alphaphi_acc(const DV& d, int f) { d.alphaPhi[f] += d.subW * d.alphaPhiUn[f]; }
// the last limiter iteration's face pass, the limited flux and its share of the step's alpha flux
// in one pass over the faces (mules_face + mules_phipsi + alphaphi_acc: same operations, same order)
HD void b_mules_face_final(const DV& d, int f) {
    double v;
    if (f < d.nI) {
        int P = d.own[f], N = d.nei[f];
        double l = d.lambda[f];
        const double pc = d.phiCorr[f];
        if (pc > 0) l = dmin(l, dmin(d.lambdap[P], d.lambdam[N]));
        else l = dmin(l, dmin(d.lambdam[P], d.lambdap[N]));
        d.lambda[f] = l;
        v = d.phiBD[f] + l * pc;
    } else
        v = d.phiBD[f] + 1.0 * 0.0;
    d.alphaPhiUn[f] = v;
    d.alphaPhi[f] += d.subW * v;
}

template <int WT> HD void b_mules_update(const DV& d, int c) {
    double div = 0;
    const double V = d.V[c], a0 = d.alpha0[c];
    if constexpr (WT > 0) {
        int e[WT], o[WT];
        load_slots<WT>(d, c, e, o);
        double pv[WT];
#pragma unroll
        for (int k = 0; k < WT; k++) pv[k] = d.alphaPhiUn[e[k] >= 0 ? e[k] >> 1 : 0];
#pragma unroll
        for (int k = 0; k < WT; k++) {
            if (e[k] < 0) continue;
            if (e[k] & 1) div -= pv[k]; else div += pv[k];
        }
    } else {
        FOR_CELL_FACES(d, c)
            if (isN) div -= d.alphaPhiUn[f]; else div += d.alphaPhiUn[f];
        END_CELL_FACES
    }
    double psiIf = div / V;
    const double rdt = s_rdt(d);
    d.alpha[c] = (V * a0 * rdt / V - psiIf) / rdt;
}

HD void b_mixture_cell(const DV& d, int c) { d.rho[c] = d.alpha[c] * d.rho1 + (1.0 - d.alpha[c]) * d.rho2; }
HD void b_mixture_bnd(const DV& d, int b) { d.rho_b[b] = d.alpha_b[b] * d.rho1 + (1.0 - d.alpha_b[b]) * d.rho2; }
HD void b_rhophi(const DV& d, int f) { d.rhoPhi[f] = d.alphaPhi[f] * (d.rho1 - d.rho2) + d.phi[f] * d.rho2; }

HD double mu_of(const DV& d, double a, double r) {
    double la = dmin(dmax(a, 0.0), 1.0);
    double mu = la * d.rho1 * d.nu1 + (1.0 - la) * d.rho2 * d.nu2;
    double nu = mu / (la * d.rho1 + (1.0 - la) * d.rho2);
    return r * nu;
}

// ---- S4 momentum matrix ---------------------------------------------------------------------
template <int WT> HD void b_grad_U(const DV& d, int c) {
    double g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if constexpr (WT > 0) {
        // load first (indices, then every gathered operand of all slots), accumulate in slot order
        int e[WT], o[WT];
        load_slots<WT>(d, c, e, o);
        const double uc[3] = {d.U[3 * c], d.U[3 * c + 1], d.U[3 * c + 2]};
        double wv[WT], S[WT][3], uo[WT][3];
#pragma unroll
        for (int k = 0; k < WT; k++) {
            const bool live = e[k] >= 0;
            const int f = live ? e[k] >> 1 : 0;
            wv[k] = d.w[f];
            S[k][0] = d.Sf[3 * f]; S[k][1] = d.Sf[3 * f + 1]; S[k][2] = d.Sf[3 * f + 2];
            const double* q = f >= d.nI ? &d.U_b[3 * (f - d.nI)] : &d.U[3 * (live ? o[k] : c)];
            uo[k][0] = q[0]; uo[k][1] = q[1]; uo[k][2] = q[2];
        }
#pragma unroll
        for (int k = 0; k < WT; k++) {
            if (e[k] < 0) continue;
            const int f = e[k] >> 1;
            const bool isN = e[k] & 1;
            if (f < d.nI) {
                const double wl = wv[k];
                double uf[3];
                // uf = w U_P + (1 - w) U_N with P the face's owner
                for (int j = 0; j < 3; j++) uf[j] = isN ? wl * uo[k][j] + (1.0 - wl) * uc[j] : wl * uc[j] + (1.0 - wl) * uo[k][j];
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) {
                        double v = S[k][i] * uf[j];
                        if (isN) g[3 * i + j] -= v; else g[3 * i + j] += v;
                    }
            } else {
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) g[3 * i + j] += S[k][i] * uo[k][j];
            }
        }
    } else {
        FOR_CELL_FACES(d, c)
            if (f < d.nI) {
                int P = isN ? o : c, N = isN ? c : o;
                double wl = d.w[f];
                double uf[3];
                for (int j = 0; j < 3; j++) uf[j] = wl * d.U[3 * P + j] + (1.0 - wl) * d.U[3 * N + j];
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) {
                        double v = d.Sf[3 * f + i] * uf[j];
                        if (isN) g[3 * i + j] -= v; else g[3 * i + j] += v;
                    }
            } else {
                const double* ub = &d.U_b[3 * (f - d.nI)];
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) g[3 * i + j] += d.Sf[3 * f + i] * ub[j];
            }
        END_CELL_FACES
    }
    const double V = d.V[c];
    for (int k = 0; k < 9; k++) d.gradU[9 * c + k] = g[k] / V;
}

HD double vanLeerV_limiter(double flux, const double* uP, const double* uN, const double* gP, const double* gN, const double* dd) {
    double gv[3] = {uN[0] - uP[0], uN[1] - uP[1], uN[2] - uP[2]};
    double gradf = dot3(gv, gv);
    const double* g = flux > 0 ? gP : gN;
    double dg[3];
    for (int j = 0; j < 3; j++) dg[j] = dd[0] * g[j] + dd[1] * g[3 + j] + dd[2] * g[6 + j];
    double gradcf = dot3(gv, dg);
    double r;
    if (fabs(gradcf) >= 1000 * fabs(gradf)) r = 2 * 1000 * sign_(gradcf) * sign_(gradf) - 1;
    else r = 2 * (gradcf / gradf) - 1;
    return (r + fabs(r)) / (1 + fabs(r));
}

HD void dev_tensor(double mu, const double* g, double* T) {
    double tr = g[0] + g[4] + g[8];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T[3 * i + j] = mu * (g[3 * j + i] - (i == j ? (2.0 / 3.0) * tr : 0.0));
}

HD void b_mom_face(const DV& d, int f) {
    int P = d.own[f], N = d.nei[f];
    double F = d.rhoPhi[f];
    const double *gP = &d.gradU[9 * P], *gN = &d.gradU[9 * N];
    double lim = vanLeerV_limiter(F, &d.U[3 * P], &d.U[3 * N], gP, gN, &d.dPN[3 * f]);
    double wl = d.w[f];
    double wf = lim * wl + (1.0 - lim) * pos0_(F);
    double cl = -wf * F, cu = cl + F;
    double muP = mu_of(d, d.alpha[P], d.rho[P]), muN = mu_of(d, d.alpha[N], d.rho[N]);
    double muf = wl * muP + (1.0 - wl) * muN;
    double lc = muf * d.magSf[f] * d.dc[f];
    d.mLower[f] = cl - lc;
    d.mUpper[f] = cu - lc;
    double TP[9], TN[9];
    dev_tensor(muP, gP, TP);
    dev_tensor(muN, gN, TN);
    for (int j = 0; j < 3; j++) {
        double gcorr = 0, dev = 0;
        for (int i = 0; i < 3; i++) {
            double gf = wl * gP[3 * i + j] + (1.0 - wl) * gN[3 * i + j];
            gcorr += d.corrVec[3 * f + i] * gf;
            dev += d.Sf[3 * f + i] * (wl * TP[3 * i + j] + (1.0 - wl) * TN[3 * i + j]);
        }
        d.mExpl[3 * f + j] = muf * d.magSf[f] * gcorr + dev;
    }
}

HD void b_mom_bnd(const DV& d, int b) {
    int f = d.nI + b, P = d.own[f];
    double m = d.magSf[f];
    double n[3] = {d.Sf[3 * f] / m, d.Sf[3 * f + 1] / m, d.Sf[3 * f + 2] / m};
    double F = d.rhoPhi[f];
    double mub = mu_of(d, d.alpha_b[b], d.rho_b[b]);
    double lc = mub * m;
    double dcf = d.dc[f];
    const double *Ub = &d.U_b[3 * b], *Uc = &d.U[3 * P];
    double vIC[3], vBC[3], gIC[3], gBC[3];
    if (d.bcU[b] == 0) {
        for (int k = 0; k < 3; k++) {
            vIC[k] = 0.0; vBC[k] = Ub[k];
            gIC[k] = -dcf; gBC[k] = dcf * Ub[k];
        }
    } else {
        for (int k = 0; k < 3; k++) {
            double vfkk = d.phi[f] < 0 ? 1.0 - n[k] * n[k] : 0.0;
            double sTD = sqrt(fabs(vfkk));
            vIC[k] = 1.0 - sTD;
            vBC[k] = Ub[k] - vIC[k] * Uc[k];
            gIC[k] = -dcf * sTD;
            double sn = (Ub[k] - Uc[k]) * dcf;
            gBC[k] = sn - gIC[k] * Uc[k];
        }
    }
    for (int k = 0; k < 3; k++) {
        d.mBIC[3 * b + k] = F * vIC[k] - lc * gIC[k];
        d.mBBC[3 * b + k] = -F * vBC[k] + lc * gBC[k];
    }
    const double* gc = &d.gradU[9 * P];
    double gb[9], T[9];
    for (int j = 0; j < 3; j++) {
        double ng = n[0] * gc[j] + n[1] * gc[3 + j] + n[2] * gc[6 + j];
        double sn = (Ub[j] - Uc[j]) * dcf;
        for (int i = 0; i < 3; i++) gb[3 * i + j] = gc[3 * i + j] + n[i] * (sn - ng);
    }
    dev_tensor(mub, gb, T);
    for (int j = 0; j < 3; j++) {
        double dev = 0;
        for (int i = 0; i < 3; i++) dev += d.Sf[3 * f + i] * T[3 * i + j];
        d.mExpl[3 * f + j] = dev;
    }
}

// row assembly + UEqn.relax(1): fvMatrix::relax enforces diagonal dominance even with the
// reference's relaxation factor 1 (fvSolution:89-95): D = max(|D + sum_b max|iC||, sum|offdiag|)
// - sum_b min(iC), the difference goes to the source with the current U [OF13-MEM].  This is
// what keeps A = D/V positive when a water-laden mass flux crosses an air cell.
template <int WT> HD void b_mom_cell(const DV& d, int c) {
    double ds = 0, so = 0, bmax = 0, bmin = 0, src[3] = {0, 0, 0};
    FOR_CELL_FACES(d, c)
        if (f < d.nI) {
            if (isN) {
                ds -= d.mUpper[f];
                so += fabs(d.mLower[f]);
                for (int k = 0; k < 3; k++) src[k] -= d.mExpl[3 * f + k];
            } else {
                ds -= d.mLower[f];
                so += fabs(d.mUpper[f]);
                for (int k = 0; k < 3; k++) src[k] += d.mExpl[3 * f + k];
            }
        } else {
            const double* iC = &d.mBIC[3 * (f - d.nI)];
            bmax += dmax(fabs(iC[0]), dmax(fabs(iC[1]), fabs(iC[2])));
            bmin += dmin(iC[0], dmin(iC[1], iC[2]));
            for (int k = 0; k < 3; k++) src[k] += d.mExpl[3 * f + k];
        }
    END_CELL_FACES
    double V = d.V[c];
    const double rdt = s_rdt(d);
    double D0 = rdt * d.rho[c] * V + ds;
    double D = D0 + bmax;
    D = dmax(fabs(D), so);
    D = D - bmin;
    d.mDiag[c] = D;
    for (int k = 0; k < 3; k++) d.mSource[3 * c + k] = (rdt * d.rho0[c] * d.U0[3 * c + k] * V + src[k]) + (D - D0) * d.U[3 * c + k];
}

// ---- S5 pressure corrector ---------------------------------------------------------------------
template <int WT> HD void b_HbyA(const DV& d, int c) {
    double D = d.mDiag[c];
    double hb[3] = {0, 0, 0}, ldu[3] = {0, 0, 0}, bbc[3] = {0, 0, 0};
    const double* Uc = &d.U[3 * c];
    if constexpr (WT > 0) {
        // load first: the off-diagonal coefficient and the neighbour's velocity of every internal slot
        int e[WT], o[WT];
        load_slots<WT>(d, c, e, o);
        double a[WT], un[WT][3];
#pragma unroll
        for (int k = 0; k < WT; k++) {
            const int f = e[k] >= 0 ? e[k] >> 1 : 0;
            const bool internal = e[k] >= 0 && f < d.nI;
            const int fi = internal ? f : 0, oc = internal ? o[k] : c;
            a[k] = (e[k] & 1) ? d.mLower[fi] : d.mUpper[fi];
            un[k][0] = d.U[3 * oc]; un[k][1] = d.U[3 * oc + 1]; un[k][2] = d.U[3 * oc + 2];
        }
#pragma unroll
        for (int k = 0; k < WT; k++) {
            if (e[k] < 0) continue;
            const int f = e[k] >> 1;
            if (f < d.nI) {
                for (int q = 0; q < 3; q++) ldu[q] -= a[k] * un[k][q];
            } else {
                int b = f - d.nI;
                double av = (d.mBIC[3 * b] + d.mBIC[3 * b + 1] + d.mBIC[3 * b + 2]) / 3.0;
                D += av;
                for (int q = 0; q < 3; q++) {
                    hb[q] += (av - d.mBIC[3 * b + q]) * Uc[q];
                    bbc[q] += d.mBBC[3 * b + q];
                }
            }
        }
    } else {
        FOR_CELL_FACES(d, c)
            if (f < d.nI) {
                double a = isN ? d.mLower[f] : d.mUpper[f];
                for (int k = 0; k < 3; k++) ldu[k] -= a * d.U[3 * o + k];
            } else {
                int b = f - d.nI;
                double av = (d.mBIC[3 * b] + d.mBIC[3 * b + 1] + d.mBIC[3 * b + 2]) / 3.0;
                D += av;
                for (int k = 0; k < 3; k++) {
                    hb[k] += (av - d.mBIC[3 * b + k]) * Uc[k];
                    bbc[k] += d.mBBC[3 * b + k];
                }
            }
        END_CELL_FACES
    }
    double V = d.V[c];
    double A = D / V;
    double r = 1.0 / A;
    d.rAU[c] = r;
    for (int k = 0; k < 3; k++) {
        double H = hb[k] + (ldu[k] + d.mSource[3 * c + k]);
        H += bbc[k];
        d.HbyA[3 * c + k] = r * (H / V);
    }
}

HD void b_HbyA_bnd(const DV& d, int b) {
    int P = d.own[d.nI + b];
    for (int k = 0; k < 3; k++) d.HbyA_b[3 * b + k] = d.bcU[b] == 0 ? d.U_b[3 * b + k] : d.HbyA[3 * P + k];
}

HD void b_phiHbyA(const DV& d, int f) {
    int P = d.own[f];
    const double* S = &d.Sf[3 * f];
    double flux, ddtCorr, rhorAUf, snGradRho, raf;
    if (f < d.nI) {
        int N = d.nei[f];
        double wf = d.w[f];
        raf = wf * d.rAU[P] + (1.0 - wf) * d.rAU[N];
        double hf[3], u0f[3], gr[3];
        for (int k = 0; k < 3; k++) {
            hf[k] = wf * d.HbyA[3 * P + k] + (1.0 - wf) * d.HbyA[3 * N + k];
            u0f[k] = wf * d.U0[3 * P + k] + (1.0 - wf) * d.U0[3 * N + k];
            gr[k] = wf * d.grad[3 * P + k] + (1.0 - wf) * d.grad[3 * N + k];
        }
        flux = dot3(S, hf);
        double phiUf0 = dot3(S, &d.Uf0[3 * f]);
        double pc = phiUf0 - dot3(S, u0f);
        double coeff = 1.0 - dmin(fabs(pc) / (fabs(phiUf0) + SMALL), 1.0);
        ddtCorr = coeff * s_rdt(d) * pc;
        rhorAUf = wf * (d.rho[P] * d.rAU[P]) + (1.0 - wf) * (d.rho[N] * d.rAU[N]);
        snGradRho = d.dc[f] * (d.rho[N] - d.rho[P]) + dot3(&d.corrVec[3 * f], gr);
    } else {
        int b = f - d.nI;
        raf = d.rAU[P];
        flux = dot3(S, &d.HbyA_b[3 * b]);
        double phiUf0 = dot3(S, &d.Uf0[3 * f]);
        double pc = phiUf0 - dot3(S, &d.U0_b[3 * b]);
        double coeff = d.bcU[b] == 0 ? 0.0 : 1.0 - dmin(fabs(pc) / (fabs(phiUf0) + SMALL), 1.0);
        ddtCorr = coeff * s_rdt(d) * pc;
        rhorAUf = d.rho_b[b] * d.rAU[P];
        snGradRho = d.dc[f] * (d.rho_b[b] - d.rho[P]);
    }
    d.rAUf[f] = raf;
    const double st = d.stf ? d.stf[f] : 0.0;
    double pg = (st - d.ghf[f] * snGradRho) * raf * d.magSf[f];
    d.phig[f] = pg;
    double ph = (flux + rhorAUf * ddtCorr) + pg;
    d.phiHbyA[f] = ph;
    if (f >= d.nI) {
        int b = f - d.nI;
        if (d.bcP[b] == 0) d.pGrad_b[b] = (ph - dot3(S, &d.U_b[3 * b])) / (d.magSf[f] * raf);
    }
}

HD void b_p_face(const DV& d, int f) {
    int P = d.own[f], N = d.nei[f];
    double c = d.rAUf[f] * d.magSf[f];
    d.pUpper[f] = c * d.dc[f];
    double wl = d.w[f];
    double g[3];
    for (int k = 0; k < 3; k++) g[k] = wl * d.grad[3 * P + k] + (1.0 - wl) * d.grad[3 * N + k];
    d.pCorrFlux[f] = c * dot3(&d.corrVec[3 * f], g);
}

template <int WT> HD void b_p_cell(const DV& d, int c) {
    double dg = 0, divPhi = 0, divCorr = 0, bDiag = 0, bSrc = 0;
    FOR_CELL_FACES(d, c)
        if (f < d.nI) {
            dg += d.pUpper[f];
            if (isN) {
                divPhi -= d.phiHbyA[f];
                divCorr -= d.pCorrFlux[f];
            } else {
                divPhi += d.phiHbyA[f];
                divCorr += d.pCorrFlux[f];
            }
        } else {
            int b = f - d.nI;
            divPhi += d.phiHbyA[f];
            double cc = d.rAUf[f] * d.magSf[f];
            if (d.bcP[b] == 1) {
                bDiag += cc * d.dc[f];
                bSrc += cc * d.dc[f] * d.p_rgh_b[b];
            } else
                bSrc += cc * d.pGrad_b[b];
        }
    END_CELL_FACES
    dg += bDiag;
    double s = (divCorr - divPhi) + bSrc;
    if (d.needRef && c == d.refCell) {
        s += dg * d.p_rgh[c];
        dg += dg;
    }
    d.pDiag[c] = dg;
    d.pSource[c] = s;
}

HD void b_flux(const DV& d, int f) {
    int P = d.own[f];
    double fl;
    if (f < d.nI) fl = d.pUpper[f] * (d.p_rgh[d.nei[f]] - d.p_rgh[P]) + d.pCorrFlux[f];
    else {
        int b = f - d.nI;
        double c = d.rAUf[f] * d.magSf[f];
        fl = d.bcP[b] == 1 ? c * d.dc[f] * (d.p_rgh_b[b] - d.p_rgh[P]) : c * d.pGrad_b[b];
    }
    d.phi[f] = d.phiHbyA[f] - fl;
    d.rec[f] = (d.phig[f] - fl) / d.rAUf[f];
}

template <int WT> HD void b_U_recon(const DV& d, int c) {
    double T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, rv[3] = {0, 0, 0};
    if constexpr (WT > 0) {
        int e[WT], o[WT];
        load_slots<WT>(d, c, e, o);
        // two faces in flight at a time (their five operands each), accumulated in slot order: the
        // all-four-at-once form needs 87 registers (2 CTAs per SM, 23 % of the warps resident)
#pragma unroll
        for (int k0 = 0; k0 < WT; k0 += 2) {
            double mv[2], S[2][3], rc[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int k = k0 + q < WT ? k0 + q : WT - 1;
                const int f = e[k] >= 0 ? e[k] >> 1 : 0;
                mv[q] = d.magSf[f];
                S[q][0] = d.Sf[3 * f]; S[q][1] = d.Sf[3 * f + 1]; S[q][2] = d.Sf[3 * f + 2];
                rc[q] = d.rec[f];
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (k0 + q >= WT || e[k0 + q] < 0) continue;
                double m = mv[q];
                double sh[3] = {S[q][0] / m, S[q][1] / m, S[q][2] / m};
                double ssf = rc[q];
                for (int i = 0; i < 3; i++) {
                    for (int j = 0; j < 3; j++) T[3 * i + j] += sh[i] * S[q][j];
                    rv[i] += sh[i] * ssf;
                }
            }
        }
    } else {
        FOR_CELL_FACES(d, c)
            double m = d.magSf[f];
            double sh[3] = {d.Sf[3 * f] / m, d.Sf[3 * f + 1] / m, d.Sf[3 * f + 2] / m};
            double ssf = d.rec[f];
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++) T[3 * i + j] += sh[i] * d.Sf[3 * f + j];
                rv[i] += sh[i] * ssf;
            }
        END_CELL_FACES
    }
    double xx = T[0], xy = T[1], xz = T[2], yx = T[3], yy = T[4], yz = T[5], zx = T[6], zy = T[7], zz = T[8];
    double det = xx * (yy * zz - yz * zy) - xy * (yx * zz - yz * zx) + xz * (yx * zy - yy * zx);
    double inv[9] = {yy * zz - zy * yz, xz * zy - xy * zz, xy * yz - xz * yy,
                     zx * yz - yx * zz, xx * zz - xz * zx, yx * xz - xx * yz,
                     yx * zy - yy * zx, xy * zx - xx * zy, xx * yy - yx * xy};
    for (int i = 0; i < 3; i++) {
        double r = (inv[3 * i] / det) * rv[0] + (inv[3 * i + 1] / det) * rv[1] + (inv[3 * i + 2] / det) * rv[2];
        d.U[3 * c + i] = d.HbyA[3 * c + i] + d.rAU[c] * r;
    }
}

HD void b_Uf(const DV& d, int f) {
    double uf[3];
    if (f < d.nI) {
        int P = d.own[f], N = d.nei[f];
        double wl = d.w[f];
        for (int k = 0; k < 3; k++) uf[k] = wl * d.U[3 * P + k] + (1.0 - wl) * d.U[3 * N + k];
    } else
        for (int k = 0; k < 3; k++) uf[k] = d.U_b[3 * (f - d.nI) + k];
    double m = d.magSf[f];
    double n[3] = {d.Sf[3 * f] / m, d.Sf[3 * f + 1] / m, d.Sf[3 * f + 2] / m};
    double a = d.phi[f] / m - dot3(n, uf);
    for (int k = 0; k < 3; k++) d.Uf[3 * f + k] = uf[k] + n[k] * a;
    d.phi[f] -= d.meshPhi[f];
}

HD void b_p(const DV& d, int c) { d.p[c] = d.p_rgh[c] + d.rho[c] * d.gh[c]; }
HD void b_p_shift(const DV& d, int c) {
    double pc = d.p[c] + d.pRefShift;
    d.p[c] = pc;
    d.p_rgh[c] = pc - d.rho[c] * d.gh[c];
}

// ---- S2 rigid mesh motion ------------------------------------------------------------------------
// translation: swept volume of a rigidly translated face = Sf . dT (exact)
HD void b_meshphi_trans(const DV& d, int f) {
    const double dT[3] = {s_dT(d, 0), s_dT(d, 1), s_dT(d, 2)};
    d.meshPhi[f] = dot3(&d.Sf[3 * f], dT) / s_dt(d);
}

HD void rot3(const double* R, const double* v, double* out);
HD void rigid_point(const double* R, const double* T, const double* cofg, const double* p0, double* out) {
    double q[3] = {p0[0] - cofg[0], p0[1] - cofg[1], p0[2] - cofg[2]};
    for (int k = 0; k < 3; k++) out[k] = (R[3 * k] * q[0] + R[3 * k + 1] * q[1] + R[3 * k + 2] * q[2]) + cofg[k] + T[k];
}
HD void cross3_(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
// volume swept by a triangle whose vertices move linearly a->a1, b->b1, c->c1 (exact for the
// ruled prism): mean displacement . (A0 + A1/2 + A2/3), A(s) = A0 + s A1 + s^2 A2
HD double tri_swept(const double* a, const double* b, const double* c, const double* a1, const double* b1, const double* c1) {
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
    cross3_(e1, e2, A0);
    cross3_(e1, g2, A1a);
    cross3_(g1, e2, A1b);
    cross3_(g1, g2, A2);
    double s = 0;
    for (int k = 0; k < 3; k++) s += dm[k] * (0.5 * A0[k] + 0.25 * (A1a[k] + A1b[k]) + (1.0 / 6.0) * A2[k]);
    return s;
}
// general rigid motion (rotation): meshPhi = swept volume / deltaT, face fanned about its centre
// [OF13-MEM: face::sweptVol], point positions from the old / new rigid transforms
HD void b_meshphi_rot(const DV& d, int f) {
    int s0 = d.fOff[f], n = d.fOff[f + 1] - s0;
    const int* l = &d.fLab[s0];
    double sv = 0;
    if (n == 3) {
        double a[3], b[3], c[3], a1[3], b1[3], c1[3];
        rigid_point(d.Rold, d.To, d.cofg, &d.points0[3 * l[0]], a); rigid_point(d.R, d.Tn, d.cofg, &d.points0[3 * l[0]], a1);
        rigid_point(d.Rold, d.To, d.cofg, &d.points0[3 * l[1]], b); rigid_point(d.R, d.Tn, d.cofg, &d.points0[3 * l[1]], b1);
        rigid_point(d.Rold, d.To, d.cofg, &d.points0[3 * l[2]], c); rigid_point(d.R, d.Tn, d.cofg, &d.points0[3 * l[2]], c1);
        sv = tri_swept(a, b, c, a1, b1, c1);
    } else {
        double fc0[3], fc1[3];
        rigid_point(d.Rold, d.To, d.cofg, &d.Cf0[3 * f], fc0);
        rigid_point(d.R, d.Tn, d.cofg, &d.Cf0[3 * f], fc1);
        for (int i = 0; i < n; i++) {
            int j = (i + 1) % n;
            double b[3], c[3], b1[3], c1[3];
            rigid_point(d.Rold, d.To, d.cofg, &d.points0[3 * l[i]], b); rigid_point(d.R, d.Tn, d.cofg, &d.points0[3 * l[i]], b1);
            rigid_point(d.Rold, d.To, d.cofg, &d.points0[3 * l[j]], c); rigid_point(d.R, d.Tn, d.cofg, &d.points0[3 * l[j]], c1);
            sv += tri_swept(fc0, b, c, fc1, b1, c1);
        }
    }
    d.meshPhi[f] = sv / d.dt;
}

HD void rot3(const double* R, const double* v, double* out) {
    for (int k = 0; k < 3; k++) out[k] = R[3 * k] * v[0] + R[3 * k + 1] * v[1] + R[3 * k + 2] * v[2];
}
// rotation: re-orient the body-frame face vectors; gh/ghf follow the moved centres
HD void b_rotate_face(const DV& d, int f) {
    rot3(d.R, &d.Sf0[3 * f], &d.Sf[3 * f]);
    if (f < d.nI) {
        rot3(d.R, &d.dPN0[3 * f], &d.dPN[3 * f]);
        rot3(d.R, &d.corrVec0[3 * f], &d.corrVec[3 * f]);
    }
}
HD void b_gh_face(const DV& d, int f) {
    double q[3] = {d.Cf0[3 * f] - d.cofg[0], d.Cf0[3 * f + 1] - d.cofg[1], d.Cf0[3 * f + 2] - d.cofg[2]}, x[3];
    rot3(d.R, q, x);
    for (int k = 0; k < 3; k++) x[k] = x[k] + d.cofg[k] + s_Tn(d, k);
    d.ghf[f] = dot3(d.g, x);
}
HD void b_gh_cell(const DV& d, int c) {
    double q[3] = {d.C0[3 * c] - d.cofg[0], d.C0[3 * c + 1] - d.cofg[1], d.C0[3 * c + 2] - d.cofg[2]}, x[3];
    rot3(d.R, q, x);
    for (int k = 0; k < 3; k++) x[k] = x[k] + d.cofg[k] + s_Tn(d, k);
    d.gh[c] = dot3(d.g, x);
}

// halo send buffer: the owner-cell value behind every processor face, in ghost order
HD void b_pack_halo(const DV& d, int j) {
    int c = d.procOwner[j];
    for (int k = 0; k < d.xnc; k++) d.xbuf[(size_t)j * d.xnc + k] = d.xsrc[(size_t)c * d.xnc + k];
}

// face arrays between device order (processor faces after the internal ones) and file order;
// procOwner temporarily carries the permutation dev -> file
HD void b_face_to_file(const DV& d, int fd) {
    int ff = d.procOwner[fd];
    for (int k = 0; k < d.xnc; k++) d.xbuf[(size_t)ff * d.xnc + k] = d.xsrc[(size_t)fd * d.xnc + k];
}
HD void b_file_to_face(const DV& d, int fd) {
    int ff = d.procOwner[fd];
    for (int k = 0; k < d.xnc; k++) d.xbuf[(size_t)fd * d.xnc + k] = d.xsrc[(size_t)ff * d.xnc + k];
}

DEF_KERNEL(pack_halo, DV)
DEF_KERNEL(face_to_file, DV)
DEF_KERNEL(file_to_face, DV)
DEF_KERNEL_W(courant)
DEF_KERNEL(alpha_bc, DV)
DEF_KERNEL(U_bc, DV)
DEF_KERNEL(p_total, DV)
DEF_KERNEL(p_evaluate, DV)
DEF_KERNEL_W(grad_scalar)
DEF_KERNEL(alpha_flux, DV)
DEF_KERNEL_W(mules_setup)
DEF_KERNEL_W(mules_cell)
DEF_KERNEL(mules_face, DV)
DEF_KERNEL(mules_phipsi, DV)
DEF_KERNEL(mules_face_final, DV)
DEF_KERNEL(alphaphi_acc, DV)
DEF_KERNEL_W(mules_update)
DEF_KERNEL(nhat_face, DV)
DEF_KERNEL_W(curvature)
DEF_KERNEL(stf_face, DV)
DEF_KERNEL(mixture_cell, DV)
DEF_KERNEL(mixture_bnd, DV)
DEF_KERNEL(rhophi, DV)
DEF_KERNEL_WB(grad_U, 3)
DEF_KERNEL(mom_face, DV)
DEF_KERNEL(mom_bnd, DV)
DEF_KERNEL_W(mom_cell)
DEF_KERNEL_WB(HbyA, 3)
DEF_KERNEL(HbyA_bnd, DV)
DEF_KERNEL(phiHbyA, DV)
DEF_KERNEL(p_face, DV)
DEF_KERNEL_W(p_cell)
DEF_KERNEL(flux, DV)
DEF_KERNEL_WB(U_recon, 3)  // 9 + 3 accumulators and 2 x 5 operands in flight
DEF_KERNEL(Uf, DV)
DEF_KERNEL(p, DV)
DEF_KERNEL(p_shift, DV)
DEF_KERNEL(meshphi_trans, DV)
DEF_KERNEL(meshphi_rot, DV)
DEF_KERNEL(rotate_face, DV)
DEF_KERNEL(gh_face, DV)
DEF_KERNEL(gh_cell, DV)

}  // namespace tpp
