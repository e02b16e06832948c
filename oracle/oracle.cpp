// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's arithmetic for the grad(U) -> strain rate -> nu ->
// mixture -> heat-source path of kckincaid/solidStateAMFoam (chtMultiRegionInterFoam).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product (solidstateamfoam_b200/) never does.
//
// PARITY PARTLY PINNED.  The reference ships no tests, fixtures or example cases and the solver
// cannot be built here (it needs OpenFOAM v1806, absent offline).  Two tiers therefore:
//  * PINNED against the reference's own source text: the closed-form model expressions -- SheppardWright, FKS,
//    NortonHoff, the mixture's mu/nu/k/cp/rho/muf/nuf, constant/exponentialSlipStick qvisc/qfric, both
//    rotational-slip boundary conditions.  Their translation units are compiled UNMODIFIED from /root/reference
//    against stand-in OpenFOAM headers (oracle/ofstub, oracle/Makefile target `ref` -> oracle/_ref) and this file
//    must match them bit for bit (tests/test_reference_sources.py; outputs frozen in tests/golden/
//    reference_sources.npz for machines without the reference tree).
//  * PARITY UNPINNED: everything that is OpenFOAM library behaviour rather than reference source -- the
//    Gauss-linear gradient, face interpolation, boundary/processor-patch evaluation, field-algebra semantics
//    (also assumed by the stand-in headers), interfaceProperties::calculateK, and fluid/calculateStrain.H.
//    These are restatements: the closed-form expressions follow
// the cited reference lines operator by operator (C++ left-to-right association, one rounded
// FP64 operation per operator, no FMA: build with -ffp-contract=off), and the OpenFOAM
// semantics it relies on (Gauss-linear gradient, symm/mag, max/min argument order, boundary
// evaluation) are restated from OpenFOAM v1806 as described in SURVEY.md Appendix B.  They are
// pinned instead by (1) mpmath >=50-digit known-answer vectors (tests/golden), (2) analytic
// gradient cases, (3) two independent forms agreeing bit for bit:
//   form 0 "faithful": field-at-a-time, one loop + one temporary per operator over internal
//                      AND boundary values, the pass structure of the reference (this is also
//                      the timed CPU baseline);
//   form 1 "fused":    one per-cell / per-face evaluation of the same expression trees.
//
// Field storage: AoS [nCells + nBnd][ncomp], internal values then boundary values.

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/ssam_gpu.h"  // POD parameter structs + field ids only

namespace {

constexpr double SMALL = 1e-15;
constexpr double VSMALL = 1e-300;

typedef std::vector<double> Fld;

struct Case {
    int32_t nCells = 0, nIF = 0, nFaces = 0, nBnd = 0, nPatches = 0, nTot = 0;
    std::vector<int32_t> owner, neighbour, patchStart, patchSize, patchKind, patchNbrRank, patchUbc;
    std::vector<double> Sf, w, V, C, Cf_b, magSf_b, deltaCoeffs_b;
    Fld f[SSAM_F_COUNT];
    bool set[SSAM_F_COUNT] = {false};
    double fricCentre[3] = {0, 0, 0};
    double fricArea = 0;
    std::string err;
    // derived integer addressing (built lazily)
    std::vector<int32_t> ownerStart, extraStart, extraFace, extraOther, cellFacesStart, cellFaces, haloSend,
        haloOffsets, fricCells, faceColour;
    bool addrBuilt = false;
    Fld nHatf;  // surfaceScalarField: internal faces then boundary faces
    Fld muf, nuf;
    Fld phi, phic;     // face flux (input) and the alphaEqn.H compression coefficient
    Fld procNbrC;      // [3*nBnd] cell centre of the neighbour rank's cell behind each processor face (else 0)
};

int ncompOf(int field) {
    return (field == SSAM_F_U || field == SSAM_F_GRADALPHA || field == SSAM_F_DIVDEVRHOREFF) ? 3
                                                                                            : (field == SSAM_F_GRADU ? 9 : 1);
}

Fld& fieldRef(Case& c, int field) {
    Fld& v = c.f[field];
    if (v.empty()) v.assign(size_t(c.nTot) * ncompOf(field), 0.0);
    return v;
}

// ---- OpenFOAM scalar semantics ---------------------------------------------------------------
inline double ofMax(double a, double b) { return (a > b) ? a : b; }  // Foam::max(a,b)
inline double ofMin(double a, double b) { return (a < b) ? a : b; }  // Foam::min(a,b)

// ::pow as the reference calls it: a libm call with a run-time exponent.  Without the barrier gcc
// folds pow(x, 2) into x*x in the per-cell form, which differs from glibc's pow by 1 ulp for about
// one input in 10^4.
__attribute__((noinline)) double libmPow(double x, double y) {
    volatile double yy = y;
    return std::pow(x, yy);
}

// ---- field-at-a-time operators (form 0): each is one pass + one temporary ---------------------
#define LOOP(n) for (size_t i = 0, n_ = (n); i < n_; ++i)
Fld mulSF(double s, const Fld& a) { Fld r(a.size()); LOOP(a.size()) r[i] = s * a[i]; return r; }
Fld mulFS(const Fld& a, double s) { Fld r(a.size()); LOOP(a.size()) r[i] = a[i] * s; return r; }
Fld mulFF(const Fld& a, const Fld& b) { Fld r(a.size()); LOOP(a.size()) r[i] = a[i] * b[i]; return r; }
Fld divFS(const Fld& a, double s) { Fld r(a.size()); LOOP(a.size()) r[i] = a[i] / s; return r; }
Fld divSF(double s, const Fld& a) { Fld r(a.size()); LOOP(a.size()) r[i] = s / a[i]; return r; }
Fld divFF(const Fld& a, const Fld& b) { Fld r(a.size()); LOOP(a.size()) r[i] = a[i] / b[i]; return r; }
Fld addFF(const Fld& a, const Fld& b) { Fld r(a.size()); LOOP(a.size()) r[i] = a[i] + b[i]; return r; }
Fld addFS(const Fld& a, double s) { Fld r(a.size()); LOOP(a.size()) r[i] = a[i] + s; return r; }
Fld addSF(double s, const Fld& a) { Fld r(a.size()); LOOP(a.size()) r[i] = s + a[i]; return r; }
Fld subFS(const Fld& a, double s) { Fld r(a.size()); LOOP(a.size()) r[i] = a[i] - s; return r; }
Fld subSF(double s, const Fld& a) { Fld r(a.size()); LOOP(a.size()) r[i] = s - a[i]; return r; }
Fld powFS(const Fld& a, double s) { Fld r(a.size()); LOOP(a.size()) r[i] = libmPow(a[i], s); return r; }
Fld expF(const Fld& a) { Fld r(a.size()); LOOP(a.size()) r[i] = std::exp(a[i]); return r; }
Fld logF(const Fld& a) { Fld r(a.size()); LOOP(a.size()) r[i] = std::log(a[i]); return r; }
Fld asinhF(const Fld& a) { Fld r(a.size()); LOOP(a.size()) r[i] = std::asinh(a[i]); return r; }
Fld maxFS(const Fld& a, double s) { Fld r(a.size()); LOOP(a.size()) r[i] = ofMax(a[i], s); return r; }
Fld maxSF(double s, const Fld& a) { Fld r(a.size()); LOOP(a.size()) r[i] = ofMax(s, a[i]); return r; }
Fld minFS(const Fld& a, double s) { Fld r(a.size()); LOOP(a.size()) r[i] = ofMin(a[i], s); return r; }
Fld minSF(double s, const Fld& a) { Fld r(a.size()); LOOP(a.size()) r[i] = ofMin(s, a[i]); return r; }

// ---- tensor algebra (OpenFOAM SymmTensorI.H / TensorI.H restated) ------------
inline double magSymm(const double* g) {
    const double xx = g[0], xy = 0.5 * (g[1] + g[3]), xz = 0.5 * (g[2] + g[6]);
    const double yy = g[4], yz = 0.5 * (g[5] + g[7]), zz = g[8];
    const double ms = xx * xx + 2 * (xy * xy) + 2 * (xz * xz) + yy * yy + 2 * (yz * yz) + zz * zz;
    return std::sqrt(ms);
}
inline double magVec(double x, double y, double z) { return std::sqrt(x * x + y * y + z * z); }

// ================================================================================================
// fvc::grad(U), gradSchemes "Gauss linear" (OpenFOAM gaussGrad<vector>::calcGrad/gradf and
// surfaceInterpolationScheme::interpolate restated). Called by the reference at
// fluid/TEqn.H:3, stickModel.C:60, NortonHoff.C:62 (via viscosityModel::strainRate()).
// ================================================================================================
void gradBoundaryCorrect(const Case& c, const Fld& U, Fld& G) {
    // step 5: extrapolatedCalculated patches = adjacent cell value; processor patches keep the
    //         neighbour rank's adjacent-cell value (delivered in G's boundary part by the halo step)
    // step 6: non-coupled patches: g_b += n (x) (snGrad(U)_b - (n . g_b))
    for (int p = 0; p < c.nPatches; ++p) {
        const bool coupled = c.patchKind[p] == SSAM_PATCH_PROCESSOR;
        for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) {
            const int32_t b = f - c.nIF, P = c.owner[f];
            double* gb = &G[size_t(c.nCells + b) * 9];
            if (coupled) continue;  // filled by oracle halo exchange (or left as is)
            const double* gP = &G[size_t(P) * 9];
            for (int k = 0; k < 9; ++k) gb[k] = gP[k];
            const double ms = c.magSf_b[b];
            const double n[3] = {c.Sf[3 * size_t(f)] / ms, c.Sf[3 * size_t(f) + 1] / ms, c.Sf[3 * size_t(f) + 2] / ms};
            double sn[3];
            const double* Ub = &U[size_t(c.nCells + b) * 3];
            const double* UP = &U[size_t(P) * 3];
            for (int j = 0; j < 3; ++j)
                sn[j] = (c.patchUbc[p] == SSAM_BC_ZERO_GRADIENT) ? 0.0 : c.deltaCoeffs_b[b] * (Ub[j] - UP[j]);
            double ng[3];
            for (int j = 0; j < 3; ++j) ng[j] = n[0] * gb[j] + n[1] * gb[3 + j] + n[2] * gb[6 + j];
            double d[3];
            for (int j = 0; j < 3; ++j) d[j] = sn[j] - ng[j];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) gb[3 * i + j] += n[i] * d[j];
        }
    }
}

inline void faceValue(const Case& c, const Fld& U, int32_t f, double* Uf) {
    // internal faces: lambda*(vf[P] - vf[N]) + vf[N]  (surfaceInterpolationScheme::interpolate)
    const double* UP = &U[size_t(c.owner[f]) * 3];
    const double* UN = &U[size_t(c.neighbour[f]) * 3];
    const double w = c.w[f];
    for (int j = 0; j < 3; ++j) Uf[j] = w * (UP[j] - UN[j]) + UN[j];
}

inline void boundaryFaceValue(const Case& c, const Fld& U, int p, int32_t f, double* Uf) {
    const int32_t b = f - c.nIF;
    const double* Ub = &U[size_t(c.nCells + b) * 3];
    if (c.patchKind[p] == SSAM_PATCH_PROCESSOR) {
        const double* UP = &U[size_t(c.owner[f]) * 3];
        const double w = c.w[f];
        for (int j = 0; j < 3; ++j) Uf[j] = w * UP[j] + (1.0 - w) * Ub[j];
    } else {
        for (int j = 0; j < 3; ++j) Uf[j] = Ub[j];
    }
}

void gradFaithful(const Case& c, const Fld& U, Fld& G) {
    // pass 1: face interpolation (writes a full surfaceVectorField)
    Fld Uf(size_t(c.nFaces) * 3);
    for (int32_t f = 0; f < c.nIF; ++f) faceValue(c, U, f, &Uf[size_t(f) * 3]);
    for (int p = 0; p < c.nPatches; ++p)
        for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f)
            boundaryFaceValue(c, U, p, f, &Uf[size_t(f) * 3]);
    // pass 2: owner/neighbour scatter
    std::fill(G.begin(), G.begin() + size_t(c.nCells) * 9, 0.0);
    for (int32_t f = 0; f < c.nIF; ++f) {
        const double* S = &c.Sf[size_t(f) * 3];
        const double* u = &Uf[size_t(f) * 3];
        double* go = &G[size_t(c.owner[f]) * 9];
        double* gn = &G[size_t(c.neighbour[f]) * 9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double t = S[i] * u[j];
                go[3 * i + j] += t;
                gn[3 * i + j] -= t;
            }
    }
    // pass 3: boundary faces, patch by patch
    for (int p = 0; p < c.nPatches; ++p)
        for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) {
            const double* S = &c.Sf[size_t(f) * 3];
            const double* u = &Uf[size_t(f) * 3];
            double* go = &G[size_t(c.owner[f]) * 9];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) go[3 * i + j] += S[i] * u[j];
        }
    // pass 4: divide by cell volume
    for (int32_t cc = 0; cc < c.nCells; ++cc)
        for (int k = 0; k < 9; ++k) G[size_t(cc) * 9 + k] /= c.V[cc];
    gradBoundaryCorrect(c, U, G);
}

void buildAddressing(Case& c);

void gradFused(Case& c, const Fld& U, Fld& G) {
    // per-cell gather over the cell's faces in ascending face index (the order the scatter of
    // gradFaithful produces in every cell), then the same boundary treatment
    buildAddressing(c);
    std::vector<int32_t> bPatch(c.nBnd);
    for (int p = 0; p < c.nPatches; ++p)
        for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) bPatch[f - c.nIF] = p;
    for (int32_t cc = 0; cc < c.nCells; ++cc) {
        double g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int32_t k = c.cellFacesStart[cc]; k < c.cellFacesStart[cc + 1]; ++k) {
            const uint32_t e = uint32_t(c.cellFaces[k]);
            const int32_t f = int32_t(e & 0x7fffffffu);
            const bool isNei = (e >> 31) != 0;
            double Uf[3];
            if (f < c.nIF)
                faceValue(c, U, f, Uf);
            else
                boundaryFaceValue(c, U, bPatch[f - c.nIF], f, Uf);
            const double* S = &c.Sf[size_t(f) * 3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    const double t = S[i] * Uf[j];
                    if (isNei)
                        g[3 * i + j] -= t;
                    else
                        g[3 * i + j] += t;
                }
        }
        for (int k = 0; k < 9; ++k) G[size_t(cc) * 9 + k] = g[k] / c.V[cc];
    }
    gradBoundaryCorrect(c, U, G);
}

// factor*mag(symm(gradU)) on internal + boundary values
void strainFromGrad(const Case& c, const Fld& G, double factor, Fld& out, int form) {
    if (form == 0) {
        // symm -> mag -> scale as separate passes (values identical to the fused form)
        Fld m(c.nTot);
        for (int32_t i = 0; i < c.nTot; ++i) m[i] = magSymm(&G[size_t(i) * 9]);
        out = mulSF(factor, m);
    } else {
        for (int32_t i = 0; i < c.nTot; ++i) out[i] = factor * magSymm(&G[size_t(i) * 9]);
    }
}

// ================================================================================================
// integer addressing (the GPU-built tables must equal these byte for byte)
// ================================================================================================
void buildAddressing(Case& c) {
    if (c.addrBuilt) return;
    const int32_t nc = c.nCells;
    // lduAddressing::ownerStart: faces owned by cell i are [ownerStart[i], ownerStart[i+1])
    c.ownerStart.assign(nc + 1, 0);
    for (int32_t f = 0; f < c.nIF; ++f) c.ownerStart[c.owner[f] + 1]++;
    for (int32_t i = 0; i < nc; ++i) c.ownerStart[i + 1] += c.ownerStart[i];
    // extra list: neighbour-side internal faces (ascending) then boundary faces (ascending)
    c.extraStart.assign(nc + 1, 0);
    for (int32_t f = 0; f < c.nIF; ++f) c.extraStart[c.neighbour[f] + 1]++;
    for (int32_t f = c.nIF; f < c.nFaces; ++f) c.extraStart[c.owner[f] + 1]++;
    for (int32_t i = 0; i < nc; ++i) c.extraStart[i + 1] += c.extraStart[i];
    const int32_t nExtra = c.extraStart[nc];
    c.extraFace.assign(nExtra, 0);
    c.extraOther.assign(nExtra, 0);
    std::vector<int32_t> fill(c.extraStart.begin(), c.extraStart.end() - 1);
    for (int32_t f = 0; f < c.nIF; ++f) {
        const int32_t k = fill[c.neighbour[f]]++;
        c.extraFace[k] = f;
        c.extraOther[k] = c.owner[f];
    }
    for (int p = 0; p < c.nPatches; ++p)
        for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) {
            const int32_t k = fill[c.owner[f]]++;
            c.extraFace[k] = f;
            c.extraOther[k] = (c.patchKind[p] == SSAM_PATCH_PROCESSOR) ? -2 : -1;
        }
    // full cell->face CSR ascending: neighbour-side (bit 31), owned internal, boundary
    c.cellFacesStart.assign(nc + 1, 0);
    for (int32_t i = 0; i < nc; ++i)
        c.cellFacesStart[i + 1] =
            c.cellFacesStart[i] + (c.ownerStart[i + 1] - c.ownerStart[i]) + (c.extraStart[i + 1] - c.extraStart[i]);
    c.cellFaces.assign(c.cellFacesStart[nc], 0);
    for (int32_t i = 0; i < nc; ++i) {
        int32_t k = c.cellFacesStart[i];
        int32_t e = c.extraStart[i];
        for (; e < c.extraStart[i + 1] && c.extraFace[e] < c.nIF; ++e)
            c.cellFaces[k++] = int32_t(uint32_t(c.extraFace[e]) | 0x80000000u);
        for (int32_t f = c.ownerStart[i]; f < c.ownerStart[i + 1]; ++f) c.cellFaces[k++] = f;
        for (; e < c.extraStart[i + 1]; ++e) c.cellFaces[k++] = c.extraFace[e];
    }
    // halo send maps: faceCells of each processor patch in patch-face order
    c.haloOffsets.assign(1, 0);
    c.haloSend.clear();
    for (int p = 0; p < c.nPatches; ++p) {
        if (c.patchKind[p] != SSAM_PATCH_PROCESSOR) continue;
        for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) c.haloSend.push_back(c.owner[f]);
        c.haloOffsets.push_back(int32_t(c.haloSend.size()));
    }
    // greedy face colouring: faces ascending, colour = smallest non-negative integer not yet used
    // by any lower-numbered face of either adjacent cell (scatter variant of the gradient)
    c.faceColour.assign(c.nIF, -1);
    {
        std::vector<uint64_t> used(nc, 0);
        for (int32_t f = 0; f < c.nIF; ++f) {
            const uint64_t u = used[c.owner[f]] | used[c.neighbour[f]];
            int col = 0;
            while (col < 63 && ((u >> col) & 1)) ++col;
            c.faceColour[f] = col;
            used[c.owner[f]] |= (uint64_t(1) << col);
            used[c.neighbour[f]] |= (uint64_t(1) << col);
        }
    }
    c.addrBuilt = true;
}

// ================================================================================================
// viscosity models
// ================================================================================================
// --- SheppardWright (viscoPlasticModels/SheppardWright/SheppardWright.C) -------------------------
namespace sw {
// :64   Z = epsilonDotEq_*exp(Q_/(R_*temp_))
Fld Z(const ssam_viscosity_params& p, const Fld& T, const Fld& e) { return mulFF(e, expF(divSF(p.Q, mulSF(p.R, T)))); }
// :74   sigmaR_*asinh(pow(Z()/A_, 1/n_))
Fld sigmaf(const ssam_viscosity_params& p, const Fld& T, const Fld& e) {
    return mulSF(p.sigmaR, asinhF(powFS(divFS(Z(p, T, e), p.A), 1 / p.n)));
}
// :91   sigmaR_*asinh(pow(e01*exp(Q_/(R_*temp_))/A_, 1/n_))
Fld sigmay(const ssam_viscosity_params& p, const Fld& T) {
    return mulSF(p.sigmaR, asinhF(powFS(divFS(mulSF(0.1, expF(divSF(p.Q, mulSF(p.R, T)))), p.A), 1 / p.n)));
}
// :109-117  max(nuMin_, min(nuMax_, sigmaf()/(3.0*rho_*max(epsilonDotEq_, eDotMin_))))
Fld calcNu(const ssam_viscosity_params& p, const Fld& T, const Fld& e) {
    return maxSF(p.nuMin, minSF(p.nuMax, divFF(sigmaf(p, T, e), mulSF(3.0 * p.rho, maxFS(e, p.eDotMin)))));
}
// per-element form of the same trees
inline void cell(const ssam_viscosity_params& p, double T, double e, double& nu, double& sf, double& sy) {
    const double ex = std::exp(p.Q / (p.R * T));
    const double Zv = e * ex;
    sf = p.sigmaR * std::asinh(libmPow(Zv / p.A, 1 / p.n));
    sy = p.sigmaR * std::asinh(libmPow((0.1 * ex) / p.A, 1 / p.n));
    nu = ofMax(p.nuMin, ofMin(p.nuMax, sf / ((3.0 * p.rho) * ofMax(e, p.eDotMin))));
}
}  // namespace sw

// --- FKS (viscoPlasticModels/FKS/FKS.C) -----------------------------------------------------------
namespace fks {
// :108,:112  Tlim = min(max(T_, Ta_), Tm_); (Tlim_ - Ta_)/(Tm_ - Ta_)
Fld Tstar(const ssam_viscosity_params& p, const Fld T /* by value, FKS.C:105 */) {
    Fld Tlim = minFS(maxFS(T, p.Ta), p.Tm);
    return divFS(subFS(Tlim, p.Ta), p.Tm - p.Ta);
}
// :73  (a1_ + (a2_*1.571))*temp_/temp_
Fld H(const ssam_viscosity_params& p, const Fld& T) { return divFF(mulSF(p.a1 + (p.a2 * 1.571), T), T); }
// :89  1.0 + (b1_*pow((Th_ + Ts_), b2_))*(b3_*Foam::log(edot_/e0_ + VSMALL))
Fld Lambda(const ssam_viscosity_params& p, const Fld Th /* by value, :79 */, const Fld& e) {
    return addSF(1.0, mulFF(mulSF(p.b1, powFS(addFS(Th, SMALL), p.b2)), mulSF(p.b3, logF(addFS(divFS(e, p.e0), VSMALL)))));
}
// :99  1.0 - 1.0/pow(1.0 + exp(-c1_*Th_), 1/c2_)
Fld Theta(const ssam_viscosity_params& p, const Fld Th /* by value, :95 */) {
    return subSF(1.0, divSF(1.0, powFS(addSF(1.0, expF(mulSF(-p.c1, Th))), 1 / p.c2)));
}
// :129  H()*Lambda(Th_)*Theta(Th_)
Fld sigmaf(const ssam_viscosity_params& p, const Fld& T, const Fld& e) {
    Fld Th = Tstar(p, T);
    return mulFF(mulFF(H(p, T), Lambda(p, Th, e)), Theta(p, Th));
}
// :152  a1_*Theta(Th_)*(1.0 + (b1_*pow((Th_ + Ts_), b2_))*(b3_*Foam::log(e01_/e0_ + VSMALL)))
Fld sigmay(const ssam_viscosity_params& p, const Fld& T) {
    Fld Th = Tstar(p, T);
    const double lg = p.b3 * std::log(0.1 / p.e0 + VSMALL);
    return mulFF(mulSF(p.a1, Theta(p, Th)), addSF(1.0, mulFS(mulSF(p.b1, powFS(addFS(Th, SMALL), p.b2)), lg)));
}
// :172-181
Fld calcNu(const ssam_viscosity_params& p, const Fld& T, const Fld& e) {
    return maxSF(p.nuMin, minSF(p.nuMax, divFF(sigmaf(p, T, e), mulSF(3.0 * p.rho, maxFS(e, p.eDotMin)))));
}
inline void cell(const ssam_viscosity_params& p, double T, double e, double& nu, double& sf, double& sy) {
    const double Tlim = ofMin(ofMax(T, p.Ta), p.Tm);
    const double Th = (Tlim - p.Ta) / (p.Tm - p.Ta);
    const double Hh = ((p.a1 + (p.a2 * 1.571)) * T) / T;
    const double pw = p.b1 * libmPow(Th + SMALL, p.b2);
    const double Lam = 1.0 + pw * (p.b3 * std::log(e / p.e0 + VSMALL));
    const double Th_ = 1.0 - 1.0 / libmPow(1.0 + std::exp((-p.c1) * Th), 1 / p.c2);
    sf = (Hh * Lam) * Th_;
    sy = (p.a1 * Th_) * (1.0 + pw * (p.b3 * std::log(0.1 / p.e0 + VSMALL)));
    nu = ofMax(p.nuMin, ofMin(p.nuMax, sf / ((3.0 * p.rho) * ofMax(e, p.eDotMin))));
}
}  // namespace fks

// --- NortonHoff (viscoPlasticModels/NortonHoff/NortonHoff.C:53-66) ---------------------------------
namespace nh {
// (mu0_/rho_)*pow(sqrt(3.0)*max(strainRate(),strainRateEffMin_)/strainRateUnits_, (m_-1))
Fld calcNu(const ssam_viscosity_params& p, const Fld& sr) {
    return maxSF(p.nuMin,
                 minSF(p.nuMax, mulSF(p.mu0 / p.rho, powFS(divFS(mulSF(std::sqrt(3.0), maxFS(sr, p.strainRateEffMin)), 1.0), p.m - 1))));
}
inline double cell(const ssam_viscosity_params& p, double sr) {
    return ofMax(p.nuMin, ofMin(p.nuMax, (p.mu0 / p.rho) * libmPow((std::sqrt(3.0) * ofMax(sr, p.strainRateEffMin)) / 1.0, p.m - 1)));
}
}  // namespace nh

int viscosityCorrect(Case& c, int phase, const ssam_viscosity_params& p, int form) {
    Fld& nu = fieldRef(c, phase == 1 ? SSAM_F_NU1 : SSAM_F_NU2);
    c.set[phase == 1 ? SSAM_F_NU1 : SSAM_F_NU2] = true;
    const size_t n = c.nTot;
    switch (p.model) {
        case SSAM_VISC_NEWTONIAN:
            for (size_t i = 0; i < n; ++i) nu[i] = p.nu;  // Newtonian::correct() is a no-op on a constant nu
            return 0;
        case SSAM_VISC_SHEPPARD_WRIGHT:
        case SSAM_VISC_FKS: {
            if (!c.set[SSAM_F_T] || !c.set[SSAM_F_EPSDOT]) {
                c.err = "lookupObject: temp / epsilonDotEq not registered";
                return SSAM_ERR_MISSING_FIELD;
            }
            const Fld& T = c.f[SSAM_F_T];
            const Fld& e = c.f[SSAM_F_EPSDOT];
            Fld& sf = fieldRef(c, SSAM_F_SIGMAF);
            Fld& sy = fieldRef(c, SSAM_F_SIGMAY);
            c.set[SSAM_F_SIGMAF] = c.set[SSAM_F_SIGMAY] = true;
            if (form == 0) {
                if (p.model == SSAM_VISC_SHEPPARD_WRIGHT) {  // SheppardWright.H:135-137
                    nu = sw::calcNu(p, T, e);
                    sf = sw::sigmaf(p, T, e);
                    sy = sw::sigmay(p, T);
                } else {  // FKS.H:158-160
                    sf = fks::sigmaf(p, T, e);
                    sy = fks::sigmay(p, T);
                    nu = fks::calcNu(p, T, e);
                }
            } else {
                for (size_t i = 0; i < n; ++i) {
                    if (p.model == SSAM_VISC_SHEPPARD_WRIGHT)
                        sw::cell(p, T[i], e[i], nu[i], sf[i], sy[i]);
                    else
                        fks::cell(p, T[i], e[i], nu[i], sf[i], sy[i]);
                }
            }
            return 0;
        }
        case SSAM_VISC_NORTON_HOFF: {
            if (!c.set[SSAM_F_U]) {
                c.err = "U not set";
                return SSAM_ERR_MISSING_FIELD;
            }
            // viscosityModel::strainRate() = sqrt(2.0)*mag(symm(fvc::grad(U_)))
            Fld& G = fieldRef(c, SSAM_F_GRADU);
            Fld& sr = fieldRef(c, SSAM_F_STRAINRATE);
            if (form == 0)
                gradFaithful(c, c.f[SSAM_F_U], G);
            else
                gradFused(c, c.f[SSAM_F_U], G);
            strainFromGrad(c, G, std::sqrt(2.0), sr, form);
            c.set[SSAM_F_GRADU] = c.set[SSAM_F_STRAINRATE] = true;
            if (form == 0)
                nu = nh::calcNu(p, sr);
            else
                for (size_t i = 0; i < n; ++i) nu[i] = nh::cell(p, sr[i]);
            return 0;
        }
        default:
            c.err = "Unknown viscosityModel type";
            return SSAM_ERR_UNKNOWN_MODEL;
    }
}

// ================================================================================================
// mixture (chtMultiRegionInterFoam/incompressibleTwoPhaseThermalMixture/...C)
// ================================================================================================
Fld limitedAlpha(const Fld& a) { return minFS(maxFS(a, 0.0), 1.0); }  // :49,:138
inline double la1(double a) { return ofMin(ofMax(a, 0.0), 1.0); }

// :149-150  la*la*rho1*nu1 + (1-la)*(1-la)*rho2*nu2
Fld muFaithful(const Case& c, const ssam_mixture_params& p) {
    Fld la = limitedAlpha(c.f[SSAM_F_ALPHA1]);
    Fld oml = subSF(1.0, la);
    return addFF(mulFF(mulFS(mulFF(la, la), p.rho1), c.f[SSAM_F_NU1]),
                 mulFF(mulFS(mulFF(oml, subSF(1.0, la)), p.rho2), c.f[SSAM_F_NU2]));
}
inline double muCell(const ssam_mixture_params& p, double a, double nu1, double nu2) {
    const double la = la1(a);
    return ((la * la) * p.rho1) * nu1 + (((1.0 - la) * (1.0 - la)) * p.rho2) * nu2;
}
inline double rhoMixCell(const ssam_mixture_params& p, double a) {
    const double la = la1(a);
    return la * p.rho1 + (1.0 - la) * p.rho2;
}

int needNu(Case& c) {
    if (!c.set[SSAM_F_ALPHA1] || !c.set[SSAM_F_NU1] || !c.set[SSAM_F_NU2]) {
        c.err = "alpha1 / nu1 / nu2 not available";
        return SSAM_ERR_MISSING_FIELD;
    }
    return 0;
}

int mixtureMu(Case& c, const ssam_mixture_params& p, int form) {
    if (int e = needNu(c)) return e;
    Fld& mu = fieldRef(c, SSAM_F_MUEFF);
    if (form == 0)
        mu = muFaithful(c, p);
    else
        for (int32_t i = 0; i < c.nTot; ++i) mu[i] = muCell(p, c.f[SSAM_F_ALPHA1][i], c.f[SSAM_F_NU1][i], c.f[SSAM_F_NU2][i]);
    c.set[SSAM_F_MUEFF] = true;
    return 0;
}

// :46-53  nu_ = mu()/(limitedAlpha1*rho1_ + (scalar(1) - limitedAlpha1)*rho2_)
// fvc::interpolate(vf) with the linear scheme (surfaceInterpolationScheme::interpolate restated)
Fld interpolateLinear(const Case& c, const Fld& vf) {
    Fld sf(c.nFaces);
    for (int32_t f = 0; f < c.nIF; ++f) sf[f] = c.w[f] * (vf[c.owner[f]] - vf[c.neighbour[f]]) + vf[c.neighbour[f]];
    for (int p = 0; p < c.nPatches; ++p)
        for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) {
            const double vb = vf[c.nCells + (f - c.nIF)];
            sf[f] = (c.patchKind[p] == SSAM_PATCH_PROCESSOR) ? c.w[f] * vf[c.owner[f]] + (1.0 - c.w[f]) * vb : vb;
        }
    return sf;
}

// muf() incompressibleTwoPhaseThermalMixture.C:163-180, nuf() :183-202
int mixtureFaceVisc(Case& c, const ssam_mixture_params& p, bool nuf) {
    if (!c.set[SSAM_F_ALPHA1] || !c.set[SSAM_F_NU1] || !c.set[SSAM_F_NU2]) {
        c.err = "alpha1 / nu1 / nu2 not set";
        return SSAM_ERR_MISSING_FIELD;
    }
    const Fld alpha1f = minFS(maxFS(interpolateLinear(c, c.f[SSAM_F_ALPHA1]), 0.0), 1.0);  // :166-169, :186-189
    const Fld n1f = interpolateLinear(c, c.f[SSAM_F_NU1]), n2f = interpolateLinear(c, c.f[SSAM_F_NU2]);
    // alpha1f*rho1_*interpolate(nu1) + (scalar(1) - alpha1f)*rho2_*interpolate(nu2)      :176-177, :197-198
    Fld num = addFF(mulFF(mulFS(alpha1f, p.rho1), n1f), mulFF(mulFS(subSF(1.0, alpha1f), p.rho2), n2f));
    if (!nuf) {
        c.muf = num;
        return 0;
    }
    // /(alpha1f*rho1_ + (scalar(1) - alpha1f)*rho2_)                                      :199
    c.nuf = divFF(num, addFF(mulFS(alpha1f, p.rho1), mulFS(subSF(1.0, alpha1f), p.rho2)));
    return 0;
}

int mixtureCalcNu(Case& c, const ssam_mixture_params& p, int form) {
    if (int e = needNu(c)) return e;
    Fld& nu = fieldRef(c, SSAM_F_NU);
    if (form == 0) {
        Fld la = limitedAlpha(c.f[SSAM_F_ALPHA1]);
        nu = divFF(muFaithful(c, p), addFF(mulFS(la, p.rho1), mulFS(subSF(1.0, la), p.rho2)));
    } else {
        for (int32_t i = 0; i < c.nTot; ++i) {
            const double a = c.f[SSAM_F_ALPHA1][i];
            nu[i] = muCell(p, a, c.f[SSAM_F_NU1][i], c.f[SSAM_F_NU2][i]) / rhoMixCell(p, a);
        }
    }
    c.set[SSAM_F_NU] = true;
    return 0;
}

// k1()/k2()  :205-343
int unitsShift(const ssam_conductivity& k, double& shift, std::string& err) {
    shift = -273.0;  // celsConv_ (:226)
    if (k.units == SSAM_UNITS_KELVIN)
        shift = 0.0;
    else if (k.units != SSAM_UNITS_CELSIUS) {
        err = "Unknown conductivity function units. Choose Kelvin or Celsius.";
        return SSAM_ERR_UNITS;
    }
    return 0;
}
int kPhaseFaithful(const ssam_conductivity& k, const Fld& T, Fld& out, std::string& err) {
    if (k.kModel == SSAM_K_CUBIC) {
        double cc;
        if (int e = unitsShift(k, cc, err)) return e;
        Fld Tl = minFS(maxFS(addFS(T, cc), k.Tmin), k.Tmax);  // :244
        // :253  k0*(Tl/Tl) + k1*Tl + k2*pow(Tl,2) + k3*pow(Tl,3)
        out = addFF(addFF(addFF(mulSF(k.k0, divFF(Tl, Tl)), mulSF(k.k1, Tl)), mulSF(k.k2, powFS(Tl, 2))), mulSF(k.k3, powFS(Tl, 3)));
    } else {
        out = mulSF(k.k, divFF(T, T));  // :269
    }
    return 0;
}
inline double kPhaseCell(const ssam_conductivity& k, double shift, double T) {
    if (k.kModel == SSAM_K_CUBIC) {
        const double Tl = ofMin(ofMax(T + shift, k.Tmin), k.Tmax);
        return k.k0 * (Tl / Tl) + k.k1 * Tl + k.k2 * libmPow(Tl, 2) + k.k3 * libmPow(Tl, 3);
    }
    return k.k * (T / T);
}

int mixtureK(Case& c, const ssam_mixture_params& p, int form) {
    if (!c.set[SSAM_F_ALPHA1] || !c.set[SSAM_F_T]) {
        c.err = "alpha1 / temp not available";
        return SSAM_ERR_MISSING_FIELD;
    }
    Fld& k = fieldRef(c, SSAM_F_K);
    const Fld& T = c.f[SSAM_F_T];
    const Fld& a = c.f[SSAM_F_ALPHA1];
    double s1 = 0, s2 = 0;
    if (p.k1.kModel == SSAM_K_CUBIC)
        if (int e = unitsShift(p.k1, s1, c.err)) return e;
    if (p.k2.kModel == SSAM_K_CUBIC)
        if (int e = unitsShift(p.k2, s2, c.err)) return e;
    if (form == 0) {
        Fld la = limitedAlpha(a);  // :348-351
        Fld k1, k2;
        kPhaseFaithful(p.k1, T, k1, c.err);
        kPhaseFaithful(p.k2, T, k2, c.err);
        k = addFF(mulFF(la, k1), mulFF(subSF(1.0, la), k2));  // :359
    } else {
        for (int32_t i = 0; i < c.nTot; ++i) {
            const double la = la1(a[i]);
            k[i] = la * kPhaseCell(p.k1, s1, T[i]) + (1.0 - la) * kPhaseCell(p.k2, s2, T[i]);
        }
    }
    c.set[SSAM_F_K] = true;
    return 0;
}

int mixtureLinear(Case& c, int outField, double v1, double v2, int form) {  // cp() :378, rho() :397
    if (!c.set[SSAM_F_ALPHA1]) {
        c.err = "alpha1 not available";
        return SSAM_ERR_MISSING_FIELD;
    }
    Fld& out = fieldRef(c, outField);
    const Fld& a = c.f[SSAM_F_ALPHA1];
    if (form == 0) {
        Fld la = limitedAlpha(a);
        out = addFF(mulFS(la, v1), mulFS(subSF(1.0, la), v2));
    } else {
        for (int32_t i = 0; i < c.nTot; ++i) {
            const double la = la1(a[i]);
            out[i] = la * v1 + (1.0 - la) * v2;
        }
    }
    c.set[outField] = true;
    return 0;
}

// multiRegionVoF/alphaEqnSubCycle.H:41-42 with alpha2 = 1.0 - alpha1 (alphaEqn.H:144,217)
int solverRhoRhoCp(Case& c, const ssam_mixture_params& p, int form) {
    if (!c.set[SSAM_F_ALPHA1]) {
        c.err = "alpha1 not available";
        return SSAM_ERR_MISSING_FIELD;
    }
    const Fld& a1 = c.f[SSAM_F_ALPHA1];
    Fld& rho = fieldRef(c, SSAM_F_RHO_SOLVER);
    Fld& rcp = fieldRef(c, SSAM_F_RHOCP);
    if (form == 0) {
        Fld a2 = subSF(1.0, a1);
        rho = addFF(mulFS(a1, p.rho1), mulFS(a2, p.rho2));
        rcp = addFF(mulFS(mulFS(a1, p.rho1), p.cp1), mulFS(mulFS(a2, p.rho2), p.cp2));
    } else {
        for (int32_t i = 0; i < c.nTot; ++i) {
            const double a2 = 1.0 - a1[i];
            rho[i] = a1[i] * p.rho1 + a2 * p.rho2;
            rcp[i] = (a1[i] * p.rho1) * p.cp1 + (a2 * p.rho2) * p.cp2;
        }
    }
    c.set[SSAM_F_RHO_SOLVER] = c.set[SSAM_F_RHOCP] = true;
    return 0;
}

// ================================================================================================
// friction (frictionModels/stickModels/{constant,exponential}SlipStick/*.C)
// ================================================================================================
// r(): local area-weighted sums over the friction patch (constantSlipStick.C:118-131)
void patchSums(const Case& c, int patch, double out[4]) {
    double cx = 0, cy = 0, cz = 0, area = VSMALL;
    for (int32_t f = c.patchStart[patch]; f < c.patchStart[patch] + c.patchSize[patch]; ++f) {
        const int32_t b = f - c.nIF;
        const double ms = c.magSf_b[b];
        cx += c.Cf_b[3 * size_t(b)] * ms;
        cy += c.Cf_b[3 * size_t(b) + 1] * ms;
        cz += c.Cf_b[3 * size_t(b) + 2] * ms;
        area += ms;
    }
    out[0] = cx;
    out[1] = cy;
    out[2] = cz;
    out[3] = area;
}

// calcQvisc :58   3.0*betaTQ_*muEff_*pow(eDotEq_, 2)
int qvisc(Case& c, const ssam_stick_params& p, int form) {
    if (!c.set[SSAM_F_MUEFF] || !c.set[SSAM_F_EPSDOT]) {
        c.err = "lookupObject: muEff / epsilonDotEq not registered";
        return SSAM_ERR_MISSING_FIELD;
    }
    Fld& q = fieldRef(c, SSAM_F_QVISC);
    const Fld& mu = c.f[SSAM_F_MUEFF];
    const Fld& e = c.f[SSAM_F_EPSDOT];
    if (form == 0)
        q = mulFF(mulSF(3.0 * p.betaTQ, mu), powFS(e, 2));
    else
        for (int32_t i = 0; i < c.nTot; ++i) q[i] = ((3.0 * p.betaTQ) * mu[i]) * libmPow(e[i], 2);
    c.set[SSAM_F_QVISC] = true;
    return 0;
}

// calcQfric: constantSlipStick.C:62-107, exponentialSlipStick.C:62-108,150-165.
// sums = cross-rank totals of patchSums (reduce(), constantSlipStick.C:134-135) or NULL (serial).
int qfric(Case& c, const ssam_stick_params& p, const double* sums, int form) {
    if (p.fricPatch < 0 || p.fricPatch >= c.nPatches) {
        c.err = "fricPatch not found";
        return SSAM_ERR_INVALID;
    }
    if (!c.set[SSAM_F_SIGMAY]) {
        c.err = "lookupObject: sigmay not registered (NortonHoff registers none, SURVEY D-1)";
        return SSAM_ERR_MISSING_FIELD;
    }
    if (!c.set[SSAM_F_ALPHA1] || !c.set[SSAM_F_P]) {
        c.err = "lookupObject: alpha / p not registered";
        return SSAM_ERR_MISSING_FIELD;
    }
    double s[4];
    if (sums)
        std::memcpy(s, sums, sizeof(s));
    else
        patchSums(c, p.fricPatch, s);
    // :138  patchCenter_ = patchCenter_ / patchArea_
    const double ctr[3] = {s[0] / s[3], s[1] / s[3], s[2] / s[3]};
    std::memcpy(c.fricCentre, ctr, sizeof(ctr));
    c.fricArea = s[3];
    const size_t n = c.nTot;
    const Fld& alpha = c.f[SSAM_F_ALPHA1];
    const Fld& sy = c.f[SSAM_F_SIGMAY];
    const Fld& pr = c.f[SSAM_F_P];
    Fld& q = fieldRef(c, SSAM_F_QFRIC);
    const double magOmega = magVec(p.omega[0], p.omega[1], p.omega[2]);
    c.fricCells.assign(c.owner.begin() + c.patchStart[p.fricPatch],
                       c.owner.begin() + c.patchStart[p.fricPatch] + c.patchSize[p.fricPatch]);
    // :145 rad_ = mesh.C() - C_  (cell centres, and face centres on the boundary)
    auto radMag = [&](size_t i) {
        const double* x = (i < size_t(c.nCells)) ? &c.C[3 * i] : &c.Cf_b[3 * (i - c.nCells)];
        return magVec(x[0] - ctr[0], x[1] - ctr[1], x[2] - ctr[2]);
    };
    if (form == 0) {
        Fld magR(n);
        for (size_t i = 0; i < n; ++i) magR[i] = radMag(i);
        Fld fricCell = mulFS(alpha, 0.0);  // :79
        for (int32_t cc : c.fricCells) fricCell[cc] = 1.0;
        Fld inner;
        if (p.model == SSAM_STICK_CONSTANT) {
            // :100 (1-delta_)*betaTQ_*sigmay_/sqrt(3.0) + delta_*muf_*p_
            inner = addFF(divFS(mulSF((1.0 - p.delta) * p.betaTQ, sy), std::sqrt(3.0)), mulSF(p.delta * p.muf, pr));
        } else if (p.model == SSAM_STICK_EXPONENTIAL) {
            // delta(r) :156, muf(r) :164 -- the reference evaluates delta() three times
            auto delta = [&]() { return subSF(1.0, expF(divFS(mulSF(-magOmega, magR), (p.delta0 * p.omega0) * p.Rs))); };
            Fld muf = mulSF(p.mu0, expF(mulFF(mulFS(mulSF(-p.lambda, delta()), magOmega), magR)));
            inner = addFF(divFS(mulFF(mulFS(subSF(1.0, delta()), p.betaTQ), sy), std::sqrt(3.0)), mulFF(mulFF(delta(), muf), pr));
        } else {
            c.err = "Unknown stickModel type";
            return SSAM_ERR_UNKNOWN_MODEL;
        }
        q = mulFF(mulFF(mulFF(fricCell, alpha), inner), mulSF(magOmega, magR));
    } else {
        if (p.model != SSAM_STICK_CONSTANT && p.model != SSAM_STICK_EXPONENTIAL) {
            c.err = "Unknown stickModel type";
            return SSAM_ERR_UNKNOWN_MODEL;
        }
        std::vector<char> isFric(c.nCells, 0);
        for (int32_t cc : c.fricCells) isFric[cc] = 1;
        for (size_t i = 0; i < n; ++i) {
            const double fc = (i < size_t(c.nCells) && isFric[i]) ? 1.0 : alpha[i] * 0.0;
            const double mr = radMag(i);
            double inner;
            if (p.model == SSAM_STICK_CONSTANT) {
                inner = (((1.0 - p.delta) * p.betaTQ) * sy[i]) / std::sqrt(3.0) + (p.delta * p.muf) * pr[i];
            } else {
                const double d = 1.0 - std::exp(((-magOmega) * mr) / ((p.delta0 * p.omega0) * p.Rs));
                const double mf = p.mu0 * std::exp((((-p.lambda) * d) * magOmega) * mr);
                inner = (((1.0 - d) * p.betaTQ) * sy[i]) / std::sqrt(3.0) + (d * mf) * pr[i];
            }
            q[i] = ((fc * alpha[i]) * inner) * (magOmega * mr);
        }
    }
    c.set[SSAM_F_QFRIC] = true;
    return 0;
}

// boundaryConditions/*RotationalSlip*/…FvPatchVectorField.C updateCoeffs()
int bcRotSlip(Case& c, const ssam_rotslip_params& p) {
    if (p.patch < 0 || p.patch >= c.nPatches) {
        c.err = "patch out of range";
        return SSAM_ERR_INVALID;
    }
    Fld& U = fieldRef(c, SSAM_F_U);
    const double mo = magVec(p.omega[0], p.omega[1], p.omega[2]);
    const double oh[3] = {p.omega[0] / mo, p.omega[1] / mo, p.omega[2] / mo};  // omega_/mag(omega_)
    for (int32_t f = c.patchStart[p.patch]; f < c.patchStart[p.patch] + c.patchSize[p.patch]; ++f) {
        const int32_t b = f - c.nIF;
        const double* r = &c.Cf_b[3 * size_t(b)];
        const double dp = oh[0] * r[0] + oh[1] * r[1] + oh[2] * r[2];              // (omegaHat & r)
        const double d[3] = {r[0] - dp * oh[0], r[1] - dp * oh[1], r[2] - dp * oh[2]};  // r - (..)*omegaHat
        const double tv[3] = {p.omega[1] * d[2] - p.omega[2] * d[1], p.omega[2] * d[0] - p.omega[0] * d[2],
                              p.omega[0] * d[1] - p.omega[1] * d[0]};  // omega_ ^ d
        double dl;
        if (p.model == SSAM_ROTSLIP_CONSTANT)
            dl = p.delta;
        else {
            const double R = magVec(d[0], d[1], d[2]);
            dl = 1.0 - std::exp(((-mo) * R) / ((p.delta0 * p.omega0) * p.R0));  // …exponential…C:140
        }
        double* Ub = &U[size_t(c.nCells + b) * 3];
        for (int j = 0; j < 3; ++j) Ub[j] = tv[j] * (1 - dl) + p.Vt[j] * dl;  // tangVel*(1-delta) + Vt_*delta
    }
    return 0;
}

int strainRate(Case& c, double factor, int outField, int form) {
    if (!c.set[SSAM_F_U]) {
        c.err = "U not set";
        return SSAM_ERR_MISSING_FIELD;
    }
    Fld& G = fieldRef(c, SSAM_F_GRADU);
    if (form == 0)
        gradFaithful(c, c.f[SSAM_F_U], G);
    else
        gradFused(c, c.f[SSAM_F_U], G);
    c.set[SSAM_F_GRADU] = true;
    Fld& out = fieldRef(c, outField);
    strainFromGrad(c, G, factor, out, form);
    c.set[outField] = true;
    return 0;
}

// ================================================================================================
// interfaceProperties::correct() -> calculateK()   
// Called by the reference at immiscibleIncompressibleTwoPhaseThermalMixture.H:80; the class is OpenFOAM
// v1806 library code (src/transportModels/interfaceProperties/interfaceProperties.C), not part of the
// reference tree, so this is a restatement of its published algorithm:
//   const volVectorField gradAlpha(fvc::grad(alpha1_, "nHat"));          // Gauss linear
//   surfaceVectorField gradAlphaf(fvc::interpolate(gradAlpha));          // linear
//   surfaceVectorField nHatfv(gradAlphaf/(mag(gradAlphaf) + deltaN_));
//   correctContactAngle(...);                                            // no alphaContactAngle patches: no-op
//   nHatf_ = nHatfv & Sf;
//   K_ = -fvc::div(nHatf_);                                              // surfaceIntegrate, then boundary evaluate
// Stages, so that a multi-rank test can put the processor-patch evaluations in between:
//   1 cell gradient (+ extrapolated non-coupled boundary values)   [halo GRADALPHA]
//   2 gaussGrad::correctBoundaryConditions, face normals, nHatf, cell divergence, non-coupled K_b   [halo CURVATURE]
// ================================================================================================
int interfaceCorrect(Case& c, double deltaN, int stage) {
    if (!c.set[SSAM_F_ALPHA1]) {
        c.err = "alpha1 not set";
        return SSAM_ERR_MISSING_FIELD;
    }
    const Fld& al = c.f[SSAM_F_ALPHA1];
    Fld& G = fieldRef(c, SSAM_F_GRADALPHA);
    Fld& K = fieldRef(c, SSAM_F_CURVATURE);
    const int32_t nC = c.nCells;
    auto coupled = [&](int p) { return c.patchKind[p] == SSAM_PATCH_PROCESSOR; };
    if (stage & 1) {
        // gaussGrad<scalar>::gradf: face values, owner += Sf*ssf, neighbour -= Sf*ssf, boundary +=, /V
        Fld af(c.nFaces);
        for (int32_t f = 0; f < c.nIF; ++f) {
            const double aP = al[c.owner[f]], aN = al[c.neighbour[f]];
            af[f] = c.w[f] * (aP - aN) + aN;
        }
        for (int p = 0; p < c.nPatches; ++p)
            for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) {
                const double ab = al[nC + (f - c.nIF)];
                af[f] = coupled(p) ? c.w[f] * al[c.owner[f]] + (1.0 - c.w[f]) * ab : ab;
            }
        std::fill(G.begin(), G.begin() + size_t(nC) * 3, 0.0);
        for (int32_t f = 0; f < c.nIF; ++f)
            for (int i = 0; i < 3; ++i) {
                const double t = c.Sf[3 * size_t(f) + i] * af[f];
                G[size_t(c.owner[f]) * 3 + i] += t;
                G[size_t(c.neighbour[f]) * 3 + i] -= t;
            }
        for (int32_t f = c.nIF; f < c.nFaces; ++f)
            for (int i = 0; i < 3; ++i) G[size_t(c.owner[f]) * 3 + i] += c.Sf[3 * size_t(f) + i] * af[f];
        for (int32_t cc = 0; cc < nC; ++cc)
            for (int i = 0; i < 3; ++i) G[size_t(cc) * 3 + i] /= c.V[cc];
        // extrapolatedCalculated: patch value = adjacent cell value (processor patches: neighbour rank's, by halo)
        for (int p = 0; p < c.nPatches; ++p) {
            if (coupled(p)) continue;
            for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f)
                for (int i = 0; i < 3; ++i) G[size_t(nC + f - c.nIF) * 3 + i] = G[size_t(c.owner[f]) * 3 + i];
        }
        c.set[SSAM_F_GRADALPHA] = true;
    }
    if (stage & 2) {
        // gaussGrad::correctBoundaryConditions: g_b += n*(snGrad(alpha)_b - (n & g_b)) on non-coupled patches
        for (int p = 0; p < c.nPatches; ++p) {
            if (coupled(p)) continue;
            for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) {
                const int32_t b = f - c.nIF, P = c.owner[f];
                double* gb = &G[size_t(nC + b) * 3];
                const double ms = c.magSf_b[b];
                const double n[3] = {c.Sf[3 * size_t(f)] / ms, c.Sf[3 * size_t(f) + 1] / ms, c.Sf[3 * size_t(f) + 2] / ms};
                const double sn = c.deltaCoeffs_b[b] * (al[nC + b] - al[P]);  // fvPatchField::snGrad()
                const double ng = n[0] * gb[0] + n[1] * gb[1] + n[2] * gb[2];
                const double d = sn - ng;
                for (int i = 0; i < 3; ++i) gb[i] += n[i] * d;
            }
        }
        // fvc::interpolate(gradAlpha), nHatfv, nHatf_
        c.nHatf.assign(c.nFaces, 0.0);
        auto normalFlux = [&](const double* gf, int32_t f) {
            const double m = std::sqrt(gf[0] * gf[0] + gf[1] * gf[1] + gf[2] * gf[2]);
            const double den = m + deltaN;
            const double nv[3] = {gf[0] / den, gf[1] / den, gf[2] / den};
            const double* S = &c.Sf[3 * size_t(f)];
            return nv[0] * S[0] + nv[1] * S[1] + nv[2] * S[2];
        };
        for (int32_t f = 0; f < c.nIF; ++f) {
            const double* gP = &G[size_t(c.owner[f]) * 3];
            const double* gN = &G[size_t(c.neighbour[f]) * 3];
            double gf[3];
            for (int i = 0; i < 3; ++i) gf[i] = c.w[f] * (gP[i] - gN[i]) + gN[i];
            c.nHatf[f] = normalFlux(gf, f);
        }
        for (int p = 0; p < c.nPatches; ++p)
            for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) {
                const double* gb = &G[size_t(nC + f - c.nIF) * 3];
                double gf[3];
                if (coupled(p)) {
                    const double* gP = &G[size_t(c.owner[f]) * 3];
                    for (int i = 0; i < 3; ++i) gf[i] = c.w[f] * gP[i] + (1.0 - c.w[f]) * gb[i];
                } else {
                    for (int i = 0; i < 3; ++i) gf[i] = gb[i];
                }
                c.nHatf[f] = normalFlux(gf, f);
            }
        // fvc::div(nHatf_) = fvc::surfaceIntegrate: owner +=, neighbour -=, boundary +=, /V; K_ = -(...)
        Fld div(nC, 0.0);
        for (int32_t f = 0; f < c.nIF; ++f) {
            div[c.owner[f]] += c.nHatf[f];
            div[c.neighbour[f]] -= c.nHatf[f];
        }
        for (int32_t f = c.nIF; f < c.nFaces; ++f) div[c.owner[f]] += c.nHatf[f];
        for (int32_t cc = 0; cc < nC; ++cc) K[cc] = -(div[cc] / c.V[cc]);
        for (int p = 0; p < c.nPatches; ++p) {
            if (coupled(p)) continue;  // processor patches: neighbour rank's cell value, by halo
            for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) K[nC + f - c.nIF] = K[c.owner[f]];
        }
        c.set[SSAM_F_CURVATURE] = true;
    }
    return 0;
}

// ================================================================================================
// multiRegionVoF/alphaEqn.H:58-87 -- the interface-compression coefficient of the alpha equation:
//     surfaceScalarField phic(thermo.cAlpha()*mag(phi/mesh.magSf()));                                   :58
//     if (icAlpha > 0) { phic *= (1.0 - icAlpha);                                                       :63
//                        phic += (thermo.cAlpha()*icAlpha)*fvc::interpolate(mag(U)); }                  :64
//     if (scAlpha > 0) { phic += scAlpha*mag(mesh.delta() & fvc::interpolate(symm(fvc::grad(U)))); }    :70-71
//     non-coupled boundary patches: phic == 0                                                           :79-87
// OpenFOAM pieces restated (PARITY UNPINNED): magSf = mag(Sf) (+ VSMALL, fvMesh::makeMagSf);
// linear interpolation per component; symm(T) per SymmTensorI.H; Vector & SymmTensor =
// (v.x*xx + v.y*xy + v.z*xz, v.x*xy + v.y*yy + v.z*yz, v.x*xz + v.y*yz + v.z*zz); mesh.delta(): internal faces
// C[neighbour] - C[owner], processor faces (Cf - Cn) - (nbrCf - nbrCn) (processorFvPatch::delta()).
// Inputs: phi (face flux), U incl. boundary values, GRADU incl. boundary values (scAlpha > 0), the neighbour cell
// centres of the processor faces (c.procNbrC, orc_exchange_cell_centres).
// ================================================================================================
int alphaCompressionCoeff(Case& c, double cAlpha, double icAlpha, double scAlpha) {
    if (c.phi.size() != size_t(c.nFaces)) {
        c.err = "phi not set";
        return SSAM_ERR_MISSING_FIELD;
    }
    if ((icAlpha > 0 && !c.set[SSAM_F_U]) || (scAlpha > 0 && !c.set[SSAM_F_GRADU])) {
        c.err = "U / gradU not set";
        return SSAM_ERR_MISSING_FIELD;
    }
    const int32_t nC = c.nCells;
    auto magSfOf = [&](int32_t f) {
        const double* S = &c.Sf[3 * size_t(f)];
        return std::sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]) + VSMALL;
    };
    Fld phic(c.nFaces);
    for (int32_t f = 0; f < c.nFaces; ++f) phic[f] = cAlpha * std::fabs(c.phi[f] / magSfOf(f));
    auto coupled = [&](int p) { return c.patchKind[p] == SSAM_PATCH_PROCESSOR; };
    if (icAlpha > 0) {
        const Fld& U = c.f[SSAM_F_U];
        Fld magU(c.nTot);
        for (int32_t i = 0; i < c.nTot; ++i) magU[i] = magVec(U[3 * size_t(i)], U[3 * size_t(i) + 1], U[3 * size_t(i) + 2]);
        const Fld magUf = interpolateLinear(c, magU);
        const double s = cAlpha * icAlpha;
        for (int32_t f = 0; f < c.nFaces; ++f) {
            phic[f] *= (1.0 - icAlpha);
            phic[f] += s * magUf[f];
        }
    }
    if (scAlpha > 0) {
        const Fld& G = c.f[SSAM_F_GRADU];
        auto symmOf = [&](int32_t i, double* S6) {  // xx, xy, xz, yy, yz, zz
            const double* g = &G[9 * size_t(i)];
            S6[0] = g[0]; S6[1] = 0.5 * (g[1] + g[3]); S6[2] = 0.5 * (g[2] + g[6]);
            S6[3] = g[4]; S6[4] = 0.5 * (g[5] + g[7]); S6[5] = g[8];
        };
        auto term = [&](const double* d, const double* S6) {
            const double vx = d[0] * S6[0] + d[1] * S6[1] + d[2] * S6[2];
            const double vy = d[0] * S6[1] + d[1] * S6[3] + d[2] * S6[4];
            const double vz = d[0] * S6[2] + d[1] * S6[4] + d[2] * S6[5];
            return scAlpha * magVec(vx, vy, vz);
        };
        for (int32_t f = 0; f < c.nIF; ++f) {
            double sP[6], sN[6], sf[6], d[3];
            symmOf(c.owner[f], sP);
            symmOf(c.neighbour[f], sN);
            for (int k = 0; k < 6; ++k) sf[k] = c.w[f] * (sP[k] - sN[k]) + sN[k];
            for (int k = 0; k < 3; ++k) d[k] = c.C[3 * size_t(c.neighbour[f]) + k] - c.C[3 * size_t(c.owner[f]) + k];
            phic[f] += term(d, sf);
        }
        for (int p = 0; p < c.nPatches; ++p) {
            if (!coupled(p)) continue;  // the others are zeroed below
            for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) {
                const int32_t b = f - c.nIF, P = c.owner[f];
                double sP[6], sB[6], sf[6], d[3];
                symmOf(P, sP);
                symmOf(nC + b, sB);
                for (int k = 0; k < 6; ++k) sf[k] = c.w[f] * sP[k] + (1.0 - c.w[f]) * sB[k];
                for (int k = 0; k < 3; ++k) {
                    const double cf = c.Cf_b[3 * size_t(b) + k];
                    d[k] = (cf - c.C[3 * size_t(P) + k]) - (cf - (c.procNbrC.empty() ? 0.0 : c.procNbrC[3 * size_t(b) + k]));
                }
                phic[f] += term(d, sf);
            }
        }
    }
    for (int p = 0; p < c.nPatches; ++p) {
        if (coupled(p)) continue;
        for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) phic[f] = 0.0;
    }
    c.phic.swap(phic);
    return 0;
}

// ================================================================================================
// Explicit part of turbulence.divDevRhoReff(rho, U) (fluid/UEqn.H:13).  OpenFOAM v1806 library code
// (linearViscousStress<BasicTurbulenceModel>::divDevRhoReff), restated -- PARITY UNPINNED:
//     - fvc::div((alpha*rho*nuEff())*dev2(T(fvc::grad(U))))        [- fvm::laplacian(...) stays in OpenFOAM]
// alpha = 1, nuEff = nu (laminar), fvc::div(vf) = surfaceIntegrate(Sf & interpolate(vf)), linear interpolation.
// dev2(t) = t - twoThirdsI*tr(t); Vector & Tensor = (v.x*t.xx + v.y*t.yx + v.z*t.zx, ...).
// ================================================================================================
int divDevRhoReffExplicit(Case& c) {
    if (!c.set[SSAM_F_GRADU] || !c.set[SSAM_F_RHO_SOLVER] || !c.set[SSAM_F_NU]) {
        c.err = "gradU / rho / nu not set";
        return SSAM_ERR_MISSING_FIELD;
    }
    const Fld& G = c.f[SSAM_F_GRADU];
    const Fld coeff = mulFF(c.f[SSAM_F_RHO_SOLVER], c.f[SSAM_F_NU]);  // rho*nuEff
    Fld X(size_t(c.nTot) * 9);
    for (int32_t i = 0; i < c.nTot; ++i) {
        const double* g = &G[size_t(i) * 9];
        const double t[9] = {g[0], g[3], g[6], g[1], g[4], g[7], g[2], g[5], g[8]};  // T(gradU)
        const double ii = (2.0 / 3.0) * (t[0] + t[4] + t[8]);                        // twoThirdsI*tr(t)
        double d[9];
        for (int k = 0; k < 9; ++k) d[k] = t[k];
        d[0] = t[0] - ii; d[4] = t[4] - ii; d[8] = t[8] - ii;                         // dev2
        for (int k = 0; k < 9; ++k) X[size_t(i) * 9 + k] = coeff[i] * d[k];
    }
    // Sf & interpolate(X) per face, surfaceIntegrate, negate
    Fld& R = fieldRef(c, SSAM_F_DIVDEVRHOREFF);
    std::fill(R.begin(), R.end(), 0.0);
    auto flux = [&](const double* xf, int32_t f, double* out) {
        const double* S = &c.Sf[3 * size_t(f)];
        for (int j = 0; j < 3; ++j) out[j] = S[0] * xf[j] + S[1] * xf[3 + j] + S[2] * xf[6 + j];
    };
    Fld phi(size_t(c.nFaces) * 3);
    for (int32_t f = 0; f < c.nIF; ++f) {
        const double* xP = &X[size_t(c.owner[f]) * 9];
        const double* xN = &X[size_t(c.neighbour[f]) * 9];
        double xf[9];
        for (int k = 0; k < 9; ++k) xf[k] = c.w[f] * (xP[k] - xN[k]) + xN[k];
        flux(xf, f, &phi[size_t(f) * 3]);
    }
    for (int p = 0; p < c.nPatches; ++p)
        for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f) {
            const double* xb = &X[size_t(c.nCells + f - c.nIF) * 9];
            double xf[9];
            if (c.patchKind[p] == SSAM_PATCH_PROCESSOR) {
                const double* xP = &X[size_t(c.owner[f]) * 9];
                for (int k = 0; k < 9; ++k) xf[k] = c.w[f] * xP[k] + (1.0 - c.w[f]) * xb[k];
            } else {
                for (int k = 0; k < 9; ++k) xf[k] = xb[k];
            }
            flux(xf, f, &phi[size_t(f) * 3]);
        }
    for (int32_t f = 0; f < c.nIF; ++f)
        for (int j = 0; j < 3; ++j) {
            R[size_t(c.owner[f]) * 3 + j] += phi[size_t(f) * 3 + j];
            R[size_t(c.neighbour[f]) * 3 + j] -= phi[size_t(f) * 3 + j];
        }
    for (int32_t f = c.nIF; f < c.nFaces; ++f)
        for (int j = 0; j < 3; ++j) R[size_t(c.owner[f]) * 3 + j] += phi[size_t(f) * 3 + j];
    for (int32_t cc = 0; cc < c.nCells; ++cc)
        for (int j = 0; j < 3; ++j) R[size_t(cc) * 3 + j] = -(R[size_t(cc) * 3 + j] / c.V[cc]);
    for (int p = 0; p < c.nPatches; ++p) {
        if (c.patchKind[p] == SSAM_PATCH_PROCESSOR) continue;  // neighbour rank's cell value, by halo
        for (int32_t f = c.patchStart[p]; f < c.patchStart[p] + c.patchSize[p]; ++f)
            for (int j = 0; j < 3; ++j) R[size_t(c.nCells + f - c.nIF) * 3 + j] = R[size_t(c.owner[f]) * 3 + j];
    }
    c.set[SSAM_F_DIVDEVRHOREFF] = true;
    return 0;
}

// interfaceProperties constructor: deltaN_ = 1e-8/pow(average(alpha1.mesh().V()), 1.0/3.0)
double interfaceDeltaN(const std::vector<const Case*>& ranks) {
    double sum = 0;
    long long n = 0;
    for (const Case* c : ranks) {
        double s = 0;
        for (int32_t i = 0; i < c->nCells; ++i) s += c->V[i];
        sum += s;
        n += c->nCells;
    }
    return 1e-8 / libmPow(sum / double(n), 1.0 / 3.0);
}

// ================================================================================================
// fluid/calculateStrain.H:1-35,77-94 (the explicit stress / strain post-processing)
// ================================================================================================
// OpenFOAM tensor semantics restated: T() transpose, tr = xx+yy+zz, Tensor - SphericalTensor touches the
// diagonal only, A && B = sum_ij A_ij*B_ij in row-major order, dev(T) = T - (1/3)*tr(T)*I.
inline void stressStrainPoint(const ssam_stress_strain_params& p, const double* g, double eps, double T, double mu,
                              double prgh, double& Z, double& sf, double& dEq, double& svm) {
    Z = eps * std::exp(p.Q / (p.R * T));                                  // :12
    sf = p.sigmaR * std::asinh(libmPow(Z / p.A, 1 / p.n));                // :13
    double dE[9], eD[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double sym = 0.5 * (g[3 * i + j] + g[3 * j + i]);       // 0.5*(gradU + gradU.T())
            dE[3 * i + j] = sym * p.deltaT;                               // :18
            eD[3 * i + j] = sym;                                          // :79
        }
    const double third = 1.0 / 3.0;
    // :22  dDevEpsilon = dEpsilon - 1.0/3.0*tr(dEpsilon)*I
    const double s1 = third * (dE[0] + dE[4] + dE[8]);
    double dD[9];
    for (int k = 0; k < 9; ++k) dD[k] = dE[k];
    dD[0] = dE[0] - s1; dD[4] = dE[4] - s1; dD[8] = dE[8] - s1;
    // :26  J2s = 0.5*(pow(tr(dDevEpsilon),2) - (dDevEpsilon && dDevEpsilon))
    double dd = dD[0] * dD[0];
    for (int k = 1; k < 9; ++k) dd += dD[k] * dD[k];
    const double J2s = 0.5 * (libmPow(dD[0] + dD[4] + dD[8], 2) - dd);
    const double J2 = std::sqrt(libmPow(J2s, 2));                         // :29
    dEq = std::sqrt((4.0 / 3.0) * J2);                                    // :35
    // :82-92  dev(epsilonDot), shear stress, total stress, its deviator, von Mises
    const double s2 = third * (eD[0] + eD[4] + eD[8]);
    double sh[9];
    for (int k = 0; k < 9; ++k) sh[k] = eD[k];
    sh[0] = eD[0] - s2; sh[4] = eD[4] - s2; sh[8] = eD[8] - s2;
    const double twoMu = 2.0 * mu;
    for (int k = 0; k < 9; ++k) sh[k] = twoMu * sh[k];                    // :85
    double st[9];
    for (int k = 0; k < 9; ++k) st[k] = sh[k];
    st[0] = prgh + sh[0]; st[4] = prgh + sh[4]; st[8] = prgh + sh[8];     // :88  p_rgh*I + shearStress
    const double s3 = third * (st[0] + st[4] + st[8]);
    st[0] -= s3; st[4] -= s3; st[8] -= s3;                                // :91
    double ss = st[0] * st[0];
    for (int k = 1; k < 9; ++k) ss += st[k] * st[k];
    svm = std::sqrt(2.3 / 3.0) * std::sqrt(ss);                           // :94  (2.3/3.0 as written)
}

int stressStrain(Case& c, const ssam_stress_strain_params& p, int form) {
    const int needF[] = {SSAM_F_U, SSAM_F_EPSDOT, SSAM_F_T, SSAM_F_MUEFF, SSAM_F_PRGH};
    for (int f : needF)
        if (!c.set[f]) {
            c.err = "calculateStrain: U / epsilonDotEq / temp / muEff / p_rgh not available";
            return SSAM_ERR_MISSING_FIELD;
        }
    Fld& G = fieldRef(c, SSAM_F_GRADU);
    if (form == 0)
        gradFaithful(c, c.f[SSAM_F_U], G);
    else
        gradFused(c, c.f[SSAM_F_U], G);
    c.set[SSAM_F_GRADU] = true;
    const size_t n = c.nTot;
    const Fld &eps = c.f[SSAM_F_EPSDOT], &T = c.f[SSAM_F_T], &mu = c.f[SSAM_F_MUEFF], &pr = c.f[SSAM_F_PRGH];
    Fld &Z = fieldRef(c, SSAM_F_Z), &sf = fieldRef(c, SSAM_F_SIGMAF_PP), &dEq = fieldRef(c, SSAM_F_DEPSEQ),
        &eEq = fieldRef(c, SSAM_F_EPSEQ), &svm = fieldRef(c, SSAM_F_SIGMAVM);
    if (form == 1) {
        for (size_t i = 0; i < n; ++i) stressStrainPoint(p, &G[9 * i], eps[i], T[i], mu[i], pr[i], Z[i], sf[i], dEq[i], svm[i]);
    } else {
        // field-at-a-time with tensor temporaries, one pass per operator as calculateStrain.H writes them
        Z = mulFF(eps, expF(divSF(p.Q, mulSF(p.R, T))));
        sf = mulSF(p.sigmaR, asinhF(powFS(divFS(Z, p.A), 1 / p.n)));
        auto tens = [&]() { return Fld(9 * n); };
        Fld GT = tens(), sum = tens(), sym = tens(), dE = tens();
        for (size_t i = 0; i < n; ++i)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) GT[9 * i + 3 * a + b] = G[9 * i + 3 * b + a];
        LOOP(9 * n) sum[i] = G[i] + GT[i];
        LOOP(9 * n) sym[i] = 0.5 * sum[i];
        LOOP(9 * n) dE[i] = sym[i] * p.deltaT;
        auto trF = [&](const Fld& t) { Fld r(n); LOOP(n) r[i] = t[9 * i] + t[9 * i + 4] + t[9 * i + 8]; return r; };
        auto minusSph = [&](const Fld& t, const Fld& sph) {
            Fld r = t;
            LOOP(n) { r[9 * i] = t[9 * i] - sph[i]; r[9 * i + 4] = t[9 * i + 4] - sph[i]; r[9 * i + 8] = t[9 * i + 8] - sph[i]; }
            return r;
        };
        auto ddot = [&](const Fld& a, const Fld& b) {
            Fld r(n);
            LOOP(n) { double s = a[9 * i] * b[9 * i]; for (int k = 1; k < 9; ++k) s += a[9 * i + k] * b[9 * i + k]; r[i] = s; }
            return r;
        };
        Fld dD = minusSph(dE, mulSF(1.0 / 3.0, trF(dE)));
        Fld dd = ddot(dD, dD);
        Fld t2 = powFS(trF(dD), 2);
        Fld diff(n);
        LOOP(n) diff[i] = t2[i] - dd[i];
        Fld J2s = mulSF(0.5, diff);
        Fld J2p = powFS(J2s, 2), J2(n);
        LOOP(n) J2[i] = std::sqrt(J2p[i]);
        Fld a43 = mulSF(4.0 / 3.0, J2);
        LOOP(n) dEq[i] = std::sqrt(a43[i]);
        Fld eDev = minusSph(sym, mulSF(1.0 / 3.0, trF(sym)));
        Fld twoMu = mulSF(2.0, mu), sh = tens();
        LOOP(n) for (int k = 0; k < 9; ++k) sh[9 * i + k] = twoMu[i] * eDev[9 * i + k];
        Fld st = sh;
        LOOP(n) { st[9 * i] = pr[i] + sh[9 * i]; st[9 * i + 4] = pr[i] + sh[9 * i + 4]; st[9 * i + 8] = pr[i] + sh[9 * i + 8]; }
        Fld sdev = minusSph(st, mulSF(1.0 / 3.0, trF(st)));
        Fld ss = ddot(sdev, sdev), rt(n);
        LOOP(n) rt[i] = std::sqrt(ss[i]);
        svm = mulSF(std::sqrt(2.3 / 3.0), rt);
    }
    if (!c.set[SSAM_F_EPSEQ]) std::fill(eEq.begin(), eEq.end(), 0.0);  // createStressStrain.H: initial value zero
    for (size_t i = 0; i < n; ++i) eEq[i] += dEq[i];                     // :40  epsilonEq += dEpsilonEq
    for (int f : {SSAM_F_Z, SSAM_F_SIGMAF_PP, SSAM_F_DEPSEQ, SSAM_F_EPSEQ, SSAM_F_SIGMAVM}) c.set[f] = true;
    return 0;
}

// the fused update = the staged calls in the order documented in include/ssam_gpu.h
int update(Case& c, const ssam_update_params& p, const double* sums, int form) {
    int e;
    if ((e = strainRate(c, std::sqrt(2.0 / 3.0), SSAM_F_EPSDOT, form))) return e;
    if ((e = viscosityCorrect(c, 1, p.visc1, form))) return e;
    if ((e = viscosityCorrect(c, 2, p.visc2, form))) return e;
    if ((e = mixtureCalcNu(c, p.mix, form))) return e;
    if ((e = mixtureMu(c, p.mix, form))) return e;
    if ((e = qvisc(c, p.stick, form))) return e;
    if (p.stick.fricPatch >= 0)
        if ((e = qfric(c, p.stick, sums, form))) return e;
    if ((e = mixtureLinear(c, SSAM_F_CP, p.mix.cp1, p.mix.cp2, form))) return e;
    if ((e = mixtureK(c, p.mix, form))) return e;
    if ((e = mixtureLinear(c, SSAM_F_RHO, p.mix.rho1, p.mix.rho2, form))) return e;
    return 0;
}

}  // namespace

extern "C" {

typedef struct orc_case orc_case;
#define CASE(h) (*reinterpret_cast<Case*>(h))

orc_case* orc_create(const ssam_mesh_desc* m, const int32_t* patchUbc) {
    Case* c = new Case;
    c->nCells = m->nCells;
    c->nIF = m->nInternalFaces;
    c->nFaces = m->nFaces;
    c->nBnd = m->nFaces - m->nInternalFaces;
    c->nPatches = m->nPatches;
    c->nTot = c->nCells + c->nBnd;
    c->owner.assign(m->owner, m->owner + m->nFaces);
    c->neighbour.assign(m->neighbour, m->neighbour + m->nInternalFaces);
    c->patchStart.assign(m->patchStart, m->patchStart + m->nPatches);
    c->patchSize.assign(m->patchSize, m->patchSize + m->nPatches);
    c->patchKind.assign(m->patchKind, m->patchKind + m->nPatches);
    c->patchNbrRank.assign(m->patchNbrRank, m->patchNbrRank + m->nPatches);
    if (patchUbc)
        c->patchUbc.assign(patchUbc, patchUbc + m->nPatches);
    else
        c->patchUbc.assign(m->nPatches, SSAM_BC_VALUE);
    c->Sf.assign(m->Sf, m->Sf + 3 * size_t(m->nFaces));
    c->w.assign(m->weights, m->weights + m->nFaces);
    c->V.assign(m->V, m->V + m->nCells);
    c->C.assign(m->C, m->C + 3 * size_t(m->nCells));
    c->Cf_b.assign(m->Cf_b, m->Cf_b + 3 * size_t(c->nBnd));
    c->magSf_b.assign(m->magSf_b, m->magSf_b + c->nBnd);
    c->deltaCoeffs_b.assign(m->deltaCoeffs_b, m->deltaCoeffs_b + c->nBnd);
    return reinterpret_cast<orc_case*>(c);
}
void orc_destroy(orc_case* h) { delete reinterpret_cast<Case*>(h); }
const char* orc_last_error(orc_case* h) { return CASE(h).err.c_str(); }

// pointer to the AoS storage [nCells+nBnd][ncomp]; marks the field as set when `mark` != 0
double* orc_field(orc_case* h, int field, int mark) {
    Case& c = CASE(h);
    Fld& v = fieldRef(c, field);
    if (mark) c.set[field] = true;
    return v.data();
}
int orc_ntot(orc_case* h) { return CASE(h).nTot; }

int orc_bc_rotational_slip(orc_case* h, const ssam_rotslip_params* p) { return bcRotSlip(CASE(h), *p); }
int orc_grad_u(orc_case* h, int form) {
    Case& c = CASE(h);
    if (!c.set[SSAM_F_U]) return SSAM_ERR_MISSING_FIELD;
    Fld& G = fieldRef(c, SSAM_F_GRADU);
    if (form == 0)
        gradFaithful(c, c.f[SSAM_F_U], G);
    else
        gradFused(c, c.f[SSAM_F_U], G);
    c.set[SSAM_F_GRADU] = true;
    return 0;
}
int orc_strain_rate(orc_case* h, double factor, int outField, int form) { return strainRate(CASE(h), factor, outField, form); }
// boundary part only: after the processor-patch values of GRADU were exchanged
int orc_strain_from_grad(orc_case* h, double factor, int outField, int form) {
    Case& c = CASE(h);
    Fld& out = fieldRef(c, outField);
    strainFromGrad(c, c.f[SSAM_F_GRADU], factor, out, form);
    c.set[outField] = true;
    return 0;
}
int orc_viscosity_correct(orc_case* h, int phase, const ssam_viscosity_params* p, int form) {
    return viscosityCorrect(CASE(h), phase, *p, form);
}
int orc_mixture_calc_nu(orc_case* h, const ssam_mixture_params* p, int form) { return mixtureCalcNu(CASE(h), *p, form); }
int orc_mixture_mu(orc_case* h, const ssam_mixture_params* p, int form) { return mixtureMu(CASE(h), *p, form); }
int orc_mixture_rho(orc_case* h, const ssam_mixture_params* p, int form) {
    return mixtureLinear(CASE(h), SSAM_F_RHO, p->rho1, p->rho2, form);
}
int orc_mixture_cp(orc_case* h, const ssam_mixture_params* p, int form) {
    return mixtureLinear(CASE(h), SSAM_F_CP, p->cp1, p->cp2, form);
}
int orc_mixture_k(orc_case* h, const ssam_mixture_params* p, int form) { return mixtureK(CASE(h), *p, form); }
int orc_solver_rho_rhocp(orc_case* h, const ssam_mixture_params* p, int form) { return solverRhoRhoCp(CASE(h), *p, form); }
int orc_stress_strain(orc_case* h, const ssam_stress_strain_params* p, int form) { return stressStrain(CASE(h), *p, form); }
void orc_friction_patch_sums(orc_case* h, int patch, double* out4) { patchSums(CASE(h), patch, out4); }
int orc_friction_correct(orc_case* h, const ssam_stick_params* p, const double* sums, int form) {
    Case& c = CASE(h);
    if (int e = qvisc(c, *p, form)) return e;
    if (p->fricPatch < 0) return 0;
    return qfric(c, *p, sums, form);
}
void orc_friction_patch_centre(orc_case* h, double* centre3, double* area) {
    std::memcpy(centre3, CASE(h).fricCentre, 3 * sizeof(double));
    *area = CASE(h).fricArea;
}
// fvc::interpolate(field) (linear) into out[nFaces]
int orc_face_interpolate(orc_case* h, int field, double* out) {
    Case& c = CASE(h);
    if (field < 0 || field >= SSAM_F_COUNT || ncompOf(field) != 1 || !c.set[field]) {
        c.err = "scalar field not set";
        return SSAM_ERR_MISSING_FIELD;
    }
    const Fld sf = interpolateLinear(c, c.f[field]);
    std::memcpy(out, sf.data(), sf.size() * sizeof(double));
    return 0;
}
int orc_mixture_face_visc(orc_case* h, const ssam_mixture_params* p, int nuf) { return mixtureFaceVisc(CASE(h), *p, nuf != 0); }
double* orc_face_field(orc_case* h, int which) {
    Case& c = CASE(h);
    return which == SSAM_FF_NHATF ? c.nHatf.data() : (which == SSAM_FF_MUF ? c.muf.data() : c.nuf.data());
}
// min(vf).value(), max(vf).value() (fluid/TEqn.H:42-43): over internal and boundary values
void orc_field_min_max(orc_case** cases, int nRanks, int field, double* mn, double* mx) {
    double lo = 0, hi = 0;
    bool first = true;
    for (int r = 0; r < nRanks; ++r) {
        const Case& c = CASE(cases[r]);
        for (double v : c.f[field]) {
            if (first) { lo = hi = v; first = false; }
            lo = ofMin(lo, v);
            hi = ofMax(hi, v);
        }
    }
    *mn = lo;
    *mx = hi;
}
int orc_div_dev_rho_reff_explicit(orc_case* h) { return divDevRhoReffExplicit(CASE(h)); }
int orc_set_phi(orc_case* h, const double* phi) {
    Case& c = CASE(h);
    c.phi.assign(phi, phi + c.nFaces);
    return 0;
}
int orc_alpha_compression_coeff(orc_case* h, double cAlpha, double icAlpha, double scAlpha) {
    return alphaCompressionCoeff(CASE(h), cAlpha, icAlpha, scAlpha);
}
double* orc_phic(orc_case* h) { return CASE(h).phic.data(); }
// processorPolyPatch::neighbFaceCellCentres(): the cell centre behind each processor face on the neighbour rank
int orc_exchange_cell_centres(orc_case** cases, int nRanks) {
    for (int r = 0; r < nRanks; ++r) {
        Case& me = CASE(cases[r]);
        me.procNbrC.assign(size_t(me.nBnd) * 3, 0.0);
        for (int p = 0; p < me.nPatches; ++p) {
            if (me.patchKind[p] != SSAM_PATCH_PROCESSOR) continue;
            const int q = me.patchNbrRank[p];
            if (q < 0 || q >= nRanks) return SSAM_ERR_COMM;
            Case& nb = CASE(cases[q]);
            int pq = -1;
            for (int pp = 0; pp < nb.nPatches; ++pp)
                if (nb.patchKind[pp] == SSAM_PATCH_PROCESSOR && nb.patchNbrRank[pp] == r) pq = pp;
            if (pq < 0 || nb.patchSize[pq] != me.patchSize[p]) return SSAM_ERR_COMM;
            for (int32_t i = 0; i < me.patchSize[p]; ++i) {
                const int32_t b = me.patchStart[p] - me.nIF + i, nbrCell = nb.owner[nb.patchStart[pq] + i];
                for (int k = 0; k < 3; ++k) me.procNbrC[size_t(b) * 3 + k] = nb.C[size_t(nbrCell) * 3 + k];
            }
        }
    }
    return 0;
}
int orc_interface_correct(orc_case* h, double deltaN, int stage) { return interfaceCorrect(CASE(h), deltaN, stage); }
double* orc_interface_nhatf(orc_case* h) { return CASE(h).nHatf.data(); }
double orc_interface_delta_n(orc_case** cases, int nRanks) {
    std::vector<const Case*> v;
    for (int r = 0; r < nRanks; ++r) v.push_back(&CASE(cases[r]));
    return interfaceDeltaN(v);
}
int orc_update(orc_case* h, const ssam_update_params* p, const double* sums, int form) {
    return update(CASE(h), *p, sums, form);
}

// integer addressing
int64_t orc_addressing_size(orc_case* h, int which) {
    Case& c = CASE(h);
    buildAddressing(c);
    switch (which) {
        case SSAM_ADDR_OWNER_START: return c.ownerStart.size();
        case SSAM_ADDR_EXTRA_START: return c.extraStart.size();
        case SSAM_ADDR_EXTRA_FACE: return c.extraFace.size();
        case SSAM_ADDR_EXTRA_OTHER: return c.extraOther.size();
        case SSAM_ADDR_CELL_FACES_START: return c.cellFacesStart.size();
        case SSAM_ADDR_CELL_FACES: return c.cellFaces.size();
        case SSAM_ADDR_HALO_SEND_CELLS: return c.haloSend.size();
        case SSAM_ADDR_HALO_PATCH_OFFSETS: return c.haloOffsets.size();
        case SSAM_ADDR_FRIC_CELLS: return c.fricCells.size();
        case SSAM_ADDR_FACE_COLOUR: return c.faceColour.size();
    }
    return -1;
}
int orc_addressing_get(orc_case* h, int which, int32_t* out) {
    Case& c = CASE(h);
    buildAddressing(c);
    const std::vector<int32_t>* v = nullptr;
    switch (which) {
        case SSAM_ADDR_OWNER_START: v = &c.ownerStart; break;
        case SSAM_ADDR_EXTRA_START: v = &c.extraStart; break;
        case SSAM_ADDR_EXTRA_FACE: v = &c.extraFace; break;
        case SSAM_ADDR_EXTRA_OTHER: v = &c.extraOther; break;
        case SSAM_ADDR_CELL_FACES_START: v = &c.cellFacesStart; break;
        case SSAM_ADDR_CELL_FACES: v = &c.cellFaces; break;
        case SSAM_ADDR_HALO_SEND_CELLS: v = &c.haloSend; break;
        case SSAM_ADDR_HALO_PATCH_OFFSETS: v = &c.haloOffsets; break;
        case SSAM_ADDR_FRIC_CELLS: v = &c.fricCells; break;
        case SSAM_ADDR_FACE_COLOUR: v = &c.faceColour; break;
        default: return SSAM_ERR_INVALID;
    }
    if (!v->empty()) std::memcpy(out, v->data(), v->size() * sizeof(int32_t));
    return 0;
}

// ---- multi-rank emulation (one case per rank, in-memory halo copies) ------------------------------
// processorFvPatchField::evaluate: boundary values on processor patches <- neighbour's faceCells values
int orc_halo_exchange(orc_case** cases, int nRanks, int field) {
    const int nc = ncompOf(field);
    for (int r = 0; r < nRanks; ++r) {
        Case& me = CASE(cases[r]);
        Fld& mine = fieldRef(me, field);
        for (int p = 0; p < me.nPatches; ++p) {
            if (me.patchKind[p] != SSAM_PATCH_PROCESSOR) continue;
            const int q = me.patchNbrRank[p];
            Case& nb = CASE(cases[q]);
            int pq = -1;
            for (int pp = 0; pp < nb.nPatches; ++pp)
                if (nb.patchKind[pp] == SSAM_PATCH_PROCESSOR && nb.patchNbrRank[pp] == r) pq = pp;
            if (pq < 0 || nb.patchSize[pq] != me.patchSize[p]) return SSAM_ERR_COMM;
            const Fld& theirs = fieldRef(nb, field);
            for (int32_t i = 0; i < me.patchSize[p]; ++i) {
                const int32_t b = me.patchStart[p] - me.nIF + i;
                const int32_t nbrCell = nb.owner[nb.patchStart[pq] + i];
                for (int k = 0; k < nc; ++k) mine[size_t(me.nCells + b) * nc + k] = theirs[size_t(nbrCell) * nc + k];
            }
        }
    }
    return 0;
}

// Timed CPU baseline: every rank's faithful (or fused) update on its own thread, halo copies of U
// before and of gradU after the gradient (what gaussGrad's correctBoundaryConditions() exchanges).
// Returns 0 on success.  Threads = nRanks (one mesh partition per worker, like one MPI rank per core).
int orc_update_multi(orc_case** cases, int nRanks, const ssam_update_params* p, int form) {
    // patch-centre reduce() (constantSlipStick.C:134-135): sum of per-rank partial sums in rank order
    double sums[4] = {0, 0, 0, 0};
    bool haveFric = p->stick.fricPatch >= 0;
    if (haveFric) {
        for (int r = 0; r < nRanks; ++r) {
            double s[4];
            patchSums(CASE(cases[r]), p->stick.fricPatch, s);
            for (int k = 0; k < 4; ++k) sums[k] = (r == 0) ? s[k] : sums[k] + s[k];
        }
    }
    if (int e = orc_halo_exchange(cases, nRanks, SSAM_F_U)) return e;
    std::vector<int> rc(nRanks, 0);
    auto phase1 = [&](int r) {
        Case& c = CASE(cases[r]);
        Fld& G = fieldRef(c, SSAM_F_GRADU);
        if (form == 0)
            gradFaithful(c, c.f[SSAM_F_U], G);
        else
            gradFused(c, c.f[SSAM_F_U], G);
        c.set[SSAM_F_GRADU] = true;
    };
    auto phase2 = [&](int r) {
        Case& c = CASE(cases[r]);
        int e = 0;
        Fld& eps = fieldRef(c, SSAM_F_EPSDOT);
        strainFromGrad(c, c.f[SSAM_F_GRADU], std::sqrt(2.0 / 3.0), eps, form);
        c.set[SSAM_F_EPSDOT] = true;
        const ssam_update_params& q = *p;
        // NortonHoff recomputes its own gradient inside viscosityCorrect (needs U halos only)
        if (!e) e = viscosityCorrect(c, 1, q.visc1, form);
        if (!e) e = viscosityCorrect(c, 2, q.visc2, form);
        if (!e) e = mixtureCalcNu(c, q.mix, form);
        if (!e) e = mixtureMu(c, q.mix, form);
        if (!e) e = qvisc(c, q.stick, form);
        if (!e && haveFric) e = qfric(c, q.stick, sums, form);
        if (!e) e = mixtureLinear(c, SSAM_F_CP, q.mix.cp1, q.mix.cp2, form);
        if (!e) e = mixtureK(c, q.mix, form);
        if (!e) e = mixtureLinear(c, SSAM_F_RHO, q.mix.rho1, q.mix.rho2, form);
        rc[r] = e;
    };
    auto runAll = [&](auto fn) {
        if (nRanks == 1) {
            fn(0);
            return;
        }
        std::vector<std::thread> th;
        for (int r = 0; r < nRanks; ++r) th.emplace_back(fn, r);
        for (auto& t : th) t.join();
    };
    runAll(phase1);
    if (int e = orc_halo_exchange(cases, nRanks, SSAM_F_GRADU)) return e;
    runAll(phase2);
    for (int r = 0; r < nRanks; ++r)
        if (rc[r]) return rc[r];
    return 0;
}

}  // extern "C"
