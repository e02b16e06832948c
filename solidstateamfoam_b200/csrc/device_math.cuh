// Per-element arithmetic of the hot path, written once and used by the staged kernels, the fused
// kernel and the boundary-face kernel, so "fused == staged" holds by construction.
//
// Every expression keeps the reference's operator order (C++ left-to-right; one rounded FP64
// operation per operator).  The translation unit is compiled with -fmad=false, so nvcc never
// contracts a*b+c into an FMA here -- OpenFOAM's stock gcc build has none.
// Division and sqrt are IEEE round-to-nearest in CUDA FP64, so everything except exp/log/pow/
// asinh is bit-identical to the host.
#pragma once
#include <cstdint>

#include "fastmath.cuh"
#include "ssam_gpu.h"

namespace ssam {

constexpr double kSMALL = 1e-15;    // Foam::SMALL
constexpr double kVSMALL = 1e-300;  // Foam::VSMALL

// Foam::max / Foam::min: (a > b) ? a : b and (a < b) ? a : b, argument order as written
// (fmax/fmin differ for NaN)
SSAM_DEV double ofMax(double a, double b) { return (a > b) ? a : b; }
SSAM_DEV double ofMin(double a, double b) { return (a < b) ? a : b; }

// x/x evaluated without a divide: exactly 1 for finite non-zero x, NaN for 0, +-inf, NaN --
// bit-identical to the IEEE quotient the reference computes in k0*(T/T) and k*(T/T)
// (incompressibleTwoPhaseThermalMixture.C:253,269).
SSAM_DEV double selfRatio(double x) {
    const unsigned h = unsigned(__double2hiint(x)) & 0x7fffffffu;  // integer pipe, not FP64
    const bool ok = (h < 0x7ff00000u) && ((h | unsigned(__double2loint(x))) != 0u);
    return ok ? 1.0 : __longlong_as_double(0x7ff8000000000000LL);
}

// ---- tensor algebra (OpenFOAM SymmTensorI.H) -------------------------------------
// mag(symm(T)) with T row-major xx,xy,xz,yx,yy,yz,zx,zy,zz
SSAM_DEV double magSymm(const double (&g)[9]) {
    const double xx = g[0], xy = 0.5 * (g[1] + g[3]), xz = 0.5 * (g[2] + g[6]);
    const double yy = g[4], yz = 0.5 * (g[5] + g[7]), zz = g[8];
    const double ms = xx * xx + 2 * (xy * xy) + 2 * (xz * xz) + yy * yy + 2 * (yz * yz) + zz * zz;
    return sqrt(ms);
}
SSAM_DEV double magVec(double x, double y, double z) { return sqrt(x * x + y * y + z * z); }

// ---- viscosity models --------------------------------------------------------------------------------
struct ViscOut {
    double nu, sigmaf, sigmay;
};

// Parameter-only sub-expressions.  The reference evaluates these dimensionedScalar expressions on the
// host (libm) once per call; so do we (deriveVisc() in ssam_abi.cu), instead of once per cell.
struct ViscDerived {
    double invN;        // 1/n_                                   SheppardWright.C:74
    double threeRho;    // 3.0*rho_                               SheppardWright.C:115, FKS.C:178
    double logA;        // log(A_)          (log-domain form of pow(Z/A, 1/n))
    double log01;       // log(0.1)         (e01 in sigmay(), SheppardWright.C:85)
    double Hcoef;       // a1_ + (a2_*1.571)                      FKS.C:73
    double TmTa;        // Tm_ - Ta_                              FKS.C:112
    double negC1;       // -c1_                                   FKS.C:99
    double invC2;       // 1/c2_                                  FKS.C:99
    double b3LogY;      // b3_*log(e01_/e0_ + VSMALL)             FKS.C:152
    double mu0OverRho;  // mu0_/rho_                              NortonHoff.C:60
    double mMinus1;     // m_ - 1                                 NortonHoff.C:63
    double sqrt3;       // sqrt(3.0)                              NortonHoff.C:62
    double invE0;       // 1/e0_
    int b2IsOne;        // pow(x, 1.0) == x exactly
    int invC2IsOne;
};

// SheppardWright.C:51-118; correct() SheppardWright.H:133-138.
// pow(Z/A, 1/n) with Z = eDot*exp(Q/(R T)) is evaluated in the log domain,
//   log(Z/A) = log(eDot) + Q/(R T) - log(A),
// which needs neither exp(Q/RT) nor a pow(): abs. error of the exponent ~6e-15 -> relative ~2e-15
// after the 1/n, the same size as the reference's own rounding of exp(58).
SSAM_DEV double swAsinhOfLog(double t) {
    // asinh(exp(t)): for exp(t) >= 2^28, asinh(P) = log(2P) = t + ln2 to double precision
    if (t > 19.5) return t + kLn2;
    return fastAsinh(fastExp(t));
}
SSAM_DEV ViscOut sheppardWright(const ssam_viscosity_params& p, const ViscDerived& d, double T, double e) {
    ViscOut o;
    const double q = divFast(p.Q, p.R * T);  // Q_/(R_*temp_)  :64
    const double qa = q - d.logA;
    o.sigmaf = p.sigmaR * swAsinhOfLog((fastLog(e) + qa) * d.invN);   // :74
    o.sigmay = p.sigmaR * swAsinhOfLog((d.log01 + qa) * d.invN);      // :91
    o.nu = ofMax(p.nuMin, ofMin(p.nuMax, divFast(o.sigmaf, d.threeRho * ofMax(e, p.eDotMin))));  // :109-117
    return o;
}

// FKS.C:63-183; correct() FKS.H:156-161
SSAM_DEV ViscOut fks(const ssam_viscosity_params& p, const ViscDerived& d, double T, double e) {
    ViscOut o;
    const double Tlim = ofMin(ofMax(T, p.Ta), p.Tm);                      // :108
    // :112  true quotient: Th feeds exp(-c1*Th) under the 1 - 1/z cancellation below, where a 1-ulp difference in
    // Th is amplified by c1*Th/Theta (DESIGN.md 7 "conditioning")
    const double Th = divExact(Tlim - p.Ta, d.TmTa);
    const double H = d.Hcoef * selfRatio(T);                              // :73  (X*T)/T: X to 1 ulp, NaN at T = 0
    const double base = Th + kSMALL;
    const double pw = p.b1 * (d.b2IsOne ? base : fastPow(base, p.b2));    // :89 / :152
    const double Lam = 1.0 + pw * (p.b3 * fastLog(e * d.invE0 + kVSMALL));  // :89
    const double z = 1.0 + fastExp(d.negC1 * Th);                         // in (1, 2]
    // :99  1.0 - 1.0/pow(z, 1/c2) cancels as exp(-c1*Th) -> 0 (T -> Tm, large c1): the quotient is the IEEE one, so
    // that with c2 = 1 Theta carries the reference's own bits whenever exp() rounds z = 1 + e the same way
    const double Theta = 1.0 - divExact(1.0, d.invC2IsOne ? z : fastPow(z, d.invC2));
    o.sigmaf = (H * Lam) * Theta;                                         // :129
    o.sigmay = (p.a1 * Theta) * (1.0 + pw * d.b3LogY);                    // :152
    o.nu = ofMax(p.nuMin, ofMin(p.nuMax, divFast(o.sigmaf, d.threeRho * ofMax(e, p.eDotMin))));  // :172-181
    return o;
}

// NortonHoff.C:53-66, sr = viscosityModel::strainRate() = sqrt(2)*mag(symm(grad U)); the division by
// strainRateUnits_ = 1 is the identity
SSAM_DEV double nortonHoff(const ssam_viscosity_params& p, const ViscDerived& d, double sr) {
    return ofMax(p.nuMin,
                 ofMin(p.nuMax, d.mu0OverRho * fastPow(d.sqrt3 * ofMax(sr, p.strainRateEffMin), d.mMinus1)));
}

// ---- mixture (incompressibleTwoPhaseThermalMixture.C) ----------------------------------------------
SSAM_DEV double limitedAlpha(double a) { return ofMin(ofMax(a, 0.0), 1.0); }  // :49,:138

SSAM_DEV double mixMu(const ssam_mixture_params& p, double la, double nu1, double nu2) {  // :149-150
    return ((la * la) * p.rho1) * nu1 + (((1.0 - la) * (1.0 - la)) * p.rho2) * nu2;
}
SSAM_DEV double mixLinear(double la, double v1, double v2) { return la * v1 + (1.0 - la) * v2; }  // :359,:378,:397

// k1()/k2() :205-343.  shift = 0 (Kelvin) or -273.0 (Celsius), resolved on the host.
SSAM_DEV double phaseK(const ssam_conductivity& k, double shift, double T) {
    if (k.kModel == SSAM_K_CUBIC) {
        const double Tl = ofMin(ofMax(T + shift, k.Tmin), k.Tmax);  // :244
        // :253  k0*(Tl/Tl) + k1*Tl + k2*pow(Tl,2) + k3*pow(Tl,3); pow(x,2|3) as products: glibc's pow
        // differs from the correctly rounded product by <= 1 ulp (inside the 1e-11 gate)
        return k.k0 * selfRatio(Tl) + k.k1 * Tl + k.k2 * (Tl * Tl) + k.k3 * ((Tl * Tl) * Tl);
    }
    return k.k * selfRatio(T);  // :269
}

// ---- fluid/calculateStrain.H:1-35,77-94: explicit stress / strain post-processing --------------------
// OpenFOAM tensor semantics: T() transpose, tr = xx+yy+zz, Tensor - SphericalTensor touches the diagonal
// only, A && B = sum_ij A_ij*B_ij in row-major order.  pow(x,2) as a product (<= 1 ulp from libm's pow).
struct StressStrainOut {
    double Z, sigmaf, dEpsilonEq, sigmavm;
};
SSAM_DEV StressStrainOut stressStrain(const ssam_stress_strain_params& p, double logA, double invN,
                                      const double (&g)[9], double eps, double T, double mu, double prgh) {
    StressStrainOut o;
    const double q = divFast(p.Q, p.R * T);
    o.Z = eps * fastExp(q);                                               // :12
    o.sigmaf = p.sigmaR * swAsinhOfLog((fastLog(eps) + (q - logA)) * invN);  // :13 in the log domain
    double dE[9], eD[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double sym = 0.5 * (g[3 * i + j] + g[3 * j + i]);
            dE[3 * i + j] = sym * p.deltaT;                               // :18
            eD[3 * i + j] = sym;                                          // :79
        }
    const double third = 1.0 / 3.0;
    const double s1 = third * (dE[0] + dE[4] + dE[8]);                    // :22
    dE[0] = dE[0] - s1; dE[4] = dE[4] - s1; dE[8] = dE[8] - s1;
    double dd = dE[0] * dE[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) dd += dE[k] * dE[k];
    const double trD = dE[0] + dE[4] + dE[8];
    const double J2s = 0.5 * (trD * trD - dd);                            // :26
    const double J2 = sqrt(J2s * J2s);                                    // :29
    o.dEpsilonEq = sqrt((4.0 / 3.0) * J2);                                // :35
    const double s2 = third * (eD[0] + eD[4] + eD[8]);                    // dev(epsilonDot) :82
    eD[0] = eD[0] - s2; eD[4] = eD[4] - s2; eD[8] = eD[8] - s2;
    const double twoMu = 2.0 * mu;
#pragma unroll
    for (int k = 0; k < 9; ++k) eD[k] = twoMu * eD[k];                    // shearStress :85
    eD[0] = prgh + eD[0]; eD[4] = prgh + eD[4]; eD[8] = prgh + eD[8];     // sigmaTot :88
    const double s3 = third * (eD[0] + eD[4] + eD[8]);
    eD[0] -= s3; eD[4] -= s3; eD[8] -= s3;                                // sigmadev :91
    double ss = eD[0] * eD[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) ss += eD[k] * eD[k];
    o.sigmavm = sqrt(2.3 / 3.0) * sqrt(ss);                               // :94
    return o;
}

// ---- friction (constantSlipStick.C / exponentialSlipStick.C) ---------------------------------------
SSAM_DEV double qviscOf(double betaTQ, double mu, double e) { return ((3.0 * betaTQ) * mu) * (e * e); }  // :58

// qfric at a friction cell (fricCell_ = 1): constantSlipStick.C:100, exponentialSlipStick.C:100,156,164
SSAM_DEV double qfricOf(const ssam_stick_params& p, double magOmega, double alpha, double sigmay, double pr,
                        double magR) {
    double inner;
    if (p.model == SSAM_STICK_CONSTANT) {
        inner = divExact(((1.0 - p.delta) * p.betaTQ) * sigmay, sqrt(3.0)) + (p.delta * p.muf) * pr;
    } else {
        const double d = 1.0 - fastExp(divExact((-magOmega) * magR, (p.delta0 * p.omega0) * p.Rs));
        const double mf = p.mu0 * fastExp((((-p.lambda) * d) * magOmega) * magR);
        inner = divExact(((1.0 - d) * p.betaTQ) * sigmay, sqrt(3.0)) + (d * mf) * pr;
    }
    return ((1.0 * alpha) * inner) * (magOmega * magR);
}

}  // namespace ssam
