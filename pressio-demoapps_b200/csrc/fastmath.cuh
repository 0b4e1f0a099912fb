// fastmath.cuh -- leaf arithmetic restated for the FP64 pipe, shared by the structured kernels:
// branch-free reciprocal / square root from the MUFU seeds (no IEEE slow-path subroutines), WENO3/5 face values with
// one reciprocal per face, WENO3/5 reconstruction gradients and the Euler / shallow-water Rusanov fluxes and flux
// Jacobians without a single IEEE division.  Same formulas as physics.cuh (which keeps the reference's operation
// order for the graph-driven kernels); results agree to a few ulp.
#pragma once
#include <type_traits>

#include "physics.cuh"

namespace pda {
namespace dev {

// ---------------------------------------------------------------------------------------------------------------
// Branch-free FP64 reciprocal and square root: MUFU seed (>= 20 good bits) + Newton steps; relative error ~2 ulp.
// Arguments are positive normal numbers here (densities, Roe averages, sums of squared smoothness indicators >=
// eps^2), so the IEEE slow paths (denormals, inf, signed zero) that make `/` and sqrt() a subroutine call with a
// divergent branch are not needed.  The reference's tolerance (1e-12) is four orders above this error.
// ---------------------------------------------------------------------------------------------------------------
PDA_DEVFN double rcpFast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
PDA_DEVFN double sqrtFast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // two Newton steps on y ~ x^-1/2:  y <- y + y*(1 - x*y*y)/2
  double t = x * y;
  double e = fma(-t, y, 1.0);
  y = fma(0.5 * y, e, y);
  t = x * y;
  e = fma(-t, y, 1.0);
  y = fma(0.5 * y, e, y);
  // s = x*y with one residual correction
  double s = x * y;
  const double r = fma(-s, s, x);
  s = fma(r, 0.5 * y, s);
  return (x == 0.0) ? 0.0 : s;   // x == 0: the seed is inf and the chain NaN; x < 0 stays NaN like sqrt()
}

// sqrt(k) for the wave speed smax = sqrt(|v_roe|^2) + a, where k may be arbitrarily small (gas almost at rest: far
// from a blast the velocities are numerical dust and their squares reach the denormal range, where the ftz seed of
// sqrtFast is inf and the Newton chain NaN).  Below 1e-200 the square root is < 1e-100 and vanishes next to the sound
// speed a in double precision, so such arguments are treated as zero: one compare + select, still branch-free.
PDA_DEVFN double sqrtFastTiny(double x) {
  return sqrtFast((x < 1.0e-200) ? 0.0 : x);
}

// ---------------------------------------------------------------------------------------------------------------
// WENO5 (Jiang-Shu) at one face from the six cells around it, both sides (impl/weno5.hpp:56-178):
// q = (a,b,c,d,e,f) = cells i-3..i+2, face between c and d; uNeg from (a..e), uPos from (b..f).
//   * smoothness indicators from first/second differences; E_k = 4*(eps + B_k), the common factor 16 cancels in the
//     weights: w_k = c_k/E_k^2 / sum;  multiplied through by E_0^2 E_1^2 E_2^2 -> no division per weight;
//   * candidate polynomials in difference form around c / d (fewer operations, less cancellation);
//   * uNeg = Nn/Dn, uPos = Np/Dp with ONE reciprocal: r = 1/(Dn*Dp).
// ~75 FP64 instructions per (face, dof) instead of ~135.
// ---------------------------------------------------------------------------------------------------------------
PDA_DEVFN void weno5FaceFast(const double* q, double& uNeg, double& uPos) {
  constexpr double k133 = 13.0 / 3.0, eps4 = 4.0e-6, s6 = 1.0 / 6.0;
  const double c = q[2], d = q[3];
  const double d0 = q[1] - q[0], d1 = c - q[1], d2 = d - c, d3 = q[4] - d, d4 = q[5] - q[4];
  const double tb = d1 - d0, tc = d2 - d1, td = d3 - d2, te = d4 - d3;
  const double tb2 = tb * tb, tc2 = tc * tc, td2 = td * td, te2 = te * te;
  // neg side: s0 = a-4b+3c = 3 d1 - d0 ; s1 = b-d = -(d1+d2) ; s2 = 3c-4d+e = d3 - 3 d2
  const double sn0 = fma(3.0, d1, -d0), sn1 = d1 + d2, sn2 = fma(-3.0, d2, d3);
  // pos side: s0 = b-4c+3d = 3 d2 - d1 ; s1 = c-e = -(d2+d3) ; s2 = 3d-4e+f = d4 - 3 d3
  const double sp0 = fma(3.0, d2, -d1), sp1 = d2 + d3, sp2 = fma(-3.0, d3, d4);
  const double En0 = fma(sn0, sn0, fma(k133, tb2, eps4));
  const double En1 = fma(sn1, sn1, fma(k133, tc2, eps4));
  const double En2 = fma(sn2, sn2, fma(k133, td2, eps4));
  const double Ep0 = fma(sp0, sp0, fma(k133, tc2, eps4));
  const double Ep1 = fma(sp1, sp1, fma(k133, td2, eps4));
  const double Ep2 = fma(sp2, sp2, fma(k133, te2, eps4));
  const double Gn0 = En0 * En0, Gn1 = En1 * En1, Gn2 = En2 * En2;
  const double Gp0 = Ep0 * Ep0, Gp1 = Ep1 * Ep1, Gp2 = Ep2 * Ep2;
  // candidates: p(a,b,c) = c + (5 d1 - 2 d0)/6 ; p(b,c,d) = c + (2 d2 + d1)/6 ; p(c,d,e) = d - (2 d2 + d3)/6 ;
  //             p(d,e,f) = d + (2 d4 - 5 d3)/6
  const double pabc = fma(s6, fma(5.0, d1, -2.0 * d0), c);
  const double pbcd = fma(s6, fma(2.0, d2, d1), c);
  const double pcde = fma(-s6, fma(2.0, d2, d3), d);
  const double pdef = fma(s6, fma(2.0, d4, -5.0 * d3), d);
  // neg: linear weights (1,6,3)/10 ; pos: (3,6,1)/10
  const double n0 = Gn1 * Gn2, n1 = 6.0 * (Gn0 * Gn2), n2 = 3.0 * (Gn0 * Gn1);
  const double m0 = 3.0 * (Gp1 * Gp2), m1 = 6.0 * (Gp0 * Gp2), m2 = Gp0 * Gp1;
  const double Dn = n0 + (n1 + n2), Dp = m0 + (m1 + m2);
  const double Nn = fma(n0, pabc, fma(n1, pbcd, n2 * pcde));
  const double Np = fma(m0, pbcd, fma(m1, pcde, m2 * pdef));
  const double r = rcpFast(Dn * Dp);
  uNeg = Nn * (Dp * r);
  uPos = Np * (Dn * r);
}

// WENO3 at one face from the four cells around it (impl/weno3.hpp:56-114), both sides with ONE reciprocal
PDA_DEVFN void weno3FaceFast(const double* q, double& uNeg, double& uPos) {
  const double b = q[0], c = q[1], d = q[2], e = q[3];
  const double dbc = b - c, dcd = c - d, dde = d - e;
  const double Eb = fma(dbc, dbc, kWenoEps), Ec = fma(dcd, dcd, kWenoEps), Ed = fma(dde, dde, kWenoEps);
  const double Gb = Eb * Eb, Gc = Ec * Ec, Gd = Ed * Ed;
  const double pm = 0.5 * (c + d);
  const double p0 = 0.5 * fma(3.0, c, -b), p1 = 0.5 * fma(3.0, d, -e);
  // neg: w0 = Gc/(Gc + 2 Gb) on p0, rest on pm ; pos: w0 = 2 Gd/(2 Gd + Gc) on pm, rest on p1
  const double Dn = fma(2.0, Gb, Gc), Dp = fma(2.0, Gd, Gc);
  const double Nn = fma(Gc, p0, 2.0 * Gb * pm);
  const double Np = fma(2.0 * Gd, pm, Gc * p1);
  const double r = rcpFast(Dn * Dp);
  uNeg = Nn * (Dp * r);
  uPos = Np * (Dn * r);
}

template <int S> PDA_DEVFN void reconFaceFast(const double* q, double& uNeg, double& uPos) {
  if constexpr (S == 7) weno5FaceFast(q, uNeg, uPos);
  else if constexpr (S == 5) weno3FaceFast(q, uNeg, uPos);
  else Recon<S>::face(q, uNeg, uPos);
}

// Rusanov flux of the DIM-dimensional Euler equations along AX with the fast reciprocal / square root
// (impl/euler_rusanov_flux_values_function.hpp:54-208)
template <int DIM, int AX>
PDA_DEVFN void eulerFluxFast(double gamma, const double* qL, const double* qR, double* F) {
  constexpr int N = DIM + 2;
  const double gm1 = gamma - 1.0;
  const double rL = qL[0], rR = qR[0];
  const double iL = rcpFast(rL), iR = rcpFast(rR);
  double vL[DIM], vR[DIM];
  double kL = 0.0, kR = 0.0;
#pragma unroll
  for (int m = 0; m < DIM; ++m) {
    vL[m] = qL[1 + m] * iL; vR[m] = qR[1 + m] * iR;
    kL = fma(vL[m], vL[m], kL); kR = fma(vR[m], vR[m], kR);
  }
  const double pL = gm1 * fma(-0.5 * rL, kL, qL[N - 1]);
  const double pR = gm1 * fma(-0.5 * rR, kR, qR[N - 1]);
  const double HL = (qL[N - 1] + pL) * iL;
  const double HR = (qR[N - 1] + pR) * iR;
  const double mL = rL * vL[AX], mR = rR * vR[AX];
  const double RT = sqrtFast(rR * iL);
  const double iRT = rcpFast(1.0 + RT);
  double k = 0.0;
#pragma unroll
  for (int m = 0; m < DIM; ++m) { const double v = fma(RT, vR[m], vL[m]) * iRT; k = fma(v, v, k); }
  const double H = fma(RT, HR, HL) * iRT;
  const double a = sqrtFast(gm1 * fma(-0.5, k, H));
  const double smax = sqrtFastTiny(k) + a;
  const double pS = pL + pR;
  F[0] = 0.5 * fma(smax, rL - rR, mL + mR);
#pragma unroll
  for (int m = 0; m < DIM; ++m)
    F[1 + m] = 0.5 * (fma(smax, qL[1 + m] - qR[1 + m], fma(mL, vL[m], mR * vR[m])) + ((m == AX) ? pS : 0.0));
  F[N - 1] = 0.5 * fma(smax, qL[N - 1] - qR[N - 1], fma(mL, HL, mR * HR));
}

// shallow-water Rusanov flux (impl/swe_rusanov_flux_values_function.hpp:54-97) with the fast reciprocal / square
// root: depths are positive normal numbers, the 1e-30 guards of the reference vanish in double precision next to them
template <int AX>
PDA_DEVFN void sweFluxFast(double g, const double* qL, const double* qR, double* F) {
  const double hL = qL[0], hR = qR[0];
  const double iL = rcpFast(hL), iR = rcpFast(hR);
  const double uL = qL[1] * iL, vL = qL[2] * iL;
  const double uR = qR[1] * iR, vR = qR[2] * iR;
  const double unL = (AX == 0) ? uL : vL, unR = (AX == 0) ? uR : vR;
  const double pS = 0.5 * g * fma(hL, hL, hR * hR);
  const double sL = sqrtFast(hL), sR = sqrtFast(hR);
  const double um = fma(unL, sL, unR * sR) * rcpFast(sL + sR);
  const double smax = fabs(um) + sqrtFast(g * (0.5 * (hL + hR)));
  const double mL = hL * unL, mR = hR * unR;
  F[0] = 0.5 * fma(smax, qL[0] - qR[0], mL + mR);
  F[1] = 0.5 * (fma(smax, qL[1] - qR[1], fma(mL, uL, mR * uR)) + ((AX == 0) ? pS : 0.0));
  F[2] = 0.5 * (fma(smax, qL[2] - qR[2], fma(mL, vL, mR * vR)) + ((AX == 1) ? pS : 0.0));
}

// the same flux with 1/h and sqrt(h) of the two states handed in: with a FIRST-ORDER reconstruction the face states are
// cell values, and a cell's 1/h and sqrt(h) serve its four faces (identical operations -> identical bits as sweFluxFast)
template <int AX>
PDA_DEVFN void sweFluxFastPre(double g, const double* qL, const double* qR, double iL, double sL, double iR, double sR, double* F) {
  const double hL = qL[0], hR = qR[0];
  const double uL = qL[1] * iL, vL = qL[2] * iL;
  const double uR = qR[1] * iR, vR = qR[2] * iR;
  const double unL = (AX == 0) ? uL : vL, unR = (AX == 0) ? uR : vR;
  const double pS = 0.5 * g * fma(hL, hL, hR * hR);
  const double um = fma(unL, sL, unR * sR) * rcpFast(sL + sR);
  const double smax = fabs(um) + sqrtFast(g * (0.5 * (hL + hR)));
  const double mL = hL * unL, mR = hR * unR;
  F[0] = 0.5 * fma(smax, qL[0] - qR[0], mL + mR);
  F[1] = 0.5 * (fma(smax, qL[1] - qR[1], fma(mL, uL, mR * uR)) + ((AX == 0) ? pS : 0.0));
  F[2] = 0.5 * (fma(smax, qL[2] - qR[2], fma(mL, vL, mR * vR)) + ((AX == 1) ? pS : 0.0));
}

template <class Phys, int AX>
PDA_DEVFN void faceFlux2d(const Phys& phys, const double* uN, const double* uP, double* F) {
  if constexpr (std::is_same<Phys, Euler<1>>::value || std::is_same<Phys, Euler<2>>::value || std::is_same<Phys, Euler<3>>::value)
    eulerFluxFast<Phys::dim, AX>(phys.gamma, uN, uP, F);
  else if constexpr (std::is_same<Phys, Swe2d>::value) sweFluxFast<AX>(phys.g, uN, uP, F);
  else phys.template flux<AX>(uN, uP, F);
}


// 1/sqrt(x) for positive normal x: MUFU seed + two Newton steps
PDA_DEVFN double rsqrtFast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  e = fma(-x * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  return y;
}

// ---------------------------------------------------------------------------------------------------------------
// Reconstruction gradients (impl/weno5.hpp:180-434, impl/weno3.hpp:116-246) in the well-conditioned form of
// physics.cuh (c_k = d(alpha_k)/dB_k (p_k - u)/S), every division replaced: r_k = 1/(eps+B_k) once per candidate,
// alpha_k = c_k r_k^2, 1/S once per side.  4 reciprocals per side instead of 7 IEEE divisions.
// ---------------------------------------------------------------------------------------------------------------
PDA_DEVFN void weno5FaceValGradFast(const double* q, double& uNeg, double& uPos, double* gNeg, double* gPos) {
  const double a = q[0], b = q[1], c = q[2], d = q[3], e = q[4], f = q[5];
  constexpr double k13 = 13.0 / 12.0, s6 = 1.0 / 6.0, k136 = 13.0 / 6.0, k133 = 13.0 / 3.0;
  {
    const double p0 = (2.0 * a - 7.0 * b + 11.0 * c) * s6;
    const double p1 = (-b + 5.0 * c + 2.0 * d) * s6;
    const double p2 = (2.0 * c + 5.0 * d - e) * s6;
    const double t0 = a - 2.0 * b + c, s0 = a - 4.0 * b + 3.0 * c;
    const double t1 = b - 2.0 * c + d, s1 = b - d;
    const double t2 = c - 2.0 * d + e, s2 = 3.0 * c - 4.0 * d + e;
    const double r0 = rcpFast(kWenoEps + (k13 * t0 * t0 + 0.25 * s0 * s0));
    const double r1 = rcpFast(kWenoEps + (k13 * t1 * t1 + 0.25 * s1 * s1));
    const double r2 = rcpFast(kWenoEps + (k13 * t2 * t2 + 0.25 * s2 * s2));
    const double a0 = 0.1 * r0 * r0, a1 = 0.6 * r1 * r1, a2 = 0.3 * r2 * r2;
    const double invS = rcpFast(a0 + a1 + a2);
    const double w0 = a0 * invS, w1 = a1 * invS, w2 = a2 * invS;
    const double u = w0 * p0 + w1 * p1 + w2 * p2;
    uNeg = u;
    const double c0 = -2.0 * w0 * r0 * (p0 - u);
    const double c1 = -2.0 * w1 * r1 * (p1 - u);
    const double c2 = -2.0 * w2 * r2 * (p2 - u);
    gNeg[0] = c0 * (k136 * t0 + 0.5 * s0) + w0 * (1.0 / 3.0);
    gNeg[1] = c0 * (-k133 * t0 - 2.0 * s0) + c1 * (k136 * t1 + 0.5 * s1) + w0 * (-7.0 / 6.0) + w1 * (-1.0 / 6.0);
    gNeg[2] = c0 * (k136 * t0 + 1.5 * s0) + c1 * (-k133 * t1) + c2 * (k136 * t2 + 1.5 * s2)
            + w0 * (11.0 / 6.0) + w1 * (5.0 / 6.0) + w2 * (1.0 / 3.0);
    gNeg[3] = c1 * (k136 * t1 - 0.5 * s1) + c2 * (-k133 * t2 - 2.0 * s2) + w1 * (1.0 / 3.0) + w2 * (5.0 / 6.0);
    gNeg[4] = c2 * (k136 * t2 + 0.5 * s2) + w2 * (-1.0 / 6.0);
    gNeg[5] = 0.0;
  }
  {
    const double p0 = (-b + 5.0 * c + 2.0 * d) * s6;
    const double p1 = (2.0 * c + 5.0 * d - e) * s6;
    const double p2 = (11.0 * d - 7.0 * e + 2.0 * f) * s6;
    const double t0 = b - 2.0 * c + d, s0 = b - 4.0 * c + 3.0 * d;
    const double t1 = c - 2.0 * d + e, s1 = c - e;
    const double t2 = d - 2.0 * e + f, s2 = 3.0 * d - 4.0 * e + f;
    const double r0 = rcpFast(kWenoEps + (k13 * t0 * t0 + 0.25 * s0 * s0));
    const double r1 = rcpFast(kWenoEps + (k13 * t1 * t1 + 0.25 * s1 * s1));
    const double r2 = rcpFast(kWenoEps + (k13 * t2 * t2 + 0.25 * s2 * s2));
    const double a0 = 0.3 * r0 * r0, a1 = 0.6 * r1 * r1, a2 = 0.1 * r2 * r2;
    const double invS = rcpFast(a0 + a1 + a2);
    const double w0 = a0 * invS, w1 = a1 * invS, w2 = a2 * invS;
    const double u = w0 * p0 + w1 * p1 + w2 * p2;
    uPos = u;
    const double c0 = -2.0 * w0 * r0 * (p0 - u);
    const double c1 = -2.0 * w1 * r1 * (p1 - u);
    const double c2 = -2.0 * w2 * r2 * (p2 - u);
    gPos[0] = 0.0;
    gPos[1] = c0 * (k136 * t0 + 0.5 * s0) + w0 * (-1.0 / 6.0);
    gPos[2] = c0 * (-k133 * t0 - 2.0 * s0) + c1 * (k136 * t1 + 0.5 * s1) + w0 * (5.0 / 6.0) + w1 * (1.0 / 3.0);
    gPos[3] = c0 * (k136 * t0 + 1.5 * s0) + c1 * (-k133 * t1) + c2 * (k136 * t2 + 1.5 * s2)
            + w0 * (1.0 / 3.0) + w1 * (5.0 / 6.0) + w2 * (11.0 / 6.0);
    gPos[4] = c1 * (k136 * t1 - 0.5 * s1) + c2 * (-k133 * t2 - 2.0 * s2) + w1 * (-1.0 / 6.0) + w2 * (-7.0 / 6.0);
    gPos[5] = c2 * (k136 * t2 + 0.5 * s2) + w2 * (1.0 / 3.0);
  }
}

PDA_DEVFN void weno5FaceGradFast(const double* q, double* gNeg, double* gPos) {
  double uN, uP;
  weno5FaceValGradFast(q, uN, uP, gNeg, gPos);
}

PDA_DEVFN void weno3FaceValGradFast(const double* q, double& uNeg, double& uPos, double* gNeg, double* gPos) {
  const double b = q[0], c = q[1], d = q[2], e = q[3];
  const double pm = 0.5 * (c + d);
  {
    const double p0 = 0.5 * (3.0 * c - b);
    const double r0 = rcpFast(kWenoEps + sq(b - c)), r1 = rcpFast(kWenoEps + sq(c - d));
    const double a0 = (1.0 / 3.0) * r0 * r0, a1 = (2.0 / 3.0) * r1 * r1;
    const double invS = rcpFast(a0 + a1);
    const double w0 = a0 * invS, w1 = a1 * invS;
    const double u = w0 * p0 + w1 * pm;
    uNeg = u;
    const double h0 = -4.0 * w0 * r0 * (b - c) * (p0 - u);
    const double h1 = -4.0 * w1 * r1 * (c - d) * (pm - u);
    gNeg[0] = h0 - 0.5 * w0;
    gNeg[1] = -h0 + h1 + 1.5 * w0 + 0.5 * w1;
    gNeg[2] = -h1 + 0.5 * w1;
    gNeg[3] = 0.0;
  }
  {
    const double p1 = 0.5 * (3.0 * d - e);
    const double r0 = rcpFast(kWenoEps + sq(c - d)), r1 = rcpFast(kWenoEps + sq(d - e));
    const double a0 = (2.0 / 3.0) * r0 * r0, a1 = (1.0 / 3.0) * r1 * r1;
    const double invS = rcpFast(a0 + a1);
    const double w0 = a0 * invS, w1 = a1 * invS;
    const double u = w0 * pm + w1 * p1;
    uPos = u;
    const double h0 = -4.0 * w0 * r0 * (c - d) * (pm - u);
    const double h1 = -4.0 * w1 * r1 * (d - e) * (p1 - u);
    gPos[0] = 0.0;
    gPos[1] = h0 + 0.5 * w0;
    gPos[2] = -h0 + h1 + 0.5 * w0 + 1.5 * w1;
    gPos[3] = -h1 - 0.5 * w1;
  }
}

PDA_DEVFN void weno3FaceGradFast(const double* q, double* gNeg, double* gPos) {
  double uN, uP;
  weno3FaceValGradFast(q, uN, uP, gNeg, gPos);
}

// face values AND their gradients in one pass: the gradient formulas already hold the reconstructed value
// (u = sum w_k p_k), so kernels that need both skip the separate value reconstruction (~75 FP64 instructions per
// (face, dof) for WENO5)
template <int S> PDA_DEVFN void reconFaceValGradFast(const double* q, double& uNeg, double& uPos, double* gNeg, double* gPos) {
  if constexpr (S == 7) weno5FaceValGradFast(q, uNeg, uPos, gNeg, gPos);
  else if constexpr (S == 5) weno3FaceValGradFast(q, uNeg, uPos, gNeg, gPos);
  else Recon<S>::faceGrad(q, uNeg, uPos, gNeg, gPos);
}

template <int S> PDA_DEVFN void reconFaceGradFast(const double* q, double* gNeg, double* gPos) {
  if constexpr (S == 7) weno5FaceGradFast(q, gNeg, gPos);
  else if constexpr (S == 5) weno3FaceGradFast(q, gNeg, gPos);
  else { double t0, t1; Recon<S>::faceGrad(q, t0, t1, gNeg, gPos); }
}

// ---------------------------------------------------------------------------------------------------------------
// Euler Rusanov flux Jacobians JL = dF/dqL, JR = dF/dqR (row-major [N][N]) along AX
// (impl/euler_rusanov_flux_jacobian_function.hpp:54-406; same terms as Euler<DIM>::fluxJac in physics.cuh) with
// reciprocals / square roots from the fast primitives: 0 IEEE divisions, 0 IEEE square roots.
// ---------------------------------------------------------------------------------------------------------------
template <int DIM, int AX>
PDA_DEVFN void eulerPhysicalHalfJac(double gamma, double* J, const double* vel, double k2, double H, double un, double gm1) {
  constexpr int N = DIM + 2;
  J[0] = 0.0;
#pragma unroll
  for (int j = 0; j < DIM; ++j) J[1 + j] = (j == AX) ? 0.5 : 0.0;
  J[N - 1] = 0.0;
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    const double ni = (i == AX) ? 1.0 : 0.0;
    J[(1 + i) * N + 0] = 0.5 * (0.5 * gm1 * k2 * ni - vel[i] * un);
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
      const double nj = (j == AX) ? 1.0 : 0.0;
      J[(1 + i) * N + 1 + j] = 0.5 * (vel[i] * nj - gm1 * vel[j] * ni + ((i == j) ? un : 0.0));
    }
    J[(1 + i) * N + N - 1] = 0.5 * gm1 * ni;
  }
  J[(N - 1) * N + 0] = 0.5 * ((0.5 * gm1 * k2 - H) * un);
#pragma unroll
  for (int j = 0; j < DIM; ++j) {
    const double nj = (j == AX) ? 1.0 : 0.0;
    J[(N - 1) * N + 1 + j] = 0.5 * (H * nj - gm1 * vel[j] * un);
  }
  J[(N - 1) * N + N - 1] = 0.5 * gamma * un;
}

template <int DIM, int AX>
PDA_DEVFN void eulerFluxJacFast(double gamma, const double* qL, const double* qR, double* JL, double* JR) {
  constexpr int N = DIM + 2;
  const double gm1 = gamma - 1.0;
  const double rL = qL[0], rR = qR[0];
  const double iL = rcpFast(rL), iR = rcpFast(rR);   // densities are far above the reference's 1e-30 guard
  double vL[DIM], vR[DIM], v[DIM];
  double kL = 0.0, kR = 0.0;
#pragma unroll
  for (int m = 0; m < DIM; ++m) {
    vL[m] = qL[1 + m] * iL; vR[m] = qR[1 + m] * iR;
    kL += vL[m] * vL[m]; kR += vR[m] * vR[m];
  }
  const double pL = gm1 * (qL[DIM + 1] - 0.5 * rL * kL);
  const double pR = gm1 * (qR[DIM + 1] - 0.5 * rR * kR);
  const double HL = (qL[DIM + 1] + pL) * iL;
  const double HR = (qR[DIM + 1] + pR) * iR;
  const double unL = vL[AX], unR = vR[AX];
  const double RT = sqrtFast(rR * iL);
  const double r = rL * RT;                    // sqrt(rR*rL)
  const double iRT = rcpFast(1.0 + RT);
  double k = 0.0, dotL = 0.0, dotR = 0.0;
#pragma unroll
  for (int m = 0; m < DIM; ++m) {
    v[m] = (vL[m] + RT * vR[m]) * iRT;
    k += v[m] * v[m]; dotL += vL[m] * v[m]; dotR += vR[m] * v[m];
  }
  const double H = (HL + RT * HR) * iRT;
  const double a2 = gm1 * (H - 0.5 * k);
  const double ia = rsqrtFast(a2);
  const double a = a2 * ia;
  const double vmag2 = k + kEs;
  const double ivmag = rsqrtFast(vmag2);
  const double smax = vmag2 * ivmag + a;       // sqrt(k + 1e-30) + a: equals sqrt(k) + a to rounding
  double gL[N], gR[N];
  const double iLr = rcpFast(rL + r), iRr = rcpFast(rR + r);
  {
    double sL = 0.0, sR = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      const double rel = v[m] * ivmag;
      sL -= 0.5 * (vL[m] + v[m]) * rel;
      sR -= 0.5 * (vR[m] + v[m]) * rel;
      gL[1 + m] = iLr * (rel - 0.5 * (gm1 * (v[m] + gm1 * vL[m])) * ia);
      gR[1 + m] = iRr * (rel - 0.5 * (gm1 * (v[m] + gm1 * vR[m])) * ia);
    }
    // aL2/gm1 = HL - 0.5 kL
    gL[0] = iLr * (sL + 0.5 * gm1 * ia * (0.5 * (vmag2 + dotL) + 0.5 * (HL - H) - (HL - 0.5 * kL) + 0.5 * (gamma - 2.0) * kL));
    gR[0] = iRr * (sR + 0.5 * gm1 * ia * (0.5 * (vmag2 + dotR) + 0.5 * (HR - H) - (HR - 0.5 * kR) + 0.5 * (gamma - 2.0) * kR));
    gL[N - 1] = 0.5 * iLr * gamma * gm1 * ia;
    gR[N - 1] = 0.5 * iRr * gamma * gm1 * ia;
  }
  eulerPhysicalHalfJac<DIM, AX>(gamma, JL, vL, kL, HL, unL, gm1);
  eulerPhysicalHalfJac<DIM, AX>(gamma, JR, vR, kR, HR, unR, gm1);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double dq = 0.5 * (qL[i] - qR[i]);
    JL[i * N + i] += 0.5 * smax;
    JR[i * N + i] -= 0.5 * smax;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      JL[i * N + j] += gL[j] * dq;
      JR[i * N + j] += gR[j] * dq;
    }
  }
}

// Jacobian-VECTOR product of the Euler Rusanov flux: D[c] = JL dL[c] + JR dR[c] for NC direction pairs, without forming
// the two N x N matrices.  Same terms as eulerFluxJacFast:  JL = A(qL)/2 + smax/2 I + dq (x) gL,  JR = A(qR)/2 - smax/2 I
// + dq (x) gR  with dq = (qL - qR)/2 and gL/R = d(smax)/dqL/R, so
//     D = 1/2 A(qL) dL + 1/2 A(qR) dR + smax/2 (dL - dR) + dq (gL.dL + gR.dR),
// and A(q) d is the directional derivative of the physical flux in closed form (one dot product v.dm per side).
// ~125 FP64 instructions fewer per face than matrices + products, and 2 N^2 fewer live registers.
template <int DIM, int AX, int NC>
PDA_DEVFN void eulerFluxJvpFast(double gamma, const double* qL, const double* qR, const double (*dL)[DIM + 2],
                                const double (*dR)[DIM + 2], double (*D)[DIM + 2]) {
  constexpr int N = DIM + 2;
  const double gm1 = gamma - 1.0;
  const double rL = qL[0], rR = qR[0];
  const double iL = rcpFast(rL), iR = rcpFast(rR);
  double vL[DIM], vR[DIM], v[DIM];
  double kL = 0.0, kR = 0.0;
#pragma unroll
  for (int m = 0; m < DIM; ++m) {
    vL[m] = qL[1 + m] * iL; vR[m] = qR[1 + m] * iR;
    kL += vL[m] * vL[m]; kR += vR[m] * vR[m];
  }
  const double pL = gm1 * (qL[DIM + 1] - 0.5 * rL * kL);
  const double pR = gm1 * (qR[DIM + 1] - 0.5 * rR * kR);
  const double HL = (qL[DIM + 1] + pL) * iL;
  const double HR = (qR[DIM + 1] + pR) * iR;
  const double unL = vL[AX], unR = vR[AX];
  const double RT = sqrtFast(rR * iL);
  const double r = rL * RT;
  const double iRT = rcpFast(1.0 + RT);
  double k = 0.0, dotL = 0.0, dotR = 0.0;
#pragma unroll
  for (int m = 0; m < DIM; ++m) {
    v[m] = (vL[m] + RT * vR[m]) * iRT;
    k += v[m] * v[m]; dotL += vL[m] * v[m]; dotR += vR[m] * v[m];
  }
  const double H = (HL + RT * HR) * iRT;
  const double a2 = gm1 * (H - 0.5 * k);
  const double ia = rsqrtFast(a2);
  const double a = a2 * ia;
  const double vmag2 = k + kEs;
  const double ivmag = rsqrtFast(vmag2);
  const double smax = vmag2 * ivmag + a;
  double gL[N], gR[N];
  const double iLr = rcpFast(rL + r), iRr = rcpFast(rR + r);
  {
    double sL = 0.0, sR = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      const double rel = v[m] * ivmag;
      sL -= 0.5 * (vL[m] + v[m]) * rel;
      sR -= 0.5 * (vR[m] + v[m]) * rel;
      gL[1 + m] = iLr * (rel - 0.5 * (gm1 * (v[m] + gm1 * vL[m])) * ia);
      gR[1 + m] = iRr * (rel - 0.5 * (gm1 * (v[m] + gm1 * vR[m])) * ia);
    }
    gL[0] = iLr * (sL + 0.5 * gm1 * ia * (0.5 * (vmag2 + dotL) + 0.5 * (HL - H) - (HL - 0.5 * kL) + 0.5 * (gamma - 2.0) * kL));
    gR[0] = iRr * (sR + 0.5 * gm1 * ia * (0.5 * (vmag2 + dotR) + 0.5 * (HR - H) - (HR - 0.5 * kR) + 0.5 * (gamma - 2.0) * kR));
    gL[N - 1] = 0.5 * iLr * gamma * gm1 * ia;
    gR[N - 1] = 0.5 * iRr * gamma * gm1 * ia;
  }
  // A(q) d for one side: rows of eulerPhysicalHalfJac times d (the factor 1/2 is applied by the caller)
  auto physJvp = [&](const double* vel, double k2, double Hs, double un, const double* d, double* out) {
    double vdm = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) vdm += vel[m] * d[1 + m];
    const double dmn = d[1 + AX];
    out[0] = dmn;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      double t = fma(-vel[i] * un, d[0], fma(vel[i], dmn, un * d[1 + i]));
      if (i == AX) t += gm1 * (fma(0.5 * k2, d[0], d[N - 1]) - vdm);
      out[1 + i] = t;
    }
    out[N - 1] = fma((0.5 * gm1 * k2 - Hs) * un, d[0], fma(Hs, dmn, fma(-gm1 * un, vdm, gamma * un * d[N - 1])));
  };
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double aL[N], aR[N];
    physJvp(vL, kL, HL, unL, dL[c], aL);
    physJvp(vR, kR, HR, unR, dR[c], aR);
    double gd = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) gd += gL[j] * dL[c][j] + gR[j] * dR[c][j];
#pragma unroll
    for (int i = 0; i < N; ++i)
      D[c][i] = 0.5 * ((aL[i] + aR[i]) + smax * (dL[c][i] - dR[c][i]) + (qL[i] - qR[i]) * gd);
  }
}

// shallow-water Rusanov flux Jacobians (impl/swe_rusanov_flux_jacobian_function.hpp:54-136; same terms as
// Swe2d::fluxJac in physics.cuh): 1/h, 1/sqrt(h), 1/(sL+sR) and 1/sqrt(g(hL+hR)) once each, no IEEE division
template <int AX>
PDA_DEVFN void sweFluxJacFast(double g, const double* qL, const double* qR, double* JL, double* JR) {
  constexpr double nx = (AX == 0) ? 1.0 : 0.0, ny = (AX == 1) ? 1.0 : 0.0;
  const double hL = qL[0], hR = qR[0];
  const double iL = rcpFast(hL), iR = rcpFast(hR);
  const double uL = qL[1] * iL, vL = qL[2] * iL, uR = qR[1] * iR, vR = qR[2] * iR;
  const double unL = (AX == 0) ? uL : vL, unR = (AX == 0) ? uR : vR;
  const double sL = sqrtFast(hL), sR = sqrtFast(hR);
  const double isL = sL * iL, isR = sR * iR;            // 1/sqrt(h)
  const double ss = sL + sR;
  const double iss = rcpFast(ss);
  const double hsun = sL * unL + sR * unR + kEs;
  const double ahs = fabs(hsun);
  const double sgn = (hsun < 0.0) ? -1.0 : 1.0;         // hsun/|hsun|
  const double um = (unL * sL + unR * sR) * iss;
  const double smax = fabs(um) + sqrtFast(g * (0.5 * (hL + hR)));
  const double termL = unL * iL, termR = unR * iR;      // (n.q)/h^2
  const double gterm = g * rsqrtFast(g * (hL + hR)) * 0.35355339059327373;   // g / (2^(3/2) sqrt(g (hL+hR)))
  const double iss2 = iss * iss;
  double dL[3], dR[3];
  dL[0] = -0.5 * ahs * isL * iss2 + (0.5 * unL * isL - sL * termL) * sgn * iss + gterm;
  dL[1] = nx * sgn * isL * iss;
  dL[2] = ny * sgn * isL * iss;
  dR[0] = -0.5 * ahs * isR * iss2 + (0.5 * unR * isR - sR * termR) * sgn * iss + gterm;
  dR[1] = nx * sgn * isR * iss;
  dR[2] = ny * sgn * isR * iss;
  const double d0 = qR[0] - qL[0], d1 = qR[1] - qL[1], d2 = qR[2] - qL[2];

  JL[0] = -0.5 * dL[0] * d0 + 0.5 * (nx * uL + ny * vL - qL[0] * termL) + 0.5 * smax;
  JL[1] = 0.5 * nx - 0.5 * dL[1] * d0;
  JL[2] = 0.5 * ny - 0.5 * dL[2] * d0;
  JL[3] = 0.5 * (g * nx * qL[0] - qL[1] * termL) - 0.5 * dL[0] * d1;
  JL[4] = nx * uL + 0.5 * ny * vL + 0.5 * smax - 0.5 * dL[1] * d1;
  JL[5] = 0.5 * ny * uL - 0.5 * dL[2] * d1;
  JL[6] = 0.5 * (g * ny * qL[0] - qL[2] * termL) - 0.5 * dL[0] * d2;
  JL[7] = 0.5 * nx * vL - 0.5 * dL[1] * d2;
  JL[8] = ny * vL + 0.5 * nx * uL + 0.5 * smax - 0.5 * dL[2] * d2;

  JR[0] = -0.5 * dR[0] * d0 + 0.5 * (nx * uR + ny * vR - qR[0] * termR) - 0.5 * smax;
  JR[1] = 0.5 * nx - 0.5 * dR[1] * d0;
  JR[2] = 0.5 * ny - 0.5 * dR[2] * d0;
  JR[3] = 0.5 * (g * nx * qR[0] - qR[1] * termR) - 0.5 * dR[0] * d1;
  JR[4] = nx * uR + 0.5 * ny * vR - 0.5 * smax - 0.5 * dR[1] * d1;
  JR[5] = 0.5 * ny * uR - 0.5 * dR[2] * d1;
  JR[6] = 0.5 * (g * ny * qR[0] - qR[2] * termR) - 0.5 * dR[0] * d2;
  JR[7] = 0.5 * nx * vR - 0.5 * dR[1] * d2;
  JR[8] = ny * vR + 0.5 * nx * uR - 0.5 * smax - 0.5 * dR[2] * d2;
}

template <class Phys, int AX>
PDA_DEVFN void faceFluxJac2d(const Phys& phys, const double* uN, const double* uP, double* JN, double* JP) {
  if constexpr (std::is_same<Phys, Euler<1>>::value || std::is_same<Phys, Euler<2>>::value || std::is_same<Phys, Euler<3>>::value)
    eulerFluxJacFast<Phys::dim, AX>(phys.gamma, uN, uP, JN, JP);
  else if constexpr (std::is_same<Phys, Swe2d>::value) sweFluxJacFast<AX>(phys.g, uN, uP, JN, JP);
  else phys.template fluxJac<AX>(uN, uP, JN, JP);
}

}  // namespace dev
}  // namespace pda
