// physics.cuh -- device leaf arithmetic of the finite-volume hot path (FP64, sm_100a).
//
// What the reference computes (cited per function) restated for a GPU thread: everything lives in registers,
// loops are compile-time unrolled, and the division count is cut (one division per one-sided WENO
// reconstruction instead of four, reciprocals shared inside the Rusanov flux) -- FP64 divides are ~10x an FMA on
// B200 and the velocity kernels are FP64-pipe bound.  Results agree with the reference to rounding
// (<= 1e-12 relative), not bit-for-bit; the parity tests in tests/ pin that.
#pragma once
#include <cuda_runtime.h>

namespace pda {
namespace dev {

#define PDA_DEVFN __device__ __forceinline__

constexpr double kWenoEps = 1.0e-6;   // impl/weno5.hpp:107, impl/weno3.hpp:89
constexpr double kEs = 1.0e-30;       // impl/euler_rusanov_flux_values_function.hpp:66

PDA_DEVFN double sq(double v) { return v * v; }

// ---------------------------------------------------------------------------------------------------------------
// Reconstruction at ONE face from the S-1 cell values around it:  q[0..S-2], the face sits between q[h-1] and q[h]
// (h = (S-1)/2).  uNeg = state on the minus side of the face (reference: uMinusHalfNeg / uPlusHalfNeg),
// uPos = state on the plus side.  Reference: weno.hpp:300-328 -> impl/weno5.hpp:56-178, impl/weno3.hpp:56-114;
// first order: functor_reconstruct_from_state.hpp (uNeg = left cell, uPos = right cell).
// ---------------------------------------------------------------------------------------------------------------
template <int S> struct Recon;

template <> struct Recon<3> {
  PDA_DEVFN static void face(const double* q, double& uNeg, double& uPos) { uNeg = q[0]; uPos = q[1]; }
  // gradients w.r.t. q[0..1]
  PDA_DEVFN static void faceGrad(const double* q, double& uNeg, double& uPos, double* gNeg, double* gPos) {
    uNeg = q[0]; uPos = q[1];
    gNeg[0] = 1.0; gNeg[1] = 0.0;
    gPos[0] = 0.0; gPos[1] = 1.0;
  }
};

// WENO3 (Jiang-Shu, eps = 1e-6, linear weights 1/3, 2/3).  q = (b,c,d,e) = cells (i-2,i-1,i,i+1), face (i-1 | i).
template <> struct Recon<5> {
  PDA_DEVFN static void face(const double* q, double& uNeg, double& uPos) {
    const double b = q[0], c = q[1], d = q[2], e = q[3];
    const double pm = 0.5 * (c + d);
    {
      const double p0 = 0.5 * (3.0 * c - b);
      const double D0 = sq(kWenoEps + sq(b - c));
      const double D1 = sq(kWenoEps + sq(c - d));
      // w0 = (1/3)/D0 / ((1/3)/D0 + (2/3)/D1) = D1 / (D1 + 2 D0)
      const double n0 = D1, n1 = 2.0 * D0;
      uNeg = (n0 * p0 + n1 * pm) / (n0 + n1);
    }
    {
      const double p1 = 0.5 * (3.0 * d - e);
      const double D0 = sq(kWenoEps + sq(c - d));
      const double D1 = sq(kWenoEps + sq(d - e));
      const double n0 = 2.0 * D1, n1 = D0;
      uPos = (n0 * pm + n1 * p1) / (n0 + n1);
    }
  }
  // gradients w.r.t. q[0..3]; layout of impl/weno3.hpp:116-246 (gNeg[3] = 0, gPos[0] = 0)
  PDA_DEVFN static void faceGrad(const double* q, double& uNeg, double& uPos, double* gNeg, double* gPos) {
    const double b = q[0], c = q[1], d = q[2], e = q[3];
    const double pm = 0.5 * (c + d);
    {
      const double p0 = 0.5 * (3.0 * c - b);
      const double e0 = kWenoEps + sq(b - c), e1 = kWenoEps + sq(c - d);
      const double a0 = (1.0 / 3.0) / (e0 * e0), a1 = (2.0 / 3.0) / (e1 * e1);
      const double invS = 1.0 / (a0 + a1);
      const double w0 = a0 * invS, w1 = a1 * invS;
      const double u = w0 * p0 + w1 * pm;
      // d(alpha_k)/dB_k = -2 alpha_k/(eps+B_k);  dB0 = 2(b-c)(db - dc), dB1 = 2(c-d)(dc - dd)
      const double h0 = -2.0 * a0 / e0 * 2.0 * (b - c) * invS * (p0 - u);
      const double h1 = -2.0 * a1 / e1 * 2.0 * (c - d) * invS * (pm - u);
      gNeg[0] = h0 - 0.5 * w0;
      gNeg[1] = -h0 + h1 + 1.5 * w0 + 0.5 * w1;
      gNeg[2] = -h1 + 0.5 * w1;
      gNeg[3] = 0.0;
      uNeg = u;
    }
    {
      const double p1 = 0.5 * (3.0 * d - e);
      const double e0 = kWenoEps + sq(c - d), e1 = kWenoEps + sq(d - e);
      const double a0 = (2.0 / 3.0) / (e0 * e0), a1 = (1.0 / 3.0) / (e1 * e1);
      const double invS = 1.0 / (a0 + a1);
      const double w0 = a0 * invS, w1 = a1 * invS;
      const double u = w0 * pm + w1 * p1;
      const double h0 = -2.0 * a0 / e0 * 2.0 * (c - d) * invS * (pm - u);
      const double h1 = -2.0 * a1 / e1 * 2.0 * (d - e) * invS * (p1 - u);
      gPos[0] = 0.0;
      gPos[1] = h0 + 0.5 * w0;
      gPos[2] = -h0 + h1 + 0.5 * w0 + 1.5 * w1;
      gPos[3] = -h1 - 0.5 * w1;
      uPos = u;
    }
  }
};

// WENO5 (Jiang-Shu, eps = 1e-6, linear weights 1/10, 6/10, 3/10).
// q = (a,b,c,d,e,f) = cells (i-3 .. i+2), face (i-1 | i).  impl/weno5.hpp:56-178.
template <> struct Recon<7> {
  PDA_DEVFN static void face(const double* q, double& uNeg, double& uPos) {
    const double a = q[0], b = q[1], c = q[2], d = q[3], e = q[4], f = q[5];
    constexpr double k13 = 13.0 / 12.0;
    constexpr double s6 = 1.0 / 6.0;
    // shared second differences and candidate polynomials
    const double tb = a - 2.0 * b + c;   // centred at b
    const double tc = b - 2.0 * c + d;   // centred at c
    const double td = c - 2.0 * d + e;   // centred at d
    const double te = d - 2.0 * e + f;   // centred at e
    const double tc2 = k13 * tc * tc, td2 = k13 * td * td;
    const double pcd = (-b + 5.0 * c + 2.0 * d) * s6;   // p1(neg) == p0(pos)
    const double pde = (2.0 * c + 5.0 * d - e) * s6;    // p2(neg) == p1(pos)
    {
      const double p0 = (2.0 * a - 7.0 * b + 11.0 * c) * s6;
      const double B0 = k13 * tb * tb + 0.25 * sq(a - 4.0 * b + 3.0 * c);
      const double B1 = tc2 + 0.25 * sq(b - d);
      const double B2 = td2 + 0.25 * sq(3.0 * c - 4.0 * d + e);
      const double D0 = sq(kWenoEps + B0), D1 = sq(kWenoEps + B1), D2 = sq(kWenoEps + B2);
      // alpha_k = c_k / D_k  ->  w_k = c_k prod_{j!=k} D_j / sum(...)   (one division instead of four)
      const double n0 = D1 * D2, n1 = 6.0 * (D0 * D2), n2 = 3.0 * (D0 * D1);
      uNeg = (n0 * p0 + n1 * pcd + n2 * pde) / (n0 + n1 + n2);
    }
    {
      const double p2 = (11.0 * d - 7.0 * e + 2.0 * f) * s6;
      const double B0 = tc2 + 0.25 * sq(b - 4.0 * c + 3.0 * d);
      const double B1 = td2 + 0.25 * sq(c - e);
      const double B2 = k13 * te * te + 0.25 * sq(3.0 * d - 4.0 * e + f);
      const double D0 = sq(kWenoEps + B0), D1 = sq(kWenoEps + B1), D2 = sq(kWenoEps + B2);
      const double n0 = 3.0 * (D1 * D2), n1 = 6.0 * (D0 * D2), n2 = D0 * D1;
      uPos = (n0 * pcd + n1 * pde + n2 * p2) / (n0 + n1 + n2);
    }
  }

  // one-sided WENO5 with gradient on 5 points (v0..v4), linear weights (c0,c1,c2), polynomial coefficients fixed by
  // `MINUS` (true: value at the right edge of the centre cell v2, false: value at the left edge of cell v2... see use).
  // Generic helper: candidates p_k = sum_m P[k][m] v_m, smoothness B_k from (t_k, s_k).
  // gradients w.r.t. q[0..5]; layout of impl/weno5.hpp:180-434 (gNeg[5] = 0, gPos[0] = 0).
  PDA_DEVFN static void faceGrad(const double* q, double& uNeg, double& uPos, double* gNeg, double* gPos) {
    const double a = q[0], b = q[1], c = q[2], d = q[3], e = q[4], f = q[5];
    constexpr double k13 = 13.0 / 12.0;
    constexpr double s6 = 1.0 / 6.0;
    constexpr double k136 = 13.0 / 6.0, k133 = 13.0 / 3.0;
    {
      // stencil (a,b,c,d,e)
      const double p0 = (2.0 * a - 7.0 * b + 11.0 * c) * s6;
      const double p1 = (-b + 5.0 * c + 2.0 * d) * s6;
      const double p2 = (2.0 * c + 5.0 * d - e) * s6;
      const double t0 = a - 2.0 * b + c, s0 = a - 4.0 * b + 3.0 * c;
      const double t1 = b - 2.0 * c + d, s1 = b - d;
      const double t2 = c - 2.0 * d + e, s2 = 3.0 * c - 4.0 * d + e;
      const double e0 = kWenoEps + (k13 * t0 * t0 + 0.25 * s0 * s0);
      const double e1 = kWenoEps + (k13 * t1 * t1 + 0.25 * s1 * s1);
      const double e2 = kWenoEps + (k13 * t2 * t2 + 0.25 * s2 * s2);
      const double a0 = 0.1 / (e0 * e0), a1 = 0.6 / (e1 * e1), a2 = 0.3 / (e2 * e2);
      const double invS = 1.0 / (a0 + a1 + a2);
      const double w0 = a0 * invS, w1 = a1 * invS, w2 = a2 * invS;
      const double u = w0 * p0 + w1 * p1 + w2 * p2;
      // c_k = d(alpha_k)/dB_k * (p_k - u) / S      (sum_k dw_k p_k = (1/S) sum_k dalpha_k (p_k - u))
      const double c0 = -2.0 * a0 / e0 * invS * (p0 - u);
      const double c1 = -2.0 * a1 / e1 * invS * (p1 - u);
      const double c2 = -2.0 * a2 / e2 * invS * (p2 - u);
      gNeg[0] = c0 * (k136 * t0 + 0.5 * s0) + w0 * (1.0 / 3.0);
      gNeg[1] = c0 * (-k133 * t0 - 2.0 * s0) + c1 * (k136 * t1 + 0.5 * s1) + w0 * (-7.0 / 6.0) + w1 * (-1.0 / 6.0);
      gNeg[2] = c0 * (k136 * t0 + 1.5 * s0) + c1 * (-k133 * t1) + c2 * (k136 * t2 + 1.5 * s2)
              + w0 * (11.0 / 6.0) + w1 * (5.0 / 6.0) + w2 * (1.0 / 3.0);
      gNeg[3] = c1 * (k136 * t1 - 0.5 * s1) + c2 * (-k133 * t2 - 2.0 * s2) + w1 * (1.0 / 3.0) + w2 * (5.0 / 6.0);
      gNeg[4] = c2 * (k136 * t2 + 0.5 * s2) + w2 * (-1.0 / 6.0);
      gNeg[5] = 0.0;
      uNeg = u;
    }
    {
      // stencil (b,c,d,e,f)
      const double p0 = (-b + 5.0 * c + 2.0 * d) * s6;
      const double p1 = (2.0 * c + 5.0 * d - e) * s6;
      const double p2 = (11.0 * d - 7.0 * e + 2.0 * f) * s6;
      const double t0 = b - 2.0 * c + d, s0 = b - 4.0 * c + 3.0 * d;
      const double t1 = c - 2.0 * d + e, s1 = c - e;
      const double t2 = d - 2.0 * e + f, s2 = 3.0 * d - 4.0 * e + f;
      const double e0 = kWenoEps + (k13 * t0 * t0 + 0.25 * s0 * s0);
      const double e1 = kWenoEps + (k13 * t1 * t1 + 0.25 * s1 * s1);
      const double e2 = kWenoEps + (k13 * t2 * t2 + 0.25 * s2 * s2);
      const double a0 = 0.3 / (e0 * e0), a1 = 0.6 / (e1 * e1), a2 = 0.1 / (e2 * e2);
      const double invS = 1.0 / (a0 + a1 + a2);
      const double w0 = a0 * invS, w1 = a1 * invS, w2 = a2 * invS;
      const double u = w0 * p0 + w1 * p1 + w2 * p2;
      const double c0 = -2.0 * a0 / e0 * invS * (p0 - u);
      const double c1 = -2.0 * a1 / e1 * invS * (p1 - u);
      const double c2 = -2.0 * a2 / e2 * invS * (p2 - u);
      gPos[0] = 0.0;
      gPos[1] = c0 * (k136 * t0 + 0.5 * s0) + w0 * (-1.0 / 6.0);
      gPos[2] = c0 * (-k133 * t0 - 2.0 * s0) + c1 * (k136 * t1 + 0.5 * s1) + w0 * (5.0 / 6.0) + w1 * (1.0 / 3.0);
      gPos[3] = c0 * (k136 * t0 + 1.5 * s0) + c1 * (-k133 * t1) + c2 * (k136 * t2 + 1.5 * s2)
              + w0 * (1.0 / 3.0) + w1 * (5.0 / 6.0) + w2 * (11.0 / 6.0);
      gPos[4] = c1 * (k136 * t1 - 0.5 * s1) + c2 * (-k133 * t2 - 2.0 * s2) + w1 * (-1.0 / 6.0) + w2 * (-7.0 / 6.0);
      gPos[5] = c2 * (k136 * t2 + 0.5 * s2) + w2 * (1.0 / 3.0);
      uPos = u;
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Euler, Rusanov flux with Roe-averaged wave speed.  DIM = 1,2,3 -> 3,4,5 dofs (rho, rho*vel[DIM], rho*E).
// Reference: impl/euler_rusanov_flux_values_function.hpp:54-208 (values),
//            impl/euler_rusanov_flux_jacobian_function.hpp:54-406 (Jacobians incl. d(smax)/dq).
// The face normal is the unit vector of axis AX, so un = vel[AX].
// ---------------------------------------------------------------------------------------------------------------
template <int DIM>
struct Euler {
  static constexpr int dim = DIM;
  static constexpr int ndpc = DIM + 2;
  double gamma;

  template <int AX>
  PDA_DEVFN void flux(const double* qL, const double* qR, double* F) const {
    const double gm1 = gamma - 1.0;
    const double rL = qL[0], rR = qR[0];
    // u = q/(r + es) and H = (E+p)/r use different denominators in the reference; r + 1e-30 == r for every
    // representable density above ~1e-14, so the second reciprocal is only taken when it differs.
    const double rLe = rL + kEs, rRe = rR + kEs;
    const double iLe = 1.0 / rLe, iRe = 1.0 / rRe;
    const double iL = (rLe == rL) ? iLe : 1.0 / rL;
    const double iR = (rRe == rR) ? iRe : 1.0 / rR;
    double vL[DIM], vR[DIM];
    double kL = 0.0, kR = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      vL[m] = qL[1 + m] * iLe; vR[m] = qR[1 + m] * iRe;
      kL += vL[m] * vL[m]; kR += vR[m] * vR[m];
    }
    const double pL = gm1 * (qL[DIM + 1] - 0.5 * rL * kL);
    const double pR = gm1 * (qR[DIM + 1] - 0.5 * rR * kR);
    const double HL = (qL[DIM + 1] + pL) * iL;
    const double HR = (qR[DIM + 1] + pR) * iR;
    const double mL = rL * vL[AX], mR = rR * vR[AX];   // rho * un

    const double RT = sqrt(rR * iL);
    const double iRT = 1.0 / (1.0 + RT);
    double k = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) { const double v = (vL[m] + RT * vR[m]) * iRT; k += v * v; }
    const double H = (HL + RT * HR) * iRT;
    const double a = sqrt(gm1 * (H - 0.5 * k));
    const double smax = sqrt(k) + a;   // 1D: sqrt(u*u) == |u|

    F[0] = 0.5 * (mL + mR + smax * (qL[0] - qR[0]));
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      const double pl = (m == AX) ? pL : 0.0, pr = (m == AX) ? pR : 0.0;
      F[1 + m] = 0.5 * ((mL * vL[m] + pl) + (mR * vR[m] + pr) + smax * (qL[1 + m] - qR[1 + m]));
    }
    F[DIM + 1] = 0.5 * (mL * HL + mR * HR + smax * (qL[DIM + 1] - qR[DIM + 1]));
  }

  // JL = dF/dqL, JR = dF/dqR, row-major [ndpc][ndpc]
  template <int AX>
  PDA_DEVFN void fluxJac(const double* qL, const double* qR, double* JL, double* JR) const {
    constexpr int N = ndpc;
    const double gm1 = gamma - 1.0;
    const double rL = qL[0], rR = qR[0];
    const double iLe = 1.0 / (rL + kEs), iRe = 1.0 / (rR + kEs);
    double vL[DIM], vR[DIM], v[DIM];
    double kL = 0.0, kR = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      vL[m] = qL[1 + m] * iLe; vR[m] = qR[1 + m] * iRe;
      kL += vL[m] * vL[m]; kR += vR[m] * vR[m];
    }
    const double pL = gm1 * (qL[DIM + 1] - 0.5 * rL * kL);
    const double pR = gm1 * (qR[DIM + 1] - 0.5 * rR * kR);
    const double HL = (qL[DIM + 1] + pL) / rL;
    const double HR = (qR[DIM + 1] + pR) / rR;
    const double aL2 = gm1 * (HL - 0.5 * kL);   // aL*aL (the reference squares a sqrt)
    const double aR2 = gm1 * (HR - 0.5 * kR);
    const double unL = vL[AX], unR = vR[AX];

    const double r = sqrt(rR * rL);
    const double RT = sqrt(rR / rL);
    const double iRT = 1.0 / (1.0 + RT);
    double k = 0.0, dotL = 0.0, dotR = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      v[m] = (vL[m] + RT * vR[m]) * iRT;
      k += v[m] * v[m]; dotL += vL[m] * v[m]; dotR += vR[m] * v[m];
    }
    const double H = (HL + RT * HR) * iRT;
    const double a = sqrt(gm1 * (H - 0.5 * k));
    const double smax = sqrt(k) + a;
    const double vmag2 = k + kEs;
    const double ivmag = 1.0 / sqrt(vmag2);
    const double ia = 1.0 / a;

    double gL[N], gR[N];
    const double iLr = 1.0 / (rL + r), iRr = 1.0 / (rR + r);
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
      gL[0] = iLr * (sL + 0.5 * gm1 * ia * (0.5 * (vmag2 + dotL) + 0.5 * (HL - H) - aL2 / gm1 + 0.5 * (gamma - 2.0) * kL));
      gR[0] = iRr * (sR + 0.5 * gm1 * ia * (0.5 * (vmag2 + dotR) + 0.5 * (HR - H) - aR2 / gm1 + 0.5 * (gamma - 2.0) * kR));
      gL[N - 1] = 0.5 * iLr * gamma * gm1 * ia;
      gR[N - 1] = 0.5 * iRr * gamma * gm1 * ia;
    }

    fillPhysical<AX>(JL, vL, kL, HL, unL, gm1);
    fillPhysical<AX>(JR, vR, kR, HR, unR, gm1);
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

 private:
  // 0.5 * d(F_phys . n)/dq  for one side
  template <int AX>
  PDA_DEVFN void fillPhysical(double* J, const double* vel, double k2, double H, double un, double gm1) const {
    constexpr int N = ndpc;
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
};

// ---------------------------------------------------------------------------------------------------------------
// 2D shallow water (h, hu, hv), Rusanov.  impl/swe_rusanov_flux_values_function.hpp:54-97,
// impl/swe_rusanov_flux_jacobian_function.hpp:54-136; Coriolis source impl/swe_2d_prob_class.hpp:984-1012.
// ---------------------------------------------------------------------------------------------------------------
struct Swe2d {
  static constexpr int dim = 2;
  static constexpr int ndpc = 3;
  double g;
  double coriolis;

  template <int AX>
  PDA_DEVFN void flux(const double* qL, const double* qR, double* F) const {
    const double hL = qL[0], hR = qR[0];
    const double iL = 1.0 / (hL + kEs), iR = 1.0 / (hR + kEs);
    const double uL = qL[1] * iL, vL = qL[2] * iL;
    const double uR = qR[1] * iR, vR = qR[2] * iR;
    const double unL = (AX == 0) ? uL : vL, unR = (AX == 0) ? uR : vR;
    const double pL = 0.5 * g * hL * hL, pR = 0.5 * g * hR * hR;
    const double sL = sqrt(hL), sR = sqrt(hR);
    const double hm = 0.5 * (hL + hR);
    const double um = (unL * sL + unR * sR) / (sL + sR + kEs);
    const double smax = fabs(um) + sqrt(g * hm);
    const double mL = hL * unL, mR = hR * unR;
    F[0] = 0.5 * (mL + mR + smax * (qL[0] - qR[0]));
    F[1] = 0.5 * ((mL * uL + (AX == 0 ? pL : 0.0)) + (mR * uR + (AX == 0 ? pR : 0.0)) + smax * (qL[1] - qR[1]));
    F[2] = 0.5 * ((mL * vL + (AX == 1 ? pL : 0.0)) + (mR * vR + (AX == 1 ? pR : 0.0)) + smax * (qL[2] - qR[2]));
  }

  template <int AX>
  PDA_DEVFN void fluxJac(const double* qL, const double* qR, double* JL, double* JR) const {
    constexpr double nx = (AX == 0) ? 1.0 : 0.0, ny = (AX == 1) ? 1.0 : 0.0;
    const double hL = qL[0], hR = qR[0];
    const double uL = qL[1] / (hL + kEs), vL = qL[2] / (hL + kEs);
    const double uR = qR[1] / (hR + kEs), vR = qR[2] / (hR + kEs);
    const double unL = uL * nx + vL * ny, unR = uR * nx + vR * ny;
    const double hm = 0.5 * (hL + hR);
    const double sL = sqrt(hL), sR = sqrt(hR);
    const double um = (unL * sL + unR * sR) / (sL + sR + kEs);
    const double smax = fabs(um) + fabs(sqrt(g * hm));
    const double termL = (nx * qL[1] + ny * qL[2]) / (qL[0] * qL[0]);
    const double termR = (nx * qR[1] + ny * qR[2]) / (qR[0] * qR[0]);
    const double hsun = sL * unL + sR * unR + kEs;
    const double ahs = fabs(hsun);
    const double ss = sL + sR;
    const double gterm = g / (2.8284271247461903 /* 2^(3/2) */ * sqrt(g * (hL + hR)));
    double dL[3], dR[3];
    dL[0] = -ahs / (2.0 * sL * ss * ss) + (0.5 * unL / sL - sL * termL) * hsun / (ss * ahs) + gterm;
    dL[1] = nx * hsun / (sL * ss * ahs);
    dL[2] = ny * hsun / (sL * ss * ahs);
    dR[0] = -ahs / (2.0 * sR * ss * ss) + (0.5 * unR / sR - sR * termR) * hsun / (ss * ahs) + gterm;
    dR[1] = nx * hsun / (sR * ss * ahs);
    dR[2] = ny * hsun / (sR * ss * ahs);
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
};


// ---------------------------------------------------------------------------------------------------------------
// Extra (non-flux) terms of a problem.  The flux families above have none; the advection-diffusion families add a
// first-order central diffusion term per axis, dD[ax] * (u_{+1} - 2u + u_{-1}) with dD = D * hInv^2
// (advection_diffusion_2d_prob_class.hpp:1167-1201), and point terms (reaction, source) evaluated at the cell.
// ---------------------------------------------------------------------------------------------------------------
template <class Phys> struct PhysTraits {
  static constexpr bool hasDiffusion = false;   // per-axis second-difference term
};

// 2D Burgers (u, v), Rusanov.  impl/advection_diffusion_2d_flux_functions.hpp:54-114.
struct Burgers2d {
  static constexpr int dim = 2;
  static constexpr int ndpc = 2;
  double dD[2];   // diffusion * dxInv^2, diffusion * dyInv^2

  template <int AX>
  PDA_DEVFN void flux(const double* qL, const double* qR, double* F) const {
    constexpr double n0 = (AX == 0) ? 1.0 : 0.0, n1 = (AX == 1) ? 1.0 : 0.0;
    const double a0 = fmax(fabs(qL[0]), fabs(qR[0]));
    const double a1 = fmax(fabs(qL[1]), fabs(qR[1]));
    double f0 = a0 * (qL[0] - qR[0]);
    f0 += n0 * (qL[0] * qL[0] + qR[0] * qR[0]);
    f0 += n1 * (qL[0] * qL[1] + qR[0] * qR[1]);
    double f1 = a1 * (qL[1] - qR[1]);
    f1 += n0 * (qL[0] * qL[1] + qR[0] * qR[1]);
    f1 += n1 * (qL[1] * qL[1] + qR[1] * qR[1]);
    F[0] = f0 * 0.25;
    F[1] = f1 * 0.25;
  }

  template <int AX>
  PDA_DEVFN void fluxJac(const double* qL, const double* qR, double* JL, double* JR) const {
    constexpr double n0 = (AX == 0) ? 1.0 : 0.0, n1 = (AX == 1) ? 1.0 : 0.0;
    if (fabs(qL[0]) > fabs(qR[0])) {
      JL[0] = (2.0 * qL[0] - qR[0]) * copysign(1.0, qL[0]) + n0 * 2.0 * qL[0] + n1 * qL[1];
      JR[0] = n0 * 2.0 * qR[0] + n1 * qR[1] - fabs(qL[0]);
    } else {
      JL[0] = n0 * 2.0 * qL[0] + n1 * qL[1] + fabs(qR[0]);
      JR[0] = (qL[0] - 2.0 * qR[0]) * copysign(1.0, qR[0]) + n0 * 2.0 * qR[0] + n1 * qR[1];
    }
    JL[0] *= 0.25; JR[0] *= 0.25;
    if (fabs(qL[1]) > fabs(qR[1])) {
      JL[3] = (2.0 * qL[1] - qR[1]) * copysign(1.0, qL[1]) + n0 * qL[0] + n1 * 2.0 * qL[1];
      JR[3] = n0 * qR[0] + n1 * 2.0 * qR[1] - fabs(qL[1]);
    } else {
      JL[3] = n0 * qL[0] + n1 * 2.0 * qL[1] + fabs(qR[1]);
      JR[3] = (qL[1] - 2.0 * qR[1]) * copysign(1.0, qR[1]) + n0 * qR[0] + n1 * 2.0 * qR[1];
    }
    JL[3] *= 0.25; JR[3] *= 0.25;
    JL[1] = n1 * qL[0] * 0.25; JL[2] = n0 * qL[1] * 0.25;
    JR[1] = n1 * qR[0] * 0.25; JR[2] = n0 * qR[1] * 0.25;
  }
};
template <> struct PhysTraits<Burgers2d> { static constexpr bool hasDiffusion = true; };

// Linear advection of one scalar with constant velocity a: the reference's "Rusanov" branch is the upwind flux of a
// POSITIVE velocity, F = a * uNeg (impl/advection_1d_mixins.hpp:79-94, advection_diffusion_reaction_2d_flux_mixin.hpp).
//   DIM 1: Advection1d::PeriodicLinear (no extra terms: dD = sigma = 0, no source);
//   DIM 2: AdvectionDiffusionReaction2d::ProblemA: + D lap(u) - sigma u + f(x,y,t)
//          (advection_diffusion_reaction_2d_prob_class.hpp:485-512); f is a per-cell table (device pointer, indexed
//          by sample-mesh row) or the constant srcConst when the table is null (default source: 1).
template <int DIM>
struct LinAdv {
  static constexpr int dim = DIM;
  static constexpr int ndpc = 1;
  double a[DIM];
  double dD[DIM];
  double sigma;
  double srcConst;
  const double* srcTable;

  template <int AX>
  PDA_DEVFN void flux(const double* qL, const double* /*qR*/, double* F) const { F[0] = qL[0] * a[AX]; }
  template <int AX>
  PDA_DEVFN void fluxJac(const double* /*qL*/, const double* /*qR*/, double* JL, double* JR) const {
    JL[0] = a[AX];
    JR[0] = 0.0;
  }
};
template <> struct PhysTraits<LinAdv<2>> { static constexpr bool hasDiffusion = true; };

}  // namespace dev
}  // namespace pda
