// kernels_applymarch2d.cuh -- matrix-free J*b of the 2D families on full lattices for ONE contiguous operand column (vector
// operands, columns of a column-major operand), on the skeleton of the 2D velocity kernel (kernels_march2d.cuh): a warp
// owns a strip of 32-2h columns and marches along y with a REGISTER ring of 2h+1 rows of the state AND of the operand;
// the x stencil comes from the neighbouring lanes by shuffle; every face once: the front y-face tangent flux of row j is
// the back one of row j+1, a lane computes the left face of its cell and receives the right one from lane+1.
// Per face: reconstruction values + gradients in one pass, two dot products with the operand stencil, and the tangent
// of the Rusanov flux (Euler: one closed-form Jacobian-vector product, eulerFluxJvpFast; shallow water, Burgers,
// advection-diffusion-reaction: their small flux Jacobians times the directional derivatives, plus the point / diffusion
// terms applied to the operand) -- the velocity kernel with (value, tangent) pairs.  Replaces Eigen's J * b (adapter_cpp.hpp:231-259) for the Newton-Krylov J*v; the tile kernel
// k_applyjac_lattice2d (kernels_applylattice.cuh) keeps the other families and row-major multi-column operands.
#pragma once
#include "kernels_march2d.cuh"

namespace pda {
namespace dev {

// tangent of the face flux along AX: closed-form Jacobian-vector product for Euler, flux Jacobians times the two
// directional derivatives for the small systems (9 / 4 / 1 entries per side)
template <class Phys, int AX>
PDA_DEVFN void faceFluxTangent(const Phys& phys, const double* uN, const double* uP, const double* dN, const double* dP,
                               double* D) {
  constexpr int N = Phys::ndpc;
  if constexpr (std::is_same<Phys, Euler<2>>::value) {
    double a[1][4], b[1][4], out[1][4];
#pragma unroll
    for (int d = 0; d < 4; ++d) { a[0][d] = dN[d]; b[0][d] = dP[d]; }
    eulerFluxJvpFast<2, AX, 1>(phys.gamma, uN, uP, a, b, out);
#pragma unroll
    for (int d = 0; d < 4; ++d) D[d] = out[0][d];
  } else {
    double JN[N * N], JP[N * N];
    faceFluxJac2d<Phys, AX>(phys, uN, uP, JN, JP);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += JN[k * N + j] * dN[j] + JP[k * N + j] * dP[j];
      D[k] = s;
    }
  }
}

// tangent flux of one face from the 2h stencil values of the state (q) and of the operand (b) per dof
template <class Phys, int S, int AX>
PDA_DEVFN void faceTangentOf(const Phys& phys, const double (*q)[Phys::ndpc], const double (*b)[Phys::ndpc], double* D) {
  constexpr int N = Phys::ndpc;
  constexpr int M = S - 1;
  double uN[N], uP[N], dN[N], dP[N];
#pragma unroll
  for (int d = 0; d < N; ++d) {
    double s[M], gN[M], gP[M];
#pragma unroll
    for (int o = 0; o < M; ++o) s[o] = q[o][d];
    reconFaceValGradFast<S>(s, uN[d], uP[d], gN, gP);
    double sN = 0.0, sP = 0.0;
#pragma unroll
    for (int o = 0; o < M; ++o) { sN = fma(gN[o], b[o][d], sN); sP = fma(gP[o], b[o][d], sP); }
    dN[d] = sN; dP[d] = sP;
  }
  faceFluxTangent<Phys, AX>(phys, uN, uP, dN, dP, D);
}

template <class Phys, int S>
__global__ void __launch_bounds__(128, 2)
k_applyjac_march2d(Phys phys, LatticeDesc L, Deltas dl, const double* __restrict__ U, const double* __restrict__ B,
                   double* __restrict__ Rout, int LY) {
  constexpr int N = Phys::ndpc;
  constexpr int h = (S - 1) / 2;
  constexpr int W = 32 - 2 * h;
  constexpr int R = 2 * h + 1;
  constexpr int M = 2 * h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nx = L.n[0], ny = L.n[1];
  const int lo0 = L.per[0] ? 0 : L.meshHalo, hi0 = L.per[0] ? nx : nx - L.meshHalo;
  const int yb = L.per[1] ? 0 : L.meshHalo, ye = L.per[1] ? ny : ny - L.meshHalo;
  const int nStrips = (hi0 - lo0 + W - 1) / W;
  const int wid = blockIdx.x * 4 + warp;
  const int strip = wid % nStrips, chunk = wid / nStrips;
  const int j0 = yb + chunk * LY;
  if (j0 >= ye) return;
  const int j1 = min(j0 + LY, ye);
  const int x = lo0 + strip * W - h + lane;
  int xc = x;
  if (L.per[0]) { xc %= nx; if (xc < 0) xc += nx; }
  else xc = (xc < 0) ? 0 : (xc >= nx ? nx - 1 : xc);
  const bool outLane = (lane >= h) && (lane <= 31 - h) && (x < hi0);

  auto rowOff = [&](int r) -> int64_t {
    if (L.per[1]) { r %= ny; if (r < 0) r += ny; }
    else r = (r < 0) ? 0 : (r >= ny ? ny - 1 : r);
    return ((int64_t)r * nx + xc) * N;
  };
  double q[R][N], b[R][N];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int64_t off = rowOff(j0 - 1 - h + i);
    loadCell<N>(U + off, q[i]);
    loadCell<N>(B + off, b[i]);
  }
  double DyB[N];
#pragma unroll
  for (int d = 0; d < N; ++d) DyB[d] = 0.0;

  for (int j = j0 - 1; j < j1; ++j) {
    const bool ghost = (j < j0);
    double nq[N], nb[N];
    const bool more = (j + 1 < j1);
    if (more) {
      const int64_t off = rowOff(j + 1 + h);
      loadCell<N>(U + off, nq);
      loadCell<N>(B + off, nb);
    }
    // ---- front y face (j+1/2): rows j-h+1 .. j+h
    double DyF[N];
    faceTangentOf<Phys, S, 1>(phys, q + 1, b + 1, DyF);
    if (!ghost) {
      // ---- x left face of this lane's cell from the neighbouring lanes' row-j values
      double sq[M][N], sb[M][N], Dx[N];
#pragma unroll
      for (int o = 0; o < M; ++o)
#pragma unroll
        for (int d = 0; d < N; ++d) {
          sq[o][d] = (o == h) ? q[h][d] : __shfl_sync(0xffffffffu, q[h][d], (lane + o - h) & 31);
          sb[o][d] = (o == h) ? b[h][d] : __shfl_sync(0xffffffffu, b[h][d], (lane + o - h) & 31);
        }
      faceTangentOf<Phys, S, 0>(phys, sq, sb, Dx);
      double v[N];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const double DxR = __shfl_down_sync(0xffffffffu, Dx[d], 1);
        v[d] = dl.hInv[0] * (Dx[d] - DxR);
        v[d] += dl.hInv[1] * (DyB[d] - DyF[d]);
      }
      if constexpr (std::is_same<Phys, Swe2d>::value || PhysTraits<Phys>::hasDiffusion || std::is_same<Phys, LinAdv<2>>::value) {
        // point terms and diffusion: the entries addExtraJacInner adds to J, applied to the operand.  The "slot" handed to
        // it is the graph column itself (0 self, 1 left, 2 front, 3 right, 4 back); the operand values of those cells
        // sit in the ring (front / back) and in the neighbouring lanes (left / right)
        double nbv[5][N];
#pragma unroll
        for (int d = 0; d < N; ++d) {
          nbv[0][d] = b[h][d];
          nbv[1][d] = __shfl_sync(0xffffffffu, b[h][d], (lane - 1) & 31);
          nbv[2][d] = b[h + 1][d];
          nbv[3][d] = __shfl_sync(0xffffffffu, b[h][d], (lane + 1) & 31);
          nbv[4][d] = b[h - 1][d];
        }
        uint8_t ident[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) ident[c] = (uint8_t)c;
        addExtraJacInner<Phys>(phys, q[h], ident, [&](int k, int col, int jj, double xv) { v[k] += xv * nbv[col][jj]; });
      }
      if (outLane) storeBlockRow<N>(Rout + ((int64_t)j * nx + x) * N, v);
    }
#pragma unroll
    for (int d = 0; d < N; ++d) DyB[d] = DyF[d];
#pragma unroll
    for (int i = 0; i < R - 1; ++i)
#pragma unroll
      for (int d = 0; d < N; ++d) { q[i][d] = q[i + 1][d]; b[i][d] = b[i + 1][d]; }
    if (more) {
#pragma unroll
      for (int d = 0; d < N; ++d) { q[R - 1][d] = nq[d]; b[R - 1][d] = nb[d]; }
    }
  }
}

}  // namespace dev

// PDA_APPLY2D_MARCH=0 keeps the tile kernel (A/B measurements)
inline bool applyMarch2dEnabled() {
  static const bool on = [] { const char* e = std::getenv("PDA_APPLY2D_MARCH"); return !(e && e[0] == '0'); }();
  return on;
}

template <class Phys, int S>
void launchApplyMarch2d(const Phys& phys, const dev::LatticeDesc& L, const dev::Deltas& dl, const double* dU, const double* dB,
                             double* dR, cudaStream_t st) {
  constexpr int h = (S - 1) / 2, W = 32 - 2 * h;
  const int lo0 = L.per[0] ? 0 : L.meshHalo, hi0 = L.per[0] ? L.n[0] : L.n[0] - L.meshHalo;
  const int yb = L.per[1] ? 0 : L.meshHalo, ye = L.per[1] ? L.n[1] : L.n[1] - L.meshHalo;
  if (hi0 <= lo0 || ye <= yb) return;
  const int64_t nStrips = (hi0 - lo0 + W - 1) / W;
  int LY = 64;
  while (LY > 8 && nStrips * ((ye - yb + LY - 1) / LY) < (int64_t)148 * 8 * 4) LY /= 2;
  const int64_t tasks = nStrips * ((ye - yb + LY - 1) / LY);
  dev::k_applyjac_march2d<Phys, S><<<(unsigned)((tasks + 3) / 4), 128, 0, st>>>(phys, L, dl, dU, dB, dR, LY);
}

}  // namespace pda
