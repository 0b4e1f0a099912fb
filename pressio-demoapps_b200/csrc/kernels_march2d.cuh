// kernels_march2d.cuh -- structured 2D velocity with every face flux computed ONCE, registers and shuffles only.
//
// Replaces the reference's inner-cell velocity loops on full 2D meshes (euler_2d_prob_class.hpp:991-1044,
// swe_2d_prob_class.hpp:928-981, advection_diffusion_2d_prob_class.hpp inner loop), which evaluate both faces of
// both axes per cell (each face flux twice).
//
// A warp owns a strip of 32 columns and MARCHES along y:
//   * lane l holds column x0-h+l; the y stencil of that column is a register ring of 2h+1 rows (rows j-h..j+h), one
//     new row per step fetched one step ahead with a single coalesced (AoS, 256-bit for 4 dofs) load per lane;
//   * the front y-face flux of row j is the back flux of row j+1 (registers) -> y faces once;
//   * the x stencil of row j comes from the neighbouring lanes by warp shuffle; each lane computes the LEFT face of
//     its cell, the right face is lane+1's left face (one more shuffle) -> x faces once.  The 2h edge lanes of a warp
//     only feed stencils (32-2h output cells per warp and step);
//   * no shared memory, no barrier; V written once per cell, coalesced.
// Euler uses the branch-free reciprocal / square root arithmetic of the 3D kernel (kernels_tiled.cuh).
// Works for scheme stencils 3/5/7, periodic or not per axis (cells whose MESH stencil leaves a non-periodic domain
// belong to the near-boundary kernel), diffusion / point terms of the advection-diffusion families, and slab-local
// storage (2D multi-GPU slabs).
#pragma once
#include "fastmath.cuh"
#include "cellmath.cuh"
#include "kernels_lattice.cuh"
#include "kernels_jaclattice.cuh"

namespace pda {
namespace dev {

// First-order instances of the small systems are latency-bound at 16 warps/SM (ncu, SWE 4096^2: FP64 pipe 52 %, stall
// "wait" 2.5 of 6 cycles per issue, DRAM 2 TB/s): they fit 80 registers -> 24 warps/SM
template <class Phys, int S> struct March2dOcc { static constexpr int minCtas = (S == 7) ? 4 : ((Phys::ndpc <= 3) ? (S == 3 ? 6 : 5) : 1); };

template <class Phys, int S, int CELLMASK>
__global__ void __launch_bounds__(128, March2dOcc<Phys, S>::minCtas)
k_velocity_march2d(Phys phys, LatticeDesc L, Deltas dl, const double* __restrict__ U, double* __restrict__ V, int LY) {
  constexpr int N = Phys::ndpc;
  constexpr int h = (S - 1) / 2;
  constexpr int W = 32 - 2 * h;      // output cells per warp and step
  constexpr int R = 2 * h + 1;       // ring rows j-h .. j+h
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nx = L.n[0], ny = L.n[1];
  const int lo0 = L.per[0] ? 0 : L.meshHalo, hi0 = L.per[0] ? nx : nx - L.meshHalo;
  int yb = L.planeBegin, ye = L.planeEnd;
  if (!L.slab && !L.per[1]) { yb = max(yb, L.meshHalo); ye = min(ye, ny - L.meshHalo); }
  const int nStrips = (hi0 - lo0 + W - 1) / W;
  const int wid = blockIdx.x * 4 + warp;
  const int strip = wid % nStrips, chunk = wid / nStrips;
  const int j0 = yb + chunk * LY;
  if (j0 >= ye) return;
  const int j1 = min(j0 + LY, ye);
  const int x = lo0 + strip * W - h + lane;
  int xc = x;
  if (L.per[0]) { xc %= nx; if (xc < 0) xc += nx; }   // edge lanes of tiny periodic meshes may wrap more than once
  else xc = (xc < 0) ? 0 : (xc >= nx ? nx - 1 : xc);
  const bool outLane = (lane >= h) && (lane <= 31 - h) && (x < hi0);

  auto rowPtr = [&](int r) -> const double* {
    int sr;
    if (L.slab) sr = max(r + L.haloPlanes, 0);   // the ghost step's lowest row is loaded but never used
    else if (L.per[1]) { sr = r % ny; if (sr < 0) sr += ny; }
    else sr = (r < 0) ? 0 : (r >= ny ? ny - 1 : r);
    return U + ((int64_t)sr * nx + xc) * N;
  };
  auto yFace = [&](const double (*q)[N], double* F) {   // q[0..2h-1] = rows around the face
    double uN[N], uP[N];
#pragma unroll
    for (int d = 0; d < N; ++d) {
      double s[2 * h];
#pragma unroll
      for (int o = 0; o < 2 * h; ++o) s[o] = q[o][d];
      reconFaceFast<S>(s, uN[d], uP[d]);
    }
    faceFlux2d<Phys, 1>(phys, uN, uP, F);
  };

  // the march starts one row early ("ghost" step: only the y face below the first row), so that EVERY y face comes
  // from the same code -> a row's result does not depend on where chunks / slabs are cut (bit-exact decompositions)
  // shallow water, first order: the face states are cell values, so 1/h and sqrt(h) are computed once per cell (when its
  // row enters the ring) instead of once per adjacent face: the kernel is FP64-bound (ncu: pipe 61 %, DRAM 2.3 TB/s)
  constexpr bool kSwePre = std::is_same<Phys, Swe2d>::value && S == 3;
  double q[R][N];
  double pre[R][2];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    loadCell<N>(rowPtr(j0 - 1 - h + i), q[i]);
    if constexpr (kSwePre) { pre[i][0] = rcpFast(q[i][0]); pre[i][1] = sqrtFast(q[i][0]); }
  }
  (void)pre;
  double FyB[N];
#pragma unroll
  for (int d = 0; d < N; ++d) FyB[d] = 0.0;
  // WENO3 / WENO5: CELL-based edge values (cellmath.cuh) along both axes -- the smoothness indicators of a cell serve its
  // two edges, so they are computed once per cell instead of once per adjacent face (57 instead of 76 FP64 instructions
  // per dof for WENO5).  y: the ring holds the whole stencil of cell j+1 at step j: face j+1/2 = (eR(j) carried, eL(j+1));
  // x: a lane computes (eL, eR) of its own cell from the neighbouring lanes, the left face is (eR of lane-1, own eL).
  constexpr bool kCellY = (S >= 5) && (CELLMASK & 1), kCellX = (S >= 5) && (CELLMASK & 2);
  double eRc[N];   // front-edge value of the current row's cell (carried from the previous step)
  if constexpr (kCellY) {
#pragma unroll
    for (int d = 0; d < N; ++d) {
      double s[2 * h - 1], eLx;
#pragma unroll
      for (int o = 0; o < 2 * h - 1; ++o) s[o] = q[1 + o][d];     // cell j0-1: rows j0-h .. j0+h-2
      cellEdgesFast2<S>(s, eLx, eRc[d]);
    }
  }
  (void)eRc;

  for (int j = j0 - 1; j < j1; ++j) {
    const bool ghost = (j < j0);
    double nxt[N];
    const bool more = (j + 1 < j1);
    if (more) loadCell<N>(rowPtr(j + 1 + h), nxt);   // lands while this row is computed
    // ---- front y face (j+1/2): rows j-h+1 .. j+h
    double FyF[N];
    if constexpr (kSwePre) sweFluxFastPre<1>(phys.g, q[1], q[2], pre[1][0], pre[1][1], pre[2][0], pre[2][1], FyF);
    else if constexpr (kCellY) {
      double uN[N], uP[N];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        double s[2 * h - 1], eR1;
#pragma unroll
        for (int o = 0; o < 2 * h - 1; ++o) s[o] = q[2 + o][d];   // cell j+1: rows j-h+2 .. j+h
        cellEdgesFast2<S>(s, uP[d], eR1);
        uN[d] = eRc[d];
        eRc[d] = eR1;
      }
      faceFlux2d<Phys, 1>(phys, uN, uP, FyF);
    } else yFace(q + 1, FyF);
    if (ghost) {
#pragma unroll
      for (int d = 0; d < N; ++d) FyB[d] = FyF[d];
#pragma unroll
      for (int i = 0; i < R - 1; ++i) {
#pragma unroll
        for (int d = 0; d < N; ++d) q[i][d] = q[i + 1][d];
        if constexpr (kSwePre) { pre[i][0] = pre[i + 1][0]; pre[i][1] = pre[i + 1][1]; }
      }
      if (more) {
#pragma unroll
        for (int d = 0; d < N; ++d) q[R - 1][d] = nxt[d];
        if constexpr (kSwePre) { pre[R - 1][0] = rcpFast(nxt[0]); pre[R - 1][1] = sqrtFast(nxt[0]); }
      }
      continue;
    }
    // ---- x left face of this lane's cell from the neighbouring lanes' row-j values
    double Fx[N];
    if constexpr (kSwePre) {
      double qL[N];
#pragma unroll
      for (int d = 0; d < N; ++d) qL[d] = __shfl_sync(0xffffffffu, q[h][d], (lane - 1) & 31);
      const double iLft = __shfl_sync(0xffffffffu, pre[h][0], (lane - 1) & 31);
      const double sLft = __shfl_sync(0xffffffffu, pre[h][1], (lane - 1) & 31);
      sweFluxFastPre<0>(phys.g, qL, q[h], iLft, sLft, pre[h][0], pre[h][1], Fx);
    } else if constexpr (kCellX) {
      double uN[N], uP[N];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        double s[2 * h - 1], eRx;
#pragma unroll
        for (int o = 0; o < 2 * h - 1; ++o)
          s[o] = (o == h - 1) ? q[h][d] : __shfl_sync(0xffffffffu, q[h][d], (lane + o - (h - 1)) & 31);
        cellEdgesFast2<S>(s, uP[d], eRx);                 // own cell: left-edge value = uPos of my left face
        uN[d] = __shfl_up_sync(0xffffffffu, eRx, 1);     // right-edge value of the cell to the left
      }
      faceFlux2d<Phys, 0>(phys, uN, uP, Fx);
    } else {
      double uN[N], uP[N];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        double s[2 * h];
#pragma unroll
        for (int o = 0; o < 2 * h; ++o)
          s[o] = (o == h) ? q[h][d] : __shfl_sync(0xffffffffu, q[h][d], (lane + o - h) & 31);
        reconFaceFast<S>(s, uN[d], uP[d]);
      }
      faceFlux2d<Phys, 0>(phys, uN, uP, Fx);
    }
    double v[N];
#pragma unroll
    for (int d = 0; d < N; ++d) {
      const double FxR = __shfl_down_sync(0xffffffffu, Fx[d], 1);
      v[d] = dl.hInv[0] * (Fx[d] - FxR);
      v[d] += dl.hInv[1] * (FyB[d] - FyF[d]);
      FyB[d] = FyF[d];
    }
    if constexpr (PhysTraits<Phys>::hasDiffusion) {
      // dD[ax] * (u+ - 2u + u-), x then y, after the flux balances (advection_diffusion_2d_prob_class.hpp:1167-1201)
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const double uL = __shfl_sync(0xffffffffu, q[h][d], (lane - 1) & 31);
        const double uR = __shfl_sync(0xffffffffu, q[h][d], (lane + 1) & 31);
        v[d] += phys.dD[0] * (uR - 2.0 * q[h][d] + uL);
      }
#pragma unroll
      for (int d = 0; d < N; ++d) v[d] += phys.dD[1] * (q[h + 1][d] - 2.0 * q[h][d] + q[h - 1][d]);
    }
    if (outLane) {
      const int64_t vIdx = (int64_t)j * nx + x;
      addForcing<Phys>(phys, q[h], v, (int32_t)vIdx);
      storeBlockRow<N>(V + vIdx * N, v);
    }
    // ---- rotate the ring
#pragma unroll
    for (int i = 0; i < R - 1; ++i) {
#pragma unroll
      for (int d = 0; d < N; ++d) q[i][d] = q[i + 1][d];
      if constexpr (kSwePre) { pre[i][0] = pre[i + 1][0]; pre[i][1] = pre[i + 1][1]; }
    }
    if (more) {
#pragma unroll
      for (int d = 0; d < N; ++d) q[R - 1][d] = nxt[d];
      if constexpr (kSwePre) { pre[R - 1][0] = rcpFast(nxt[0]); pre[R - 1][1] = sqrtFast(nxt[0]); }
    }
  }
}

}  // namespace dev

template <class Phys, int S>
void launchMarch2d(const Phys& phys, const dev::LatticeDesc& L, const dev::Deltas& dl, const double* dU, double* dV,
                   cudaStream_t st) {
  static_assert(Phys::dim == 2, "2D kernel");
  constexpr int h = (S - 1) / 2, W = 32 - 2 * h;
  const int lo0 = L.per[0] ? 0 : L.meshHalo, hi0 = L.per[0] ? L.n[0] : L.n[0] - L.meshHalo;
  int yb = L.planeBegin, ye = L.planeEnd;
  if (!L.slab && !L.per[1]) { yb = std::max(yb, L.meshHalo); ye = std::min(ye, L.n[1] - L.meshHalo); }
  if (hi0 <= lo0 || ye <= yb) return;
  const int64_t nStrips = (hi0 - lo0 + W - 1) / W;
  // y chunks: long enough to amortise the ring fill (2h+1 rows, one extra y face), short enough for >= 4 waves
  int LY = 64;
  while (LY > 8 && nStrips * ((ye - yb + LY - 1) / LY) < (int64_t)148 * 16 * 4) LY /= 2;
  const int64_t tasks = nStrips * ((ye - yb + LY - 1) / LY);
  if constexpr (std::is_same<Phys, dev::Euler<2>>::value && S >= 5) {
    // 2D Euler: which axes use the cell-based edge values (PDA_MARCH2D_CELL = 0..3, bit 0: y, bit 1: x).  Default x only:
    // measured on the reference's Mach-10 double-Mach-reflection WENO5 fixture, whose absolute tolerance floor is 7 ulp
    // of the energy flux, the worst velocity entry sits at 0.83 of the tolerance face-based, 0.71 with cell-based x,
    // 1.19 with cell-based y and 1.07 with both (cfg 2 velocity: 0.314 / 0.296 / 0.297 / 0.280 ms): the form along the
    // marching axis is kept face-based for this family, every other family uses cell-based edges on both axes
    static const int mask = [] { const char* e = std::getenv("PDA_MARCH2D_CELL"); return e ? (std::atoi(e) & 3) : 2; }();
    const unsigned g = (unsigned)((tasks + 3) / 4);
    switch (mask) {
      case 1: dev::k_velocity_march2d<Phys, S, 1><<<g, 128, 0, st>>>(phys, L, dl, dU, dV, LY); break;
      case 2: dev::k_velocity_march2d<Phys, S, 2><<<g, 128, 0, st>>>(phys, L, dl, dU, dV, LY); break;
      case 3: dev::k_velocity_march2d<Phys, S, 3><<<g, 128, 0, st>>>(phys, L, dl, dU, dV, LY); break;
      default: dev::k_velocity_march2d<Phys, S, 0><<<g, 128, 0, st>>>(phys, L, dl, dU, dV, LY);
    }
  } else {
    dev::k_velocity_march2d<Phys, S, 3><<<(unsigned)((tasks + 3) / 4), 128, 0, st>>>(phys, L, dl, dU, dV, LY);
  }
}

}  // namespace pda
