// kernels_applylattice.cuh -- matrix-free applyJacobian on 2D full lattices: R = J(U) * B without ever storing J.
//
// Replaces adapter_cpp.hpp:231-259 (evaluate J into a scratch SparseMatrix, then J*B) for the inner rows.  With
//   D_f[k] = sum_j JN[k][j] * dN[j] + JP[k][j] * dP[j],   dN[j] = sum_m d(uNeg_j)/d(q_m) * b_m[j]   (dP alike)
// -- the directional derivative of face f's flux along the operand column b -- the row block of cell c is
//   R_c[k] = sum_axes hInv * (D_c[k] - D_{c+1}[k])  (+ point / diffusion terms),
// so a face contributes ONE N-vector per operand column: no CSR values, no 1.7 kB/cell of stores, no second pass
// reading them back.  Same lane layout as k_jacobian_lattice2d (a lane owns one face, 16-lane lines, 15x15 tiles, x
// phase then y phase, the x part of R parked in shared memory); up to NC operand columns per pass share the
// reconstruction gradients and flux Jacobians.  Near-boundary rows (first-order Jacobian with ghost factors) go
// through the assembled path restricted to those rows.
#pragma once
#include "kernels_jaclattice.cuh"

namespace pda {
namespace dev {

template <class Phys, int NC>
struct ApplyLat2d {
  static constexpr int N = Phys::ndpc;
  static constexpr int WARPS = 8, THREADS = 32 * WARPS, T = 15;
  static constexpr int RS = N * NC + 1;                       // padded per-cell stride of the parked x part
  static constexpr size_t smemBytes = (size_t)T * T * RS * sizeof(double);
};

template <class Phys, int S, int AX, int NC>
PDA_DEVFN void applyLatLine(const Phys& phys, const LatticeDesc& L, double hInv, const double* __restrict__ U,
                            const double* __restrict__ B, int ncols, int c0, int64_t ldbRow, int64_t ldbCol,
                            double* __restrict__ R, int64_t ldrRow, int64_t ldrCol, int a, int o, bool owns,
                            int cellLocal, double* __restrict__ sR) {
  constexpr int N = Phys::ndpc;
  constexpr int h = (S - 1) / 2;
  using K = ApplyLat2d<Phys, NC>;
  const int nx = L.n[0], ny = L.n[1];
  const int nA = L.n[AX], perA = L.per[AX];
  const int nc = min(NC, ncols - c0);
  const bool vec4 = (ldbCol == 1) && ((ldbRow & 3) == 0) && ((c0 & 3) == 0) && (nc == 4) &&
                    ((reinterpret_cast<uintptr_t>(B) & 31) == 0);

  int64_t off[S - 1];
  double q[S - 1][N];
#pragma unroll
  for (int m = 0; m < S - 1; ++m) {
    int c = a - h + m;
    if (perA) { c %= nA; if (c < 0) c += nA; }
    else c = (c < 0) ? 0 : (c >= nA ? nA - 1 : c);
    const int64_t gid = (AX == 0) ? (int64_t)o * nx + c : (int64_t)c * nx + o;
    off[m] = gid * N;
    loadCell<N>(U + off[m], q[m]);
  }
  // face states + directional derivatives of the reconstructions along every operand column
  double un[N], up[N], dN[NC][N], dP[NC][N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double qd[S - 1], gN[S - 1], gP[S - 1];
#pragma unroll
    for (int m = 0; m < S - 1; ++m) qd[m] = q[m][j];
    reconFaceValGradFast<S>(qd, un[j], up[j], gN, gP);
    if (NC == 4 && vec4) {
      // row-major operand, 4 aligned columns: the 4 values of a (cell, dof) are one 32-byte sector -> LDG.256
      double sN[4] = {0.0, 0.0, 0.0, 0.0}, sP[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int m = 0; m < S - 1; ++m) {
        double b[4];
        loadCell<4>(B + (off[m] + j) * ldbRow + c0, b);
#pragma unroll
        for (int c = 0; c < 4; ++c) { sN[c] = fma(gN[m], b[c], sN[c]); sP[c] = fma(gP[m], b[c], sP[c]); }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) { dN[c][j] = sN[c & 3]; dP[c][j] = sP[c & 3]; }
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double sN = 0.0, sP = 0.0;
        if (c < nc) {
#pragma unroll
          for (int m = 0; m < S - 1; ++m) {
            const double b = __ldg(B + (off[m] + j) * ldbRow + (int64_t)(c0 + c) * ldbCol);
            sN = fma(gN[m], b, sN);
            sP = fma(gP[m], b, sP);
          }
        }
        dN[c][j] = sN; dP[c][j] = sP;
      }
    }
  }
  double r[NC][N];
  if constexpr (std::is_same<Phys, Euler<2>>::value) {
    // Euler: the Jacobian-vector product of the Rusanov flux in closed form (no N x N matrices, fastmath.cuh)
    double D[NC][N];
    eulerFluxJvpFast<2, AX, NC>(phys.gamma, un, up, dN, dP, D);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int k = 0; k < N; ++k) r[c][k] = hInv * (D[c][k] - __shfl_down_sync(0xffffffffu, D[c][k], 1));
  } else {
    double JN[N * N], JP[N * N];
    faceFluxJac2d<Phys, AX>(phys, un, up, JN, JP);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int k = 0; k < N; ++k) {
        double d = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) d += JN[k * N + j] * dN[c][j] + JP[k * N + j] * dP[c][j];
        r[c][k] = hInv * (d - __shfl_down_sync(0xffffffffu, d, 1));
      }
  }
  if (!owns) return;
  double* mine = sR + cellLocal * K::RS;
  if (AX == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int k = 0; k < N; ++k) mine[c * N + k] = r[c][k];
    return;
  }
  const int64_t gidSelf = (int64_t)a * nx + o;   // y phase: a = row, o = column
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int k = 0; k < N; ++k) r[c][k] += mine[c * N + k];
  if constexpr (std::is_same<Phys, Swe2d>::value || PhysTraits<Phys>::hasDiffusion || std::is_same<Phys, LinAdv<2>>::value) {
    // point terms and diffusion: the same entries addExtraJacInner adds to J, applied to the operand.  The "slot"
    // handed to it is the graph column itself (0 self, 1 left, 2 front, 3 right, 4 back).
    auto wrapI = [&](int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); };
    const int i = o, jr = a;
    int64_t nb[5];
    nb[0] = gidSelf;
    nb[1] = (int64_t)jr * nx + wrapI(i - 1, nx);
    nb[2] = (int64_t)wrapI(jr + 1, ny) * nx + i;
    nb[3] = (int64_t)jr * nx + wrapI(i + 1, nx);
    nb[4] = (int64_t)wrapI(jr - 1, ny) * nx + i;
    uint8_t ident[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) ident[c] = (uint8_t)c;
    addExtraJacInner<Phys>(phys, U + gidSelf * N, ident, [&](int k, int col, int j, double x) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < nc) r[c][k] += x * __ldg(B + (nb[col] * N + j) * ldbRow + (int64_t)(c0 + c) * ldbCol);
    });
  }
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < nc) {
#pragma unroll
      for (int k = 0; k < N; ++k) R[(gidSelf * N + k) * ldrRow + (int64_t)(c0 + c) * ldrCol] = r[c][k];
    }
}

template <class Phys, int S, int NC>
__global__ void __launch_bounds__(ApplyLat2d<Phys, NC>::THREADS)
k_applyjac_lattice2d(Phys phys, LatticeDesc L, Deltas dl, const double* __restrict__ U, const double* __restrict__ B,
                     int ncols, int c0, int64_t ldbRow, int64_t ldbCol, double* __restrict__ R, int64_t ldrRow,
                     int64_t ldrCol) {
  using K = ApplyLat2d<Phys, NC>;
  extern __shared__ __align__(16) double sR[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int line = 2 * warp + (lane >> 4), f = lane & 15;
  const int lo0 = L.per[0] ? 0 : L.meshHalo, hi0 = L.per[0] ? L.n[0] : L.n[0] - L.meshHalo;
  const int lo1 = L.per[1] ? 0 : L.meshHalo, hi1 = L.per[1] ? L.n[1] : L.n[1] - L.meshHalo;
  const int I0 = lo0 + K::T * blockIdx.x, J0 = lo1 + K::T * blockIdx.y;
  {
    const int j = J0 + line, a = I0 + f;
    const bool owns = (line < K::T) && (f < K::T) && (j < hi1) && (a < hi0);
    applyLatLine<Phys, S, 0, NC>(phys, L, dl.hInv[0], U, B, ncols, c0, ldbRow, ldbCol, R, ldrRow, ldrCol, a,
                                 min(j, hi1 - 1), owns, min(line, K::T - 1) * K::T + min(f, K::T - 1), sR);
  }
  __syncthreads();
  {
    const int i = I0 + line, a = J0 + f;
    const bool owns = (line < K::T) && (f < K::T) && (i < hi0) && (a < hi1);
    applyLatLine<Phys, S, 1, NC>(phys, L, dl.hInv[1], U, B, ncols, c0, ldbRow, ldbCol, R, ldrRow, ldrCol, a,
                                 min(i, hi0 - 1), owns, min(f, K::T - 1) * K::T + min(line, K::T - 1), sR);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// 3D: the same directional-derivative formulation on full 3D lattices (Euler3d).  There is no assembled alternative
// at scale: the Jacobian of a 512^3 WENO5 problem has 6.4e10 stored entries (0.5 TB, beyond the reference's int32
// indexing), while J*v needs the state, the operand and the result only.
// Tile = 7 x 7 x 7 cells; a line (7 cells + the closing face) occupies 8 lanes, a warp carries 4 lines; phases x, y, z
// with the partial result parked in shared memory in between.
// Tried and dropped (session 14): `prefetch.global.L1` of the stencil the next (phase, round) will read, against the
// long_scoreboard stalls ncu reports (2.5 of 8 cycles per issue): 81 -> 94 ms at 512^3 -- the 24 prefetch instructions per
// line compete for the same LSU slots as the loads they were meant to hide.
// ---------------------------------------------------------------------------------------------------------------
template <int NC>
struct ApplyLat3d {
  static constexpr int N = 5, T = 7, THREADS = 224;   // 28 lines of 8 lanes per round, 49 lines per phase: 2 rounds
  static constexpr int GROUPS = THREADS / 8;
  static constexpr int RS = N * NC + 1;
  static constexpr size_t smemBytes = (size_t)T * T * T * RS * sizeof(double);
};

// One copy of the line code serves the three phases (`ax` is a run-time, warp-uniform argument): the kernel's
// instruction footprint is a third of the per-axis instantiation (ncu on the first version: `no_instruction`
// 2.2 of 9.4 stall cycles per issue).  The flux Jacobians are always taken along x of a state whose momentum
// components 1 and 1+ax are swapped -- the Euler flux along axis ax of q is P F_x(P q) with P that swap -- so
// D = P (JN' P dN + JP' P dP): four cheap swaps instead of three copies of eulerFluxJacFast.
PDA_DEVFN void swapMomentum(int ax, double* v) {
  const double v1 = v[1], v2 = v[2], v3 = v[3];
  v[1] = (ax == 0) ? v1 : ((ax == 1) ? v2 : v3);
  v[2] = (ax == 1) ? v1 : v2;
  v[3] = (ax == 2) ? v1 : v3;
}

template <int S, int NC>
PDA_DEVFN void applyLatLine3(double gamma, const LatticeDesc& L, int ax, double hInv, const double* __restrict__ U,
                             const double* __restrict__ B, int ncols, int c0, int64_t ldbRow, int64_t ldbCol,
                             double* __restrict__ R, int64_t ldrRow, int64_t ldrCol, int a, int o1, int o2, bool owns,
                             int cellLocal, double* __restrict__ sR) {
  constexpr int N = 5;
  constexpr int h = (S - 1) / 2;
  using K = ApplyLat3d<NC>;
  const int64_t nx = L.n[0], ny = L.n[1];
  const int nA = (ax == 0) ? L.n[0] : ((ax == 1) ? L.n[1] : L.n[2]);
  const int perA = (ax == 0) ? L.per[0] : ((ax == 1) ? L.per[1] : L.per[2]);
  const int nc = min(NC, ncols - c0);
  // (a, o1, o2) -> (x, y, z): ax = 0: a = x, (o1, o2) = (y, z); ax = 1: a = y, (o1, o2) = (x, z); ax = 2: a = z, (x, y)
  const int64_t lineBase = (ax == 0) ? ((int64_t)o2 * ny + o1) * nx : ((ax == 1) ? (int64_t)o2 * ny * nx + o1 : (int64_t)o2 * nx + o1);
  const int64_t lineStride = (ax == 0) ? 1 : ((ax == 1) ? nx : nx * ny);
  auto gidOf = [&](int c) -> int64_t { return lineBase + (int64_t)c * lineStride; };
  int64_t off[S - 1];
  double q[S - 1][N];
#pragma unroll
  for (int m = 0; m < S - 1; ++m) {
    int c = a - h + m;
    if (perA) { c %= nA; if (c < 0) c += nA; }
    else c = (c < 0) ? 0 : (c >= nA ? nA - 1 : c);
    off[m] = gidOf(c) * N;
    loadCell<N>(U + off[m], q[m]);
  }
  double un[N], up[N], dN[NC][N], dP[NC][N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double qd[S - 1], gN[S - 1], gP[S - 1];
#pragma unroll
    for (int m = 0; m < S - 1; ++m) qd[m] = q[m][j];
    reconFaceValGradFast<S>(qd, un[j], up[j], gN, gP);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double sN = 0.0, sP = 0.0;
      if (c < nc) {
#pragma unroll
        for (int m = 0; m < S - 1; ++m) {
          const double b = __ldg(B + (off[m] + j) * ldbRow + (int64_t)(c0 + c) * ldbCol);
          sN = fma(gN[m], b, sN);
          sP = fma(gP[m], b, sP);
        }
      }
      dN[c][j] = sN; dP[c][j] = sP;
    }
  }
  // D = P (JN' P dN + JP' P dP) as ONE Jacobian-vector product of the x-direction flux (no N x N matrices)
  swapMomentum(ax, un);
  swapMomentum(ax, up);
#pragma unroll
  for (int c = 0; c < NC; ++c) { swapMomentum(ax, dN[c]); swapMomentum(ax, dP[c]); }
  double d[NC][N];
  eulerFluxJvpFast<3, 0, NC>(gamma, un, up, dN, dP, d);
  double r[NC][N];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    swapMomentum(ax, d[c]);
#pragma unroll
    for (int k = 0; k < N; ++k) r[c][k] = hInv * (d[c][k] - __shfl_down_sync(0xffffffffu, d[c][k], 1));
  }
  if (!owns) return;
  double* mine = sR + cellLocal * K::RS;
  if (ax == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int k = 0; k < N; ++k) mine[c * N + k] = r[c][k];
  } else if (ax == 1) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int k = 0; k < N; ++k) mine[c * N + k] += r[c][k];
  } else {
    const int64_t gidSelf = gidOf(a);
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c < nc) {
#pragma unroll
        for (int k = 0; k < N; ++k)
          R[(gidSelf * N + k) * ldrRow + (int64_t)(c0 + c) * ldrCol] = mine[c * N + k] + r[c][k];   // (x + y) + z
      }
  }
}

template <int S, int NC>
__global__ void __launch_bounds__(ApplyLat3d<NC>::THREADS, (NC == 1 ? 2 : 1))
k_applyjac_lattice3d(double gamma, LatticeDesc L, Deltas dl, const double* __restrict__ U, const double* __restrict__ B,
                     int ncols, int c0, int64_t ldbRow, int64_t ldbCol, double* __restrict__ R, int64_t ldrRow,
                     int64_t ldrCol) {
  using K = ApplyLat3d<NC>;
  constexpr int T = K::T;
  extern __shared__ __align__(16) double sR3[];
  const int tid = threadIdx.x;
  const int grp = tid >> 3, f = tid & 7;   // K::GROUPS lines of 8 lanes per round
  int lo[3], hi[3];
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    lo[ax] = L.per[ax] ? 0 : L.meshHalo;
    hi[ax] = L.per[ax] ? L.n[ax] : L.n[ax] - L.meshHalo;
  }
  const int O[3] = {lo[0] + T * (int)blockIdx.x, lo[1] + T * (int)blockIdx.y, lo[2] + T * (int)blockIdx.z};
#pragma unroll 1
  for (int ax = 0; ax < 3; ++ax) {
    // origin / upper bound along the line axis and along the two other axes (ascending)
    const int Oa = (ax == 0) ? O[0] : ((ax == 1) ? O[1] : O[2]);
    const int O1 = (ax == 0) ? O[1] : O[0], O2 = (ax == 2) ? O[1] : O[2];
    const int hia = (ax == 0) ? hi[0] : ((ax == 1) ? hi[1] : hi[2]);
    const int hi1 = (ax == 0) ? hi[1] : hi[0], hi2 = (ax == 2) ? hi[1] : hi[2];
    const double hInv = (ax == 0) ? dl.hInv[0] : ((ax == 1) ? dl.hInv[1] : dl.hInv[2]);
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
      const int l = round * K::GROUPS + grp;     // line index in the tile: (u, v) offsets along the other axes
      const int u = l % T, v = l / T;
      const int a = Oa + f;
      const int c1 = O1 + u, c2 = O2 + v;
      const bool owns = (l < T * T) && (f < T) && (a < hia) && (c1 < hi1) && (c2 < hi2);
      const int cc1 = min(c1, hi1 - 1), cc2 = min(c2, hi2 - 1);
      const int uu = min(u, T - 1), vv = min(v, T - 1), ff = min(f, T - 1);
      // tile-local linear index (z*T + y)*T + x
      const int lx = (ax == 0) ? ff : uu;
      const int ly = (ax == 0) ? uu : ((ax == 1) ? ff : vv);
      const int lz = (ax == 2) ? ff : vv;
      const int cellLocal = (lz * T + ly) * T + lx;
      applyLatLine3<S, NC>(gamma, L, ax, hInv, U, B, ncols, c0, ldbRow, ldbCol, R, ldrRow, ldrCol, a, cc1, cc2, owns,
                           cellLocal, sR3);
    }
    __syncthreads();
  }
}

}  // namespace dev
}  // namespace pda
