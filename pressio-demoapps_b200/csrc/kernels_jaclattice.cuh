// kernels_jaclattice.cuh -- velocity + Jacobian of the inner rows of a 2D full lattice, every FACE computed once.
//
// Replaces the reference's inner-cell velocity+Jacobian loop (euler_2d_prob_class.hpp:633-720, swe_2d_prob_class.hpp
// :556-640) and its scatter (mixin_directional_flux_balance_jacobian.hpp:142-284), where each cell evaluates both
// faces of every axis (each face's reconstruction gradients and flux Jacobians twice) and every CSR entry is a
// read-modify-write through SparseMatrix::coeffRef.
//
// A lane owns ONE face: the 32 lanes of a warp sit on 32 consecutive faces of one mesh line (a row for the x phase,
// a column for the y phase), i.e. on the left faces of 32 consecutive cells; lane l also owns cell l of the line
// (31 cells per warp task: the last lane only supplies the right face of cell 30).  With
//      t[k][m][j] = JN[k][j] * d(uNeg_j)/d(q_m) + JP[k][j] * d(uPos_j)/d(q_m)         (q_m = m-th cell of the face stencil)
// the block of row-cell c at stencil position P is   hInv * ( t_c[k][P][j] - t_{c+1}[k][P-1][j] ):  the face's own
// products and its right neighbour's, fetched by ONE warp shuffle -- one set of products per face serves both cells.
//   * every (row k, position P) is N contiguous doubles (one 32-byte sector for Euler2d): stored straight to the CSR
//     value array, each entry written once, no shared-memory staging, no memset, no read-modify-write;
//     the self block alone meets both axes: the x phase stores it, the y phase adds to it (an L2 hit, 1/13 of the data);
//   * flux Jacobians (2 N^2 doubles) live in a thread-private shared-memory column, the row loop over k is rolled:
//     small code (the instruction cache was the first limiter of the staged kernel) and fewer live registers;
//   * a CTA = 31 x 31 cell tile: phase x = its 31 rows, barrier, phase y = its 31 columns (same code, other stride).
#pragma once
#include "kernels_generic.cuh"
#include "kernels_lattice.cuh"

namespace pda {
namespace dev {

template <class Phys, int S>
struct JacLat2d {
  static constexpr int N = Phys::ndpc;
  static constexpr int WARPS = 8;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int T = 31;                                        // cells per warp task, tile edge
  static constexpr size_t smemBytes = (size_t)2 * N * N * THREADS * sizeof(double);
};

struct JacLatTables {
  const int32_t* cellBase;   // [cells] offset of the cell's first CSR row
  const uint8_t* cellSlots;  // [cells][nslotCols] block position of graph column c
  int32_t nslotCols;
  int32_t rowLen;            // entries per CSR row of an inner cell
};

// one mesh line of one axis: lane l = face between cells (a-1, a), a = a0 + lane, and owner of cell a
template <class Phys, int S, int AX>
PDA_DEVFN void jacLatLine(const Phys& phys, const LatticeDesc& L, const JacLatTables& jt, double hInv,
                          const double* __restrict__ U, double* __restrict__ V, double* __restrict__ Jv,
                          int a0, int hiA, int o, double* __restrict__ sJ /* + tid */, int lane) {
  constexpr int N = Phys::ndpc;
  constexpr int h = (S - 1) / 2;
  constexpr int THREADS = JacLat2d<Phys, S>::THREADS;
  const int nx = L.n[0];
  const int nA = L.n[AX], perA = L.per[AX];
  const int a = a0 + lane;
  const bool owns = (lane < 31) && (a < hiA);

  // cells of the face stencil: coordinate a-h+m, m = 0..S-2 (wrapped on periodic axes, clamped for idle lanes)
  int64_t off[S - 1];
#pragma unroll
  for (int m = 0; m < S - 1; ++m) {
    int c = a - h + m;
    if (perA) c = (c < 0) ? c + nA : (c >= nA ? c - nA : c);
    else c = (c < 0) ? 0 : (c >= nA ? nA - 1 : c);
    const int64_t gid = (AX == 0) ? (int64_t)o * nx + c : (int64_t)c * nx + o;
    off[m] = gid * N;
  }

  // ---- face states, flux, flux Jacobians
  double F[N];
  {
    double un[N], up[N];
#pragma unroll
    for (int d = 0; d < N; ++d) {
      double q[S - 1];
#pragma unroll
      for (int m = 0; m < S - 1; ++m) q[m] = U[off[m] + d];
      Recon<S>::face(q, un[d], up[d]);
    }
    double JN[N * N], JP[N * N];
    phys.template flux<AX>(un, up, F);
    phys.template fluxJac<AX>(un, up, JN, JP);
#pragma unroll
    for (int e = 0; e < N * N; ++e) { sJ[e * THREADS] = JN[e]; sJ[(N * N + e) * THREADS] = JP[e]; }
  }
  // ---- reconstruction gradients of every dof
  double gN[N][S - 1], gP[N][S - 1];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double q[S - 1], t0, t1;
#pragma unroll
    for (int m = 0; m < S - 1; ++m) q[m] = U[off[m] + j];
    Recon<S>::faceGrad(q, t0, t1, gN[j], gP[j]);
  }

  // ---- owned cell: gid, CSR base, slots of its S stencil positions along this axis
  const int aa = owns ? a : ((a0 < hiA) ? a0 : 0);
  const int64_t gidSelf = (AX == 0) ? (int64_t)o * nx + aa : (int64_t)aa * nx + o;
  const uint8_t* slots = jt.cellSlots + gidSelf * jt.nslotCols;
  int sl[S];
  sl[h] = slots[0];
#pragma unroll
  for (int l = 0; l < h; ++l) {
    sl[h - 1 - l] = slots[gcol<2>(sideMinus<AX>(), l)];
    sl[h + 1 + l] = slots[gcol<2>(sidePlus<AX>(), l)];
  }
  double* jBase = Jv + jt.cellBase[gidSelf];
  const int rowLen = jt.rowLen;

  // ---- velocity: hInv (F_left - F_right)
  {
    double v[N];
#pragma unroll
    for (int d = 0; d < N; ++d) v[d] = hInv * (F[d] - __shfl_down_sync(0xffffffffu, F[d], 1));
    if (owns && V) {
      double* out = V + gidSelf * N;
      if (AX == 0) {
#pragma unroll
        for (int d = 0; d < N; ++d) out[d] = v[d];
      } else {
#pragma unroll
        for (int d = 0; d < N; ++d) v[d] += out[d];
        if constexpr (PhysTraits<Phys>::hasDiffusion) {
          // first-layer neighbours as a graph row (left, front, right, back): inner cells have them all
          const int i = (int)(gidSelf % nx), jrow = (int)(gidSelf / nx), ny = L.n[1];
          auto wrapI = [&](int c, int n) { return c < 0 ? c + n : (c >= n ? c - n : c); };
          int32_t row5[5];
          row5[0] = (int32_t)gidSelf;
          row5[1] = jrow * nx + wrapI(i - 1, nx);
          row5[2] = wrapI(jrow + 1, ny) * nx + i;
          row5[3] = jrow * nx + wrapI(i + 1, nx);
          row5[4] = wrapI(jrow - 1, ny) * nx + i;
          addDiffusionInner<Phys>(phys, row5, U, v);
        }
        addForcing<Phys>(phys, U + gidSelf * N, v, (int32_t)gidSelf);
#pragma unroll
        for (int d = 0; d < N; ++d) out[d] = v[d];
      }
    }
  }

  // ---- Jacobian rows k = 0..N-1 (rolled), positions and dofs unrolled
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    double jn[N], jp[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { jn[j] = hInv * sJ[(k * N + j) * THREADS]; jp[j] = hInv * sJ[(N * N + k * N + j) * THREADS]; }
    double* rowp = jBase + (int64_t)k * rowLen;
#pragma unroll
    for (int P = 0; P < S; ++P) {
      double val[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        // own products at m = P (left face of the owned cell); partner's at m = P-1 (its face is my right face)
        double mine = 0.0, give = 0.0;
        if (P <= S - 2) mine = jn[j] * gN[j][P] + jp[j] * gP[j][P];
        if (P >= 1) give = jn[j] * gN[j][P - 1] + jp[j] * gP[j][P - 1];
        const double got = (P >= 1) ? __shfl_down_sync(0xffffffffu, give, 1) : 0.0;
        val[j] = mine - got;
      }
      if (owns) {
        double* dst = rowp + sl[P] * N;
        if (AX != 0 && P == h) {
#pragma unroll
          for (int j = 0; j < N; ++j) val[j] += dst[j];
        }
        if constexpr (N == 4) {
          reinterpret_cast<double2*>(dst)[0] = make_double2(val[0], val[1]);
          reinterpret_cast<double2*>(dst)[1] = make_double2(val[2], val[3]);
        } else if constexpr (N == 2) {
          reinterpret_cast<double2*>(dst)[0] = make_double2(val[0], val[1]);
        } else {
#pragma unroll
          for (int j = 0; j < N; ++j) dst[j] = val[j];
        }
      }
    }
  }
  // ---- point terms and diffusion (after both flux phases: the reference adds them last)
  if (AX != 0 && owns) {
    addExtraJacInner<Phys>(phys, U + gidSelf * N, slots, [&](int k, int slot, int j, double val) {
      jBase[(int64_t)k * rowLen + slot * N + j] += val;
    });
  }
}

template <class Phys, int S>
__global__ void __launch_bounds__(JacLat2d<Phys, S>::THREADS, 1)
k_jacobian_lattice2d(Phys phys, LatticeDesc L, Deltas dl, JacLatTables jt, const double* __restrict__ U,
                     double* __restrict__ V, double* __restrict__ Jv) {
  using K = JacLat2d<Phys, S>;
  extern __shared__ double sJall[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* sJ = sJall + tid;
  const int lo0 = L.per[0] ? 0 : L.meshHalo, hi0 = L.per[0] ? L.n[0] : L.n[0] - L.meshHalo;
  const int lo1 = L.per[1] ? 0 : L.meshHalo, hi1 = L.per[1] ? L.n[1] : L.n[1] - L.meshHalo;
  const int I0 = lo0 + K::T * blockIdx.x, J0 = lo1 + K::T * blockIdx.y;
  // phase x: rows J0 .. J0+30
  for (int task = warp; task < K::T; task += K::WARPS) {
    const int j = J0 + task;
    if (j >= hi1) break;
    jacLatLine<Phys, S, 0>(phys, L, jt, dl.hInv[0], U, V, Jv, I0, hi0, j, sJ, lane);
  }
  __syncthreads();   // self blocks and V of the tile are in place (block-scope visibility of the global stores)
  // phase y: columns I0 .. I0+30
  for (int task = warp; task < K::T; task += K::WARPS) {
    const int i = I0 + task;
    if (i >= hi0) break;
    jacLatLine<Phys, S, 1>(phys, L, jt, dl.hInv[1], U, V, Jv, J0, hi1, i, sJ, lane);
  }
}

}  // namespace dev
}  // namespace pda
