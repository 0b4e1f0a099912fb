// kernels_jacobian.cuh -- inner-row Jacobian assembly with every CSR value written exactly once, coalesced.
//
// Replaces the reference's inner-cell velocity+Jacobian loop (euler_2d_prob_class.hpp:633-720,
// euler_3d_prob_class.hpp:410-523) and its scatter (mixin_directional_flux_balance_jacobian.hpp:142-284), where
// every entry is a read-modify-write through SparseMatrix::coeffRef (a binary search each) on zeroed memory.
//
// The ndpc rows of a cell share one column pattern and are consecutive in the CSR arrays, so the whole row block
// of an inner cell is ONE contiguous chunk of ndpc * ndpc * (1 + dim*(S-1)) doubles (1664 B for 2D Euler WENO5),
// and consecutive inner cells are adjacent chunks.  The kernel is bound by those stores (1.7 kB/cell against 64 B
// of state), so:
//   * the chunk of a cell is assembled in SHARED memory by 2*dim threads, one per (axis, face): each computes its
//     face's flux, flux Jacobians and reconstruction gradients; the two faces of an axis sit in adjacent lanes and
//     swap the contributions to the stencil positions they share by warp shuffle, so every entry is stored ONCE
//     (no zeroing, no read-modify-write); axes live in different warps (no divergence) and meet only in the self
//     block, which they update in x,y,z order between barriers -- the reference's accumulation order;
//   * the CTA then streams the chunks to HBM with coalesced 8-byte stores, each value written once
//     (no memset, no read-modify-write in HBM: algorithmic bytes = nnz*8).
// Requires the regular inner-row pattern (no coincident neighbours); tiny periodic meshes whose neighbours merge keep
// the read-modify-write kernel (k_jacobian_inner_rows).
#pragma once
#include "fastmath.cuh"
#include "kernels_generic.cuh"

namespace pda {
namespace dev {

template <class Phys, int S>
struct JacStage {
  static constexpr int N = Phys::ndpc;
  static constexpr int DIM = Phys::dim;
  static constexpr int NBLK = 1 + DIM * (S - 1);
  static constexpr int ROWLEN = N * NBLK;            // entries per CSR row
  static constexpr int CHUNK = N * ROWLEN;           // doubles per cell
  static constexpr int STRIDE = CHUNK | 1;           // odd stride: conflict-free 8-byte accesses across lanes
  // cells per CTA (multiple of 16: a warp = 16 cells x 2 faces of one axis): as many as fit ~110 KB
  static constexpr int FIT = 110 * 1024 / (STRIDE * 8);
  static constexpr int CELLS = FIT >= 64 ? 64 : (FIT >= 48 ? 48 : (FIT >= 32 ? 32 : 16));
  static constexpr int THREADS = CELLS * 2 * DIM;
  static constexpr size_t smemBytes = (size_t)CELLS * STRIDE * 8;
  // CTAs per SM the register allocation is capped for: as many as the staged chunks leave room for (<= 4)
  static constexpr int FIT_CTAS = (int)(220 * 1024 / smemBytes);
  static constexpr int WANT_CTAS = FIT_CTAS > 4 ? 4 : (FIT_CTAS < 1 ? 1 : FIT_CTAS);
  // measured (B200, session 11): the cap pays where the spills stay small -- 3D WENO3 160^3 8.10 -> 5.84 ms, cfg 4 WENO3
  // 0.254 -> 0.227 ms -- and loses for the WENO5 instances (3D 4.54 -> 5.24 ms with 284 bytes of spills): those stay uncapped
  static constexpr int MIN_CTAS = (S == 7) ? 1 : ((THREADS * WANT_CTAS <= 512) ? WANT_CTAS : (512 / THREADS < 1 ? 1 : 512 / THREADS));
};

// Rusanov flux and flux Jacobians along a RUN-TIME axis.  Euler and shallow water: the flux along axis a is
// P F_x(P q) with P the swap of momentum components 1 and 1+a, so ONE copy of the x-direction code serves every axis
// (JN = P JN' P: rows and columns swapped); the scalar families keep their two small per-axis instances.
template <class Phys>
PDA_DEVFN void faceFluxAndJacAxis(const Phys& phys, int axis, double* un, double* up, double* F, double* JN, double* JP) {
  constexpr int N = Phys::ndpc, DIM = Phys::dim;
  constexpr bool kSwap = std::is_same<Phys, Euler<2>>::value || std::is_same<Phys, Euler<3>>::value || std::is_same<Phys, Swe2d>::value;
  if constexpr (DIM == 1) {
    faceFlux2d<Phys, 0>(phys, un, up, F);
    faceFluxJac2d<Phys, 0>(phys, un, up, JN, JP);
  } else if constexpr (kSwap) {
    auto swp = [&](double& a, double& b, bool on) { const double t = a; a = on ? b : a; b = on ? t : b; };
    auto swapVec = [&](double* v) {
      swp(v[1], v[2], axis == 1);
      if constexpr (DIM == 3) swp(v[1], v[3], axis == 2);
    };
    swapVec(un); swapVec(up);
    faceFlux2d<Phys, 0>(phys, un, up, F);
    faceFluxJac2d<Phys, 0>(phys, un, up, JN, JP);
    swapVec(F);
#pragma unroll
    for (int k = 0; k < N; ++k) { swapVec(JN + k * N); swapVec(JP + k * N); }        // columns
#pragma unroll
    for (int j = 0; j < N; ++j) {                                                    // rows
      swp(JN[1 * N + j], JN[2 * N + j], axis == 1); swp(JP[1 * N + j], JP[2 * N + j], axis == 1);
      if constexpr (DIM == 3) { swp(JN[1 * N + j], JN[3 * N + j], axis == 2); swp(JP[1 * N + j], JP[3 * N + j], axis == 2); }
    }
  } else {
    if (axis == 0) { faceFlux2d<Phys, 0>(phys, un, up, F); faceFluxJac2d<Phys, 0>(phys, un, up, JN, JP); }
    else { faceFlux2d<Phys, 1>(phys, un, up, F); faceFluxJac2d<Phys, 1>(phys, un, up, JN, JP); }
  }
}

// one (axis, face) role of one cell; `face` 0 = left/back/bottom face, 1 = right/front/top.
// ONE copy of the code for every axis (run-time `axis`, warp-uniform) and every dof: the loops over the dofs are ROLLED
// (the WENO5 3D instance of the per-axis, fully unrolled version was 325 KB of SASS and stalled on instruction fetch);
// the arrays a rolled loop would index at run time (flux Jacobian column j, self-block column j) are ROTATED by one
// column per iteration instead, so that they stay in registers.
template <class Phys, int S>
PDA_DEVFN void jacobianFaceRole(const Phys& phys, int axis, const int32_t* __restrict__ row, const double* __restrict__ U,
                                double hInv, int face, double* __restrict__ my, const uint8_t* __restrict__ slots,
                                double* selfAcc /*[N*N], L lane only*/, double* dF /*[N] F_L - F_R, L lane only*/) {
  constexpr int N = Phys::ndpc;
  constexpr int h = (S - 1) / 2;
  constexpr int DIM = Phys::dim;
  constexpr int ROWLEN = JacStage<Phys, S>::ROWLEN;
  const int sm = (axis == 0) ? 0 : (axis == 1 ? 3 : 4), sp = (axis == 0) ? 2 : (axis == 1 ? 1 : 5);
  // graph column of stencil position pos: 0..h-1 = minus layers (far..near), h = self, h+1.. = plus layers
  auto colOf = [&](int pos) -> int {
    return (pos == h) ? 0 : (pos < h ? gcol<DIM>(sm, h - 1 - pos) : gcol<DIM>(sp, pos - h - 1));
  };
  int sl[S];   // block slot of each stencil position
#pragma unroll
  for (int p = 0; p < S; ++p) sl[p] = slots[colOf(p)];
  int64_t fc[S - 1];   // the face's stencil: cells at positions face .. face+S-2 (element offsets into U)
#pragma unroll
  for (int p = 0; p < S - 1; ++p) fc[p] = (int64_t)row[colOf(p + face)] * N;
  const double sgn = (face == 0) ? hInv : -hInv;
  double un[N], up[N];
#pragma unroll
  for (int d = 0; d < N; ++d) { un[d] = 0.0; up[d] = 0.0; }
#pragma unroll 1
  for (int d = 0; d < N; ++d) {
    double q[S - 1];
#pragma unroll
    for (int p = 0; p < S - 1; ++p) q[p] = U[fc[p] + d];
    double a, b;
    reconFaceFast<S>(q, a, b);
#pragma unroll
    for (int i = 0; i < N - 1; ++i) { un[i] = un[i + 1]; up[i] = up[i + 1]; }
    un[N - 1] = a; up[N - 1] = b;
  }
  double F[N], JN[N * N], JP[N * N];
  faceFluxAndJacAxis<Phys>(phys, axis, un, up, F, JN, JP);
#pragma unroll
  for (int d = 0; d < N; ++d) {
    const double other = __shfl_xor_sync(0xffffffffu, F[d], 1);
    dF[d] = F[d] - other;   // meaningful in the L lane: F_L - F_R
  }
#pragma unroll
  for (int i = 0; i < N * N; ++i) { JN[i] *= sgn; JP[i] *= sgn; selfAcc[i] = 0.0; }
  // chain rule: d(flux)/d(u_p) = JN * diag(d uNeg/d u_p) + JP * diag(d uPos/d u_p).  The L lane owns stencil
  // positions 0..h, the R lane h+1..S-1; position p receives the L face's column m = p and the R face's m = p-1.
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    double q[S - 1], gN[S - 1], gP[S - 1];
#pragma unroll
    for (int p = 0; p < S - 1; ++p) q[p] = U[fc[p] + j];
    reconFaceGradFast<S>(q, gN, gP);
    double selfCol[N];
#pragma unroll
    for (int k = 0; k < N; ++k) selfCol[k] = 0.0;
#pragma unroll
    for (int s = 0; s <= h; ++s) {
      const int pOwnL = s, pOwnR = h + 1 + s;          // positions owned by the L / R lane in this round
      const bool rHas = (pOwnR <= S - 1);
      // gradient columns (compile-time indices, selected by the lane's face):
      //   L lane: own position pOwnL -> column pOwnL ; partner's position pOwnR -> column pOwnR (if <= S-2)
      //   R lane: own position pOwnR -> column pOwnR-1 ; partner's position pOwnL -> column pOwnL-1 (if >= 0)
      const int cLm = pOwnL, cLt = (pOwnR <= S - 2) ? pOwnR : 0;
      const int cRm = rHas ? pOwnR - 1 : 0, cRt = (pOwnL >= 1) ? pOwnL - 1 : 0;
      const bool LtValid = (pOwnR <= S - 2), RtValid = (pOwnL >= 1);
      const double gNm = (face == 0) ? gN[cLm] : (rHas ? gN[cRm] : 0.0);
      const double gPm = (face == 0) ? gP[cLm] : (rHas ? gP[cRm] : 0.0);
      const double gNt = (face == 0) ? (LtValid ? gN[cLt] : 0.0) : (RtValid ? gN[cRt] : 0.0);
      const double gPt = (face == 0) ? (LtValid ? gP[cLt] : 0.0) : (RtValid ? gP[cRt] : 0.0);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double jn = JN[k * N], jp = JP[k * N];   // column j (rotated to the front)
        const double mine = jn * gNm + jp * gPm;
        const double theirs = jn * gNt + jp * gPt;
        const double got = __shfl_xor_sync(0xffffffffu, theirs, 1);
        const double val = mine + got;
        if (face == 0) {
          if (pOwnL == h) selfCol[k] = val;                                // self block: stored in axis order later
          else my[k * ROWLEN + sl[pOwnL] * N + j] = val;
        } else if (rHas) {
          my[k * ROWLEN + sl[pOwnR] * N + j] = val;
        }
      }
    }
    // rotate: column j+1 of the flux Jacobians to the front, this iteration's self-block column in at the back
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int i = 0; i < N - 1; ++i) {
        JN[k * N + i] = JN[k * N + i + 1]; JP[k * N + i] = JP[k * N + i + 1];
        selfAcc[k * N + i] = selfAcc[k * N + i + 1];
      }
      selfAcc[k * N + N - 1] = selfCol[k];
    }
  }
}

// zero the CSR values of the given cells (near-boundary rows are assembled by read-modify-write); one warp per cell
__global__ void k_zero_cell_chunks(const int32_t* __restrict__ base, const int32_t* __restrict__ len, int32_t n,
                                   int ndpc, double* __restrict__ Jv) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  double* dst = Jv + base[warp];
  const int cnt = len[warp] * ndpc;
  for (int e = lane; e < cnt; e += 32) dst[e] = 0.0;
}

template <class Phys, int S, int MINB>
__global__ void __launch_bounds__(JacStage<Phys, S>::THREADS, MINB)
k_jacobian_inner_staged(Phys phys, RowSet rs, Deltas dl, const double* __restrict__ U, double* __restrict__ V,
                        double* __restrict__ Jv, JacLayout jl) {
  using J = JacStage<Phys, S>;
  constexpr int N = J::N, DIM = J::DIM, CELLS = J::CELLS, CHUNK = J::CHUNK, STRIDE = J::STRIDE, ROWLEN = J::ROWLEN;
  constexpr int THREADS = J::THREADS;
  extern __shared__ double sJ[];
  __shared__ int64_t sBase[CELLS];
  __shared__ double sV[CELLS * N];
  const int t = threadIdx.x;
  const int axis = t / (2 * CELLS);            // warp-uniform (CELLS is a multiple of 16)
  const int c = (t % (2 * CELLS)) >> 1;
  const int face = t & 1;
  const int32_t r0 = blockIdx.x * CELLS + c;
  const bool valid = r0 < rs.n;
  const int32_t r = valid ? r0 : rs.n - 1;     // out-of-range roles recompute the last row (they must take part in
                                               // the shuffles) and simply do not publish it
  const int32_t* row = rs.graph + (int64_t)r * rs.ncols;
  const uint8_t* slots = jl.slot + (int64_t)r * jl.nslotCols;
  double* my = sJ + (size_t)c * STRIDE;
  if (axis == 0 && face == 0) sBase[c] = valid ? (int64_t)jl.base[r] : -1;
  double selfAcc[N * N], dF[N];
  const double hInvAx = (axis == 0) ? dl.hInv[0] : (axis == 1 ? dl.hInv[1] : dl.hInv[2]);
  jacobianFaceRole<Phys, S>(phys, axis, row, U, hInvAx, face, my, slots, selfAcc, dF);
  // self block and velocity: x, then y, then z (the reference's accumulation order), one barrier apart
  const int sSelf = slots[0];
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    if (axis == a && face == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k)
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double* dst = my + k * ROWLEN + sSelf * N + j;
          if (a == 0) *dst = selfAcc[k * N + j]; else *dst += selfAcc[k * N + j];
        }
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const double term = dl.hInv[a] * dF[d];
        if (a == 0) sV[c * N + d] = term; else sV[c * N + d] += term;
      }
      if (a == DIM - 1) {   // source terms last (swe_2d_prob_class.hpp:984-1012)
        const double* uSelf = U + (int64_t)row[0] * N;
        double v[N];
#pragma unroll
        for (int d = 0; d < N; ++d) v[d] = sV[c * N + d];
        addDiffusionInner<Phys>(phys, row, U, v);
        addForcing<Phys>(phys, uSelf, v, valid ? rs.rowIds[r] : 0);
        addExtraJacInner<Phys>(phys, uSelf, slots, [&](int k, int slot, int j, double val) {
          my[k * ROWLEN + slot * N + j] += val;
        });
        if (V && valid) {
          double* out = V + (int64_t)rs.rowIds[r] * N;
#pragma unroll
          for (int d = 0; d < N; ++d) out[d] = v[d];
        }
      }
    }
    __syncthreads();
  }
  // stream the chunks out: one warp per cell at a time, lanes across the chunk (coalesced; adjacent cells are
  // adjacent in HBM for inner rows in natural order)
  const int warp = t >> 5, lane = t & 31, nwarps = THREADS / 32;
  for (int cc = warp; cc < CELLS; cc += nwarps) {
    const int64_t base = sBase[cc];
    if (base < 0) continue;
    const double* src = sJ + (size_t)cc * STRIDE;
    double* dst = Jv + base;
    for (int e = lane; e < CHUNK; e += 32) dst[e] = src[e];
  }
}

}  // namespace dev
}  // namespace pda
