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
};

// one (axis AX, face) role of one cell; `face` 0 = left/back/bottom face, 1 = right/front/top
template <class Phys, int S, int AX>
PDA_DEVFN void jacobianFaceRole(const Phys& phys, const int32_t* __restrict__ row, const double* __restrict__ U,
                                double hInv, int face, double* __restrict__ my, const uint8_t* __restrict__ slots,
                                double* selfAcc /*[N*N], L lane only*/, double* dF /*[N] F_L - F_R, L lane only*/) {
  constexpr int N = Phys::ndpc;
  constexpr int h = (S - 1) / 2;
  constexpr int DIM = Phys::dim;
  constexpr int ROWLEN = JacStage<Phys, S>::ROWLEN;
  int32_t cells[S];
  stencilCells<DIM, S, AX>(row, cells);
  int sl[S];   // block slot of each stencil position
  sl[h] = slots[0];
#pragma unroll
  for (int L = 0; L < h; ++L) {
    sl[h - 1 - L] = slots[gcol<DIM>(sideMinus<AX>(), L)];
    sl[h + 1 + L] = slots[gcol<DIM>(sidePlus<AX>(), L)];
  }
  const double sgn = (face == 0) ? hInv : -hInv;
  double un[N], up[N];
#pragma unroll
  for (int d = 0; d < N; ++d) {
    double q[S - 1];
#pragma unroll
    for (int p = 0; p < S - 1; ++p) q[p] = U[(int64_t)cells[p + face] * N + d];
    reconFaceFast<S>(q, un[d], up[d]);
  }
  double F[N], JN[N * N], JP[N * N];
  faceFlux2d<Phys, AX>(phys, un, up, F);
  faceFluxJac2d<Phys, AX>(phys, un, up, JN, JP);
#pragma unroll
  for (int d = 0; d < N; ++d) {
    const double other = __shfl_xor_sync(0xffffffffu, F[d], 1);
    dF[d] = F[d] - other;   // meaningful in the L lane: F_L - F_R
  }
  // chain rule: d(flux)/d(u_p) = JN * diag(d uNeg/d u_p) + JP * diag(d uPos/d u_p).  The L lane owns stencil
  // positions 0..h, the R lane h+1..S-1; position p receives the L face's column m = p and the R face's m = p-1.
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double q[S - 1], gN[S - 1], gP[S - 1];
#pragma unroll
    for (int p = 0; p < S - 1; ++p) q[p] = U[(int64_t)cells[p + face] * N + j];
    reconFaceGradFast<S>(q, gN, gP);
#pragma unroll
    for (int s = 0; s <= h; ++s) {
      constexpr int dummy = 0; (void)dummy;
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
        const double jn = sgn * JN[k * N + j], jp = sgn * JP[k * N + j];
        const double mine = jn * gNm + jp * gPm;
        const double theirs = jn * gNt + jp * gPt;
        const double got = __shfl_xor_sync(0xffffffffu, theirs, 1);
        const double val = mine + got;
        if (face == 0) {
          if (pOwnL == h) selfAcc[k * N + j] = val;                       // self block: stored in axis order later
          else my[k * ROWLEN + sl[pOwnL] * N + j] = val;
        } else if (rHas) {
          my[k * ROWLEN + sl[pOwnR] * N + j] = val;
        }
      }
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

template <class Phys, int S>
__global__ void __launch_bounds__(JacStage<Phys, S>::THREADS)
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
  if (axis == 0) jacobianFaceRole<Phys, S, 0>(phys, row, U, dl.hInv[0], face, my, slots, selfAcc, dF);
  if constexpr (DIM >= 2) { if (axis == 1) jacobianFaceRole<Phys, S, 1>(phys, row, U, dl.hInv[1], face, my, slots, selfAcc, dF); }
  if constexpr (DIM >= 3) { if (axis == 2) jacobianFaceRole<Phys, S, 2>(phys, row, U, dl.hInv[2], face, my, slots, selfAcc, dF); }
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
