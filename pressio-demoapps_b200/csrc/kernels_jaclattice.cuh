// kernels_jaclattice.cuh -- velocity + Jacobian of the inner rows of a 2D full lattice, every FACE computed once.
//
// Replaces the reference's inner-cell velocity+Jacobian loop (euler_2d_prob_class.hpp:633-720, swe_2d_prob_class.hpp
// :556-640) and its scatter (mixin_directional_flux_balance_jacobian.hpp:142-284), where each cell evaluates both
// faces of every axis (each face's reconstruction gradients and flux Jacobians twice) and every CSR entry is a
// read-modify-write through SparseMatrix::coeffRef.
//
// A lane owns ONE face: the 32 lanes of a warp sit on 32 consecutive faces of one mesh line (a row for the x phase,
// a column for the y phase), i.e. on the left faces of 32 consecutive cells; lane l also owns cell l of the line
// (the last lane of a line only supplies the right face of the last cell).  With
//      t[k][m][j] = JN[k][j] * d(uNeg_j)/d(q_m) + JP[k][j] * d(uPos_j)/d(q_m)         (q_m = m-th cell of the face stencil)
// the block of row-cell c at stencil position P is   hInv * ( t_c[k][P][j] - t_{c+1}[k][P-1][j] ):  the face's own
// products and its right neighbour's, fetched by ONE warp shuffle -- one set of products per face serves both cells.
//   * every (row k, position P) is N contiguous doubles (one 32-byte sector for Euler2d): stored straight to the CSR
//     value array, each entry written once, no shared-memory staging, no memset, no read-modify-write;
//     the self block alone meets both axes: its x part is parked in shared memory until the y phase completes it;
//   * flux Jacobians (2 N^2 doubles) live in a thread-private shared-memory column, the row loop over k is rolled:
//     small code (the instruction cache was the first limiter of the staged kernel) and fewer live registers;
//   * a CTA = 15 x 15 cell tile, a warp carries two lines of 16 faces: phase x = the tile's rows, barrier, phase y = its
//     columns (same code, other stride); the x part of the self blocks and of V waits in shared memory in between.
#pragma once
#include "fastmath.cuh"
#include "kernels_generic.cuh"
#include "kernels_lattice.cuh"

namespace pda {
namespace dev {

template <class Phys, int S>
struct JacLat2d {
  static constexpr int N = Phys::ndpc;
  static constexpr int WARPS = 8;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int NBLK = 1 + 2 * (S - 1);
  static constexpr int ROWLEN = N * NBLK;        // entries per CSR row of an inner cell
  static constexpr int CHUNK = N * ROWLEN;       // doubles per cell (its N rows are consecutive in the CSR arrays)
  // 4 dofs: a block row is one aligned 32-byte sector -> stored straight to HBM (STG.256).  Other sizes (24-, 16-,
  // 8-byte block rows) would turn into partial-sector writes (4x the L2 write transactions): the cell chunks are
  // STAGED in shared memory and streamed out coalesced, a tile row of cells being one contiguous range of the CSR
  // value array.
  static constexpr bool STAGE = (N != 4);
  static constexpr int CSTRIDE = CHUNK | 1;      // odd stride: conflict-free across lanes
  static constexpr size_t kBudget = 200 * 1024;
  static constexpr size_t stagedBytes(int t) { return (size_t)(2 * N * N * THREADS + t * t * CSTRIDE + t * t * N) * sizeof(double); }
  // tile edge in cells; a warp carries two lines of 16 faces (<= 15 cells each)
  static constexpr int T = !STAGE ? 15 : (stagedBytes(15) <= kBudget ? 15 : (stagedBytes(11) <= kBudget ? 11 : 7));
  static constexpr int SELF = N * N + 1;        // (direct mode) padded self-block stride: conflict-free along both tile axes
  static constexpr int oSelf = 2 * N * N * THREADS;           // direct: [T*T][SELF] x-phase part of the self blocks
  static constexpr int oChunk = oSelf;                        // staged: [T*T][CSTRIDE] the cells' chunks
  static constexpr int oV = oSelf + T * T * (STAGE ? CSTRIDE : SELF);   // [T*T][N] x-phase part of the velocity
  static constexpr size_t smemBytes = (size_t)(oV + T * T * N) * sizeof(double);
  // small systems in direct mode fit two CTAs per SM in 128 registers; Euler (96 gradient registers) does not
  static constexpr int MIN_CTAS = (N == 4) ? 2 : 1;
};

struct JacLatTables {
  const int32_t* cellBase;   // [cells] offset of the cell's first CSR row
  const uint4* cellSlots;    // [cells] 16 bytes: block position of graph column c (padded copy of the slot table)
  int32_t rowLen;            // entries per CSR row of an inner cell
};

template <int N> PDA_DEVFN void loadCell(const double* __restrict__ p, double* q) {
  if constexpr (N == 4) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(q[0]), "=d"(q[1]), "=d"(q[2]), "=d"(q[3]) : "l"(p));
  } else if constexpr (N == 2) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    q[0] = v.x; q[1] = v.y;
  } else {
#pragma unroll
    for (int d = 0; d < N; ++d) q[d] = __ldg(p + d);
  }
}
template <int N> PDA_DEVFN void storeBlockRow(double* __restrict__ dst, const double* v) {
  if constexpr (N == 4) {   // one full 32-byte sector per store (STG.256)
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
  } else if constexpr (N == 2) {
    *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) dst[j] = v[j];
  }
}

// one mesh line of one axis: lane f of a 16-lane group = face between cells (a-1, a), a = a0 + f, and owner of
// cell a (f < 15).  `o` = coordinate of the line on the other axis, `cellLocal` = tile-local index of the owned cell.
// `a` may lie one past the inner region (that lane supplies the right face of the last cell): only stencil
// coordinates are clamped, never the face position.
template <class Phys, int S, int AX>
PDA_DEVFN void jacLatLine(const Phys& phys, const LatticeDesc& L, const JacLatTables& jt, double hInv,
                          const double* __restrict__ U, double* __restrict__ V, double* __restrict__ Jv,
                          int a, int o, bool owns, int cellLocal, double* __restrict__ sAll, int tid) {
  constexpr int N = Phys::ndpc;
  constexpr int h = (S - 1) / 2;
  using K = JacLat2d<Phys, S>;
  constexpr int THREADS = K::THREADS;
  double* sJ = sAll + tid;
  const int nx = L.n[0];
  const int nA = L.n[AX], perA = L.per[AX];

  // cells of the face stencil: coordinate a-h+m, m = 0..S-2 (wrapped on periodic axes, clamped for idle lanes)
  double q[S - 1][N];
#pragma unroll
  for (int m = 0; m < S - 1; ++m) {
    int c = a - h + m;
    if (perA) { c %= nA; if (c < 0) c += nA; }   // idle lanes of a ragged tile on a small mesh may wrap twice
    else c = (c < 0) ? 0 : (c >= nA ? nA - 1 : c);
    const int64_t gid = (AX == 0) ? (int64_t)o * nx + c : (int64_t)c * nx + o;
    loadCell<N>(U + gid * N, q[m]);
  }
  // owned cell: gid, CSR base, block slots (one 16-byte load)
  const int64_t gidSelf = (AX == 0) ? (int64_t)o * nx + a : (int64_t)a * nx + o;
  const uint4 sv = owns ? __ldg(jt.cellSlots + gidSelf) : make_uint4(0, 0, 0, 0);
  const int32_t base = owns ? __ldg(jt.cellBase + gidSelf) : 0;
  auto slotOf = [&](int c) -> int {
    const unsigned w = (c < 4) ? sv.x : (c < 8 ? sv.y : (c < 12 ? sv.z : sv.w));
    return (int)((w >> (8 * (c & 3))) & 0xffu);
  };

  // ---- face states, flux, flux Jacobians (parked in this thread's shared-memory column)
  double F[N];
  {
    double un[N], up[N];
#pragma unroll
    for (int d = 0; d < N; ++d) {
      double qd[S - 1];
#pragma unroll
      for (int m = 0; m < S - 1; ++m) qd[m] = q[m][d];
      reconFaceFast<S>(qd, un[d], up[d]);
    }
    double JN[N * N], JP[N * N];
    faceFlux2d<Phys, AX>(phys, un, up, F);
    faceFluxJac2d<Phys, AX>(phys, un, up, JN, JP);
#pragma unroll
    for (int e = 0; e < N * N; ++e) { sJ[e * THREADS] = hInv * JN[e]; sJ[(N * N + e) * THREADS] = hInv * JP[e]; }
  }
  // ---- reconstruction gradients of every dof
  double gN[N][S - 1], gP[N][S - 1];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double qd[S - 1];
#pragma unroll
    for (int m = 0; m < S - 1; ++m) qd[m] = q[m][j];
    reconFaceGradFast<S>(qd, gN[j], gP[j]);
  }

  int sl[S];
  sl[h] = slotOf(0);
#pragma unroll
  for (int l = 0; l < h; ++l) {
    sl[h - 1 - l] = slotOf(gcol<2>(sideMinus<AX>(), l));
    sl[h + 1 + l] = slotOf(gcol<2>(sidePlus<AX>(), l));
  }
  double* jBase = Jv + base;
  const int rowLen = jt.rowLen;
  double* sSelf = sAll + K::oSelf + cellLocal * K::SELF;
  double* sChunk = sAll + K::oChunk + cellLocal * K::CSTRIDE;
  double* sV = sAll + K::oV + cellLocal * N;

  // ---- velocity: hInv (F_left - F_right); x phase parks its part in shared memory, y phase completes and stores
  {
    double v[N];
#pragma unroll
    for (int d = 0; d < N; ++d) v[d] = hInv * (F[d] - __shfl_down_sync(0xffffffffu, F[d], 1));
    if (owns) {
      if (AX == 0) {
#pragma unroll
        for (int d = 0; d < N; ++d) sV[d] = v[d];
      } else {
#pragma unroll
        for (int d = 0; d < N; ++d) v[d] += sV[d];
        if constexpr (PhysTraits<Phys>::hasDiffusion) {
          // first-layer neighbours as a graph row (left, front, right, back): inner cells have them all
          const int i = o, jrow = a, ny = L.n[1];
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
        if (V) {
          double* out = V + gidSelf * N;
#pragma unroll
          for (int d = 0; d < N; ++d) out[d] = v[d];
        }
      }
    }
  }

  // ---- Jacobian rows k = 0..N-1 (rolled), positions and dofs unrolled
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    double jn[N], jp[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { jn[j] = sJ[(k * N + j) * THREADS]; jp[j] = sJ[(N * N + k * N + j) * THREADS]; }
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
        if constexpr (K::STAGE) {
          double* dst = sChunk + k * K::ROWLEN + sl[P] * N;
          if (P == h && AX != 0) {
#pragma unroll
            for (int j = 0; j < N; ++j) dst[j] += val[j];   // self block: x part already there (x, then y)
          } else {
#pragma unroll
            for (int j = 0; j < N; ++j) dst[j] = val[j];
          }
        } else if (P == h && AX == 0) {   // self block: x part waits in shared memory for the y part (the reference's order)
#pragma unroll
          for (int j = 0; j < N; ++j) sSelf[k * N + j] = val[j];
        } else {
          if (P == h) {
#pragma unroll
            for (int j = 0; j < N; ++j) val[j] += sSelf[k * N + j];
          }
          storeBlockRow<N>(rowp + sl[P] * N, val);
        }
      }
    }
  }
  // ---- point terms and diffusion (after both flux phases: the reference adds them last); a few entries of rows
  //      this tile has just written (same thread, or x-phase threads before the barrier): read-modify-write in L2
  if (AX != 0 && owns) {
    uint8_t slots[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) slots[c] = (uint8_t)slotOf(c);
    addExtraJacInner<Phys>(phys, U + gidSelf * N, slots, [&](int k, int slot, int j, double x) {
      if constexpr (K::STAGE) sChunk[k * K::ROWLEN + slot * N + j] += x;
      else jBase[(int64_t)k * rowLen + slot * N + j] += x;
    });
  }
}

template <class Phys, int S>
__global__ void __launch_bounds__(JacLat2d<Phys, S>::THREADS, JacLat2d<Phys, S>::MIN_CTAS)
k_jacobian_lattice2d(Phys phys, LatticeDesc L, Deltas dl, JacLatTables jt, const double* __restrict__ U,
                     double* __restrict__ V, double* __restrict__ Jv) {
  using K = JacLat2d<Phys, S>;
  extern __shared__ __align__(16) double sAll[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int line = 2 * warp + (lane >> 4), f = lane & 15;
  const int lo0 = L.per[0] ? 0 : L.meshHalo, hi0 = L.per[0] ? L.n[0] : L.n[0] - L.meshHalo;
  const int lo1 = L.per[1] ? 0 : L.meshHalo, hi1 = L.per[1] ? L.n[1] : L.n[1] - L.meshHalo;
  const int I0 = lo0 + K::T * blockIdx.x, J0 = lo1 + K::T * blockIdx.y;
  {   // phase x: line = row J0+line, lanes along x
    const int j = J0 + line, a = I0 + f;
    const bool owns = (line < K::T) && (f < K::T) && (j < hi1) && (a < hi0);
    jacLatLine<Phys, S, 0>(phys, L, jt, dl.hInv[0], U, V, Jv, a, min(j, hi1 - 1), owns,
                           min(line, K::T - 1) * K::T + min(f, K::T - 1), sAll, tid);
  }
  __syncthreads();
  {   // phase y: line = column I0+line, lanes along y
    const int i = I0 + line, a = J0 + f;
    const bool owns = (line < K::T) && (f < K::T) && (i < hi0) && (a < hi1);
    jacLatLine<Phys, S, 1>(phys, L, jt, dl.hInv[1], U, V, Jv, a, min(i, hi0 - 1), owns,
                           min(f, K::T - 1) * K::T + min(line, K::T - 1), sAll, tid);
  }
  if constexpr (K::STAGE) {
    // stream the staged chunks out: the cells of a tile row are consecutive inner cells, i.e. ONE contiguous range of
    // the CSR value array -> lanes across it, coalesced, every value written once
    __syncthreads();
    const int nvalid = min(K::T, hi0 - I0);
    for (int r = warp; r < K::T; r += K::WARPS) {
      const int j = J0 + r;
      if (j >= hi1) break;
      double* dst = Jv + jt.cellBase[(int64_t)j * L.n[0] + I0];
      const double* src = sAll + K::oChunk + (size_t)(r * K::T) * K::CSTRIDE;
      for (int e = lane; e < nvalid * K::CHUNK; e += 32) {
        const int c = e / K::CHUNK, w = e - c * K::CHUNK;
        dst[e] = src[c * K::CSTRIDE + w];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// First-order scheme on 2D full lattices: velocity + Jacobian by Y-MARCHING warps (the layout of k_velocity_march2d).
//
// With a first-order reconstruction the face states are the two adjacent cells, so a block of the Jacobian IS a scaled
// flux Jacobian: for cell (i,j) with left/right faces l, r and back/front faces b, f
//     left nb  hx JN(l)      right nb  -hx JP(r)      back nb  hy JN(b)      front nb  -hy JP(f)
//     self     hx (JP(l) - JN(r))  +  hy (JP(b) - JN(f))                       (x part, then y part: the reference's order)
// (swe_2d_prob_class.hpp:556-640, euler_2d_prob_class.hpp:633-720 with the first-order functor; scatter
// mixin_directional_flux_balance_jacobian.hpp:142-284).  The kernel is bound by the stores (45 doubles out per 3 in for
// shallow water), and the tile kernel above reaches 25 % of the HBM peak on it (one CTA of 8 warps per SM: its tile of
// staged chunks fills the shared memory).  Here:
//   * a warp owns a strip of 30 columns (+ one feeder lane per side) and marches along y; every face is evaluated ONCE:
//     a lane computes the left face of its cell and receives the right face from lane+1 by shuffle; the front face of
//     row j is carried in registers as the back face of row j+1;
//   * the lanes assemble their cells' chunks in a per-warp shared-memory slab (32 x CHUNK, odd stride) and the warp
//     streams the 30 chunks -- consecutive inner cells = ONE contiguous range of the CSR value array -- with coalesced
//     stores, every value written once (no memset, no read-modify-write);
//   * no CTA barrier; 4 warps per CTA, several CTAs per SM.
template <class Phys>
struct JacMarchFo {
  static constexpr int N = Phys::ndpc;
  static constexpr int ROWLEN = N * 5;
  static constexpr int CHUNK = N * ROWLEN;
  static constexpr int STRIDE = CHUNK | 1;
  static constexpr int WARPS = 4;
  static constexpr int W = 30;
  static constexpr size_t smemBytes = (size_t)WARPS * 32 * STRIDE * sizeof(double);
  // CTAs per SM the register allocation is capped for: shallow water at 4 (128 registers) spills 108 bytes and runs
  // 2.03 ms at 4096^2, at 3 (167 registers, no spills, 12 warps/SM) 1.71 ms
  static constexpr int MIN_CTAS = (N <= 2) ? 4 : (N == 3 ? 3 : 2);
};

template <class Phys, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_jacobian_march2d_fo(Phys phys, LatticeDesc L, Deltas dl, JacLatTables jt, const double* __restrict__ U,
                      double* __restrict__ V, double* __restrict__ Jv, int LY) {
  using K = JacMarchFo<Phys>;
  constexpr int N = K::N, ROWLEN = K::ROWLEN, CHUNK = K::CHUNK, STRIDE = K::STRIDE, W = K::W;
  extern __shared__ __align__(16) double sAll[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const slab = sAll + (size_t)warp * 32 * STRIDE;
  double* const mine = slab + lane * STRIDE;
  const int nx = L.n[0], ny = L.n[1];
  const int lo0 = L.per[0] ? 0 : L.meshHalo, hi0 = L.per[0] ? nx : nx - L.meshHalo;
  const int yb = L.per[1] ? 0 : L.meshHalo, ye = L.per[1] ? ny : ny - L.meshHalo;
  const int nStrips = (hi0 - lo0 + W - 1) / W;
  const int wid = blockIdx.x * K::WARPS + warp;
  const int strip = wid % nStrips, chunk = wid / nStrips;
  const int j0 = yb + chunk * LY;
  if (j0 >= ye) return;
  const int j1 = min(j0 + LY, ye);
  const int xFirst = lo0 + strip * W;          // first output column of the strip (lane 1)
  const int x = xFirst - 1 + lane;
  int xc = x;
  if (L.per[0]) { xc %= nx; if (xc < 0) xc += nx; }
  else xc = (xc < 0) ? 0 : (xc >= nx ? nx - 1 : xc);
  const bool outLane = (lane >= 1) && (lane <= W) && (x < hi0);
  const int nOut = min(W, hi0 - xFirst);       // output cells of this strip
  const double hx = dl.hInv[0], hy = dl.hInv[1];

  auto rowIdx = [&](int r) -> int {
    if (L.per[1]) { r %= ny; if (r < 0) r += ny; return r; }
    return (r < 0) ? 0 : (r >= ny ? ny - 1 : r);
  };
  auto rowPtr = [&](int r) -> const double* { return U + ((int64_t)rowIdx(r) * nx + xc) * N; };

  double qc[N], qn[N];                         // rows j and j+1 of this lane's column
  loadCell<N>(rowPtr(j0 - 1), qc);
  loadCell<N>(rowPtr(j0), qn);
  double FyB[N];                               // flux through the back face of the current row
#pragma unroll
  for (int d = 0; d < N; ++d) FyB[d] = 0.0;
  auto slotsOf = [&](int j) -> uint4 { return outLane ? __ldg(jt.cellSlots + (int64_t)j * nx + x) : make_uint4(0, 0, 0, 0); };
  auto slotIn = [](const uint4& sv, int c) -> int {
    const unsigned w = (c < 4) ? sv.x : (c < 8 ? sv.y : (c < 12 ? sv.z : sv.w));
    return (int)((w >> (8 * (c & 3))) & 0xffu);
  };
  uint4 sv = make_uint4(0, 0, 0, 0);           // block slots of the current row's cell (set by the previous step)

  for (int j = j0 - 1; j < j1; ++j) {
    const bool ghost = (j < j0);
    double nxt[N];
    const bool more = (j + 1 < j1);
    if (more) loadCell<N>(rowPtr(j + 2), nxt);           // lands while this row is computed
    // ---- front y face (j+1/2): uNeg = row j, uPos = row j+1
    double FyF[N], JNf[N * N], JPf[N * N];
    {
      double a[N], b[N];
#pragma unroll
      for (int d = 0; d < N; ++d) { a[d] = qc[d]; b[d] = qn[d]; }
      faceFlux2d<Phys, 1>(phys, a, b, FyF);
      faceFluxJac2d<Phys, 1>(phys, a, b, JNf, JPf);
#pragma unroll
      for (int e = 0; e < N * N; ++e) { JNf[e] *= hy; JPf[e] *= hy; }
    }
    if (!ghost) {
      const int64_t gidSelf = (int64_t)j * nx + x;
      const int s0 = slotIn(sv, 0) * N, sL = slotIn(sv, 1) * N, sF = slotIn(sv, 2) * N, sR = slotIn(sv, 3) * N;
      // ---- x left face of this lane's cell: uNeg = lane-1's cell, uPos = mine.  The back-face terms of this row (back
      //      block, hy JP(b) in the self block) are already in the slab: the previous step stored them.
      double Fx[N], v[N];
      {
        double a[N], b[N], JNx[N * N], JPx[N * N];
#pragma unroll
        for (int d = 0; d < N; ++d) { b[d] = qc[d]; a[d] = __shfl_up_sync(0xffffffffu, qc[d], 1); }
        faceFlux2d<Phys, 0>(phys, a, b, Fx);
        faceFluxJac2d<Phys, 0>(phys, a, b, JNx, JPx);
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
          for (int jj = 0; jj < N; ++jj) {
            const double jn = hx * JNx[k * N + jj], jp = hx * JPx[k * N + jj];
            const double jnR = __shfl_down_sync(0xffffffffu, jn, 1), jpR = __shfl_down_sync(0xffffffffu, jp, 1);
            mine[k * ROWLEN + sL + jj] = jn;
            mine[k * ROWLEN + sR + jj] = -jpR;
            mine[k * ROWLEN + sF + jj] = -JPf[k * N + jj];
            double* self = mine + k * ROWLEN + s0 + jj;
            *self = (jp - jnR) + (*self - JNf[k * N + jj]);     // x part, then y part (the reference's order)
          }
      }
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const double FxR = __shfl_down_sync(0xffffffffu, Fx[d], 1);
        v[d] = hx * (Fx[d] - FxR);
        v[d] += hy * (FyB[d] - FyF[d]);
      }
      if (outLane) {
        if constexpr (PhysTraits<Phys>::hasDiffusion) {
          auto wrapI = [&](int c, int n) { return c < 0 ? c + n : (c >= n ? c - n : c); };
          int32_t row5[5];
          row5[0] = (int32_t)gidSelf;
          row5[1] = j * nx + wrapI(x - 1, nx);
          row5[2] = wrapI(j + 1, ny) * nx + x;
          row5[3] = j * nx + wrapI(x + 1, nx);
          row5[4] = wrapI(j - 1, ny) * nx + x;
          addDiffusionInner<Phys>(phys, row5, U, v);
        }
        addForcing<Phys>(phys, qc, v, (int32_t)gidSelf);
        if (V) {
          double* out = V + gidSelf * N;
#pragma unroll
          for (int d = 0; d < N; ++d) out[d] = v[d];
        }
        uint8_t slots[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) slots[c] = (uint8_t)slotIn(sv, c);
        addExtraJacInner<Phys>(phys, qc, slots, [&](int k, int slot, int jj, double val) {
          mine[k * ROWLEN + slot * N + jj] += val;
        });
      }
      __syncwarp();
      // ---- stream the strip's chunks out (lanes 1..nOut): one contiguous range of the CSR value array
      if (nOut > 0) {
        double* dst = Jv + __ldg(jt.cellBase + (int64_t)j * nx + xFirst);
        const int total = nOut * CHUNK;
        for (int e = lane; e < total; e += 32) {
          const int c = e / CHUNK, w = e - c * CHUNK;
          dst[e] = slab[(c + 1) * STRIDE + w];
        }
      }
      __syncwarp();
    }
    // ---- this front face is the back face of row j+1: its terms go into the (now free) slab row right away, so no
    //      flux Jacobian is carried in registers across steps
    if (more) {
      sv = slotsOf(j + 1);
      const int s0 = slotIn(sv, 0) * N, sB = slotIn(sv, 4) * N;
#pragma unroll
      for (int k = 0; k < N; ++k)
#pragma unroll
        for (int jj = 0; jj < N; ++jj) {
          mine[k * ROWLEN + sB + jj] = JNf[k * N + jj];
          mine[k * ROWLEN + s0 + jj] = JPf[k * N + jj];
        }
    }
#pragma unroll
    for (int d = 0; d < N; ++d) { FyB[d] = FyF[d]; qc[d] = qn[d]; }
    if (more) {
#pragma unroll
      for (int d = 0; d < N; ++d) qn[d] = nxt[d];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// WENO3 / WENO5 on the SMALL systems (1-3 dofs per cell): the y-marching layout of k_jacobian_march2d_fo with the chain rule
// in it.  With  t_f[k][m][j] = h ( JN[k][j] d(uNeg_j)/d(q_m) + JP[k][j] d(uPos_j)/d(q_m) )  for face f and the m-th cell of its
// stencil, the block of a cell at stencil position P of an axis is  t_lower[k][P][j] - t_upper[k][P-1][j]  (lower = left /
// back face, upper = right / front face of the cell) -- same algebra as jacLatLine above.  Per step and lane: ONE x face
// (the right face's products arrive from lane+1 by shuffle) and ONE y face (the front face; its products are the back-face
// terms of row j+1 and are written into the lane's slab row for the next step as soon as the current row has been streamed
// out).  Chunks are assembled in a per-warp slab and streamed out coalesced; no tile, no CTA barrier, CTAs of 2 warps so that
// several fit next to each other (the tile kernel keeps ONE 8-warp CTA per SM: its 15 x 15 tile of chunks fills the shared memory).
template <class Phys, int S>
struct JacMarchWeno {
  static constexpr int N = Phys::ndpc;
  static constexpr int h = (S - 1) / 2;
  static constexpr int NBLK = 1 + 2 * (S - 1);
  static constexpr int ROWLEN = N * NBLK;
  static constexpr int CHUNK = N * ROWLEN;
  static constexpr int STRIDE = CHUNK | 1;
  static constexpr int WARPS = 2;
  static constexpr int W = 32 - 2 * h;
  static constexpr size_t smemBytes = (size_t)WARPS * 32 * STRIDE * sizeof(double);
  static constexpr int FIT = (int)(220 * 1024 / smemBytes);
  static constexpr int MIN_CTAS = FIT > 5 ? 5 : (FIT < 1 ? 1 : FIT);   // 5 CTAs of 64 threads: 204 registers
};

template <class Phys, int S>
__global__ void __launch_bounds__(32 * JacMarchWeno<Phys, S>::WARPS, JacMarchWeno<Phys, S>::MIN_CTAS)
k_jacobian_march2d_weno(Phys phys, LatticeDesc L, Deltas dl, JacLatTables jt, const double* __restrict__ U,
                        double* __restrict__ V, double* __restrict__ Jv, int LY) {
  using K = JacMarchWeno<Phys, S>;
  constexpr int N = K::N, h = K::h, ROWLEN = K::ROWLEN, CHUNK = K::CHUNK, STRIDE = K::STRIDE, W = K::W;
  constexpr int M = 2 * h;            // cells per face stencil
  constexpr int R = 2 * h + 1;        // ring rows j-h .. j+h
  extern __shared__ __align__(16) double sAll[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const slab = sAll + (size_t)warp * 32 * STRIDE;
  double* const mine = slab + lane * STRIDE;
  const int nx = L.n[0], ny = L.n[1];
  const int lo0 = L.per[0] ? 0 : L.meshHalo, hi0 = L.per[0] ? nx : nx - L.meshHalo;
  const int yb = L.per[1] ? 0 : L.meshHalo, ye = L.per[1] ? ny : ny - L.meshHalo;
  const int nStrips = (hi0 - lo0 + W - 1) / W;
  const int wid = blockIdx.x * K::WARPS + warp;
  const int strip = wid % nStrips, chunk = wid / nStrips;
  const int j0 = yb + chunk * LY;
  if (j0 >= ye) return;
  const int j1 = min(j0 + LY, ye);
  const int xFirst = lo0 + strip * W;          // first output column of the strip (lane h)
  const int x = xFirst - h + lane;
  int xc = x;
  if (L.per[0]) { xc %= nx; if (xc < 0) xc += nx; }
  else xc = (xc < 0) ? 0 : (xc >= nx ? nx - 1 : xc);
  const bool outLane = (lane >= h) && (lane <= 31 - h) && (x < hi0);
  const int nOut = min(W, hi0 - xFirst);
  const double hx = dl.hInv[0], hy = dl.hInv[1];

  auto rowPtr = [&](int r) -> const double* {
    if (L.per[1]) { r %= ny; if (r < 0) r += ny; }
    else r = (r < 0) ? 0 : (r >= ny ? ny - 1 : r);
    return U + ((int64_t)r * nx + xc) * N;
  };
  auto slotsOf = [&](int j) -> uint4 { return outLane ? __ldg(jt.cellSlots + (int64_t)j * nx + x) : make_uint4(0, 0, 0, 0); };
  auto slotIn = [](const uint4& sv, int c) -> int {
    const unsigned w = (c < 4) ? sv.x : (c < 8 ? sv.y : (c < 12 ? sv.z : sv.w));
    return (int)((w >> (8 * (c & 3))) & 0xffu);
  };
  // slab offset of the block at stencil position P of axis AX (self at P = h)
  auto posSlot = [&](const uint4& sv, int ax, int P) -> int {
    const int sm = (ax == 0) ? 0 : 3, sp = (ax == 0) ? 2 : 1;
    const int c = (P == h) ? 0 : (P < h ? gcol<2>(sm, h - 1 - P) : gcol<2>(sp, P - h - 1));
    return slotIn(sv, c) * N;
  };

  double q[R][N];
#pragma unroll
  for (int i = 0; i < R; ++i) loadCell<N>(rowPtr(j0 - 1 - h + i), q[i]);
  double FyB[N];
#pragma unroll
  for (int d = 0; d < N; ++d) FyB[d] = 0.0;
  uint4 sv = make_uint4(0, 0, 0, 0);

  for (int j = j0 - 1; j < j1; ++j) {
    const bool ghost = (j < j0);
    double nxt[N];
    const bool more = (j + 1 < j1);
    if (more) loadCell<N>(rowPtr(j + 1 + h), nxt);
    // ---- x left face of this lane's cell FIRST (stencil from the neighbouring lanes' row-j values): its temporaries are
    //      dead before the y face is evaluated, whose products must stay live until the row has been streamed out
    double Fx[N];
    if (!ghost) {
      double uN[N], uP[N], gN[N][M], gP[N][M];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        double sq[M];
#pragma unroll
        for (int o = 0; o < M; ++o)
          sq[o] = (o == h) ? q[h][d] : __shfl_sync(0xffffffffu, q[h][d], (lane + o - h) & 31);
        reconFaceValGradFast<S>(sq, uN[d], uP[d], gN[d], gP[d]);
      }
      double JN[N * N], JP[N * N];
      faceFlux2d<Phys, 0>(phys, uN, uP, Fx);
      faceFluxJac2d<Phys, 0>(phys, uN, uP, JN, JP);
      int sx[R];
#pragma unroll
      for (int P = 0; P < R; ++P) sx[P] = posSlot(sv, 0, P);
#pragma unroll
      for (int k = 0; k < N; ++k)
#pragma unroll
        for (int jj = 0; jj < N; ++jj) {
          const double jn = hx * JN[k * N + jj], jp = hx * JP[k * N + jj];
          double own[M], rgt[M];
#pragma unroll
          for (int m = 0; m < M; ++m) {
            own[m] = jn * gN[jj][m] + jp * gP[jj][m];
            rgt[m] = __shfl_down_sync(0xffffffffu, own[m], 1);
          }
#pragma unroll
          for (int P = 0; P < R; ++P) {
            const double val = ((P < M) ? own[P] : 0.0) - ((P >= 1) ? rgt[P - 1] : 0.0);
            double* dst = mine + k * ROWLEN + sx[P] + jj;
            if (P == h) *dst += val;      // self block: the back-face term of the y axis is already in the slab
            else *dst = val;
          }
        }
    }
    // ---- front y face (j+1/2): rows j-h+1 .. j+h = ring rows 1 .. 2h
    double FyF[N], tF[N * N][M];
    {
      double uN[N], uP[N], gN[N][M], gP[N][M];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        double sq[M];
#pragma unroll
        for (int o = 0; o < M; ++o) sq[o] = q[1 + o][d];
        reconFaceValGradFast<S>(sq, uN[d], uP[d], gN[d], gP[d]);
      }
      double JN[N * N], JP[N * N];
      faceFlux2d<Phys, 1>(phys, uN, uP, FyF);
      faceFluxJac2d<Phys, 1>(phys, uN, uP, JN, JP);
#pragma unroll
      for (int k = 0; k < N; ++k)
#pragma unroll
        for (int jj = 0; jj < N; ++jj) {
          const double jn = hy * JN[k * N + jj], jp = hy * JP[k * N + jj];
#pragma unroll
          for (int m = 0; m < M; ++m) tF[k * N + jj][m] = jn * gN[jj][m] + jp * gP[jj][m];
        }
    }
    if (!ghost) {
      const int64_t gidSelf = (int64_t)j * nx + x;
      // ---- y blocks: the back-face terms t_b[P] (P = 0..2h-1) are already in the slab; position P gets - t_f[P-1]
#pragma unroll
      for (int P = 1; P <= M; ++P) {
        const int sl = posSlot(sv, 1, P);
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
          for (int jj = 0; jj < N; ++jj) {
            double* dst = mine + k * ROWLEN + sl + jj;
            if (P == M) *dst = -tF[k * N + jj][M - 1];
            else *dst -= tF[k * N + jj][P - 1];
          }
      }
      double v[N];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const double FxR = __shfl_down_sync(0xffffffffu, Fx[d], 1);
        v[d] = hx * (Fx[d] - FxR);
        v[d] += hy * (FyB[d] - FyF[d]);
      }
      if (outLane) {
        if constexpr (PhysTraits<Phys>::hasDiffusion) {
          auto wrapI = [&](int c, int n) { return c < 0 ? c + n : (c >= n ? c - n : c); };
          int32_t row5[5];
          row5[0] = (int32_t)gidSelf;
          row5[1] = j * nx + wrapI(x - 1, nx);
          row5[2] = wrapI(j + 1, ny) * nx + x;
          row5[3] = j * nx + wrapI(x + 1, nx);
          row5[4] = wrapI(j - 1, ny) * nx + x;
          addDiffusionInner<Phys>(phys, row5, U, v);
        }
        addForcing<Phys>(phys, q[h], v, (int32_t)gidSelf);
        if (V) {
          double* out = V + gidSelf * N;
#pragma unroll
          for (int d = 0; d < N; ++d) out[d] = v[d];
        }
        uint8_t slots[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) slots[c] = (uint8_t)slotIn(sv, c);
        addExtraJacInner<Phys>(phys, q[h], slots, [&](int k, int slot, int jj, double val) {
          mine[k * ROWLEN + slot * N + jj] += val;
        });
      }
      __syncwarp();
      if (nOut > 0) {
        double* dst = Jv + __ldg(jt.cellBase + (int64_t)j * nx + xFirst);
        const int total = nOut * CHUNK;
        for (int e = lane; e < total; e += 32) {
          const int c = e / CHUNK, w = e - c * CHUNK;
          dst[e] = slab[(c + h) * STRIDE + w];
        }
      }
      __syncwarp();
    }
    // ---- this front face is the back face of row j+1: its products go into the (now free) slab row, positions 0..2h-1
    if (more) {
      sv = slotsOf(j + 1);
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const int sl = posSlot(sv, 1, m);
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
          for (int jj = 0; jj < N; ++jj) mine[k * ROWLEN + sl + jj] = tF[k * N + jj][m];
      }
    }
#pragma unroll
    for (int d = 0; d < N; ++d) FyB[d] = FyF[d];
#pragma unroll
    for (int i = 0; i < R - 1; ++i)
#pragma unroll
      for (int d = 0; d < N; ++d) q[i][d] = q[i + 1][d];
    if (more) {
#pragma unroll
      for (int d = 0; d < N; ++d) q[R - 1][d] = nxt[d];
    }
  }
}

}  // namespace dev
}  // namespace pda
