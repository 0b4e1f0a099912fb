// kernels_generic.cuh -- graph-driven kernels: one thread per cell, neighbours gathered through the connectivity
// graph (sample meshes, near-boundary rows of any mesh, and every problem on small meshes).
//
// Reference loop nests replaced: ghost fill (euler_2d_prob_class.hpp:465-565 + fillers), inner-cell
// velocity (:991-1044), near-boundary velocity (:1046-1113), inner velocity+Jacobian (:633-720 + scatter
// mixin_directional_flux_balance_jacobian.hpp:142-284), near-boundary first-order Jacobian (:723-989, :287-371),
// Gray-Scott (diffusion_reaction_2d_prob_class.hpp:459-529), SWE Coriolis (swe_2d_prob_class.hpp:984-1012).
#pragma once
#include <cstdint>

#include <type_traits>

#include "fastmath.cuh"
#include "kernel_types.cuh"
#include "physics.cuh"

namespace pda {
namespace dev {

// graph column of layer L on side s; device twin of pda::graphCol
template <int DIM> PDA_DEVFN int gcol(int side, int layer) {
  if (DIM == 1) return 1 + 2 * layer + (side == 2 ? 1 : 0);
  return 1 + (DIM == 2 ? 4 : 6) * layer + side;
}
template <int AX> PDA_DEVFN constexpr int sideMinus() { return AX == 0 ? 0 : (AX == 1 ? 3 : 4); }
template <int AX> PDA_DEVFN constexpr int sidePlus() { return AX == 0 ? 2 : (AX == 1 ? 1 : 5); }

// ------------------------------------------------------------------------------------------------ ghost fill
// One recipe per (near-bd row, side, layer): ghost[d] = mul[mode][d] * U[src*ndpc+d] + add[mode][d].
// kind 1 (double Mach reflection, top wall): mode is chosen at run time from the shock position at time t
// (euler_2d_ghost_filler_double_mach_reflection.hpp:112-125,176-189,...).
struct GhostRecipe {
  int32_t src;     // source cell (stencil-mesh id); -1 = ghost not needed (neighbour exists)
  int16_t mode;    // index into the mul/add tables
  int16_t kind;    // 0 affine, 1 DMR top wall
};
constexpr int kMaxGhostModes = 12;
struct GhostTables {
  double mul[kMaxGhostModes][5];
  double add[kMaxGhostModes][5];
  // DMR: dist = x - wedge - speed*t - slope*(y + (layer+1)*dy) < 0 -> modePost else modePre
  double dmrWedge, dmrSpeed, dmrSlope, dy;
  int32_t dmrModePost, dmrModePre;
};

template <int NDPC>
__global__ void k_ghost_fill(const GhostRecipe* __restrict__ rec, const double2* __restrict__ nbXY, int32_t nNb,
                             int nsides, int h, GhostTables tab, const double* __restrict__ U, GhostView gv,
                             double t) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)nNb * nsides * h;
  if (tid >= total) return;
  const int layer = (int)(tid % h);
  const int side = (int)((tid / h) % nsides);
  const int32_t r = (int32_t)(tid / ((int64_t)h * nsides));
  const GhostRecipe gr = rec[tid];
  if (gr.src < 0) return;
  int mode = gr.mode;
  if (gr.kind == 1) {
    const double2 xy = nbXY[r];
    const double yIn = __dadd_rn(xy.y, __dmul_rn((double)(layer + 1), tab.dy));
    double dist = __dsub_rn(xy.x, tab.dmrWedge);
    dist = __dsub_rn(dist, __dmul_rn(tab.dmrSpeed, t));
    dist = __dsub_rn(dist, __dmul_rn(tab.dmrSlope, yIn));
    mode = (dist < 0.0) ? tab.dmrModePost : tab.dmrModePre;
  }
  double* out = gv.g[side] + (int64_t)r * gv.stride + layer * NDPC;
  const double* in = U + (int64_t)gr.src * NDPC;
#pragma unroll
  for (int d = 0; d < NDPC; ++d) {
    // constant components are ASSIGNED like the reference does (0 * NaN would poison a Dirichlet ghost)
    const double mul = tab.mul[mode][d];
    out[d] = (mul == 0.0) ? tab.add[mode][d] : mul * in[d] + tab.add[mode][d];
  }
}

// ------------------------------------------------------------------------------------------------ helpers
// value of dof d in the stencil cell `c` (or the ghost of `layer` on `side` when c == -1)
template <int NDPC, bool NEARBD>
PDA_DEVFN double stencilVal(const double* __restrict__ U, int32_t c, const GhostView& gv, int32_t nbRow, int side,
                            int layer, int d) {
  if (NEARBD && c < 0) return gv.g[side][(int64_t)nbRow * gv.stride + layer * NDPC + d];
  return U[(int64_t)c * NDPC + d];
}

// gather the S stencil cell ids of one axis: pos 0..h-1 = minus layers (far..near), h = self, h+1.. = plus layers
template <int DIM, int S, int AX>
PDA_DEVFN void stencilCells(const int32_t* __restrict__ row, int32_t* cells) {
  constexpr int h = (S - 1) / 2;
  cells[h] = row[0];
#pragma unroll
  for (int L = 0; L < h; ++L) {
    cells[h - 1 - L] = row[gcol<DIM>(sideMinus<AX>(), L)];
    cells[h + 1 + L] = row[gcol<DIM>(sidePlus<AX>(), L)];
  }
}

// accumulate hInv*(F_L - F_R) of axis AX into v[].  Problems with a diffusion term also get dD[AX]*(u+ - 2u + u-):
// near-boundary rows add it right after the axis' flux balance, inner rows after all flux balances (`diff` receives
// the term) -- the two accumulation orders of advection_diffusion_2d_prob_class.hpp:959-1000 / 1050-1056,1190-1194.
template <class Phys, int S, int AX, bool NEARBD>
PDA_DEVFN void axisVelocity(const Phys& phys, const int32_t* __restrict__ row, const double* __restrict__ U,
                            const GhostView& gv, int32_t nbRow, double hInv, double* v, double* diff) {
  constexpr int N = Phys::ndpc;
  constexpr int h = (S - 1) / 2;
  int32_t cells[S];
  stencilCells<Phys::dim, S, AX>(row, cells);
  double uLn[N], uLp[N], uRn[N], uRp[N], dterm[N];
#pragma unroll
  for (int d = 0; d < N; ++d) {
    double q[S];
#pragma unroll
    for (int p = 0; p < S; ++p) {
      const int layer = (p < h) ? (h - 1 - p) : (p - h - 1);
      const int side = (p < h) ? sideMinus<AX>() : sidePlus<AX>();
      q[p] = stencilVal<N, NEARBD>(U, cells[p], gv, nbRow, side, layer, d);
    }
    // inner rows: division-free leaf arithmetic (fastmath.cuh); near-boundary rows (< 1 % of the cells, ghost states
    // in the stencil) keep the reference's operation order
    if constexpr (NEARBD) { Recon<S>::face(q, uLn[d], uLp[d]); Recon<S>::face(q + 1, uRn[d], uRp[d]); }
    else { reconFaceFast<S>(q, uLn[d], uLp[d]); reconFaceFast<S>(q + 1, uRn[d], uRp[d]); }
    if constexpr (PhysTraits<Phys>::hasDiffusion) dterm[d] = phys.dD[AX] * (q[h + 1] - 2.0 * q[h] + q[h - 1]);
  }
  double FL[N], FR[N];
  if constexpr (NEARBD) {
    phys.template flux<AX>(uLn, uLp, FL);
    phys.template flux<AX>(uRn, uRp, FR);
  } else {
    faceFlux2d<Phys, AX>(phys, uLn, uLp, FL);
    faceFlux2d<Phys, AX>(phys, uRn, uRp, FR);
  }
#pragma unroll
  for (int d = 0; d < N; ++d) {
    v[d] += hInv * (FL[d] - FR[d]);
    if constexpr (PhysTraits<Phys>::hasDiffusion) {
      if (NEARBD) v[d] += dterm[d]; else diff[d] = dterm[d];
    }
  }
  (void)dterm; (void)diff;
}

// point terms added after the flux balances: SWE Coriolis (swe_2d_prob_class.hpp:984-1012), ADR source + reaction
// (advection_diffusion_reaction_2d_prob_class.hpp:504-509).  `sampleRow` indexes the per-cell source table.
template <class Phys> PDA_DEVFN void addForcing(const Phys&, const double*, double*, int32_t) {}
template <> PDA_DEVFN void addForcing<Swe2d>(const Swe2d& phys, const double* u, double* v, int32_t) {
  v[1] -= phys.coriolis * u[2] / u[0];
  v[2] += phys.coriolis * u[1] / u[0];
}
template <> PDA_DEVFN void addForcing<LinAdv<2>>(const LinAdv<2>& phys, const double* u, double* v, int32_t sampleRow) {
  v[0] += phys.srcTable ? phys.srcTable[sampleRow] : phys.srcConst;
  v[0] -= phys.sigma * u[0];
}

// diffusion term of an inner row gathered through the graph (rows whose first-layer neighbours all exist)
template <class Phys>
PDA_DEVFN void addDiffusionInner(const Phys& phys, const int32_t* __restrict__ row, const double* __restrict__ U, double* v) {
  if constexpr (PhysTraits<Phys>::hasDiffusion) {
    constexpr int N = Phys::ndpc, DIM = Phys::dim;
    const double* c = U + (int64_t)row[0] * N;
#pragma unroll
    for (int ax = 0; ax < DIM; ++ax) {
      const int sm = (ax == 0) ? 0 : (ax == 1 ? 3 : 4), sp = (ax == 0) ? 2 : (ax == 1 ? 1 : 5);
      const double* l = U + (int64_t)row[gcol<DIM>(sm, 0)] * N;
      const double* r = U + (int64_t)row[gcol<DIM>(sp, 0)] * N;
#pragma unroll
      for (int d = 0; d < N; ++d) v[d] += phys.dD[ax] * (r[d] - 2.0 * c[d] + l[d]);
    }
  }
}

// d(point terms + diffusion)/dU of an inner row: `put(k, slot, j, value)` ADDS value to entry (row k, block slot, col j)
template <class Phys, class Put>
PDA_DEVFN void addExtraJacInner(const Phys& phys, const double* u, const uint8_t* __restrict__ slots, Put&& put) {
  constexpr int N = Phys::ndpc, DIM = Phys::dim;
  (void)N; (void)DIM; (void)u; (void)slots;
  if constexpr (std::is_same<Phys, Swe2d>::value) {
    const double f = phys.coriolis;
    put(1, slots[0], 0, f * u[2] / (u[0] * u[0]));
    put(1, slots[0], 2, -f / u[0]);
    put(2, slots[0], 1, f / u[0]);
    put(2, slots[0], 0, -f * u[1] / (u[0] * u[0]));
  }
  if constexpr (PhysTraits<Phys>::hasDiffusion) {
    double self = 0.0;
#pragma unroll
    for (int ax = 0; ax < DIM; ++ax) self += -2.0 * phys.dD[ax];
#pragma unroll
    for (int k = 0; k < N; ++k) put(k, slots[0], k, self);
#pragma unroll
    for (int ax = 0; ax < DIM; ++ax) {
      const int sm = (ax == 0) ? 0 : (ax == 1 ? 3 : 4), sp = (ax == 0) ? 2 : (ax == 1 ? 1 : 5);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        put(k, slots[gcol<DIM>(sm, 0)], k, phys.dD[ax]);
        put(k, slots[gcol<DIM>(sp, 0)], k, phys.dD[ax]);
      }
    }
  }
  if constexpr (std::is_same<Phys, LinAdv<2>>::value) put(0, slots[0], 0, -phys.sigma);
}


// ------------------------------------------------------------------------------------------------ velocity
template <class Phys, int S, bool NEARBD>
__global__ void __launch_bounds__(128)
k_velocity_rows(Phys phys, RowSet rs, Deltas dl, const double* __restrict__ U, double* __restrict__ V, GhostView gv) {
  constexpr int N = Phys::ndpc;
  constexpr int DIM = Phys::dim;
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rs.n) return;
  const int32_t* row = rs.graph + (int64_t)r * rs.ncols;
  double v[N], diff[DIM][N];
#pragma unroll
  for (int d = 0; d < N; ++d) v[d] = 0.0;
  axisVelocity<Phys, S, 0, NEARBD>(phys, row, U, gv, r, dl.hInv[0], v, diff[0]);
  if constexpr (DIM >= 2) axisVelocity<Phys, S, 1, NEARBD>(phys, row, U, gv, r, dl.hInv[1], v, diff[1]);
  if constexpr (DIM >= 3) axisVelocity<Phys, S, 2, NEARBD>(phys, row, U, gv, r, dl.hInv[2], v, diff[2]);
  if constexpr (PhysTraits<Phys>::hasDiffusion && !NEARBD) {
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int d = 0; d < N; ++d) v[d] += diff[a][d];
  }
  (void)diff;
  const int32_t sampleRow = rs.rowIds[r];
  addForcing<Phys>(phys, U + (int64_t)row[0] * N, v, sampleRow);
  double* out = V + (int64_t)sampleRow * N;
#pragma unroll
  for (int d = 0; d < N; ++d) out[d] = v[d];
}

// ------------------------------------------------------------------------------------------------ Jacobian
// Where the blocks of a cell live in the CSR value array: all ndpc rows of a cell share one column pattern, so
// entry (k, block slot s, j) sits at  base + k*len + s*ndpc + j.
// (struct JacLayout: kernel_types.cuh)

template <int N>
PDA_DEVFN void addBlockColumn(double* __restrict__ Jv, int64_t base, int32_t len, int slot, int j, const double* col) {
#pragma unroll
  for (int k = 0; k < N; ++k) Jv[base + (int64_t)k * len + slot * N + j] += col[k];
}

// inner rows, scheme S: velocity and Jacobian (values accumulated into zero-initialised Jv)
template <class Phys, int S, int AX>
PDA_DEVFN void axisJacobianInner(const Phys& phys, const int32_t* __restrict__ row, const double* __restrict__ U,
                                 double hInv, double* v, double* __restrict__ Jv, int64_t base, int32_t len,
                                 const uint8_t* __restrict__ slots) {
  constexpr int N = Phys::ndpc;
  constexpr int h = (S - 1) / 2;
  constexpr int DIM = Phys::dim;
  int32_t cells[S];
  stencilCells<DIM, S, AX>(row, cells);
  // graph column of each stencil position (for the slot lookup)
  int cols[S];
  cols[h] = 0;
#pragma unroll
  for (int L = 0; L < h; ++L) {
    cols[h - 1 - L] = gcol<DIM>(sideMinus<AX>(), L);
    cols[h + 1 + L] = gcol<DIM>(sidePlus<AX>(), L);
  }
#pragma unroll
  for (int face = 0; face < 2; ++face) {   // 0: left face (stencil pos 0..S-2), 1: right face (pos 1..S-1)
    const double sgn = (face == 0) ? hInv : -hInv;
    double un[N], up[N];
#pragma unroll
    for (int d = 0; d < N; ++d) {
      double q[S - 1];
#pragma unroll
      for (int p = 0; p < S - 1; ++p) q[p] = U[(int64_t)cells[p + face] * N + d];
      Recon<S>::face(q, un[d], up[d]);
    }
    double F[N], JN[N * N], JP[N * N];
    phys.template flux<AX>(un, up, F);
    phys.template fluxJac<AX>(un, up, JN, JP);
#pragma unroll
    for (int d = 0; d < N; ++d) v[d] += sgn * F[d];
    // chain rule: d(flux)/d(u_m) = JN * diag(d uNeg/d u_m) + JP * diag(d uPos/d u_m)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double q[S - 1], gN[S - 1], gP[S - 1], t0, t1;
#pragma unroll
      for (int p = 0; p < S - 1; ++p) q[p] = U[(int64_t)cells[p + face] * N + j];
      Recon<S>::faceGrad(q, t0, t1, gN, gP);
#pragma unroll
      for (int m = 0; m < S - 1; ++m) {
        double col[N];
#pragma unroll
        for (int k = 0; k < N; ++k) col[k] = sgn * (JN[k * N + j] * gN[m] + JP[k * N + j] * gP[m]);
        addBlockColumn<N>(Jv, base, len, slots[cols[m + face]], j, col);
      }
    }
  }
}

template <class Phys, int S>
__global__ void __launch_bounds__(128)
k_jacobian_inner_rows(Phys phys, RowSet rs, Deltas dl, const double* __restrict__ U, double* __restrict__ V,
                      double* __restrict__ Jv, JacLayout jl) {
  constexpr int N = Phys::ndpc;
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rs.n) return;
  const int32_t* row = rs.graph + (int64_t)r * rs.ncols;
  const int64_t base = jl.base[r];
  const int32_t len = jl.len[r];
  const uint8_t* slots = jl.slot + (int64_t)r * jl.nslotCols;
  double v[N];
#pragma unroll
  for (int d = 0; d < N; ++d) v[d] = 0.0;
  axisJacobianInner<Phys, S, 0>(phys, row, U, dl.hInv[0], v, Jv, base, len, slots);
  if constexpr (Phys::dim >= 2) axisJacobianInner<Phys, S, 1>(phys, row, U, dl.hInv[1], v, Jv, base, len, slots);
  if constexpr (Phys::dim >= 3) axisJacobianInner<Phys, S, 2>(phys, row, U, dl.hInv[2], v, Jv, base, len, slots);
  const double* uSelf = U + (int64_t)row[0] * N;
  addDiffusionInner<Phys>(phys, row, U, v);
  addForcing<Phys>(phys, uSelf, v, rs.rowIds[r]);
  addExtraJacInner<Phys>(phys, uSelf, slots, [&](int k, int slot, int j, double val) {
    Jv[base + (int64_t)k * len + slot * N + j] += val;
  });
  if (V) {
    double* out = V + (int64_t)rs.rowIds[r] * N;
#pragma unroll
    for (int d = 0; d < N; ++d) out[d] = v[d];
  }
}

// near-boundary rows: FIRST-ORDER Jacobian whatever the velocity scheme is ("DifferentScheme" path), missing
// first-layer neighbours folded into the self block with per-dof factors
// (mixin_directional_flux_balance_jacobian.hpp:287-371).  The velocity of these rows is k_velocity_rows<.,S,true>.
template <class Phys, int AX>
PDA_DEVFN void axisJacobianNearBd(const Phys& phys, const int32_t* __restrict__ row, const double* __restrict__ U,
                                  const GhostView& gv, int32_t nbRow, double hInv, const double* __restrict__ fac,
                                  double* __restrict__ Jv, int64_t base, int32_t len,
                                  const uint8_t* __restrict__ slots) {
  constexpr int N = Phys::ndpc;
  constexpr int DIM = Phys::dim;
  const int cl = gcol<DIM>(sideMinus<AX>(), 0), cr = gcol<DIM>(sidePlus<AX>(), 0);
  const int32_t l0 = row[cl], r0 = row[cr];
  double qL[N], qC[N], qR[N];
#pragma unroll
  for (int d = 0; d < N; ++d) {
    qL[d] = stencilVal<N, true>(U, l0, gv, nbRow, sideMinus<AX>(), 0, d);
    qC[d] = U[(int64_t)row[0] * N + d];
    qR[d] = stencilVal<N, true>(U, r0, gv, nbRow, sidePlus<AX>(), 0, d);
  }
  double JN[N * N], JP[N * N];
  const int sSelf = slots[0];
  // left face: flux(qL, qC)
  phys.template fluxJac<AX>(qL, qC, JN, JP);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double col[N];
#pragma unroll
    for (int k = 0; k < N; ++k) col[k] = hInv * JP[k * N + j];
    addBlockColumn<N>(Jv, base, len, sSelf, j, col);
#pragma unroll
    for (int k = 0; k < N; ++k) col[k] = hInv * JN[k * N + j] * (l0 >= 0 ? 1.0 : fac[j]);
    addBlockColumn<N>(Jv, base, len, (l0 >= 0) ? slots[cl] : sSelf, j, col);
  }
  // right face: flux(qC, qR)
  phys.template fluxJac<AX>(qC, qR, JN, JP);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double col[N];
#pragma unroll
    for (int k = 0; k < N; ++k) col[k] = -hInv * JN[k * N + j];
    addBlockColumn<N>(Jv, base, len, sSelf, j, col);
#pragma unroll
    for (int k = 0; k < N; ++k) col[k] = -hInv * JP[k * N + j] * (r0 >= 0 ? 1.0 : fac[j]);
    addBlockColumn<N>(Jv, base, len, (r0 >= 0) ? slots[cr] : sSelf, j, col);
  }
}

template <class Phys>
__global__ void __launch_bounds__(128)
k_jacobian_nearbd_rows(Phys phys, RowSet rs, Deltas dl, const double* __restrict__ U, double* __restrict__ Jv,
                       JacLayout jl, GhostView gv, const double* __restrict__ factors /*[n][dim][ndpc]*/) {
  constexpr int N = Phys::ndpc;
  constexpr int DIM = Phys::dim;
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rs.n) return;
  const int32_t* row = rs.graph + (int64_t)r * rs.ncols;
  const int64_t base = jl.base[r];
  const int32_t len = jl.len[r];
  const uint8_t* slots = jl.slot + (int64_t)r * jl.nslotCols;
  const double* fac = factors + (int64_t)r * DIM * N;
  axisJacobianNearBd<Phys, 0>(phys, row, U, gv, r, dl.hInv[0], fac, Jv, base, len, slots);
  if constexpr (DIM >= 2) axisJacobianNearBd<Phys, 1>(phys, row, U, gv, r, dl.hInv[1], fac + N, Jv, base, len, slots);
  if constexpr (DIM >= 3) axisJacobianNearBd<Phys, 2>(phys, row, U, gv, r, dl.hInv[2], fac + 2 * N, Jv, base, len, slots);
  // point terms: Coriolis / reaction on the self block
  const double* uSelf = U + (int64_t)row[0] * N;
  if constexpr (!PhysTraits<Phys>::hasDiffusion) {
    addExtraJacInner<Phys>(phys, uSelf, slots, [&](int k, int slot, int j, double val) {
      Jv[base + (int64_t)k * len + slot * N + j] += val;
    });
  } else {
    // diffusion on a near-boundary row (advection_diffusion_2d_prob_class.hpp:755-792,
    // advection_diffusion_reaction_2d_prob_class.hpp:693-721): existing first-layer neighbours get dD, a missing one
    // folds -dD into the self entry
    double self = 0.0;
#pragma unroll
    for (int ax = 0; ax < DIM; ++ax) self += -2.0 * phys.dD[ax];
    if constexpr (std::is_same<Phys, LinAdv<2>>::value) self -= phys.sigma;
#pragma unroll
    for (int ax = 0; ax < DIM; ++ax) {
      const int sm = (ax == 0) ? 0 : (ax == 1 ? 3 : 4), sp = (ax == 0) ? 2 : (ax == 1 ? 1 : 5);
      for (int side : {sm, sp}) {
        const int c = gcol<DIM>(side, 0);
        if (row[c] >= 0) {
#pragma unroll
          for (int k = 0; k < N; ++k) Jv[base + (int64_t)k * len + slots[c] * N + k] += phys.dD[ax];
        } else {
          self += -phys.dD[ax];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) Jv[base + (int64_t)k * len + slots[0] * N + k] += self;
  }
}

// ------------------------------------------------------------------------------------------------ Gray-Scott
// diffusion_reaction_2d_prob_class.hpp:459-529: V is ASSIGNED; J: 2x2 self block + diagonal neighbour blocks
// (off-diagonal entries of the neighbour blocks are stored zeros).
struct GrayScottParams { double Du, Dv, F, k, dxInvSq, dyInvSq; };

__global__ void k_gray_scott_rows(GrayScottParams gp, RowSet rs, const double* __restrict__ U,
                                  double* __restrict__ V, double* __restrict__ Jv, JacLayout jl) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rs.n) return;
  const int32_t* row = rs.graph + (int64_t)r * rs.ncols;
  const int64_t c = row[0], l = row[1], f = row[2], rt = row[3], b = row[4];
  const double u = U[2 * c], w = U[2 * c + 1];
  const double uDx = gp.Du * gp.dxInvSq, uDy = gp.Du * gp.dyInvSq;
  const double vDx = gp.Dv * gp.dxInvSq, vDy = gp.Dv * gp.dyInvSq;
  const double uvv = u * w * w;
  if (V) {
    const int64_t o = (int64_t)rs.rowIds[r] * 2;
    V[o] = gp.F * (1.0 - u) - uvv + uDx * (U[2 * rt] - 2.0 * u + U[2 * l]) + uDy * (U[2 * b] - 2.0 * u + U[2 * f]);
    V[o + 1] = -(gp.F + gp.k) * w + uvv + vDx * (U[2 * rt + 1] - 2.0 * w + U[2 * l + 1]) +
               vDy * (U[2 * b + 1] - 2.0 * w + U[2 * f + 1]);
  }
  if (Jv) {
    const int64_t base = jl.base[r];
    const int32_t len = jl.len[r];
    const uint8_t* slots = jl.slot + (int64_t)r * jl.nslotCols;
    double* r0 = Jv + base;
    double* r1 = Jv + base + len;
    const int s = slots[0];
    r0[2 * s] += -2.0 * uDx - 2.0 * uDy - w * w - gp.F;
    r0[2 * s + 1] -= 2.0 * u * w;
    r1[2 * s] += w * w;
    r1[2 * s + 1] += -2.0 * vDx - 2.0 * vDy + 2.0 * u * w - (gp.F + gp.k);
    const int sl = slots[1], sf = slots[2], sr = slots[3], sb = slots[4];
    r0[2 * sl] += uDx; r0[2 * sf] += uDy; r0[2 * sr] += uDx; r0[2 * sb] += uDy;
    r1[2 * sl + 1] += vDx; r1[2 * sf + 1] += vDy; r1[2 * sr + 1] += vDx; r1[2 * sb + 1] += vDy;
  }
}

// ------------------------------------------------------------------------------------------------ diffusion-reaction A
// DiffusionReaction1d::ProblemA / DiffusionReaction2d::ProblemA (one dof):  ds/dt = D lap(s) + k s^2 + f(x[,y],t)
// (diffusion_reaction_1d_prob_class.hpp:211-302, diffusion_reaction_2d_prob_class.hpp:306-456).  One thread per
// sample row, all rows in one launch.  The ghost of a missing neighbour is -s(self) (homogeneous Dirichlet at the
// wall: diffusion_reaction_1d_ghost_filler.hpp:85-94), so no ghost arrays are needed.  f is a per-row table.
struct DiffReacParams { double dD[2]; double reaction; };

template <int DIM>
__global__ void k_diffreac_rows(DiffReacParams pr, RowSet rs, const double* __restrict__ U, const double* __restrict__ src,
                                double* __restrict__ V, double* __restrict__ Jv, JacLayout jl) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rs.n) return;
  const int32_t* row = rs.graph + (int64_t)r * rs.ncols;
  const double u = U[row[0]];
  // neighbour columns in accumulation order: x: left (1), right (DIM==1 ? 2 : 3); y: back (4), front (2)
  const int cl = 1, cr = (DIM == 1) ? 2 : 3, cb = 4, cf = 2;
  const int32_t il = row[cl], ir = row[cr];
  const double ul = il >= 0 ? U[il] : -u, ur = ir >= 0 ? U[ir] : -u;
  bool nearBd = (il < 0) || (ir < 0);
  int32_t ib = 0, ifr = 0;
  double ub = 0.0, uf = 0.0;
  if (DIM == 2) {
    ib = row[cb]; ifr = row[cf];
    ub = ib >= 0 ? U[ib] : -u; uf = ifr >= 0 ? U[ifr] : -u;
    nearBd = nearBd || (ib < 0) || (ifr < 0);
  }
  const int32_t sampleRow = rs.rowIds[r];
  if (V) {
    double v = src[sampleRow];
    v += pr.reaction * u * u;
    v += pr.dD[0] * (ur - 2.0 * u + ul);
    if (DIM == 2) v += pr.dD[1] * (uf - 2.0 * u + ub);
    V[sampleRow] = v;
  }
  if (Jv) {
    const int64_t base = jl.base[r];
    const uint8_t* slots = jl.slot + (int64_t)r * jl.nslotCols;
    double* Jr = Jv + base;
    const double twoK = pr.reaction * 2.0;
    if (DIM == 1) {
      Jr[slots[0]] += (nearBd ? -3.0 * pr.dD[0] : -2.0 * pr.dD[0]) + twoK * u;
      if (il >= 0) Jr[slots[cl]] += pr.dD[0];
      if (ir >= 0) Jr[slots[cr]] += pr.dD[0];
    } else {
      double self = -2.0 * pr.dD[0] - 2.0 * pr.dD[1] + twoK * u;
      // reference order of the missing-neighbour folds: left, front, right, back
      if (il >= 0) Jr[slots[cl]] += pr.dD[0]; else self += -pr.dD[0];
      if (ifr >= 0) Jr[slots[cf]] += pr.dD[1]; else self += -pr.dD[1];
      if (ir >= 0) Jr[slots[cr]] += pr.dD[0]; else self += -pr.dD[0];
      if (ib >= 0) Jr[slots[cb]] += pr.dD[1]; else self += -pr.dD[1];
      Jr[slots[0]] += self;
    }
  }
}

// Gray-Scott velocity on a fully periodic full lattice: neighbours by index arithmetic, no graph in HBM
// (HBM-bound: 32 B/cell; rows of 128 consecutive x cells per CTA keep the y neighbours' loads coalesced)
__global__ void __launch_bounds__(128)
k_gray_scott_lattice(GrayScottParams gp, int32_t nx, int32_t ny, const double2* __restrict__ U, double2* __restrict__ V) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t j = blockIdx.y;
  if (i >= nx) return;
  const int32_t il = (i == 0) ? nx - 1 : i - 1, ir = (i == nx - 1) ? 0 : i + 1;
  const int32_t jb = (j == 0) ? ny - 1 : j - 1, jf = (j == ny - 1) ? 0 : j + 1;
  const double2 c = U[(int64_t)j * nx + i];
  const double2 l = U[(int64_t)j * nx + il], rt = U[(int64_t)j * nx + ir];
  const double2 b = U[(int64_t)jb * nx + i], f = U[(int64_t)jf * nx + i];
  const double uDx = gp.Du * gp.dxInvSq, uDy = gp.Du * gp.dyInvSq;
  const double vDx = gp.Dv * gp.dxInvSq, vDy = gp.Dv * gp.dyInvSq;
  const double uvv = c.x * c.y * c.y;
  double2 out;
  out.x = gp.F * (1.0 - c.x) - uvv + uDx * (rt.x - 2.0 * c.x + l.x) + uDy * (b.x - 2.0 * c.x + f.x);
  out.y = -(gp.F + gp.k) * c.y + uvv + vDx * (rt.y - 2.0 * c.y + l.y) + vDy * (b.y - 2.0 * c.y + f.y);
  V[(int64_t)j * nx + i] = out;
}

// Gray-Scott velocity + Jacobian on a periodic full lattice: one thread per cell, 128 consecutive cells per CTA.  The
// 20 stored entries of a cell (2 rows x 5 blocks x 2, explicit zeros included) are assembled in shared memory at
// their sorted block positions and the CTA streams its 128 chunks -- one contiguous 20 KB range of the CSR value
// array -- out coalesced: every value written once, no memset, no read-modify-write.
__global__ void __launch_bounds__(128)
k_gray_scott_lattice_jac(GrayScottParams gp, int32_t nx, int32_t ny, const double2* __restrict__ U,
                         double2* __restrict__ V, double* __restrict__ Jv, const int32_t* __restrict__ cellBase,
                         const uint4* __restrict__ cellSlots) {
  constexpr int CH = 20, STR = 21;
  __shared__ double sJ[128 * STR];
  const int64_t ncell = (int64_t)nx * ny;
  const int64_t gid0 = (int64_t)blockIdx.x * 128;
  const int64_t gid = gid0 + threadIdx.x;
  if (gid < ncell) {
    const int32_t i = (int32_t)(gid % nx), j = (int32_t)(gid / nx);
    const int32_t il = (i == 0) ? nx - 1 : i - 1, ir = (i == nx - 1) ? 0 : i + 1;
    const int32_t jb = (j == 0) ? ny - 1 : j - 1, jf = (j == ny - 1) ? 0 : j + 1;
    const double2 c = U[gid];
    const double2 l = U[(int64_t)j * nx + il], rt = U[(int64_t)j * nx + ir];
    const double2 b = U[(int64_t)jb * nx + i], f = U[(int64_t)jf * nx + i];
    const double uDx = gp.Du * gp.dxInvSq, uDy = gp.Du * gp.dyInvSq;
    const double vDx = gp.Dv * gp.dxInvSq, vDy = gp.Dv * gp.dyInvSq;
    const double u = c.x, w = c.y;
    const double uvv = u * w * w;
    if (V) {
      double2 out;
      out.x = gp.F * (1.0 - u) - uvv + uDx * (rt.x - 2.0 * u + l.x) + uDy * (b.x - 2.0 * u + f.x);
      out.y = -(gp.F + gp.k) * w + uvv + vDx * (rt.y - 2.0 * w + l.y) + vDy * (b.y - 2.0 * w + f.y);
      V[gid] = out;
    }
    const uint4 sv = __ldg(cellSlots + gid);   // graph columns: 0 self, 1 left, 2 front (j+1), 3 right, 4 back (j-1)
    const int s0 = sv.x & 0xff, sl = (sv.x >> 8) & 0xff, sf = (sv.x >> 16) & 0xff, sr = (sv.x >> 24) & 0xff, sb = sv.y & 0xff;
    double* r0 = sJ + threadIdx.x * STR;
    double* r1 = r0 + 10;
    // the reference accumulates into zeroed entries: 0 + x (diffusion_reaction_2d_prob_class.hpp:486-527)
    r0[2 * s0] = -2.0 * uDx - 2.0 * uDy - w * w - gp.F;
    r0[2 * s0 + 1] = -(2.0 * u * w);
    r1[2 * s0] = w * w;
    r1[2 * s0 + 1] = -2.0 * vDx - 2.0 * vDy + 2.0 * u * w - (gp.F + gp.k);
    r0[2 * sl] = uDx; r0[2 * sl + 1] = 0.0; r1[2 * sl] = 0.0; r1[2 * sl + 1] = vDx;
    r0[2 * sr] = uDx; r0[2 * sr + 1] = 0.0; r1[2 * sr] = 0.0; r1[2 * sr + 1] = vDx;
    r0[2 * sf] = uDy; r0[2 * sf + 1] = 0.0; r1[2 * sf] = 0.0; r1[2 * sf + 1] = vDy;
    r0[2 * sb] = uDy; r0[2 * sb + 1] = 0.0; r1[2 * sb] = 0.0; r1[2 * sb + 1] = vDy;
  }
  __syncthreads();
  const int nvalid = (int)min((int64_t)128, ncell - gid0);
  double* dst = Jv + cellBase[gid0];
  for (int e = threadIdx.x; e < nvalid * CH; e += 128) {
    const int cc = e / CH, w = e - cc * CH;
    dst[e] = sJ[cc * STR + w];
  }
}

// ------------------------------------------------------------------------------------------------ J * B
// applyJacobian (adapter_cpp.hpp:231-259): R = J * B with the fixed CSR pattern.  The N rows of a cell share one
// column pattern made of N-wide blocks (entry (k, block b, j) at base + k*len + b*N + j, column id_b*N + j), so the
// product is done per CELL:
//   * row-major operands with >= 8 columns: one warp per cell, LANES ACROSS THE OPERAND COLUMNS -- every B row is one
//     coalesced read shared by the N rows of the cell, J entries are warp-uniform (vector) loads, no reduction;
//     cells are visited in a tile-major order on lattices so that the B rows of stencil neighbours hit in L1;
//   * everything else (vectors, few columns): one warp per row, lanes across the row's entries, J read once for up
//     to 8 columns, shuffle reduction.  Column-major operands with many columns are transposed around kernel 1.
// smem: per warp two slots of `slotDoubles` doubles (the J chunk of the current cell and of the next one in flight)
template <int N>
__global__ void __launch_bounds__(256)
k_spmm_cells_rowmajor(int32_t ncells, const int32_t* __restrict__ order, const int32_t* __restrict__ cellBase,
                      const int32_t* __restrict__ cellLen, const int32_t* __restrict__ colidx,
                      const double* __restrict__ vals, const double* __restrict__ B, int nB, double* __restrict__ R,
                      int slotDoubles) {
  // a CTA owns 64 consecutive cells of the visiting order (one 8x8 lattice tile): warp w takes cells w, w+8, ...
  // The cell's J chunk (N rows x len, contiguous in the CSR value array) streams into shared memory with cp.async,
  // one cell ahead of the arithmetic; the arithmetic then reads it as warp-uniform (broadcast) shared loads.
  extern __shared__ __align__(16) double sChunks[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const slot0 = sChunks + (size_t)(2 * warp) * slotDoubles;
  constexpr int PIECE = (N % 2 == 0) ? 2 : 1;   // doubles per cp.async (16-byte pieces need an even chunk offset)

  auto cellOf = [&](int t) -> int32_t {
    const int64_t w = (int64_t)blockIdx.x * 64 + t * 8 + warp;
    if (t >= 8 || w >= ncells) return -1;
    return order ? order[w] : (int32_t)w;
  };
  auto prefetch = [&](int32_t cell, double* dst) {
    if (cell >= 0) {
      const double* src = vals + cellBase[cell];
      const int n = N * cellLen[cell];
      for (int e = lane * PIECE; e < n; e += 32 * PIECE) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + e);
        if constexpr (PIECE == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + e) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(src + e) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  int32_t cell = cellOf(0);
  prefetch(cell, slot0);
  for (int t = 0; t < 8 && cell >= 0; ++t) {
    const int32_t next = cellOf(t + 1);
    prefetch(next, slot0 + ((t + 1) & 1) * slotDoubles);
    const int64_t base = cellBase[cell];
    const int32_t len = cellLen[cell];
    const int nblk = len / N;
    // first column of every block: one load by the first nblk lanes, broadcast by shuffle (nblk <= 19)
    const int32_t myCol = (lane < nblk) ? __ldg(colidx + base + lane * N) : 0;
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");   // this cell's chunk has landed (the next one may be in flight)
    __syncwarp();
    const double* sj = slot0 + (t & 1) * slotDoubles;
    for (int c0 = 0; c0 < nB; c0 += 32) {
      const int c = c0 + lane;
      const bool active = c < nB;
      const double* Bc = B + (active ? c : 0);
      double acc[N];
#pragma unroll
      for (int k = 0; k < N; ++k) acc[k] = 0.0;
#pragma unroll 4
      for (int b = 0; b < nblk; ++b) {
        const int32_t col0 = __shfl_sync(0xffffffffu, myCol, b);
        const double* bp = Bc + (int64_t)col0 * nB;
        double bv[N];
#pragma unroll
        for (int j = 0; j < N; ++j) bv[j] = __ldg(bp + (int64_t)j * nB);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const double* jv = sj + k * len + b * N;
          if constexpr (N % 2 == 0) {   // even N: the block row is 16-byte aligned in the slot -> LDS.128
#pragma unroll
            for (int j = 0; j < N; j += 2) {
              const double2 a = *reinterpret_cast<const double2*>(jv + j);
              acc[k] += a.x * bv[j];
              acc[k] += a.y * bv[j + 1];
            }
          } else {
#pragma unroll
            for (int j = 0; j < N; ++j) acc[k] += jv[j] * bv[j];
          }
        }
      }
      if (active) {
#pragma unroll
        for (int k = 0; k < N; ++k) R[((int64_t)cell * N + k) * nB + c] = acc[k];
      }
    }
    __syncwarp();   // everyone is done with this slot before the prefetch two cells ahead overwrites it
    cell = next;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

__global__ void __launch_bounds__(256)
k_spmm_rows_fewcols(int32_t nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                    const double* __restrict__ vals, const double* __restrict__ B, int nB, int64_t ldbRow,
                    int64_t ldbCol, double* __restrict__ R, int64_t ldrRow, int64_t ldrCol) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= nrows) return;
  const int32_t b = rowptr[warp], e = rowptr[warp + 1];
  for (int c0 = 0; c0 < nB; c0 += 8) {
    const int nb = min(8, nB - c0);
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.0;
    for (int32_t p = b + lane; p < e; p += 32) {
      const double v = vals[p];
      const double* brow = B + (int64_t)colidx[p] * ldbRow + (int64_t)c0 * ldbCol;
#pragma unroll
      for (int c = 0; c < 8; ++c) if (c < nb) acc[c] += v * brow[c * ldbCol];
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < nb) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        if (lane == 0) R[(int64_t)warp * ldrRow + (int64_t)(c0 + c) * ldrCol] = acc[c];
      }
    }
  }
}

// Few operand columns (J*v of Newton-Krylov solvers): one warp per CELL.  The cell's N rows are one contiguous chunk
// of N*len values: lanes stride over the chunk (coalesced, every J value and column index read once) in batches of
// four independent entries (all loads of a batch in flight together); entry e belongs to row e / len (compares, no
// division); per-row partial sums in registers, one shuffle reduction per row and column.
template <int N, int NB>
__global__ void __launch_bounds__(256)
k_spmm_cells_fewcols(int32_t ncells, const int32_t* __restrict__ order, const int32_t* __restrict__ cellBase,
                     const int32_t* __restrict__ cellLen, const int32_t* __restrict__ colidx,
                     const double* __restrict__ vals, const double* __restrict__ B,
                     int c0, int64_t ldbRow, int64_t ldbCol, double* __restrict__ R, int64_t ldrRow, int64_t ldrCol) {
  const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= ncells) return;
  const int cell = order ? order[w] : w;   // `order`: optional list of cells (e.g. the near-boundary rows only)
  const int64_t base = cellBase[cell];
  const int32_t len = cellLen[cell];
  const int n = N * len;
  const double* Bc = B + (int64_t)c0 * ldbCol;
  double acc[N][NB];
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int c = 0; c < NB; ++c) acc[k][c] = 0.0;
  for (int e0 = lane; e0 < n; e0 += 128) {
    double v[4];
    int32_t col[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + 32 * u;
      const bool ok = e < n;
      v[u] = ok ? __ldg(vals + base + e) : 0.0;
      col[u] = ok ? __ldg(colidx + base + e) : 0;
    }
    double bv[4][NB];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < NB; ++c) bv[u][c] = __ldg(Bc + (int64_t)col[u] * ldbRow + c * ldbCol);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + 32 * u;
      int k = 0;
#pragma unroll
      for (int kk = 1; kk < N; ++kk) k += (e >= kk * len) ? 1 : 0;
#pragma unroll
      for (int kk = 0; kk < N; ++kk) {
        const double w = (kk == k) ? v[u] : 0.0;
#pragma unroll
        for (int c = 0; c < NB; ++c) acc[kk][c] = fma(w, bv[u][c], acc[kk][c]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      double a = acc[k][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) R[((int64_t)cell * N + k) * ldrRow + (int64_t)(c0 + c) * ldrCol] = a;
    }
}

// out[c][r] <- in[r][c]  (in: rows x cols row-major); 32x32 tiles through shared memory, linear tile index
__global__ void __launch_bounds__(256)
k_transpose(const double* __restrict__ in, int64_t rows, int64_t cols, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int64_t tilesC = (cols + 31) / 32;
  const int64_t r0 = ((int64_t)blockIdx.x / tilesC) * 32, c0 = ((int64_t)blockIdx.x % tilesC) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = in[r * cols + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) out[c * rows + r] = tile[tx][i];
  }
}

// the same for SKINNY matrices (a few columns / a few rows: row-major multi-column operands around the single-column
// J*v kernel): one thread per long-axis index, the short axis in a loop -- both sides stay coalesced, no 32x32 tiles that
// would be 1/16 full
__global__ void __launch_bounds__(256)
k_split_columns(const double* __restrict__ in, int64_t rows, int nc, double* __restrict__ out) {   // [rows][nc] -> [nc][rows]
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  for (int c = 0; c < nc; ++c) out[(int64_t)c * rows + r] = in[r * nc + c];
}
__global__ void __launch_bounds__(256)
k_merge_columns(const double* __restrict__ in, int64_t rows, int nc, double* __restrict__ out) {   // [nc][rows] -> [rows][nc]
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  for (int c = 0; c < nc; ++c) out[r * nc + c] = in[(int64_t)c * rows + r];
}

}  // namespace dev
}  // namespace pda
