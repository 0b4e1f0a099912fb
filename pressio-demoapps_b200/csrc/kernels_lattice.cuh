// kernels_lattice.cuh -- structured-lattice fast path: full meshes in natural ordering need no connectivity graph
// in HBM; neighbours are index arithmetic (periodic wrap or "cell is near a physical boundary -> left to the
// graph-driven near-boundary kernel").
//
// Replaces the reference's inner-cell velocity loop (euler_3d_prob_class.hpp:861-927, euler_2d_prob_class.hpp:991-1044,
// swe_2d_prob_class.hpp:928-981) for rows = graphRowsOfCellsAwayFromBd() of a full mesh.
#pragma once
#include <cstdint>

#include "kernels_generic.cuh"
#include "mesh.hpp"

namespace pda {

namespace dev {

struct LatticeDesc {
  int32_t n[3];        // cells per axis (unused axes = 1)
  int32_t per[3];      // periodic flags
  int32_t planeBegin;  // slowest-axis planes [planeBegin, planeEnd) are evaluated
  int32_t planeEnd;
  int32_t haloPlanes;  // slab mode: U carries this many extra planes below plane 0 (and above the last)
  int32_t slab;        // 1: no wrap along the slowest axis, U/V are slab-local; 2: peer mode (below)
  int32_t meshHalo;    // (mesh stencil-1)/2: rows this close to a physical boundary are near-boundary rows
  // peer mode (slab == 2, multi-GPU over NVLink peer memory): U holds the owned planes only; the h planes below plane 0
  // / above plane n-1 live in library-owned halo buffers that the ring neighbours fill by peer copies, each followed
  // by a flag write.  A CTA that needs a halo plane spins on the flag until it carries this evaluation's epoch.
  const double* haloLo;
  const double* haloHi;
  const uint32_t* flagLo;
  const uint32_t* flagHi;
  uint32_t epoch;
};

// index of the cell `off` steps away along one axis; -1 if outside a non-periodic axis
PDA_DEVFN int32_t shiftIdx(int32_t idx, int off, int32_t n, int32_t periodic) {
  int32_t j = idx + off;
  if (j < 0) j = periodic ? j + n : -1;
  else if (j >= n) j = periodic ? j - n : -1;
  return j;
}

// v1: one thread per cell, both faces of every axis computed by the owning cell (like the reference).
template <class Phys, int S, int AX>
PDA_DEVFN void latticeAxis(const Phys& phys, const double* __restrict__ U, const int64_t* cellOfPos, double hInv, double* v,
                           double* diff) {
  constexpr int N = Phys::ndpc;
  constexpr int h = (S - 1) / 2;
  double uLn[N], uLp[N], uRn[N], uRp[N];
#pragma unroll
  for (int d = 0; d < N; ++d) {
    double q[S];
#pragma unroll
    for (int p = 0; p < S; ++p) q[p] = U[cellOfPos[p] * N + d];
    Recon<S>::face(q, uLn[d], uLp[d]);
    Recon<S>::face(q + 1, uRn[d], uRp[d]);
    if constexpr (PhysTraits<Phys>::hasDiffusion) diff[d] = phys.dD[AX] * (q[h + 1] - 2.0 * q[h] + q[h - 1]);
  }
  (void)diff;
  double FL[N], FR[N];
  phys.template flux<AX>(uLn, uLp, FL);
  phys.template flux<AX>(uRn, uRp, FR);
#pragma unroll
  for (int d = 0; d < N; ++d) v[d] += hInv * (FL[d] - FR[d]);
}

template <class Phys, int S>
__global__ void __launch_bounds__(128)
k_velocity_lattice_v1(Phys phys, LatticeDesc L, Deltas dl, const double* __restrict__ U, double* __restrict__ V) {
  constexpr int N = Phys::ndpc;
  constexpr int DIM = Phys::dim;
  constexpr int h = (S - 1) / 2;
  const int64_t planeCells = (DIM == 3) ? (int64_t)L.n[0] * L.n[1] : (DIM == 2 ? L.n[0] : 1);
  const int64_t nwork = planeCells * (L.planeEnd - L.planeBegin);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nwork) return;
  int32_t ijk[3] = {0, 0, 0};
  {
    const int64_t inPlane = tid % planeCells;
    const int32_t plane = L.planeBegin + (int32_t)(tid / planeCells);
    if (DIM == 1) ijk[0] = plane;
    else if (DIM == 2) { ijk[0] = (int32_t)inPlane; ijk[1] = plane; }
    else { ijk[0] = (int32_t)(inPlane % L.n[0]); ijk[1] = (int32_t)(inPlane / L.n[0]); ijk[2] = plane; }
  }
  // cells whose stencil leaves a non-periodic domain belong to the near-boundary kernel
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    const bool wrap = L.per[a] || (L.slab && a == DIM - 1);
    if (!wrap && (ijk[a] < L.meshHalo || ijk[a] >= L.n[a] - L.meshHalo)) return;
  }
  const int64_t strideU[3] = {1, L.n[0], (int64_t)L.n[0] * L.n[1]};
  auto cellIndex = [&](int32_t i, int32_t j, int32_t k) -> int64_t {
    if (DIM == 1) return L.slab ? i + L.haloPlanes : i;
    if (DIM == 2) return (int64_t)(L.slab ? j + L.haloPlanes : j) * L.n[0] + i;
    return ((int64_t)(L.slab ? k + L.haloPlanes : k) * L.n[1] + j) * L.n[0] + i;
  };
  (void)strideU;
  double v[N], diff[DIM][N];
#pragma unroll
  for (int d = 0; d < N; ++d) v[d] = 0.0;
  int64_t pos[S];
  // x
#pragma unroll
  for (int p = 0; p < S; ++p) {
    const bool slabAxis = L.slab && DIM == 1;
    const int32_t ii = slabAxis ? ijk[0] + (p - h) : shiftIdx(ijk[0], p - h, L.n[0], 1);
    pos[p] = cellIndex(ii, ijk[1], ijk[2]);
  }
  latticeAxis<Phys, S, 0>(phys, U, pos, dl.hInv[0], v, diff[0]);
  if constexpr (DIM >= 2) {
#pragma unroll
    for (int p = 0; p < S; ++p) {
      const bool slabAxis = L.slab && DIM == 2;
      const int32_t jj = slabAxis ? ijk[1] + (p - h) : shiftIdx(ijk[1], p - h, L.n[1], 1);
      pos[p] = cellIndex(ijk[0], jj, ijk[2]);
    }
    latticeAxis<Phys, S, 1>(phys, U, pos, dl.hInv[1], v, diff[1]);
  }
  if constexpr (DIM >= 3) {
#pragma unroll
    for (int p = 0; p < S; ++p) {
      const int32_t kk = L.slab ? ijk[2] + (p - h) : shiftIdx(ijk[2], p - h, L.n[2], 1);
      pos[p] = cellIndex(ijk[0], ijk[1], kk);
    }
    latticeAxis<Phys, S, 2>(phys, U, pos, dl.hInv[2], v, diff[2]);
  }
  if constexpr (PhysTraits<Phys>::hasDiffusion) {
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int d = 0; d < N; ++d) v[d] += diff[a][d];
  }
  (void)diff;
  const int64_t self = cellIndex(ijk[0], ijk[1], ijk[2]);
  // V is indexed without halo planes
  const int64_t vIdx = (DIM == 1) ? ijk[0] : (DIM == 2 ? (int64_t)ijk[1] * L.n[0] + ijk[0]
                                                      : ((int64_t)ijk[2] * L.n[1] + ijk[1]) * L.n[0] + ijk[0]);
  addForcing<Phys>(phys, U + self * N, v, (int32_t)vIdx);
  double* out = V + vIdx * N;
#pragma unroll
  for (int d = 0; d < N; ++d) out[d] = v[d];
}

}  // namespace dev

inline bool latticeKernelAvailable(int family, int dim, int /*S*/) {
  // Euler 1/2/3D, SWE, Burgers, ADR and 1D advection go through the structured kernels; the diffusion-reaction
  // family has its own lattice kernel (k_diffreac_lattice)
  return family == 1 || family == 2 || family == 3 || family == 4 || family == 6 || family == 7 || family == 8;
  (void)dim;
}

// 3D: tiled, face-sharing kernel (kernels_tiled.cuh)
template <class Phys, int S>
void launchLattice3dTiled(const Phys& phys, const dev::LatticeDesc& L, const dev::Deltas& dl, const double* dU,
                          double* dV, cudaStream_t st);

// 2D: y-marching, face-sharing kernel (kernels_march2d.cuh)
template <class Phys, int S>
void launchMarch2d(const Phys& phys, const dev::LatticeDesc& L, const dev::Deltas& dl, const double* dU, double* dV,
                   cudaStream_t st);

template <class Phys, int S>
void launchLatticeVelocity(const Phys& phys, const Mesh& m, const dev::Deltas& dl, const double* dU, double* dV,
                           cudaStream_t st, int32_t planeBegin, int32_t planeEnd, int /*flags*/) {
  dev::LatticeDesc L;
  for (int a = 0; a < 3; ++a) { L.n[a] = m.n[a]; L.per[a] = m.periodic[a] ? 1 : 0; }
  L.planeBegin = planeBegin; L.planeEnd = planeEnd; L.haloPlanes = 0; L.slab = 0; L.meshHalo = m.halo();
  L.haloLo = L.haloHi = nullptr; L.flagLo = L.flagHi = nullptr; L.epoch = 0;
  int64_t planeCells = 1;
  for (int a = 0; a < Phys::dim - 1; ++a) planeCells *= m.n[a];
  const int64_t nwork = planeCells * (planeEnd - planeBegin);
  if (nwork <= 0) return;
  if constexpr (Phys::dim == 3) {
    if (m.n[0] >= 2 * m.halo() && m.n[1] >= 2 * m.halo() && m.n[2] >= 2 * m.halo()) {
      launchLattice3dTiled<Phys, S>(phys, L, dl, dU, dV, st);
      return;
    }
  }
  if constexpr (Phys::dim == 2) {
    launchMarch2d<Phys, S>(phys, L, dl, dU, dV, st);
    return;
  }
  const int block = 128;
  dev::k_velocity_lattice_v1<Phys, S><<<(unsigned)((nwork + block - 1) / block), block, 0, st>>>(phys, L, dl, dU, dV);
}

// slab-local evaluation of owned planes [pBegin, pEnd) (local numbering, 0 = first owned plane)
template <class Phys, int S>
void launchLatticeVelocitySlab(const Phys& phys, const Mesh& m, const dev::Deltas& dl, const double* dUlocal,
                               double* dVowned, cudaStream_t st, int32_t nOwned, int32_t pBegin, int32_t pEnd) {
  dev::LatticeDesc L;
  for (int a = 0; a < 3; ++a) { L.n[a] = m.n[a]; L.per[a] = m.periodic[a] ? 1 : 0; }
  L.n[Phys::dim - 1] = nOwned;
  L.planeBegin = pBegin; L.planeEnd = pEnd; L.haloPlanes = (S - 1) / 2; L.slab = 1; L.meshHalo = m.halo();
  L.haloLo = L.haloHi = nullptr; L.flagLo = L.flagHi = nullptr; L.epoch = 0;
  int64_t planeCells = 1;
  for (int a = 0; a < Phys::dim - 1; ++a) planeCells *= m.n[a];
  const int64_t nwork = planeCells * (pEnd - pBegin);
  if (nwork <= 0) return;
  if constexpr (Phys::dim == 3) {
    if (m.n[0] >= 2 * m.halo() && m.n[1] >= 2 * m.halo()) {
      launchLattice3dTiled<Phys, S>(phys, L, dl, dUlocal, dVowned, st);
      return;
    }
  }
  if constexpr (Phys::dim == 2) {
    launchMarch2d<Phys, S>(phys, L, dl, dUlocal, dVowned, st);
    return;
  }
  const int block = 128;
  dev::k_velocity_lattice_v1<Phys, S><<<(unsigned)((nwork + block - 1) / block), block, 0, st>>>(phys, L, dl, dUlocal, dVowned);
}

}  // namespace pda
