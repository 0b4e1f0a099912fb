// kernels_tiled2.cuh -- second generation of the headline kernel (structured 3D Euler velocity, every face flux once).
//
// Same tile / TMA / z-marching skeleton as k_euler3d_velocity_tiled (kernels_tiled.cuh; reference loop replaced:
// euler_3d_prob_class.hpp:861-927), with the FP64 instruction count per cell cut by 13 %:
//   * z and x faces use CELL-based WENO edge values (cellmath.cuh): the three smoothness indicators of a cell serve its
//     left AND right edge, so they are computed once per cell instead of once per adjacent face.
//       z: the thread that marches up a column sees cell k+1's whole stencil in its private ring; it computes
//          (eL, eR)(k+1), forms face k+1/2 from the carried eR(k) and eL(k+1), and carries eR(k+1) on (shared memory
//          slot, no registers across the step);
//       x: a lane computes (eL, eR) of its own cell from the plane tile; the left face of cell i is (eR(i-1) by
//          warp shuffle, eL(i)); the two tile-boundary faces of a row come from the edge warp, which evaluates the
//          same cell-based code for the two cells on either side (identical bits to an interior face);
//       y: face-based as before (the y neighbours live in other warps: sharing would need a second barrier);
//   * the 8-instruction square root (cellmath.cuh) in the Rusanov flux: 3 x 5 instructions fewer per face;
//   * ONE copy each of the cell reconstruction and the face reconstruction (two of the flux); inside a phase the five
//     dofs form one straight-line block so that their dependency chains interleave (a first version that branched per
//     dof ran at 58 % FP64-pipe utilisation against 73 %: profiles/ncu_velocity_r02.txt).
// FP64 instructions per cell: z 5*57+104, x 5*57+104, y 5*76+104 = 1262 (+ edge warp 9 %) against 1485 (+ 9.5 %).
#pragma once
#include <cstdint>
#include <cstdlib>

#include "cellmath.cuh"
#include "func_attrs.hpp"
#include "kernels_tiled.cuh"

namespace pda {
namespace dev {

template <int S, int TY>
struct Tile3dSmem2 {
  static constexpr int h = (S - 1) / 2;
  static constexpr int TX = 32;
  static constexpr int HX = (h + 1) & ~1;          // x halo rounded up to an even cell count: 16-byte aligned rows
  static constexpr int PX = TX + 2 * HX, PY = TY + 2 * h;
  static constexpr int R = 2 * h;                  // z ring: planes k+2-h .. k+h in use at step k, k+h+1 in flight
  template <int N> static constexpr int fluxDoubles() { return N * (TY + 1) * TX + 2 * N * TY; }
  template <int N> static constexpr size_t bytes() {
    return sizeof(double) * (size_t)(N * PY * PX + R * N * TY * TX + 2 * fluxDoubles<N>() + N * TY * TX + 2);
  }
};

template <int S, int TY, bool PEER>
__global__ void __launch_bounds__(32 * (TY + 1), (TY <= 7 ? 2 : 1))
k_euler3d_velocity_tiled2(double gamma, LatticeDesc L, Deltas dl, const double* __restrict__ U, double* __restrict__ V,
                          int LZ, int useTma) {
  constexpr int N = 5;
  using T = Tile3dSmem2<S, TY>;
  constexpr int h = T::h, hc = h - 1, TX = T::TX, HX = T::HX, PX = T::PX, PY = T::PY, R = T::R;
  constexpr int NT = TX * (TY + 1);
  constexpr int NQ = (2 * h > 1) ? 2 * h : 1;      // stencil values per reconstruction: 2h (face) or 2h-1 (cell)
  constexpr int oP = 0;                            // [PY][PX][N]      current plane with x/y halo
  constexpr int oZ = oP + N * PY * PX;             // [R][TY][TX][N]   thread-private z columns
  constexpr int kFx = T::template fluxDoubles<N>();   // one exchange buffer: [TY+1][TX][N] y faces + [TY][2][N] tile-edge x faces
  constexpr int oFy0 = oZ + R * N * TY * TX;       // two exchange buffers (step parity, see kernels_tiled.cuh)
  constexpr int oE = oFy0 + 2 * kFx;               // [TY][TX][N] carried right-edge value eR(k) of the z march
  constexpr int oBar = oE + N * TY * TX;           // mbarrier
  constexpr int slotStride = N * TY * TX;

  extern __shared__ __align__(16) double smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(&smem[oBar + (oBar & 1)]);

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TX + tx;
  const bool edgeWarp = (ty == TY);
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int zc = PEER ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;   // peer mode: halo-dependent chunks last
  const int k0 = L.planeBegin + zc * LZ;
  const int k1 = min(k0 + LZ, L.planeEnd);
  const int nx = L.n[0], ny = L.n[1], nz = L.n[2];
  const int perZ = L.per[2];

  auto planeOf = [&](int p) -> int64_t { return L.slab ? (int64_t)(p + L.haloPlanes) : (int64_t)fixIdx(p, nz, perZ); };
  const int64_t rowStride = (int64_t)nx * N, planeStride = (int64_t)nx * ny * N;
  auto planeBase = [&](int p) -> const double* {
    if constexpr (PEER) {
      if (p >= 0 && p < nz) return U + (int64_t)p * planeStride;
      const uint32_t* flag = (p < 0) ? L.flagLo : L.flagHi;
      if (tx == 0) {
        unsigned seen;
        do {
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(flag) : "memory");
          if (seen != L.epoch) __nanosleep(200);
        } while (seen != L.epoch);
      }
      __syncwarp();
      return (p < 0) ? L.haloLo + (int64_t)(p + h) * planeStride : L.haloHi + (int64_t)(p - nz) * planeStride;
    }
    return U + planeOf(p) * planeStride;
  };

  const int ci = min(x0 + tx, nx - 1), cj = min(y0 + min(ty, TY - 1), ny - 1);
  const int64_t colOff = ((int64_t)cj * nx + ci) * N;
  const int tyc = min(ty, TY - 1);
  const int zMine = oZ + (tyc * TX + tx) * N;      // + slot*slotStride + d
  const int eMine = oE + (tyc * TX + tx) * N;
  auto fetchColumn = [&](int p, int slot) {
    const double* src = planeBase(p) + colOff;
    const int off = zMine + slot * slotStride;
#pragma unroll
    for (int d = 0; d < N; ++d) cpAsync8(&smem[off + d], src + d);
    cpAsyncCommit();
  };

  auto loadPlane = [&](int p) {
    const double* src = U + planeOf(p) * planeStride;
    if (useTma) {
      if (!edgeWarp) return;
      unsigned bytes = 0;
      int segG[3], segD[3], segL[3], nseg = 0;
      if (tx < PY) {
        int start = x0 - HX, remaining = PX, dcol = 0;
        while (remaining > 0 && nseg < 3) {
          int g = start;
          if (g < 0) {
            if (L.per[0]) g += nx;
            else { const int skip = min(remaining, -g); start += skip; dcol += skip; remaining -= skip; continue; }
          } else if (g >= nx) {
            if (L.per[0]) g -= nx; else break;
          }
          const int len = min(remaining, nx - g);
          segG[nseg] = g; segD[nseg] = dcol; segL[nseg] = len; ++nseg;
          bytes += (unsigned)len * (N * 8);
          start += len; dcol += len; remaining -= len;
        }
      }
      unsigned total = bytes;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      if (tx == 0) mbarArriveExpectTx(bar, total);
      __syncwarp();
      if (tx < PY) {
        const int gy = wrapIdx(y0 - h + tx, ny, L.per[1]);
        for (int sI = 0; sI < nseg; ++sI)
          bulkCopyG2S(&smem[oP + (tx * PX + segD[sI]) * N], src + (int64_t)gy * rowStride + (int64_t)segG[sI] * N,
                      (unsigned)segL[sI] * (N * 8), bar);
      }
    } else {
      for (int e = tid; e < PY * (TX + 2 * h) * N; e += NT) {
        const int r = e / ((TX + 2 * h) * N);
        const int rem = e - r * ((TX + 2 * h) * N);
        const int cc = rem / N;
        const int d = rem - cc * N;
        const int gy = wrapIdx(y0 - h + r, ny, L.per[1]);
        const int gx = wrapIdx(x0 - h + cc, nx, L.per[0]);
        cpAsync8(&smem[oP + (r * PX + (HX - h) + cc) * N + d], src + (int64_t)gy * rowStride + (int64_t)gx * N + d);
      }
      cpAsyncCommit();
    }
  };

  if (tid == 0) mbarInit(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();

  // ---- prologue: ring <- planes k0-h .. k0+h-2 (what the first step, k = k0-2, reads); slot of plane p = (p-(k0-h)) mod R
  if (!edgeWarp) {
#pragma unroll
    for (int o = 0; o < 2 * h - 1; ++o) fetchColumn(k0 - h + o, o);
  }
  loadPlane(k0);
  int slot0 = 0;          // ring slot of plane k+2-h at step k
  unsigned parity = 0;    // mbarrier phase of the plane awaited next

  double Fz[N];   // flux through the bottom face of the current cell
#pragma unroll
  for (int d = 0; d < N; ++d) Fz[d] = 0.0;

  const bool inX = (x0 + tx < nx) && (L.per[0] || (x0 + tx >= L.meshHalo && x0 + tx < nx - L.meshHalo));
  const bool inY = !edgeWarp && (y0 + ty < ny) && (L.per[1] || (y0 + ty >= L.meshHalo && y0 + ty < ny - L.meshHalo));
  // edge warp, x phase: lanes [0,TY) own the RIGHT tile-boundary face of row `lane`, lanes [TY,2TY) the LEFT one
  const int eRow = (tx < TY) ? tx : min(tx - TY, TY - 1);
  const int eSide = (tx < TY) ? 1 : 0;             // 1: face between columns TX-1 | TX ; 0: between -1 | 0

  for (int k = k0 - 2; k < k1; ++k) {
    const bool ghostA = (k == k0 - 2), ghostB = (k == k0 - 1), ghost = (k < k0);
    const int oFy = oFy0 + (k & 1) * kFx;
    const int oXe = oFy + N * (TY + 1) * TX;
    if (!edgeWarp) {
      cpAsyncWaitAll();   // this thread's column cell of plane k+h (fetched one step ago) has landed
      if (k + 1 < k1) { int sl = slot0 + 2 * h - 1; if (sl >= R) sl -= R; fetchColumn(k + 1 + h, sl); }
    }
    // Two CELL-based phases, then one FACE-based phase; inside a phase the five dofs are one straight-line block (the
    // compiler interleaves their dependency chains), control flow only between phases:
    //   cell warps: Z (every step), X, then Y;   edge warp: the two cells of its tile-boundary x face (XEA, XEB), then YE
    double v[N], dFx[N], uN[N], uP[N];
    const int nCellPhases = edgeWarp ? (ghost ? 0 : 2) : (ghost ? 1 : 2);
#pragma unroll 1
    for (int it = 0; it < nCellPhases; ++it) {
      const bool phZ = !edgeWarp && it == 0;
      if (!phZ && (edgeWarp ? it == 0 : it == 1)) {   // first phase of the step that reads the plane tile
        if (useTma) { mbarWait(bar, parity); parity ^= 1u; }
        else { cpAsyncWaitAll(); __syncthreads(); }
      }
      int offs[NQ];
      if (phZ) {
        int sl = slot0;
#pragma unroll
        for (int o = 0; o < 2 * h - 1; ++o) { offs[o] = zMine + sl * slotStride; sl = (sl + 1 == R) ? 0 : sl + 1; }
      } else {
        // cell whose edge values are wanted, in tile columns: own cell (X); left / right cell of the boundary face (XEA / XEB)
        const int row = edgeWarp ? (eRow + h) : (ty + h);
        const int cc = edgeWarp ? ((eSide ? TX - 1 : -1) + it) : tx;
#pragma unroll
        for (int o = 0; o < 2 * h - 1; ++o) offs[o] = oP + (row * PX + (cc + HX - hc + o)) * N;
      }
      double eL[N], eR[N];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        double q[NQ];
#pragma unroll
        for (int o = 0; o < 2 * h - 1; ++o) q[o] = smem[offs[o] + d];
        cellEdgesFast<S>(q, eL[d], eR[d]);
      }
      if (phZ) {                    // face k+1/2 = (eR(k) carried, eL(k+1)); carry eR(k+1)
#pragma unroll
        for (int d = 0; d < N; ++d) { uP[d] = eL[d]; uN[d] = smem[eMine + d]; smem[eMine + d] = eR[d]; }
        if (ghostA) continue;       // no face yet
      } else if (!edgeWarp) {       // X: left face of my cell = (eR of lane-1, my eL)
#pragma unroll
        for (int d = 0; d < N; ++d) { uP[d] = eL[d]; uN[d] = __shfl_up_sync(0xffffffffu, eR[d], 1); }
      } else if (it == 0) {         // XEA: left cell of the boundary face
#pragma unroll
        for (int d = 0; d < N; ++d) uN[d] = eR[d];
        continue;
      } else {                      // XEB: right cell
#pragma unroll
        for (int d = 0; d < N; ++d) uP[d] = eL[d];
      }
      double F[N];
      eulerFlux3dFast8(gamma, phZ ? 2 : 0, uN, uP, F);
      if (phZ) {
#pragma unroll
        for (int d = 0; d < N; ++d) { v[d] = dl.hInv[2] * (Fz[d] - F[d]); Fz[d] = F[d]; }   // z term, added last
      } else if (!edgeWarp) {
        // FxL - FxR: lane 0's own left flux is meaningless (no lane -1) and lane 31 has no lane +1: both tile-boundary
        // fluxes come from the edge warp after the barrier (x + 0 is exact until then)
#pragma unroll
        for (int d = 0; d < N; ++d) {
          const double r = __shfl_down_sync(0xffffffffu, F[d], 1);
          dFx[d] = ((tx == 0) ? 0.0 : F[d]) - ((tx == TX - 1) ? 0.0 : r);
        }
      } else if (tx < 2 * TY) {
#pragma unroll
        for (int d = 0; d < N; ++d) smem[oXe + (eRow * 2 + eSide) * N + d] = F[d];
      }
    }
    if (!ghost) {   // FACE-based phase: y back face of my cell (cell warps) / y faces of row TY (edge warp)
      const int row0 = edgeWarp ? TY : ty;
      int offs[NQ];
#pragma unroll
      for (int o = 0; o < 2 * h; ++o) offs[o] = oP + ((row0 + o) * PX + (tx + HX)) * N;
#pragma unroll
      for (int d = 0; d < N; ++d) {
        double q[NQ];
#pragma unroll
        for (int o = 0; o < 2 * h; ++o) q[o] = smem[offs[o] + d];
        reconFaceFast<S>(q, uN[d], uP[d]);
      }
      double F[N];
      eulerFlux3dFast8(gamma, 1, uN, uP, F);
#pragma unroll
      for (int d = 0; d < N; ++d) smem[oFy + (row0 * TX + tx) * N + d] = F[d];
    }
    slot0 = (slot0 + 1 == R) ? 0 : slot0 + 1;
    if (ghost) continue;   // ghost steps: only the carried edge value / the bottom flux of the first plane

    __syncthreads();                      // fluxes exchanged; nobody reads the plane buffer any more
    if (k + 1 < k1) loadPlane(k + 1);     // lands while the next z face is computed
    if (edgeWarp) continue;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      if (tx == 0) dFx[d] += smem[oXe + (ty * 2 + 0) * N + d];
      if (tx == TX - 1) dFx[d] -= smem[oXe + (ty * 2 + 1) * N + d];
      const double FyB = smem[oFy + (ty * TX + tx) * N + d];
      const double FyF = smem[oFy + ((ty + 1) * TX + tx) * N + d];
      // V = hx(FxL-FxR) + hy(FyB-FyF) + hz(FzB-FzT): same x,y,z accumulation order as the reference
      v[d] = (dl.hInv[0] * dFx[d] + dl.hInv[1] * (FyB - FyF)) + v[d];
    }
    const bool inZ = L.slab || perZ || (k >= L.meshHalo && k < nz - L.meshHalo);
    if (inX && inY && inZ) {
      double* out = V + (((int64_t)k * ny + (y0 + ty)) * nx + (x0 + tx)) * N;
#pragma unroll
      for (int d = 0; d < N; ++d) out[d] = v[d];
    }
  }
}

}  // namespace dev

// which generation the 3D lattice velocity uses (PDA_TILED_V2=0 selects the first one: A/B measurements)
inline bool tiledV2Enabled() {
  static const bool on = [] { const char* e = std::getenv("PDA_TILED_V2"); return !(e && e[0] == '0'); }();
  return on;
}

template <class Phys, int S, int TY>
void launchLattice3dTiled2T(const Phys& phys, const dev::LatticeDesc& L, const dev::Deltas& dl, const double* dU,
                            double* dV, cudaStream_t st) {
  using T = dev::Tile3dSmem2<S, TY>;
  constexpr size_t smem = T::template bytes<5>();
  auto kern = (L.slab == 2) ? dev::k_euler3d_velocity_tiled2<S, TY, true> : dev::k_euler3d_velocity_tiled2<S, TY, false>;
  ensureFuncAttrs(kern, (int)smem, true);
  const int planes = L.planeEnd - L.planeBegin;
  if (planes <= 0) return;
  const int gx = (L.n[0] + 31) / 32, gy = (L.n[1] + TY - 1) / TY;
  // z chunks: long enough to amortise the two ghost steps, short enough to fill the SMs with >= 4 waves
  constexpr int ctasPerSm = (TY <= 7) ? 2 : 1;
  static const int lzStart = [] { const char* e = std::getenv("PDA_TILED_LZ"); const int v = e ? std::atoi(e) : 0; return v >= 8 ? v : 64; }();
  int LZ = lzStart;
  // long slabs: 128-plane chunks halve the ghost-step share as long as >= 16 waves remain (512^3: 14.84 -> 14.76 ms)
  if (lzStart == 64 && L.slab != 2 && (int64_t)gx * gy * (planes / 128) >= 148 * ctasPerSm * 16) LZ = 128;
  while (LZ > 8 && (int64_t)gx * gy * ((planes + LZ - 1) / LZ) < 148 * ctasPerSm * 4) LZ /= 2;
  // peer mode: at least 2 z chunks, so that no CTA needs both halos (the chunk that needs the lower halo at its start is
  // scheduled last); PDA_PEER_MINCHUNKS=1 lifts that for experiments
  static const int minChunks = [] { const char* e = std::getenv("PDA_PEER_MINCHUNKS"); const int v = e ? std::atoi(e) : 2; return v >= 1 ? v : 2; }();
  if (L.slab == 2) while (LZ > 8 && (planes + LZ - 1) / LZ < minChunks) LZ /= 2;
  const int gz = (planes + LZ - 1) / LZ;
  dim3 grid(gx, gy, gz), block(32, TY + 1);
  const int useTma = (L.n[0] % 2 == 0) && (L.n[0] >= T::PX) && ((reinterpret_cast<uintptr_t>(dU) & 15) == 0);
  kern<<<grid, block, smem, st>>>(phys.gamma, L, dl, dU, dV, LZ, useTma);
}

template <class Phys, int S>
void launchLattice3dTiled2(const Phys& phys, const dev::LatticeDesc& L, const dev::Deltas& dl, const double* dU,
                           double* dV, cudaStream_t st) {
  static_assert(Phys::dim == 3 && Phys::ndpc == 5, "Euler3d kernel");
  // TY = 7: 7 cell warps + 1 edge warp = 256 threads, 2 CTAs per SM at 128 registers (default).
  // TY = 15 (PDA_TILED_TY=15): 15 + 1 warps = 512 threads, ONE CTA per SM: half the edge-warp and y-halo overhead, a
  // 16-warp barrier -- an experiment, WENO5 only
  static const int ty = [] { const char* e = std::getenv("PDA_TILED_TY"); return e ? std::atoi(e) : 7; }();
  if constexpr (S == 7) {
    if (ty == 15 && L.slab != 2) { launchLattice3dTiled2T<Phys, S, 15>(phys, L, dl, dU, dV, st); return; }
  }
  launchLattice3dTiled2T<Phys, S, 7>(phys, L, dl, dU, dV, st);
}

}  // namespace pda
