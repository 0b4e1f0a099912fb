// kernels_tiled.cuh -- the headline kernel: structured 3D Euler velocity with every face flux computed ONCE.
//
// Replaces the reference's inner-cell velocity loop (euler_3d_prob_class.hpp:861-927), which computes both faces of
// every axis per cell (each face flux twice).  The kernel is bound by the FP64 pipe (WENO5: ~33 flop/B,
// B200 balance ~5 flop/B), so the design minimises FP64 instructions per cell and keeps that pipe fed:
//
//   * a CTA owns a 32 x TY tile of (x,y) columns and MARCHES along z over LZ planes: the z-face flux of step k is
//     the bottom flux of step k+1 (registers) -> z faces computed once;
//   * the current plane, with its x/y stencil halo, is staged in shared memory by TMA (cp.async.bulk, one copy per
//     tile row, completion on an mbarrier, issued one z-face computation ahead of its use);
//   * x faces: each lane computes the left face of its cell, the right face arrives by warp shuffle from lane+1;
//     y faces: each thread computes the back face, the front face is read from shared memory after a barrier;
//     the TY + 32 tile-edge faces that no cell thread owns are computed by a ninth "edge" warp (two face tasks per
//     step against three for the cell warps: off the critical path, barriers see balanced work);
//   * the z stencil of a column lives in a thread-private shared-memory ring (2h planes + one in flight, fetched
//     by cp.async one step ahead): no barrier and no registers on the z path;
//   * ONE copy of the face code (reconstruction + Rusanov) serves all four face tasks through a task loop, so the
//     hot loop stays resident in the instruction cache;
//   * leaf arithmetic restated for the FP64 pipe: difference-form WENO5 with one reciprocal per (face, dof) for BOTH
//     sides, branch-free Newton reciprocal / square root from the MUFU seeds (no IEEE slow paths, no calls);
//   * V is written once per cell (no zero + "+=" passes, mixin_directional_flux_balance.hpp:68-84).
//
// Works for scheme stencils 3/5/7, periodic or not per axis (cells whose mesh stencil leaves a non-periodic domain
// belong to the near-boundary kernel and are not stored), and for slab-local storage (multi-GPU).
#pragma once
#include <cstdint>
#include <cstdlib>

#include "fastmath.cuh"
#include "func_attrs.hpp"
#include "kernels_lattice.cuh"

namespace pda {
namespace dev {

PDA_DEVFN void cpAsync8(void* smemDst, const void* gmemSrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmemSrc) : "memory");
}
PDA_DEVFN void cpAsyncCommit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
PDA_DEVFN void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

PDA_DEVFN int32_t fixIdx(int32_t idx, int32_t n, int32_t periodic) {   // |idx| may exceed n by less than n
  if (periodic) return idx < 0 ? idx + n : (idx >= n ? idx - n : idx);
  return idx < 0 ? 0 : (idx >= n ? n - 1 : idx);
}
// x/y halo of a ragged tile on a small periodic mesh may wrap more than once (the values only feed lanes that do not
// store, but the address must stay inside the allocation)
PDA_DEVFN int32_t wrapIdx(int32_t idx, int32_t n, int32_t periodic) {
  if (periodic) { idx %= n; return idx < 0 ? idx + n : idx; }
  return idx < 0 ? 0 : (idx >= n ? n - 1 : idx);
}

// 3D Euler Rusanov flux along axis `ax` (runtime), impl/euler_rusanov_flux_values_function.hpp:151-208:
// F = 1/2 (F(qL) + F(qR) + smax (qL - qR)), smax = |v_roe| + a_roe.   ~115 FP64 instructions, no calls.
PDA_DEVFN void eulerFlux3dFast(double gamma, int ax, const double* qL, const double* qR, double* F) {
  const double gm1 = gamma - 1.0;
  const double rL = qL[0], rR = qR[0];
  const double iL = rcpFast(rL), iR = rcpFast(rR);
  const double uL = qL[1] * iL, vL = qL[2] * iL, wL = qL[3] * iL;
  const double uR = qR[1] * iR, vR = qR[2] * iR, wR = qR[3] * iR;
  const double kL = fma(wL, wL, fma(vL, vL, uL * uL));
  const double kR = fma(wR, wR, fma(vR, vR, uR * uR));
  const double pL = gm1 * fma(-0.5 * rL, kL, qL[4]);
  const double pR = gm1 * fma(-0.5 * rR, kR, qR[4]);
  const double HL = (qL[4] + pL) * iL;
  const double HR = (qR[4] + pR) * iR;
  const double unL = (ax == 0) ? uL : ((ax == 1) ? vL : wL);
  const double unR = (ax == 0) ? uR : ((ax == 1) ? vR : wR);
  const double mL = rL * unL, mR = rR * unR;
  // Roe averages
  const double RT = sqrtFast(rR * iL);
  const double iRT = rcpFast(1.0 + RT);
  const double u = fma(RT, uR, uL) * iRT, v = fma(RT, vR, vL) * iRT, w = fma(RT, wR, wL) * iRT;
  const double H = fma(RT, HR, HL) * iRT;
  const double k = fma(w, w, fma(v, v, u * u));
  const double a = sqrtFast(gm1 * fma(-0.5, k, H));
  const double smax = sqrtFastTiny(k) + a;
  const double pS = pL + pR;
  F[0] = 0.5 * fma(smax, rL - rR, mL + mR);
  F[1] = 0.5 * (fma(smax, qL[1] - qR[1], fma(mL, uL, mR * uR)) + ((ax == 0) ? pS : 0.0));
  F[2] = 0.5 * (fma(smax, qL[2] - qR[2], fma(mL, vL, mR * vR)) + ((ax == 1) ? pS : 0.0));
  F[3] = 0.5 * (fma(smax, qL[3] - qR[3], fma(mL, wL, mR * wR)) + ((ax == 2) ? pS : 0.0));
  F[4] = 0.5 * fma(smax, qL[4] - qR[4], fma(mL, HL, mR * HR));
}

// ---- TMA (bulk async copy) + mbarrier wrappers: sm_90+ PTX, SASS UBLKCP / SYNCS
PDA_DEVFN unsigned smemU32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
PDA_DEVFN void mbarInit(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smemU32(bar)), "r"(count) : "memory");
}
PDA_DEVFN void mbarArriveExpectTx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smemU32(bar)), "r"(bytes) : "memory");
}
PDA_DEVFN void mbarWait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smemU32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
PDA_DEVFN void bulkCopyG2S(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smemU32(dst)), "l"(src), "r"(bytes), "r"(smemU32(bar)) : "memory");
}

template <int S, int TY>
struct Tile3dSmem {
  static constexpr int h = (S - 1) / 2;
  static constexpr int TX = 32;
  static constexpr int HX = (h + 1) & ~1;          // x halo rounded up to an even cell count: 16-byte aligned rows
  static constexpr int PX = TX + 2 * HX, PY = TY + 2 * h;
  static constexpr int R = 2 * h + 1;              // z ring: planes k-h+1 .. k+h in use, k+h+1 in flight
  template <int N> static constexpr size_t bytes() {
    return sizeof(double) * (size_t)(N * PY * PX + R * N * TY * TX + 2 * (N * (TY + 1) * TX + N * TY) + 2);
  }
};

// Thread block = TY "cell" warps (one tile row each) + ONE "edge" warp that (a) issues the TMA bulk copies of the
// next plane and (b) computes the tile-edge faces nobody owns (y face row TY, x face column 32): two face tasks per
// step against three for the cell warps, so it never sits on the critical path.
// All shared-memory tiles are AoS ([..][cell][dof], 40-byte cells): a lane stride of 40 bytes is conflict-free for
// 8-byte accesses, the dof index becomes an immediate offset, and a row of the plane tile is ONE contiguous range
// of the AoS state in HBM -> one cp.async.bulk per row (two where the periodic wrap splits it).
template <int S, int TY, bool PEER>
__global__ void __launch_bounds__(32 * (TY + 1), (TY <= 7 ? 2 : 1))
k_euler3d_velocity_tiled(double gamma, LatticeDesc L, Deltas dl, const double* __restrict__ U, double* __restrict__ V,
                         int LZ, int useTma) {
  constexpr int N = 5;
  using T = Tile3dSmem<S, TY>;
  constexpr int h = T::h, TX = T::TX, HX = T::HX, PX = T::PX, PY = T::PY, R = T::R;
  constexpr int NT = TX * (TY + 1);
  constexpr int oP = 0;                        // [PY][PX][N]      current plane with x/y halo
  constexpr int oZ = oP + N * PY * PX;         // [R][TY][TX][N]   thread-private z columns
  // flux exchange buffers, DOUBLE-BUFFERED by step parity: a warp may run a whole step ahead of the neighbour that still
  // reads its previous flux (only one barrier per step), so consecutive steps must not share a buffer
  constexpr int kFx = N * (TY + 1) * TX + N * TY;  // one exchange buffer: [TY+1][TX][N] y-face + [TY][N] tile-edge x-face fluxes
  constexpr int oFy0 = oZ + R * N * TY * TX;
  constexpr int oBar = oFy0 + 2 * kFx;         // mbarrier (8 bytes)
  constexpr int slotStride = N * TY * TX;

  extern __shared__ __align__(16) double smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(&smem[oBar + (oBar & 1)]);

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TX + tx;
  const bool edgeWarp = (ty == TY);
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  // peer mode: z chunks run in reverse order, so the chunk that needs the upper halo (at its END) starts first and
  // the one that needs the lower halo (at its START) last: the neighbours' pushes land while interior work runs
  const int zc = PEER ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int k0 = L.planeBegin + zc * LZ;
  const int k1 = min(k0 + LZ, L.planeEnd);
  const int nx = L.n[0], ny = L.n[1], nz = L.n[2];
  const int perZ = L.per[2];

  // storage plane of lattice plane p (slab: halo planes precede plane 0, no wrap)
  auto planeOf = [&](int p) -> int64_t { return L.slab ? (int64_t)(p + L.haloPlanes) : (int64_t)fixIdx(p, nz, perZ); };
  const int64_t rowStride = (int64_t)nx * N, planeStride = (int64_t)nx * ny * N;
  // base of plane p for the z-column fetches; peer mode: planes outside [0,nz) come from the halo buffers, once the
  // neighbour's flag carries this evaluation's epoch (lane 0 polls with acquire.sys, the warp follows)
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

  // own column (clamped into the domain for threads of a ragged tile; the edge warp has none)
  const int ci = min(x0 + tx, nx - 1), cj = min(y0 + min(ty, TY - 1), ny - 1);
  const int64_t colOff = ((int64_t)cj * nx + ci) * N;
  const int zMine = oZ + (min(ty, TY - 1) * TX + tx) * N;   // + slot*slotStride + d
  // thread-private async copy of this column's cell of plane p into ring slot `slot`
  auto fetchColumn = [&](int p, int slot) {
    const double* src = planeBase(p) + colOff;
    const int off = zMine + slot * slotStride;
#pragma unroll
    for (int d = 0; d < N; ++d) cpAsync8(&smem[off + d], src + d);
    cpAsyncCommit();
  };

  // plane p (with x/y halo) -> sP.  TMA path: lane r of the edge warp copies tile row r (cells x0-HX .. x0+TX+HX-1)
  // with one bulk copy per contiguous segment; fallback: every thread issues 8-byte cp.async copies.
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

  // ---- prologue: ring <- planes (k0-1)-h+1 .. (k0-1)+h ; slot of plane p = (p - (k0-h)) mod R
  if (!edgeWarp) {
#pragma unroll
    for (int o = 0; o < 2 * h; ++o) fetchColumn(k0 - h + o, o);
  }
  loadPlane(k0);
  int slot0 = 0;          // ring slot of plane k-h+1 at step k
  unsigned parity = 0;    // mbarrier phase of the plane awaited next

  double Fz[N];   // flux through the bottom face of the current cell
#pragma unroll
  for (int d = 0; d < N; ++d) Fz[d] = 0.0;

  const bool inX = (x0 + tx < nx) && (L.per[0] || (x0 + tx >= L.meshHalo && x0 + tx < nx - L.meshHalo));
  const bool inY = !edgeWarp && (y0 + ty < ny) && (L.per[1] || (y0 + ty >= L.meshHalo && y0 + ty < ny - L.meshHalo));

  for (int k = k0 - 1; k < k1; ++k) {
    const bool ghost = (k < k0);
    const int oFy = oFy0 + (k & 1) * kFx;
    const int oXe = oFy + N * (TY + 1) * TX;
    if (!edgeWarp) {
      cpAsyncWaitAll();   // this thread's column cell of plane k+h (fetched one step ago) has landed
      // prefetch the plane the NEXT z face needs into the free slot (slot0 + 2h) mod R
      if (k + 1 < k1) { int sl = slot0 + 2 * h; if (sl >= R) sl -= R; fetchColumn(k + 1 + h, sl); }
    }
    // face tasks: 0 = z face k+1/2 (thread-private), 1 = y back face, 2 = x left face (cell warps);
    //             3 = y faces of row TY, 4 = x faces of column TX (edge warp)
    const int tBegin = edgeWarp ? 3 : 0;
    const int tEnd = edgeWarp ? (ghost ? 3 : 5) : (ghost ? 1 : 3);
    double v[N], dFx[N];
#pragma unroll 1
    for (int task = tBegin; task < tEnd; ++task) {
      if (task == 1 || task == 3) {   // plane k (issued one z-face computation ago) must have landed
        if (useTma) { mbarWait(bar, parity); parity ^= 1u; }
        else { cpAsyncWaitAll(); __syncthreads(); }
      }
      int offs[2 * h];
      int ax;
      if (task == 0) {
        int sl = slot0;
#pragma unroll
        for (int o = 0; o < 2 * h; ++o) { offs[o] = zMine + sl * slotStride; sl = (sl + 1 == R) ? 0 : sl + 1; }
        ax = 2;
      } else if (task == 1 || task == 3) {
        const int row0 = (task == 1) ? ty : TY;
#pragma unroll
        for (int o = 0; o < 2 * h; ++o) offs[o] = oP + ((row0 + o) * PX + (tx + HX)) * N;
        ax = 1;
      } else {
        const int row = (task == 2) ? (ty + h) : (min(tx, TY - 1) + h);
        const int col0 = ((task == 2) ? tx : TX) + (HX - h);
#pragma unroll
        for (int o = 0; o < 2 * h; ++o) offs[o] = oP + (row * PX + col0 + o) * N;
        ax = 0;
      }
      double uN[N], uP[N], F[N];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        double q[2 * h];
#pragma unroll
        for (int o = 0; o < 2 * h; ++o) q[o] = smem[offs[o] + d];
        reconFaceFast<S>(q, uN[d], uP[d]);
      }
      eulerFlux3dFast(gamma, ax, uN, uP, F);
      if (task == 0) {
#pragma unroll
        for (int d = 0; d < N; ++d) { v[d] = dl.hInv[2] * (Fz[d] - F[d]); Fz[d] = F[d]; }   // z term, added last
      } else if (task == 1) {
#pragma unroll
        for (int d = 0; d < N; ++d) smem[oFy + (ty * TX + tx) * N + d] = F[d];
      } else if (task == 2) {
        // FxL - FxR, the right face being lane+1's left face (lane 31: FxL - 0 until the barrier: exact)
#pragma unroll
        for (int d = 0; d < N; ++d) {
          const double r = __shfl_down_sync(0xffffffffu, F[d], 1);
          dFx[d] = F[d] - ((tx == TX - 1) ? 0.0 : r);
        }
      } else if (task == 3) {
#pragma unroll
        for (int d = 0; d < N; ++d) smem[oFy + (TY * TX + tx) * N + d] = F[d];
      } else if (tx < TY) {
#pragma unroll
        for (int d = 0; d < N; ++d) smem[oXe + tx * N + d] = F[d];
      }
    }
    slot0 = (slot0 + 1 == R) ? 0 : slot0 + 1;
    if (ghost) continue;   // ghost step: only the bottom flux of the first plane

    __syncthreads();                      // fluxes exchanged; nobody reads the plane buffer any more
    if (k + 1 < k1) loadPlane(k + 1);     // lands while the next z face is computed
    if (edgeWarp) continue;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      if (tx == TX - 1) dFx[d] -= smem[oXe + ty * N + d];
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

// second generation (kernels_tiled2.cuh): cell-based WENO edges along z and x; PDA_TILED_V2=0 keeps this file's kernel
bool tiledV2Enabled();
template <class Phys, int S>
void launchLattice3dTiled2(const Phys& phys, const dev::LatticeDesc& L, const dev::Deltas& dl, const double* dU,
                           double* dV, cudaStream_t st);

template <class Phys, int S>
void launchLattice3dTiled(const Phys& phys, const dev::LatticeDesc& L, const dev::Deltas& dl, const double* dU,
                          double* dV, cudaStream_t st) {
  static_assert(Phys::dim == 3 && Phys::ndpc == 5, "Euler3d kernel");
  if (tiledV2Enabled()) { launchLattice3dTiled2<Phys, S>(phys, L, dl, dU, dV, st); return; }
  constexpr int TY = 7;   // 7 cell warps + 1 edge warp = 256 threads, 2 CTAs per SM at 128 registers
  using T = dev::Tile3dSmem<S, TY>;
  constexpr size_t smem = T::template bytes<5>();
  auto kern = (L.slab == 2) ? dev::k_euler3d_velocity_tiled<S, TY, true> : dev::k_euler3d_velocity_tiled<S, TY, false>;
  ensureFuncAttrs(kern, (int)smem, true);   // per device (func_attrs.hpp)
  const int planes = L.planeEnd - L.planeBegin;
  if (planes <= 0) return;
  const int gx = (L.n[0] + 31) / 32, gy = (L.n[1] + TY - 1) / TY;
  // z chunks: long enough to amortise the ghost step (1/LZ extra z faces), short enough to fill 148 SMs x 2 CTAs
  // (PDA_TILED_LZ overrides the start value: tuning only)
  static const int lzStart = [] { const char* e = std::getenv("PDA_TILED_LZ"); const int v = e ? std::atoi(e) : 0; return v >= 8 ? v : 64; }();
  int LZ = lzStart;
  while (LZ > 8 && (int64_t)gx * gy * ((planes + LZ - 1) / LZ) < 148 * 2 * 4) LZ /= 2;
  // peer mode: at least two chunks, so that no CTA needs both halos and the pushes overlap interior work
  if (L.slab == 2) while (LZ > 8 && (planes + LZ - 1) / LZ < 2) LZ /= 2;
  const int gz = (planes + LZ - 1) / LZ;
  dim3 grid(gx, gy, gz), block(32, TY + 1);
  // TMA rows need 16-byte aligned segments: even cell counts (40-byte cells) and an aligned base
  const int useTma = (L.n[0] % 2 == 0) && (L.n[0] >= T::PX) && ((reinterpret_cast<uintptr_t>(dU) & 15) == 0);
  kern<<<grid, block, smem, st>>>(phys.gamma, L, dl, dU, dV, LZ, useTma);
}

}  // namespace pda
