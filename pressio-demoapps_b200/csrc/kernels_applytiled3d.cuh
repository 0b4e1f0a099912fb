// kernels_applytiled3d.cuh -- matrix-free J*b of the 3D Euler problems on full lattices for ONE contiguous operand column
// (vector operands, columns of a column-major operand: the Newton-Krylov J*v), on the skeleton of the headline velocity
// kernel (kernels_tiled.cuh): tiles staged by TMA, z marching, every face once, 100 % of the lanes on cells.
//
// Replaces Eigen's  J * b  (adapter_cpp.hpp:231-259) where no Jacobian can be stored (512^3 WENO5: 6.4e10 entries).
// With dN[j] = sum_m d(uNeg_j)/d(q_m) b_m[j] (directional derivative of the reconstruction along b) the row block of a
// cell is  R_c = sum_axes hInv (D_lower - D_upper),  D_f = JN dN + JP dP  the directional derivative of face f's
// Rusanov flux -- evaluated as ONE Jacobian-vector product (eulerFluxJvpFast: no N x N matrices).  That is the velocity
// kernel with (value, tangent) pairs: the state AND the operand travel through the same tiles / z rings, the flux
// exchange buffers carry tangents only.
//
// The first matrix-free kernel (k_applyjac_lattice3d, kernels_applylattice.cuh: 7^3 tiles, a line of 7 cells + closing
// face on 8 lanes, every stencil cell loaded from global memory by the lane that needs it) ran the 512^3 WENO5 J*v in 79
// ms: 7/8 x 49/56 of the lanes useful, FP64 pipe 52 %, long_scoreboard 2.5 of 8 stall cycles per issue.  It stays for
// row-major multi-column operands.
#pragma once
#include "kernels_applylattice.cuh"
#include "kernels_tiled.cuh"

namespace pda {
namespace dev {

template <int S, int TY>
struct ApplyTile3dSmem {
  static constexpr int N = 5;
  static constexpr int h = (S - 1) / 2;
  static constexpr int TX = 32;
  static constexpr int HX = (h + 1) & ~1;
  static constexpr int PX = TX + 2 * HX, PY = TY + 2 * h;
  static constexpr int R = 2 * h + 1;
  static constexpr int kPlane = N * PY * PX;            // one field's plane tile
  static constexpr int kRing = R * N * TY * TX;         // one field's z ring
  static constexpr int kFx = N * (TY + 1) * TX + N * TY;
  static constexpr size_t bytes = sizeof(double) * (size_t)(2 * kPlane + 2 * kRing + 2 * kFx + 2);
};

template <int S, int TY>
__global__ void __launch_bounds__(32 * (TY + 1), 1)
k_applyjac_tiled3d(double gamma, LatticeDesc L, Deltas dl, const double* __restrict__ U, const double* __restrict__ B,
                   double* __restrict__ Rout, int LZ, int useTma) {
  constexpr int N = 5;
  using T = ApplyTile3dSmem<S, TY>;
  constexpr int h = T::h, TX = T::TX, HX = T::HX, PX = T::PX, PY = T::PY, R = T::R;
  constexpr int NT = TX * (TY + 1);
  constexpr int M = 2 * h;
  constexpr int oP = 0;                         // [PY][PX][N]   current plane of the STATE with x/y halo
  constexpr int oPB = oP + T::kPlane;           //               ... of the OPERAND
  constexpr int oZ = oPB + T::kPlane;           // [R][TY][TX][N] thread-private z columns of the state
  constexpr int oZB = oZ + T::kRing;            //               ... of the operand
  constexpr int kFx = T::kFx;                   // tangent-flux exchange buffers, double-buffered by step parity
  constexpr int oFy0 = oZB + T::kRing;
  constexpr int oBar = oFy0 + 2 * kFx;
  constexpr int slotStride = N * TY * TX;

  extern __shared__ __align__(16) double smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(&smem[oBar + (oBar & 1)]);

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TX + tx;
  const bool edgeWarp = (ty == TY);
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int k0 = L.planeBegin + (int)blockIdx.z * LZ;
  const int k1 = min(k0 + LZ, L.planeEnd);
  const int nx = L.n[0], ny = L.n[1], nz = L.n[2];
  const int perZ = L.per[2];
  auto planeOf = [&](int p) -> int64_t { return (int64_t)fixIdx(p, nz, perZ); };
  const int64_t rowStride = (int64_t)nx * N, planeStride = (int64_t)nx * ny * N;

  const int ci = min(x0 + tx, nx - 1), cj = min(y0 + min(ty, TY - 1), ny - 1);
  const int64_t colOff = ((int64_t)cj * nx + ci) * N;
  const int zMine = oZ + (min(ty, TY - 1) * TX + tx) * N;
  auto fetchColumn = [&](int p, int slot) {
    const int64_t g = planeOf(p) * planeStride + colOff;
    const int off = zMine + slot * slotStride;
#pragma unroll
    for (int d = 0; d < N; ++d) cpAsync8(&smem[off + d], U + g + d);
#pragma unroll
    for (int d = 0; d < N; ++d) cpAsync8(&smem[off + (oZB - oZ) + d], B + g + d);
    cpAsyncCommit();
  };

  auto loadPlane = [&](int p) {
    const int64_t pb = planeOf(p) * planeStride;
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
          bytes += 2u * (unsigned)len * (N * 8);
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
        for (int sI = 0; sI < nseg; ++sI) {
          const int64_t g = pb + (int64_t)gy * rowStride + (int64_t)segG[sI] * N;
          const int dst = (tx * PX + segD[sI]) * N;
          bulkCopyG2S(&smem[oP + dst], U + g, (unsigned)segL[sI] * (N * 8), bar);
          bulkCopyG2S(&smem[oPB + dst], B + g, (unsigned)segL[sI] * (N * 8), bar);
        }
      }
    } else {
      for (int e = tid; e < PY * (TX + 2 * h) * N; e += NT) {
        const int r = e / ((TX + 2 * h) * N);
        const int rem = e - r * ((TX + 2 * h) * N);
        const int cc = rem / N;
        const int d = rem - cc * N;
        const int gy = wrapIdx(y0 - h + r, ny, L.per[1]);
        const int gx = wrapIdx(x0 - h + cc, nx, L.per[0]);
        const int64_t g = pb + (int64_t)gy * rowStride + (int64_t)gx * N + d;
        const int dst = (r * PX + (HX - h) + cc) * N + d;
        cpAsync8(&smem[oP + dst], U + g);
        cpAsync8(&smem[oPB + dst], B + g);
      }
      cpAsyncCommit();
    }
  };

  if (tid == 0) mbarInit(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();

  if (!edgeWarp) {
#pragma unroll
    for (int o = 0; o < 2 * h; ++o) fetchColumn(k0 - h + o, o);
  }
  loadPlane(k0);
  int slot0 = 0;
  unsigned parity = 0;

  double Dz[N];   // tangent flux through the bottom face of the current cell
#pragma unroll
  for (int d = 0; d < N; ++d) Dz[d] = 0.0;

  const bool inX = (x0 + tx < nx) && (L.per[0] || (x0 + tx >= L.meshHalo && x0 + tx < nx - L.meshHalo));
  const bool inY = !edgeWarp && (y0 + ty < ny) && (L.per[1] || (y0 + ty >= L.meshHalo && y0 + ty < ny - L.meshHalo));

  for (int k = k0 - 1; k < k1; ++k) {
    const bool ghost = (k < k0);
    const int oFy = oFy0 + (k & 1) * kFx;
    const int oXe = oFy + N * (TY + 1) * TX;
    if (!edgeWarp) {
      cpAsyncWaitAll();
      if (k + 1 < k1) { int sl = slot0 + 2 * h; if (sl >= R) sl -= R; fetchColumn(k + 1 + h, sl); }
    }
    // face tasks: 0 = z face k+1/2 (thread-private), 1 = y back face, 2 = x left face (cell warps);
    //             3 = y faces of row TY, 4 = x faces of column TX (edge warp)
    const int tBegin = edgeWarp ? 3 : 0;
    const int tEnd = edgeWarp ? (ghost ? 3 : 5) : (ghost ? 1 : 3);
    double v[N], dFx[N];
#pragma unroll 1
    for (int task = tBegin; task < tEnd; ++task) {
      if (task == 1 || task == 3) {
        if (useTma) { mbarWait(bar, parity); parity ^= 1u; }
        else { cpAsyncWaitAll(); __syncthreads(); }
      }
      int offs[M];
      int ax, offB;
      if (task == 0) {
        int sl = slot0;
#pragma unroll
        for (int o = 0; o < M; ++o) { offs[o] = zMine + sl * slotStride; sl = (sl + 1 == R) ? 0 : sl + 1; }
        ax = 2; offB = oZB - oZ;
      } else if (task == 1 || task == 3) {
        const int row0 = (task == 1) ? ty : TY;
#pragma unroll
        for (int o = 0; o < M; ++o) offs[o] = oP + ((row0 + o) * PX + (tx + HX)) * N;
        ax = 1; offB = oPB - oP;
      } else {
        const int row = (task == 2) ? (ty + h) : (min(tx, TY - 1) + h);
        const int col0 = ((task == 2) ? tx : TX) + (HX - h);
#pragma unroll
        for (int o = 0; o < M; ++o) offs[o] = oP + (row * PX + col0 + o) * N;
        ax = 0; offB = oPB - oP;
      }
      // face states and their directional derivatives along the operand
      double uN[N], uP[N], dN[1][N], dP[1][N];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        double q[M], gN[M], gP[M];
#pragma unroll
        for (int o = 0; o < M; ++o) q[o] = smem[offs[o] + d];
        reconFaceValGradFast<S>(q, uN[d], uP[d], gN, gP);
        double sN = 0.0, sP = 0.0;
#pragma unroll
        for (int o = 0; o < M; ++o) {
          const double b = smem[offs[o] + offB + d];
          sN = fma(gN[o], b, sN);
          sP = fma(gP[o], b, sP);
        }
        dN[0][d] = sN; dP[0][d] = sP;
      }
      // tangent of the Rusanov flux along `ax`: P JVP_x(P uN, P uP; P dN, P dP) with the momentum swap P
      swapMomentum(ax, uN); swapMomentum(ax, uP); swapMomentum(ax, dN[0]); swapMomentum(ax, dP[0]);
      double F[1][N];
      eulerFluxJvpFast<3, 0, 1>(gamma, uN, uP, dN, dP, F);
      swapMomentum(ax, F[0]);
      if (task == 0) {
#pragma unroll
        for (int d = 0; d < N; ++d) { v[d] = dl.hInv[2] * (Dz[d] - F[0][d]); Dz[d] = F[0][d]; }
      } else if (task == 1) {
#pragma unroll
        for (int d = 0; d < N; ++d) smem[oFy + (ty * TX + tx) * N + d] = F[0][d];
      } else if (task == 2) {
#pragma unroll
        for (int d = 0; d < N; ++d) {
          const double r = __shfl_down_sync(0xffffffffu, F[0][d], 1);
          dFx[d] = F[0][d] - ((tx == TX - 1) ? 0.0 : r);
        }
      } else if (task == 3) {
#pragma unroll
        for (int d = 0; d < N; ++d) smem[oFy + (TY * TX + tx) * N + d] = F[0][d];
      } else if (tx < TY) {
#pragma unroll
        for (int d = 0; d < N; ++d) smem[oXe + tx * N + d] = F[0][d];
      }
    }
    slot0 = (slot0 + 1 == R) ? 0 : slot0 + 1;
    if (ghost) continue;

    __syncthreads();
    if (k + 1 < k1) loadPlane(k + 1);
    if (edgeWarp) continue;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      if (tx == TX - 1) dFx[d] -= smem[oXe + ty * N + d];
      const double FyB = smem[oFy + (ty * TX + tx) * N + d];
      const double FyF = smem[oFy + ((ty + 1) * TX + tx) * N + d];
      v[d] = (dl.hInv[0] * dFx[d] + dl.hInv[1] * (FyB - FyF)) + v[d];      // x, y, z: the reference's accumulation order
    }
    const bool inZ = perZ || (k >= L.meshHalo && k < nz - L.meshHalo);
    if (inX && inY && inZ) {
      double* out = Rout + (((int64_t)k * ny + (y0 + ty)) * nx + (x0 + tx)) * N;
#pragma unroll
      for (int d = 0; d < N; ++d) out[d] = v[d];
    }
  }
}

}  // namespace dev

// PDA_APPLY3D_TILED=0 keeps the line kernel (A/B measurements)
inline bool applyTiled3dEnabled() {
  static const bool on = [] { const char* e = std::getenv("PDA_APPLY3D_TILED"); return !(e && e[0] == '0'); }();
  return on;
}

// one contiguous operand column: dB / dR are AoS fields like the state
template <int S>
void launchApplyTiled3d(double gamma, const dev::LatticeDesc& L, const dev::Deltas& dl, const double* dU, const double* dB,
                        double* dR, cudaStream_t st) {
  constexpr int TY = 7;
  using T = dev::ApplyTile3dSmem<S, TY>;
  auto kern = dev::k_applyjac_tiled3d<S, TY>;
  ensureFuncAttrs(kern, (int)T::bytes, true);
  const int planes = L.planeEnd - L.planeBegin;
  if (planes <= 0) return;
  const int gx = (L.n[0] + 31) / 32, gy = (L.n[1] + TY - 1) / TY;
  int LZ = 64;
  while (LZ > 8 && (int64_t)gx * gy * ((planes + LZ - 1) / LZ) < 148 * 4) LZ /= 2;   // one CTA per SM, >= 4 waves
  const int gz = (planes + LZ - 1) / LZ;
  dim3 grid(gx, gy, gz), block(32, TY + 1);
  const int useTma = (L.n[0] % 2 == 0) && (L.n[0] >= T::PX) && ((reinterpret_cast<uintptr_t>(dU) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dB) & 15) == 0);
  kern<<<grid, block, T::bytes, st>>>(gamma, L, dl, dU, dB, dR, LZ, useTma);
}

}  // namespace pda
