// kernels_reforder.hpp -- launch interface of the REFERENCE-ORDER Jacobian kernels (kernels_reforder.cu, compiled
// with -fmad=false).  Selected with pda_problem_set_option(p, "jacobian_order", "reference"); see that file.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace pda {
namespace dev {

struct RowSet;
struct JacLayout;
struct GhostView;

// everything the reference-order kernels need to know about the problem (plain data, passed by value)
struct RefOrderParams {
  int family;          // PDA_FAMILY_* (1 Euler1d, 2 Euler2d, 3 Euler3d, 4 Swe2d, 6 Burgers, 7 ADR 2D, 8 advection 1D)
  int ndpc, S, dim;
  double gamma;        // Euler
  double gravity, coriolis;   // shallow water
  double adv[2];       // linear advection velocity per axis
  double diffusion;    // Burgers / ADR
  double sigma;        // ADR reaction
  double dInv[3];
  const double* src;   // ADR per-sample-row source table (device) or null -> 1.0
};

// inner rows: velocity (optional) + Jacobian with the reference's formulas, operation order and accumulation order;
// Jv must be zero on entry for these rows.  near-boundary rows: first-order Jacobian with ghost factors.
void launchRefOrderInner(const RefOrderParams& P, const int32_t* graph, const int32_t* rowIds, int32_t nRows, int ncols,
                         const double* U, double* V, double* Jv, const int32_t* jBase, const int32_t* jLen,
                         const uint8_t* jSlot, int nslotCols, cudaStream_t st);
void launchRefOrderNearBd(const RefOrderParams& P, const int32_t* graph, const int32_t* rowIds, int32_t nRows, int ncols,
                          const double* U, double* Jv, const int32_t* jBase, const int32_t* jLen, const uint8_t* jSlot,
                          int nslotCols, double* const ghost[6], int ghostStride, const double* factors,
                          cudaStream_t st);
// velocity only, reference operation order (option "velocity_order" = "reference"); nearBd rows read ghost states
void launchRefOrderVelocity(const RefOrderParams& P, const int32_t* graph, const int32_t* rowIds, int32_t nRows, int ncols,
                            const double* U, double* V, double* const ghost[6], int ghostStride, bool nearBd,
                            cudaStream_t st);
// the diffusion-reaction families (no reconstruction; the reference's operation order, nothing contracted)
void launchRefOrderGrayScott(const double gs[4], double dxInv, double dyInv, const int32_t* graph, const int32_t* rowIds,
                             int32_t nRows, int ncols, const double* U, double* V, double* Jv, const int32_t* jBase,
                             const int32_t* jLen, const uint8_t* jSlot, int nslotCols, cudaStream_t st);
void launchRefOrderDiffReac(int dim, double D, double kR, double dxInv, double dyInv, const int32_t* graph,
                            const int32_t* rowIds, int32_t nRows, int ncols, const double* U, const double* src, double* V,
                            double* Jv, const int32_t* jBase, const int32_t* jLen, const uint8_t* jSlot, int nslotCols,
                            cudaStream_t st);
// test hook: out[i] = the device restatement of glibc's pow(x[i], y)
void launchGlibcPow(const double* x, double y, double* out, int64_t n, cudaStream_t st);

}  // namespace dev
}  // namespace pda
