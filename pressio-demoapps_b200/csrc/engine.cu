// engine.cu -- Problem: host logic (parameters, initial conditions, CSR pattern, ghost recipes) and the CUDA
// launches of the velocity / Jacobian evaluation.  No CPU fallback: without a device every evaluation throws.
#include "engine.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

#include "common.hpp"
#include "func_attrs.hpp"
#include "kernels_generic.cuh"
#include "kernels_lattice.cuh"
#include "kernels_tiled.cuh"
#include "kernels_tiled2.cuh"
#include "kernels_jacobian.cuh"
#include "kernels_jaclattice.cuh"
#include "kernels_march2d.cuh"
#include "kernels_applylattice.cuh"
#include "kernels_applytiled3d.cuh"
#include "kernels_applymarch2d.cuh"
#include "kernels_reforder.hpp"

namespace pda {

namespace {

#define PDA_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      throw Error(kCuda, std::string(#call) + " failed: " + cudaGetErrorString(e_));                     \
  } while (0)


template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  void alloc(size_t count) {
    if (count <= n && p) return;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    if (count == 0) return;
    PDA_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void upload(const std::vector<T>& v) {
    alloc(v.size());
    if (!v.empty()) PDA_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
};

template <class T>
struct PinnedBuf {
  T* p = nullptr;
  size_t n = 0;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  void alloc(size_t count) {
    if (count <= n && p) return;
    if (p) cudaFreeHost(p);
    p = nullptr; n = 0;
    if (count == 0) return;
    PDA_CUDA(cudaMallocHost(&p, count * sizeof(T)));
    n = count;
  }
};

// family / problem ids: include/pda_b200.h
enum { F_EULER1D = 1, F_EULER2D = 2, F_EULER3D = 3, F_SWE2D = 4, F_DIFFREAC2D = 5, F_ADVDIFF2D = 6,
       F_ADVDIFFREAC2D = 7, F_ADVECTION1D = 8, F_DIFFREAC1D = 9 };
enum { E2_PERIODIC = 0, E2_KH = 1, E2_SEDOV_FULL = 2, E2_SEDOV_SYM = 3, E2_RIEMANN = 4, E2_NORMAL_SHOCK = 5,
       E2_DMR = 6, E2_CROSS_SHOCK = 7, E2_NEUMANN = 8 };
enum { BC_DIRICHLET = 0, BC_NEUMANN = 1, BC_REFLECTIVE = 2, BC_CALLBACK = 3 };

// Euler2d parameter indices (impl/euler_2d_parametrization_helpers.hpp:59-72)
enum { icNormalShockMach = 0, icCrossDensity = 1, icCrossInletX = 2, icCrossBottomY = 3, icRiem1TRP = 4,
       icRiem2TRP = 5, icRiem2TRU = 6, icRiem2TRV = 7, icRiem2TRD = 8, icRiem2BLP = 9 };

int euler2dIcIndex(int prob, int icFlag, const std::string& s) {
  if (prob == E2_NORMAL_SHOCK && icFlag == 1 && s == "mach") return icNormalShockMach;
  if (prob == E2_CROSS_SHOCK && icFlag == 1) {
    if (s == "crossShockDensity") return icCrossDensity;
    if (s == "crossShockInletXVel") return icCrossInletX;
    if (s == "crossShockBottomYVel") return icCrossBottomY;
  }
  if (prob == E2_RIEMANN && icFlag == 1 && s == "riemannTopRightPressure") return icRiem1TRP;
  if (prob == E2_RIEMANN && icFlag == 2) {
    if (s == "riemannTopRightPressure") return icRiem2TRP;
    if (s == "riemannTopRightXVel") return icRiem2TRU;
    if (s == "riemannTopRightYVel") return icRiem2TRV;
    if (s == "riemannTopRightDensity") return icRiem2TRD;
    if (s == "riemannBotLeftPressure") return icRiem2BLP;
  }
  return -1;
}
// Swe2d (impl/swe_2d_parametrization_helpers.hpp:59-72)
int sweIcIndex(int icFlag, const std::string& s) {
  if (icFlag == 1) {
    if (s == "pulseMagnitude") return 0;
    if (s == "pulseX") return 1;
    if (s == "pulseY") return 2;
  } else if (icFlag == 2) {
    if (s == "pulseMagnitude1") return 3;
    if (s == "pulseX1") return 4;
    if (s == "pulseY1") return 5;
    if (s == "pulseMagnitude2") return 6;
    if (s == "pulseX2") return 7;
    if (s == "pulseY2") return 8;
  }
  return -1;
}

// euler_compute_energy.hpp: E = p/(gamma-1) + 0.5 rho |v|^2
template <int NV>
double energyFromPrim(double gm1Inv, double rho, const double* vel, double p) {
  double k = 0;
  for (int m = 0; m < NV; ++m) k += vel[m] * vel[m];
  return p * gm1Inv + 0.5 * rho * k;
}

// impl/euler_rankine_hugoniot.hpp:55-148 (pre-shock gas at rest)
void postShockFromPreshockAtRest(double post[4], const double pre[4], double angle, double mach, double gamma) {
  const double rho0 = pre[0], p0 = pre[3], v0 = 0.0;
  const double m2 = mach * mach;
  const double rho1 = rho0 * (gamma + 1.0) * m2 / (2.0 + (gamma - 1.0) * m2);
  const double p1 = p0 * (1.0 + 2.0 * gamma / (gamma + 1.0) * (m2 - 1.0));
  const double a0 = std::sqrt(gamma * p0 / rho0);
  const double a1 = std::sqrt(gamma * p1 / rho1);
  const double num = (1.0 + 0.5 * (gamma - 1.0) * m2);
  const double den = gamma * m2 - 0.5 * (gamma - 1.0);
  const double mrel = std::sqrt(num / den);
  const double v1 = mrel * a1 - mach * a0 + v0;
  post[0] = rho1;
  post[1] = -v1 * std::cos(angle);
  post[2] = -v1 * std::sin(angle);
  post[3] = p1;
}

}  // namespace

// =============================================================================================== device state
struct DeviceRowSet {
  DevBuf<int32_t> graph, rowIds;
  DevBuf<int32_t> jBase, jLen;
  DevBuf<uint8_t> jSlot;
  int32_t n = 0;
  std::vector<int32_t> hostRowIds;
  dev::RowSet view(int ncols) const { return dev::RowSet{graph.p, rowIds.p, n, ncols}; }
  dev::JacLayout jac(int nslotCols) const { return dev::JacLayout{jBase.p, jLen.p, jSlot.p, nslotCols}; }
};

struct DeviceState {
  cudaStream_t stream = nullptr;
  DeviceRowSet inner, nearBd;
  bool innerViaLattice = false;
  bool windowLattice = false;   // slab window (Mesh::makeWindow): inner rows through the structured kernels on the local lattice
  // ghosts
  DevBuf<double> ghost[6];
  DevBuf<dev::GhostRecipe> recipes;
  DevBuf<double2> nbXY;
  DevBuf<double> factors;
  dev::GhostTables tables{};
  int nsides = 0, hS = 0;
  bool haveGhosts = false;
  bool jacTablesReady = false;        // near-boundary rows
  bool innerJacTablesReady = false;   // inner rows (graph-driven Jacobian kernels, reference-order mode)
  bool innerRowsReady = false;
  // scratch owned by the problem (host-pointer entry points, applyJacobian)
  DevBuf<double> dU, dV, dJ, dB, dR;
  DevBuf<int32_t> dRowptr, dColidx;
  DevBuf<double> stAux, stK[4];   // stepper work vectors
  PinnedBuf<double> hostU, hostGhost[4];   // host-functor boundary conditions (slow path)
  // J*B per cell (k_spmm_cells_rowmajor): per-cell CSR base / row length, lattice visiting order, transposed operands
  DevBuf<int32_t> dCellBase, dCellLen, dCellOrder;
  DevBuf<double> dBt, dRt;
  bool spmmReady = false, spmmFewReady = false;
  int32_t spmmMaxLen = 0;
  // lattice Jacobian kernel: per-cell CSR base and block-slot tables (indexed by gid)
  DevBuf<int32_t> latBase;
  DevBuf<uint4> latSlots;
  bool latJacReady = false;
  DevBuf<double> src;          // per-sample-row source table (diffusion-reaction ProblemA, ADR ProblemA)
  bool srcReady = false;
  // host-pointer pipeline (velocityHost on large lattices)
  static constexpr int kMaxChunks = 128;
  cudaStream_t sH2D = nullptr, sD2H = nullptr;
  cudaEvent_t evIn[kMaxChunks] = {}, evOut[kMaxChunks] = {};
  void ensurePipeline() {
    if (sH2D) return;
    PDA_CUDA(cudaStreamCreateWithFlags(&sH2D, cudaStreamNonBlocking));
    PDA_CUDA(cudaStreamCreateWithFlags(&sD2H, cudaStreamNonBlocking));
    for (int i = 0; i < kMaxChunks; ++i) {
      PDA_CUDA(cudaEventCreateWithFlags(&evIn[i], cudaEventDisableTiming));
      PDA_CUDA(cudaEventCreateWithFlags(&evOut[i], cudaEventDisableTiming));
    }
  }
  ~DeviceState() {
    if (stream) cudaStreamDestroy(stream);
    if (sH2D) cudaStreamDestroy(sH2D);
    if (sD2H) cudaStreamDestroy(sD2H);
    for (int i = 0; i < kMaxChunks; ++i) { if (evIn[i]) cudaEventDestroy(evIn[i]); if (evOut[i]) cudaEventDestroy(evOut[i]); }
    for (int i = 0; i < 2; ++i) {
      if (peer.opened[i] && peer.remote[i] && !(i == 1 && peer.remote[1] == peer.remote[0])) cudaIpcCloseMemHandle(peer.remote[i]);
      if (peer.sPush[i]) cudaStreamDestroy(peer.sPush[i]);
      if (peer.evPushed[i]) cudaEventDestroy(peer.evPushed[i]);
    }
    if (peer.evReady) cudaEventDestroy(peer.evReady);
    if (peer.base) cudaFree(peer.base);
  }
  // ---- slab peer mode: library-owned halo buffers the ring neighbours push into over NVLink (engine.cu, bottom)
  struct PeerHalo {
    unsigned char* base = nullptr;        // [2 parities][lo,hi][h planes] doubles | flags | epoch table
    size_t haloDoubles = 0;               // h * planeDofs
    size_t flagsOff = 0, tableOff = 0, bytes = 0;
    unsigned char* remote[2] = {nullptr, nullptr};   // base of the lower / upper neighbour's buffer (peer-mapped)
    bool opened[2] = {false, false};      // remote[i] came from cudaIpcOpenMemHandle
    bool connected = false;
    cudaStream_t sPush[2] = {nullptr, nullptr};
    cudaEvent_t evReady = nullptr, evPushed[2] = {nullptr, nullptr};
    uint32_t epoch = 0;
    static constexpr size_t kFlagStride = 128, kTableEntries = 65536;
    double* halo(unsigned char* b, int par, int side) const { return reinterpret_cast<double*>(b) + (size_t)(par * 2 + side) * haloDoubles; }
    uint32_t* flag(unsigned char* b, int par, int side) const { return reinterpret_cast<uint32_t*>(b + flagsOff + (size_t)(par * 2 + side) * kFlagStride); }
  } peer;
  dev::GhostView ghostView(int ndpc) const {
    dev::GhostView gv;
    for (int s = 0; s < 6; ++s) gv.g[s] = ghost[s].p;
    gv.stride = ndpc * hS;
    return gv;
  }
};

// =============================================================================================== construction
Problem::Problem(Mesh* mesh, int family, int problemId, int recon, int icFlag, int nparams,
                 const char* const* names, const double* values, int device)
    : mesh_(mesh), family_(family), probId_(problemId), recon_(recon), icFlag_(icFlag), device_(device) {
  if (!mesh) throw Error(kInvalid, "create_problem: null mesh");
  if (recon < 0 || recon > 2) throw Error(kInvalid, "create_problem: invalid reconstruction enum");
  S_ = 3 + 2 * recon;
  switch (family) {
    case F_EULER1D: dim_ = 1; ndpc_ = 3; if (problemId < 0 || problemId > 3) throw Error(kInvalid, "Euler1d: invalid problem enum"); break;
    case F_EULER2D: dim_ = 2; ndpc_ = 4; if (problemId < 0 || problemId > 8) throw Error(kInvalid, "Euler2d: invalid problem enum"); break;
    case F_EULER3D: dim_ = 3; ndpc_ = 5; if (problemId < 0 || problemId > 1) throw Error(kInvalid, "Euler3d: invalid problem enum"); break;
    case F_SWE2D: dim_ = 2; ndpc_ = 3; if (problemId < 0 || problemId > 1) throw Error(kInvalid, "Swe2d: invalid problem enum"); break;
    case F_DIFFREAC2D:
      dim_ = 2; ndpc_ = (problemId == 1) ? 2 : 1;
      if (problemId < 0 || problemId > 1) throw Error(kInvalid, "DiffusionReaction2d: invalid problem enum");
      S_ = 3;
      break;
    case F_ADVDIFF2D: dim_ = 2; ndpc_ = 2; if (problemId < 0 || problemId > 1) throw Error(kInvalid, "AdvectionDiffusion2d: invalid problem enum"); break;
    case F_ADVDIFFREAC2D: dim_ = 2; ndpc_ = 1; if (problemId != 0) throw Error(kInvalid, "advection-diffusion-reaction2d: invalid problem enum"); break;
    case F_ADVECTION1D: dim_ = 1; ndpc_ = 1; if (problemId != 0) throw Error(kInvalid, "advection: invalid problem enum"); break;
    case F_DIFFREAC1D: dim_ = 1; ndpc_ = 1; S_ = 3; if (problemId != 0) throw Error(kInvalid, "1D diffusion-reaction: invalid problem enum"); break;
    default: throw Error(kUnsupported, "create_problem: problem family not available in this engine");
  }
  if (mesh->dim != dim_) throw Error(kInvalid, "create_problem: mesh dimensionality does not match the problem");
  // the reference does not check this (schemes_info.hpp:94-102 is unused) and reads garbage columns; we refuse
  if (mesh->stencil < S_) throw Error(kInvalid, "create_problem: mesh stencil size too small for the reconstruction scheme");
  if (family == F_DIFFREAC2D && mesh->stencil != 3)
    throw Error(kInvalid, "DiffusionReaction2d currently, only supports 3-pt stencil");   // diffusion_reaction_2d_prob_class.hpp:307-309
  if (family == F_DIFFREAC1D && mesh->stencil != 3)
    throw Error(kInvalid, "DiffusionReaction1d currently only supports 3-pt stencil");    // diffusion_reaction_1d_prob_class.hpp:103-105
  if ((int64_t)mesh->nStencil * ndpc_ > INT32_MAX) throw Error(kTooLarge, "create_problem: dof count exceeds int32");

  // ---- parameters
  if (family == F_EULER2D) {
    physParams_ = {1.4};
    icParams_ = {9, 0.1, 10., 1., 0.4, 1.5, 0.0, 0.0, 1.5, 0.029};
    const bool valid = (problemId == E2_RIEMANN) ? (icFlag == 1 || icFlag == 2) : (problemId == E2_NEUMANN ? true : icFlag == 1);
    if (!valid) throw Error(kInvalid, "Euler2d: invalid icFlag for the given problem enum");
    if (nparams > 0 && problemId != E2_RIEMANN && problemId != E2_NORMAL_SHOCK && problemId != E2_CROSS_SHOCK)
      throw Error(kInvalid, "Euler2d: custom parametrization only valid for Euler2d::{Riemann, NormalShock}");
    for (int i = 0; i < nparams; ++i) {
      const std::string s = names[i];
      if (s == "gamma") { physParams_[0] = values[i]; continue; }
      const int idx = euler2dIcIndex(problemId, icFlag, s);
      if (idx < 0) throw Error(kInvalid, "Euler2d: invalid parameter name '" + s + "'");
      icParams_[idx] = values[i];
    }
    gamma_ = physParams_[0];
    if (problemId == E2_RIEMANN && icFlag == 2 && icParams_[icRiem2TRP] <= icParams_[icRiem2BLP])
      throw Error(kInvalid, "INVALID: riemannTopRightPressure <= riemannBotLeftPressure");
  } else if (family == F_SWE2D) {
    physParams_ = {9.8, -3.0};
    icParams_ = {1.0 / 8, 1, 1, 1.0 / 10, -2, -2, 1.0 / 8, 2, 2};
    if (icFlag < 1 || icFlag > 2) throw Error(kInvalid, "2D swe: invalid icFlag");
    for (int i = 0; i < nparams; ++i) {
      const std::string s = names[i];
      if (s == "gravity") { physParams_[0] = values[i]; continue; }
      if (s == "coriolis") { physParams_[1] = values[i]; continue; }
      const int idx = sweIcIndex(icFlag, s);
      if (idx < 0) {
        // names valid for the OTHER icFlag are accepted by the reference's name check and then written out of
        // range; here they are rejected like any unknown name
        throw Error(kInvalid, "2D swe: one or more params in user-provided map is invalid ('" + s + "')");
      }
      icParams_[idx] = values[i];
    }
    customBcs_ = (problemId == 1);
  } else if (family == F_DIFFREAC2D && problemId == 1) {
    for (int i = 0; i < nparams; ++i) {
      const std::string s = names[i];
      if (s == "Du") gs_[0] = values[i];
      else if (s == "Dv") gs_[1] = values[i];
      else if (s == "F") gs_[2] = values[i];
      else if (s == "k") gs_[3] = values[i];
      else throw Error(kInvalid, "GrayScott: invalid parameter name '" + s + "'");
    }
  } else if (family == F_DIFFREAC2D || family == F_DIFFREAC1D) {
    // ProblemA: D = k = 0.01 (diffusion_reaction1d.hpp:118-121, diffusion_reaction2d.hpp:144-150)
    physParams_ = {0.01, 0.01};
    for (int i = 0; i < nparams; ++i) {
      const std::string s = names[i];
      if (s == "diffusion") physParams_[0] = values[i];
      else if (s == "reaction") physParams_[1] = values[i];
      else throw Error(kInvalid, "diffusion-reaction ProblemA: invalid parameter name '" + s + "'");
    }
  } else if (family == F_ADVDIFF2D) {
    // impl/advection_diffusion_2d_parametrization_helpers.hpp:69-82
    physParams_ = {0.00001};
    icParams_ = {0.5, 0.15, 0.0, -0.2};
    if (icFlag != 1) throw Error(kInvalid, "AdvectionDiffusion2d: invalid icFlag");
    for (int i = 0; i < nparams; ++i) {
      const std::string s = names[i];
      if (s == "diffusion") physParams_[0] = values[i];
      else if (s == "pulseMagnitude") icParams_[0] = values[i];
      else if (s == "pulseSpread") icParams_[1] = values[i];
      else if (s == "pulseX") icParams_[2] = values[i];
      else if (s == "pulseY") icParams_[3] = values[i];
      else throw Error(kInvalid, "AdvectionDiffusion2d: invalid parameter name '" + s + "'");
    }
  } else if (family == F_ADVDIFFREAC2D) {
    // advection_diffusion_reaction2d.hpp:124-128: ux, uy, diffusion, sigma; default source f = 1
    physParams_ = {0.5 * std::cos(M_PI / 3), 0.5 * std::sin(M_PI / 3), 0.001, 1.0};
    for (int i = 0; i < nparams; ++i) {
      const std::string s = names[i];
      if (s == "ux") physParams_[0] = values[i];
      else if (s == "uy") physParams_[1] = values[i];
      else if (s == "diffusion") physParams_[2] = values[i];
      else if (s == "sigma") physParams_[3] = values[i];
      else throw Error(kInvalid, "advection-diffusion-reaction2d: invalid parameter name '" + s + "'");
    }
  } else if (family == F_ADVECTION1D) {
    // advection1d.hpp:88-92,140-152: velocity (default 1), ic 1..4
    physParams_ = {1.0};
    if (icFlag < 1 || icFlag > 4) throw Error(kInvalid, "advection1d: invalid ic");
    for (int i = 0; i < nparams; ++i) {
      const std::string s = names[i];
      if (s == "velocity") physParams_[0] = values[i];
      else throw Error(kInvalid, "advection1d: invalid parameter name '" + s + "'");
    }
  } else if (nparams > 0) {
    throw Error(kInvalid, "create_problem: this problem takes no user parameters");
  }
}

Problem::~Problem() = default;

double Problem::queryParameter(const std::string& name) const {
  if (family_ == F_EULER1D || family_ == F_EULER3D) {
    if (name == "gamma") return gamma_;
  } else if (family_ == F_EULER2D) {
    if (name == "gamma") return physParams_[0];
    const int idx = euler2dIcIndex(probId_, icFlag_, name);
    if (idx >= 0) return icParams_[idx];
  } else if (family_ == F_SWE2D) {
    if (name == "gravity") return physParams_[0];
    if (name == "coriolis") return physParams_[1];
    const int idx = sweIcIndex(icFlag_, name);
    if (idx >= 0) return icParams_[idx];
  } else if (family_ == F_DIFFREAC1D || (family_ == F_DIFFREAC2D && probId_ == 0)) {
    if (name == "diffusion") return physParams_[0];
    if (name == "reaction") return physParams_[1];
  } else if (family_ == F_ADVDIFF2D) {
    static const char* icn[4] = {"pulseMagnitude", "pulseSpread", "pulseX", "pulseY"};
    if (name == "diffusion") return physParams_[0];
    for (int i = 0; i < 4; ++i) if (name == icn[i]) return icParams_[i];
  } else if (family_ == F_ADVDIFFREAC2D) {
    static const char* pn[4] = {"ux", "uy", "diffusion", "sigma"};
    for (int i = 0; i < 4; ++i) if (name == pn[i]) return physParams_[i];
  } else if (family_ == F_ADVECTION1D) {
    if (name == "velocity") return physParams_[0];
  } else if (family_ == F_DIFFREAC2D) {
    if (name == "Du") return gs_[0];
    if (name == "Dv") return gs_[1];
    if (name == "F") return gs_[2];
    if (name == "k") return gs_[3];
  }
  throw Error(kInvalid, "queryParameter: unknown parameter '" + name + "'");
}

void Problem::setOption(const std::string& name, const std::string& value) {
  if (name == "jacobian_order") {
    // "fast" (default): face-sharing kernels, well-conditioned reconstruction gradients (closer to the exact Jacobian
    // than the reference, not within 1e-12 of it).  "reference": kernels_reforder.cu -- the reference's formulas,
    // operation order and accumulation order, bit for bit.
    if (value == "fast") refOrderJac_ = false;
    else if (value == "reference") refOrderJac_ = true;
    else throw Error(kInvalid, "set_option: jacobian_order must be \"fast\" or \"reference\"");
    return;
  }
  if (name == "velocity_order") {
    // "fast" (default): face-sharing kernels, division-free leaf arithmetic (values within 1e-12 / 1e-10 of the
    // reference at the reference's mesh sizes).  "reference": one thread per row, the reference's operation order,
    // every operation individually rounded: identical values at ANY mesh size (where hInv * ulp(flux) exceeds 1e-10
    // -- Mach-10 flows on fine meshes -- no other evaluation order can stay inside the tolerance, the reference's
    // own compiled with FMA contraction included).
    if (value == "fast") refOrderVel_ = false;
    else if (value == "reference") refOrderVel_ = true;
    else throw Error(kInvalid, "set_option: velocity_order must be \"fast\" or \"reference\"");
    return;
  }
  if (name == "order") {   // both at once
    setOption("jacobian_order", value);
    setOption("velocity_order", value);
    return;
  }
  throw Error(kInvalid, "set_option: unknown option '" + name + "'");
}

std::string Problem::getOption(const std::string& name) const {
  if (name == "jacobian_order") return refOrderJac_ ? "reference" : "fast";
  if (name == "velocity_order") return refOrderVel_ ? "reference" : "fast";
  throw Error(kInvalid, "get_option: unknown option '" + name + "'");
}

void Problem::setBc(int side, int kind, const double* values) {
  const bool ok = (family_ == F_SWE2D && probId_ == 1) || family_ == F_ADVDIFF2D ||
                  (family_ == F_EULER2D && (probId_ == E2_RIEMANN || probId_ == E2_NORMAL_SHOCK));
  if (!ok)
    throw Error(kInvalid, "custom BCs only valid for Swe2d::CustomBCs, Euler2d::{Riemann, NormalShock} and AdvectionDiffusion2d");
  if (side < 0 || side > 3) throw Error(kInvalid, "set_bc: invalid side");
  if (kind < 0 || kind > 2) throw Error(kInvalid, "set_bc: invalid kind");
  bc_[side] = BcRule{};
  bc_[side].kind = kind;
  if (kind == BC_DIRICHLET) {
    if (!values) throw Error(kInvalid, "set_bc: Dirichlet needs ndpc values");
    for (int d = 0; d < ndpc_; ++d) bc_[side].values[d] = values[d];
  }
  customBcs_ = true;
  if (dev_) { dev_->haveGhosts = false; dev_->jacTablesReady = false; }
}

void Problem::setBcCallback(int side, BcGhostFn ghost, BcFactorFn factors, void* user) {
  const bool ok = (family_ == F_SWE2D && probId_ == 1) || family_ == F_ADVDIFF2D ||
                  (family_ == F_EULER2D && (probId_ == E2_RIEMANN || probId_ == E2_NORMAL_SHOCK));
  if (!ok)
    throw Error(kInvalid, "custom BCs only valid for Swe2d::CustomBCs, Euler2d::{Riemann, NormalShock} and AdvectionDiffusion2d");
  if (side < 0 || side > 3) throw Error(kInvalid, "set_bc_callback: invalid side");
  if (!ghost) throw Error(kInvalid, "set_bc_callback: null ghost functor");
  bc_[side] = BcRule{};
  bc_[side].kind = BC_CALLBACK;
  bc_[side].ghostFn = ghost;
  bc_[side].factorFn = factors;
  bc_[side].user = user;
  customBcs_ = true;
  if (dev_) { dev_->haveGhosts = false; dev_->jacTablesReady = false; }
}

void Problem::setBcPointer(int side, void* user) {
  // setBCPointer(rloc, ptr) hands ptr to the functor of that side (custom_bc_holder.hpp:89-103 -> the user's
  // setInternalPtr); here the functor's state IS the `user` argument of its callbacks
  if (side < 0 || side > 3) throw Error(kInvalid, "set_bc_pointer: invalid side");
  if (bc_[side].kind != BC_CALLBACK) throw Error(kInvalid, "set_bc_pointer: no host functor installed on this side");
  bc_[side].user = user;   // the functors run at every evaluation: nothing cached depends on it
}

// =============================================================================================== initial condition
void Problem::initialCondition(double* U) const {
  Mesh& m = *mesh_;
  const int32_t n = m.nStencil;
  // coordinates as the reference sees them; for a big lattice they are produced per axis, never per cell
  std::vector<double> cx, cy, cz;
  const bool lat = m.lattice && !m.haveCoords;
  if (lat) {
    cx.resize(m.n[0]); cy.resize(m.n[1]); cz.resize(m.n[2]);
    for (int32_t i = 0; i < m.n[0]; ++i) cx[i] = m.latticeCoord(0, i);
    for (int32_t j = 0; j < m.n[1]; ++j) cy[j] = m.latticeCoord(1, j);
    for (int32_t k = 0; k < m.n[2]; ++k) cz[k] = m.latticeCoord(2, k);
  }
  const int32_t nx = m.n[0], ny = m.n[1];
  auto X = [&](int32_t i) { return lat ? cx[i % nx] : m.x[i]; };
  auto Y = [&](int32_t i) { return lat ? cy[(i / nx) % ny] : m.y[i]; };
  auto Z = [&](int32_t i) { return lat ? cz[i / (nx * ny)] : m.z[i]; };
  const double gm1Inv = 1.0 / (gamma_ - 1.0);

  if (family_ == F_EULER1D) {   // impl/euler_1d_initial_condition.hpp:55-183
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
      double prim[3] = {0, 0, 0};
      const double x = X(i);
      switch (probId_) {
        case 0: prim[0] = 1.0 + 0.2 * std::sin(M_PI * x); prim[1] = 1.0; prim[2] = 1.0; break;
        case 1:
          if (x <= 0.0) { prim[0] = 1.0; prim[1] = 0.0; prim[2] = 1.0; }
          if (x > 0.0) { prim[0] = 0.125; prim[1] = 0.0; prim[2] = 0.1; }
          break;
        case 2:
          if (x <= 0.0) { prim[0] = 0.445; prim[1] = 0.698; prim[2] = 3.528; }
          else if (x > 0.0) { prim[0] = 0.5; prim[1] = 0.0; prim[2] = 0.571; }
          break;
        case 3:
          if (x <= -4.0) { prim[0] = 27.0 / 7.0; prim[1] = 2.629369; prim[2] = 31.0 / 3.0; }
          else { prim[0] = 1.0 + (1.0 / 5.0) * std::sin(5.0 * x); prim[1] = 0.0; prim[2] = 1.0; }
          break;
      }
      U[3 * i] = prim[0];
      U[3 * i + 1] = prim[0] * prim[1];
      U[3 * i + 2] = energyFromPrim<1>(gm1Inv, prim[0], &prim[1], prim[2]);
    }
    return;
  }

  if (family_ == F_EULER2D) {   // impl/euler_2d_initial_condition.hpp
    const double gamma = gamma_;
    const double gm1 = gamma - 1.0;
    auto put = [&](int32_t i, const double prim[4]) {
      U[4 * i] = prim[0];
      U[4 * i + 1] = prim[0] * prim[1];
      U[4 * i + 2] = prim[0] * prim[2];
      U[4 * i + 3] = energyFromPrim<2>(gm1Inv, prim[0], &prim[1], prim[3]);
    };
    switch (probId_) {
      case E2_PERIODIC:
#pragma omp parallel for schedule(static)
        for (int32_t i = 0; i < n; ++i) {
          const double prim[4] = {1.0 + (1.0 / 5.0) * std::sin(M_PI * (X(i) + Y(i))), 1.0, 1.0, 1.0};
          put(i, prim);
        }
        return;
      case E2_KH:
#pragma omp parallel for schedule(static)
        for (int32_t i = 0; i < n; ++i) {
          const double freq = 4., mag = 0.025;
          const double pert = mag * std::cos(2. * 3.14159265 / 10. * freq * X(i));
          double* s = U + 4 * (int64_t)i;
          if (Y(i) > -2 + pert && Y(i) < 2 + pert) { s[0] = 2.; s[1] = s[0] * 0.5; }
          else { s[0] = 1.; s[1] = -s[0] * 0.5; }
          s[2] = 0.;
          s[3] = 2.5 / gm1 + 0.5 / s[0] * (s[1] * s[1] + s[2] * s[2]);
        }
        return;
      case E2_SEDOV_FULL:
      case E2_SEDOV_SYM: {
        const bool sym = probId_ == E2_SEDOV_SYM;
        const double sRad = (sym ? 3. : 2.0) * std::min(m.d[0], m.d[1]);
#pragma omp parallel for schedule(static)
        for (int32_t i = 0; i < n; ++i) {
          const double r = std::sqrt(X(i) * X(i) + Y(i) * Y(i));
          double prim[4] = {1.0, 0.0, 0.0, 0.0};
          if (r <= sRad) prim[3] = sym ? gm1 * 0.851072 / (M_PI * sRad * sRad) : gm1 / (M_PI * sRad * sRad);
          else prim[3] = sym ? 2.5e-5 : 5.e-5;
          put(i, prim);
        }
        return;
      }
      case E2_RIEMANN: {
        if (icFlag_ == 1) {
          const double trp = icParams_[icRiem1TRP];
          const double x0 = 0.5, y0 = 0.5;
          // the reference keeps `prim` across iterations (stale value if a centre sits on x == x0 with y < y0,
          // SURVEY C-13); the loop is therefore serial here too
          double prim[4] = {0, 0, 0, 0};
          for (int32_t i = 0; i < n; ++i) {
            const double x = X(i), y = Y(i);
            if (x >= x0 && y >= y0) { prim[0] = 0.5313; prim[1] = 0; prim[2] = 0; prim[3] = trp; }
            else if (x < x0 && y >= y0) { prim[0] = 1; prim[1] = 0.7276; prim[2] = 0; prim[3] = 1; }
            else if (x < x0 && y < y0) { prim[0] = 0.8; prim[1] = 0; prim[2] = 0; prim[3] = 1; }
            else if (x > x0 && y < y0) { prim[0] = 1; prim[1] = 0; prim[2] = 0.7276; prim[3] = 1; }
            put(i, prim);
          }
        } else {
          const double p1 = icParams_[icRiem2TRP], u1 = icParams_[icRiem2TRU], v1 = icParams_[icRiem2TRV];
          const double rho1 = icParams_[icRiem2TRD], p3 = icParams_[icRiem2BLP];
          const double x0 = 0.8, y0 = 0.8;
          const double eps = (gamma - 1.0) / (gamma + 1.0);
          const double fac13 = p1 / p3;
          const double fac43 = (1.0 / (2.0 * (1.0 + 2.0 * eps))) *
                               (eps * (fac13 + 1.0) + std::sqrt(std::pow(eps * (fac13 + 1.0), 2) + 4.0 * (1.0 + 2.0 * eps) * fac13));
          const double p2 = fac43 * p3, p4 = p2;
          const double v2 = v1, u4 = u1;
          const double rho2 = rho1 * (p2 / p1 + eps) / (1 + eps * p2 / rho1);
          const double rho4 = rho2;
          const double psi21 = (p2 - p1) * (rho2 - rho1) / (rho2 * rho1);
          const double u2 = std::sqrt(psi21) + u1;
          const double psi41 = (p4 - p1) * (rho4 - rho1) / (rho4 * rho1);
          const double v4 = std::sqrt(psi41) + v1;
          const double u3 = u2, v3 = v4;
          const double rho3 = rho2 * (p3 - p2) / ((p3 - p2) - psi41 * rho2);
          double prim[4] = {0, 0, 0, 0};
          for (int32_t i = 0; i < n; ++i) {
            const double x = X(i), y = Y(i);
            if (x >= x0 && y >= y0) { prim[0] = rho1; prim[1] = u1; prim[2] = v1; prim[3] = p1; }
            else if (x < x0 && y >= y0) { prim[0] = rho2; prim[1] = u2; prim[2] = v2; prim[3] = p2; }
            else if (x < x0 && y < y0) { prim[0] = rho3; prim[1] = u3; prim[2] = v3; prim[3] = p3; }
            else if (x > x0 && y < y0) { prim[0] = rho4; prim[1] = u4; prim[2] = v4; prim[3] = p4; }
            put(i, prim);
          }
        }
        return;
      }
      case E2_NORMAL_SHOCK:
      case E2_DMR: {
        const bool dmr = probId_ == E2_DMR;
        const double mach = dmr ? 10.0 : icParams_[icNormalShockMach];
        const double angle = dmr ? M_PI / 6.0 : 0.0;
        const double slope = std::tan(angle);
        const double pre[4] = {gamma, 0.0, 0.0, 1.0};
        double post[4];
        postShockFromPreshockAtRest(post, pre, dmr ? -angle : angle, mach, gamma);
#pragma omp parallel for schedule(static)
        for (int32_t i = 0; i < n; ++i) {
          const double xShock = dmr ? (1.0 / 6.0 + slope * Y(i)) : 1.0 / 6.0;
          put(i, (X(i) < xShock) ? post : pre);
        }
        return;
      }
      case E2_CROSS_SHOCK: {
        const double prim[4] = {icParams_[icCrossDensity], icParams_[icCrossInletX], 0., 1.};
#pragma omp parallel for schedule(static)
        for (int32_t i = 0; i < n; ++i) put(i, prim);
        return;
      }
      case E2_NEUMANN:   // euler_2d_prob_class.hpp:438-440: returns an uninitialised vector; zeros here
        std::memset(U, 0, sizeof(double) * (size_t)n * 4);
        return;
    }
  }

  if (family_ == F_EULER3D) {   // impl/euler_3d_initial_condition.hpp:57-143
    if (probId_ == 0) {
#pragma omp parallel for schedule(static)
      for (int32_t i = 0; i < n; ++i) {
        const double rho = 1.0 + 0.2 * std::sin(M_PI * (X(i) + Y(i) + Z(i)));
        const double vel[3] = {1.0, 1.0, 1.0};
        double* s = U + 5 * (int64_t)i;
        s[0] = rho; s[1] = rho; s[2] = rho; s[3] = rho;
        s[4] = energyFromPrim<3>(gm1Inv, rho, vel, 1.0);
      }
    } else {
      const double sRad = 3. * std::min(m.d[0], std::min(m.d[1], m.d[2]));
      const double gm1 = gamma_ - 1.;
#pragma omp parallel for schedule(static)
      for (int32_t i = 0; i < n; ++i) {
        const double r = std::sqrt(X(i) * X(i) + Y(i) * Y(i) + Z(i) * Z(i));
        const double p = (r <= sRad) ? (3. * gm1 * 0.851072) / (4. * M_PI * sRad * sRad * sRad) : 2.5e-5;
        const double vel[3] = {0, 0, 0};
        double* s = U + 5 * (int64_t)i;
        s[0] = 1.0; s[1] = 0.0; s[2] = 0.0; s[3] = 0.0;
        s[4] = energyFromPrim<3>(gm1Inv, 1.0, vel, p);
      }
    }
    return;
  }

  if (family_ == F_SWE2D) {   // impl/swe_2d_initial_condition.hpp:55-107
    const double* ic = icParams_.data();
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
      double hgt;
      if (icFlag_ == 1) {
        const double dx1 = X(i) - ic[1], dy1 = Y(i) - ic[2];
        const double r = std::sqrt(dx1 * dx1 + dy1 * dy1);
        hgt = 1.0 + ic[0] * std::exp(-(r * r));
      } else {
        const double dx1 = X(i) - ic[4], dy1 = Y(i) - ic[5];
        const double r1 = std::sqrt(dx1 * dx1 + dy1 * dy1);
        const double dx2 = X(i) - ic[7], dy2 = Y(i) - ic[8];
        const double r2 = std::sqrt(dx2 * dx2 + dy2 * dy2);
        hgt = 1.0 + ic[3] * std::exp(-(r1 * r1)) + ic[6] * std::exp(-(r2 * r2));
      }
      U[3 * (int64_t)i] = hgt; U[3 * (int64_t)i + 1] = 0.0; U[3 * (int64_t)i + 2] = 0.0;
    }
    return;
  }

  if (family_ == F_DIFFREAC1D || family_ == F_ADVDIFFREAC2D || (family_ == F_DIFFREAC2D && probId_ == 0)) {
    // zero state: diffusion_reaction_1d_prob_class.hpp:110-118, diffusion_reaction_2d_prob_class.hpp:144-149,
    // advection_diffusion_reaction_2d_initial_condition.hpp:57-61
    std::memset(U, 0, sizeof(double) * (size_t)n * ndpc_);
    return;
  }
  if (family_ == F_ADVDIFF2D) {   // impl/advection_diffusion_2d_initial_condition.hpp:54-78
    const double mag = icParams_[0], spread = icParams_[1], x0 = icParams_[2], y0 = icParams_[3];
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
      const double dx = X(i) - x0, dy = Y(i) - y0;
      const double dxSq = dx * dx, dySq = dy * dy;
      const double v = mag * std::exp(-(dxSq + dySq) / spread);
      U[2 * (int64_t)i] = v; U[2 * (int64_t)i + 1] = v;
    }
    return;
  }
  if (family_ == F_ADVECTION1D) {   // impl/advection_1d_prob_class.hpp:111-156
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
      const double x = X(i);
      double r = 0.0;
      if (icFlag_ == 1) r = std::sin(M_PI * x);
      else if (icFlag_ == 2) {
        const double dx1Sq = (x - 1.2) * (x - 1.2), dx2Sq = (x - 2.5) * (x - 2.5);
        r = 0.8 * std::exp(-200.0 * dx1Sq / 16.0) + std::exp(-100.0 * dx2Sq / 36.0);
      } else if (icFlag_ == 3) {
        const double delta = 0.5 * 0.5;
        const double dx1Sq = (x - 2.) * (x - 2.), dx2Sq = (x - 3.0) * (x - 3.0);
        r = std::exp(-dx1Sq / delta) + 0.5 * std::exp(-dx2Sq / delta);
      } else {
        r = std::tanh(8.0 * (x - 1.0)) - std::tanh(8.0 * (x - 3.0));
      }
      U[i] = r;
    }
    return;
  }

  if (family_ == F_DIFFREAC2D) {   // diffusion_reaction_2d_prob_class.hpp:141-178 (Gray-Scott)
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
      const bool in = std::abs(X(i)) < 0.1 && std::abs(Y(i)) < 0.1;
      U[2 * (int64_t)i] = in ? 0.5 : 1.0;
      U[2 * (int64_t)i + 1] = in ? 0.25 : 0.0;
    }
    return;
  }
  throw Error(kUnsupported, "initialCondition: family not supported");
}

// =============================================================================================== CSR pattern
// Same triplet rule as EigenApp::initializeJacobian (euler_2d_prob_class.hpp:223-237,315-387; swe, euler1d/3d
// alike; diffusion_reaction_2d_prob_class.hpp:180-224): inner rows = self + the (S-1)*dim scheme neighbours,
// near-boundary rows = self + existing first-layer neighbours; setFromTriplets sorts columns and merges duplicates.
void Problem::buildPattern() {
  if (havePattern_) return;
  Mesh& m = *mesh_;
  m.ensureGraph();
  m.ensureRows();
  const int nc = m.ncols();
  const int nnbInner = (S_ - 1) * dim_;
  const int nnbFirst = 2 * dim_;
  const bool allFirst = (family_ == F_DIFFREAC2D || family_ == F_DIFFREAC1D);
  const int32_t ns = m.nSample;
  slotCols_ = nnbInner + 1;
  std::vector<uint8_t> isNb(ns, 0);
  for (int32_t r : m.rowsNearBd) isNb[r] = 1;

  cellBase_.assign(ns, 0);
  cellLen_.assign(ns, 0);
  slots_.assign((size_t)ns * slotCols_, 0xFF);
  std::vector<int32_t> nblk(ns);
  bool merged = false;
  // pass 1: blocks per cell
#pragma omp parallel for schedule(static) reduction(|| : merged)
  for (int32_t r = 0; r < ns; ++r) {
    const int32_t* row = &m.graph[(size_t)r * nc];
    const bool first = allFirst || isNb[r];
    const int ncand = first ? nnbFirst : nnbInner;
    int32_t ids[20];
    int cnt = 0;
    ids[cnt++] = row[0];
    for (int c = 1; c <= ncand; ++c) if (row[c] >= 0) ids[cnt++] = row[c];
    std::sort(ids, ids + cnt);
    const int u = (int)(std::unique(ids, ids + cnt) - ids);
    if (u != cnt) merged = true;
    nblk[r] = u;
  }
  mergedNeighbors_ = merged;
  int64_t nnz = 0;
  for (int32_t r = 0; r < ns; ++r) nnz += (int64_t)nblk[r] * ndpc_ * ndpc_;
  if (nnz > INT32_MAX)
    throw Error(kTooLarge, "jacobian: nnz = " + std::to_string(nnz) + " does not fit the reference's int32 index type");
  rowptr_.assign((size_t)ns * ndpc_ + 1, 0);
  {
    int64_t acc = 0;
    for (int32_t r = 0; r < ns; ++r) {
      cellBase_[r] = (int32_t)acc;
      cellLen_[r] = nblk[r] * ndpc_;
      for (int k = 0; k < ndpc_; ++k) { rowptr_[(size_t)r * ndpc_ + k] = (int32_t)acc; acc += cellLen_[r]; }
    }
    rowptr_[(size_t)ns * ndpc_] = (int32_t)acc;
  }
  colidx_.assign((size_t)nnz, 0);
#pragma omp parallel for schedule(static)
  for (int32_t r = 0; r < ns; ++r) {
    const int32_t* row = &m.graph[(size_t)r * nc];
    const bool first = allFirst || isNb[r];
    const int ncand = first ? nnbFirst : nnbInner;
    int32_t ids[20];
    int cnt = 0;
    ids[cnt++] = row[0];
    for (int c = 1; c <= ncand; ++c) if (row[c] >= 0) ids[cnt++] = row[c];
    std::sort(ids, ids + cnt);
    cnt = (int)(std::unique(ids, ids + cnt) - ids);
    for (int k = 0; k < ndpc_; ++k) {
      int32_t* out = &colidx_[(size_t)cellBase_[r] + (size_t)k * cellLen_[r]];
      for (int b = 0; b < cnt; ++b)
        for (int j = 0; j < ndpc_; ++j) out[b * ndpc_ + j] = ids[b] * ndpc_ + j;
    }
    uint8_t* sl = &slots_[(size_t)r * slotCols_];
    for (int c = 0; c <= ncand; ++c) {
      if (row[c] < 0) continue;
      sl[c] = (uint8_t)(std::lower_bound(ids, ids + cnt, row[c]) - ids);
    }
  }
  havePattern_ = true;
}

int64_t Problem::jacobianNnz() {
  buildPattern();
  return (int64_t)colidx_.size();
}

void Problem::jacobianPattern(int32_t* rowptr, int32_t* colidx) {
  buildPattern();
  if (rowptr) std::memcpy(rowptr, rowptr_.data(), rowptr_.size() * sizeof(int32_t));
  if (colidx) std::memcpy(colidx, colidx_.data(), colidx_.size() * sizeof(int32_t));
}

// =============================================================================================== device set-up
void Problem::ensureDevice() {
  if (dev_) {   // every caller may allocate right after: the current device must be this problem's
    PDA_CUDA(cudaSetDevice(device_));
    return;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error(kNoDevice, "no usable CUDA device: the B200 engine has no CPU fallback");
  }
  if (device_ < 0 || device_ >= ndev) throw Error(kInvalid, "create_problem: invalid CUDA device index");
  PDA_CUDA(cudaSetDevice(device_));
  auto ds = std::make_unique<DeviceState>();
  PDA_CUDA(cudaStreamCreateWithFlags(&ds->stream, cudaStreamNonBlocking));

  Mesh& m = *mesh_;
  const int nc = m.ncols();
  // ---- near-boundary rows (compact graph)
  std::vector<int32_t> nb;
  m.nearBdRows(nb);
  ds->nearBd.n = (int32_t)nb.size();
  if (!nb.empty()) {
    std::vector<int32_t> g((size_t)nb.size() * nc);
    for (size_t r = 0; r < nb.size(); ++r) {
      if (m.haveGraph) std::memcpy(&g[r * nc], &m.graph[(size_t)nb[r] * nc], sizeof(int32_t) * nc);
      else m.latticeRow(nb[r], &g[r * nc]);
    }
    ds->nearBd.graph.upload(g);
    ds->nearBd.rowIds.upload(nb);
  }
  ds->nearBd.hostRowIds = nb;
  // ---- inner rows: structured kernels on lattices, compact graph otherwise
  ds->innerViaLattice = m.lattice && (latticeKernelAvailable(family_, dim_, S_) ||
                                      (family_ == F_DIFFREAC2D && probId_ == 1 && m.fullyPeriodic));
  ds->hS = (S_ - 1) / 2;
  ds->nsides = (dim_ == 1) ? 3 : 2 * dim_;
  dev_ = std::move(ds);
  // slab window: velocity (2D / 3D) and the 2D Jacobian of the inner rows run the structured kernels on the window's
  // local lattice; everything else (near-boundary rows, other families, reference-order mode) stays graph-driven
  dev_->windowLattice = m.window && dim_ >= 2 && m.n[0] >= 2 * m.halo() && m.n[1] >= 2 * m.halo() &&
                        (dim_ == 2 || m.n[2] >= 2 * m.halo()) &&
                        (family_ == F_EULER2D || family_ == F_EULER3D || family_ == F_SWE2D || family_ == F_ADVDIFF2D);
  if (!dev_->innerViaLattice) ensureInnerRows();
}

// compact graph of the inner rows (graph-driven kernels): always for non-lattice meshes, on demand for lattices
// (Jacobian assembly of a lattice still goes through the graph kernels)
void Problem::ensureInnerRows() {
  DeviceState& ds = *dev_;
  if (ds.innerRowsReady) return;
  Mesh& m = *mesh_;
  const int nc = m.ncols();
  m.ensureGraph();
  m.ensureRows();
  std::vector<int32_t> rows;
  if (family_ == F_DIFFREAC2D || family_ == F_DIFFREAC1D) { rows.resize(m.nSample); for (int32_t r = 0; r < m.nSample; ++r) rows[r] = r; }
  else rows = m.rowsInner;
  ds.inner.n = (int32_t)rows.size();
  if (!rows.empty()) {
    std::vector<int32_t> g((size_t)rows.size() * nc);
    for (size_t r = 0; r < rows.size(); ++r) std::memcpy(&g[r * nc], &m.graph[(size_t)rows[r] * nc], sizeof(int32_t) * nc);
    ds.inner.graph.upload(g);
    ds.inner.rowIds.upload(rows);
  }
  ds.inner.hostRowIds = rows;
  ds.innerRowsReady = true;
}

// ghost recipes: which cell (and which sign / constant) every needed ghost layer copies -- per-problem rules of the
// reference's ghost fillers, evaluated ONCE on the host because they only depend on the static graph.
void Problem::buildGhostRecipes() {
  DeviceState& ds = *dev_;
  if (ds.haveGhosts) return;
  Mesh& m = *mesh_;
  const int nc = m.ncols();
  const int h = ds.hS;
  const int nsides = ds.nsides;
  const std::vector<int32_t>& nb = ds.nearBd.hostRowIds;
  const int32_t nNb = (int32_t)nb.size();
  dev::GhostTables& T = ds.tables;
  std::memset(&T, 0, sizeof T);
  int nModes = 0;
  auto addMode = [&](std::initializer_list<double> mul, std::initializer_list<double> add) {
    if (nModes >= dev::kMaxGhostModes) throw Error(kUnsupported, "ghost modes exhausted");
    int d = 0; for (double v : mul) T.mul[nModes][d++] = v;
    d = 0; for (double v : add) T.add[nModes][d++] = v;
    return nModes++;
  };
  const int mCopy = addMode({1, 1, 1, 1, 1}, {0, 0, 0, 0, 0});
  const int mNeg1 = addMode({1, -1, 1, 1, 1}, {0, 0, 0, 0, 0});
  const int mNeg2 = addMode({1, 1, -1, 1, 1}, {0, 0, 0, 0, 0});
  const int mNeg3 = addMode({1, 1, 1, -1, 1}, {0, 0, 0, 0, 0});
  const int mNegAll = addMode({-1, -1, -1, -1, -1}, {0, 0, 0, 0, 0});
  const int mZero = addMode({0, 0, 0, 0, 0}, {0, 0, 0, 0, 0});
  int mDirich[6] = {-1, -1, -1, -1, -1, -1};

  // which filler applies (0 = none)
  enum { NONE, NAIVE, PROPER, CUSTOM } style = NONE;
  // per-side mode for mirror fills
  int sideMode[6] = {mCopy, mCopy, mCopy, mCopy, mCopy, mCopy};
  bool dmr = false, cross = false;
  int mPost = -1, mPre = -1, mCrossL = -1, mCrossB0a = -1, mCrossB0b = -1, mCrossB1a = -1, mCrossB1b = -1;

  if (customBcs_) {
    style = CUSTOM;
    for (int s = 0; s < 4; ++s) {
      if (bc_[s].kind < 0) throw Error(kInvalid, "custom BCs: pda_problem_set_bc must be called for all four sides before evaluating");
      if (bc_[s].kind == BC_DIRICHLET) {
        const double* v = bc_[s].values;
        mDirich[s] = addMode({0, 0, 0, 0, 0}, {v[0], v[1], v[2], v[3], v[4]});
      }
    }
  } else if (family_ == F_EULER1D) {
    style = (probId_ == 0) ? NONE : NAIVE;                       // euler_1d_prob_class.hpp:315-335
  } else if (family_ == F_EULER2D) {
    switch (probId_) {
      case E2_SEDOV_FULL: case E2_RIEMANN: case E2_NEUMANN: style = PROPER; break;   // Ghost2dNeumannFiller
      case E2_SEDOV_SYM: style = NAIVE; sideMode[0] = mNeg1; sideMode[3] = mNeg2; break;
      case E2_NORMAL_SHOCK: style = NAIVE; sideMode[1] = mNeg2; sideMode[3] = mNeg2; break;
      case E2_DMR: {
        style = NAIVE; dmr = true;
        // euler_2d_ghost_filler_double_mach_reflection.hpp:265-291
        const double gamma = gamma_, gm1Inv = 1.0 / (gamma - 1.0);
        const double pre[4] = {gamma, 0.0, 0.0, 1.0};
        double post[4];
        postShockFromPreshockAtRest(post, pre, -(M_PI / 6.), 10.0, gamma);
        const double preS[4] = {pre[0], pre[0] * pre[1], pre[0] * pre[2], energyFromPrim<2>(gm1Inv, pre[0], &pre[1], pre[3])};
        const double postS[4] = {post[0], post[0] * post[1], post[0] * post[2], energyFromPrim<2>(gm1Inv, post[0], &post[1], post[3])};
        mPost = addMode({0, 0, 0, 0, 0}, {postS[0], postS[1], postS[2], postS[3], 0});
        mPre = addMode({0, 0, 0, 0, 0}, {preS[0], preS[1], preS[2], preS[3], 0});
        T.dmrWedge = 1.0 / 6.0;
        T.dmrSpeed = 10.0 / std::cos(M_PI / 6.);
        T.dmrSlope = std::tan(M_PI / 6.);
        T.dy = m.d[1];
        T.dmrModePost = mPost; T.dmrModePre = mPre;
        break;
      }
      case E2_CROSS_SHOCK: {
        style = NAIVE; cross = true;
        const double rho = icParams_[icCrossDensity], uin = icParams_[icCrossInletX], vb = icParams_[icCrossBottomY];
        const double vel[2] = {uin, 0.0};
        const double E = energyFromPrim<2>(1.0 / (gamma_ - 1.0), rho, vel, 1.0);
        mCrossL = addMode({0, 0, 0, 0, 0}, {rho, rho * uin, rho * 0.0, E, 0});
        mCrossB0a = addMode({1, 0, 0, 1, 0}, {0, rho * uin, 0., 0, 0});
        mCrossB0b = addMode({1, 0, 0, 1, 0}, {0, rho * uin, rho * vb, 0, 0});
        mCrossB1a = addMode({0, 0, 0, 1, 0}, {rho, rho * uin, 0., 0, 0});
        mCrossB1b = addMode({0, 0, 0, 1, 0}, {rho, rho * uin, rho * vb, 0, 0});
        break;
      }
      default: style = NONE;   // PeriodicSmooth, KelvinHelmholtz: no ghosts (euler_2d_prob_class.hpp:562-564)
    }
  } else if (family_ == F_EULER3D) {
    if (probId_ == 1) { style = NAIVE; sideMode[0] = mNeg1; sideMode[3] = mNeg2; sideMode[4] = mNeg3; }
  } else if (family_ == F_SWE2D) {
    style = PROPER; sideMode[0] = mNeg1; sideMode[2] = mNeg1; sideMode[1] = mNeg2; sideMode[3] = mNeg2;
  } else if (family_ == F_ADVDIFF2D) {
    // BurgersOutflow (advection_diffusion_2d_ghost_filler_outflow.hpp:102-235): left/back homogeneous Dirichlet,
    // right/front layer-aware mirror copy; BurgersPeriodic has no ghosts
    if (probId_ == 1) { style = PROPER; sideMode[0] = mZero; sideMode[3] = mZero; }
  } else if (family_ == F_ADVDIFFREAC2D) {
    // advection_diffusion_reaction_2d_ghost_filler_problemA.hpp:92-147: layer 0 <- -self, layer k <- -(opposite k-1)
    style = NAIVE;
    for (int sd = 0; sd < 4; ++sd) sideMode[sd] = mNegAll;
  }

  if (nNb > 0 && style == NONE && !m.fullyPeriodic && family_ != F_DIFFREAC2D && family_ != F_DIFFREAC1D)
    throw Error(kInvalid, "this problem requires a fully periodic mesh (no ghost filler in the reference)");

  std::vector<dev::GhostRecipe> rec((size_t)nNb * nsides * h, dev::GhostRecipe{-1, 0, 0});
  std::vector<double2> xy(nNb);
  std::vector<double> fac((size_t)nNb * dim_ * ndpc_, 1.0);
  std::vector<int32_t> rowBuf(nc);
  static const int oppSide[6] = {2, 3, 0, 1, 5, 4};
  for (int32_t r = 0; r < nNb; ++r) {
    const int32_t* row;
    if (m.haveGraph) row = &m.graph[(size_t)nb[r] * nc];
    else { m.latticeRow(nb[r], rowBuf.data()); row = rowBuf.data(); }
    const int32_t self = row[0];
    double cxv, cyv;
    if (m.haveCoords) { cxv = m.x[self]; cyv = m.y[self]; }
    else { cxv = m.latticeCoord(0, self % m.n[0]); cyv = m.latticeCoord(1, (self / m.n[0]) % m.n[1]); }
    xy[r] = make_double2(cxv, cyv);
    auto nbr = [&](int side, int L) { return row[graphCol(dim_, side, L)]; };
    auto hasBd = [&](int side) {   // hasBd{Left,..}{1,2,3}d of the MESH stencil (mesh_ccu.hpp:162-296)
      for (int L = 0; L < m.halo(); ++L) if (nbr(side, L) == -1) return true;
      return false;
    };
    for (int si = 0; si < nsides; ++si) {
      if (dim_ == 1 && si == 1) continue;
      const int side = si;
      const int opp = oppSide[side];
      for (int L = 0; L < h; ++L) {
        if (nbr(side, L) != -1) continue;
        dev::GhostRecipe& g = rec[((size_t)r * nsides + si) * h + L];
        int32_t src = self;
        if (style == NAIVE) {
          src = (L == 0) ? self : nbr(opp, L - 1);
        } else if (style == PROPER) {
          if (L == 0) src = self;
          else if (L == 1) src = (nbr(side, 0) == -1) ? nbr(opp, 0) : nbr(side, 0);
          else {
            const int32_t s1 = nbr(side, 1), s0 = nbr(side, 0);
            if (s1 != -1 && s0 != -1) src = s1;
            else if (s1 == -1 && s0 != -1) src = self;
            else if (s1 == -1 && s0 == -1) src = nbr(opp, 1);
            else src = self;
          }
        } else if (style == CUSTOM) {
          // device-expressible rules: Dirichlet = constant state, homogeneous Neumann = the cell's own state,
          // reflective = layer-aware mirror with the normal momentum negated
          if (bc_[side].kind == BC_REFLECTIVE) {
            if (L == 0) src = self;
            else if (L == 1) src = (nbr(side, 0) == -1) ? nbr(opp, 0) : nbr(side, 0);
            else {
              const int32_t s1 = nbr(side, 1), s0 = nbr(side, 0);
              src = (s1 != -1 && s0 != -1) ? s1 : ((s1 == -1 && s0 != -1) ? self : nbr(opp, 1));
            }
          }
        }
        if (src < 0) src = self;   // degenerate (mesh narrower than the stencil): the reference reads out of bounds
        if (style == CUSTOM && bc_[side].kind == BC_CALLBACK) continue;   // filled on the host (runHostBcCallbacks)
        g.src = src;
        g.kind = 0;
        g.mode = (int16_t)sideMode[side];
        if (style == CUSTOM) {
          const int k = bc_[side].kind;
          g.mode = (int16_t)((k == BC_DIRICHLET) ? mDirich[side]
                             : (k == BC_REFLECTIVE) ? ((side == 0 || side == 2) ? mNeg1 : mNeg2) : mCopy);
        }
        if (dmr) {
          if (side == 1) { g.kind = 1; g.mode = (int16_t)mPre; }
          if (side == 3) g.mode = (int16_t)((cxv < 1.0 / 6.0) ? mCopy : mNeg2);
        }
        if (cross) {
          if (side == 0) g.mode = (int16_t)mCrossL;
          if (side == 3) g.mode = (int16_t)((L == 0) ? (cxv < 0.5 ? mCrossB0a : mCrossB0b) : (cxv < 0.5 ? mCrossB1a : mCrossB1b));
        }
      }
    }
    // ---- first-order Jacobian factors per axis (fillJacFactorsForCellBd: euler_2d_prob_class.hpp:1115-1224,
    //      swe_2d_prob_class.hpp:839-855, euler_3d_prob_class.hpp:1017-1044, euler_1d_prob_class.hpp:608-616)
    for (int ax = 0; ax < dim_; ++ax) {
      double* f = &fac[((size_t)r * dim_ + ax) * ndpc_];
      for (int d = 0; d < ndpc_; ++d) f[d] = 1.0;
      const int sm = minusSide(ax), sp = plusSide(ax);
      if (style == CUSTOM) {
        // fillJacFactorsCustomBCs (custom_bcs_functions.hpp:60-103): minus side first, plus side overwrites
        for (int s : {sm, sp}) {
          if (!hasBd(s)) continue;
          const int k = bc_[s].kind;
          if (k == BC_CALLBACK) {
            if (bc_[s].factorFn) bc_[s].factorFn(bc_[s].user, row, cxv, cyv, ndpc_, f);
            continue;
          }
          for (int d = 0; d < ndpc_; ++d) f[d] = (k == BC_DIRICHLET) ? 0.0 : 1.0;
          if (k == BC_REFLECTIVE) f[1 + ax] = -1.0;
        }
      } else if (family_ == F_SWE2D) {
        f[1 + ax] = -1.0;
      } else if (family_ == F_EULER2D) {
        if (probId_ == E2_SEDOV_SYM) { if (hasBd(sm)) f[1 + ax] = -1.0; }
        else if (probId_ == E2_NORMAL_SHOCK) { if (ax == 1) f[2] = -1.0; }
        else if (probId_ == E2_DMR) {
          if (ax == 1) {
            if (hasBd(3) && cxv < 1.0 / 6.0) { /* neumann */ }
            else if (hasBd(3) && cxv >= 1.0 / 6.0) f[2] = -1.0;
            else for (int d = 0; d < ndpc_; ++d) f[d] = 0.0;
          }
        } else if (probId_ == E2_CROSS_SHOCK) {
          if (ax == 0) { if (hasBd(0)) for (int d = 0; d < ndpc_; ++d) f[d] = 0.0; }
          else { if (hasBd(3)) { f[1] = 0.0; f[2] = 0.0; } }
        }
      } else if (family_ == F_EULER3D) {
        if (probId_ == 1 && hasBd(sm)) f[1 + ax] = -1.0;
      } else if (family_ == F_ADVDIFF2D) {
        // advection_diffusion_2d_prob_class.hpp:1119-1155: minus side (left/back) Dirichlet -> 0, else Neumann -> 1
        if (hasBd(sm)) for (int d = 0; d < ndpc_; ++d) f[d] = 0.0;
      } else if (family_ == F_ADVDIFFREAC2D) {
        f[0] = -1.0;   // advection_diffusion_reaction_2d_prob_class.hpp:634-635
      }
    }
  }
  for (int s = 0; s < 6; ++s) {
    ds.ghost[s].alloc((size_t)std::max<int32_t>(nNb, 1) * h * ndpc_);
    PDA_CUDA(cudaMemset(ds.ghost[s].p, 0, ds.ghost[s].n * sizeof(double)));
  }
  ds.recipes.upload(rec);
  ds.nbXY.upload(xy);
  ds.factors.upload(fac);
  ds.haveGhosts = true;
}

// =============================================================================================== evaluation
namespace {

template <class F>
void dispatchScheme(int S, F&& f) {
  switch (S) {
    case 3: f(std::integral_constant<int, 3>{}); break;
    case 5: f(std::integral_constant<int, 5>{}); break;
    case 7: f(std::integral_constant<int, 7>{}); break;
    default: throw Error(kInvalid, "invalid scheme stencil");
  }
}

inline int gridFor(int64_t n, int block) { return (int)((n + block - 1) / block); }
// operands up to this many columns take the matrix-free inner-row kernel (PDA_FUSED_APPLY_MAX_COLS overrides: tuning)
// PDA_JAC_FO_MARCH=0 keeps the tile kernel for first-order schemes (A/B measurements)
inline bool jacFoMarchEnabled() {
  static const bool on = [] { const char* e = std::getenv("PDA_JAC_FO_MARCH"); return !(e && e[0] == '0'); }();
  return on;
}

// PDA_JAC_WENO_MARCH=0 keeps the tile kernel for WENO3/5 on the 1-3 dof systems (A/B measurements)
inline bool jacWenoMarchEnabled() {
  static const bool on = [] { const char* e = std::getenv("PDA_JAC_WENO_MARCH"); return !(e && e[0] == '0'); }();
  return on;
}

inline int fusedApplyMaxCols() {
  static const int v = [] { const char* e = std::getenv("PDA_FUSED_APPLY_MAX_COLS"); return e ? std::atoi(e) : 12; }();
  return v;
}

}  // namespace

void Problem::evaluateDev(const double* dU, double t, double* dV, double* dJ, void* streamV) {
  ensureDevice();
  PDA_CUDA(cudaSetDevice(device_));
  DeviceState& ds = *dev_;
  cudaStream_t st = (cudaStream_t)streamV;   // NULL = the legacy default stream (CUDA convention)
  Mesh& m = *mesh_;
  const int nc = m.ncols();
  dev::Deltas dl{{m.dInv[0], m.dInv[1], m.dInv[2]}};

  // reference-order mode (pda_problem_set_option "jacobian_order" = "reference", kernels_reforder.cu): every row through
  // the one-thread-per-row kernels that keep the reference's formulas, operation order and accumulation order.  The
  // diffusion-reaction families have no reconstruction gradients: their kernels already follow the reference's order.
  const bool refOrder = dJ && refOrderJac_ && family_ != F_DIFFREAC2D && family_ != F_DIFFREAC1D;
  const bool refVel = !dJ && refOrderVel_ && family_ != F_DIFFREAC2D && family_ != F_DIFFREAC1D;
  if (refVel) ensureInnerRows();
  // inner rows of a 2D full lattice: face-sharing lattice kernel (kernels_jaclattice.cuh); everything else: staged
  // graph-driven kernel
  const bool jacLattice = dJ && !refOrder && m.lattice && dim_ == 2 && ds.innerViaLattice && family_ != F_DIFFREAC2D;
  if (dJ) {
    buildPattern();
    const bool gsLat = family_ == F_DIFFREAC2D && probId_ == 1 && ds.innerViaLattice && m.fullyPeriodic;
    if (((jacLattice || gsLat) && !mergedNeighbors_) && !ds.latJacReady) {
      ds.latBase.upload(cellBase_);
      {   // slot table padded to 16 bytes per cell: one vector load per cell in the kernel
        std::vector<uint4> padded(cellBase_.size(), make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
        if (slotCols_ > 16) throw Error(kUnsupported, "lattice Jacobian: slot table wider than 16 columns");
        for (size_t r = 0; r < cellBase_.size(); ++r)
          std::memcpy(&padded[r], &slots_[r * slotCols_], (size_t)slotCols_);
        ds.latSlots.upload(padded);
      }
      ds.latJacReady = true;
    }
    auto fill = [&](DeviceRowSet& rs) {
      std::vector<int32_t> base(rs.n), len(rs.n);
      std::vector<uint8_t> sl((size_t)rs.n * slotCols_);
      for (int32_t r = 0; r < rs.n; ++r) {
        const int32_t row = rs.hostRowIds[r];
        base[r] = cellBase_[row]; len[r] = cellLen_[row];
        std::memcpy(&sl[(size_t)r * slotCols_], &slots_[(size_t)row * slotCols_], slotCols_);
      }
      rs.jBase.upload(base); rs.jLen.upload(len); rs.jSlot.upload(sl);
    };
    if (!ds.jacTablesReady) { fill(ds.nearBd); ds.jacTablesReady = true; }
    const bool innerTables = refOrder || !((jacLattice || gsLat) && !mergedNeighbors_);
    if (innerTables && !ds.innerJacTablesReady) { ensureInnerRows(); fill(ds.inner); ds.innerJacTablesReady = true; }
    // the staged inner-row kernel writes every value once; only rows assembled by read-modify-write need zeros
    if (refOrder) {
      PDA_CUDA(cudaMemsetAsync(dJ, 0, colidx_.size() * sizeof(double), st));   // the reference zeroes J, then accumulates
    } else if (gsLat && !mergedNeighbors_) {
      // the Gray-Scott lattice kernel writes every value once
    } else if (family_ == F_DIFFREAC2D || family_ == F_DIFFREAC1D || mergedNeighbors_) {
      PDA_CUDA(cudaMemsetAsync(dJ, 0, colidx_.size() * sizeof(double), st));
    } else if (ds.nearBd.n > 0) {
      dev::k_zero_cell_chunks<<<gridFor((int64_t)ds.nearBd.n * 32, 256), 256, 0, st>>>(ds.nearBd.jBase.p, ds.nearBd.jLen.p,
                                                                                     ds.nearBd.n, ndpc_, dJ);
      ++launches_;
    }
  }

  // ---- diffusion-reaction families in reference-order mode: the -fmad=false twins of the kernels below
  if ((family_ == F_DIFFREAC2D || family_ == F_DIFFREAC1D) && ((dJ && refOrderJac_) || (!dJ && refOrderVel_))) {
    ensureInnerRows();   // all rows of these families
    if (dJ && !ds.innerJacTablesReady) {
      std::vector<int32_t> base(ds.inner.n), len(ds.inner.n);
      std::vector<uint8_t> sl((size_t)ds.inner.n * slotCols_);
      for (int32_t r = 0; r < ds.inner.n; ++r) {
        const int32_t row = ds.inner.hostRowIds[r];
        base[r] = cellBase_[row]; len[r] = cellLen_[row];
        std::memcpy(&sl[(size_t)r * slotCols_], &slots_[(size_t)row * slotCols_], slotCols_);
      }
      ds.inner.jBase.upload(base); ds.inner.jLen.upload(len); ds.inner.jSlot.upload(sl);
      ds.innerJacTablesReady = true;
    }
    if (dJ) PDA_CUDA(cudaMemsetAsync(dJ, 0, colidx_.size() * sizeof(double), st));
    if (family_ == F_DIFFREAC2D && probId_ == 1) {
      if (!m.fullyPeriodic) throw Error(kInvalid, "GrayScott requires a periodic mesh");
      dev::launchRefOrderGrayScott(gs_, m.dInv[0], m.dInv[1], ds.inner.graph.p, ds.inner.rowIds.p, ds.inner.n, nc, dU, dV, dJ,
                                   ds.inner.jBase.p, ds.inner.jLen.p, ds.inner.jSlot.p, slotCols_, st);
    } else {
      ensureSource();
      dev::launchRefOrderDiffReac(dim_, physParams_[0], physParams_[1], m.dInv[0], m.dInv[1], ds.inner.graph.p,
                                  ds.inner.rowIds.p, ds.inner.n, nc, dU, ds.src.p, dV, dJ, ds.inner.jBase.p, ds.inner.jLen.p,
                                  ds.inner.jSlot.p, slotCols_, st);
    }
    ++launches_;
    PDA_CUDA(cudaGetLastError());
    return;
  }
  // ---- Gray-Scott: one fused kernel over all rows
  if (family_ == F_DIFFREAC2D && probId_ == 1) {
    dev::GrayScottParams gp{gs_[0], gs_[1], gs_[2], gs_[3], m.dInv[0] * m.dInv[0], m.dInv[1] * m.dInv[1]};
    if (!m.fullyPeriodic) throw Error(kInvalid, "GrayScott requires a periodic mesh");
    if (ds.innerViaLattice && dJ && !mergedNeighbors_) {
      const int64_t ncell = (int64_t)m.n[0] * m.n[1];
      dev::k_gray_scott_lattice_jac<<<(unsigned)((ncell + 127) / 128), 128, 0, st>>>(
          gp, m.n[0], m.n[1], reinterpret_cast<const double2*>(dU), reinterpret_cast<double2*>(dV), dJ, ds.latBase.p, ds.latSlots.p);
    } else if (ds.innerViaLattice && !dJ) {
      dim3 grid((unsigned)((m.n[0] + 127) / 128), (unsigned)m.n[1]);
      dev::k_gray_scott_lattice<<<grid, 128, 0, st>>>(gp, m.n[0], m.n[1], reinterpret_cast<const double2*>(dU),
                                                     reinterpret_cast<double2*>(dV));
    } else {
      ensureInnerRows();
      dev::k_gray_scott_rows<<<gridFor(ds.inner.n, 128), 128, 0, st>>>(gp, ds.inner.view(nc), dU, dV, dJ, ds.inner.jac(slotCols_));
    }
    ++launches_;
    PDA_CUDA(cudaGetLastError());
    return;
  }
  // ---- diffusion-reaction ProblemA (1D / 2D): one fused kernel over all rows, source from the per-row table
  if (family_ == F_DIFFREAC1D || family_ == F_DIFFREAC2D) {
    ensureSource();
    dev::DiffReacParams pr{{physParams_[0] * (m.dInv[0] * m.dInv[0]), physParams_[0] * (m.dInv[1] * m.dInv[1])}, physParams_[1]};
    if (dim_ == 1)
      dev::k_diffreac_rows<1><<<gridFor(ds.inner.n, 128), 128, 0, st>>>(pr, ds.inner.view(nc), dU, ds.src.p, dV, dJ, ds.inner.jac(slotCols_));
    else
      dev::k_diffreac_rows<2><<<gridFor(ds.inner.n, 128), 128, 0, st>>>(pr, ds.inner.view(nc), dU, ds.src.p, dV, dJ, ds.inner.jac(slotCols_));
    ++launches_;
    PDA_CUDA(cudaGetLastError());
    return;
  }

  const int32_t nNb = ds.nearBd.n;
  dev::GhostView gv{};
  if (nNb > 0) {
    buildGhostRecipes();
    gv = ds.ghostView(ndpc_);
    const int64_t tot = (int64_t)nNb * ds.nsides * ds.hS;
    auto launchGhost = [&](auto ndpcTag) {
      constexpr int N = decltype(ndpcTag)::value;
      dev::k_ghost_fill<N><<<gridFor(tot, 128), 128, 0, st>>>(ds.recipes.p, ds.nbXY.p, nNb, ds.nsides, ds.hS, ds.tables, dU, gv, t);
    };
    switch (ndpc_) {
      case 1: launchGhost(std::integral_constant<int, 1>{}); break;
      case 2: launchGhost(std::integral_constant<int, 2>{}); break;
      case 3: launchGhost(std::integral_constant<int, 3>{}); break;
      case 4: launchGhost(std::integral_constant<int, 4>{}); break;
      case 5: launchGhost(std::integral_constant<int, 5>{}); break;
    }
    ++launches_;
    if (customBcs_) {
      bool anyCb = false;
      for (int sd = 0; sd < 4; ++sd) anyCb = anyCb || bc_[sd].kind == BC_CALLBACK;
      if (anyCb) runHostBcCallbacks(dU, st);
    }
  }

  auto run = [&](auto phys) {
    using Phys = decltype(phys);
    dispatchScheme(S_, [&](auto sTag) {
      constexpr int S = decltype(sTag)::value;
      if (refOrder || refVel) {
        dev::RefOrderParams P{};
        P.family = family_; P.ndpc = ndpc_; P.S = S_; P.dim = dim_; P.gamma = gamma_;
        for (int a = 0; a < 3; ++a) P.dInv[a] = m.dInv[a];
        if (family_ == F_SWE2D) { P.gravity = physParams_[0]; P.coriolis = physParams_[1]; }
        if (family_ == F_ADVDIFF2D) P.diffusion = physParams_[0];
        if (family_ == F_ADVDIFFREAC2D) {
          P.adv[0] = physParams_[0]; P.adv[1] = physParams_[1]; P.diffusion = physParams_[2]; P.sigma = physParams_[3];
          P.src = srcUser_ ? ds.src.p : nullptr;
        }
        if (family_ == F_ADVECTION1D) P.adv[0] = physParams_[0];
        if (refVel) {
          if (nNb > 0) {
            dev::launchRefOrderVelocity(P, ds.nearBd.graph.p, ds.nearBd.rowIds.p, nNb, nc, dU, dV, gv.g, gv.stride, true, st);
            ++launches_;
          }
          if (ds.inner.n > 0) {
            dev::launchRefOrderVelocity(P, ds.inner.graph.p, ds.inner.rowIds.p, ds.inner.n, nc, dU, dV, nullptr, 0, false, st);
            ++launches_;
          }
          return;
        }
        if (nNb > 0) {
          if (dV) {
            dev::launchRefOrderVelocity(P, ds.nearBd.graph.p, ds.nearBd.rowIds.p, nNb, nc, dU, dV, gv.g, gv.stride, true, st);
            ++launches_;
          }
          dev::launchRefOrderNearBd(P, ds.nearBd.graph.p, ds.nearBd.rowIds.p, nNb, nc, dU, dJ, ds.nearBd.jBase.p,
                                    ds.nearBd.jLen.p, ds.nearBd.jSlot.p, slotCols_, gv.g, gv.stride, ds.factors.p, st);
          ++launches_;
        }
        if (ds.inner.n > 0) {
          dev::launchRefOrderInner(P, ds.inner.graph.p, ds.inner.rowIds.p, ds.inner.n, nc, dU, dV, dJ, ds.inner.jBase.p,
                                   ds.inner.jLen.p, ds.inner.jSlot.p, slotCols_, st);
          ++launches_;
        }
        return;
      }
      if (nNb > 0) {
        if (dV) {
          dev::k_velocity_rows<Phys, S, true><<<gridFor(nNb, 128), 128, 0, st>>>(phys, ds.nearBd.view(nc), dl, dU, dV, gv);
          ++launches_;
        }
        if (dJ) {
          dev::k_jacobian_nearbd_rows<Phys><<<gridFor(nNb, 128), 128, 0, st>>>(phys, ds.nearBd.view(nc), dl, dU, dJ,
                                                                               ds.nearBd.jac(slotCols_), gv, ds.factors.p);
          ++launches_;
        }
      }
      const int64_t winShift = m.window ? (int64_t)m.winHLo * m.winPlaneCells * ndpc_ : 0;   // V rows = owned cells only
      if (ds.innerViaLattice && !dJ) {
        launchLatticeVelocity<Phys, S>(phys, m, dl, dU, dV, st, 0, m.n[dim_ - 1], 0);
        ++launches_;
      } else if (ds.windowLattice && !dJ) {
        if constexpr (Phys::dim >= 2 && !std::is_same<Phys, dev::LinAdv<2>>::value) {
          launchLatticeVelocity<Phys, S>(phys, m, dl, dU, dV - winShift, st, 0, m.n[dim_ - 1], 0);
          ++launches_;
        }
      } else if (ds.windowLattice && dJ && dim_ == 2 && !mergedNeighbors_ && !skipInnerJacobian_) {
        if constexpr (Phys::dim == 2 && !std::is_same<Phys, dev::LinAdv<2>>::value) {
          using JL = dev::JacLat2d<Phys, S>;
          auto kern = dev::k_jacobian_lattice2d<Phys, S>;
          ensureFuncAttrs(kern, (int)JL::smemBytes);
          if (!ds.latJacReady) {   // tables indexed by LOCAL cell id
            std::vector<int32_t> base((size_t)m.nStencil, 0);
            std::vector<uint4> padded((size_t)m.nStencil, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
            if (slotCols_ > 16) throw Error(kUnsupported, "lattice Jacobian: slot table wider than 16 columns");
            const size_t off = (size_t)m.winHLo * (size_t)m.winPlaneCells;
            for (size_t r = 0; r < cellBase_.size(); ++r) {
              base[off + r] = cellBase_[r];
              std::memcpy(&padded[off + r], &slots_[r * slotCols_], (size_t)slotCols_);
            }
            ds.latBase.upload(base);
            ds.latSlots.upload(padded);
            ds.latJacReady = true;
          }
          dev::LatticeDesc L;
          for (int a = 0; a < 3; ++a) { L.n[a] = m.n[a]; L.per[a] = m.periodic[a] ? 1 : 0; }
          L.planeBegin = 0; L.planeEnd = m.n[1]; L.haloPlanes = 0; L.slab = 0; L.meshHalo = m.halo();
          L.haloLo = L.haloHi = nullptr; L.flagLo = L.flagHi = nullptr; L.epoch = 0;
          const int w0 = L.per[0] ? m.n[0] : m.n[0] - 2 * m.halo(), w1 = m.n[1] - 2 * m.halo();
          if (w0 > 0 && w1 > 0) {
            dev::JacLatTables jt{ds.latBase.p, ds.latSlots.p, ndpc_ * (1 + dim_ * (S - 1))};
            dim3 grid((unsigned)((w0 + JL::T - 1) / JL::T), (unsigned)((w1 + JL::T - 1) / JL::T));
            kern<<<grid, JL::THREADS, JL::smemBytes, st>>>(phys, L, dl, jt, dU, dV ? dV - winShift : nullptr, dJ);
            ++launches_;
          }
        }
      } else if (jacLattice && !mergedNeighbors_) {
        if constexpr (Phys::dim == 2) {
          using JL = dev::JacLat2d<Phys, S>;
          auto kern = dev::k_jacobian_lattice2d<Phys, S>;
          ensureFuncAttrs(kern, (int)JL::smemBytes);
          dev::LatticeDesc L;
          for (int a = 0; a < 3; ++a) { L.n[a] = m.n[a]; L.per[a] = m.periodic[a] ? 1 : 0; }
          L.planeBegin = 0; L.planeEnd = m.n[1]; L.haloPlanes = 0; L.slab = 0; L.meshHalo = m.halo();
          L.haloLo = L.haloHi = nullptr; L.flagLo = L.flagHi = nullptr; L.epoch = 0;
          const int w0 = L.per[0] ? m.n[0] : m.n[0] - 2 * m.halo(), w1 = L.per[1] ? m.n[1] : m.n[1] - 2 * m.halo();
          if (w0 > 0 && w1 > 0 && !skipInnerJacobian_) {
            dev::JacLatTables jt{ds.latBase.p, ds.latSlots.p, ndpc_ * (1 + dim_ * (S - 1))};
            if (S == 3 && jacFoMarchEnabled()) {
              // first-order scheme: y-marching warps, every face once, chunks streamed out per strip row
              using JM = dev::JacMarchFo<Phys>;
              auto kfo = dev::k_jacobian_march2d_fo<Phys, JM::MIN_CTAS>;
              ensureFuncAttrs(kfo, (int)JM::smemBytes);
              const int64_t nStrips = (w0 + JM::W - 1) / JM::W;
              int LY = 64;
              while (LY > 8 && nStrips * ((w1 + LY - 1) / LY) < (int64_t)148 * 16 * 4) LY /= 2;
              const int64_t tasks = nStrips * ((w1 + LY - 1) / LY);
              kfo<<<(unsigned)((tasks + JM::WARPS - 1) / JM::WARPS), 32 * JM::WARPS, JM::smemBytes, st>>>(phys, L, dl, jt, dU, dV, dJ, LY);
            } else if (S > 3 && Phys::ndpc <= 3 && jacWenoMarchEnabled()) {
              // WENO on the small systems: same marching layout with the chain rule in it
              if constexpr (S > 3 && Phys::ndpc <= 3) {
                using JM = dev::JacMarchWeno<Phys, S>;
                auto kw = dev::k_jacobian_march2d_weno<Phys, S>;
                ensureFuncAttrs(kw, (int)JM::smemBytes);
                const int64_t nStrips = (w0 + JM::W - 1) / JM::W;
                int LY = 64;
                while (LY > 8 && nStrips * ((w1 + LY - 1) / LY) < (int64_t)148 * 16 * 4) LY /= 2;
                const int64_t tasks = nStrips * ((w1 + LY - 1) / LY);
                kw<<<(unsigned)((tasks + JM::WARPS - 1) / JM::WARPS), 32 * JM::WARPS, JM::smemBytes, st>>>(phys, L, dl, jt, dU, dV, dJ, LY);
              }
            } else {
              dim3 grid((unsigned)((w0 + JL::T - 1) / JL::T), (unsigned)((w1 + JL::T - 1) / JL::T));
              kern<<<grid, JL::THREADS, JL::smemBytes, st>>>(phys, L, dl, jt, dU, dV, dJ);
            }
            ++launches_;
          }
        }
      } else if (ds.inner.n > 0 && !(skipInnerJacobian_ && dJ)) {
        if (dJ && !mergedNeighbors_) {
          // staged assembly: every value of the inner rows written once, coalesced (no memset needed for them)
          using JS = dev::JacStage<Phys, S>;
          // register cap: MIN_CTAS CTAs per SM (more warps, some spills) or 1 (PDA_JAC_STAGED_OCC=0: A/B measurements)
          static const bool occ = [] { const char* e = std::getenv("PDA_JAC_STAGED_OCC"); return !(e && e[0] == '0'); }();
          auto kern = occ ? dev::k_jacobian_inner_staged<Phys, S, JS::MIN_CTAS> : dev::k_jacobian_inner_staged<Phys, S, 1>;
          ensureFuncAttrs(kern, (int)JS::smemBytes);
          kern<<<gridFor(ds.inner.n, JS::CELLS), JS::THREADS, JS::smemBytes, st>>>(phys, ds.inner.view(nc), dl, dU, dV, dJ,
                                                                               ds.inner.jac(slotCols_));
        } else if (dJ) {
          dev::k_jacobian_inner_rows<Phys, S><<<gridFor(ds.inner.n, 128), 128, 0, st>>>(phys, ds.inner.view(nc), dl, dU, dV, dJ,
                                                                                        ds.inner.jac(slotCols_));
        } else {
          // (a variant with one lane per (cell, axis, face) -- 2*dim lanes per cell, fluxes merged by shuffle -- was measured on
          // cfg 4 and is SLOWER: 41 -> 47 us (WENO3), 54 -> 64 us (WENO5): every face gathers its own stencil, 4(S-1) instead
          // of 2S cells per cell, and the gathers, not the arithmetic, are what this kernel waits for)
          dev::k_velocity_rows<Phys, S, false><<<gridFor(ds.inner.n, 128), 128, 0, st>>>(phys, ds.inner.view(nc), dl, dU, dV, gv);
        }
        ++launches_;
      }
    });
  };
  switch (family_) {
    case F_EULER1D: run(dev::Euler<1>{gamma_}); break;
    case F_EULER2D: run(dev::Euler<2>{gamma_}); break;
    case F_EULER3D: run(dev::Euler<3>{gamma_}); break;
    case F_SWE2D: run(dev::Swe2d{physParams_[0], physParams_[1]}); break;
    case F_ADVDIFF2D:
      run(dev::Burgers2d{{physParams_[0] * (m.dInv[0] * m.dInv[0]), physParams_[0] * (m.dInv[1] * m.dInv[1])}});
      break;
    case F_ADVDIFFREAC2D: {
      const double D = physParams_[2];
      if (srcUser_) ensureSource();
      run(dev::LinAdv<2>{{physParams_[0], physParams_[1]}, {D * (m.dInv[0] * m.dInv[0]), D * (m.dInv[1] * m.dInv[1])},
                         physParams_[3], 1.0, srcUser_ ? ds.src.p : nullptr});
      break;
    }
    case F_ADVECTION1D: run(dev::LinAdv<1>{{physParams_[0]}, {0.0}, 0.0, 0.0, nullptr}); break;
    default: throw Error(kUnsupported, "family not supported on device");
  }
  PDA_CUDA(cudaGetLastError());
}

// Host-functor boundary conditions (PDA_BC_HOST_CALLBACK): the reference's fillGhostsUseCustomFunctors
// (custom_bcs_functions.hpp:107-164) with the functor running on the host over a copy of the state.  Slow path by
// construction (state D2H + ghost rows H2D per evaluation); the device-expressible rules never leave HBM.
void Problem::runHostBcCallbacks(const double* dU, void* streamV) {
  DeviceState& ds = *dev_;
  cudaStream_t st = (cudaStream_t)streamV;
  Mesh& m = *mesh_;
  const int nc = m.ncols();
  const int h = ds.hS;
  const std::vector<int32_t>& nb = ds.nearBd.hostRowIds;
  const int32_t nNb = (int32_t)nb.size();
  const size_t nU = (size_t)nDofStencil();
  ds.hostU.alloc(nU);
  PDA_CUDA(cudaMemcpyAsync(ds.hostU.p, dU, nU * sizeof(double), cudaMemcpyDeviceToHost, st));
  PDA_CUDA(cudaStreamSynchronize(st));
  const size_t rowLen = (size_t)h * ndpc_;
  std::vector<int32_t> rowBuf(nc);
  for (int side = 0; side < 4; ++side) {
    if (bc_[side].kind != BC_CALLBACK) continue;
    ds.hostGhost[side].alloc((size_t)std::max<int32_t>(nNb, 1) * rowLen);
    std::memset(ds.hostGhost[side].p, 0, (size_t)std::max<int32_t>(nNb, 1) * rowLen * sizeof(double));
    const double width = (side == 0 || side == 2) ? m.d[0] : m.d[1];
    for (int32_t r = 0; r < nNb; ++r) {
      const int32_t* row;
      if (m.haveGraph) row = &m.graph[(size_t)nb[r] * nc];
      else { m.latticeRow(nb[r], rowBuf.data()); row = rowBuf.data(); }
      bool hasBd = false;
      for (int L = 0; L < m.halo(); ++L) hasBd = hasBd || row[graphCol(dim_, side, L)] == -1;
      if (!hasBd) continue;
      const int32_t self = row[0];
      double cxv, cyv;
      if (m.haveCoords) { cxv = m.x[self]; cyv = m.y[self]; }
      else { cxv = m.latticeCoord(0, self % m.n[0]); cyv = m.latticeCoord(1, (self / m.n[0]) % m.n[1]); }
      bc_[side].ghostFn(bc_[side].user, r, row, cxv, cyv, ds.hostU.p, ndpc_, width, ds.hostGhost[side].p + (size_t)r * rowLen);
    }
    PDA_CUDA(cudaMemcpyAsync(ds.ghost[side].p, ds.hostGhost[side].p, (size_t)nNb * rowLen * sizeof(double),
                             cudaMemcpyHostToDevice, st));
  }
}

// per-row source table f(x[,y]) of the ProblemA families: the reference's default functors
// (diffusion_reaction1d.hpp:86-97, diffusion_reaction2d.hpp:103-114) tabulated once, or the caller's values
// (pda_problem_set_source: a host functor evaluated by the binding at the current time)
void Problem::ensureSource() {
  DeviceState& ds = *dev_;
  if (ds.srcReady) return;
  Mesh& m = *mesh_;
  if (!srcUser_) {
    m.ensureGraph();
    const int nc = m.ncols();
    srcHost_.resize(m.nSample);
    for (int32_t r = 0; r < m.nSample; ++r) {
      const int32_t c = m.graph[(size_t)r * nc];
      double x, y = 0.0;
      if (m.haveCoords) { x = m.x[c]; if (dim_ > 1) y = m.y[c]; }
      else { x = m.latticeCoord(0, c % m.n[0]); if (dim_ > 1) y = m.latticeCoord(1, (c / m.n[0]) % m.n[1]); }
      if (family_ == F_DIFFREAC1D) srcHost_[r] = std::sin(M_PI * x) * x * x * 4. * std::cos(4. * M_PI * x);
      else if (family_ == F_DIFFREAC2D) srcHost_[r] = std::sin(M_PI * x * (y - 0.2)) * 4. * std::sin(4. * M_PI * y * x);
      else srcHost_[r] = 1.0;
    }
  }
  ds.src.upload(srcHost_);
  ds.srcReady = true;
}

void Problem::setSource(const double* values) {
  const bool ok = family_ == F_DIFFREAC1D || family_ == F_ADVDIFFREAC2D || (family_ == F_DIFFREAC2D && probId_ == 0);
  if (!ok) throw Error(kInvalid, "set_source: this problem has no source term");
  if (!values) throw Error(kInvalid, "set_source: null pointer");
  srcHost_.assign(values, values + mesh_->nSample);
  srcUser_ = true;
  if (dev_) dev_->srcReady = false;
}

// structured velocity of the slowest-axis planes [p0,p1) of a fully periodic lattice (host pipeline, slab interior)
void Problem::evaluatePlanes(const double* dU, double /*t*/, double* dV, void* streamV, int32_t p0, int32_t p1) {
  DeviceState& ds = *dev_;
  cudaStream_t st = (cudaStream_t)streamV;   // NULL = the legacy default stream (CUDA convention)
  Mesh& m = *mesh_;
  dev::Deltas dl{{m.dInv[0], m.dInv[1], m.dInv[2]}};
  auto run = [&](auto phys) {
    using Phys = decltype(phys);
    dispatchScheme(S_, [&](auto sTag) {
      constexpr int S = decltype(sTag)::value;
      launchLatticeVelocity<Phys, S>(phys, m, dl, dU, dV, st, p0, p1, 0);
      ++launches_;
    });
  };
  switch (family_) {
    case F_EULER1D: run(dev::Euler<1>{gamma_}); break;
    case F_EULER2D: run(dev::Euler<2>{gamma_}); break;
    case F_EULER3D: run(dev::Euler<3>{gamma_}); break;
    case F_SWE2D: run(dev::Swe2d{physParams_[0], physParams_[1]}); break;
    case F_ADVDIFF2D:
      run(dev::Burgers2d{{physParams_[0] * (m.dInv[0] * m.dInv[0]), physParams_[0] * (m.dInv[1] * m.dInv[1])}});
      break;
    default: throw Error(kUnsupported, "family not supported on device");
  }
  PDA_CUDA(cudaGetLastError());
}

void Problem::velocityDev(const double* dU, double t, double* dV, void* stream) {
  if (!dU || !dV) throw Error(kInvalid, "velocity: null pointer");
  evaluateDev(dU, t, dV, nullptr, stream);
}

void Problem::velocityAndJacobianDev(const double* dU, double t, double* dV, double* dJ, void* stream) {
  if (!dU || !dJ) throw Error(kInvalid, "jacobian: null pointer");
  ensureDevice();
  if (!dV) {   // jacobian(U,t,J): the reference evaluates into its internal m_rhs (adapter_cpp.hpp:215-221)
    dev_->dV.alloc((size_t)nDofSample());
    dV = dev_->dV.p;
  }
  evaluateDev(dU, t, dV, dJ, stream);
}

void Problem::velocityHost(const double* U, double t, double* V) {
  if (!U || !V) throw Error(kInvalid, "velocity: null pointer");
  ensureDevice();
  PDA_CUDA(cudaSetDevice(device_));
  DeviceState& ds = *dev_;
  Mesh& m = *mesh_;
  const size_t nU = (size_t)nDofStencil(), nV = (size_t)nDofSample();
  ds.dU.alloc(nU); ds.dV.alloc(nV);
  // Host buffers are handed to cudaMemcpyAsync as they are: pinned buffers (cudaHostAlloc / cudaHostRegister /
  // torch pin_memory) stream at PCIe speed and overlap, pageable ones are staged by the driver.
  const int64_t nPlanes = m.n[dim_ - 1];
  // only the families evaluatePlanes dispatches: periodic Gray-Scott / ADR lattices take the single-shot path
  const bool planeFamily = family_ == F_EULER2D || family_ == F_EULER3D || family_ == F_SWE2D || family_ == F_ADVDIFF2D;
  const bool pipelined = !refOrderVel_ && planeFamily && ds.innerViaLattice && ds.nearBd.n == 0 && m.fullyPeriodic && dim_ >= 2 &&
                         (int64_t)m.nSample >= (int64_t)(1 << 22) && nPlanes >= 16;
  if (!pipelined) {
    PDA_CUDA(cudaMemcpyAsync(ds.dU.p, U, nU * sizeof(double), cudaMemcpyHostToDevice, ds.stream));
    evaluateDev(ds.dU.p, t, ds.dV.p, nullptr, ds.stream);
    PDA_CUDA(cudaMemcpyAsync(V, ds.dV.p, nV * sizeof(double), cudaMemcpyDeviceToHost, ds.stream));
    PDA_CUDA(cudaStreamSynchronize(ds.stream));
    return;
  }
  // ---- large periodic lattice: chunks of planes flow H2D -> kernel -> D2H on three streams, so the call costs
  //      max(H2D, D2H) instead of H2D + kernel + D2H.  Chunk c needs the planes of chunk c+1 (upper stencil halo)
  //      and c-1; the wrap-around planes of chunk 0 come from the last chunk, which is therefore uploaded first.
  ds.ensurePipeline();
  const int h = (S_ - 1) / 2;
  // Small chunks keep the pipeline's head (three uploads before the first kernel) and tail (the last two downloads)
  // short: the call then costs the box's full-duplex copy time (tools/pcie_peak.py: 114 ms for 5.37 GB each way)
  // plus a few chunk times; the kernels (16 ms in total) hide behind the copies even at 6 planes per launch.
  // PDA_HOST_CHUNKS overrides the chunk count (tuning only).
  static const int chunkLimit = [] {
    const char* e = std::getenv("PDA_HOST_CHUNKS");
    const int v = e ? std::atoi(e) : 0;
    return (v >= 4 && v <= DeviceState::kMaxChunks) ? v : 86;
  }();
  const int nChunks = (int)std::max<int64_t>(4, std::min<int64_t>(chunkLimit, nPlanes / std::max(4, 2 * h)));
  const size_t planeDofs = nU / (size_t)nPlanes;
  auto c0 = [&](int c) { return (int32_t)((int64_t)nPlanes * c / nChunks); };
  auto h2d = [&](int c) {
    const size_t off = (size_t)c0(c) * planeDofs, cnt = (size_t)(c0(c + 1) - c0(c)) * planeDofs;
    PDA_CUDA(cudaMemcpyAsync(ds.dU.p + off, U + off, cnt * sizeof(double), cudaMemcpyHostToDevice, ds.sH2D));
    PDA_CUDA(cudaEventRecord(ds.evIn[c], ds.sH2D));
  };
  auto compute = [&](int c) {
    for (int w : {c - 1, c, c + 1}) PDA_CUDA(cudaStreamWaitEvent(ds.stream, ds.evIn[(w + nChunks) % nChunks], 0));
    evaluatePlanes(ds.dU.p, t, ds.dV.p, ds.stream, c0(c), c0(c + 1));
    PDA_CUDA(cudaEventRecord(ds.evOut[c], ds.stream));
    PDA_CUDA(cudaStreamWaitEvent(ds.sD2H, ds.evOut[c], 0));
    const size_t off = (size_t)c0(c) * planeDofs, cnt = (size_t)(c0(c + 1) - c0(c)) * planeDofs;
    PDA_CUDA(cudaMemcpyAsync(V + off, ds.dV.p + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, ds.sD2H));
  };
  // the previous call's kernels/copies on ds.stream must be done before dU is overwritten
  PDA_CUDA(cudaEventRecord(ds.evOut[0], ds.stream));
  PDA_CUDA(cudaStreamWaitEvent(ds.sH2D, ds.evOut[0], 0));
  h2d(nChunks - 1);
  h2d(0);
  for (int c = 1; c < nChunks - 1; ++c) {
    h2d(c);
    compute(c - 1);
  }
  compute(nChunks - 2);
  compute(nChunks - 1);
  PDA_CUDA(cudaStreamSynchronize(ds.sD2H));
  PDA_CUDA(cudaStreamSynchronize(ds.stream));
}

void Problem::velocityAndJacobianHost(const double* U, double t, double* V, double* Jvalues) {
  if (!U || !Jvalues) throw Error(kInvalid, "jacobian: null pointer");
  ensureDevice();
  PDA_CUDA(cudaSetDevice(device_));
  buildPattern();
  DeviceState& ds = *dev_;
  const size_t nU = (size_t)nDofStencil(), nV = (size_t)nDofSample(), nnz = colidx_.size();
  ds.dU.alloc(nU); ds.dV.alloc(nV); ds.dJ.alloc(nnz);
  PDA_CUDA(cudaMemcpyAsync(ds.dU.p, U, nU * sizeof(double), cudaMemcpyHostToDevice, ds.stream));
  evaluateDev(ds.dU.p, t, ds.dV.p, ds.dJ.p, ds.stream);
  if (V) PDA_CUDA(cudaMemcpyAsync(V, ds.dV.p, nV * sizeof(double), cudaMemcpyDeviceToHost, ds.stream));
  PDA_CUDA(cudaMemcpyAsync(Jvalues, ds.dJ.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, ds.stream));
  PDA_CUDA(cudaStreamSynchronize(ds.stream));
}

void Problem::applyJacobianDev(const double* dU, const double* dB, int ncols, int layout, double t, double* dR,
                               void* streamV) {
  if (!dU || !dB || !dR) throw Error(kInvalid, "applyJacobian: null pointer");
  if (ncols < 1) throw Error(kInvalid, "applyJacobian: ncols must be >= 1");
  ensureDevice();
  PDA_CUDA(cudaSetDevice(device_));
  DeviceState& ds = *dev_;
  cudaStream_t st = (cudaStream_t)streamV;   // NULL = the legacy default stream (CUDA convention)
  const int32_t nrows = nDofSample();
  const int64_t nJc = nDofStencil();
  Mesh& mm = *mesh_;
  const bool fused3d = !refOrderJac_ && mm.lattice && dim_ == 3 && ds.innerViaLattice && family_ == F_EULER3D &&
                       mm.n[0] >= 2 * mm.halo() && mm.n[1] >= 2 * mm.halo() && mm.n[2] >= 2 * mm.halo();
  const bool fused = !refOrderJac_ && (fused3d ||
                     (mm.lattice && dim_ == 2 && ds.innerViaLattice && ncols <= fusedApplyMaxCols() &&
                      (family_ == F_EULER2D || family_ == F_SWE2D || family_ == F_ADVDIFF2D || family_ == F_ADVDIFFREAC2D)));
  // the assembled Jacobian (pattern, values scratch) is needed unless every row is matrix-free (periodic lattices)
  if (!(fused && ds.nearBd.n == 0)) {
    buildPattern();
    ds.dV.alloc((size_t)nDofSample());
    ds.dJ.alloc(colidx_.size());
    if (!ds.dRowptr.p) { ds.dRowptr.upload(rowptr_); ds.dColidx.upload(colidx_); }
  }
  if (fused) {
    // ---- matrix-free inner rows (kernels_applylattice.cuh); near-boundary rows through the assembled path
    const int64_t ldbRow = (layout == 1) ? ncols : 1, ldbCol = (layout == 1) ? 1 : nJc;
    const int64_t ldrRow = (layout == 1) ? ncols : 1, ldrCol = (layout == 1) ? 1 : nrows;
    // near-boundary rows: assembled (first-order Jacobian with ghost factors) and multiplied per column
    auto nearBdRows = [&]() {
      if (ds.nearBd.n > 0) {
        skipInnerJacobian_ = true;
        try { evaluateDev(dU, t, ds.dV.p, ds.dJ.p, st); } catch (...) { skipInnerJacobian_ = false; throw; }
        skipInnerJacobian_ = false;
        if (!ds.spmmFewReady) { ds.dCellBase.upload(cellBase_); ds.dCellLen.upload(cellLen_); ds.spmmFewReady = true; }
        const unsigned gridNb = (unsigned)gridFor((int64_t)ds.nearBd.n * 32, 256);
        auto nbLaunch = [&](auto nTag) {
          constexpr int NN = decltype(nTag)::value;
          for (int c0 = 0; c0 < ncols; ++c0) {
            dev::k_spmm_cells_fewcols<NN, 1><<<gridNb, 256, 0, st>>>(ds.nearBd.n, ds.nearBd.rowIds.p, ds.dCellBase.p, ds.dCellLen.p,
                                                                     ds.dColidx.p, ds.dJ.p, dB, c0, ldbRow, ldbCol, dR, ldrRow, ldrCol);
            ++launches_;
          }
        };
        switch (ndpc_) {
          case 1: nbLaunch(std::integral_constant<int, 1>{}); break;
          case 2: nbLaunch(std::integral_constant<int, 2>{}); break;
          case 3: nbLaunch(std::integral_constant<int, 3>{}); break;
          case 4: nbLaunch(std::integral_constant<int, 4>{}); break;
          default: nbLaunch(std::integral_constant<int, 5>{}); break;
        }
      }
    };
    // row-major operands with several columns on 3D lattices: transposed around the tiled single-column kernel (the inner
    // kernel then fills whole columns of a scratch result, so the near-boundary rows are written afterwards)
    const bool rowMajor3dTiled = fused3d && layout == 1 && ncols > 1 && applyTiled3dEnabled();
    if (!rowMajor3dTiled) nearBdRows();
    dev::Deltas dl{{mm.dInv[0], mm.dInv[1], mm.dInv[2]}};
    dev::LatticeDesc L;
    for (int a = 0; a < 3; ++a) { L.n[a] = mm.n[a]; L.per[a] = mm.periodic[a] ? 1 : 0; }
    L.planeBegin = 0; L.planeEnd = mm.n[dim_ - 1]; L.haloPlanes = 0; L.slab = 0; L.meshHalo = mm.halo();
    L.haloLo = L.haloHi = nullptr; L.flagLo = L.flagHi = nullptr; L.epoch = 0;
    const int w0 = L.per[0] ? mm.n[0] : mm.n[0] - 2 * mm.halo(), w1 = L.per[1] ? mm.n[1] : mm.n[1] - 2 * mm.halo();
    if (fused3d) {
      const int w2 = L.per[2] ? mm.n[2] : mm.n[2] - 2 * mm.halo();
      if (w0 > 0 && w1 > 0 && w2 > 0) {
        dispatchScheme(S_, [&](auto sTag) {
          constexpr int S = decltype(sTag)::value;
          if (ldbRow == 1 && ldrRow == 1 && applyTiled3dEnabled()) {
            // contiguous operand columns (vector, column-major): the tiled (value, tangent) kernel, one launch per column
            for (int c = 0; c < ncols; ++c) {
              launchApplyTiled3d<S>(gamma_, L, dl, dU, dB + (int64_t)c * ldbCol, dR + (int64_t)c * ldrCol, st);
              ++launches_;
            }
            return;
          }
          if (rowMajor3dTiled) {
            auto transpose = [&](const double* in, int64_t rows, int64_t cols, double* out) {
              const int64_t tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
              dev::k_transpose<<<(unsigned)tiles, 256, 0, st>>>(in, rows, cols, out);
              ++launches_;
            };
            ds.dBt.alloc((size_t)nJc * ncols);
            ds.dRt.alloc((size_t)nrows * ncols);
            const bool skinny = ncols <= 16;
            if (skinny) { dev::k_split_columns<<<gridFor(nJc, 256), 256, 0, st>>>(dB, nJc, ncols, ds.dBt.p); ++launches_; }
            else transpose(dB, nJc, ncols, ds.dBt.p);            // [nJc][ncols] -> [ncols][nJc]
            for (int c = 0; c < ncols; ++c) {
              launchApplyTiled3d<S>(gamma_, L, dl, dU, ds.dBt.p + (int64_t)c * nJc, ds.dRt.p + (int64_t)c * nrows, st);
              ++launches_;
            }
            if (skinny) { dev::k_merge_columns<<<gridFor(nrows, 256), 256, 0, st>>>(ds.dRt.p, nrows, ncols, dR); ++launches_; }
            else transpose(ds.dRt.p, ncols, nrows, dR);          // [ncols][nrows] -> [nrows][ncols]
            return;
          }
          auto pass = [&](auto ncTag, int c0) {
            constexpr int NC = decltype(ncTag)::value;
            using AK = dev::ApplyLat3d<NC>;
            auto kern = dev::k_applyjac_lattice3d<S, NC>;
            ensureFuncAttrs(kern, (int)AK::smemBytes);
            dim3 grid((unsigned)((w0 + AK::T - 1) / AK::T), (unsigned)((w1 + AK::T - 1) / AK::T), (unsigned)((w2 + AK::T - 1) / AK::T));
            kern<<<grid, AK::THREADS, AK::smemBytes, st>>>(gamma_, L, dl, dU, dB, ncols, c0, ldbRow, ldbCol, dR, ldrRow, ldrCol);
            ++launches_;
          };
          int c0 = 0;
          for (; c0 + 2 <= ncols; c0 += 2) pass(std::integral_constant<int, 2>{}, c0);
          if (c0 < ncols) pass(std::integral_constant<int, 1>{}, c0);
        });
      }
    } else if (w0 > 0 && w1 > 0) {
      auto runPhys = [&](auto phys) {
        using Phys = decltype(phys);
        dispatchScheme(S_, [&](auto sTag) {
          constexpr int S = decltype(sTag)::value;
          // ONE operand column: the y-marching (value, tangent) kernel; the vector loads of 2- and 4-dof cells need a
          // 32-byte aligned operand / result (operands with several columns keep the tile kernel: it shares the
          // reconstruction gradients between 4 columns)
          if (ncols == 1 && applyMarch2dEnabled() &&
              ((reinterpret_cast<uintptr_t>(dB) | reinterpret_cast<uintptr_t>(dR)) & 31) == 0) {
            launchApplyMarch2d<Phys, S>(phys, L, dl, dU, dB, dR, st);
            ++launches_;
            return;
          }
          constexpr int NC = 4;
          using AK = dev::ApplyLat2d<Phys, NC>;
          auto kern = dev::k_applyjac_lattice2d<Phys, S, NC>;
          ensureFuncAttrs(kern, (int)AK::smemBytes);
          dim3 grid((unsigned)((w0 + AK::T - 1) / AK::T), (unsigned)((w1 + AK::T - 1) / AK::T));
          for (int c0 = 0; c0 < ncols; c0 += NC) {
            kern<<<grid, AK::THREADS, AK::smemBytes, st>>>(phys, L, dl, dU, dB, ncols, c0, ldbRow, ldbCol, dR, ldrRow, ldrCol);
            ++launches_;
          }
        });
      };
      switch (family_) {
        case F_EULER2D: runPhys(dev::Euler<2>{gamma_}); break;
        case F_SWE2D: runPhys(dev::Swe2d{physParams_[0], physParams_[1]}); break;
        case F_ADVDIFF2D:
          runPhys(dev::Burgers2d{{physParams_[0] * (mm.dInv[0] * mm.dInv[0]), physParams_[0] * (mm.dInv[1] * mm.dInv[1])}});
          break;
        default: {
          const double D = physParams_[2];
          runPhys(dev::LinAdv<2>{{physParams_[0], physParams_[1]}, {D * (mm.dInv[0] * mm.dInv[0]), D * (mm.dInv[1] * mm.dInv[1])},
                                 physParams_[3], 1.0, nullptr});
        }
      }
    }
    if (rowMajor3dTiled) nearBdRows();
    PDA_CUDA(cudaGetLastError());
    return;
  }
  evaluateDev(dU, t, ds.dV.p, ds.dJ.p, st);
  if (ncols >= 8) {
    // operands with many columns: per-cell kernel on row-major data (column-major operands are transposed around it)
    if (!ds.spmmReady) {
      ds.dCellBase.upload(cellBase_);
      ds.dCellLen.upload(cellLen_);
      Mesh& m = *mesh_;
      if (m.lattice && dim_ >= 2 && m.nSample == m.nStencil) {
        // tile-major visiting order: the B rows of a cell's stencil neighbours stay in L1 between cells
        const int32_t nx = m.n[0], ny = m.n[1], nz = (dim_ == 3) ? m.n[2] : 1;
        const int32_t T = (dim_ == 2) ? 8 : 4, Tz = (dim_ == 3) ? T : 1;
        std::vector<int32_t> order;
        order.reserve((size_t)m.nSample);
        for (int32_t k0 = 0; k0 < nz; k0 += Tz)
          for (int32_t j0 = 0; j0 < ny; j0 += T)
            for (int32_t i0 = 0; i0 < nx; i0 += T)
              for (int32_t k = k0; k < std::min(nz, k0 + Tz); ++k)
                for (int32_t j = j0; j < std::min(ny, j0 + T); ++j)
                  for (int32_t i = i0; i < std::min(nx, i0 + T); ++i) order.push_back((k * ny + j) * nx + i);
        ds.dCellOrder.upload(order);
      }
      ds.spmmReady = true;
    }
    const double* Bp = dB;
    double* Rp = dR;
    auto transpose = [&](const double* in, int64_t rows, int64_t cols, double* out) {
      const int64_t tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
      dev::k_transpose<<<(unsigned)tiles, 256, 0, st>>>(in, rows, cols, out);
      ++launches_;
    };
    if (layout != 1) {   // column-major: memory is [ncols][nJc]
      ds.dBt.alloc((size_t)nJc * ncols);
      ds.dRt.alloc((size_t)nrows * ncols);
      transpose(dB, ncols, nJc, ds.dBt.p);
      Bp = ds.dBt.p; Rp = ds.dRt.p;
    }
    const int32_t ncells = mesh_->nSample;
    const int32_t* order = ds.dCellOrder.p;
    const unsigned grid = (unsigned)gridFor(ncells, 64);   // 64 cells (one tile of the visiting order) per CTA
    // shared memory: two chunk slots per warp, sized for the longest row block of this problem
    int32_t maxLen = 0;
    if (ds.spmmMaxLen == 0) { for (int32_t l : cellLen_) maxLen = std::max(maxLen, l); ds.spmmMaxLen = maxLen; }
    const int slotDoubles = (ndpc_ * ds.spmmMaxLen + 1) & ~1;
    const size_t smem = (size_t)16 * slotDoubles * sizeof(double);
    auto launch = [&](auto nTag) {
      constexpr int NN = decltype(nTag)::value;
      auto kern = dev::k_spmm_cells_rowmajor<NN>;
      ensureFuncAttrs(kern, (int)smem);
      kern<<<grid, 256, smem, st>>>(ncells, order, ds.dCellBase.p, ds.dCellLen.p, ds.dColidx.p, ds.dJ.p, Bp, ncols, Rp, slotDoubles);
    };
    switch (ndpc_) {
      case 1: launch(std::integral_constant<int, 1>{}); break;
      case 2: launch(std::integral_constant<int, 2>{}); break;
      case 3: launch(std::integral_constant<int, 3>{}); break;
      case 4: launch(std::integral_constant<int, 4>{}); break;
      default: launch(std::integral_constant<int, 5>{}); break;
    }
    ++launches_;
    if (layout != 1) transpose(ds.dRt.p, nrows, ncols, dR);
  } else {
    // row-major: B[r*ncols + c]; col-major: B[c*rows + r]
    const int64_t ldbRow = (layout == 1) ? ncols : 1, ldbCol = (layout == 1) ? 1 : nJc;
    const int64_t ldrRow = (layout == 1) ? ncols : 1, ldrCol = (layout == 1) ? 1 : nrows;
    if (!ds.spmmFewReady) { ds.dCellBase.upload(cellBase_); ds.dCellLen.upload(cellLen_); ds.spmmFewReady = true; }
    const int32_t ncells = mesh_->nSample;
    const unsigned grid = (unsigned)gridFor((int64_t)ncells * 32, 256);
    auto launch = [&](auto nTag) {
      constexpr int NN = decltype(nTag)::value;
      // columns in groups of 4, 2, 1 (J is re-read per group: operands with >= 8 columns take the other kernel)
      int c0 = 0;
      while (c0 < ncols) {
        const int left = ncols - c0;
        if (left >= 4) { dev::k_spmm_cells_fewcols<NN, 4><<<grid, 256, 0, st>>>(ncells, nullptr, ds.dCellBase.p, ds.dCellLen.p, ds.dColidx.p, ds.dJ.p, dB, c0, ldbRow, ldbCol, dR, ldrRow, ldrCol); c0 += 4; }
        else if (left >= 2) { dev::k_spmm_cells_fewcols<NN, 2><<<grid, 256, 0, st>>>(ncells, nullptr, ds.dCellBase.p, ds.dCellLen.p, ds.dColidx.p, ds.dJ.p, dB, c0, ldbRow, ldbCol, dR, ldrRow, ldrCol); c0 += 2; }
        else { dev::k_spmm_cells_fewcols<NN, 1><<<grid, 256, 0, st>>>(ncells, nullptr, ds.dCellBase.p, ds.dCellLen.p, ds.dColidx.p, ds.dJ.p, dB, c0, ldbRow, ldbCol, dR, ldrRow, ldrCol); c0 += 1; }
        ++launches_;
      }
    };
    switch (ndpc_) {
      case 1: launch(std::integral_constant<int, 1>{}); break;
      case 2: launch(std::integral_constant<int, 2>{}); break;
      case 3: launch(std::integral_constant<int, 3>{}); break;
      case 4: launch(std::integral_constant<int, 4>{}); break;
      default: launch(std::integral_constant<int, 5>{}); break;
    }
  }
  PDA_CUDA(cudaGetLastError());
}

void Problem::applyJacobianHost(const double* U, const double* B, int ncols, int layout, double t, double* R) {
  if (!U || !B || !R) throw Error(kInvalid, "applyJacobian: null pointer");
  if (ncols < 1) throw Error(kInvalid, "applyJacobian: ncols must be >= 1");
  ensureDevice();
  PDA_CUDA(cudaSetDevice(device_));
  DeviceState& ds = *dev_;
  const size_t nU = (size_t)nDofStencil(), nR = (size_t)nDofSample();
  ds.dU.alloc(nU); ds.dB.alloc(nU * ncols); ds.dR.alloc(nR * ncols);
  PDA_CUDA(cudaMemcpyAsync(ds.dU.p, U, nU * sizeof(double), cudaMemcpyHostToDevice, ds.stream));
  PDA_CUDA(cudaMemcpyAsync(ds.dB.p, B, nU * ncols * sizeof(double), cudaMemcpyHostToDevice, ds.stream));
  applyJacobianDev(ds.dU.p, ds.dB.p, ncols, layout, t, ds.dR.p, ds.stream);
  PDA_CUDA(cudaMemcpyAsync(R, ds.dR.p, nR * ncols * sizeof(double), cudaMemcpyDeviceToHost, ds.stream));
  PDA_CUDA(cudaStreamSynchronize(ds.stream));
}

// ----------------------------------------------------------------------------------------------- steppers
// Explicit time stepping with the state resident in HBM.  Same stage arithmetic as the steppers the
// reference's tests use (tests_cpp/pressio/include/pressio/ode/impl/ode_explicit_stepper_without_mass_matrix.hpp:
// ForwardEuler :168-186, SSPRungeKutta3 :230-281, RungeKutta4 :284-340) and the RK2 of the reference's Python module
// (pressiodemoapps/__init__.py:90-113), so the reference's gold files are reproduced end to end on the GPU.
namespace {
__global__ void k_lincomb3(int64_t n, double* __restrict__ out, double a, const double* __restrict__ x, double b,
                           const double* __restrict__ y, double c, const double* __restrict__ z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r = a * x[i] + b * y[i];
  if (z) r += c * z[i];
  out[i] = r;
}
__global__ void k_rk4_combine(int64_t n, double* __restrict__ y, double c1, const double* __restrict__ k1, double c2,
                              const double* __restrict__ k2, const double* __restrict__ k3, const double* __restrict__ k4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = y[i] + c1 * k1[i] + c2 * k2[i] + c2 * k3[i] + c1 * k4[i];
}
}  // namespace

void Problem::advanceDev(int scheme, double* dU, double t0, double dt, int32_t nsteps, void* streamV) {
  if (!dU) throw Error(kInvalid, "advance: null pointer");
  if (scheme < 0 || scheme > 3) throw Error(kInvalid, "advance: invalid stepper enum");
  if (nsteps < 0) throw Error(kInvalid, "advance: negative step count");
  if (mesh_->nSample != mesh_->nStencil) throw Error(kInvalid, "advance: time stepping needs a full mesh (sample mesh == stencil mesh)");
  ensureDevice();
  PDA_CUDA(cudaSetDevice(device_));
  DeviceState& ds = *dev_;
  cudaStream_t st = (cudaStream_t)streamV;
  const int64_t n = nDofStencil();
  ds.stAux.alloc((size_t)n);
  for (int i = 0; i < (scheme == 2 ? 4 : 1); ++i) ds.stK[i].alloc((size_t)n);
  const unsigned grid = (unsigned)((n + 255) / 256);
  auto lin = [&](double* out, double a, const double* x, double b, const double* y, double c, const double* z) {
    k_lincomb3<<<grid, 256, 0, st>>>(n, out, a, x, b, y, c, z);
    ++launches_;
  };
  double* aux = ds.stAux.p;
  double* k1 = ds.stK[0].p;
  double t = t0;
  for (int32_t s = 0; s < nsteps; ++s, t += dt) {   // time accumulates like ode_advance_n_steps.hpp:114
    switch (scheme) {
      case 0:   // forward Euler: y += dt f(y, t)
        evaluateDev(dU, t, k1, nullptr, st);
        lin(dU, 1.0, dU, dt, k1, 0.0, nullptr);
        break;
      case 1: { // RK2 (Heun) of the Python module
        evaluateDev(dU, t, k1, nullptr, st);
        lin(aux, 1.0, dU, dt, k1, 0.0, nullptr);
        double* k2 = aux;   // f(aux) may not alias its input: use a second buffer
        ds.stK[1].alloc((size_t)n);
        k2 = ds.stK[1].p;
        evaluateDev(aux, t + dt, k2, nullptr, st);
        lin(dU, 1.0, dU, 0.5 * dt, k2, 0.5 * dt, k1);
        break;
      }
      case 2: { // RK4
        double *k2 = ds.stK[1].p, *k3 = ds.stK[2].p, *k4 = ds.stK[3].p;
        const double half = dt / 2.0;
        evaluateDev(dU, t, k1, nullptr, st);
        lin(aux, 1.0, dU, half, k1, 0.0, nullptr);
        evaluateDev(aux, t + half, k2, nullptr, st);
        lin(aux, 1.0, dU, half, k2, 0.0, nullptr);
        evaluateDev(aux, t + half, k3, nullptr, st);
        lin(aux, 1.0, dU, dt, k3, 0.0, nullptr);
        evaluateDev(aux, t + dt, k4, nullptr, st);
        k_rk4_combine<<<grid, 256, 0, st>>>(n, dU, dt / 6.0, k1, dt / 3.0, k2, k3, k4);
        ++launches_;
        break;
      }
      default: { // SSPRK3
        evaluateDev(dU, t, k1, nullptr, st);
        lin(aux, 1.0, dU, dt, k1, 0.0, nullptr);                       // u1 = u + dt f(u, t)
        evaluateDev(aux, t + dt, k1, nullptr, st);
        lin(aux, 0.25, aux, 0.75, dU, 0.25 * dt, k1);                  // u2 = 1/4 u1 + 3/4 u + 1/4 dt f(u1, t+dt)
        evaluateDev(aux, t + dt / 2.0, k1, nullptr, st);
        lin(dU, 1.0 / 3.0, dU, 2.0 / 3.0, aux, (2.0 / 3.0) * dt, k1);  // u = 1/3 u + 2/3 u2 + 2/3 dt f(u2, t+dt/2)
        break;
      }
    }
  }
  PDA_CUDA(cudaGetLastError());
}

void Problem::advanceHost(int scheme, double* U, double t0, double dt, int32_t nsteps) {
  if (!U) throw Error(kInvalid, "advance: null pointer");
  ensureDevice();
  PDA_CUDA(cudaSetDevice(device_));
  DeviceState& ds = *dev_;
  const size_t n = (size_t)nDofStencil();
  ds.dU.alloc(n);
  PDA_CUDA(cudaMemcpyAsync(ds.dU.p, U, n * sizeof(double), cudaMemcpyHostToDevice, ds.stream));
  advanceDev(scheme, ds.dU.p, t0, dt, nsteps, ds.stream);
  PDA_CUDA(cudaMemcpyAsync(U, ds.dU.p, n * sizeof(double), cudaMemcpyDeviceToHost, ds.stream));
  PDA_CUDA(cudaStreamSynchronize(ds.stream));
}

void Problem::ghosts(int side, double* out) {
  if (!dev_ || !dev_->haveGhosts) throw Error(kInvalid, "ghosts: no evaluation with ghost cells has run yet");
  if (side < 0 || side >= 6) throw Error(kInvalid, "ghosts: invalid side");
  PDA_CUDA(cudaSetDevice(device_));
  DeviceState& ds = *dev_;
  PDA_CUDA(cudaDeviceSynchronize());
  const size_t n = (size_t)ds.nearBd.n * ds.hS * ndpc_;
  if (n) PDA_CUDA(cudaMemcpy(out, ds.ghost[side].p, n * sizeof(double), cudaMemcpyDeviceToHost));
}

// =============================================================================================== slab (multi-GPU)
void Problem::makeSlab(int rank, int nranks) {
  Mesh& m = *mesh_;
  if (!m.lattice || !m.fullyPeriodic) throw Error(kUnsupported, "slab decomposition needs a fully periodic full lattice");
  if (nranks < 1 || rank < 0 || rank >= nranks) throw Error(kInvalid, "slab: invalid rank / nranks");
  const int32_t nk = m.n[dim_ - 1];
  if (nk % nranks != 0) throw Error(kInvalid, "slab: the slowest axis must be divisible by the number of ranks");
  const int32_t per = nk / nranks;
  if (per < (S_ - 1) / 2) throw Error(kInvalid, "slab: fewer planes per rank than the stencil halo");
  if (!latticeKernelAvailable(family_, dim_, S_)) throw Error(kUnsupported, "slab: no structured kernel for this problem");
  slab_ = true;
  slabRank_ = rank; slabRanks_ = nranks;
  slabK0_ = rank * per;
  slabK1_ = slabK0_ + per;
}

void Problem::slabExtent(int32_t* k0, int32_t* k1, int32_t* halo, int64_t* planeDofs) const {
  if (!slab_) throw Error(kInvalid, "not a slab problem");
  const Mesh& m = *mesh_;
  int64_t plane = 1;
  for (int a = 0; a < dim_ - 1; ++a) plane *= m.n[a];
  if (k0) *k0 = slabK0_;
  if (k1) *k1 = slabK1_;
  if (halo) *halo = (S_ - 1) / 2;
  if (planeDofs) *planeDofs = plane * ndpc_;
}

void Problem::slabInitialCondition(double* Uowned) const {
  if (!slab_) throw Error(kInvalid, "not a slab problem");
  if (!(family_ == F_EULER3D && probId_ == 0) && !(family_ == F_EULER2D && probId_ == E2_PERIODIC))
    throw Error(kUnsupported, "slab initial condition: only the periodic smooth Euler problems");
  const Mesh& m = *mesh_;
  const double gm1Inv = 1.0 / (gamma_ - 1.0);
  std::vector<double> cx(m.n[0]), cy(m.n[1]), cz(m.n[2]);
  for (int32_t i = 0; i < m.n[0]; ++i) cx[i] = m.latticeCoord(0, i);
  for (int32_t j = 0; j < m.n[1]; ++j) cy[j] = m.latticeCoord(1, j);
  for (int32_t k = 0; k < m.n[2]; ++k) cz[k] = m.latticeCoord(2, k);
  if (dim_ == 3) {
#pragma omp parallel for schedule(static)
    for (int32_t k = slabK0_; k < slabK1_; ++k)
      for (int32_t j = 0; j < m.n[1]; ++j)
        for (int32_t i = 0; i < m.n[0]; ++i) {
          const double rho = 1.0 + 0.2 * std::sin(M_PI * (cx[i] + cy[j] + cz[k]));
          const double vel[3] = {1.0, 1.0, 1.0};
          double* s = Uowned + 5 * (((int64_t)(k - slabK0_) * m.n[1] + j) * m.n[0] + i);
          s[0] = rho; s[1] = rho; s[2] = rho; s[3] = rho;
          s[4] = energyFromPrim<3>(gm1Inv, rho, vel, 1.0);
        }
  } else {
#pragma omp parallel for schedule(static)
    for (int32_t j = slabK0_; j < slabK1_; ++j)
      for (int32_t i = 0; i < m.n[0]; ++i) {
        const double rho = 1.0 + (1.0 / 5.0) * std::sin(M_PI * (cx[i] + cy[j]));
        const double vel[2] = {1.0, 1.0};
        double* s = Uowned + 4 * ((int64_t)(j - slabK0_) * m.n[0] + i);
        s[0] = rho; s[1] = rho; s[2] = rho;
        s[3] = energyFromPrim<2>(gm1Inv, rho, vel, 1.0);
      }
  }
}

void Problem::slabVelocityDev(const double* dUlocal, double /*t*/, double* dVowned, void* streamV, bool boundary) {
  if (!slab_) throw Error(kInvalid, "not a slab problem");
  ensureDevice();
  PDA_CUDA(cudaSetDevice(device_));
  DeviceState& ds = *dev_;
  cudaStream_t st = (cudaStream_t)streamV;   // NULL = the legacy default stream (CUDA convention)
  Mesh& m = *mesh_;
  dev::Deltas dl{{m.dInv[0], m.dInv[1], m.dInv[2]}};
  const int h = (S_ - 1) / 2;
  const int32_t nOwned = slabK1_ - slabK0_;
  auto run = [&](auto phys) {
    using Phys = decltype(phys);
    dispatchScheme(S_, [&](auto sTag) {
      constexpr int S = decltype(sTag)::value;
      // local plane index p in [0, nOwned): interior planes [h, nOwned-h) need no halo data
      if (!boundary) {
        if (nOwned - 2 * h > 0) { launchLatticeVelocitySlab<Phys, S>(phys, m, dl, dUlocal, dVowned, st, nOwned, h, nOwned - h); ++launches_; }
      } else {
        const int32_t hi0 = std::max<int32_t>(h, nOwned - h);
        launchLatticeVelocitySlab<Phys, S>(phys, m, dl, dUlocal, dVowned, st, nOwned, 0, std::min<int32_t>(h, nOwned)); ++launches_;
        if (hi0 < nOwned) { launchLatticeVelocitySlab<Phys, S>(phys, m, dl, dUlocal, dVowned, st, nOwned, hi0, nOwned); ++launches_; }
      }
    });
  };
  switch (family_) {
    case F_EULER2D: run(dev::Euler<2>{gamma_}); break;
    case F_EULER3D: run(dev::Euler<3>{gamma_}); break;
    default: throw Error(kUnsupported, "slab: family not supported");
  }
  PDA_CUDA(cudaGetLastError());
}

// ----------------------------------------------------------------------------------------------- slab: peer mode
// The halo exchange without a collective library: every rank owns a halo buffer (two parities x {lower, upper} x h
// planes) that its ring neighbours fill with copy-engine peer copies over NVLink, each followed by a 4-byte copy of
// the evaluation's epoch into a flag.  The velocity kernel is ONE launch over all owned planes; only the CTAs whose z
// chunk touches a halo plane poll the flag (ld.acquire.sys), and they are scheduled so that interior chunks run while
// the pushes are in flight.  No SM is needed to signal (a signalling kernel could starve behind the polling CTAs).
// Double buffering by epoch parity makes the buffers race-free without a "done reading" handshake: a neighbour can
// push epoch e+2 only after its kernel e+1 completed, which needed my push e+1, which my stream ordered after my
// kernel e -- the last reader of the parity-e buffers.
namespace {
void ensurePeerBuffer(DeviceState& ds, size_t haloDoubles) {
  auto& ph = ds.peer;
  if (ph.base) return;
  ph.haloDoubles = haloDoubles;
  ph.flagsOff = ((4 * haloDoubles * sizeof(double)) + 255) & ~size_t(255);
  ph.tableOff = ph.flagsOff + 4 * DeviceState::PeerHalo::kFlagStride;
  ph.bytes = ph.tableOff + DeviceState::PeerHalo::kTableEntries * sizeof(uint32_t);
  PDA_CUDA(cudaMalloc(&ph.base, ph.bytes));
  PDA_CUDA(cudaMemset(ph.base + ph.flagsOff, 0, 4 * DeviceState::PeerHalo::kFlagStride));
  std::vector<uint32_t> table(DeviceState::PeerHalo::kTableEntries);
  for (size_t i = 0; i < table.size(); ++i) table[i] = (uint32_t)i;
  PDA_CUDA(cudaMemcpy(ph.base + ph.tableOff, table.data(), table.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  for (int i = 0; i < 2; ++i) {
    PDA_CUDA(cudaStreamCreateWithFlags(&ph.sPush[i], cudaStreamNonBlocking));
    PDA_CUDA(cudaEventCreateWithFlags(&ph.evPushed[i], cudaEventDisableTiming));
  }
  PDA_CUDA(cudaEventCreateWithFlags(&ph.evReady, cudaEventDisableTiming));
  PDA_CUDA(cudaDeviceSynchronize());
}
}  // namespace

void Problem::slabPeerHandle(unsigned char handle[64]) {
  if (!slab_) throw Error(kInvalid, "not a slab problem");
  if (dim_ != 3) throw Error(kUnsupported, "slab peer mode: 3D lattices only (2D slabs use the send/recv halo path)");
  ensureDevice();
  PDA_CUDA(cudaSetDevice(device_));
  int64_t planeDofs; int32_t h;
  slabExtent(nullptr, nullptr, &h, &planeDofs);
  ensurePeerBuffer(*dev_, (size_t)h * (size_t)planeDofs);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hd;
  PDA_CUDA(cudaIpcGetMemHandle(&hd, dev_->peer.base));
  std::memcpy(handle, &hd, 64);
}

void Problem::slabPeerConnect(const unsigned char* handles) {
  if (!handles) throw Error(kInvalid, "slab_peer_connect: null handles");
  unsigned char mine[64];
  slabPeerHandle(mine);   // allocates
  auto& ph = dev_->peer;
  const int lo = (slabRank_ - 1 + slabRanks_) % slabRanks_, hi = (slabRank_ + 1) % slabRanks_;
  const int nb[2] = {lo, hi};
  for (int i = 0; i < 2; ++i) {
    if (nb[i] == slabRank_) { ph.remote[i] = ph.base; continue; }
    if (i == 1 && hi == lo) { ph.remote[1] = ph.remote[0]; ph.opened[1] = true; continue; }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handles + (size_t)nb[i] * 64, 64);
    void* ptr = nullptr;
    PDA_CUDA(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    ph.remote[i] = static_cast<unsigned char*>(ptr);
    ph.opened[i] = true;
  }
  ph.connected = true;
}

void Problem::slabPeerConnectLocal(Problem* lo, Problem* hi) {
  if (!lo || !hi) throw Error(kInvalid, "slab_peer_connect_local: null neighbour");
  unsigned char tmp[64];
  Problem* nb[2] = {lo, hi};
  slabPeerHandle(tmp);
  for (int i = 0; i < 2; ++i) {
    if (!nb[i]->slab_ || nb[i]->S_ != S_ || nb[i]->mesh_->n[0] != mesh_->n[0] || nb[i]->mesh_->n[1] != mesh_->n[1])
      throw Error(kInvalid, "slab_peer_connect_local: neighbour is not a slab of the same lattice");
    nb[i]->slabPeerHandle(tmp);
    if (nb[i]->device_ != device_) {
      PDA_CUDA(cudaSetDevice(device_));
      int can = 0;
      PDA_CUDA(cudaDeviceCanAccessPeer(&can, device_, nb[i]->device_));
      if (!can) throw Error(kUnsupported, "slab_peer_connect_local: no peer access between the two devices");
      const cudaError_t e = cudaDeviceEnablePeerAccess(nb[i]->device_, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PDA_CUDA(e);
      (void)cudaGetLastError();
    }
    dev_->peer.remote[i] = nb[i]->dev_->peer.base;
    dev_->peer.opened[i] = false;
  }
  dev_->peer.connected = true;
}

// one evaluation = one epoch: enqueue the pushes of my boundary planes (copy engine, after `ready` fired) ...
void Problem::peerPush(const double* dU, void* readyEvent0, void* readyEvent1) {
  DeviceState& ds = *dev_;
  auto& ph = ds.peer;
  Mesh& m = *mesh_;
  const int h = (S_ - 1) / 2;
  const int32_t nOwned = slabK1_ - slabK0_;
  const size_t planeDofs = (size_t)m.n[0] * m.n[1] * ndpc_;
  const size_t haloBytes = (size_t)h * planeDofs * sizeof(double);
  ++ph.epoch;
  const int par = (int)(ph.epoch & 1u);
  const uint32_t val = ph.epoch & 0xffffu;
  const unsigned char* valSrc = ph.base + ph.tableOff + sizeof(uint32_t) * val;
  // [1] my top h planes -> upper neighbour's LOWER halo ; [0] my bottom h planes -> lower neighbour's UPPER halo
  for (int i = 1; i >= 0; --i) {
    PDA_CUDA(cudaStreamWaitEvent(ph.sPush[i], (cudaEvent_t)readyEvent0, 0));
    if (readyEvent1) PDA_CUDA(cudaStreamWaitEvent(ph.sPush[i], (cudaEvent_t)readyEvent1, 0));
    const double* src = (i == 1) ? dU + (size_t)(nOwned - h) * planeDofs : dU;
    const int side = (i == 1) ? 0 : 1;
    PDA_CUDA(cudaMemcpyAsync(ph.halo(ph.remote[i], par, side), src, haloBytes, cudaMemcpyDefault, ph.sPush[i]));
    PDA_CUDA(cudaMemcpyAsync(ph.flag(ph.remote[i], par, side), valSrc, sizeof(uint32_t), cudaMemcpyDefault, ph.sPush[i]));
    PDA_CUDA(cudaEventRecord(ph.evPushed[i], ph.sPush[i]));
  }
}

// ... and launch the kernel over owned planes [p0, p1) of the current epoch
void Problem::peerLaunch(const double* dU, double* dV, void* streamV, int32_t p0, int32_t p1) {
  DeviceState& ds = *dev_;
  auto& ph = ds.peer;
  cudaStream_t st = (cudaStream_t)streamV;
  Mesh& m = *mesh_;
  const int32_t nOwned = slabK1_ - slabK0_;
  const int par = (int)(ph.epoch & 1u);
  dev::Deltas dl{{m.dInv[0], m.dInv[1], m.dInv[2]}};
  dev::LatticeDesc L;
  for (int a = 0; a < 3; ++a) { L.n[a] = m.n[a]; L.per[a] = m.periodic[a] ? 1 : 0; }
  L.n[2] = nOwned;
  L.planeBegin = p0; L.planeEnd = p1; L.haloPlanes = 0; L.slab = 2; L.meshHalo = m.halo();
  L.haloLo = ph.halo(ph.base, par, 0); L.haloHi = ph.halo(ph.base, par, 1);
  L.flagLo = ph.flag(ph.base, par, 0); L.flagHi = ph.flag(ph.base, par, 1);
  L.epoch = ph.epoch & 0xffffu;
  dispatchScheme(S_, [&](auto sTag) {
    constexpr int S = decltype(sTag)::value;
    launchLattice3dTiled<dev::Euler<3>, S>(dev::Euler<3>{gamma_}, L, dl, dU, dV, st);
    ++launches_;
  });
  PDA_CUDA(cudaGetLastError());
}

void Problem::peerCheck() {
  if (!slab_) throw Error(kInvalid, "not a slab problem");
  ensureDevice();
  if (!dev_->peer.connected) throw Error(kInvalid, "slab peer mode: pda_slab_peer_connect has not been called");
  if (family_ != F_EULER3D) throw Error(kUnsupported, "slab peer mode: Euler3d only");
  Mesh& m = *mesh_;
  if (m.n[0] < 2 * m.halo() || m.n[1] < 2 * m.halo()) throw Error(kUnsupported, "slab peer mode: mesh too small for the tiled kernel");
  PDA_CUDA(cudaSetDevice(device_));
}

void Problem::slabVelocityPeerDev(const double* dU, double /*t*/, double* dV, void* streamV) {
  if (!dU || !dV) throw Error(kInvalid, "velocity: null pointer");
  peerCheck();
  auto& ph = dev_->peer;
  cudaStream_t st = (cudaStream_t)streamV;
  // U as of everything enqueued on `st` so far is what the neighbours receive
  PDA_CUDA(cudaEventRecord(ph.evReady, st));
  peerPush(dU, ph.evReady, nullptr);
  peerLaunch(dU, dV, streamV, 0, slabK1_ - slabK0_);
  // the caller may overwrite U once `st` reaches this point: both pushes must have left by then
  for (int i = 0; i < 2; ++i) PDA_CUDA(cudaStreamWaitEvent(st, ph.evPushed[i], 0));
}

// host-pointer flavour: chunks of owned planes flow H2D -> kernel -> D2H on three streams like velocityHost; the two
// boundary chunks are uploaded first so that the pushes to the neighbours leave while the interior chunks stream in
void Problem::slabVelocityPeerHost(const double* U, double t, double* V) {
  if (!U || !V) throw Error(kInvalid, "velocity: null pointer");
  peerCheck();
  DeviceState& ds = *dev_;
  auto& ph = ds.peer;
  Mesh& m = *mesh_;
  const int h = (S_ - 1) / 2;
  const int32_t nOwned = slabK1_ - slabK0_;
  const size_t planeDofs = (size_t)m.n[0] * m.n[1] * ndpc_;
  const size_t n = (size_t)nOwned * planeDofs;
  ds.dU.alloc(n); ds.dV.alloc(n);
  ds.ensurePipeline();
  const int nChunks = (int)std::min<int64_t>(DeviceState::kMaxChunks, nOwned / std::max(8, 2 * h));
  // the previous call's kernels / pushes must be done before dU is overwritten
  PDA_CUDA(cudaEventRecord(ds.evOut[0], ds.stream));
  PDA_CUDA(cudaStreamWaitEvent(ds.sH2D, ds.evOut[0], 0));
  if (ph.epoch > 0) for (int i = 0; i < 2; ++i) PDA_CUDA(cudaStreamWaitEvent(ds.sH2D, ph.evPushed[i], 0));
  if (nChunks < 3) {
    PDA_CUDA(cudaMemcpyAsync(ds.dU.p, U, n * sizeof(double), cudaMemcpyHostToDevice, ds.sH2D));
    PDA_CUDA(cudaEventRecord(ds.evIn[0], ds.sH2D));
    PDA_CUDA(cudaStreamWaitEvent(ds.stream, ds.evIn[0], 0));
    slabVelocityPeerDev(ds.dU.p, t, ds.dV.p, ds.stream);
    PDA_CUDA(cudaMemcpyAsync(V, ds.dV.p, n * sizeof(double), cudaMemcpyDeviceToHost, ds.stream));
    PDA_CUDA(cudaStreamSynchronize(ds.stream));
    return;
  }
  auto c0 = [&](int c) { return (int32_t)((int64_t)nOwned * c / nChunks); };
  auto h2d = [&](int c) {
    const size_t off = (size_t)c0(c) * planeDofs, cnt = (size_t)(c0(c + 1) - c0(c)) * planeDofs;
    PDA_CUDA(cudaMemcpyAsync(ds.dU.p + off, U + off, cnt * sizeof(double), cudaMemcpyHostToDevice, ds.sH2D));
    PDA_CUDA(cudaEventRecord(ds.evIn[c], ds.sH2D));
  };
  auto compute = [&](int c) {
    for (int w = std::max(0, c - 1); w <= std::min(nChunks - 1, c + 1); ++w) PDA_CUDA(cudaStreamWaitEvent(ds.stream, ds.evIn[w], 0));
    peerLaunch(ds.dU.p, ds.dV.p, ds.stream, c0(c), c0(c + 1));
    PDA_CUDA(cudaEventRecord(ds.evOut[c], ds.stream));
    PDA_CUDA(cudaStreamWaitEvent(ds.sD2H, ds.evOut[c], 0));
    const size_t off = (size_t)c0(c) * planeDofs, cnt = (size_t)(c0(c + 1) - c0(c)) * planeDofs;
    PDA_CUDA(cudaMemcpyAsync(V + off, ds.dV.p + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, ds.sD2H));
  };
  h2d(nChunks - 1);
  h2d(0);
  peerPush(ds.dU.p, ds.evIn[nChunks - 1], ds.evIn[0]);   // boundary planes are on the device: send them off
  for (int c = 1; c < nChunks - 1; ++c) {
    h2d(c);
    compute(c - 1);
  }
  compute(nChunks - 2);
  compute(nChunks - 1);
  PDA_CUDA(cudaStreamSynchronize(ds.sD2H));
  PDA_CUDA(cudaStreamSynchronize(ds.stream));
}

}  // namespace pda
