// engine.hpp -- the problem object behind the C-ABI: host-side mirror of the reference's
// PublicProblemEigenMixinCpp<EigenApp<...>> (include/pressiodemoapps/adapter_cpp.hpp:59-264): sizes, initial
// condition, fixed CSR pattern, and the velocity / Jacobian / applyJacobian evaluations, which run as CUDA kernels.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "mesh.hpp"

namespace pda {

struct DeviceState;   // everything that lives in HBM (engine.cu)

using BcGhostFn = void (*)(void*, int32_t, const int32_t*, double, double, const double*, int, double, double*);
using BcFactorFn = void (*)(void*, const int32_t*, double, double, int, double*);

struct BcRule {
  int kind = -1;               // PDA_BC_* or -1 (unset)
  double values[5] = {0, 0, 0, 0, 0};
  BcGhostFn ghostFn = nullptr;     // kind 3 (host callback)
  BcFactorFn factorFn = nullptr;
  void* user = nullptr;
};

class Problem {
 public:
  Problem(Mesh* mesh, int family, int problemId, int recon, int icFlag, int nparams, const char* const* names,
          const double* values, int device);
  ~Problem();

  int ndpc() const { return ndpc_; }
  int dim() const { return dim_; }
  int schemeStencil() const { return S_; }
  int32_t nDofSample() const { return mesh_->nSample * ndpc_; }
  int32_t nDofStencil() const { return mesh_->nStencil * ndpc_; }
  double queryParameter(const std::string& name) const;
  void initialCondition(double* U) const;
  void setBc(int side, int kind, const double* values);
  void setBcCallback(int side, BcGhostFn ghost, BcFactorFn factors, void* user);
  void setBcPointer(int side, void* user);   // setBCPointer (euler_2d_prob_class.hpp:213-216)
  void setSource(const double* values);   // nSample doubles (host)
  void setOption(const std::string& name, const std::string& value);   // pda_problem_set_option
  std::string getOption(const std::string& name) const;

  int64_t jacobianNnz();
  void jacobianPattern(int32_t* rowptr, int32_t* colidx);

  void velocityHost(const double* U, double t, double* V);
  void velocityAndJacobianHost(const double* U, double t, double* V, double* Jvalues);
  void applyJacobianHost(const double* U, const double* B, int ncols, int layout, double t, double* R);
  void velocityDev(const double* dU, double t, double* dV, void* stream);
  void velocityAndJacobianDev(const double* dU, double t, double* dV, double* dJ, void* stream);
  void applyJacobianDev(const double* dU, const double* dB, int ncols, int layout, double t, double* dR, void* stream);
  // device-resident explicit time stepping (U never leaves HBM between evaluations): engine.cu "steppers"
  void advanceDev(int scheme, double* dU, double t0, double dt, int32_t nsteps, void* stream);
  void advanceHost(int scheme, double* U, double t0, double dt, int32_t nsteps);
  void ghosts(int side, double* out);
  int64_t launchCount() const { return launches_; }

  // slab decomposition (multi-GPU): see engine_slab in engine.cu
  void makeSlab(int rank, int nranks);
  void slabExtent(int32_t* k0, int32_t* k1, int32_t* halo, int64_t* planeDofs) const;
  void slabInitialCondition(double* Uowned) const;
  void slabVelocityDev(const double* dUlocal, double t, double* dVowned, void* stream, bool boundary);
  // peer-memory halo exchange fused with the evaluation (engine.cu, "slab: peer mode")
  void slabPeerHandle(unsigned char handle[64]);
  void slabPeerConnect(const unsigned char* handles);           // nranks x 64 bytes, indexed by rank
  void slabPeerConnectLocal(Problem* lo, Problem* hi);          // same-process neighbours (tests, single-process multi-GPU)
  void slabVelocityPeerDev(const double* dUowned, double t, double* dVowned, void* stream);
  void slabVelocityPeerHost(const double* Uowned, double t, double* Vowned);

 private:
  friend struct DeviceState;
  void buildPattern();
  void ensureDevice();
  void ensureInnerRows();
  void buildGhostRecipes();
  void ensureSource();
  void runHostBcCallbacks(const double* dU, void* stream);
  void evaluateDev(const double* dU, double t, double* dV, double* dJ, void* stream);
  void evaluatePlanes(const double* dU, double t, double* dV, void* stream, int32_t p0, int32_t p1);
  void peerCheck();
  void peerPush(const double* dU, void* readyEvent0, void* readyEvent1);
  void peerLaunch(const double* dU, double* dV, void* stream, int32_t p0, int32_t p1);

  Mesh* mesh_;
  int family_, probId_, recon_, icFlag_;
  int dim_ = 0, ndpc_ = 0, S_ = 3;
  int device_ = 0;
  // parameters (reference defaults: impl/*_parametrization_helpers.hpp)
  double gamma_ = 1.4;
  std::vector<double> icParams_, physParams_;
  double gs_[4] = {0.0002, 0.00005, 0.042, 0.062};   // Du, Dv, F, k (diffusion_reaction2d.hpp:152-155)
  BcRule bc_[6];
  bool customBcs_ = false;
  std::vector<double> srcHost_;   // per-sample-row source values (ProblemA families)
  bool srcUser_ = false;

  // fixed CSR pattern (host), built on first request
  bool havePattern_ = false;
  bool mergedNeighbors_ = false;         // coincident neighbours on tiny periodic meshes
  std::vector<int32_t> rowptr_, colidx_;
  std::vector<int32_t> cellBase_, cellLen_;   // per sample cell: rowptr of its first row, entries per row
  std::vector<uint8_t> slots_;                // [nSample][slotCols_]
  int slotCols_ = 0;

  // slab
  bool slab_ = false;
  int32_t slabK0_ = 0, slabK1_ = 0;
  int slabRank_ = 0, slabRanks_ = 1;

  bool refOrderJac_ = false;         // option "jacobian_order" = "reference" (kernels_reforder.cu)
  bool refOrderVel_ = false;         // option "velocity_order" = "reference"
  bool skipInnerJacobian_ = false;   // applyJacobian, matrix-free inner rows: assemble the near-boundary rows only
  int64_t launches_ = 0;
  std::unique_ptr<DeviceState> dev_;
};

}  // namespace pda
