// gradient.cu -- GradientEvaluator: face list on the host, one-sided finite differences on the device (gradient.hpp).
#include <cuda_runtime.h>

#include <string>

#include "common.hpp"
#include "gradient.hpp"

namespace pda {

#define PDA_GCUDA(call)                                                                                   \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      throw Error(kCuda, std::string(#call) + " failed: " + cudaGetErrorString(e_));                      \
  } while (0)

namespace {

enum : int32_t { kAxisY = 1, kBackward = 2 };

// One thread per (face, dof).  The arithmetic is the reference's, operation by operation
// (gradient_2d.hpp:92-103), with explicit round-to-nearest intrinsics so that nothing is contracted into an FMA:
//   two points   forward  (-f0 + f1)/h                 backward ( f0 - f1)/h
//   three points forward  (-2 f0 + 3 f1 - 1 f2)/h      backward ( 2 f0 - 3 f1 + 1 f2)/h
// f1, f2 = first / second neighbour inwards (a missing neighbour contributes 0 like gradient_2d.hpp:86-89).
__global__ void __launch_bounds__(128)
k_boundary_face_normal_gradient(const int4* __restrict__ table, int32_t nFaces, int ndpc, int threePoint, double hx,
                                double hy, const double* __restrict__ field, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)nFaces * ndpc) return;
  const int32_t face = (int32_t)(t / ndpc);
  const int j = (int)(t - (int64_t)face * ndpc);
  const int4 e = __ldg(table + face);
  const double h = (e.w & kAxisY) ? hy : hx;
  const double f0 = __ldg(field + (int64_t)e.x * ndpc + j);
  const double f1 = (e.y != -1) ? __ldg(field + (int64_t)e.y * ndpc + j) : 0.0;
  double num;
  if (!threePoint) {
    num = (e.w & kBackward) ? __dsub_rn(f0, f1) : __dadd_rn(-f0, f1);
  } else {
    const double f2 = (e.z != -1) ? __ldg(field + (int64_t)e.z * ndpc + j) : 0.0;
    if (e.w & kBackward) num = __dadd_rn(__dsub_rn(__dmul_rn(2.0, f0), __dmul_rn(3.0, f1)), __dmul_rn(1.0, f2));
    else num = __dsub_rn(__dadd_rn(__dmul_rn(-2.0, f0), __dmul_rn(3.0, f1)), __dmul_rn(1.0, f2));
  }
  out[t] = __ddiv_rn(num, h);
}

}  // namespace

GradientEvaluator::GradientEvaluator(Mesh& mesh, int maxNumDofPerCell) {
  if (mesh.dim != 2) throw Error(kUnsupported, "gradients currently only supported for 2D");
  if (maxNumDofPerCell < 1) throw Error(kInvalid, "GradientEvaluator: MaxNumDofPerCell must be >= 1");
  maxNdpc_ = maxNumDofPerCell;
  stencil_ = mesh.stencil;
  nStencil_ = mesh.nStencil;
  h_[0] = mesh.d[0];
  h_[1] = mesh.d[1];
  // a lattice that has not materialised its coordinates (O(cells) host memory) supplies them per boundary cell
  const bool analytic = mesh.lattice && !mesh.haveCoords;
  if (!analytic) mesh.ensureCoords();
  std::vector<int32_t> rows;
  mesh.strictlyOnBdRows(rows);
  const double dxHalf = mesh.d[0] * 0.5, dyHalf = mesh.d[1] * 0.5;
  const bool wide = mesh.ncols() >= 9;
  int32_t cc[32];
  for (int32_t rowInd : rows) {
    mesh.graphRow(rowInd, cc);
    const int32_t gid = cc[0];
    const double cx = analytic ? mesh.latticeCoord(0, gid % mesh.n[0]) : mesh.x[gid];
    const double cy = analytic ? mesh.latticeCoord(1, gid / mesh.n[0]) : mesh.y[gid];
    const double cz = analytic ? mesh.latticeCoord(2, 0) : mesh.z[gid];
    // Left, Front, Right, Back = graph columns 1..4 of the first layer (mesh_ccu.hpp:298-312)
    for (int pos = 0; pos < 4; ++pos) {
      if (cc[1 + pos] != -1) continue;
      const bool alongX = (pos == 0 || pos == 2);
      // Left / Back faces difference forwards (towards +x / +y), Right / Front backwards (gradient_2d.hpp:201-205)
      const bool backward = (pos == 2 || pos == 1);
      // gradient_2d.hpp:79-83: +1.5 h -> col 3 (x) / 2 (y), -1.5 h -> col 1 / 4, +3 h -> col 7 / 6, -3 h -> col 5 / 8
      const int c1 = alongX ? (backward ? 1 : 3) : (backward ? 4 : 2);
      const int c2 = alongX ? (backward ? 5 : 7) : (backward ? 8 : 6);
      cellGid_.push_back(gid);
      position_.push_back(pos);
      parentRow_.push_back(rowInd);
      normalDir_.push_back(alongX ? 1 : 2);
      centers_.push_back(pos == 0 ? cx - dxHalf : (pos == 2 ? cx + dxHalf : cx));
      centers_.push_back(pos == 1 ? cy + dyHalf : (pos == 3 ? cy - dyHalf : cy));
      centers_.push_back(cz);
      table_.push_back(gid);
      table_.push_back(cc[c1]);
      table_.push_back(wide ? cc[c2] : -1);
      table_.push_back((alongX ? 0 : kAxisY) | (backward ? kBackward : 0));
    }
  }
}

GradientEvaluator::~GradientEvaluator() {
  if (device_ >= 0) {
    cudaSetDevice(device_);
    if (dTable_) cudaFree(dTable_);
    if (dField_) cudaFree(dField_);
    if (dOut_) cudaFree(dOut_);
    if (stream_) cudaStreamDestroy((cudaStream_t)stream_);
  }
}

int32_t GradientEvaluator::findFace(int32_t cellGid, int position) const {
  // faces of one cell are adjacent in the list and the list is ascending in the parent row, not in the gid:
  // linear scan (the list is O(perimeter))
  for (size_t i = 0; i < cellGid_.size(); ++i)
    if (cellGid_[i] == cellGid && position_[i] == position) return (int32_t)i;
  return -1;
}

void GradientEvaluator::checkNdpc(int numDofPerCell) const {
  if (numDofPerCell < 1) throw Error(kInvalid, "GradientEvaluator: numDofPerCell must be >= 1");
  if (numDofPerCell > maxNdpc_)   // gradient.hpp:87-91
    throw Error(kInvalid, "GradientEvaluator: cannot call operator() with numDofPerCell > MaxNumDofPerCell: " +
                              std::to_string(numDofPerCell) + " > " + std::to_string(maxNdpc_));
}

void GradientEvaluator::ensureDevice() {
  if (device_ >= 0) return;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error(kNoDevice, "no usable CUDA device: the B200 engine has no CPU fallback");
  }
  int dev = 0;
  PDA_GCUDA(cudaGetDevice(&dev));
  if (!table_.empty()) {
    PDA_GCUDA(cudaMalloc(&dTable_, table_.size() * sizeof(int32_t)));
    PDA_GCUDA(cudaMemcpy(dTable_, table_.data(), table_.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  cudaStream_t s;
  PDA_GCUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  stream_ = s;
  device_ = dev;
}

void GradientEvaluator::computeDev(const double* dField, int numDofPerCell, double* dNormalGrad, void* stream) {
  checkNdpc(numDofPerCell);
  ensureDevice();
  const int64_t total = (int64_t)numFaces() * numDofPerCell;
  if (total == 0) return;
  if (!dField || !dNormalGrad) throw Error(kInvalid, "GradientEvaluator: null device pointer");
  const int threads = 128;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  k_boundary_face_normal_gradient<<<blocks, threads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(dTable_), numFaces(), numDofPerCell, stencil_ == 3 ? 0 : 1, h_[0], h_[1], dField,
      dNormalGrad);
  PDA_GCUDA(cudaGetLastError());
  ++launches_;
}

void GradientEvaluator::computeHost(const double* field, int numDofPerCell, double* normalGrad) {
  checkNdpc(numDofPerCell);
  ensureDevice();
  if (!field || (!normalGrad && numFaces() > 0)) throw Error(kInvalid, "GradientEvaluator: null pointer");
  PDA_GCUDA(cudaSetDevice(device_));
  const size_t nf = (size_t)nStencil_ * numDofPerCell, no = (size_t)numFaces() * numDofPerCell;
  if (no == 0) return;
  if (nf > fieldCap_) {
    if (dField_) cudaFree(dField_);
    dField_ = nullptr;
    PDA_GCUDA(cudaMalloc(&dField_, nf * sizeof(double)));
    fieldCap_ = nf;
  }
  if (no > outCap_) {
    if (dOut_) cudaFree(dOut_);
    dOut_ = nullptr;
    PDA_GCUDA(cudaMalloc(&dOut_, no * sizeof(double)));
    outCap_ = no;
  }
  cudaStream_t s = (cudaStream_t)stream_;
  PDA_GCUDA(cudaMemcpyAsync(dField_, field, nf * sizeof(double), cudaMemcpyHostToDevice, s));
  computeDev(dField_, numDofPerCell, dOut_, s);
  PDA_GCUDA(cudaMemcpyAsync(normalGrad, dOut_, no * sizeof(double), cudaMemcpyDeviceToHost, s));
  PDA_GCUDA(cudaStreamSynchronize(s));
}

}  // namespace pda
