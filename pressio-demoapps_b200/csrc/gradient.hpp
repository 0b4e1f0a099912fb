// gradient.hpp -- normal gradients of a cell-centred field at the faces on the domain boundary (2D).
//
// B200-native counterpart of the reference's GradientEvaluator (include/pressiodemoapps/gradient.hpp:61-121) and of
// impl::GradientEvaluatorInternal (impl/gradient_2d.hpp:141-293).  The face list is host data built once from the
// mesh (initializeForStoringNormalGradsAtBoundaryFaces, gradient_2d.hpp:233-286); the one-sided finite differences
// (face_normal_gradient_for_cell_centered_function_2d, gradient_2d.hpp:62-104) run on the device, one thread per
// (face, dof), in the reference's operation order without FMA contraction, so the results are bit-identical.
// There is no CPU evaluation path.
#pragma once
#include <cstdint>
#include <vector>

#include "mesh.hpp"

namespace pda {

class GradientEvaluator {
 public:
  // throws Error(kUnsupported, "gradients currently only supported for 2D") like gradient.hpp:71-73
  GradientEvaluator(Mesh& mesh, int maxNumDofPerCell);
  ~GradientEvaluator();
  GradientEvaluator(const GradientEvaluator&) = delete;
  GradientEvaluator& operator=(const GradientEvaluator&) = delete;

  int32_t numFaces() const { return (int32_t)cellGid_.size(); }
  int maxNumDofPerCell() const { return maxNdpc_; }
  // per face, in the order the faces were created: rows of graphRowsOfCellsStrictlyOnBd(), then Left, Front, Right, Back
  const std::vector<int32_t>& cellGid() const { return cellGid_; }
  const std::vector<int32_t>& position() const { return position_; }       // FacePosition: 0 Left 1 Front 2 Right 3 Back
  const std::vector<int32_t>& parentRow() const { return parentRow_; }     // parentCellGraphRow
  const std::vector<int32_t>& normalDirection() const { return normalDir_; }   // 1 = x, 2 = y
  const std::vector<double>& centers() const { return centers_; }          // [numFaces][3]
  // index of the face (cellGID, position), -1 when the mesh has no such boundary face (queryFace, gradient_2d.hpp:157-162)
  int32_t findFace(int32_t cellGid, int position) const;

  // normalGrad[face][dof] for field[stencilCell][dof]; device pointers, asynchronous on `stream`
  void computeDev(const double* dField, int numDofPerCell, double* dNormalGrad, void* stream);
  // host pointers: staged copy in, kernel, copy out (synchronous)
  void computeHost(const double* field, int numDofPerCell, double* normalGrad);
  int64_t launchCount() const { return launches_; }

 private:
  void checkNdpc(int numDofPerCell) const;
  void ensureDevice();

  int maxNdpc_ = 1;
  int stencil_ = 3;
  int32_t nStencil_ = 0;
  double h_[2] = {0, 0};
  std::vector<int32_t> cellGid_, position_, parentRow_, normalDir_;
  std::vector<double> centers_;
  std::vector<int32_t> table_;   // [numFaces][4]: cell, first neighbour inwards, second neighbour inwards, flags
  // device side (created on first use)
  int device_ = -1;
  int32_t* dTable_ = nullptr;
  double* dField_ = nullptr;
  double* dOut_ = nullptr;
  size_t fieldCap_ = 0, outCap_ = 0;
  void* stream_ = nullptr;
  int64_t launches_ = 0;
};

}  // namespace pda
