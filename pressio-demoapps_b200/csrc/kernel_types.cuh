// kernel_types.cuh -- plain argument structs shared by the kernel translation units
#pragma once
#include <cstdint>

namespace pda {
namespace dev {

struct RowSet {
  const int32_t* graph;   // compact [n][ncols]
  const int32_t* rowIds;  // sample-mesh row of compact row r
  int32_t n;
  int32_t ncols;
};

struct GhostView {
  double* g[6];     // per side: [numNearBd][stride]
  int32_t stride;   // ndpc * (schemeStencil-1)/2
};

struct Deltas { double hInv[3]; };

// Where the blocks of a cell live in the CSR value array: all ndpc rows of a cell share one column pattern, so
// entry (k, block slot s, j) sits at  base + k*len + s*ndpc + j.
struct JacLayout {
  const int32_t* base;   // [n] rowptr of the cell's first row
  const int32_t* len;    // [n] entries per row
  const uint8_t* slot;   // [n][nslotCols] block position of graph column c in the (sorted) row; 0xFF = absent
  int32_t nslotCols;
};

}  // namespace dev
}  // namespace pda
