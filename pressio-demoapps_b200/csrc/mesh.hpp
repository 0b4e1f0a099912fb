// mesh.hpp -- host-side cell-centred uniform mesh (connectivity graph + coordinates + row classification).
//
// B200-native counterpart of the reference's CellCenteredUniformMesh (include/pressiodemoapps/impl/mesh_ccu.hpp:67-473)
// and of its Python mesh generators (meshing_scripts/create_full_mesh.py, create_sample_mesh.py).  A full mesh in
// natural ordering is kept as a *lattice descriptor* (n, periodic flags, bounds): nothing O(cells) exists on the host
// until a caller asks for it, which is what lets a 512^3 stencil-7 mesh (10 GB of graph in the reference's format)
// be created in microseconds and evaluated by the structured kernels without any graph in HBM.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pda {

struct Mesh {
  int dim = 0;
  int stencil = 0;            // stencil size of the MESH (3,5,7); a problem may use a narrower scheme
  int32_t nSample = 0;        // rows of the graph = cells where the velocity is computed
  int32_t nStencil = 0;       // cells carrying state
  double d[3] = {0, 0, 0};    // dx,dy,dz as the reference sees them (rounded to 14 decimals by the file format)
  double dInv[3] = {0, 0, 0};
  double bounds[6] = {0, 0, 0, 0, 0, 0};
  bool hasBounds = false;

  // lattice descriptor (full mesh, natural ordering, gid = k*nx*ny + j*nx + i)
  bool lattice = false;
  int32_t n[3] = {1, 1, 1};
  bool periodic[3] = {false, false, false};

  bool isSample = false;
  bool fullyPeriodic = false;

  // slab window of a full lattice (multi-GPU shard, makeWindow): the rank owns planes [winK0, winK1) of the slowest axis;
  // local cells = [winHLo halo planes | owned planes | winHHi halo planes], plane after plane, local id = local plane *
  // planeCells + position in the plane.  Rows (sample cells) = the owned cells in that order.
  bool window = false;
  int32_t winK0 = 0, winK1 = 0, winHLo = 0, winHHi = 0, winRank = 0, winRanks = 1;
  int64_t winPlaneCells = 0;

  // materialised arrays (always present for loaded / sample meshes, lazily built for lattices)
  std::vector<double> x, y, z;                 // [nStencil]
  std::vector<int32_t> graph;                  // [nSample][ncols()]
  std::vector<int32_t> rowsInner, rowsNearBd;  // graphRowsOfCellsAwayFromBd / NearBd (mesh_ccu.hpp:385-439)
  std::vector<int32_t> stencilGids;            // sample mesh: full-mesh gid of each stencil cell
  bool haveCoords = false, haveGraph = false, haveRows = false;

  int ncols() const { return (stencil - 1) * dim + 1; }
  int halo() const { return (stencil - 1) / 2; }

  // number of near-boundary / inner rows without materialising the lists
  int64_t countNearBd() const;

  void ensureCoords();
  void ensureGraph();
  void ensureRows();

  // graph row of lattice cell gid (ncols() entries), reference column conventions
  void latticeRow(int32_t gid, int32_t* row) const;
  // lattice coordinate of index idx along axis a, rounded like the mesh files ("%.14f")
  double latticeCoord(int a, int32_t idx) const;
  // true if row has a missing (-1) neighbour within the MESH stencil along any axis (mesh_ccu.hpp:162-296)
  bool rowIsNearBd(const int32_t* row) const;
  // list of near-boundary rows in ascending row order (lattice: computed analytically)
  void nearBdRows(std::vector<int32_t>& out) const;
  // graph row r (ncols() entries): from the stored graph, or synthesised for a lattice that has none
  void graphRow(int32_t r, int32_t* row) const;
  // graphRowsOfCellsStrictlyOnBd (mesh_ccu.hpp:153-155, 441-447): the near-boundary rows with a first-layer
  // neighbour missing, ascending; 2D only (the reference fills the list for 2D meshes only)
  void strictlyOnBdRows(std::vector<int32_t>& out) const;

  static Mesh makeLattice(int dim, const int32_t n[3], const double bounds[6], const int32_t periodic[3], int stencil);
  static Mesh load(const std::string& dir);
  static Mesh makeSample(Mesh& full, const int32_t* gids, int64_t ngids);
  // shard `rank` of `nranks` of a full lattice cut into slabs along its slowest axis: a sample-mesh-like mesh whose
  // stencil cells are the owned planes plus (stencil-1)/2 halo planes per side where a neighbour rank (or the periodic
  // image) exists -- none at a physical boundary, where the usual near-boundary rows / ghost cells take over
  static Mesh makeWindow(Mesh& full, int rank, int nranks);
  static Mesh fromArrays(int dim, int stencil, int32_t nSample, int32_t nStencil, const double dxyz[3],
                         const double* x, const double* y, const double* z, const int32_t* graph);
  void write(const std::string& dir);

 private:
  void classifyFromGraph();
  void detectLattice(int32_t nx, int32_t ny, int32_t nz);
};

// graph column of the k-th layer neighbour on `side` (0 left,1 front,2 right,3 back,4 bottom,5 top) -- App. A
inline int graphCol(int dim, int side, int layer) {
  if (dim == 1) return 1 + 2 * layer + (side == 2 ? 1 : 0);  // [l0 r0 | l1 r1 | l2 r2]
  const int perLayer = (dim == 2) ? 4 : 6;
  return 1 + perLayer * layer + side;
}
// (minus side, plus side) of an axis (0=x,1=y,2=z): x -> left/right, y -> back/front, z -> bottom/top
inline int minusSide(int axis) { return axis == 0 ? 0 : (axis == 1 ? 3 : 4); }
inline int plusSide(int axis) { return axis == 0 ? 2 : (axis == 1 ? 1 : 5); }

}  // namespace pda
