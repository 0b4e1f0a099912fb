// mesh.cpp -- see mesh.hpp.  Host only (no CUDA).
#include "mesh.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "common.hpp"

namespace pda {

namespace {

// value as the reference's C++ reader sees it after the Python writer printed it with "%.14f"
// (create_full_mesh.py:157-218 -> impl/mesh_read_info.hpp:84-99, impl/mesh_read_coords.hpp:73-83)
double roundLikeFile(double v) {
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.14f", v);
  return std::strtod(buf, nullptr);
}

inline int32_t wrapIndex(int32_t idx, int32_t n, bool periodic) {
  if (idx >= 0 && idx < n) return idx;
  if (!periodic) return -1;
  idx %= n;
  if (idx < 0) idx += n;
  return idx;
}

}  // namespace

// ----------------------------------------------------------------------------------------------- lattice helpers
void Mesh::latticeRow(int32_t gid, int32_t* row) const {
  const int32_t nx = n[0], ny = n[1];
  const int32_t i = gid % nx;
  const int32_t j = (dim >= 2) ? (gid / nx) % ny : 0;
  const int32_t k = (dim == 3) ? gid / (nx * ny) : 0;
  row[0] = gid;
  const int h = halo();
  for (int L = 0; L < h; ++L) {
    const int32_t il = wrapIndex(i - (L + 1), nx, periodic[0]);
    const int32_t ir = wrapIndex(i + (L + 1), nx, periodic[0]);
    if (dim == 1) {
      row[graphCol(1, 0, L)] = il;
      row[graphCol(1, 2, L)] = ir;
      continue;
    }
    const int32_t jb = wrapIndex(j - (L + 1), ny, periodic[1]);
    const int32_t jf = wrapIndex(j + (L + 1), ny, periodic[1]);
    const int32_t base = k * nx * ny;
    row[graphCol(dim, 0, L)] = (il < 0) ? -1 : base + j * nx + il;
    row[graphCol(dim, 1, L)] = (jf < 0) ? -1 : base + jf * nx + i;
    row[graphCol(dim, 2, L)] = (ir < 0) ? -1 : base + j * nx + ir;
    row[graphCol(dim, 3, L)] = (jb < 0) ? -1 : base + jb * nx + i;
    if (dim == 3) {
      const int32_t kd = wrapIndex(k - (L + 1), n[2], periodic[2]);
      const int32_t ku = wrapIndex(k + (L + 1), n[2], periodic[2]);
      row[graphCol(3, 4, L)] = (kd < 0) ? -1 : kd * nx * ny + j * nx + i;
      row[graphCol(3, 5, L)] = (ku < 0) ? -1 : ku * nx * ny + j * nx + i;
    }
  }
}

double Mesh::latticeCoord(int a, int32_t idx) const {
  if (a >= dim) return 0.0;
  // natural_order_mesh_3d.py:53-61: ox = lo + 0.5*dx (unrounded dx), x = ox + gi*dx
  const double L = bounds[2 * a + 1] - bounds[2 * a];
  const double dxRaw = L / double(n[a]);
  const double o = bounds[2 * a] + 0.5 * dxRaw;
  return roundLikeFile(o + double(idx) * dxRaw);
}

bool Mesh::rowIsNearBd(const int32_t* row) const {
  // mesh_ccu.hpp:162-296.  1D/2D test every layer of the mesh stencil; 3D tests the second layer only when
  // stencilSize==5 (stencil 7 does not exist there); the stencil-7 extension tests all three layers.
  const int h = halo();
  const int nsides = 2 * dim;
  for (int L = 0; L < h; ++L) {
    for (int s = 0; s < nsides; ++s) {
      const int side = (dim == 1) ? (s == 0 ? 0 : 2) : s;
      if (row[graphCol(dim, side, L)] == -1) return true;
    }
  }
  return false;
}

int64_t Mesh::countNearBd() const {
  if (haveRows) return (int64_t)rowsNearBd.size();
  if (!lattice) return -1;
  const int h = halo();
  int64_t inner = 1;
  for (int a = 0; a < dim; ++a) inner *= periodic[a] ? n[a] : std::max<int32_t>(0, n[a] - 2 * h);
  return (int64_t)nSample - inner;
}

void Mesh::nearBdRows(std::vector<int32_t>& out) const {
  out.clear();
  if (haveRows) { out = rowsNearBd; return; }
  if (!lattice) throw Error(kInvalid, "mesh: row lists not available");
  if (fullyPeriodic) return;
  const int h = halo();
  auto nearAxis = [&](int a, int32_t idx) { return !periodic[a] && (idx < h || idx >= n[a] - h); };
  out.reserve((size_t)countNearBd());
  for (int32_t k = 0; k < n[2]; ++k) {
    const bool bk = (dim == 3) && nearAxis(2, k);
    for (int32_t j = 0; j < n[1]; ++j) {
      const bool bj = (dim >= 2) && nearAxis(1, j);
      const int32_t base = (k * n[1] + j) * n[0];
      if (bk || bj) {
        for (int32_t i = 0; i < n[0]; ++i) out.push_back(base + i);
      } else if (!periodic[0]) {
        for (int32_t i = 0; i < std::min<int32_t>(h, n[0]); ++i) out.push_back(base + i);
        for (int32_t i = std::max<int32_t>(h, n[0] - h); i < n[0]; ++i) out.push_back(base + i);
      }
    }
  }
}

void Mesh::graphRow(int32_t r, int32_t* row) const {
  if (haveGraph) {
    const int nc = ncols();
    for (int c = 0; c < nc; ++c) row[c] = graph[(size_t)r * nc + c];
    return;
  }
  if (!lattice) throw Error(kInvalid, "mesh: graph not available");
  latticeRow(r, row);
}

void Mesh::strictlyOnBdRows(std::vector<int32_t>& out) const {
  out.clear();
  if (dim != 2) return;
  std::vector<int32_t> nb;
  nearBdRows(nb);
  int32_t row[32];
  for (int32_t r : nb) {
    graphRow(r, row);
    if (row[1] == -1 || row[2] == -1 || row[3] == -1 || row[4] == -1) out.push_back(r);   // mesh_ccu.hpp:298-326
  }
}

void Mesh::ensureCoords() {
  if (haveCoords) return;
  if (!lattice) throw Error(kInvalid, "mesh: coordinates not available");
  std::vector<double> cx(n[0]), cy(n[1]), cz(n[2]);
  for (int32_t i = 0; i < n[0]; ++i) cx[i] = latticeCoord(0, i);
  for (int32_t j = 0; j < n[1]; ++j) cy[j] = latticeCoord(1, j);
  for (int32_t k = 0; k < n[2]; ++k) cz[k] = latticeCoord(2, k);
  x.resize(nStencil); y.resize(nStencil); z.resize(nStencil);
  size_t g = 0;
  for (int32_t k = 0; k < n[2]; ++k)
    for (int32_t j = 0; j < n[1]; ++j)
      for (int32_t i = 0; i < n[0]; ++i, ++g) { x[g] = cx[i]; y[g] = cy[j]; z[g] = cz[k]; }
  haveCoords = true;
}

void Mesh::ensureGraph() {
  if (haveGraph) return;
  if (!lattice) throw Error(kInvalid, "mesh: graph not available");
  const int nc = ncols();
  graph.resize((size_t)nSample * nc);
#pragma omp parallel for schedule(static)
  for (int32_t r = 0; r < nSample; ++r) latticeRow(r, &graph[(size_t)r * nc]);
  haveGraph = true;
}

void Mesh::ensureRows() {
  if (haveRows) return;
  if (!lattice) throw Error(kInvalid, "mesh: row lists not available");
  std::vector<int32_t> nb;
  nearBdRows(nb);
  rowsInner.clear();
  rowsInner.reserve((size_t)nSample - nb.size());
  size_t p = 0;
  for (int32_t r = 0; r < nSample; ++r) {
    if (p < nb.size() && nb[p] == r) { ++p; continue; }
    rowsInner.push_back(r);
  }
  rowsNearBd.swap(nb);
  haveRows = true;
}

void Mesh::classifyFromGraph() {
  const int nc = ncols();
  rowsInner.clear();
  rowsNearBd.clear();
  bool anyNeg = false;
  for (int32_t r = 0; r < nSample; ++r) {
    const int32_t* row = &graph[(size_t)r * nc];
    if (rowIsNearBd(row)) rowsNearBd.push_back(r); else rowsInner.push_back(r);
    for (int c = 0; c < nc && !anyNeg; ++c) anyNeg = row[c] < 0;   // checkIfFullyPeriodic, mesh_ccu.hpp:331-342
  }
  fullyPeriodic = !anyNeg;
  haveRows = true;
}

// ------------------------------------------------------------------------------------------------ constructors
Mesh Mesh::makeLattice(int dim, const int32_t nIn[3], const double bd[6], const int32_t per[3], int stencil) {
  if (dim < 1 || dim > 3) throw Error(kInvalid, "mesh: dimensionality must be 1, 2 or 3");
  if (stencil != 3 && stencil != 5 && stencil != 7) throw Error(kInvalid, "mesh: stencil size must be 3, 5 or 7");
  Mesh m;
  m.dim = dim;
  m.stencil = stencil;
  int64_t total = 1;
  for (int a = 0; a < 3; ++a) {
    m.n[a] = (a < dim) ? nIn[a] : 1;
    m.periodic[a] = (a < dim) ? (per[a] != 0) : false;
    if (m.n[a] < 1) throw Error(kInvalid, "mesh: number of cells must be positive");
    if (a < dim && m.n[a] < (stencil - 1) / 2 + 1)
      throw Error(kInvalid, "mesh: too few cells along an axis for this stencil size");
    total *= m.n[a];
    m.bounds[2 * a] = (a < dim) ? bd[2 * a] : 0.0;
    m.bounds[2 * a + 1] = (a < dim) ? bd[2 * a + 1] : 0.0;
  }
  if (total > INT32_MAX) throw Error(kTooLarge, "mesh: cell count exceeds the reference's int32 index type");
  m.hasBounds = true;
  m.nSample = m.nStencil = (int32_t)total;
  for (int a = 0; a < 3; ++a) {
    if (a < dim) {
      const double L = m.bounds[2 * a + 1] - m.bounds[2 * a];
      if (!(L > 0)) throw Error(kInvalid, "mesh: invalid bounds");
      m.d[a] = roundLikeFile(L / double(m.n[a]));
      m.dInv[a] = 1.0 / m.d[a];
    } else {
      m.d[a] = 0.0;     // the reader never sets deltas of unused axes (mesh_read_info.hpp:84-99)
      m.dInv[a] = 0.0;
    }
  }
  m.lattice = true;
  m.fullyPeriodic = true;
  for (int a = 0; a < dim; ++a) m.fullyPeriodic = m.fullyPeriodic && m.periodic[a];
  return m;
}

Mesh Mesh::fromArrays(int dim, int stencil, int32_t nSample, int32_t nStencil, const double dxyz[3],
                      const double* x, const double* y, const double* z, const int32_t* graph) {
  if (dim < 1 || dim > 3) throw Error(kInvalid, "mesh: dimensionality must be 1, 2 or 3");
  if (stencil != 3 && stencil != 5 && stencil != 7) throw Error(kInvalid, "mesh: stencil size must be 3, 5 or 7");
  Mesh m;
  m.dim = dim; m.stencil = stencil; m.nSample = nSample; m.nStencil = nStencil;
  for (int a = 0; a < 3; ++a) {
    m.d[a] = (a < dim) ? dxyz[a] : 0.0;
    m.dInv[a] = (a < dim) ? 1.0 / dxyz[a] : 0.0;
  }
  m.x.assign(x, x + nStencil);
  if (y) m.y.assign(y, y + nStencil); else m.y.assign(nStencil, 0.0);
  if (z && dim == 3) m.z.assign(z, z + nStencil); else m.z.assign(nStencil, 0.0);
  m.graph.assign(graph, graph + (size_t)nSample * m.ncols());
  for (size_t i = 0; i < m.graph.size(); ++i)
    if (m.graph[i] < -1 || m.graph[i] >= nStencil) throw Error(kInvalid, "mesh: graph entry out of range");
  m.haveCoords = m.haveGraph = true;
  m.isSample = (nSample != nStencil);
  m.classifyFromGraph();
  return m;
}

void Mesh::detectLattice(int32_t nx, int32_t ny, int32_t nz) {
  if (nx <= 0) return;
  if (dim >= 2 && ny <= 0) return;
  if (dim == 3 && nz <= 0) return;
  const int64_t total = (int64_t)nx * (dim >= 2 ? ny : 1) * (dim == 3 ? nz : 1);
  if (total != nSample || nSample != nStencil) return;
  n[0] = nx; n[1] = dim >= 2 ? ny : 1; n[2] = dim == 3 ? nz : 1;
  const int nc = ncols();
  // periodic flag of an axis = "cell 0 has a minus-side neighbour"
  const int32_t* row0 = &graph[0];
  for (int a = 0; a < 3; ++a) periodic[a] = (a < dim) && row0[graphCol(dim, minusSide(a), 0)] != -1;
  std::vector<int32_t> tmp(nc);
  bool same = true;
  for (int32_t r = 0; r < nSample && same; ++r) {
    latticeRow(r, tmp.data());
    same = std::memcmp(tmp.data(), &graph[(size_t)r * nc], sizeof(int32_t) * nc) == 0;
  }
  lattice = same;
  if (!same) { n[0] = n[1] = n[2] = 1; periodic[0] = periodic[1] = periodic[2] = false; }
}

Mesh Mesh::load(const std::string& dir) {
  Mesh m;
  int32_t nx = 0, ny = 0, nz = 0;
  {  // info.dat (impl/mesh_read_info.hpp:54-119)
    std::ifstream f(dir + "/info.dat");
    if (!f) throw Error(kIO, "file not found " + dir + "/info.dat");
    std::string line;
    bool gotB[6] = {false, false, false, false, false, false};
    while (std::getline(f, line)) {
      std::istringstream ss(line);
      std::string key, val;
      ss >> key >> val;
      if (key.empty() || val.empty()) continue;
      if (key == "dim") m.dim = std::stoi(val);
      else if (key == "dx") { m.d[0] = std::stod(val); m.dInv[0] = 1.0 / m.d[0]; }
      else if (key == "dy") { m.d[1] = std::stod(val); m.dInv[1] = 1.0 / m.d[1]; }
      else if (key == "dz") { m.d[2] = std::stod(val); m.dInv[2] = 1.0 / m.d[2]; }
      else if (key == "sampleMeshSize") m.nSample = std::stoi(val);
      else if (key == "stencilMeshSize") m.nStencil = std::stoi(val);
      else if (key == "stencilSize") m.stencil = std::stoi(val);
      else if (key == "nx") nx = std::stoi(val);
      else if (key == "ny") ny = std::stoi(val);
      else if (key == "nz") nz = std::stoi(val);
      else if (key == "xMin") { m.bounds[0] = std::stod(val); gotB[0] = true; }
      else if (key == "xMax") { m.bounds[1] = std::stod(val); gotB[1] = true; }
      else if (key == "yMin") { m.bounds[2] = std::stod(val); gotB[2] = true; }
      else if (key == "yMax") { m.bounds[3] = std::stod(val); gotB[3] = true; }
      else if (key == "zMin") { m.bounds[4] = std::stod(val); gotB[4] = true; }
      else if (key == "zMax") { m.bounds[5] = std::stod(val); gotB[5] = true; }
    }
    m.hasBounds = gotB[0] && gotB[1];
  }
  if (m.dim < 1 || m.dim > 3) throw Error(kIO, "mesh: invalid or missing dim in " + dir + "/info.dat");
  if (m.stencil != 3 && m.stencil != 5 && m.stencil != 7) throw Error(kIO, "mesh: invalid stencilSize in info.dat");
  if (m.nSample <= 0 || m.nStencil <= 0) throw Error(kIO, "mesh: empty mesh (sampleMeshSize/stencilMeshSize)");

  {  // coordinates.dat (impl/mesh_read_coords.hpp:54-87)
    std::ifstream f(dir + "/coordinates.dat");
    if (!f) throw Error(kIO, "file not found " + dir + "/coordinates.dat");
    m.x.assign(m.nStencil, 0.0); m.y.assign(m.nStencil, 0.0); m.z.assign(m.nStencil, 0.0);
    std::string line;
    while (std::getline(f, line)) {
      if (line.empty()) continue;
      const char* p = line.c_str();
      char* e = nullptr;
      const long gid = std::strtol(p, &e, 10);
      if (e == p) continue;
      if (gid < 0 || gid >= m.nStencil) throw Error(kIO, "mesh: coordinates.dat gid out of range");
      p = e; m.x[gid] = std::strtod(p, &e);
      p = e; m.y[gid] = std::strtod(p, &e);
      if (m.dim == 3) { p = e; m.z[gid] = std::strtod(p, &e); }
    }
    m.haveCoords = true;
  }
  {  // connectivity.dat (impl/mesh_read_connectivity.hpp:54-83): row order = file order
    std::ifstream f(dir + "/connectivity.dat");
    if (!f) throw Error(kIO, "file not found " + dir + "/connectivity.dat");
    const int nc = m.ncols();
    m.graph.assign((size_t)m.nSample * nc, -1);
    std::string line;
    int32_t count = 0;
    while (std::getline(f, line)) {
      if (line.empty()) continue;
      if (count >= m.nSample) throw Error(kIO, "mesh: connectivity.dat has more rows than sampleMeshSize");
      const char* p = line.c_str();
      char* e = nullptr;
      for (int c = 0; c < nc; ++c) {
        const long v = std::strtol(p, &e, 10);
        if (e == p) throw Error(kIO, "mesh: connectivity.dat row too short for the stencil size");
        if (v < -1 || v >= m.nStencil) throw Error(kIO, "mesh: connectivity.dat entry out of range");
        m.graph[(size_t)count * nc + c] = (int32_t)v;
        p = e;
      }
      ++count;
    }
    if (count != m.nSample) throw Error(kIO, "mesh: connectivity.dat has fewer rows than sampleMeshSize");
    m.haveGraph = true;
  }
  m.isSample = (m.nSample != m.nStencil);
  m.classifyFromGraph();
  m.detectLattice(nx, ny, nz);
  if (m.isSample) {
    std::ifstream f(dir + "/stencil_mesh_gids.dat");
    if (f) {
      long v;
      while (f >> v) m.stencilGids.push_back((int32_t)v);
      if ((int32_t)m.stencilGids.size() != m.nStencil) m.stencilGids.clear();
    }
  }
  return m;
}

Mesh Mesh::makeSample(Mesh& full, const int32_t* gidsIn, int64_t ngids) {
  if (full.nSample != full.nStencil) throw Error(kInvalid, "sample mesh: source must be a full mesh");
  if (ngids <= 0) throw Error(kInvalid, "sample mesh: empty list of sample cells");
  std::vector<int32_t> gids(gidsIn, gidsIn + ngids);
  std::sort(gids.begin(), gids.end());                      // create_sample_mesh.py:39-40
  gids.erase(std::unique(gids.begin(), gids.end()), gids.end());
  if (gids.front() < 0 || gids.back() >= full.nSample) throw Error(kInvalid, "sample mesh: gid out of range");

  const int nc = full.ncols();
  if (!full.lattice) full.ensureGraph();
  std::vector<int32_t> rows((size_t)gids.size() * nc);
  for (size_t r = 0; r < gids.size(); ++r) {
    if (full.lattice) full.latticeRow(gids[r], &rows[r * nc]);
    else {
      const int32_t* src = &full.graph[(size_t)gids[r] * nc];
      if (src[0] != gids[r]) throw Error(kUnsupported, "sample mesh: full mesh rows are not in gid order");
      std::memcpy(&rows[r * nc], src, sizeof(int32_t) * nc);
    }
  }
  // stencil mesh = sample cells + all existing neighbours, sorted by full-mesh gid (create_sample_mesh.py:53-66)
  std::vector<int32_t> st;
  st.reserve(rows.size());
  for (int32_t v : rows) if (v >= 0) st.push_back(v);
  std::sort(st.begin(), st.end());
  st.erase(std::unique(st.begin(), st.end()), st.end());
  auto newId = [&](int32_t g) -> int32_t {
    return (int32_t)(std::lower_bound(st.begin(), st.end(), g) - st.begin());
  };

  Mesh m;
  m.dim = full.dim; m.stencil = full.stencil;
  m.nSample = (int32_t)gids.size();
  m.nStencil = (int32_t)st.size();
  for (int a = 0; a < 3; ++a) { m.d[a] = full.d[a]; m.dInv[a] = full.dInv[a]; }
  std::memcpy(m.bounds, full.bounds, sizeof m.bounds);
  m.hasBounds = full.hasBounds;
  m.isSample = true;
  m.graph.resize(rows.size());
  for (size_t i = 0; i < rows.size(); ++i) m.graph[i] = rows[i] < 0 ? -1 : newId(rows[i]);
  m.haveGraph = true;
  m.x.resize(m.nStencil); m.y.resize(m.nStencil); m.z.resize(m.nStencil);
  if (full.lattice && !full.haveCoords) {
    const int32_t nx = full.n[0], ny = full.n[1];
    for (int32_t s = 0; s < m.nStencil; ++s) {
      const int32_t g = st[s];
      m.x[s] = full.latticeCoord(0, g % nx);
      m.y[s] = full.latticeCoord(1, (g / nx) % ny);
      m.z[s] = full.latticeCoord(2, g / (nx * ny));
    }
  } else {
    for (int32_t s = 0; s < m.nStencil; ++s) { m.x[s] = full.x[st[s]]; m.y[s] = full.y[st[s]]; m.z[s] = full.z[st[s]]; }
  }
  m.haveCoords = true;
  m.stencilGids.swap(st);
  m.classifyFromGraph();
  return m;
}

Mesh Mesh::makeWindow(Mesh& full, int rank, int nranks) {
  if (!full.lattice || full.nSample != full.nStencil) throw Error(kInvalid, "slab window: source must be a full lattice");
  if (nranks < 1 || rank < 0 || rank >= nranks) throw Error(kInvalid, "slab window: invalid rank / nranks");
  const int dim = full.dim, ax = dim - 1;
  const int32_t nk = full.n[ax];
  const int h = full.halo();
  const int32_t k0 = (int32_t)((int64_t)nk * rank / nranks), k1 = (int32_t)((int64_t)nk * (rank + 1) / nranks);
  if (k1 - k0 < h) throw Error(kInvalid, "slab window: fewer planes per rank than the stencil halo");
  const bool per = full.periodic[ax];
  int64_t planeCells = 1;
  for (int a = 0; a < ax; ++a) planeCells *= full.n[a];
  // halo planes exist where a neighbour owns them: every side of a periodic axis cut into > 1 slabs, interior sides otherwise
  const int hLo = (nranks == 1) ? 0 : ((per || rank > 0) ? h : 0);
  const int hHi = (nranks == 1) ? 0 : ((per || rank < nranks - 1) ? h : 0);
  const int32_t nOwned = k1 - k0, nLocalPlanes = hLo + nOwned + hHi;
  if ((int64_t)nLocalPlanes * planeCells > INT32_MAX) throw Error(kTooLarge, "slab window: too many local cells for int32 ids");
  // local plane of global plane p (p may be the periodic image of a plane outside [0,nk))
  auto localPlane = [&](int32_t p) -> int32_t {
    for (int s : {0, -1, 1}) {
      const int64_t q = (int64_t)p + (int64_t)s * nk;
      if (!per && s != 0) continue;
      if (q >= (int64_t)k0 - hLo && q < (int64_t)k1 + hHi) return (int32_t)(q - (k0 - hLo));
    }
    return -1;
  };
  Mesh m;
  m.dim = dim; m.stencil = full.stencil;
  m.nSample = (int32_t)(nOwned * planeCells);
  m.nStencil = (int32_t)(nLocalPlanes * planeCells);
  for (int a = 0; a < 3; ++a) { m.d[a] = full.d[a]; m.dInv[a] = full.dInv[a]; }
  std::memcpy(m.bounds, full.bounds, sizeof m.bounds);
  m.hasBounds = full.hasBounds;
  m.isSample = true;
  m.window = true;
  m.winK0 = k0; m.winK1 = k1; m.winHLo = hLo; m.winHHi = hHi; m.winRank = rank; m.winRanks = nranks;
  m.winPlaneCells = planeCells;
  // the local cells form a lattice themselves: nLocalPlanes planes, NOT periodic along the slab axis, whose "near the
  // boundary" planes at an artificial edge are exactly the halo planes (not owned, never evaluated).  The structured
  // kernels run on this descriptor (engine.cu, window fast path); m.lattice stays false: rows != cells here.
  for (int a = 0; a < 3; ++a) { m.n[a] = (a < ax) ? full.n[a] : (a == ax ? nLocalPlanes : 1); m.periodic[a] = (a < ax) ? full.periodic[a] : false; }
  const int nc = full.ncols();
  m.graph.resize((size_t)m.nSample * nc);
#pragma omp parallel for schedule(static)
  for (int32_t r = 0; r < m.nSample; ++r) {
    const int32_t gid = (int32_t)((int64_t)k0 * planeCells + r);
    int32_t row[32];
    full.latticeRow(gid, row);
    int32_t* out = &m.graph[(size_t)r * nc];
    for (int c = 0; c < nc; ++c) {
      if (row[c] < 0) { out[c] = -1; continue; }
      const int32_t pg = (int32_t)(row[c] / planeCells);
      int32_t lp;
      if (c == 0) lp = hLo + (int32_t)(r / planeCells);
      else {
        // the neighbour's plane as the stencil reaches it: own plane + offset, before any periodic wrap
        const int32_t own = k0 + (int32_t)(r / planeCells);
        int32_t off = pg - own;
        if (per) { if (off > nk / 2) off -= nk; else if (off < -(nk / 2)) off += nk; }
        lp = localPlane(own + off);
        if (lp < 0 && nranks == 1 && per) lp = ((own + off) % nk + nk) % nk;   // single slab: the wrap stays inside
      }
      if (lp < 0) throw Error(kInvalid, "slab window: a stencil neighbour falls outside the window");
      out[c] = (int32_t)((int64_t)lp * planeCells + row[c] % planeCells);
    }
  }
  m.haveGraph = true;
  m.x.resize(m.nStencil); m.y.resize(m.nStencil); m.z.resize(m.nStencil);
  m.stencilGids.resize(m.nStencil);
  const int32_t nx = full.n[0], ny = full.n[1];
#pragma omp parallel for schedule(static)
  for (int32_t s = 0; s < m.nStencil; ++s) {
    const int32_t lp = (int32_t)(s / planeCells);
    int32_t gp = k0 - hLo + lp;                        // global plane (wrapped into the domain)
    if (per) gp = ((gp % nk) + nk) % nk;
    const int32_t g = (int32_t)((int64_t)gp * planeCells + s % planeCells);
    m.stencilGids[s] = g;
    m.x[s] = full.latticeCoord(0, g % nx);
    m.y[s] = (dim >= 2) ? full.latticeCoord(1, (g / nx) % ny) : 0.0;
    m.z[s] = (dim >= 3) ? full.latticeCoord(2, g / (nx * ny)) : 0.0;
  }
  m.haveCoords = true;
  m.classifyFromGraph();
  return m;
}

// --------------------------------------------------------------------------------------------------- writer
void Mesh::write(const std::string& dir) {
  ensureGraph();
  ensureCoords();
  auto open = [&](const std::string& name) {
    FILE* f = std::fopen((dir + "/" + name).c_str(), "w");
    if (!f) throw Error(kIO, "cannot open " + dir + "/" + name + " for writing");
    return f;
  };
  {  // create_full_mesh.py:151-199 / create_sample_mesh.py:160-196
    FILE* f = open("info.dat");
    static const char* mn[3] = {"xMin", "yMin", "zMin"};
    static const char* mx[3] = {"xMax", "yMax", "zMax"};
    static const char* dn[3] = {"dx", "dy", "dz"};
    static const char* nn[3] = {"nx", "ny", "nz"};
    std::fprintf(f, "dim %1d\n", dim);
    for (int a = 0; a < dim; ++a) {
      std::fprintf(f, "%s %.14f\n", mn[a], bounds[2 * a]);
      std::fprintf(f, "%s %.14f\n", mx[a], bounds[2 * a + 1]);
    }
    for (int a = 0; a < dim; ++a) std::fprintf(f, "%s %.14f\n", dn[a], d[a]);
    std::fprintf(f, "sampleMeshSize %8d\n", nSample);
    std::fprintf(f, "stencilMeshSize %8d\n", nStencil);
    std::fprintf(f, "stencilSize %2d\n", stencil);
    if (lattice && !isSample)
      for (int a = 0; a < dim; ++a) std::fprintf(f, "%s %8d\n", nn[a], n[a]);
    std::fclose(f);
  }
  {
    FILE* f = open("connectivity.dat");
    const int nc = ncols();
    for (int32_t r = 0; r < nSample; ++r) {
      for (int c = 0; c < nc; ++c) std::fprintf(f, "%8d ", graph[(size_t)r * nc + c]);
      std::fputc('\n', f);
    }
    std::fclose(f);
  }
  {
    FILE* f = open("coordinates.dat");
    for (int32_t g = 0; g < nStencil; ++g) {
      std::fprintf(f, "%8d %.14f %.14f ", g, x[g], y[g]);
      if (dim == 3) std::fprintf(f, "%.14f ", z[g]);
      std::fputc('\n', f);
    }
    std::fclose(f);
  }
  if (isSample && !stencilGids.empty()) {
    FILE* f = open("stencil_mesh_gids.dat");
    for (int32_t g : stencilGids) std::fprintf(f, "%8d\n", g);
    std::fclose(f);
  }
}

}  // namespace pda
