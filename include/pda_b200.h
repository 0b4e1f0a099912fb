/* pda_b200.h -- C-ABI of the B200-native residual/Jacobian engine for pressio-demoapps' hot path.
 *
 * The reference (Pressio/pressio-demoapps, header-only C++17) has no FFI of its own; this header IS the seam a
 * maintainer binds (INTEGRATION.md shows the C++ shim and the pybind11/ctypes stub).  Every entry point cites the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - opaque handles, plain pointers + sizes, no C++/torch types;
 *  - every function returns a pda_status (0 = ok); pda_last_error() gives the message of the last failure on the
 *    calling thread.  The reference throws std::runtime_error / calls exit(); nothing here throws or exits;
 *  - all floating point is IEEE double, all indices int32_t, exactly like the reference
 *    (include/pressiodemoapps/mesh.hpp:76-81, impl/euler_3d_prob_class.hpp:76-80);
 *  - state / velocity layout: AoS  U[cell*ndpc + dof]  (SURVEY App. A);
 *  - "_host" flavours take host pointers (staged through pinned buffers, H2D/D2H inside the call);
 *    "_dev" flavours take device pointers and a cudaStream_t (passed as void*; NULL = the legacy default stream),
 *    enqueue their kernels on that stream and do not synchronise;
 *  - there is NO CPU fallback: evaluation entry points fail with PDA_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef PDA_B200_H_
#define PDA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pda_mesh_s*    pda_mesh;
typedef struct pda_problem_s* pda_problem;
typedef struct pda_gradient_s* pda_gradient;
typedef int pda_status;

enum {
  PDA_OK = 0,
  PDA_ERR_INVALID = 1,      /* invalid argument / enum / parameter name (reference: std::runtime_error)        */
  PDA_ERR_IO = 2,           /* missing or malformed mesh file (reference: exit(EXIT_FAILURE))                   */
  PDA_ERR_NO_DEVICE = 3,    /* no usable CUDA device: the product path never falls back to the CPU              */
  PDA_ERR_CUDA = 4,         /* a CUDA runtime call failed                                                       */
  PDA_ERR_UNSUPPORTED = 5,  /* combination the engine (or the reference) does not implement                     */
  PDA_ERR_TOO_LARGE = 6     /* nnz or dof count does not fit the reference's int32 index type                   */
};

/* problem families = the reference's problem enum TYPES */
enum {
  PDA_FAMILY_EULER1D = 1,              /* euler1d.hpp:64-69                  */
  PDA_FAMILY_EULER2D = 2,              /* euler2d.hpp:66-76                  */
  PDA_FAMILY_EULER3D = 3,              /* euler3d.hpp:65-68                  */
  PDA_FAMILY_SWE2D = 4,                /* swe2d.hpp:65-68                    */
  PDA_FAMILY_DIFFUSION_REACTION2D = 5, /* diffusion_reaction2d.hpp           */
  PDA_FAMILY_ADVECTION_DIFFUSION2D = 6, /* advection_diffusion2d.hpp:66-69 (Burgers) */
  PDA_FAMILY_ADVECTION_DIFFUSION_REACTION2D = 7, /* advection_diffusion_reaction2d.hpp:60-62 */
  PDA_FAMILY_ADVECTION1D = 8,          /* advection1d.hpp:64-66              */
  PDA_FAMILY_DIFFUSION_REACTION1D = 9  /* diffusion_reaction1d.hpp:64-67     */
};
/* problem ids inside a family keep the reference's enumerator order */
enum { PDA_EULER1D_PERIODIC_SMOOTH = 0, PDA_EULER1D_SOD = 1, PDA_EULER1D_LAX = 2, PDA_EULER1D_SHU_OSHER = 3 };
enum {
  PDA_EULER2D_PERIODIC_SMOOTH = 0, PDA_EULER2D_KELVIN_HELMHOLTZ = 1, PDA_EULER2D_SEDOV_FULL = 2,
  PDA_EULER2D_SEDOV_SYMMETRY = 3, PDA_EULER2D_RIEMANN = 4, PDA_EULER2D_NORMAL_SHOCK = 5,
  PDA_EULER2D_DOUBLE_MACH_REFLECTION = 6, PDA_EULER2D_CROSS_SHOCK = 7, PDA_EULER2D_TESTING_ONLY_NEUMANN = 8
};
enum { PDA_EULER3D_PERIODIC_SMOOTH = 0, PDA_EULER3D_SEDOV_SYMMETRY = 1 };
enum { PDA_SWE2D_SLIP_WALL = 0, PDA_SWE2D_CUSTOM_BCS = 1 };
enum { PDA_DIFFREAC2D_PROBLEM_A = 0, PDA_DIFFREAC2D_GRAY_SCOTT = 1 };
enum { PDA_ADVDIFF2D_BURGERS_PERIODIC = 0, PDA_ADVDIFF2D_BURGERS_OUTFLOW = 1 };
enum { PDA_ADVDIFFREAC2D_PROBLEM_A = 0 };
enum { PDA_ADVECTION1D_PERIODIC_LINEAR = 0 };
enum { PDA_DIFFREAC1D_PROBLEM_A = 0 };
/* InviscidFluxReconstruction (schemes_info.hpp:57-66) */
enum { PDA_FIRST_ORDER = 0, PDA_WENO3 = 1, PDA_WENO5 = 2 };
/* ghost sides, in the reference's graph-column order (GhostRelativeLocation, ghost_relative_locations.hpp) */
enum { PDA_SIDE_LEFT = 0, PDA_SIDE_FRONT = 1, PDA_SIDE_RIGHT = 2, PDA_SIDE_BACK = 3, PDA_SIDE_BOTTOM = 4, PDA_SIDE_TOP = 5 };
/* device-expressible custom boundary conditions (custom_bc_holder.hpp / custom_bcs_functions.hpp) */
enum { PDA_BC_DIRICHLET = 0, PDA_BC_HOMOG_NEUMANN = 1, PDA_BC_REFLECTIVE = 2, PDA_BC_HOST_CALLBACK = 3 };
/* operand layout for apply_jacobian (adapter_cpp.hpp:231-259: vector, col-major, row-major) */
enum { PDA_LAYOUT_COL_MAJOR = 0, PDA_LAYOUT_ROW_MAJOR = 1 };

const char* pda_last_error(void);
const char* pda_version(void);
/* number of CUDA devices the library can use (0 on a CPU-only box; never an error) */
int pda_device_count(void);

/* measurement helper (not on the evaluation path): peak FP64 FMA throughput of `device` in TFLOP/s from a pure DFMA
 * loop -- the FP64 roofline denominator bench.py reports beside the HBM one (MEASURED_PEAKS.json has no FP64 entry) */
pda_status pda_measure_fp64_peak(int device, double* tflops, double* sm_mhz_hint);
/* same probe with its own clock evidence: the SM clock the probe actually ran at (SM cycles from clock64 over elapsed
 * nanoseconds from globaltimer, inside the kernel) and the DFMA issue rate per SM and cycle that results (the pipe's
 * nominal rate is 64 per SM and cycle) -- separates "the clock dropped under FP64 load" from "the pipe cannot be fed" */
pda_status pda_measure_fp64_peak_ex(int device, double* tflops, double* sm_mhz_under_probe, double* dfma_per_sm_clk);
/* test hook of the reference-order mode: out[i] = the device restatement of glibc's pow(x[i], y) (csrc/glibc_pow.h);
 * host pointers, synchronous.  tests/ compare it bit for bit with the host libm. */
pda_status pda_test_glibc_pow(int device, const double* x, double y, double* out, int64_t n);

/* ------------------------------------------------------------------ mesh ---------------------------------------- */
/* load_cellcentered_uniform_mesh_eigen(dir)  (mesh.hpp:87-91, impl/mesh_ccu.hpp:359-448): reads info.dat,
 * coordinates.dat, connectivity.dat written by meshing_scripts/create_{full,sample}_mesh.py */
pda_status pda_mesh_load(const char* dir, pda_mesh* out);
/* native replacement of meshing_scripts/create_full_mesh.py (natural row ordering): n[3] cells,
 * bounds = {xmin,xmax,ymin,ymax,zmin,zmax}, periodic[3] flags, stencil in {3,5,7}.  dim is 1/2/3 (unused axes n=1).
 * 3D stencil 7 is supported (extension).  Nothing O(cells) is materialised until asked for. */
pda_status pda_mesh_make_lattice(int dim, const int32_t n[3], const double bounds[6], const int32_t periodic[3],
                                 int stencil, pda_mesh* out);
/* native replacement of meshing_scripts/create_sample_mesh.py: sample cells = gids (any order, sorted internally)
 * of `full`; stencil mesh = sample cells + their stencil neighbours, renumbered by ascending full-mesh gid. */
pda_status pda_mesh_make_sample(pda_mesh full, const int32_t* gids, int64_t ngids, pda_mesh* out);
/* Shard `rank` of `nranks` of a full lattice cut into slabs along its slowest axis ("each GPU owns its slab of
 * U, V and the corresponding row block of J; physical-boundary GPUs fill ghosts locally").  The result is an ordinary
 * mesh in the sense of create_sample_mesh.py -- sample cells = the owned cells, stencil cells = owned planes plus
 * (stencil-1)/2 HALO planes per side wherever a neighbouring rank (or the periodic image) owns them, none at a physical
 * boundary -- numbered plane after plane: [lower halo | owned | upper halo].  Every problem family, boundary condition,
 * the Jacobian (rows = owned dofs, LOCAL column ids) and applyJacobian work on it through the usual entry points; the
 * caller keeps the halo planes of the state (and of an applyJacobian operand) current by exchanging them with the ring
 * neighbours (pressiodemoapps.sharded: NCCL send/recv between processes, peer copies inside one process).
 * Any lattice, periodic or not, any nranks with at least (stencil-1)/2 planes per rank.
 * info = {plane_cells, k0, k1, halo_planes_below, halo_planes_above, rank, nranks, dim}. */
pda_status pda_mesh_make_slab_window(pda_mesh full, int rank, int nranks, pda_mesh* out);
pda_status pda_mesh_slab_window_info(pda_mesh m, int64_t info[8]);
/* mesh from caller arrays (graph row-major [nSample][(stencil-1)*dim+1]) */
pda_status pda_mesh_from_arrays(int dim, int stencil, int32_t nSample, int32_t nStencil, const double dxyz[3],
                                const double* x, const double* y, const double* z, const int32_t* graph,
                                pda_mesh* out);
/* write info.dat / connectivity.dat / coordinates.dat (+ stencil_mesh_gids.dat for sample meshes) byte-compatible
 * with the reference's scripts (create_full_mesh.py:151-218, create_sample_mesh.py:160-212) */
pda_status pda_mesh_write(pda_mesh m, const char* dir);
pda_status pda_mesh_free(pda_mesh m);

/* getters = CellCenteredUniformMesh accessors (impl/mesh_ccu.hpp:115-159; python: src_py/main_binder.cc:230-243) */
int     pda_mesh_dimensionality(pda_mesh m);
int     pda_mesh_stencil_size(pda_mesh m);
int32_t pda_mesh_stencil_mesh_size(pda_mesh m);
int32_t pda_mesh_sample_mesh_size(pda_mesh m);
int     pda_mesh_graph_cols(pda_mesh m);
int     pda_mesh_is_fully_periodic(pda_mesh m);
int     pda_mesh_is_lattice(pda_mesh m);
int32_t pda_mesh_num_cells_inner(pda_mesh m);
int32_t pda_mesh_num_cells_near_bd(pda_mesh m);
pda_status pda_mesh_deltas(pda_mesh m, double dxyz[3], double dxyz_inv[3]);             /* dx() .. dzInv()       */
pda_status pda_mesh_coordinates(pda_mesh m, double* x, double* y, double* z);           /* viewX/Y/Z (nStencil)  */
pda_status pda_mesh_graph(pda_mesh m, int32_t* graph);                                  /* graph()               */
pda_status pda_mesh_rows_inner(pda_mesh m, int32_t* rows);                   /* graphRowsOfCellsAwayFromBd()     */
pda_status pda_mesh_rows_near_bd(pda_mesh m, int32_t* rows);                 /* graphRowsOfCellsNearBd()         */
/* graphRowsOfCellsStrictlyOnBd() (impl/mesh_ccu.hpp:153-155, filled at :441-447 for 2D meshes only): near-boundary
 * rows whose cell has at least one face on the domain boundary; count = -1 for a null handle, 0 for 1D / 3D meshes */
int32_t pda_mesh_num_cells_strictly_on_bd(pda_mesh m);
pda_status pda_mesh_rows_strictly_on_bd(pda_mesh m, int32_t* rows);
/* sample mesh only: full-mesh gid of every stencil-mesh cell (stencil_mesh_gids.dat) */
pda_status pda_mesh_stencil_gids(pda_mesh m, int32_t* gids);

/* ------------------------------------------------------------------ problem ------------------------------------- */
/* create_problem_eigen(mesh, <enum>, recon[, icFlag][, {name:value}])   (euler1d.hpp:82-99, euler2d.hpp:88-186,
 * euler3d.hpp:80-122, swe2d.hpp:134-185, diffusion_reaction2d.hpp:259-285 with names Du,Dv,F,k).
 * The mesh must outlive the problem (reference: std::reference_wrapper, euler_2d_prob_class.hpp:1277).
 * Unlike the reference, an incompatible mesh stencil / scheme pair is rejected (schemes_info.hpp:94-102). */
pda_status pda_problem_create(pda_mesh mesh, int family, int problem_id, int recon, int ic_flag, int nparams,
                              const char* const* names, const double* values, int device, pda_problem* out);
/* Parameter names per family ({name: value} map of the reference, or its positional factory arguments):
 *   Euler2d   gamma + impl/euler_2d_parametrization_helpers.hpp names;  Swe2d  gravity, coriolis, pulse*;
 *   GrayScott Du, Dv, F, k;  DiffusionReaction{1d,2d}::ProblemA  diffusion, reaction (create_diffusion_reaction_*_problem_A_eigen);
 *   AdvectionDiffusion2d  diffusion, pulseMagnitude, pulseSpread, pulseX, pulseY (advection_diffusion_2d_parametrization_helpers.hpp);
 *   AdvectionDiffusionReaction2d  ux, uy, diffusion, sigma (create_advdiffreac_2d_problem_A_eigen);
 *   Advection1d  velocity (create_linear_advection_1d_problem_eigen), ic_flag 1..4. */

/* Source term of the ProblemA families (the reference takes a host functor f(x[,y],t), diffusion_reaction1d.hpp:70-79,
 * diffusion_reaction2d.hpp:228-237, advection_diffusion_reaction2d.hpp:70-75): the binding evaluates its functor at
 * the current time for every SAMPLE cell and hands over the table (sample_mesh_size doubles, host pointer).  Without
 * a call the reference's default functors are tabulated by the library. */
pda_status pda_problem_set_source(pda_problem p, const double* values);

/* Engine options (no counterpart in the reference: they select between implementations of the SAME interface).
 *   "jacobian_order" = "fast" (default) | "reference"
 *     fast:      face-sharing Jacobian kernels with the well-conditioned form of the WENO reconstruction gradients.
 *                Values agree with the reference to the rounding noise of the reference's OWN gradient formula
 *                (impl/weno5.hpp:180-434 cancels catastrophically: its values move by up to 180x the 1e-12/1e-10
 *                tolerance under FMA contraction alone) and sit closer to the exact Jacobian than the reference's.
 *     reference: every row through one-thread-per-row kernels that keep the reference's formulas, operation order
 *                and accumulation order, each operation individually rounded and std::pow reproduced bit for bit
 *                (csrc/kernels_reforder.cu): velocity and Jacobian values within 1e-12 relative / 1e-10 absolute of
 *                the reference on every golden fixture (in practice identical to the last bit).  Slow by design
 *                (read-modify-write like Eigen's coeffRef); applyJacobian then multiplies that assembled Jacobian. */
pda_status pda_problem_set_option(pda_problem p, const char* name, const char* value);
pda_status pda_problem_get_option(pda_problem p, const char* name, char* value, int capacity);

/* custom BCs (Swe2d::CustomBCs, Euler2d Riemann/NormalShock and AdvectionDiffusion2d custom-BC overloads): one device-expressible rule per side.
 * values: ndpc doubles (Dirichlet ghost state); ignored otherwise.  (custom_bcs_functions.hpp:60-164) */
pda_status pda_problem_set_bc(pda_problem p, int side, int kind, const double* values);
/* Arbitrary HOST functors for one side -- the reference's custom-BC functor contract, one-to-one
 * (custom_bcs_functions.hpp:107-164 ghost fill, :60-103 Jacobian factors; tests_cpp/eigen_2d_swe_custom_bcs/main.cc:6-58):
 *   ghost(user, nearBdRowId, graphRow, cellX, cellY, U (whole state, host copy), ndpc, cellWidth, ghostValues)
 *   factors(user, graphRow, cellX, cellY, ndpc, factors)            (may be NULL: factors stay 1)
 * This is the SLOW path kept for generality: every evaluation copies the state to the host, runs the functor for the
 * boundary cells of that side and uploads the ghost rows.  Device-expressible rules (pda_problem_set_bc) never leave HBM. */
typedef void (*pda_bc_ghost_fn)(void* user, int32_t near_bd_row, const int32_t* graph_row, double cell_x, double cell_y,
                                const double* U, int ndpc, double cell_width, double* ghost_values);
typedef void (*pda_bc_factor_fn)(void* user, const int32_t* graph_row, double cell_x, double cell_y, int ndpc,
                                 double* factors);
pda_status pda_problem_set_bc_callback(pda_problem p, int side, pda_bc_ghost_fn ghost, pda_bc_factor_fn factors, void* user);
/* setBCPointer(GhostRelativeLocation, ptr) (euler_2d_prob_class.hpp:213-216, swe_2d_prob_class.hpp:224-227,
 * advection_diffusion_2d_prob_class.hpp:199-202 -> custom_bc_holder.hpp:89-103 -> the functor's setInternalPtr): re-points
 * the state handed to the host functors of `side` (their `user` argument), e.g. at a neighbouring subdomain's state in
 * a Schwarz iteration.  PDA_ERR_INVALID when no host functor is installed on that side. */
pda_status pda_problem_set_bc_pointer(pda_problem p, int side, void* user);
pda_status pda_problem_free(pda_problem p);

int     pda_problem_num_dof_per_cell(pda_problem p);       /* numDofPerCell()        adapter_cpp.hpp:93-95   */
int32_t pda_problem_total_dof_sample_mesh(pda_problem p);  /* totalDofSampleMesh()   adapter_cpp.hpp:97-99   */
int32_t pda_problem_total_dof_stencil_mesh(pda_problem p); /* totalDofStencilMesh()  adapter_cpp.hpp:101-103 */
/* gamma() / gravity() / coriolis() / queryParameter(name) */
pda_status pda_problem_query_parameter(pda_problem p, const char* name, double* value);
/* initialCondition()  (host buffer, totalDofStencilMesh doubles) */
pda_status pda_problem_initial_condition(pda_problem p, double* U);

/* createJacobian(): fixed CSR pattern, bit-identical to Eigen's setFromTriplets+makeCompressed
 * (euler_2d_prob_class.hpp:223-237).  nnz first, then the arrays (rowptr: nrows+1, colidx: nnz). */
pda_status pda_problem_jacobian_nnz(pda_problem p, int64_t* nnz);
pda_status pda_problem_jacobian_pattern(pda_problem p, int32_t* rowptr, int32_t* colidx);

/* rightHandSide / operator()(U,t,V)                      adapter_cpp.hpp:162-199 */
pda_status pda_problem_velocity_host(pda_problem p, const double* U, double t, double* V);
/* rightHandSideAndJacobian / operator()(U,t,V,J,true)    adapter_cpp.hpp:170-213 ; V may be NULL (jacobian()) */
pda_status pda_problem_velocity_and_jacobian_host(pda_problem p, const double* U, double t, double* V,
                                                  double* jac_values);
/* applyJacobian(U,B,t,R): R = J(U,t) * B, B is [totalDofStencilMesh x ncols]   adapter_cpp.hpp:231-259 */
pda_status pda_problem_apply_jacobian_host(pda_problem p, const double* U, const double* B, int ncols, int layout,
                                           double t, double* R);

/* device-pointer flavours (inputs already resident in HBM; asynchronous on `stream`) */
pda_status pda_problem_velocity_dev(pda_problem p, const double* dU, double t, double* dV, void* stream);
pda_status pda_problem_velocity_and_jacobian_dev(pda_problem p, const double* dU, double t, double* dV,
                                                 double* dJvalues, void* stream);
pda_status pda_problem_apply_jacobian_dev(pda_problem p, const double* dU, const double* dB, int ncols, int layout,
                                          double t, double* dR, void* stream);

/* Explicit time stepping with the state resident in HBM: `nsteps` steps of size dt from t0, U updated in place, no
 * host round trip between evaluations.  Stage arithmetic of the steppers the reference's tests drive the problems with
 * (tests_cpp/pressio/include/pressio/ode/impl/ode_explicit_stepper_without_mass_matrix.hpp: ForwardEuler :168-186,
 * SSPRungeKutta3 :230-281, RungeKutta4 :284-340) and of the Python module's advanceRK2 (pressiodemoapps/__init__.py:90-113).
 * Full meshes only (sample mesh == stencil mesh). */
enum { PDA_STEPPER_FORWARD_EULER = 0, PDA_STEPPER_RK2 = 1, PDA_STEPPER_RK4 = 2, PDA_STEPPER_SSPRK3 = 3 };
pda_status pda_problem_advance_dev(pda_problem p, int stepper, double* dU, double t0, double dt, int32_t nsteps, void* stream);
pda_status pda_problem_advance_host(pda_problem p, int stepper, double* U, double t0, double dt, int32_t nsteps);

/* test hooks = viewGhostLeft/Front/Right/Back (euler_2d_prob_class.hpp:205-210): ghost rows after the last
 * evaluation, [numCellsNearBd][ndpc*(stencil-1)/2] */
pda_status pda_problem_ghosts(pda_problem p, int side, double* out);

/* number of kernel launches issued by this problem since creation (bench.py's gpu_launches) */
int64_t pda_problem_launch_count(pda_problem p);
/* device time of the dominant kernel in the last *_dev/_host evaluation is not measured here: callers bracket
 * with their own CUDA events on `stream`. */

/* ---------------------------------------------- slab-decomposed lattices (multi-GPU) ---------------------------- */
/* One process per GPU.  A rank owns planes [k0,k1) of the slowest axis of a full lattice; its local
 * state carries `halo` = (stencil-1)/2 extra planes on each side which the caller fills (NCCL send/recv or peer
 * copies) before the boundary planes are evaluated.  Local state layout: [(k1-k0)+2*halo planes][plane cells][ndpc]. */
pda_status pda_problem_create_slab(pda_mesh lattice, int family, int problem_id, int recon, int rank, int nranks,
                                   int device, pda_problem* out);
pda_status pda_slab_extent(pda_problem p, int32_t* k0, int32_t* k1, int32_t* halo, int64_t* plane_dofs);
/* initial condition of the owned planes (no halos), host buffer of (k1-k0)*plane_dofs doubles */
pda_status pda_slab_initial_condition(pda_problem p, double* U_owned);
/* velocity of the owned planes that do NOT need halo data (interior), then of the 2*halo boundary planes */
pda_status pda_slab_velocity_interior_dev(pda_problem p, const double* dU_local, double t, double* dV_owned, void* stream);
pda_status pda_slab_velocity_boundary_dev(pda_problem p, const double* dU_local, double t, double* dV_owned, void* stream);


/* Peer mode (3D lattices): the halo exchange fused with the evaluation over NVLink peer memory, no collective library
 * on the data path.  Every rank owns a halo buffer (2 parities x {lower, upper} x halo planes + flags) that its ring
 * neighbours fill with copy-engine peer copies followed by a flag write; ONE kernel launch evaluates all owned planes
 * and only the CTAs that touch a halo plane wait for the flag, after the interior chunks have been scheduled.
 *   1. every rank:  pda_slab_peer_handle()  -> 64-byte cudaIpcMemHandle of its buffer;
 *   2. all-gather the handles (torch.distributed / MPI / files), then pda_slab_peer_connect(all handles by rank);
 *      neighbours living in the SAME process connect with pda_slab_peer_connect_local instead;
 *   3. per evaluation: pda_slab_velocity_peer_dev(dU_owned, t, dV_owned, stream) -- dU_owned holds the owned planes
 *      only, (k1-k0)*plane_dofs doubles.  Every rank must call it the same number of times (SPMD); work enqueued on
 *      `stream` after the call (e.g. the update of U) is ordered after the outgoing copies.  Ranks sharing one
 *      process must use distinct streams. */
pda_status pda_slab_peer_handle(pda_problem p, unsigned char handle[64]);
pda_status pda_slab_peer_connect(pda_problem p, const unsigned char* handles /* nranks x 64 bytes */);
pda_status pda_slab_peer_connect_local(pda_problem p, pda_problem lower, pda_problem upper);
pda_status pda_slab_velocity_peer_dev(pda_problem p, const double* dU_owned, double t, double* dV_owned, void* stream);
/* host-pointer flavour (owned planes only, pinned buffers recommended): chunks of planes flow H2D -> kernel -> D2H on
 * three streams; the boundary chunks go first so the pushes to the neighbours overlap the interior uploads */
pda_status pda_slab_velocity_peer_host(pda_problem p, const double* U_owned, double t, double* V_owned);

/* ------------------------------------------------------------------ boundary-face gradients --------------------- */
/* GradientEvaluator<Mesh, MaxNumDofPerCell>(mesh)  (gradient.hpp:61-121, impl/gradient_2d.hpp:141-293): normal
 * gradient of a cell-centred field at every face on the domain boundary, by one-sided finite differences whose width
 * follows the mesh stencil (two points for stencil 3, three points for 5 and 7: gradient_2d.hpp:179-196).  2D only
 * (PDA_ERR_UNSUPPORTED with the reference's message otherwise, gradient.hpp:71-73).  The mesh must outlive the
 * evaluator only during this call: the face list is copied.
 * Faces are numbered in creation order: rows of graphRowsOfCellsStrictlyOnBd(), then Left, Front, Right, Back
 * (gradient_2d.hpp:243-284; the reference keeps them in an unordered_map keyed by (cellGID, FacePosition)). */
pda_status pda_gradient_create(pda_mesh mesh, int max_num_dof_per_cell, pda_gradient* out);
pda_status pda_gradient_free(pda_gradient g);
int32_t pda_gradient_num_faces(pda_gradient g);
/* the Face records (gradient_2d.hpp:109-130); any output may be NULL.  position = FacePosition (schemes_info.hpp:112-114:
 * 0 Left, 1 Front, 2 Right, 3 Back), normal_direction 1 = x, 2 = y, centers = centerCoordinates [num_faces][3] */
pda_status pda_gradient_faces(pda_gradient g, int32_t* cell_gid, int32_t* position, int32_t* parent_graph_row,
                              int32_t* normal_direction, double* centers);
/* queryFace(cellGID, FacePosition) (gradient.hpp:115-117): index of that face; PDA_ERR_INVALID when the mesh has no
 * such boundary face (the reference asserts) */
pda_status pda_gradient_query_face(pda_gradient g, int32_t cell_gid, int position, int32_t* face_index);
/* operator()(field[, numDofPerCell]) (gradient.hpp:77-93): field = [stencilMeshSize][num_dof_per_cell] row-major,
 * normal_grad = [num_faces][num_dof_per_cell].  num_dof_per_cell > max_num_dof_per_cell is refused with the
 * reference's message.  Host-pointer (staged copies, synchronous) and device-pointer + stream flavours; no CPU path. */
pda_status pda_gradient_compute_host(pda_gradient g, const double* field, int num_dof_per_cell, double* normal_grad);
pda_status pda_gradient_compute_dev(pda_gradient g, const double* d_field, int num_dof_per_cell, double* d_normal_grad,
                                    void* stream);
int64_t pda_gradient_launch_count(pda_gradient g);

#ifdef __cplusplus
}
#endif
#endif /* PDA_B200_H_ */
