"""Slab-sharded evaluation over the GPUs of one box: any full lattice -- periodic or not -- any problem
family, velocity, Jacobian row blocks, applyJacobian and device-resident explicit time stepping.

One shard = one rank's slab window (pda_mesh_make_slab_window): an ordinary problem whose state vector carries the
owned planes plus (stencil-1)/2 halo planes per side where a neighbour rank owns them,
        local state = [ lower halo | owned planes | upper halo ]          (plane after plane, AoS dofs)
Physical-boundary ranks have no halo on that side: their near-boundary rows use ghost cells exactly like a single-GPU
problem.  The ONLY exchange step of the path is the halo refresh of the state (and of an applyJacobian operand)
before an evaluation; there is no reduction, so no all-reduce is invented:

  * between processes (one process per GPU, torchrun): `exchange_halos` = NCCL send/recv with the two ring neighbours
    (torch.distributed.batch_isend_irecv on the contiguous plane ranges); the same function runs on gloo with CPU
    tensors, which is how tests/ cover the indexing without a GPU;
  * inside one process (several GPUs, or several shards on one GPU in tests): `exchange_halos_local` = peer copies.

The Jacobian of a shard has rows = owned dofs and LOCAL column ids (positions in the local state vector); with
`Shard.global_columns()` a caller maps them back to full-mesh dof ids.

The periodic 3D velocity path keeps its faster special case (create_problem_slab + peer mode: halo pushes fused with
the kernel); this module is the general one.
"""
import numpy as _np

from . import (create_problem, create_slab_window_mesh, _FAMILY, _make)  # noqa: F401


class Shard:
    """one rank's slab of a sharded problem"""

    def __init__(self, fullMesh, probEnum, *args, rank=0, nranks=1, device=0):
        self.mesh = create_slab_window_mesh(fullMesh, rank, nranks)
        self.problem = create_problem(self.mesh, probEnum, *args, device=device)
        w = self.mesh.window
        self.rank, self.nranks = w["rank"], w["nranks"]
        self.ndpc = self.problem.numDofPerCell()
        self.plane_dofs = w["plane_cells"] * self.ndpc
        self.h_lo, self.h_hi = w["halo_lo"], w["halo_hi"]
        self.n_owned_planes = w["k1"] - w["k0"]
        self.k0, self.k1 = w["k0"], w["k1"]
        self.h = (self.mesh.stencilSize() - 1) // 2
        # ring neighbours (None at a physical boundary)
        self.lower = (self.rank - 1) % self.nranks if self.h_lo else None
        self.upper = (self.rank + 1) % self.nranks if self.h_hi else None

    # ---- dof ranges inside the local state vector
    def owned(self):
        a = self.h_lo * self.plane_dofs
        return slice(a, a + self.n_owned_planes * self.plane_dofs)

    def recv_lower(self):
        return slice(0, self.h_lo * self.plane_dofs)

    def recv_upper(self):
        a = (self.h_lo + self.n_owned_planes) * self.plane_dofs
        return slice(a, a + self.h_hi * self.plane_dofs)

    def send_lower(self):
        """my bottom planes: what the LOWER neighbour needs as its upper halo"""
        a = self.h_lo * self.plane_dofs
        return slice(a, a + self.h * self.plane_dofs)

    def send_upper(self):
        """my top planes: what the UPPER neighbour needs as its lower halo"""
        b = (self.h_lo + self.n_owned_planes) * self.plane_dofs
        return slice(b - self.h * self.plane_dofs, b)

    def local_size(self):
        return self.problem.totalDofStencilMesh()

    def owned_size(self):
        return self.problem.totalDofSampleMesh()

    def global_columns(self):
        """full-mesh dof id of every local dof (column map of the shard's Jacobian)"""
        g = self.mesh.stencilMeshGids().astype(_np.int64)
        return (g[:, None] * self.ndpc + _np.arange(self.ndpc)[None, :]).ravel()

    def global_rows(self):
        """full-mesh dof id of every owned dof (row map of the shard's velocity / Jacobian)"""
        first = self.k0 * self.plane_dofs
        return _np.arange(first, first + self.owned_size(), dtype=_np.int64)

    def scatter_from_full(self, Ufull):
        """local state (owned + halo planes) cut out of a full-mesh state (numpy) -- initial data, tests"""
        return _np.ascontiguousarray(_np.asarray(Ufull)[self.global_columns()])


def _rows(t, sl, ncols_major):
    """view of the dof range `sl` of a state vector / row-major operand (first axis = dofs)"""
    return t[sl]


def exchange_halos(shard, U, dist, group=None, tag_base=0):
    """refresh the halo planes of the local state `U` (1-D tensor, or [local dofs, ncols] row-major operand) from the
    ring neighbours: my bottom planes -> lower neighbour's upper halo, my top planes -> upper neighbour's lower halo.
    torch.distributed send/recv (NCCL on GPUs, gloo on CPU tensors); returns after the receives have completed."""
    ops = []
    if shard.upper is not None:
        ops.append(dist.P2POp(dist.isend, U[shard.send_upper()], shard.upper, group))
        ops.append(dist.P2POp(dist.irecv, U[shard.recv_upper()], shard.upper, group))
    if shard.lower is not None:
        ops.append(dist.P2POp(dist.isend, U[shard.send_lower()], shard.lower, group))
        ops.append(dist.P2POp(dist.irecv, U[shard.recv_lower()], shard.lower, group))
    if not ops:
        return
    if shard.nranks == 2 and shard.lower == shard.upper and shard.lower is not None:
        # two ranks on a periodic axis: both halos come from the same peer; order the pairs so that the two sends of one
        # rank meet the two receives of the other in the same order (my upper <-> its lower first on rank 0, mirrored on 1)
        if shard.rank == 1:
            ops = ops[2:] + ops[:2]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def exchange_halos_local(shards, Us):
    """same exchange for shards living in ONE process (tests; single-process multi-GPU): plain tensor copies, which are
    peer copies over NVLink when the shards sit on different devices"""
    for s, U in zip(shards, Us):
        if s.upper is not None:
            U[s.recv_upper()].copy_(Us[s.upper][shards[s.upper].send_lower()])
        if s.lower is not None:
            U[s.recv_lower()].copy_(Us[s.lower][shards[s.lower].send_upper()])


class ShardedStepper:
    """device-resident explicit time stepping over shards: the state never leaves HBM; one halo refresh per stage.
    Stage arithmetic of pda_problem_advance_* (forward Euler, RK4, SSPRK3: the steppers the reference's tests use)."""

    def __init__(self, shard, torch, exchange):
        self.s, self.torch, self.exchange = shard, torch, exchange   # exchange(U): refresh the halos of U in place

    def _f(self, U, t, out):
        self.exchange(U)
        st = self.torch.cuda.current_stream().cuda_stream
        self.s.problem.rightHandSideDevice(U.data_ptr(), float(t), out.data_ptr(), st)

    def advance(self, stepper, U, dt, nsteps, t0=0.0):
        torch, s = self.torch, self.s
        own = s.owned()
        k = [torch.empty(s.owned_size(), dtype=torch.float64, device=U.device) for _ in range(4 if stepper == "rk4" else 1)]
        aux = U.clone()
        t = t0
        for _ in range(nsteps):
            if stepper == "euler":
                self._f(U, t, k[0])
                U[own] += dt * k[0]
            elif stepper == "rk4":
                half = dt / 2.0
                self._f(U, t, k[0])
                aux[own] = U[own] + half * k[0]
                self._f(aux, t + half, k[1])
                aux[own] = U[own] + half * k[1]
                self._f(aux, t + half, k[2])
                aux[own] = U[own] + dt * k[2]
                self._f(aux, t + dt, k[3])
                U[own] = U[own] + (dt / 6.0) * k[0] + (dt / 3.0) * k[1] + (dt / 3.0) * k[2] + (dt / 6.0) * k[3]
            elif stepper == "ssprk3":
                self._f(U, t, k[0])
                aux[own] = U[own] + dt * k[0]
                self._f(aux, t + dt, k[0])
                aux[own] = 0.25 * aux[own] + 0.75 * U[own] + (0.25 * dt) * k[0]
                self._f(aux, t + dt / 2.0, k[0])
                U[own] = (1.0 / 3.0) * U[own] + (2.0 / 3.0) * aux[own] + ((2.0 / 3.0) * dt) * k[0]
            else:
                raise ValueError("stepper must be euler, rk4 or ssprk3")
            t += dt
        self.exchange(U)
        return U
