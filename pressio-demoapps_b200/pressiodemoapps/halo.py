"""Slab decomposition plumbing (one process per GPU): which planes a rank owns and the exchange of the
stencil-width halo planes with the two ring neighbours through torch.distributed (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  The reference has no distributed layer; this is the new part of the engine.

Local state layout of a rank (matches pda_slab_* in include/pda_b200.h):
    [ halo planes from the previous rank | owned planes k0..k1 | halo planes from the next rank ]
each plane being `plane_dofs` contiguous doubles, so every message is ONE contiguous range: no pack kernel.
"""
import torch
import torch.distributed as dist


def slab_range(nplanes, rank, world):
    """planes [k0,k1) owned by `rank` (the slowest lattice axis must divide evenly, like pda_problem_create_slab)"""
    if nplanes % world:
        raise ValueError("slab: %d planes do not divide over %d ranks" % (nplanes, world))
    per = nplanes // world
    return rank * per, (rank + 1) * per


def post_halo_exchange(local, halo, plane_dofs, rank, world, group=None):
    """Start the periodic halo exchange of `local` (1-D tensor, layout above) and return the list of work handles.
    Sends: my first `halo` owned planes -> previous rank's upper halo; my last `halo` owned planes -> next rank's
    lower halo.  With world == 1 the wrap is a local copy."""
    n = halo * plane_dofs
    total = local.numel()
    lo_halo = local[:n]
    hi_halo = local[total - n:]
    first = local[n:2 * n]
    last = local[total - 2 * n: total - n]
    if world == 1:
        lo_halo.copy_(last)
        hi_halo.copy_(first)
        return []
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    # tags keep the two messages apart when prev == nxt (world == 2)
    ops = [dist.P2POp(dist.isend, first, prev, group, 0), dist.P2POp(dist.isend, last, nxt, group, 1),
           dist.P2POp(dist.irecv, hi_halo, nxt, group, 0), dist.P2POp(dist.irecv, lo_halo, prev, group, 1)]
    return dist.batch_isend_irecv(ops)


def wait_all(works):
    for w in works:
        w.wait()


def connect_peer_halo(problem, rank, world, group=None):
    """Peer mode (include/pda_b200.h): all-gather the ranks' 64-byte IPC handles of their halo buffers and map the two
    ring neighbours' buffers.  After this, `problem.slabVelocityPeerDevice` needs no collective on the data path."""
    mine = problem.peerHandle()
    if world == 1:
        problem.peerConnect([mine])
        return
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    problem.peerConnect(handles)
