"""Slab decomposition over several GPUs: the host-side mirror of `MultiArch` and of the distributed
`RectilinearGrid` constructor (reference src/Distributed/multi_architectures.jl:7-137,
distributed_grids.jl:16-60).  One process per GPU; y is split over `R` ranks, `ranks = (1, R, 1)`
(z is never split in the reference either: distributed_fft_based_poisson_solver.jl:101-102).
The NCCL communicator lives inside libocean_b200.so (ob200_comm_*); this module only distributes
the NCCL unique id with whatever process group the host already has (torch.distributed here,
MPI.bcast in the Julia shim)."""
import ctypes as C

from ._lib import lib, check
from .grids import B200


def local_size(global_size, ranks, dim=1):
    """distributed_grids.jl:36-38: the global size must divide evenly."""
    n, r = global_size[dim], ranks[dim]
    if n % r:
        raise ValueError(f"global size {n} along dimension {dim} is not divisible by {r} ranks")
    out = list(global_size)
    out[dim] = n // r
    return tuple(out)


def local_interval(interval, R, index):
    """distributed_grids.jl:40-46: equal sub-intervals, index is 0-based."""
    a, b = float(interval[0]), float(interval[1])
    d = (b - a) / R
    lo = a + index * d
    return (lo, lo + d)


def neighbors(index, R):
    """RankConnectivity with periodic wrap (multi_architectures.jl:90-137): (south, north)."""
    return ((index - 1) % R, (index + 1) % R)


class MultiArch:
    """MultiArch(child_architecture; ranks=(1, R, 1)) -- `rank`/`unique_id` are supplied by the launcher."""

    def __init__(self, child, ranks, rank, unique_id):
        if not isinstance(child, B200):
            raise TypeError("child architecture must be B200()")
        if ranks[0] != 1 or ranks[2] != 1:
            raise ValueError("only slab decomposition in y, ranks = (1, R, 1), is supported")
        self.child, self.ranks, self.local_rank = child, tuple(ranks), int(rank)
        self.R = int(ranks[1])
        self.local_index = (0, self.local_rank, 0)
        self.connectivity = neighbors(self.local_rank, self.R)
        if self.R > 1:
            check(lib.ob200_comm_init(self.R, self.local_rank, C.c_char_p(bytes(unique_id))))

    @staticmethod
    def new_unique_id():
        buf = C.create_string_buffer(128)
        check(lib.ob200_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device_index, ranks=None):
        """every rank of an initialised torch.distributed group calls this"""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        child = B200(device_index)
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            ident = torch.tensor(list(cls.new_unique_id()), dtype=torch.uint8)
        dev = torch.device("cuda", device_index) if dist.get_backend() == "nccl" else torch.device("cpu")
        ident = ident.to(dev)
        dist.broadcast(ident, 0)
        return cls(child, ranks or (1, world, 1), rank, bytes(ident.cpu().tolist()))

    def allreduce(self, values, op="sum"):
        arr = (C.c_double * len(values))(*values)
        check(lib.ob200_comm_allreduce(arr, len(values), 1 if op == "max" else 0))
        return list(arr)
