"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py
The slab-decomposed model (y split over the ranks, NCCL halo exchange + all-to-all FFT transposes) must
reproduce the single-process oracle on the same global initial condition to <= 1e-12 per step."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "clima-oceananigans.jl_b200")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import oracle as O
    import ocean_b200 as ob

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, R = dist.get_rank(), dist.get_world_size()
    arch = ob.MultiArch.from_torch_distributed(local_rank)
    N = (64, 16 * R if 16 * R >= 32 else 32, 32)
    L = (1.0, 2.0, 1.5)
    topo = ("Periodic",) * 3

    # ---- halo exchange: every rank fills its slab with its rank number + position code ------------
    gb = ob.RectilinearGrid(arch, np.float64, size=N, extent=L, topology=topo)
    f = ob.CenterField(gb)
    nl = gb.N
    jj = (np.arange(nl[1]) + rank * nl[1]).reshape(1, -1, 1)
    code = np.arange(nl[0]).reshape(-1, 1, 1) + 1000.0 * jj + 1e6 * np.arange(nl[2]).reshape(1, 1, -1)
    f.set(code + np.zeros(nl))
    ob.fill_halo_regions(f)
    p = f.parent()
    H = 3
    jg = (np.arange(-H, nl[1] + H) + rank * nl[1]) % N[1]
    ig = np.arange(-H, nl[0] + H) % N[0]
    kg = np.arange(-H, nl[2] + H) % N[2]
    want = ig.reshape(-1, 1, 1) + 1000.0 * jg.reshape(1, -1, 1) + 1e6 * kg.reshape(1, 1, -1)
    assert np.array_equal(p, want), "halo exchange (incl. edges and corners) differs from the periodic extension"

    # ---- full model vs the single-process oracle ---------------------------------------------------
    go = O.RectilinearGrid(np.float64, size=N, extent=L, topology=topo)
    mo = O.NonhydrostaticModel(go, advection=O.WENO5(), tracers=("b",), buoyancy=O.BuoyancyTracer(),
                               timestepper="RungeKutta3")
    mb = ob.NonhydrostaticModel(gb, advection=ob.WENO5(), tracers=("b",), buoyancy=ob.BuoyancyTracer(),
                                timestepper="RungeKutta3")
    rng = np.random.default_rng(5)
    vals = {}
    for n in "uvw":
        a = rng.uniform(-1, 1, N)
        vals[n] = a - a.mean()
    vals["b"] = 0.5 * go.nodes(("c", "c", "c"))[2] + 0.1 * rng.uniform(-1, 1, N)
    mo.set(**vals)
    sl = mb.grid.local_slice()
    ob.set_model(mb, **{n: v[sl] for n, v in vals.items()})
    worst = 0.0
    for step in range(3):
        mo.time_step(2e-3)
        ob.time_step(mb, 2e-3)
        for n in mo.names:
            a, b = mb.fields[n].interior(), mo.fields[n].interior[sl]
            e = float(np.max(np.abs(a - b)) / np.max(np.abs(mo.fields[n].interior)))
            worst = max(worst, e)
            assert e < 1e-12, (rank, step, n, e)
    d = mb.diagnostics()
    mx = arch.allreduce([d["max_abs_div"]], "max")[0]
    ke = arch.allreduce([d["kinetic_energy"]], "sum")[0]
    assert mx < 1e-10 and abs(ke - mo.kinetic_energy()) <= 1e-11 * mo.kinetic_energy()
    # ---- BASELINE config 3 physics on the slab decomposition: Bounded, vertically stretched z (Fourier-tridiagonal solve with the
    # y transposes around the Thomas sweep), WENO5(grid), ScalarDiffusivity, FPlane, Flux / Gradient BCs -------------------------
    zf = -1.0 + np.linspace(0.0, 1.0, N[2] + 1) ** 1.3
    kw3 = dict(size=N, x=(0, 1), y=(0, 2), z=zf, topology=("Periodic", "Periodic", "Bounded"))
    go3 = O.RectilinearGrid(np.float64, **kw3)
    gb3 = ob.RectilinearGrid(arch, np.float64, **kw3)
    mk_bcs = lambda M: {"u": {"top": M.BoundaryCondition("Flux", -1e-3)},
                        "b": {"top": M.BoundaryCondition("Flux", 1e-4), "bottom": M.BoundaryCondition("Gradient", 1e-2)}}
    mo3 = O.NonhydrostaticModel(go3, advection=O.WENO5(grid=go3), tracers=("b",), buoyancy=O.Buoyancy(O.BuoyancyTracer(), None),
                                closure=O.ScalarDiffusivity("ThreeDimensional", ν=1e-4, κ=2e-4), coriolis=O.FPlane(1e-2),
                                timestepper="RungeKutta3", boundary_conditions=mk_bcs(O))
    mb3 = ob.NonhydrostaticModel(gb3, advection=ob.WENO5(grid=gb3), tracers=("b",), buoyancy=ob.Buoyancy(ob.BuoyancyTracer(), None),
                                 closure=ob.ScalarDiffusivity("ThreeDimensional", ν=1e-4, κ=2e-4), coriolis=ob.FPlane(1e-2),
                                 timestepper="RungeKutta3", boundary_conditions=mk_bcs(ob))
    vals3 = {}
    for n in mo3.names:
        a = rng.uniform(-1, 1, mo3.fields[n].size())
        vals3[n] = a - a.mean() if n in "uvw" else 0.5 * go3.nodes(("c", "c", "c"))[2] + 0.1 * a
    mo3.set(**vals3)
    ob.set_model(mb3, **{n: v[sl] for n, v in vals3.items()})
    worst3 = 0.0
    for step in range(3):
        mo3.time_step(2e-3)
        ob.time_step(mb3, 2e-3)
        for n in mo3.names:
            a, b = mb3.fields[n].interior(), mo3.fields[n].interior[sl]
            e = float(np.max(np.abs(a - b)) / np.max(np.abs(mo3.fields[n].interior)))
            worst3 = max(worst3, e)
            assert e < 1e-12, ("C3", rank, step, n, e)
    t = torch.tensor([worst, worst3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"DIST_OK ranks={R} global={N} worst_rel_err={t[0].item():.3e} c3_bounded_stretched_worst_rel_err={t[1].item():.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
