#!/usr/bin/env python
"""bench.py -- grid-point updates/sec of one RK3 step of NonhydrostaticModel (256^3 triply periodic,
WENO5 + buoyancy tracer + FFT pressure solve, Float64) on B200, plus HBM roofline fraction of the
dominant kernel and a CPU baseline.  Contract: see the task statement / DESIGN.md section 6.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 256] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU).  One "step" = one full RK3 time step (3 stages,
3 pressure solves) of every cell of the workload.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "clima-oceananigans.jl_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

# mean DRAM bytes per tendency launch at 256^3 from the committed ncu capture (profiles/r1_summary.md)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 8.93e8

METRIC = "grid-point updates/sec (RK3 step, 256^3 WENO5+FFT)"
UNIT = "grid-point updates/s"


def synthetic_state(N, seed=2):
    """uniform(-1,1) velocities with the mean removed, b = N^2 z + noise."""
    import numpy as np
    rng = np.random.default_rng(seed)
    vals = {}
    for n in "uvw":
        a = rng.uniform(-1, 1, (N, N, N))
        vals[n] = a - a.mean()
    z = (np.arange(N) + 0.5) / N
    vals["b"] = 1e-5 * z.reshape(1, 1, N) + 1e-3 * rng.uniform(-1, 1, (N, N, N))
    return vals


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons every 20 ms.  It is started BEFORE the warm-up (the nvidia-smi
    process needs ~0.1 s to deliver its first line); mark_begin()/mark_end() bracket the timed region and the
    summary reports the samples that fall inside it (and, beside them, all samples taken under load)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append((time.perf_counter(), parts))
                if self.stop_flag:
                    break
            self.proc.terminate()
        except Exception:
            pass

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summ(samples):
            sm = sorted(int(s[0]) for _, s in samples if s[0].isdigit())
            reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for _, s in samples)]
            return (sm[len(sm) // 2] if sm else None), reasons
        inside = [x for x in self.samples if self.t0 is not None and self.t1 is not None and self.t0 <= x[0] <= self.t1 + 0.02]
        load = [x for x in self.samples if self.t0 is None or x[0] >= self.t0 - 1.0]
        med_in, reasons_in = summ(inside) if inside else (None, [])
        med_load, reasons_load = summ(load if load else self.samples)
        return {"sm_mhz": med_in if med_in is not None else med_load, "sm_max_mhz": int(self.samples[0][1][1]),
                "reasons": sorted(set(reasons_in) | set(reasons_load)), "samples": len(inside),
                "samples_under_load": len(load), "sm_mhz_under_load": med_load}


def cpu_reference_run(steps, warmup, sample_n=128):
    """The reference's CPU path is pure Julia and cannot run here (no Julia toolchain).
    What is timed is the oracle port -- the compiled OpenMP twin oracle/oracle_cpu.c, same arithmetic and
    same work per point as the reference's CPU kernels (both faces, both WENO sides per cell), on all host
    cores -- on a bounded sample of the same workload: RK3 steps of the triply periodic WENO5 + b + FFT
    model at sample_n^3 (the per-point cost does not depend on N)."""
    from oracle import cpu_twin
    N = sample_n
    vals = synthetic_state(N)
    dt = 0.1 / N
    # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)
    cores = max(cpu_twin.max_threads(), len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    args = ((N, N, N), (1.0, 1.0, 1.0), vals["u"], vals["v"], vals["w"], vals["b"])
    # the set-up (halo allocation, copies) is inside the call; time two run lengths and difference them
    t0 = time.perf_counter()
    cpu_twin.rk3_run(*args, 0, dt, project=False, nthreads=cores)
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    cpu_twin.rk3_run(*args, steps, dt, project=False, nthreads=cores)
    el = max(time.perf_counter() - t0 - t_setup, 1e-9)
    return (N ** 3 * steps / el, el / steps, cores,
            f"{steps} RK3 step(s) of the same model at {N}^3 (compiled OpenMP twin of the oracle, {cores} threads)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ftype", default="f64", choices=["f64", "f32"],
                    help="arithmetic type of the run (the headline configuration is Float64; f32 is reported for reference)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = a.size
    workload = f"C2: {N}^3 triply-periodic NonhydrostaticModel, WENO5 + tracer b + BuoyancyTracer, FFT pressure solve, RK3, Float64"

    if a.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(a.steps, 3))
        v, spstep, cores, sample = cpu_reference_run(steps, min(a.warmup, 1))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": min(a.warmup, 1), "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": workload},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import numpy as np
    import torch
    import ocean_b200 as ob
    from ocean_b200._lib import lib
    import ctypes as C

    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()
    lib.ob200_set_stream(C.c_void_p(stream.cuda_stream))
    # weak scaling: N^3 cells per GPU.  With several GPUs the GLOBAL domain N x (N*world) x N is slab-decomposed
    # in y (ranks = (1, world, 1) as in the reference's distributed benchmarks): NCCL halo exchange + all-to-all
    # transposes inside the FFT pressure solve
    if world > 1:
        arch = ob.MultiArch.from_torch_distributed(local_rank)
    else:
        arch = ob.B200(local_rank)
    FTYPE = np.float64 if a.ftype == "f64" else np.float32
    grid = ob.RectilinearGrid(arch, FTYPE, size=(N, N * world, N), extent=(1, world, 1),
                              topology=("Periodic",) * 3)
    model = ob.NonhydrostaticModel(grid, advection=ob.WENO5(FTYPE), tracers=("b",), buoyancy=ob.BuoyancyTracer(),
                                   timestepper="RungeKutta3")
    vals = synthetic_state(N, seed=2 + rank)
    ob.set_model(model, **vals)
    dt = 0.1 / N          # CFL ~ 0.1-0.3 for |u| <~ 1..3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(a.warmup):
        ob.time_step(model, dt)
    barrier()
    lib.ob200_profile_reset()
    lib.ob200_profile_enable(1)
    n0 = ob.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    ev0.record(stream)
    for _ in range(a.steps):
        ob.time_step(model, dt)
    ev1.record(stream)
    barrier()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    launches = ob.launch_count() - n0
    lib.ob200_profile_enable(0)
    phases = {}
    fft_names = ("fft_x_fwd", "fft_y", "fft_z", "fft_z_fwd", "fft_z_inv", "fft_sync", "fft_x_inv")
    for ph in ("tendency", "poisson", "halo", "pressure_correct", "hydrostatic") + fft_names:
        t, c = C.c_double(), C.c_int64()
        lib.ob200_profile_query(ph.encode(), C.byref(t), C.byref(c))
        phases[ph] = {"ms_total": t.value, "count": c.value}
    fft_phases = {k: phases.pop(k) for k in fft_names}
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = world * N ** 3 * a.steps / (ms * 1e-3)
    d = model.diagnostics()
    if world > 1:
        d["max_abs_div"] = arch.allreduce([d["max_abs_div"]], "max")[0]
        d["kinetic_energy"] = arch.allreduce([d["kinetic_energy"]], "sum")[0]
    assert np.isfinite(d["kinetic_energy"]) and d["max_abs_div"] < (1e-8 if a.ftype == "f64" else 1e-1), d

    # ---- roofline of the dominant kernel: the fused tendency+substep kernel (one launch per field) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    F = 4
    tend = phases["tendency"]
    # algorithmic words per point per launch: (4F+1)/F for stages 2,3 and (3F+1)/F for stage 1
    words = ((3 * F + 1) + 2 * (4 * F + 1)) / 3.0 / F
    W = 8 if a.ftype == "f64" else 4
    alg_bytes = words * W * N ** 3
    # the "tendency" phase brackets the F launches of a stage (they run on forked streams so that their tails overlap)
    tend_launches = a.steps * 3 * F
    avg_ms = tend["ms_total"] / max(1, tend_launches)
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "tendency+substep (per prognostic field)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH if N == 256 else None,
                "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of the 4 tendency "
                                  "launches of a stage (profiles/r1_summary.md)",
                "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "avg_launch_ms": avg_ms, "launches": tend_launches,
                "share_of_step": tend["ms_total"] / ms if ms > 0 else None,
                "whole_step": {"algorithmic_GB_per_step": 110.0 * W * N ** 3 / 1e9,
                               "achieved_GBps": 110.0 * W * N ** 3 * a.steps / (ms * 1e-3) / 1e9,
                               "frac": 110.0 * W * N ** 3 * a.steps / (ms * 1e-3) / 1e9 / peak},
                "phases_ms_per_step": {k: v["ms_total"] / a.steps for k, v in phases.items()},
                "poisson_ms_per_step": {k: v["ms_total"] / a.steps for k, v in fft_phases.items() if v["count"]}}

    # ---- e2e: host buffers in, host buffers out, every step (pinned; copies inside the timed region) ----
    e2e = None
    if not a.no_e2e:
        names = list(model.names)
        hin = {n: torch.from_numpy(np.asfortranarray(model.fields[n].parent()).ravel(order="K").copy()).pin_memory()
               for n in names}
        # two sets of pinned output buffers: step n writes set n % 2 and the host "consumes" (checksums one value of)
        # set (n - 1) % 2 while the GPU works, as a streaming caller would
        hout = [{n: torch.empty_like(hin[n]).pin_memory() for n in names} for _ in range(2)]
        nbytes = sum(t.numel() * t.element_size() for t in hin.values())
        ksteps = max(4, min(a.steps, 10))
        done_events = [torch.cuda.Event() for _ in range(2)]

        def enqueue(step):
            """one step through the C ABI with HOST buffers: H2D of the four parent arrays, time_step!, D2H of the
            four parent arrays.  Everything is asynchronous (pinned memory; uploads and downloads run on the library's
            two copy streams), so the upload of the next step and the download of the previous one overlap the kernels."""
            for n in names:
                lib.ob200_field_set_parent_async(model.fields[n].handle, C.c_void_p(hin[n].data_ptr()))
            ob.time_step(model, dt)
            for n in names:
                lib.ob200_field_get_parent_async(model.fields[n].handle, C.c_void_p(hout[step % 2][n].data_ptr()))
            lib.ob200_mark_download_batch()

        def run_pipelined(k):
            acc = 0.0
            for step in range(k):
                if step >= 2:
                    lib.ob200_sync_downloads(1)                 # the set about to be overwritten (step - 2) has been delivered
                    acc += float(hout[step % 2][names[0]][12345])
                enqueue(step)
            lib.ob200_sync()
            return acc
        run_pipelined(2)
        barrier()
        t0 = time.perf_counter()
        run_pipelined(ksteps)
        barrier()
        el = time.perf_counter() - t0
        # the unpipelined variant (one step at a time, synchronised after every step) for comparison
        t0 = time.perf_counter()
        for step in range(2):
            enqueue(step)
            lib.ob200_sync()
        el_serial = (time.perf_counter() - t0) / 2
        if world > 1:
            tt = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            el = float(tt.item())
        # resident-state variant: the call a user of the B200() architecture makes inside run!:
        # time_step!(model, dt) + the per-step scalar diagnostics read back (NaN check / CFL)
        t0 = time.perf_counter()
        for _ in range(ksteps):
            ob.time_step(model, dt)
            model.velocities["u"].reduce()
        barrier()
        el_res = time.perf_counter() - t0
        e2e = {"value": world * N ** 3 * ksteps / el, "unit": UNIT, "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "steps": ksteps,
               "note": "every step: H2D of u,v,w,b parent arrays from pinned host memory, time_step!, D2H of the same; "
                       "copies on two copy streams overlap the kernels of neighbouring steps (PCIe-bound)",
               "one_step_at_a_time": {"value": world * N ** 3 / el_serial,
                                      "note": "same, host synchronises after every step (no overlap)"},
               "resident_state": {"value": world * N ** 3 * ksteps / el_res, "d2h_bytes_per_step": 32,
                                  "note": "state stays on the device (how run! uses the architecture); per step a scalar reduction is read back"}}

    sampler.stop_flag = True
    sampler.join(timeout=2)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, spstep, cores, sample = cpu_reference_run(2, 0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.ftype, "data": "synthetic",
            "config": {"workload": workload if a.ftype == "f64" else workload.replace("Float64", "Float32"), "grid": [N, N, N], "timestepper": "RungeKutta3", "advection": "WENO5 (Z)",
                       "fields": F, "dt": dt, "l2": "inputs larger than L2 (14 fields x 144 MB)",
                       "parallelism": "single GPU" if world == 1 else
                       f"slab decomposition in y, ranks=(1,{world},1): global grid {N}x{N * world}x{N}, {N}^3 per GPU; "
                       "NCCL halo exchange + 2 all-to-all transposes per pressure solve"},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
