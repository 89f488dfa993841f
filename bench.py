#!/usr/bin/env python
"""bench.py — FP64 RHS DOF/s of the semi-discrete residual (3-D Euler, p=4 curved tets, flux
differencing) on N B200s, strong scaling over a fixed periodic mesh, plus roofline, CPU baseline and
end-to-end (host-buffer) numbers.  One JSON line on rank 0.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 ... bench.py --gpus 8 ...
    python bench.py --impl reference        # the reference algorithm on the host cores (oracle port)

A "step" is one evaluation of semi_discrete_residual! over the whole mesh (Solvers.jl:474-514).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200")]

METRIC = "FP64 RHS DOF/s (3D Euler p=4 tets, flux-diff)"
# BASELINE config 4 as a second workload (--workload advection_3d): HBM-bound, algorithmic 14 560 B per element;
# the fused pass-B kernel streams everything but the u read / u_f write of pass A
METRIC_ADV = "FP64 RHS DOF/s (3D advection p=4 tets, StandardForm)"
ALG_BYTES_ADV = 14560.0
ALG_BYTES_ADV_FUSED = 14560.0 - 280.0 - 800.0
NCU_DRAM_BYTES_ADV = {"k_adv_facets_ct": 1078.0, "k_adv_fused_ct": 7735.0}    # profiles/r2_ncu_config4.csv (196 608 elements)
# algorithmic figures per element, p=4 tet Euler
ALG_BYTES_RHS = 23200.0          # compulsory HBM traffic of one RHS
ALG_BYTES_PASS_B = 26800.0       # time_derivative kernel alone: u_q 5000 + own/nbr u_f 8000 + Λ 9000 + J_q 1000 + nJf 2400 + dudt 1400
ALG_FLOPS_RHS = 476000.0         # FMA = 2
ALG_FLOPS_PAIR = 359000.0        # the dominant kernel (k_fluxdiff_ct): volume + facet correction + interface flux + lift
# measured DRAM bytes / element (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum at 24 576 elements,
# profiles/r2_ncu_kernels.csv): pass A 9.05 kB, pair kernel 25.67 kB, projection 6.75 kB
NCU_DRAM_BYTES = {"k_nodal_ct": 9051.0, "k_fluxdiff_ct": 25668.0, "k_project_ct": 6747.0}
NCU_DRAM_BYTES_PASS_B = NCU_DRAM_BYTES["k_fluxdiff_ct"] + NCU_DRAM_BYTES["k_project_ct"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="euler_tgv_3d", choices=["euler_tgv_3d", "advection_3d"],
                    help="euler_tgv_3d: BASELINE config 5 (the headline metric); advection_3d: config 4")
    ap.add_argument("--cells", type=int, default=int(os.environ.get("SSE_BENCH_CELLS", "0")),
                    help="cubes per direction (6 tets each); default 56 -> 1 053 696 elements (config 5), 32 -> 196 608 (config 4)")
    ap.add_argument("--flux", default="lf", choices=["lf", "ec"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-cells", type=int, default=16, help="cubes per direction of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=1)
    a = ap.parse_args()
    if a.cells <= 0:
        a.cells = 56 if a.workload == "euler_tgv_3d" else 32
    return a


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.01)
        except Exception as e:          # never let monitoring kill the benchmark
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def workload_name(cells, n_e, dof, flux, workload="euler_tgv_3d"):
    if workload == "advection_3d":
        return (f"advection_3d: 3-D linear advection a = (1,1,1), ModalTensor p=4 curved tets, {cells}^3 cubes x 6 = {n_e} elements "
                f"({dof} DOF), StandardForm(skew-symmetric mapping, {'Lax-Friedrichs' if flux == 'lf' else 'central'} interface flux), "
                "DelReyWarping 0.1, conservative-curl metrics")
    return (f"euler_tgv_3d: 3-D Euler Taylor-Green vortex, ModalTensor p=4 curved tets, {cells}^3 cubes x 6 = "
            f"{n_e} elements ({dof} DOF), FluxDifferencingForm(EC two-point, "
            f"{'Lax-Friedrichs' if flux == 'lf' else 'EC'} interface), ChanWarping 1/16, conservative-curl metrics")


def build_case(cells, flux, part, device=None, workload="euler_tgv_3d"):
    from sse_b200 import cases
    if workload == "advection_3d":
        from sse_b200.assembly import REFERENCE_OPERATOR, SpatialDiscretization, StandardForm
        from sse_b200.laws import LinearAdvectionEquation, initial_data_cosine, project_function_reference
        from sse_b200.mesh import DelReyWarping, uniform_periodic_mesh
        from sse_b200.reference import ModalTensor, reference_approximation
        ra = reference_approximation(ModalTensor(4), "Tet", mapping_degree=4)
        mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 3, (cells,) * 3, DelReyWarping(0.1, (1.0,) * 3), part)
        sd = SpatialDiscretization.build(mesh, ra, "curl", need_nJq=False, device=device)
        ic = initial_data_cosine(1.0, (2 * np.pi,) * 3)
        c = cases.Case("advection_3d", LinearAdvectionEquation((1.0, 1.0, 1.0)), sd,
                       StandardForm(inviscid_numerical_flux=cases._flux(flux)), REFERENCE_OPERATOR, ic)
        u0 = project_function_reference(ic, ra, mesh.xyzq)
        mesh.xyzq = mesh.xyzf = None
        return c, u0
    from sse_b200.assembly import FluxDifferencingForm, REFERENCE_OPERATOR, SpatialDiscretization
    from sse_b200.laws import EulerEquations, project_function_reference, taylor_green_vortex
    from sse_b200.mesh import ChanWarping, uniform_periodic_mesh
    from sse_b200.reference import ModalTensor, reference_approximation
    L = 2 * np.pi
    ra = reference_approximation(ModalTensor(4), "Tet", mapping_degree=4)
    mesh = uniform_periodic_mesh(ra, ((0.0, L),) * 3, (cells,) * 3, ChanWarping(1.0 / 16.0, (L,) * 3), part)
    sd = SpatialDiscretization.build(mesh, ra, "curl", need_nJq=False, device=device)   # metrics on the GPU when given
    ic = taylor_green_vortex(1.4, 0.1)
    c = cases.Case("euler_tgv_3d", EulerEquations(3, 1.4), sd,
                   FluxDifferencingForm(inviscid_numerical_flux=cases._flux(flux)), REFERENCE_OPERATOR, ic)
    u0 = project_function_reference(ic, ra, mesh.xyzq)
    mesh.xyzq = mesh.xyzf = None          # free host memory that is no longer needed
    return c, u0


def _cpu_sample_note(cells, case, reps, t, used):
    return (f"same workload on a {cells}^3-cube sample ({case.sd.N_e} elements, {case.dof} DOF; the full 56^3 mesh needs ~15 s per "
            f"RHS on the host), {reps} RHS, {t:.3f} s each, {used} OpenMP threads = every host thread of this process's affinity "
            "mask; C/OpenMP restatement of the reference algorithm as written (Threads.@threads over elements, "
            "Solvers.jl:505-511), built -O3 -march=native on this box (the Julia reference cannot run in this image)")


def cpu_baseline(cells, flux, budget_s=12.0, workload="euler_tgv_3d"):
    """The reference algorithm (oracle port, OpenMP over elements like Threads.@threads) on a bounded
    sample of the same workload, on this box's host cores: about `budget_s` seconds of CPU work."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    c, u0 = build_case(cells, flux, None, workload=workload)
    img = c.image()
    nt = oracle.host_threads()
    t1, used, _ = oracle.time_rhs(img, u0, nt, 1, perf=True)                 # warm-up and cost estimate
    reps = int(min(max(budget_s / max(t1, 1e-6), 3), 60))
    t, used, _ = oracle.time_rhs(img, u0, nt, reps, perf=True)
    return {"value": c.dof / t, "unit": "DOF/s", "cores": used, "kind": "port", "same_config": False,
            "sample": _cpu_sample_note(cells, c, reps, t, used) + f" (best of {reps})"}


def run_reference(a):
    """--impl reference: the reference's own CPU algorithm for the path, all host threads, on a bounded sample of the workload.
    Under torchrun only rank 0 works (and it ignores OMP_NUM_THREADS=1, which torchrun exports: the thread count is explicit)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    cells = a.cpu_cells
    c, u0 = build_case(cells, a.flux, None, workload=a.workload)
    img = c.image()
    nt = oracle.host_threads()
    for _ in range(max(a.warmup, 1)):
        t, used, _ = oracle.time_rhs(img, u0, nt, 1, perf=True)
    ts = []
    for _ in range(a.steps):
        t, used, _ = oracle.time_rhs(img, u0, nt, 1, perf=True)
        ts.append(t)
    t = float(np.mean(ts))
    val = c.dof / t
    n_e_full = 6 * a.cells ** 3
    dpe = 175 if a.workload == "euler_tgv_3d" else 35
    sample = "each step = one RHS: " + _cpu_sample_note(cells, c, a.steps, t, used)
    print(json.dumps({
        "impl": "reference", "metric": METRIC if a.workload == "euler_tgv_3d" else METRIC_ADV, "value": val, "unit": "DOF/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.cells, n_e_full, n_e_full * dpe, a.flux, a.workload), "sample": sample, "same_config": False},
        "cpu_baseline": {"value": val, "unit": "DOF/s", "cores": used, "kind": "port", "same_config": False, "sample": sample},
        "e2e": {"value": val, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def bind_to_gpu_numa_node(index):
    """Pin this process to the host cores NVML reports as local to the GPU (what numactl does for a production run): the
    pinned staging buffers of the end-to-end path are then first-touched on the GPU's NUMA node.  Returns the previous mask."""
    try:
        old = os.sched_getaffinity(0)
    except AttributeError:
        return None
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 64
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= old
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    return old


def dist_parity(rank, world, local):
    """Element-partitioned residual on `world` GPUs against the single-domain oracle (rank 0 checks): config 5 and config 4
    at M = 4, through the same DistributedSolver.rhs the timed region used.  Returns the worst relative difference."""
    import torch
    import torch.distributed as dist
    from sse_b200 import cases
    from sse_b200.dist import DistributedSolver
    from sse_b200.solver import Solver
    worst = 0.0
    M = 4 if world <= 4 else 8
    for name, kw in (("euler_tgv_3d", dict(M=M, flux="lf")), ("advection_3d", dict(M=M, flux="lf"))):
        full = cases.BUILDERS[name](**kw)
        u_full = full.u0(seed=0)
        part = cases.BUILDERS[name](part=(rank, world), **kw)
        gid = part.sd.mesh.elem_gid
        s = Solver(part.image(), local)
        s.use_current_stream()
        ds = DistributedSolver(s, part.sd.mesh)
        u = torch.from_numpy(np.ascontiguousarray(u_full[gid])).cuda()
        du = s.new_state()
        for _ in range(2):
            ds.rhs(du, u)
        s.synchronize()
        got = [None] * world
        dist.all_gather_object(got, (gid, du.cpu().numpy()))
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle
            ref = oracle.rhs(full.image(), u_full, oracle.host_threads())
            out = np.empty_like(ref)
            for g, d in got:
                out[g] = d
            worst = max(worst, float(np.abs(out - ref).max() / np.abs(ref).max()))
        s.close()
    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.broadcast(t, 0)
    return float(t.item())


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    import torch
    import torch.distributed as dist
    from sse_b200.solver import Solver, fp64_peak, semi_discrete_residual
    from sse_b200.dist import DistributedSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsse_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    old_affinity = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    t_setup = time.time()
    case, u0 = build_case(a.cells, a.flux, (rank, world) if world > 1 else None, device=local, workload=a.workload)
    adv = a.workload == "advection_3d"
    img = case.image()
    solver = Solver(img, local)
    solver.set_kernel_variant(a.variant)
    solver.use_current_stream()
    mesh = case.sd.mesh
    n_e_local = case.sd.N_e
    n_e = 6 * a.cells ** 3
    dof = n_e * (35 if adv else 175)
    u = torch.from_numpy(u0).cuda()
    du = solver.new_state()
    ds = DistributedSolver(solver, mesh) if world > 1 else None
    t_setup = time.time() - t_setup

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        ds.rhs(du, u) if ds is not None else solver.rhs(du, u)      # one sse_rhs: the call a Julia residual makes

    for _ in range(max(a.warmup, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = solver.launches
    sync_all()
    e0.record()
    for i in range(a.steps):
        step()
    e1.record()
    sync_all()
    clocks = sampler.result()
    ms = e0.elapsed_time(e1)
    launches = solver.launches - launches0
    # per-kernel durations of the same residual (CUDA events between the kernels, on the stream they are launched on)
    kernel_ms = solver.profile_rhs(du, u, reps=min(max(a.steps, 3), 10)) if ds is None else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        ln = torch.tensor([launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(ln)
        launches = int(ln.item())
    ms_per_step = ms / a.steps
    value = dof / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers (pinned), H2D + D2H inside the timed region
    hu = torch.from_numpy(u0).pin_memory()
    hdu = torch.empty_like(hu).pin_memory()
    if ds is None:
        semi_discrete_residual(hdu, hu, solver)
    else:
        ds.rhs_host(hdu, hu)
    sync_all()
    # every end-to-end step is timed on its own (CUDA events around the synchronous host-buffer call); the reported time is
    # the median, so that one host-side hiccup (page faults of the pinned buffers, a noisy neighbour on the PCIe root) does not
    # decide the number; the spread is reported next to it
    e2e_times = []
    for _ in range(max(a.e2e_steps, 1)):
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        if ds is None:
            semi_discrete_residual(hdu, hu, solver)
        else:
            ds.rhs_host(hdu, hu)
        t1.record()
        sync_all()
        e2e_times.append(t0.elapsed_time(t1))
    e2e_ms = float(np.median(e2e_times))
    e2e_spread = [float(min(e2e_times)), float(max(e2e_times))]
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    state_bytes = int(u0.size) * 8

    # size-independent properties of the residual at the full size (outside every timed region): the discrete
    # conservation functionals 1' M dudt vanish for the periodic mesh, and the host-buffer call returns the device result
    checks = {}
    try:
        ds.rhs(du, u) if ds is not None else solver.rhs(du, u)
        f = torch.tensor(solver.functionals(u, du)[:int(solver.cfg.N_c)], dtype=torch.float64, device="cuda")
        sc = torch.tensor([float(du.abs().max())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(f)
            dist.all_reduce(sc, op=dist.ReduceOp.MAX)
        checks["conservation_residual_rel"] = float(f.abs().max() / (sc[0] * (1.0 if adv else (2 * np.pi) ** 3)))
        checks["host_buffer_result_equals_device"] = bool(torch.equal(hdu, du.cpu()))
        if world > 1:
            okt = torch.tensor([1.0 if checks["host_buffer_result_equals_device"] else 0.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            checks["host_buffer_result_equals_device"] = bool(okt.item() == 1.0)
    except Exception as e:
        checks["error"] = str(e)
    if world > 1:
        # parity of the partitioned path on this very communicator layout, against the single-domain oracle
        checks["dist_parity_rel"] = dist_parity(rank, world, local)
        checks["dist_parity_cases"] = "euler_tgv_3d and advection_3d (lf), M = %d, %d ranks, vs the oracle on the whole mesh" % (4 if world <= 4 else 8, world)
    if old_affinity:
        os.sched_setaffinity(0, old_affinity)            # the CPU baseline below uses every host thread again
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        out = {
            "metric": METRIC_ADV if adv else METRIC, "value": value, "unit": "DOF/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a.cells, n_e, dof, a.flux, a.workload),
                       "partition": (f"{world} slabs along z, facet-trace halos by ncclSend/ncclRecv inside sse_rhs "
                                     f"(NCCL {solver.comm_info()[2]})") if world > 1 else "single GPU",
                       "l2": "no flush: per-step inputs (metrics + state, ~24 kB/element) are far larger than the 126 MB L2",
                       "kernel_variant": solver.kernel_variant(), "setup_s": round(t_setup, 1)},
            "clocks": clocks,
            "e2e": {"value": dof / (e2e_ms * 1e-3), "unit": "DOF/s", "h2d_bytes_per_step": state_bytes * world,
                    "d2h_bytes_per_step": state_bytes * world, "ms_per_step": e2e_ms, "steps": len(e2e_times),
                    "ms_min_max": e2e_spread},
            "gpu_launches": launches,
        }
        try:
            fpeak = fp64_peak(local)
        except Exception as e:
            fpeak = None
            out["roofline_error"] = str(e)
        fp_src = "measured in this run: register-resident DFMA microbenchmark (sse_fp64_peak); MEASURED_PEAKS.json has no FP64 figure"
        if adv:
            gbs = ALG_BYTES_ADV * n_e / (ms_per_step * 1e-3) / 1e9
            out["roofline_rhs"] = {"bound": "hbm", "what": "whole residual (k_adv_facets_ct + k_adv_fused_ct)", "achieved": gbs,
                                   "peak": world * hbm_peak, "unit": "GB/s", "frac": gbs / (world * hbm_peak),
                                   "algorithmic_bytes_per_element": ALG_BYTES_ADV, "peak_source": hbm_src,
                                   "traffic": sum(NCU_DRAM_BYTES_ADV.values()) * n_e}
            if kernel_ms is not None:
                kms = {"k_adv_facets_ct": float(kernel_ms[0]), "k_adv_fused_ct": float(kernel_ms[2] + kernel_ms[3])}
                ach = ALG_BYTES_ADV_FUSED * n_e_local / (kms["k_adv_fused_ct"] * 1e-3) / 1e9
                out["roofline"] = {"bound": "hbm", "kernel": "k_adv_fused_ct (pass B in one kernel: V u, volume terms, interface flux, "
                                                             "lift, V', mass solve)", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": ach / hbm_peak, "traffic": NCU_DRAM_BYTES_ADV["k_adv_fused_ct"] * n_e_local,
                                   "kernel_ms": kms["k_adv_fused_ct"], "algorithmic_bytes_per_element": ALG_BYTES_ADV_FUSED,
                                   "peak_source": hbm_src, "kernels": kms,
                                   "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per element of this kernel (ncu "
                                                     "--set full, profiles/r2_ncu_config4.csv), scaled to this launch; it is below the "
                                                     "algorithmic figure because the kernel reads coefficient tables derived at "
                                                     "sse_create (3 instead of 9 metric terms per node) instead of the raw metrics"}
        elif fpeak:
            fl = ALG_FLOPS_RHS * n_e / (ms_per_step * 1e-3)
            # the binding bound of this path, for the whole residual (the figure BASELINE.json's >= 50 % target is about)
            out["roofline_rhs"] = {"bound": "fp64", "what": "whole residual (k_nodal_ct + k_fluxdiff_ct + k_project_ct" +
                                   (" + halo pack / unpack)" if world > 1 else ")"), "achieved": fl / 1e12,
                                   "peak": world * fpeak / 1e12, "unit": "TFLOP/s", "frac": fl / (world * fpeak),
                                   "algorithmic_flops_per_element": ALG_FLOPS_RHS, "peak_source": fp_src}
            out["roofline_hbm"] = {"bound": "hbm (not binding: 20 flop/B)", "achieved": ALG_BYTES_RHS * n_e / (ms_per_step * 1e-3) / 1e9,
                                   "peak": world * hbm_peak, "unit": "GB/s",
                                   "frac": ALG_BYTES_RHS * n_e / (ms_per_step * 1e-3) / 1e9 / (world * hbm_peak),
                                   "algorithmic_bytes_per_element": ALG_BYTES_RHS, "peak_source": hbm_src,
                                   "traffic": sum(NCU_DRAM_BYTES.values()) * n_e}
        if kernel_ms is not None and fpeak and not adv:
            flops = {"k_nodal_ct": ALG_FLOPS_RHS - ALG_FLOPS_PAIR - 36200.0, "k_fluxdiff_ct": ALG_FLOPS_PAIR, "k_project_ct": 36200.0}
            kms = {"k_nodal_ct": float(kernel_ms[0]), "k_fluxdiff_ct": float(kernel_ms[2]), "k_project_ct": float(kernel_ms[3])}
            ach = ALG_FLOPS_PAIR * n_e_local / (kms["k_fluxdiff_ct"] * 1e-3)
            out["roofline"] = {"bound": "fp64", "kernel": "k_fluxdiff_ct (dominant kernel: interface flux, volume flux differencing, "
                                                          "facet correction, lift)",
                               "achieved": ach / 1e12, "peak": fpeak / 1e12, "unit": "TFLOP/s", "frac": ach / fpeak,
                               "traffic": NCU_DRAM_BYTES["k_fluxdiff_ct"] * n_e_local, "kernel_ms": kms["k_fluxdiff_ct"],
                               "algorithmic_flops_per_element": ALG_FLOPS_PAIR, "peak_source": fp_src,
                               "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per element of this kernel from the "
                                                 "ncu --set full capture profiles/r2_ncu_kernels.csv (24 576 elements), scaled to "
                                                 "this launch",
                               "kernels": {k: {"ms": kms[k], "algorithmic_flops_per_element": flops[k],
                                               "frac_of_fp64_peak": flops[k] * n_e_local / (kms[k] * 1e-3) / fpeak,
                                               "share_of_step": kms[k] / sum(kms.values())} for k in kms}}
        out["checks"] = checks
        if not a.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_baseline(a.cpu_cells, a.flux, workload=a.workload)
            except Exception as e:
                out["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if checks.get("dist_parity_rel", 0.0) > 1e-12:
            raise SystemExit(f"partitioned residual differs from the oracle: {checks['dist_parity_rel']:.3e} > 1e-12")


if __name__ == "__main__":
    main()
