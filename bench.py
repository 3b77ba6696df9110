#!/usr/bin/env python
"""bench.py -- adjoint-RHS throughput (GDOF/s, fp64, lx=8) of the B200 path, one process per GPU.

A "step" is one fused evaluation of the hot path over one rank's brick of synthetic input:
source terms (RAMP + Brinkman + lube) -> mass matrix -> adjoint advection -> sensitivity, then
gs_op(f_i, GS_OP_ADD) on the three components incl. the NCCL shared-node exchange
(/root/reference/sources/adjoint/adjoint_pnpn.f90:669-682 + :755-757).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--ne 64] [--lx 8] [--impl reference]

Workload (BASELINE.json configs[3], the configuration the metric "GDOF/s at 1/2/4/8 B200" is quoted
on; it fits one GPU): 64^3 hexahedra per GPU, lx=8, global box (64 px) x (64 py) x (64 pz), fields keyed
on the global node lattice.  `--ne 32` gives configs[1].  All 21 fields (2.8 GB at 32^3, 22.5 GB at
64^3) are far larger than the 126 MB L2, so no L2 flush is needed between timed iterations.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the reference itself -- Fortran
on top of un-vendored Neko -- cannot be built here) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "adjoint-RHS GDOF/s (fp64, lx=8)"
UNIT = "GDOF/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ne", type=int, default=64, help="elements per direction per GPU")
    ap.add_argument("--lx", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dealias", action="store_true", help="skip the extra dealiased-operator leg (N = 1, lx = 8)")
    ap.add_argument("--dealias-ne", type=int, default=32)
    ap.add_argument("--cpu-sample-ne", type=int, default=16)
    return ap.parse_args()


def workload_name(ne, lx, n):
    tag = "configs[3]" if ne == 64 and lx == 8 else ("configs[1]" if ne == 32 and lx == 8 and n == 1 else "custom")
    return f"{tag}: synthetic box {ne}^3 hex elements per GPU, lx={lx}, random design field rho"


def config_dict(args):
    """`config` of the JSON line: the same dict, key for key, from both arms (b200 and --impl reference)."""
    from neko_top_b200 import workloads
    n = args.lx ** 3 * args.ne ** 3
    return {"workload": workload_name(args.ne, args.lx, args.gpus), "lx": args.lx,
            "elements_per_gpu": args.ne ** 3, "dof_per_gpu": n,
            "rank_grid": list(workloads.rank_grid(args.gpus)),
            "l2": "inputs (21 fields, %.1f GB per GPU) exceed the 126 MB L2; no flush needed" % (21 * n * 8 / 1e9)}


def host_cores():
    """Host cores this process may use (torch.distributed.run exports OMP_NUM_THREADS=1: ignored on purpose)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:      # pragma: no cover
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------
# clocks: sampled with NVML from a thread DURING the timed region
# ------------------------------------------------------------------------------------------------
_REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
            0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}


class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:     # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in _REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if os.environ.get("B200_BENCH_NO_CLOCKS"):
            self.ok = False
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()

    def stop(self):
        if self.ok:
            self._stop.set()
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU side: oracle on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_problem(ne_gpu, lx, sample_ne, px=1, py=1, pz=1):
    """sample_ne^3 elements cut from the corner of rank 0's brick of the same global mesh, same field
    generators (keyed on the global node lattice)."""
    import torch
    import neko_top_b200  # noqa: F401
    from neko_top_b200 import sem, workloads
    s = min(sample_ne, ne_gpu)
    brick = workloads.BoxBrick(lx=lx, ne=(s, s, s), ne_global=(ne_gpu * px, ne_gpu * py, ne_gpu * pz),
                               length=(float(px), float(py), float(pz)), name="cpu sample")
    sp = sem.Space(lx)
    x, y, z = workloads.coords(brick)
    keys = workloads.node_keys(brick)
    G, _, B = sem.geometric_factors(x, y, z, sp)
    fl = workloads.make_fields(brick, x, y, z, keys)
    f = lambda a: a.reshape(-1).numpy()
    return dict(brick=brick, sp=sp, G=[f(g) for g in G], B=f(B), v=[f(a) for a in fl.v], ub=[f(a) for a in fl.ub],
                rho=f(fl.rho), key=f(keys), desc=f"{s}^3-element corner brick of the same mesh and fields "
                                                f"({brick.n} DOF per step)")


def time_oracle(prob, steps, warmup, budget_s=None, keep=None):
    from oracle import pyoracle as orc
    orc.build()
    orc.set_num_threads(host_cores())
    b = prob["brick"]
    st = orc.RhsStep(prob["v"], prob["ub"], prob["rho"], b.lx, b.nelv, prob["sp"].dx, prob["sp"].wx, prob["G"],
                     prob["B"], prob["key"])
    for _ in range(warmup):
        st.step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        st.step()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    if keep is not None:
        keep["f"], keep["sens"] = st.f, st.sens
    return b.n * done / dt / 1e9, dt / done * 1e3, done, orc.num_threads()


def parity_vs_oracle(prob, f_dev, sens_dev, ne_gpu, oracle_out=None, tol=1e-12):
    """The GPU step's result on the corner sample against the oracle's on the same inputs.  The oracle sees the
    sample as a mesh of its own, so direct-stiffness sums are comparable on the nodes whose class lies inside
    the sample (not on its cut faces); the sensitivity is point-wise and compared everywhere."""
    import numpy as np
    import torch
    b = prob["brick"]
    s, lx, N = b.ne[0], b.lx, b.lx ** 3
    if oracle_out is None:
        oracle_out = {}
        time_oracle(prob, 1, 0, keep=oracle_out)
    e = torch.arange(b.nelv)
    ex, ey, ez = e % s, (e // s) % s, e // (s * s)
    idx = (ex + ne_gpu * (ey + ne_gpu * ez)).to(f_dev[0].device)
    p = torch.arange(lx)
    cut = []
    for d, (el, pt) in enumerate(((ex, p.view(1, 1, 1, lx)), (ey, p.view(1, 1, lx, 1)), (ez, p.view(1, lx, 1, 1)))):
        g = (el * (lx - 1)).view(-1, 1, 1, 1) + pt
        on_cut = (g == s * (lx - 1)) if s < b.ne_global[d] else torch.zeros_like(g, dtype=torch.bool)
        cut.append(on_cut.expand(b.nelv, lx, lx, lx))
    inside = ~(cut[0] | cut[1] | cut[2]).reshape(-1).numpy()
    rel = lambda a, r: float(np.linalg.norm(a - r) / max(np.linalg.norm(r), 1e-300))
    errs, worst_pt = [], (0.0, -1, -1)
    for c in range(3):
        got = f_dev[c].view(-1, N)[idx].reshape(-1).cpu().numpy()
        errs.append(rel(got[inside], oracle_out["f"][c][inside]))
        d = np.abs(got - oracle_out["f"][c]) * inside
        j = int(np.argmax(d))
        if d[j] > worst_pt[0]:
            worst_pt = (float(d[j]), c, j)
    es = rel(sens_dev.view(-1, N)[idx].reshape(-1).cpu().numpy(), oracle_out["sens"])
    worst = max(max(errs), es)
    return {"rel_l2_f": max(errs), "rel_l2_sens": es, "n_dof_checked": int(inside.sum()), "tol": tol,
            "max_abs_err_f": {"value": worst_pt[0], "component": worst_pt[1], "sample_element": worst_pt[2] // N,
                              "node": worst_pt[2] % N},
            "ok": bool(np.isfinite(worst) and worst <= tol),
            "against": "oracle/oracle.c on the " + prob["desc"] + "; f after gs on the nodes whose class lies inside "
                       "the sample, sens on all of it"}


def dealiased_leg(ne, lx, dev, steps=10, sample_nel=256, tol=1e-12):
    """Extra (N = 1): the same step with the DEALIASED adjoint operator (case.numerics.dealias = true, what the reference's
    shipped lx = 8 cases run) on an ne^3 box -- device-resident step time, and the fused right-hand side
    of the first `sample_nel` elements against oracle.c's dealiased operator."""
    import numpy as np
    import torch
    from neko_top_b200 import operators as ops, sem, workloads
    from oracle import pyoracle
    brick = workloads.config_box(ne, lx)
    sp = sem.Space(lx)
    x, y, z = workloads.coords(brick, dev)
    keys = workloads.node_keys(brick, dev)
    G, _, B = sem.geometric_factors(x, y, z, sp, chunk=8192)
    fl = workloads.make_fields(brick, x, y, z, keys)
    flat = lambda a: a.reshape(-1).contiguous()
    G, B = [flat(g) for g in G], flat(B)
    v, ub, rho = [flat(a) for a in fl.v], [flat(a) for a in fl.ub], flat(fl.rho)
    n = brick.n
    f = [torch.empty(n, device=dev, dtype=torch.float64) for _ in range(3)]
    sens = torch.empty(n, device=dev, dtype=torch.float64)
    op = ops.fused_adjoint_rhs_t(ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, G, B), device=dev.index)
    op.set_params()
    op.gs.init(flat(keys))
    lxd = 3 * lx // 2
    op.set_dealias(True)
    # parity of the element kernel (no summation) on a sample
    op.compute(v, ub, f, rho=rho, sens=sens)
    m = sample_nel * lx ** 3
    c = lambda t: t[:m].cpu().numpy()
    fo, so, _ = pyoracle.adjoint_rhs([c(a) for a in v], [c(a) for a in ub], lx, sample_nel, sp.dx,
                                     sp.wx, [c(g) for g in G], c(B), rho=c(rho), lxd=lxd)
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    err_f = max(rel(c(f[k]), fo[k]) for k in range(3))
    err_s = rel(c(sens), so)
    for _ in range(3):
        op.step(v, ub, f, rho=rho, sens=sens)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        op.step(v, ub, f, rho=rho, sens=sens)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    op.free()
    return {"value": n / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "workload": f"{ne}^3-element box, lx={lx}, lxd={lxd} (dealiased adjoint operator + Brinkman + sensitivity + "
                        "direct-stiffness summation)",
            "kernel": "advop_mma_kernel (FP64 tensor cores, mma.sync.m8n8k4.f64)" if lx == 8 else "advop_kernel",
            "parity": {"rel_l2_f": err_f, "rel_l2_sens": err_s, "tol": tol, "ok": bool(err_f <= tol and err_s <= tol),
                       "sample": f"first {sample_nel} elements, fused right-hand side vs oracle/oracle.c (lxd={lxd})"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from neko_top_b200 import workloads
    px, py, pz = workloads.rank_grid(args.gpus)
    prob = cpu_sample_problem(args.ne, args.lx, args.cpu_sample_ne, px, py, pz)
    val, ms, done, threads = time_oracle(prob, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": prob["desc"] + "; oracle/oracle.c (OpenMP over elements); the Fortran/Neko "
                                                  "reference cannot be built in this image"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# B200 side
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and therefore its pinned staging buffers, first-touch) to the CPUs NVML
    reports as local to the GPU -- what `numactl` / the MPI launcher does for a Neko rank.  Matters only for
    the host-buffer (`e2e`) path; returns the number of CPUs bound to (0: left alone)."""
    if os.environ.get("B200_BENCH_NO_AFFINITY"):
        return 0
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(local).uuid))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist
    import neko_top_b200  # noqa: F401
    from neko_top_b200 import operators as ops, sem, workloads

    N = args.gpus
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != N:
        raise SystemExit(f"--gpus {N} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {N}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: neko_top_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    if N > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lx, ne = args.lx, args.ne
    brick = workloads.config_weak(rank, N, ne, lx)
    sp = sem.Space(lx)
    x, y, z = workloads.coords(brick, dev)
    keys = workloads.node_keys(brick, dev).reshape(-1)
    G, jac, B = sem.geometric_factors(x, y, z, sp, chunk=8192)
    fl = workloads.make_fields(brick, x, y, z, keys.view(brick.nelv, lx, lx, lx))
    del x, y, z, jac
    flat = lambda a: a.reshape(-1).contiguous()
    G, B = [flat(g) for g in G], flat(B)
    v, ub, rho = [flat(a) for a in fl.v], [flat(a) for a in fl.ub], flat(fl.rho)
    del fl
    n = brick.n
    f = [torch.empty(n, device=dev, dtype=torch.float64) for _ in range(3)]
    sens = torch.empty(n, device=dev, dtype=torch.float64)

    coef = ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, G, B)
    op = ops.fused_adjoint_rhs_t(coef, device=local)
    op.set_params()                         # RAMP 0/1000/1 convex-up, K = 1, obj_scale = 1
    op.gs.init(keys)
    if N > 1:
        idb = [ops.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(idb, src=0)
        op.comm_init(idb[0], rank, N)
        # shared-node discovery behind the C ABI (device sort + ncclAllGather); the candidate mask plays the role
        # of Neko's dm_Xh%shared_dof
        cand = workloads.interface_candidates(brick, dev)
        op.gs.init_shared_from_keys(keys, cand)
        del cand
    del keys
    torch.cuda.empty_cache()

    def barrier():
        torch.cuda.synchronize()
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(val):
        if N == 1:
            return val
        t = torch.tensor([val], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(val):
        if N == 1:
            return val
        t = torch.tensor([val], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident timing: `value` ------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        op.step(v, ub, f, rho=rho, sens=sens)
    barrier()
    clk = ClockSampler(local)
    op.enable_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ops.launch_count()
    clk.start()
    barrier()
    e0.record()
    for _ in range(args.steps):
        op.step(v, ub, f, rho=rho, sens=sens)
    e1.record()
    barrier()
    clk.stop()
    launches = ops.launch_count() - l0
    ms_total = maxr(e0.elapsed_time(e1))
    elem_ms, gs_ms, nmarks = op.get_timing()
    phases = op.get_phase_timing() if os.environ.get("B200_PHASE_TIMING") else None
    op.enable_timing(False)
    ms_step = ms_total / args.steps
    n_global = sumr(float(n))
    value = n_global / (ms_step * 1e-3) / 1e9
    launches_all = int(sumr(float(launches)))
    checksum = float(sum(float(t.double().abs().sum().item()) for t in f))

    # ---- e2e: host buffers through the C ABI ----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        hv = [t.cpu().pin_memory() for t in v]
        hub = [t.cpu().pin_memory() for t in ub]
        hrho = rho.cpu().pin_memory()
        hf = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
        hs = torch.empty(n, dtype=torch.float64).pin_memory()
        op.step_host(hv, hub, hrho, hf, hs)           # warm-up (allocates the staging buffers)
        op.step_host(hv, hub, hrho, hf, hs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            op.step_host(hv, hub, hrho, hf, hs)
        torch.cuda.synchronize()
        dt = maxr(time.perf_counter() - t0)
        barrier()
        # the host-buffer path sums every class from its member list; the device path's staged sums associate 4- and
        # 8-member classes pairwise: equal up to rounding
        e2e_diff = max(float((hf[c] - f[c].cpu()).abs().max() / f[c].abs().max().cpu()) for c in range(3))
        e2e_ok = e2e_diff <= 4e-15 and bool(torch.equal(hs, sens.cpu()))
        e2e = {"value": n_global / (dt / args.e2e_steps) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(7 * n * 8), "d2h_bytes_per_step": int(4 * n * 8),
               "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
               "matches_device_path": bool(e2e_ok), "max_rel_diff_vs_device_path": e2e_diff,
               "note": "b200_adjrhs_step_host: 7 input fields H2D from pinned memory, f(3)+sens D2H, "
                       "geometry resident (set once like coef_t); bytes are per rank"}
        # what bounds e2e: the host link of this rank while ALL ranks copy at the same time (one VM, shared PCIe
        # switches / host memory): 1 GiB pinned H2D and D2H, concurrently in both directions like the pipelined step
        nb = min(n, 1 << 27)
        dbuf = torch.empty(nb, device=dev, dtype=torch.float64)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        link = 0.0
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_in):
                dbuf.copy_(hv[0][:nb], non_blocking=True)
            with torch.cuda.stream(s_out):
                hf[0][:nb].copy_(f[0][:nb], non_blocking=True)
            torch.cuda.synchronize()
            dt_link = maxr(time.perf_counter() - t0)
            link = max(link, nb * 8 / dt_link / 1e9)
        e2e["host_link"] = {"GBps_each_direction_per_rank": link, "ranks_concurrent": N,
                            "h2d_GBps_per_rank_in_step": 7 * n * 8 / (dt / args.e2e_steps) / 1e9,
                            "note": "best of 3: 1 GiB pinned H2D + 1 GiB D2H at once on EVERY rank (slowest rank counts): "
                                    "what the VM's host memory / PCIe switches give N concurrent ranks; the step moves "
                                    "7 fields in and 4 out per rank at the rate shown"}
        barrier()
        del hv, hub, hrho, hf, hs, dbuf

    # ---- roofline of the dominant kernel (rank 0) ------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bpd = sem.algorithmic_bytes_per_dof(lx, with_gs=False)
    achieved = n * bpd / (elem_ms * 1e-3) / 1e9 if elem_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(f"ne{ne}_lx{lx}", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    xs_active, xs_linked, xs_left, xs_total = op.xstage_info()
    if lx == 8:
        kname = ("adjrhs_v3_kernel<3,4,2,14,168,XS> (fused element kernel, DMMA contractions, x-face node pairs "
                 "summed in the kernel; 168 B/DOF algorithmic)") if xs_active else \
                "adjrhs_v3_kernel<3,4,2,14,168> (fused element kernel, DMMA contractions; 168 B/DOF algorithmic)"
    else:
        kname = "adjrhs_v2_kernel (fused element kernel; 168 B/DOF algorithmic)"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": ("profiles/traffic.json (ncu --set full capture of this kernel at this size; not "
                                   "re-measured in this run)") if traffic is not None else None,
                "kernel": kname,
                "gs_stage_level": int(xs_active), "gs_classes_in_pass": int(xs_left), "gs_classes_total": int(xs_total),
                "kernel_ms": elem_ms, "gs_ms": gs_ms, "algorithmic_bytes_per_launch": n * bpd,
                "peak_source": peak_src,
                "step_GBps": n * sem.algorithmic_bytes_per_dof(lx, True) / (ms_step * 1e-3) / 1e9}

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------
    cpu = None
    parity = None
    if rank == 0:
        px, py, pz = workloads.rank_grid(N)
        prob = cpu_sample_problem(ne, lx, args.cpu_sample_ne, px, py, pz)
        kept = {}
        if N == 1 and not args.no_cpu:
            cval, cms, cdone, threads = time_oracle(prob, 40, 2, budget_s=12.0, keep=kept)
            cpu = {"value": cval, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": prob["desc"] + f"; {cdone} steps of oracle/oracle.c, {cms:.1f} ms each"}
        # ---- parity of THIS run's result (every N): the last timed step against the oracle ------------
        parity = parity_vs_oracle(prob, f, sens, ne, oracle_out=kept or None)

    dealiased = None
    if rank == 0 and N == 1 and lx == 8 and not args.no_dealias:
        try:
            dealiased = dealiased_leg(args.dealias_ne, lx, dev)
        except Exception as ex:           # an extra: never takes the headline line down
            dealiased = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        line = {
            "metric": METRIC if lx == 8 else METRIC.replace("lx=8", f"lx={lx}"), "value": value, "unit": UNIT,
            "n_gpus": N, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args),
            "e2e": e2e, "gpu_launches": launches_all, "roofline": roofline, "cpu_baseline": cpu,
            "parity": parity, "clocks": clk.summary(), "checksum_abs_f": checksum, "host_cpus_bound": numa,
        }
        if dealiased is not None:
            line["dealiased"] = dealiased
        if phases:
            names = (["boundary_elements", "shared_gs", "pack", "interior_elements", "local_gs", "wait_recv+unpack"]
                     if os.environ.get("B200_EXCHANGE_OVERLAP") == "elem" else
                     ["elements", "shared_gs", "pack+exchange_issue", "local_gs", "wait_recv+unpack"])
            line["phase_ms"] = dict(zip(names, phases))
        print(json.dumps(line), flush=True)
    op.free()
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        print(f"bench.py: PARITY FAILED: {parity}", file=sys.stderr, flush=True)
        return 3
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
