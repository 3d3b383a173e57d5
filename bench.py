"""Benchmark of the discrete-GP hot path (BASELINE.json: "BO steps/s (rank-1
update + ITL argmax) at N=65k; HBM GB/s vs B200 peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c3|c3_small|c4|c4_small|c5] [--conditioning incremental|recompute]
                    [--storage lower|full] [--update-block B] [--arith fma|two_rounding] [--dtype f64|f32]

One step = score over all N candidates -> safe-set-masked arg-max -> rank-1
posterior update of the N x N covariance of every GP.

  c3  BASELINE.json configs[2] (the config the metric is quoted on): directed ITL with a fixed
      target set, N = 65,536, m = 4,096, d = 4, Matern-5/2 (l = 0.25), q = 1, rho = 0.1, fp64.
  c4  configs[3]: batched Safe BO, objective + 4 constraint GPs (q = 5), N = 65,536, VTL-PM and
      ITL-PM (the region of interest = the potential maximisers, recomputed every step), fused
      safe-set-masked arg-max; fp32 storage on one GPU (86 GB), fp64 row-sharded on >= 2.
  c5  configs[4]: N = 200,000, m = 8,192, fp64, covariance row-sharded over 8 GPUs, 500 steps.

  value  device-resident loop (index, observation table and state in HBM, no host
         synchronisation inside the timed region)
  e2e    the same steps through the drop-in API `alg.step(f)` with a HOST black box f: per
         step the arg-max record and the selected point are read back (D2H) and the
         observation is copied in from pinned memory (H2D)

With --gpus N > 1 (launched by torchrun, one rank per GPU) the covariance is row-sharded over
the ranks; the total problem is fixed ("scaling": "strong").  The default run (c3) also carries
sub-records of the other named configs that fit the box: `c4` always, `c5` at 8 GPUs.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "transductive-active-learning_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # BASELINE.json configs[2]
    "c3": dict(kind="itl", cfg=2, N=65536, m=4096, d=4, lengthscale=0.25, rho=0.1, beta=1.0, grid=16),
    "c3_small": dict(kind="itl", cfg=2, N=4096, m=256, d=4, lengthscale=0.25, rho=0.1, beta=1.0, grid=8),
    # BASELINE.json configs[4]: large-domain ITL, needs the 8 GPUs of one box (fp64: 320 GB of covariance)
    "c5": dict(kind="itl", cfg=4, N=200000, m=8192, d=4, lengthscale=0.2, rho=0.1, beta=1.0, grid=0),
    # BASELINE.json configs[3]: objective + 4 constraint GPs, potential-maximiser ROI
    "c4": dict(kind="safe", cfg=3, N=65536, d=4, q=5, grid=16, rho=0.1, beta=3.0, n_prior=10,
               lengthscales=[0.25, 0.2, 0.3, 0.25, 0.2]),
    "c4_small": dict(kind="safe", cfg=3, N=4096, d=4, q=5, grid=8, rho=0.1, beta=3.0, n_prior=10,
                     lengthscales=[0.25, 0.2, 0.3, 0.25, 0.2]),
}
METRIC = "BO steps/s (rank-1 update + ITL argmax) at N=65k"
UNIT = "steps/s"


# --------------------------------------------------------------------------- data
def make_problem(w):
    """Synthetic inputs of the named shape: one uniform point per cell of a
    grid^d lattice on [0,1)^d (on the bare lattice the reference's scalar-wise
    subset test, lib/utils.py:65-71, would pick the UNdirected rule), a fixed
    target set A, and a table of noisy observations."""
    g, d, N = w["grid"], w["d"], w["N"]
    rng = np.random.default_rng(0)
    if g:
        axes = np.meshgrid(*([np.arange(g) / g] * d), indexing="ij")
        domain = np.stack(axes, -1).reshape(-1, d)[:N] + rng.random((N, d)) / g
    else:  # uniform random domain
        domain = rng.random((N, d))
    A = np.sort(rng.choice(N, w["m"], replace=False))
    r1 = np.random.default_rng(1)
    noise = r1.standard_normal(N)
    table = (np.sin(3.0 * domain.sum(axis=1)) + w["rho"] * noise)[None, :]
    return domain, A, table


def make_problem_safe(w):
    """C4: the c3 domain, an objective sin(3 sum x) and four constraints g_i(x) = c - |x - z_i|^2
    around a safe seed (pattern: examples/unknown_constraints/bo/2d_island/experiment.py:67-87),
    a table of noisy observations [q, N] and the n_prior domain points nearest to the seed."""
    g, d, N, q = w["grid"], w["d"], w["N"], w["q"]
    rng = np.random.default_rng(0)
    axes = np.meshgrid(*([np.arange(g) / g] * d), indexing="ij")
    domain = np.stack(axes, -1).reshape(-1, d)[:N] + rng.random((N, d)) / g
    x0 = np.full(d, 0.5)
    z = x0 + 0.12 * rng.standard_normal((q - 1, d))
    cons = np.stack([0.6 - np.sum((domain - z[i]) ** 2, axis=1) for i in range(q - 1)])
    clean = np.concatenate([np.sin(3.0 * domain.sum(axis=1))[None], cons])
    table = clean + w["rho"] * np.random.default_rng(1).standard_normal((q, N))
    prior_idx = np.sort(np.argsort(np.sum((domain - x0) ** 2, axis=1))[: w["n_prior"]])
    return domain, table, prior_idx


# ------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


# -------------------------------------------------------------------- CPU baseline
def sample_workload(w, n_s):
    """The same problem family at N_s candidates (m_s = N_s m / N targets, same kernel / d / rho)."""
    ws = dict(w, N=n_s, m=max(8, n_s * w["m"] // w["N"]), grid=max(2, int(round(n_s ** (1.0 / w["d"])))))
    while ws["grid"] ** ws["d"] < n_s:
        ws["grid"] += 1
    return ws


def cpu_oracle_sample(w, n_s, steps, warmup=1):
    """Times the NumPy oracle (the reference's algorithm as written: full N x N posterior inside
    directed ITL, N x N x d sample-region mask and out-of-place N x N update every step) on a bounded
    sample of the workload.  `value` is what was MEASURED, at the sample size; the extrapolation to
    the full size by each phase's complexity (directed ITL ~ N^2 m, update and mask ~ N^2) is reported
    next to it, labelled."""
    import oracle as O
    ws = sample_workload(w, n_s)
    domain, A, table = make_problem(ws)
    noise = O.HomoscedasticNoise(1, np.array([w["rho"]]))
    distrs = O.prior_distr(noise, domain, np.zeros((0, w["d"])), np.zeros((1, 0)),
                           [O.Matern52(lengthscale=w["lengthscale"])], [O.ZeroMean()])
    model = O.MarginalModel(0, domain, distrs, beta=np.array([w["beta"]]), noise_rate=noise)
    t_itl = t_mask = t_upd = 0.0
    cpu0 = wall0 = None
    done = 0
    for t in range(warmup + steps):
        if t == warmup:
            cpu0, wall0 = time.process_time(), time.perf_counter()
        a = time.perf_counter()
        F = O.itl_directed(model, roi_idx=A)             # lib/algorithms/itl.py:33-60 (full N x N)
        b = time.perf_counter()
        mask = O.get_mask(model.domain, model.domain)     # lib/algorithms/__init__.py:80
        F = F.copy(); F[~mask] = -np.inf
        c = time.perf_counter()
        idx = model.objective_argmax_index(F)
        model.step_idx(np.array([idx]), table[:, [idx]], np.array([[w["rho"]]]))  # marginal.py:350-368
        e = time.perf_counter()
        if t >= warmup:
            t_itl += b - a; t_mask += c - b; t_upd += e - c; done += 1
    wall = time.perf_counter() - wall0
    cpu = time.process_time() - cpu0
    sN = (w["N"] / ws["N"]) ** 2
    sM = w["m"] / ws["m"]
    per_step_full = (t_itl * sN * sM + t_mask * sN + t_upd * sN) / done
    per_step_sample = wall / done
    return dict(
        value=1.0 / per_step_sample, unit=UNIT, cores=os.cpu_count(), kind="port",
        sample=(f"NumPy oracle (the reference's algorithm as written), {done} timed steps at N_s={ws['N']}, "
                f"m_s={ws['m']} (same kernel / d / rho): {per_step_sample * 1e3:.1f} ms/step measured "
                f"(directed ITL {t_itl / done * 1e3:.1f}, sample-region mask {t_mask / done * 1e3:.1f}, "
                f"arg-max + update {t_upd / done * 1e3:.1f}); `value` is this measured rate AT THE SAMPLE SIZE. "
                f"The reference as written cannot run N={w['N']} (N x N x d temporary, 3 x 34 GB matrices)."),
        sample_N=ws["N"], sample_m=ws["m"], ms_per_step_at_sample=per_step_sample * 1e3,
        extrapolated_to_full_size={"steps_per_s": 1.0 / per_step_full, "s_per_step": per_step_full,
                                   "how": "per phase: ITL x (N/N_s)^2 (m/m_s), mask and update x (N/N_s)^2; "
                                          "an extrapolation, not a measurement"},
        wall_s=wall, process_time_s=cpu)


def run_reference(args, w):
    """--impl reference: the reference's own CPU algorithm (oracle port: the JAX reference cannot be
    installed here) on the host cores.  Every step is a bounded sample of the workload (N_s candidates);
    the line's value is the measured rate at that sample size."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if w["kind"] != "itl":
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm times the ITL workloads (c3, c5)"}))
        return
    total = args.steps + args.warmup
    budget = 150.0 / max(1, total)  # seconds per step so that the run ends within minutes
    t0 = time.perf_counter()
    cpu_oracle_sample(w, 1024, 1, 0)
    t1024 = time.perf_counter() - t0
    n_s = 1024
    for cand in (2048, 4096, 8192):
        if t1024 * (cand / 1024.0) ** 3 <= budget:
            n_s = cand
    res = cpu_oracle_sample(w, n_s, args.steps, args.warmup)
    cfg = workload_config(args, w, 1)
    cfg["workload"] = (f"bounded sample of {cfg['workload']}: N_s={res['sample_N']}, m_s={res['sample_m']} "
                       f"(the reference's algorithm cannot run the full size)")
    cfg.update(sample_N=res["sample_N"], sample_m=res["sample_m"])
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / res["value"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "extrapolated_to_full_size": res["extrapolated_to_full_size"],
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------ helpers
def workload_config(args, w, world):
    s = 4 if getattr(args, "dtype", "f64") == "f32" else 8
    q = w.get("q", 1)
    if w["kind"] == "itl":
        desc = (f"BASELINE configs[{w['cfg']}]: synthetic transductive ITL, N={w['N']} candidates, target set "
                f"m={w['m']}, d={w['d']} Matern-5/2 (l={w['lengthscale']}), q=1, rho={w['rho']}, "
                f"{'fp32' if s == 4 else 'fp64'}")
    else:
        desc = (f"BASELINE configs[{w['cfg']}]: batched Safe BO, objective + {q - 1} constraint GPs, N={w['N']}, "
                f"d={w['d']} Matern-5/2 (l={w['lengthscales']}), rho={w['rho']}, beta={w['beta']}, "
                f"{w['n_prior']} prior samples at the safe seed, ROI = potential maximisers, "
                f"{'fp32' if s == 4 else 'fp64'}")
    cfg = {
        "workload": desc, "name": args.workload, "N": w["N"], "d": w["d"], "q": q,
        "update_block": int(os.environ.get("TAL_UPDATE_BLOCK", "16")),
        "update_arith": args.arith,
        "storage": ("full N x N covariance (reference layout)" if args.storage == "full" else
                    "lower-block: only the blocks on or below the diagonal are maintained (symmetric matrix)"),
        "l2": "inputs larger than L2 (covariance %.1f GB per GPU >> 126 MB): no flush needed"
              % (q * w["N"] * w["N"] * s / world / 1e9),
        "parallelism": "1 GPU" if world == 1 else f"covariance row-sharded over {world} GPUs",
    }
    if w["kind"] == "itl":
        cfg.update(m=w["m"], conditioning=args.conditioning)
    return cfg


def ncu_traffic(kernel_prefix, workload, n_gpus, storage="full", updates=None):
    """DRAM bytes per launch of a kernel from the committed ncu capture (profiles/*_kernel_traffic.json),
    or None when no capture of this workload / GPU count / instantiation is on file."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_kernel_traffic.json"))):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        if d.get("workload") != workload or d.get("n_gpus") != n_gpus or d.get("storage", "full") != storage:
            continue
        for name, k in d.get("kernels", {}).items():
            if name.startswith(kernel_prefix) and (updates is None or k.get("updates_per_launch", updates) == updates):
                best = k["dram_bytes"]
    return best


def _events():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _max_over_ranks(x, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(np.atleast_1d(np.asarray(x, dtype=np.float64)), device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()


def probe_fp64(lib, dev, kind):
    """Live fp64 peaks: kind 0 = DMMA.8x8x4 TFLOP/s, kind 1 = DFMA instructions / s (self-probed: there is
    no driver-side fp64 peak in MEASURED_PEAKS.json)."""
    import torch
    from talgp import _abi
    out = torch.zeros((1 << 20,), dtype=torch.float64, device=dev)
    blocks, iters = 148 * 8, 4096 if kind == 0 else 2000
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _abi.check(lib.tal_probe_fp64(kind, blocks, iters, C.c_void_p(out.data_ptr()), stream))
    e0, e1 = _events()
    e0.record()
    _abi.check(lib.tal_probe_fp64(kind, blocks, iters, C.c_void_p(out.data_ptr()), stream))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if kind == 0:
        return blocks * iters * 16 * 8 * 512 / (ms * 1e-3) / 1e12
    return blocks * iters * 16 * 256 / (ms * 1e-3)


def recompute_kernels(gp, roi_idx, m, hbm_peak):
    """Live timings (CUDA events) of the kernels of one full re-conditioning: TRSM and Cholesky against the
    self-probed DMMA.8x8x4 fp64 tensor peak (tal_probe_fp64), the gather against HBM."""
    import torch
    from talgp import _abi
    lib = gp.lib
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    big = m > 16384  # (C4 ITL-PM: seconds per call -- one timed call each, no dry run)

    def timed(fn, reps=3):
        if big:
            reps = 1
        else:
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = _events()
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    dmma_peak = probe_fp64(lib, gp.device, 0)
    m_pad, ncols = gp._cond_dims(m)
    S = gp._scratch("cv_S", (m_pad, m_pad)); B = gp._scratch("cv_B", (m_pad, ncols))
    work = gp._scratch("cv_work", (int(lib.tal_potrf_work_bytes(m_pad)) // 8,))
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    def gather():
        fn = lib.tal_gather_roi_lower if gp.lower else lib.tal_gather_roi
        _abi.check(fn(p(gp.cov[0]), gp.ld, gp.N, p(roi_idx), m, m_pad, 1e-5, p(S), m_pad,
                      p(B), ncols, gp.dt, stream))
    def potrf():
        gather()
        _abi.check(lib.tal_potrf_lower(p(S), m_pad, m_pad, p(work), stream))
    def trsm():
        _abi.check(lib.tal_trsm_lower(p(S), m_pad, m_pad, p(work), p(B), ncols, ncols, stream))
    t_gather = timed(gather)
    t_potrf = timed(potrf) - t_gather
    gather(); potrf()
    t_trsm = timed(trsm, reps=2)  # (solving in place repeatedly: same flops every time)
    src = "self-probed live: tal_probe_fp64 (independent DMMA.8x8x4 chains on all SMs); no driver-side fp64 peak exists"
    res = []
    fl = float(m_pad) * m_pad * ncols
    res.append({"kernel": "trsm_lower (L^-1 Sigma_AX, fp64 DMMA.8x8x4 mma.sync)", "bound": "tensor",
                "achieved": fl / (t_trsm * 1e-3) / 1e12, "peak": dmma_peak, "unit": "TFLOP/s",
                "frac": fl / (t_trsm * 1e-3) / 1e12 / dmma_peak, "flops_per_launch": fl, "ms_per_launch": t_trsm,
                "peak_source": src})
    fl = float(m_pad) ** 3 / 3
    res.append({"kernel": "potrf_lower (blocked Cholesky of Sigma_AA + jitter^2 I)", "bound": "tensor",
                "achieved": fl / (t_potrf * 1e-3) / 1e12, "peak": dmma_peak, "unit": "TFLOP/s",
                "frac": fl / (t_potrf * 1e-3) / 1e12 / dmma_peak, "flops_per_launch": fl, "ms_per_launch": t_potrf,
                "peak_source": src, "note": "latency-bound: serial chain of diagonal blocks"})
    gb = 2.0 * m * gp.ld * 8
    res.append({"kernel": "gather_roi (Sigma[A, :] rows -> dense B, Sigma_AA)", "bound": "hbm",
                "achieved": gb / (t_gather * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": gb / (t_gather * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": gb,
                "ms_per_launch": t_gather})
    return res


def pass_rooflines(gp, pass_ms, pass_updates, hbm_peak, peak_src, args, world, total_ms, dfma_per_s):
    """One roofline entry PER KERNEL INSTANTIATION of the covariance pass: the passes of the timed region are
    grouped by the number of updates they applied (rankb<16>, rankb<4>, ... are different kernels); ms is the
    mean over the launches of that instantiation (max over ranks per launch)."""
    alg_bytes = gp.update_bytes()
    s = 4 if args.dtype == "f32" else 8
    elems = alg_bytes / (2.0 * s)
    ops_per_update = 1.0 if args.arith == "fma" else 2.0
    groups = {}
    for ms, up in zip(pass_ms, pass_updates):
        groups.setdefault(int(up), []).append(ms)
    out = []
    for up, mss in sorted(groups.items(), reverse=True):
        ms = float(np.mean(mss))
        slots = 1 if gp.update_block == 1 else max(2, 1 << (max(up, 1) - 1).bit_length())
        achieved = alg_bytes / (ms * 1e-3) / 1e9
        if gp.update_block == 1:
            name = "rank1_update (tal_rank1_update%s)" % ("_lower" if gp.lower else "")
            prefix = "rank1_"
        else:
            name = ("rankb_kernel<%s, P=%d%s>: %d deferred rank-1 updates applied in one pass (tal_rankb_update, "
                    "%s storage)" % ("double" if s == 8 else "float", slots, ", FMA" if args.arith == "fma" else "",
                                     up, args.storage))
            prefix = "rankb_"
        e = {"kernel": name + (", per GPU, max over ranks" if world > 1 else ""), "bound": "hbm",
             "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
             "peak_source": peak_src,
             "traffic": ncu_traffic(prefix, args.workload, world, args.storage, slots if gp.update_block > 1 else None),
             "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms, "launches_in_timed_region": len(mss),
             "updates_per_launch": up, "instantiation_slots": slots,
             "share_of_step": float(np.sum(mss)) / total_ms, "frac_of_8TBs_datasheet": achieved / 8000.0,
             "effective_GBs_vs_full_matrix_bytes": 2.0 * gp.q * gp.N * gp.N * s / world / (ms * 1e-3) / 1e9}
        if dfma_per_s:
            ops = ops_per_update * slots * elems
            e["second_bound"] = {
                "bound": "fp64 issue (self-probed live: tal_probe_fp64, independent DFMA chains)",
                "peak_instr_per_s": dfma_per_s, "fp64_ops_per_pass": ops, "min_ms_per_pass": ops / dfma_per_s * 1e3,
                "hbm_ms_per_pass": alg_bytes / (hbm_peak * 1e9) * 1e3,
                "note": "%s arithmetic: %d fp64 instruction(s) per update and element" % (args.arith, int(ops_per_update))}
        out.append(e)
    return out


# ----------------------------------------------------------------------- ITL arm
def build_itl(w, args, dev, comm):
    import torch
    from lib.algorithms import index_roi_constructor
    from lib.algorithms.itl import ITL
    from lib.gp.kernels import stationary
    from lib.gp.means import ZeroMean
    from lib.gp.prior import prior_distr
    from lib.model.marginal import MarginalModel
    from lib.noise import HomoscedasticNoise
    domain, A, table = make_problem(w)
    nz = HomoscedasticNoise(1, np.array([w["rho"]]))
    distrs = prior_distr(nz, domain, np.zeros((0, w["d"])), np.zeros((1, 0)),
                         [stationary.Matern52(lengthscale=w["lengthscale"])], [ZeroMean()],
                         storage=args.storage, dtype=torch.float32 if args.dtype == "f32" else torch.float64,
                         comm=comm, arith=args.arith)
    model = MarginalModel(0, domain, distrs, beta=np.array([w["beta"]]), noise_rate=nz)
    alg = ITL(model, index_roi_constructor(A), conditioning=args.conditioning)
    alg.gather_F = False  # row-sharded: result["F"] stays this rank's candidates (no all-gather per step)
    alg.sample_mask()     # static; computed once (lib/algorithms/__init__.py:80)
    return domain, A, table, model, alg


def run_itl(args, w, dev, world, rank, comm, lib, hbm_peak, peak_src, steps, warmup, want_e2e=True,
            want_kernels=True):
    """The ITL workloads (c3, c5): returns the fields of the JSON line (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from lib.function import Function
    from talgp import _abi
    domain, A, table, model, alg = build_itl(w, args, dev, comm)
    gp = model._gp
    N, beta = w["N"], [w["beta"]]
    table_dev = torch.as_tensor(table, device=dev)
    roi = alg.roi
    inc = args.conditioning == "incremental"
    sample = alg.sample_mask()
    use_ctx = inc and gp._ctx_usable()
    ev = []

    def dev_step(record):
        """One step with index, observation table and state on the device.  Incremental mode: the whole launch
        sequence is ONE host call (tal_ctx_step: 3 fused kernels, 4 when sharded); the first step (and a step
        after the conditioning factor filled up) conditions on the target set from scratch."""
        st = gp._cond
        if use_ctx and st is not None and st["valid"]:
            gp.ctx_step(table_dev, N, beta, model.first_constraint, None, sample)
            return
        F = gp.score_itl_directed(roi.idx, roi.m, incremental=inc, roi_key=roi)
        rec = gp.argmax_candidates(F, None, model.first_constraint, sample)
        if record and not use_ctx:  # recompute mode: time the pass with events around the update
            idx_dev = rec[0:1]
            gp.exchange_row(idx_dev)
            gp.step_vectors(idx_dev, 0, table_dev, N, True, beta)
            e0, e1 = _events()
            e0.record()
            p = gp._pending + 1 if gp.update_block > 1 else 1
            passed = gp.rank1_update()
            e1.record()
            if passed:
                ev.append((e0, e1, p))
        else:
            gp.observe(rec[0:1], table_dev, beta, y_gather=True, y_stride=N)

    # a sharded step takes ~0.2 ms: warm up for a few hundred of them so that the clock sampler (one
    # nvidia-smi sample per 50 ms) sees the loaded GPU before and during the timed region
    warm_run = max(warmup, 200) if world > 1 else warmup
    for _ in range(warm_run):
        dev_step(False)
    gp.flush()  # the timed region starts and ends with no update pending: exactly K steps of work
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l_start, u_start = gp.l.clone(), gp.u.clone()
    lib.tal_reset_launch_count()
    ctx = gp._ctx_get() if use_ctx else None
    if ctx is not None:
        lib.tal_ctx_timing_begin(C.byref(ctx))  # CUDA events around every pass over the matrix
    t0, t1 = _events()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(steps):
        dev_step(True)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if ctx is not None:  # the updates still pending are applied inside the timed region
        ctx = gp._ctx_get()
        _abi.check(lib.tal_ctx_flush(C.byref(ctx), stream))
        gp._ctx_done(ctx)
    elif gp._pending:
        e0, e1 = _events()
        p = gp._pending
        e0.record(); gp.flush(); e1.record()
        ev.append((e0, e1, p))
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = float(_max_over_ranks(t0.elapsed_time(t1), dev, world)[0])
    launches = int(lib.tal_launch_count())
    breakdown = None
    if os.environ.get("TAL_BREAKDOWN") and ctx is not None:
        bd, nbd = (C.c_double * 6)(), C.c_longlong(0)
        _abi.check(lib.tal_ctx_breakdown(bd, C.byref(nbd)))
        bdv = _max_over_ranks([float(x) for x in bd], dev, world)
        breakdown = dict(zip(("peer_push_us", "fused_observe_us", "fused_partial_us", "fused_finish_us",
                              "pass_amortised_us", "gap_to_next_step_us"), [float(x) for x in bdv]), steps=int(nbd.value),
                         note="CUDA events between the launches of tal_ctx_observe (max over ranks per phase); "
                              "the events themselves cost ~1-2 us per phase")
    if ctx is not None:
        ms_buf, up_buf, n_buf = (C.c_double * 64)(), (C.c_longlong * 64)(), C.c_longlong(0)
        _abi.check(lib.tal_ctx_timing_read(C.byref(ctx), ms_buf, up_buf, C.byref(n_buf), stream))
        pass_ms = [float(ms_buf[i]) for i in range(int(n_buf.value))]
        pass_updates = [int(up_buf[i]) for i in range(int(n_buf.value))]
    else:
        pass_ms = [a.elapsed_time(b) for a, b, _ in ev]
        pass_updates = [p for _, _, p in ev]
    if pass_ms:
        pass_ms = [float(x) for x in _max_over_ranks(pass_ms, dev, world)]
    # the scoring pass (conditioning update), timed after the loop on scratch copies
    cond_ms = None
    if want_kernels and inc and gp._cond is not None and gp._cond["valid"] and gp._cond["append"]:
        st = gp._cond
        p_ = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        cs_tmp = st["cs"].clone()
        idx_dev = gp.argmax_record()[0:1]
        pcol = st["pcol"]

        def cond_once():
            _abi.check(lib.tal_cond_extract_col(p_(st["V"][0]), st["ncols"], st["rows"], gp.cand0, gp.n_cand,
                                                p_(idx_dev), 0, p_(pcol), stream))
            _abi.check(lib.tal_cond_append_update(
                p_(st["V"][0]), st["ncols"], st["cap"], st["m_pad"], st["rows"], st["ncols"], gp.n_cand, gp.cand0,
                p_(pcol), p_(gp.kbuf[0]), p_(gp.wbuf[0]), p_(gp.rho[0]), p_(idx_dev), 0, p_(cs_tmp[0]),
                p_(st["app"]), st["n_split"], gp.dt, stream))
        if st["rows"] + 2 <= st["cap"]:
            ts = []
            for _ in range(6):
                c0, c1 = _events()
                c0.record(); cond_once(); c1.record(); torch.cuda.synchronize()
                ts.append(c0.elapsed_time(c1))
            cond_ms = float(np.median(ts[1:]))
    status = gp.read_result(gp.argmax_record()).status
    assert status == 0, f"arg-max status {status}"
    # size-independent invariants of the state after the timed steps (full size, no oracle needed)
    rows = torch.arange(gp.row0, gp.row0 + gp.n_loc, device=dev)
    diag_ok = torch.equal(gp.cov[0, torch.arange(gp.n_loc, device=dev), rows].to(torch.float64), gp.var[0, rows])
    ok_local = (diag_ok and bool(torch.isfinite(gp.mean).all()) and bool(torch.isfinite(gp.var).all())
                and bool((gp.var > 0).all()) and bool((gp.l >= l_start).all()) and bool((gp.u <= u_start).all()))
    invariants = {"var_equals_diag_cov_bitwise__state_finite__variances_positive__bounds_are_running_intersections":
                  bool(_max_over_ranks(0.0 if ok_local else 1.0, dev, world)[0] == 0.0)}
    if world > 1:
        sel = gp.argmax_record()[0:1].clone()
        sel_all = [torch.zeros_like(sel) for _ in range(world)]
        dist.all_gather(sel_all, sel)
        digest = torch.stack([gp.mean.sum(), gp.var.sum(), gp.l.sum(), gp.u.sum()])
        dig_all = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(dig_all, digest)
        invariants["ranks_agree_on_selection"] = all(bool(torch.equal(s_, sel_all[0])) for s_ in sel_all)
        invariants["replicated_state_identical"] = all(bool(torch.equal(d_, dig_all[0])) for d_ in dig_all)
    assert all(invariants.values()), invariants
    exchange = None
    if world > 1:
        exchange = "nccl"
        if getattr(gp.comm, "peer", False):
            exchange = "nvlink-peer-memory"
            err = gp.comm.error()
            assert err == 0, f"peer exchange timed out (TAL_PEER_ERR {err})"

    # ---- end to end through the drop-in API with a host black box ---------
    e2e = None
    if want_e2e:
        class HostFunction(Function):
            q, first_constraint = 1, 1
            host_input = True  # X arrives as a host array (the model's host copy of the domain row)

            def observe(self, X):  # X: [1, d] device tensor -> D2H; returns host array [q, 1]
                x = X.cpu().numpy()
                i = model.last_argmax.idx
                noise = table[0, i] - np.sin(3.0 * domain[i].sum())  # the pre-drawn noise of point i
                return np.array([[np.sin(3.0 * x.sum()) + noise]])

        f = HostFunction()
        for _ in range(max(1, min(warmup, 3))):
            alg.step(f)
        gp.flush()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        w0 = time.perf_counter()
        a0, a1 = _events()
        a0.record()
        for _ in range(steps):
            out = alg.step(f)
        gp.flush()
        fmax = float(out["F_max"])  # D2H read of the step's result
        a1.record()
        torch.cuda.synchronize()
        e2e_ms = max(a0.elapsed_time(a1), (time.perf_counter() - w0) * 1e3)
        e2e_ms = float(_max_over_ranks(e2e_ms, dev, world)[0])
        assert np.isfinite(fmax)
        e2e = {"value": steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": 8 * model.q,
               "d2h_bytes_per_step": 32,  # the arg-max record (the selected point is a row of the host copy of the domain)
               "ms_per_step": e2e_ms / steps,
               "api": "lib.algorithms.itl.ITL.step(f) with a host black box on every rank"}
    if rank != 0:
        return None
    dfma = None
    if gp.update_block > 1:
        try:
            dfma = probe_fp64(lib, dev, 1)
        except Exception:  # pragma: no cover
            dfma = None
    roofs = pass_rooflines(gp, pass_ms, pass_updates, hbm_peak, peak_src, args, world, total_ms, dfma)
    line = {
        "value": steps / (total_ms / 1e3), "ms_per_step": total_ms / steps, "steps": steps,
        "config": dict(workload_config(args, w, world), **({"exchange": exchange} if exchange else {})),
        "gpu_launches": launches, "invariants": invariants, **({"breakdown_us": breakdown} if breakdown else {}),
        "roofline": roofs[0] if roofs else None,
        "roofline_other_instantiations": roofs[1:],
        "kernels": [],
    }
    if world > 1:
        line["warmup_steps_run"] = warm_run
    if e2e is not None:
        line["e2e"] = e2e
    if cond_ms is not None:
        st = gp._cond
        cbytes = model.q * (st["rows"] + 2.0) * st["ncols"] * 8  # one read of the factor + two appended rows
        line["kernels"].append({
            "kernel": "cond_app_partial_kernel / fused_partial_kernel (incremental target-set conditioning = the ITL "
                      "scoring pass; read-only GEMV over the factor)",
            "bound": "hbm", "achieved": cbytes / (cond_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": cbytes / (cond_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": cbytes,
            "ms_per_launch": cond_ms, "launches_per_step": 1,
            "traffic": ncu_traffic("cond_app_partial", args.workload, world, args.storage),
            "share_of_step": cond_ms / (total_ms / steps),
            "timed": "after the loop, CUDA events, same buffers (scratch copy of cs), unfused form of the same pass"})
    if want_kernels and not inc and world == 1:
        line["kernels"] += recompute_kernels(gp, roi.idx, roi.m, hbm_peak)
    return line


# ---------------------------------------------------------------------- Safe-BO arm
def run_safe(args, w, dev, world, rank, comm, lib, hbm_peak, peak_src, steps, warmup, rules=("VTL-PM", "ITL-PM"),
             itl_steps=2):
    """C4: batched Safe BO through the drop-in API (device-resident black box): every step computes the
    potential-maximiser ROI from the confidence bounds of the 5 GPs, scores all candidates (VTL-PM, then
    ITL-PM in a second model), takes the safe-set-masked arg-max (the pessimistic safe set is fused into
    the arg-max kernel) and updates all 5 covariances.

    ITL-PM: with 10 observations at the seed the potential maximisers are almost the whole domain
    (m ~ N - 10), and the safe set is not inside them, so the reference's rule is the DIRECTED one
    (lib/algorithms/itl.py:22-25): per step and GP a Cholesky of an m x m and a triangular solve of an
    m x N matrix, m^3/3 + m^2 N = 3.8e14 fp64 flop at N = 65,536 (the reference: 2 N^2 m).  It is timed for
    `itl_steps` steps without warm-up (every step re-conditions from scratch: there is nothing to warm)."""
    import torch
    import torch.distributed as dist
    from lib.algorithms import pm_roi_constructor
    from lib.algorithms.itl import ITL
    from lib.algorithms.vtl import VTL
    from lib.function import TableFunction
    from lib.gp.kernels import stationary
    from lib.gp.means import ZeroMean
    from lib.gp.prior import prior_distr
    from lib.model.marginal import MarginalModel
    from lib.noise import HomoscedasticNoise
    domain, table, prior_idx = make_problem_safe(w)
    q, N = w["q"], w["N"]
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    out = {}
    for rule in rules:
        is_itl = rule == "ITL-PM"
        steps_r, warmup_r = (max(1, min(steps, itl_steps)), 0) if (is_itl and N > 16384) else (steps, warmup)
        noise = HomoscedasticNoise(q, np.full(q, w["rho"]))
        distrs = prior_distr(noise, domain, domain[prior_idx], table[:, prior_idx],
                             [stationary.Matern52(lengthscale=l) for l in w["lengthscales"]], [ZeroMean()] * q,
                             dtype=dtype, storage=args.storage, comm=comm, arith=args.arith)
        model = MarginalModel(0, domain, distrs, beta=np.full(q, w["beta"]), noise_rate=noise)
        gp = model._gp
        alg = VTL(model, pm_roi_constructor) if rule == "VTL-PM" else ITL(model, pm_roi_constructor)
        alg.gather_F = False
        f = TableFunction(model.domain, table)
        ms_roi = []
        for _ in range(warmup_r):
            alg.step(f)
        gp.flush()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        lib.tal_reset_launch_count()
        t0, t1 = _events()
        t0.record()
        for _ in range(steps_r):
            ms_roi.append(alg.roi.m if steps_r <= 64 else 0)
            alg.step(f)
        gp.flush()
        t1.record()
        torch.cuda.synchronize()
        total_ms = float(_max_over_ranks(t0.elapsed_time(t1), dev, world)[0])
        launches = int(lib.tal_launch_count())
        # one posterior update of the 5 covariances, timed alone (the dominant kernel of this workload)
        idx_dev = gp.argmax_record()[0:1].clone()
        ts = []
        for _ in range(3):
            gp.flush()
            gp.exchange_row(idx_dev)
            gp.step_vectors(idx_dev, 0, f.y_table, N, True, [w["beta"]] * q)
            e0, e1 = _events()
            e0.record(); gp._rank1_now(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        upd_ms = float(_max_over_ranks(float(np.median(ts)), dev, world)[0])
        pess = model.pessimistic_safe_set
        r = model.last_argmax
        ok = (bool(torch.isfinite(gp.mean).all()) and bool((gp.var > 0).all()) and r.status == 0
              and bool(pess[r.idx]))
        assert ok, "C4 invariants"
        if world > 1 and getattr(gp.comm, "peer", False):
            assert gp.comm.error() == 0
        alg_bytes = gp.update_bytes()
        kernels = []
        if is_itl and world == 1 and alg.roi.m > 0:
            try:
                kernels = recompute_kernels(gp, alg.roi.idx, alg.roi.m, hbm_peak)
            except Exception as exc:  # pragma: no cover
                kernels = [{"error": repr(exc)[:200]}]
        out[rule] = {
            "kernels": kernels,
            "value": steps_r / (total_ms / 1e3), "unit": UNIT, "ms_per_step": total_ms / steps_r, "steps": steps_r,
            "warmup": warmup_r, "roi_size_per_step": ms_roi, "safe_set_size": int(pess.sum()), "gpu_launches": launches,
            "roofline": {"kernel": "rank1_update%s over the %d covariances (one pass per step: the scoring reads "
                                   "the matrix)" % ("_lower" if gp.lower else "", q),
                         "bound": "hbm", "achieved": alg_bytes / (upd_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": alg_bytes / (upd_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src, "traffic": None,
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": upd_ms,
                         "share_of_step": upd_ms / (total_ms / steps_r)},
            "invariants": {"state_finite__variances_positive__selection_in_pessimistic_safe_set": ok},
        }
        del model, alg, gp, distrs, f
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    if rank != 0:
        return None
    first = out[rules[0]]
    return {"config": workload_config(args, w, world), "rules": out,
            "value": first["value"], "ms_per_step": first["ms_per_step"],
            "api": "lib.algorithms.{vtl.VTL, itl.ITL}(model, pm_roi_constructor).step(f), device-resident black box "
                   "(TableFunction); per step one D2H of the ROI size and one of the arg-max record"}


# ----------------------------------------------------------------------- GPU arm
def run_ours(args, w):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from talgp import _abi
    lib = _abi.load()
    _abi.require_device()
    comm = None
    if world > 1:
        from talgp.dist import default_comm
        comm = default_comm()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"

    clocks = ClockSampler(local)
    clocks.start()  # sampled from the warm-up on: the timed region alone can be shorter than one sample
    head = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic"}
    if w["kind"] == "safe":
        body = run_safe(args, w, dev, world, rank, comm, lib, hbm_peak, peak_src, args.steps, args.warmup,
                        rules=tuple(args.rules.split(",")), itl_steps=args.itl_steps)
        clk = clocks.stop()
        if rank == 0:
            r = list(body["rules"].values())[0]
            line = dict(head, **body)
            line["roofline"] = r["roofline"]
            line["gpu_launches"] = r["gpu_launches"]
            line["e2e"] = {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 40,
                           "note": "the C4 loop already runs through the public API (device-resident black box)"}
            line["clocks"] = {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    body = run_itl(args, w, dev, world, rank, comm, lib, hbm_peak, peak_src, args.steps, args.warmup)
    clk = clocks.stop()
    line = None
    if rank == 0:
        line = dict(head, **body)
        line["clocks"] = {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]}
    import gc
    gc.collect()  # (model <-> ROI reference cycles keep the 34 GB matrix alive otherwise)
    torch.cuda.empty_cache()

    # ---- sub-records of the other named configs (same process, the c3 model is freed) -------------
    if args.subrecords and args.workload == "c3":
        sub = argparse.Namespace(**vars(args))
        sub.workload, sub.conditioning = "c4", "incremental"
        if world == 1:
            sub.dtype = "f32"  # 5 x 34 GB in fp64 does not fit one GPU
        try:
            # (ITL-PM at this size is 3.8e14 fp64 flop per GP and step -- tens of seconds: it has its own run,
            # `bench.py --workload c4`, whose line is kept under profiles/)
            c4 = run_safe(sub, WORKLOADS["c4"], dev, world, rank, comm, lib, hbm_peak, peak_src,
                          steps=min(args.steps, 16), warmup=2, rules=("VTL-PM",))
        except Exception as exc:  # pragma: no cover
            c4 = {"error": repr(exc)[:300]}
        if rank == 0:
            line["c4"] = c4
        gc.collect()
        torch.cuda.empty_cache()
        if world == 8:
            sub = argparse.Namespace(**vars(args))
            sub.workload, sub.conditioning, sub.dtype = "c5", "incremental", "f64"
            try:
                c5 = run_itl(sub, WORKLOADS["c5"], dev, world, rank, comm, lib, hbm_peak, peak_src, steps=500,
                             warmup=3, want_e2e=False, want_kernels=True)
            except Exception as exc:  # pragma: no cover
                c5 = {"error": repr(exc)[:300]}
            if rank == 0:
                line["c5"] = c5
            torch.cuda.empty_cache()

    # beside the headline: variants of the same workload, each a sub-run (one GPU only)
    if args.both and world == 1 and rank == 0:
        variants = []
        inc = args.conditioning == "incremental"
        common = ["--storage", args.storage, "--arith", args.arith]
        if inc:
            variants.append(("value_recompute", ["--conditioning", "recompute"] + common))
        if args.arith == "fma":
            variants.append(("value_two_rounding", ["--conditioning", args.conditioning, "--storage", args.storage,
                                                    "--arith", "two_rounding"]))
        if args.storage == "lower":
            variants.append(("value_full_storage", ["--conditioning", args.conditioning, "--storage", "full",
                                                    "--arith", args.arith]))
        if int(os.environ.get("TAL_UPDATE_BLOCK", "16")) > 1:
            variants.append(("value_update_every_step", ["--conditioning", args.conditioning, "--update-block", "1"] + common))
        if args.dtype == "f64":
            variants.append(("value_fp32_storage", ["--conditioning", args.conditioning, "--dtype", "f32"] + common))
        for key, extra in variants:
            steps = max(3, args.steps // 4) if "recompute" in extra else args.steps
            sub = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", str(steps), "--warmup", "3",
                                  "--workload", args.workload, "--no-both", "--no-cpu", "--no-subrecords"] + extra,
                                 capture_output=True, text=True)
            try:
                o = json.loads(sub.stdout.strip().splitlines()[-1])
                line[key] = {"value": o["value"], "ms_per_step": o["ms_per_step"], "e2e": o["e2e"]["value"],
                             "roofline": {k: o["roofline"][k] for k in ("kernel", "achieved", "frac", "ms_per_launch",
                                                                        "algorithmic_bytes_per_launch")},
                             "kernels": o.get("kernels", [])}
            except Exception as exc:  # pragma: no cover
                line[key] = {"error": str(exc), "stderr": sub.stderr[-300:]}
    if rank == 0 and not args.no_cpu and world == 1:
        # the CPU port next to it, and the GPU arm at the SAME sample size: a measured / measured ratio
        cpu = cpu_oracle_sample(w, args.cpu_sample_n, 3, 1)
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "sample_N", "sample_m",
                                                    "extrapolated_to_full_size")}
        try:
            ws = sample_workload(w, args.cpu_sample_n)
            sub = argparse.Namespace(**vars(args))
            same = run_itl(sub, ws, dev, 1, 0, None, lib, hbm_peak, peak_src, steps=max(args.steps, 32), warmup=3,
                           want_e2e=True, want_kernels=False)
            line["value_at_cpu_sample_size"] = {"N": ws["N"], "m": ws["m"], "value": same["value"],
                                                "e2e": same["e2e"]["value"],
                                                "ratio_to_cpu_port_measured": same["e2e"]["value"] / cpu["value"]}
        except Exception as exc:  # pragma: no cover
            line["value_at_cpu_sample_size"] = {"error": repr(exc)[:300]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--conditioning", default="incremental", choices=["incremental", "recompute"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--storage", default=os.environ.get("TAL_STORAGE", "lower"), choices=["full", "lower"],
                    help="covariance storage: full N x N (reference layout) or lower-block (symmetric)")
    ap.add_argument("--update-block", type=int, default=None,
                    help="covariance updates applied per pass over the matrix (default 16; 1 = after every observation)")
    ap.add_argument("--arith", default=os.environ.get("TAL_UPDATE_ARITH", "fma"), choices=["fma", "two_rounding"],
                    help="update arithmetic: one fused multiply-subtract per entry (fma) or product rounded first")
    ap.add_argument("--dtype", default=None, choices=["f64", "f32"],
                    help="covariance storage / arithmetic type (the reference's only mode is f64)")
    ap.add_argument("--rules", default="VTL-PM,ITL-PM", help="c4: acquisition rules to time")
    ap.add_argument("--itl-steps", type=int, default=2, help="c4 at full size: timed ITL-PM steps (tens of seconds each)")
    ap.add_argument("--cpu-sample-n", type=int, default=8192)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-both", dest="both", action="store_false", help="skip the variant sub-runs")
    ap.add_argument("--no-subrecords", dest="subrecords", action="store_false", help="skip the c4 / c5 sub-records")
    args = ap.parse_args()
    if args.update_block is not None:
        os.environ["TAL_UPDATE_BLOCK"] = str(args.update_block)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.dtype is None:
        args.dtype = "f32" if (w["kind"] == "safe" and world == 1 and w["N"] > 16384) else "f64"
    if args.impl == "reference":
        run_reference(args, w)
    else:
        if world > 1:
            args.no_cpu = True
            args.both = False
        run_ours(args, w)


if __name__ == "__main__":
    main()
