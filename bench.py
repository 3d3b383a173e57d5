"""Benchmark of the discrete-GP hot path (BASELINE.json: "BO steps/s (rank-1
update + ITL argmax) at N=65k; HBM GB/s vs B200 peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--conditioning incremental|recompute] [--workload c3|c3_small]

One step = directed-ITL score over all N candidates (fixed target set A,
|A| = m) -> safe-set-masked arg-max -> rank-1 posterior update of the N x N
covariance.  Workload "c3" is BASELINE.json configs[2]: N = 65,536, m = 4,096,
d = 4, Matérn-5/2 (l = 0.25), q = 1, rho = 0.1, fp64, synthetic data.

  value  device-resident loop (index, observation table and state in HBM, no
         host synchronisation inside the timed region)
  e2e    the same steps through the drop-in API `ITL.step(f)` with a HOST black
         box f: per step the arg-max record and the selected point are read
         back (D2H) and the observation is copied in from pinned memory (H2D)

With --gpus N > 1 (launched by torchrun, one rank per GPU) the covariance is
row-sharded over the ranks (talgp/dist.py); the total problem is fixed
("scaling": "strong").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "transductive-active-learning_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # BASELINE.json configs[2]
    "c3": dict(N=65536, m=4096, d=4, lengthscale=0.25, rho=0.1, beta=1.0, grid=16),
    "c3_small": dict(N=4096, m=256, d=4, lengthscale=0.25, rho=0.1, beta=1.0, grid=8),
    # BASELINE.json configs[4]: large-domain ITL, needs the 8 GPUs of one box (fp64: 320 GB of covariance)
    "c5": dict(N=200000, m=8192, d=4, lengthscale=0.2, rho=0.1, beta=1.0, grid=0),
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
def cpu_oracle_sample(w, n_s, steps, warmup=1):
    """Times the NumPy oracle (the reference's algorithm as written) on a bounded
    sample of the workload: the same problem family at N_s candidates with
    m_s = N_s / 16 targets, and extrapolates each phase to the full size by its
    complexity (directed ITL ~ N^2 m, posterior update and sample-region mask ~ N^2)."""
    import oracle as O
    ws = dict(w, N=n_s, m=max(8, n_s * w["m"] // w["N"]), grid=max(2, int(round(n_s ** (1.0 / w["d"])))))
    while ws["grid"] ** ws["d"] < n_s:
        ws["grid"] += 1
    domain, A, table = make_problem(ws)
    noise = O.HomoscedasticNoise(1, np.array([w["rho"]]))
    distrs = O.prior_distr(noise, domain, np.zeros((0, w["d"])), np.zeros((1, 0)),
                           [O.Matern52(lengthscale=w["lengthscale"])], [O.ZeroMean()])
    model = O.MarginalModel(0, domain, distrs, beta=np.array([w["beta"]]), noise_rate=noise)
    t_itl = t_mask = t_upd = 0.0
    cpu0 = time.process_time()
    wall0 = time.perf_counter()
    done = 0
    for t in range(warmup + steps):
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
    per_step_sample = (t_itl + t_mask + t_upd) / done
    return dict(
        value=1.0 / per_step_full, unit=UNIT, cores=os.cpu_count(), kind="port",
        sample=(f"NumPy oracle, {done} steps at N_s={ws['N']}, m_s={ws['m']} (same kernel/d/rho): "
                f"{per_step_sample * 1e3:.1f} ms/step measured (ITL {t_itl / done * 1e3:.1f}, mask "
                f"{t_mask / done * 1e3:.1f}, update+argmax {t_upd / done * 1e3:.1f}); extrapolated to N={w['N']}, "
                f"m={w['m']} by N^2*m / N^2 / N^2 -> {per_step_full:.1f} s/step. The reference as written "
                f"cannot run the full size (N x N x d temporary, 3 x 34 GB matrices)."),
        measured_ms_per_step_at_sample=per_step_sample * 1e3, wall_s=wall, process_time_s=cpu)


def run_reference(args, w):
    """--impl reference: the reference's own CPU algorithm (oracle port: the
    JAX reference cannot be installed here) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    budget = 150.0 / max(1, total)  # seconds per step so that the run ends within minutes
    # calibrate at a small size, then pick the largest sample that fits the budget (cost ~ N^3)
    t0 = time.perf_counter()
    cpu_oracle_sample(w, 1024, 1, 0)
    t1024 = time.perf_counter() - t0
    n_s = 1024
    for cand in (2048, 4096, 8192):
        if t1024 * (cand / 1024.0) ** 3 <= budget:
            n_s = cand
    res = cpu_oracle_sample(w, n_s, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / res["value"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, w, 1),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def recompute_kernels(gp, roi, hbm_peak):
    """Live timings (CUDA events) of the kernels of one full re-conditioning: TRSM and Cholesky against the
    measured DMMA.8x8x4 fp64 tensor peak (tal_probe_fp64), the column-sum-of-squares pass against HBM."""
    import ctypes as C
    import torch
    from talgp import _abi
    lib = gp.lib
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    # fp64 tensor peak: 148*8 CTAs x 256 threads of independent DMMA chains
    out = torch.zeros((1 << 20,), dtype=torch.float64, device=gp.device)
    blocks, iters = 148 * 8, 4096
    ms = timed(lambda: _abi.check(lib.tal_probe_fp64(0, blocks, iters, C.c_void_p(out.data_ptr()), stream)))
    dmma_peak = blocks * iters * 16 * 8 * 512 / (ms * 1e-3) / 1e12
    m_pad, ncols = gp._cond_dims(roi.m)
    S = gp._scratch("cv_S", (m_pad, m_pad)); B = gp._scratch("cv_B", (m_pad, ncols))
    work = gp._scratch("cv_work", (int(lib.tal_potrf_work_bytes(m_pad)) // 8,))
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    def gather():
        fn = lib.tal_gather_roi_lower if gp.lower else lib.tal_gather_roi
        _abi.check(fn(p(gp.cov[0]), gp.ld, gp.N, p(roi.idx), roi.m, m_pad, 1e-5, p(S), m_pad,
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
    res = []
    fl = float(m_pad) * m_pad * ncols
    res.append({"kernel": "trsm_lower (L^-1 Sigma_AX, fp64 DMMA.8x8x4 mma.sync)", "bound": "tensor",
                "achieved": fl / (t_trsm * 1e-3) / 1e12, "peak": dmma_peak, "unit": "TFLOP/s",
                "frac": fl / (t_trsm * 1e-3) / 1e12 / dmma_peak, "flops_per_launch": fl, "ms_per_launch": t_trsm,
                "peak_source": "measured live: tal_probe_fp64 (independent DMMA.8x8x4 chains, all SMs)"})
    fl = float(m_pad) ** 3 / 3
    res.append({"kernel": "potrf_lower (blocked Cholesky of Sigma_AA + jitter^2 I)", "bound": "tensor",
                "achieved": fl / (t_potrf * 1e-3) / 1e12, "peak": dmma_peak, "unit": "TFLOP/s",
                "frac": fl / (t_potrf * 1e-3) / 1e12 / dmma_peak, "flops_per_launch": fl, "ms_per_launch": t_potrf,
                "note": "latency-bound: serial chain of diagonal blocks"})
    gb = 2.0 * roi.m * gp.ld * 8
    res.append({"kernel": "gather_roi (Sigma[A, :] rows -> dense B, Sigma_AA)", "bound": "hbm",
                "achieved": gb / (t_gather * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": gb / (t_gather * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": gb,
                "ms_per_launch": t_gather})
    return res


def workload_config(args, w, world):
    return {
        "workload": (f"BASELINE configs[{4 if args.workload == 'c5' else 2}]: synthetic transductive ITL, "
                     f"N={w['N']} candidates, target set m={w['m']}, d={w['d']} Matern-5/2 "
                     f"(l={w['lengthscale']}), q=1, rho={w['rho']}, {'fp32' if getattr(args, 'dtype', 'f64') == 'f32' else 'fp64'}"),
        "name": args.workload, "N": w["N"], "m": w["m"], "d": w["d"], "q": 1,
        "conditioning": args.conditioning,
        "update_block": int(os.environ.get("TAL_UPDATE_BLOCK", "16")),
        "storage": ("full N x N covariance (reference layout)" if args.storage == "full" else
                    "lower-block: only the blocks on or below the diagonal are maintained (symmetric matrix)"),
        "l2": "inputs larger than L2 (covariance %.1f GB per GPU >> 126 MB): no flush needed"
              % (w["N"] * w["N"] * (4 if getattr(args, "dtype", "f64") == "f32" else 8) / world / 1e9),
        "parallelism": "1 GPU" if world == 1 else f"covariance row-sharded over {world} GPUs",
    }


def ncu_traffic(kernel_prefix, workload, n_gpus, storage="full"):
    """DRAM bytes per launch of a kernel from the committed ncu capture (profiles/*_kernel_traffic.json),
    or None when no capture of this workload / GPU count is on file."""
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
            if name.startswith(kernel_prefix):
                best = k["dram_bytes"]
    return best


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
    from talgp.engine import DeviceGP  # noqa: F401
    lib = _abi.load()
    _abi.require_device()
    domain, A, table = make_problem(w)
    N, m, rho, beta = w["N"], w["m"], w["rho"], [w["beta"]]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"

    clocks = ClockSampler(local)

    if world == 1:
        from lib.algorithms import index_roi_constructor
        from lib.algorithms.itl import ITL
        from lib.function import Function
        from lib.gp.kernels import stationary
        from lib.gp.means import ZeroMean
        from lib.gp.prior import prior_distr
        from lib.model.marginal import MarginalModel
        from lib.noise import HomoscedasticNoise

        def build_model():
            nz = HomoscedasticNoise(1, np.array([rho]))
            distrs = prior_distr(nz, domain, np.zeros((0, w["d"])), np.zeros((1, 0)),
                                 [stationary.Matern52(lengthscale=w["lengthscale"])], [ZeroMean()],
                                 storage=args.storage,
                                 dtype=torch.float32 if args.dtype == "f32" else torch.float64)
            model = MarginalModel(0, domain, distrs, beta=np.array(beta), noise_rate=nz)
            alg = ITL(model, index_roi_constructor(A), conditioning=args.conditioning)
            alg.sample_mask()  # static; computed once (lib/algorithms/__init__.py:80)
            return model, alg

        model, alg = build_model()
        gp = model._gp
        table_dev = torch.as_tensor(table, device=dev)
        roi = alg.roi
        inc = args.conditioning == "incremental"

        # ---- device-resident loop (value) -----------------------------------
        ev = []

        import ctypes as C
        sample = alg.sample_mask()
        use_ctx = inc and gp._ctx_usable()

        def dev_step(record):
            """One step with index, observation table and state on the device.  Incremental mode: the
            whole launch sequence is ONE host call (tal_ctx_step, csrc/stepctx.cu); the first step (and a
            step after the conditioning factor filled up) conditions on the target set from scratch."""
            st = gp._cond
            if use_ctx and st is not None and st["valid"]:
                gp.ctx_step(table_dev, N, beta, model.first_constraint, None, sample)
                return
            F = gp.score_itl_directed(roi.idx, roi.m, incremental=inc, roi_key=roi)
            res = gp.masked_argmax(F, None, model.first_constraint, sample)
            if record and not use_ctx:  # recompute mode: time the pass with events around the update
                idx_dev = res[0:1]
                gp.exchange_row(idx_dev)
                gp.step_vectors(idx_dev, 0, table_dev, N, True, beta)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                passed = gp.rank1_update()
                e1.record()
                if passed:
                    ev.append((e0, e1))
            else:
                gp.observe(res[0:1], table_dev, beta, y_gather=True, y_stride=N)

        clocks.start()  # sampled from the warm-up on: the timed region alone can be shorter than one sample
        for _ in range(args.warmup):
            dev_step(False)
        gp.flush()  # the timed region starts and ends with no update pending: exactly K steps of work
        torch.cuda.synchronize()
        l_start, u_start = gp.l.clone(), gp.u.clone()
        lib.tal_reset_launch_count()
        ctx = gp._ctx_get() if use_ctx else None
        if ctx is not None:
            lib.tal_ctx_timing_begin(C.byref(ctx))  # CUDA events around every pass over the matrix
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            dev_step(True)
        if ctx is not None:  # the updates still pending are applied inside the timed region
            ctx.pending = gp._pending
            _abi.check(lib.tal_ctx_flush(C.byref(ctx), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            gp._ctx_done(ctx)
        elif gp._pending:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gp.flush(); e1.record()
            ev.append((e0, e1))
        t1.record()
        torch.cuda.synchronize()
        total_ms = t0.elapsed_time(t1)
        launches = int(lib.tal_launch_count())
        if ctx is not None:
            ms_buf, n_buf = (C.c_double * 64)(), C.c_longlong(0)
            _abi.check(lib.tal_ctx_timing_read(C.byref(ctx), ms_buf, C.byref(n_buf),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            pass_ms = [float(ms_buf[i]) for i in range(int(n_buf.value))]
        else:
            pass_ms = [a.elapsed_time(b) for a, b in ev]
        ev = pass_ms
        rank1_ms = float(np.mean(pass_ms))
        rank1_total_ms = float(np.sum(pass_ms))
        # the scoring pass (conditioning update), timed after the loop on a scratch copy of cs
        cond_ms = None
        if inc and gp._cond is not None and gp._cond["valid"] and gp._cond["append"]:
            st = gp._cond
            p_ = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
            cs_tmp = st["cs"].clone()
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            idx_dev = gp.argmax_record()[0:1]

            def cond_once():
                _abi.check(lib.tal_cond_extract_col(p_(st["V"][0]), st["ncols"], st["rows"], gp.cand0, gp.n_cand,
                                                    p_(idx_dev), 0, p_(st["pcol"]), stream))
                _abi.check(lib.tal_cond_append_update(
                    p_(st["V"][0]), st["ncols"], st["cap"], st["m_pad"], st["rows"], st["ncols"], gp.n_cand, gp.cand0,
                    p_(st["pcol"]), p_(gp.kbuf[0]), p_(gp.wbuf[0]), p_(gp.rho[0]), p_(idx_dev), 0, p_(cs_tmp[0]),
                    p_(st["app"]), st["n_split"], gp.dt, stream))
            ts = []
            for _ in range(6):
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(); cond_once(); c1.record(); torch.cuda.synchronize()
                ts.append(c0.elapsed_time(c1))
            cond_ms = float(np.median(ts[1:]))
        status = gp.read_result().status
        assert status == 0, f"arg-max status {status}"
        # size-independent invariants of the state after the timed steps (full size, no oracle needed)
        invariants = {
            "var_equals_diag_cov_bitwise": bool(torch.equal(torch.diagonal(gp.cov[0, :, :N]).to(torch.float64), gp.var[0])),
            "state_finite": bool(torch.isfinite(gp.mean).all() and torch.isfinite(gp.var).all()),
            "variances_positive": bool((gp.var > 0).all()),
            "bounds_are_running_intersections": bool((gp.l >= l_start).all() and (gp.u <= u_start).all()),
        }
        assert all(invariants.values()), invariants

        # ---- end to end through the drop-in API with a host black box ---------
        class HostFunction(Function):
            q, first_constraint = 1, 1

            def observe(self, X):  # X: [1, d] device tensor -> D2H; returns host array [q, 1]
                x = X.cpu().numpy()
                i = model.last_argmax.idx
                noise = table[0, i] - np.sin(3.0 * domain[i].sum())  # the pre-drawn noise of point i
                return np.array([[np.sin(3.0 * x.sum()) + noise]])

        f = HostFunction()
        for _ in range(max(1, min(args.warmup, 3))):
            alg.step(f)
        torch.cuda.synchronize()
        e2e_steps = args.steps
        w0 = time.perf_counter()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(e2e_steps):
            out = alg.step(f)
        gp.flush()
        fmax = float(out["F_max"])  # D2H read of the step's result
        a1.record()
        torch.cuda.synchronize()
        e2e_ms = max(a0.elapsed_time(a1), (time.perf_counter() - w0) * 1e3)
        clk = clocks.stop()
        assert np.isfinite(fmax)
        h2d = 8 * model.q
        d2h = 32 + 8 * w["d"]  # arg-max record + selected point (the subset test is static here: cached)
        value = args.steps / (total_ms / 1e3)
        e2e_value = e2e_steps / (e2e_ms / 1e3)
        alg_bytes = gp.update_bytes()
        achieved = alg_bytes / (rank1_ms * 1e-3) / 1e9
        # with many updates per pass the kernel is limited by fp64 issue rather than HBM: two rounded
        # operations per update and element (kept unfused for bit-identity with the per-step update)
        second_bound = None
        if gp.update_block > 1:
            try:
                probe_out = torch.zeros((1,), dtype=torch.float64, device=dev)
                blocks, iters = 148 * 8, 2000
                pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                _abi.check(lib.tal_probe_fp64(1, blocks, iters, C.c_void_p(probe_out.data_ptr()), stream))
                pe0.record()
                _abi.check(lib.tal_probe_fp64(1, blocks, iters, C.c_void_p(probe_out.data_ptr()), stream))
                pe1.record(); torch.cuda.synchronize()
                dfma_per_s = blocks * iters * 16 * 256 / (pe0.elapsed_time(pe1) * 1e-3)  # fp64 instructions / s
                elems = alg_bytes / (2.0 * (4 if args.dtype == "f32" else 8))
                ops = 2.0 * gp.update_block * elems  # one rounded multiply + one rounded subtract per update
                second_bound = {"bound": "fp64 issue (measured live: tal_probe_fp64, independent DFMA chains)",
                                "peak_instr_per_s": dfma_per_s, "rounded_ops_per_full_pass": ops,
                                "min_ms_per_full_pass": ops / dfma_per_s * 1e3,
                                "note": "a pass applying all %d pending updates cannot be faster than this; "
                                        "the HBM time of the same pass is %.2f ms" % (
                                            gp.update_block, alg_bytes / (hbm_peak * 1e9) * 1e3)}
            except Exception as exc:  # pragma: no cover
                second_bound = {"error": str(exc)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, w, 1),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": launches, "invariants": invariants,
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
            "roofline": {"kernel": ("rank1_update (tal_rank1_update%s)" % ("_lower" if args.storage == "lower" else "")
                                    if gp.update_block == 1 else
                                    "rank-b update: %d deferred rank-1 updates per pass (tal_rankb_update, %s storage)"
                                    % (gp.update_block, args.storage)),
                         "bound": "hbm", "achieved": achieved,
                         "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "peak_source": peak_src,
                         "traffic": ncu_traffic("rank1_" if gp.update_block == 1 else "rankb_", args.workload, 1, args.storage),
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": rank1_ms,
                         "launches_in_timed_region": len(ev), "updates_per_launch": gp.update_block,
                         "share_of_step": rank1_total_ms / total_ms,
                         "frac_of_8TBs_datasheet": achieved / 8000.0,
                         "second_bound": second_bound,
                         "effective_GBs_vs_full_matrix_bytes": 2.0 * model.q * N * N * (4 if args.dtype == "f32" else 8)
                                                               / (rank1_ms * 1e-3) / 1e9},
        }
        # the scoring pass of this mode next to the dominant kernel (north_star: HBM GB/s of the
        # rank-1 update AND the scoring pass; tensor-pipe utilisation of the TRSM in recompute mode)
        line["kernels"] = []
        if cond_ms is not None:
            st = gp._cond
            cbytes = model.q * (st["rows"] + 2.0) * st["ncols"] * 8  # one read of the factor + two appended rows
            line["kernels"].append({
                "kernel": "cond_append_update (incremental target-set conditioning = the ITL scoring pass; "
                          "read-only GEMV over the factor)",
                "bound": "hbm", "achieved": cbytes / (cond_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": cbytes / (cond_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": cbytes,
                "ms_per_launch": cond_ms, "launches_per_step": 1,
                "traffic": ncu_traffic("cond_app_partial", args.workload, 1, args.storage),
                "share_of_step": cond_ms / (total_ms / args.steps),
                "timed": "after the loop, CUDA events, same buffers (scratch copy of cs)"})
        if not inc:
            line["kernels"] += recompute_kernels(gp, roi, hbm_peak)
        # beside the headline: the reference-style variants of the same workload, each a sub-run
        #   value_recompute     re-conditioning on the target set from scratch at every step (as the reference)
        #   value_full_storage  the full N x N covariance maintained (the reference's layout)
        if args.both:
            gp_update_block = gp.update_block
            del model, alg, gp, roi
            torch.cuda.empty_cache()
            variants = []
            if inc:
                variants.append(("value_recompute", ["--conditioning", "recompute", "--storage", args.storage]))
            else:
                variants.append(("value_incremental", ["--conditioning", "incremental", "--storage", args.storage]))
            if args.storage == "lower":
                variants.append(("value_full_storage", ["--conditioning", args.conditioning, "--storage", "full"]))
            if gp_update_block > 1:
                variants.append(("value_update_every_step", ["--conditioning", args.conditioning, "--storage",
                                                             args.storage, "--update-block", "1"]))
            if args.dtype == "f64":
                variants.append(("value_fp32_storage", ["--conditioning", args.conditioning, "--storage", args.storage,
                                                        "--dtype", "f32"]))
            for key, extra in variants:
                steps = max(3, args.steps // 4) if "recompute" in extra else args.steps
                sub = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", str(steps), "--warmup", "3",
                                      "--workload", args.workload, "--no-both", "--no-cpu"] + extra,
                                     capture_output=True, text=True)
                try:
                    o = json.loads(sub.stdout.strip().splitlines()[-1])
                    line[key] = {"value": o["value"], "ms_per_step": o["ms_per_step"], "e2e": o["e2e"]["value"],
                                 "roofline": {k: o["roofline"][k] for k in ("kernel", "achieved", "frac", "ms_per_launch",
                                                                            "algorithmic_bytes_per_launch")},
                                 "kernels": o.get("kernels", [])}
                except Exception as exc:  # pragma: no cover
                    line[key] = {"error": str(exc), "stderr": sub.stderr[-300:]}
        if not args.no_cpu:
            line["cpu_baseline"] = {k: v for k, v in cpu_oracle_sample(w, args.cpu_sample_n, 3, 1).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
        return

    # ---- N > 1: row-sharded -------------------------------------------------
    from talgp.dist import ShardedITLBench
    bench = ShardedITLBench(domain, A, table, w, world, rank, dev, args.conditioning, storage=args.storage)
    clocks.start()  # sampled from the warm-up on: the timed region alone can be shorter than one sample
    # a sharded step takes ~1 ms: warm up for a few hundred of them so that the clock sampler
    # (one nvidia-smi sample per 50 ms) sees the loaded GPU before and during the timed region
    warm_run = max(args.warmup, 200)
    for _ in range(warm_run):
        bench.step(False)
    bench.flush()  # the timed region starts and ends with no update pending: exactly K steps of work
    torch.cuda.synchronize()
    dist.barrier()
    lib.tal_reset_launch_count()
    bench.timing_begin()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(args.steps):
        bench.step(True)
    bench.flush(True)  # the updates still pending are applied inside the timed region
    t1.record()
    torch.cuda.synchronize()
    bench.timing_end()
    dist.barrier()
    total_ms = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
    dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    rank1_ms = torch.tensor([bench.rank1_ms()], device=dev, dtype=torch.float64)
    dist.all_reduce(rank1_ms, op=dist.ReduceOp.MAX)
    launches = int(lib.tal_launch_count())
    e2e_ms = bench.e2e(args.steps)
    clk = clocks.stop()
    # size-independent invariants at full size: every rank selected the same candidates, the replicated
    # vectors are identical on all ranks, var == diag(cov) bit for bit on the rows this rank holds
    g = bench.gp
    sel = g._combined[0:1].clone()
    sel_all = [torch.zeros_like(sel) for _ in range(world)]
    dist.all_gather(sel_all, sel)
    digest = torch.stack([g.mean.sum(), g.var.sum(), g.l.sum(), g.u.sum()])
    dig_all = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(dig_all, digest)
    rows = torch.arange(g.row0, g.row0 + g.n_loc, device=dev)
    diag_ok = torch.equal(g.cov[0, torch.arange(g.n_loc, device=dev), rows], g.var[0, rows])
    ok = torch.tensor([int(diag_ok and bool(torch.isfinite(g.var).all()) and bool((g.var > 0).all()))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    invariants = {
        "ranks_agree_on_selection": all(bool(torch.equal(s, sel_all[0])) for s in sel_all),
        "replicated_state_identical": all(bool(torch.equal(d, dig_all[0])) for d in dig_all),
        "var_equals_diag_cov_bitwise_and_positive_all_ranks": bool(int(ok) == 1),
    }
    assert all(invariants.values()), invariants
    exchange = "nccl"
    if getattr(bench.gp.comm, "peer", False):
        exchange = "nvlink-peer-memory"
        err = bench.gp.comm.error()
        assert err == 0, f"peer exchange timed out (TAL_PEER_ERR {err})"
    if os.environ.get("TAL_BREAKDOWN"):
        bd = bench.breakdown()
        if rank in (0, world - 1):
            sys.stderr.write(f"breakdown_us rank {rank} " + json.dumps(bd) + "\n")
    if rank == 0:
        total_ms, rank1_ms = float(total_ms), float(rank1_ms)
        alg_bytes = bench.gp.update_bytes()
        achieved = alg_bytes / (rank1_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": args.steps / (total_ms / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "warmup_steps_run": warm_run,
            "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(workload_config(args, w, world), exchange=exchange),
            "e2e": {"value": args.steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": 8,
                    "d2h_bytes_per_step": 32, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "invariants": invariants,
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
            "roofline": {"kernel": ("rank1_update (tal_rank1_update%s), per GPU, max over ranks" % (
                "_lower" if args.storage == "lower" else "") if bench.gp.update_block == 1 else
                "rank-b update: %d deferred rank-1 updates per pass (tal_rankb_update, %s storage), per GPU, "
                "max over ranks" % (bench.gp.update_block, args.storage)), "bound": "hbm",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "peak_source": peak_src, "traffic": ncu_traffic("rank1_", args.workload, world, args.storage),
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "ms_per_launch": rank1_ms, "updates_per_launch": bench.gp.update_block,
                         "launches_in_timed_region": len(bench.pass_ms),
                         "share_of_step": rank1_ms * len(bench.pass_ms) / total_ms,
                         "frac_of_8TBs_datasheet": achieved / 8000.0},
        }
        print(json.dumps(line), flush=True)
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
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="covariance storage / arithmetic type (the reference's only mode is f64)")
    ap.add_argument("--cpu-sample-n", type=int, default=8192)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-both", dest="both", action="store_false",
                    help="do not also time the other conditioning mode")
    args = ap.parse_args()
    if args.update_block is not None:
        os.environ["TAL_UPDATE_BLOCK"] = str(args.update_block)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            args.no_cpu = True
            args.both = False
        run_ours(args, w)


if __name__ == "__main__":
    main()
