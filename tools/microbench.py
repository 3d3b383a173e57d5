"""Per-kernel timings on one B200 (CUDA events, L2-exceeding inputs).  Writes
JSON lines to gpurun_out/microbench.jsonl.  Usage: python tools/microbench.py [what ...]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transductive-active-learning_b200"))
sys.path.insert(0, ROOT)
from talgp import _abi  # noqa: E402
from talgp.engine import DeviceGP, _ptr, _stream  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "microbench.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
PEAK = 6535.1
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    with open(OUT, "a") as fh:
        fh.write(line + "\n")


def time_ms(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def bench_rank1(N, dtype, q=1):
    gp = DeviceGP(N, q, dtype=dtype)
    gp.cov.normal_()
    gp.kbuf.normal_()
    gp.wbuf.normal_().mul_(1e-3)
    s = 8 if dtype == torch.float64 else 4
    bytes_ = 2 * q * N * gp.ld * s
    for variant in (1, 2, 3):
        med, mn = time_ms(lambda: gp.rank1_update(variant), warm=3, reps=8)
        emit(kernel="rank1_update", variant=variant, N=N, q=q, dtype=str(dtype), ms=med, ms_min=mn,
             GBs=bytes_ / med / 1e6, frac_of_measured=bytes_ / med / 1e6 / PEAK)
    # reference point: torch copy of the same bytes (what MEASURED_PEAKS measured)
    other = torch.empty_like(gp.kbuf)
    src = gp.cov[0]
    dst = torch.empty_like(src)
    med, mn = time_ms(lambda: dst.copy_(src), warm=3, reps=8)
    emit(kernel="torch_copy", N=N, dtype=str(dtype), ms=med, GBs=2 * src.numel() * s / med / 1e6)
    del dst
    return gp


def bench_scoring(gp, m):
    N = gp.N
    s = 8 if gp.dtype == torch.float64 else 4
    A = torch.as_tensor(np.sort(np.random.default_rng(0).choice(N, m, replace=False)), device=gp.device)
    gp.var.fill_(1.0)
    gp.rho.fill_(0.1)
    for rule, name in ((_abi.RULE_VTL, "vtl"), (_abi.RULE_MMITL, "mmitl")):
        med, mn = time_ms(lambda: gp.score_rows(rule, A, m), warm=2, reps=8)
        emit(kernel="score_rows/" + name, N=N, m=m, q=gp.q, dtype=str(gp.dtype), ms=med,
             GBs=gp.q * m * gp.ld * s / med / 1e6, frac_of_measured=gp.q * m * gp.ld * s / med / 1e6 / PEAK)
    med, mn = time_ms(lambda: gp.score_rows(_abi.RULE_VTL, A, m, sharded_cols=True), warm=2, reps=5)
    emit(kernel="score_cols/vtl", N=N, m=m, ms=med, GBs_alg=gp.q * m * N * s / med / 1e6)
    F = gp.score_rows(_abi.RULE_VTL, A, m)
    med, mn = time_ms(lambda: gp.masked_argmax(F, None, 1, None), warm=2, reps=10)
    emit(kernel="masked_argmax", N=N, ms=med)
    y = torch.zeros((gp.q, N), dtype=torch.float64, device=gp.device)
    idx = torch.tensor([N // 3], device=gp.device)

    def vec():
        gp.extract_row(idx)
        gp.step_vectors(idx, 0, y, N, True, [1.0] * gp.q)
    med, mn = time_ms(vec, warm=2, reps=10)
    emit(kernel="extract_row+step_vectors", N=N, ms=med)


def bench_condvar(N, m):
    lib = _abi.load()
    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    # SPD matrix with a Matérn-like spectrum: S = G G^T / m + I * 0.1
    G = torch.randn((m, 256), dtype=torch.float64, device=dev)
    S0 = G @ G.T / 256 + 0.1 * torch.eye(m, dtype=torch.float64, device=dev)
    work = torch.empty((int(lib.tal_potrf_work_bytes(m)) // 8,), dtype=torch.float64, device=dev)
    S = S0.clone()

    def potrf():
        S.copy_(S0)
        _abi.check(lib.tal_potrf_lower(_ptr(S), m, m, _ptr(work), _stream()))
    med_c, _ = time_ms(lambda: S.copy_(S0), warm=1, reps=4)
    med, mn = time_ms(potrf, warm=1, reps=4)
    emit(kernel="potrf_lower", m=m, ms=med - med_c, TFLOPs=(m ** 3 / 3) / (med - med_c) / 1e9)
    Lref = torch.linalg.cholesky(S0)
    emit(kernel="potrf_lower/check", max_rel_err=float((torch.tril(S) - Lref).abs().max() / Lref.abs().max()))
    B0 = torch.randn((m, N), dtype=torch.float64, device=dev)
    B = B0.clone()

    def trsm():
        _abi.check(lib.tal_trsm_lower(_ptr(S), m, m, _ptr(work), _ptr(B), N, N, _stream()))
    B.copy_(B0)
    trsm()
    torch.cuda.synchronize()
    ref = torch.linalg.solve_triangular(Lref, B0[:, :4096], upper=False)
    emit(kernel="trsm_lower/check", max_rel_err=float((B[:, :4096] - ref).abs().max() / ref.abs().max()))
    med, mn = time_ms(trsm, warm=1, reps=4)
    emit(kernel="trsm_lower", m=m, N=N, ms=med, TFLOPs=(m * m * N) / med / 1e9)
    t0 = time_ms(lambda: torch.linalg.solve_triangular(Lref, B0, upper=False), warm=1, reps=3)[0]
    emit(kernel="cublas_trsm(torch)", m=m, N=N, ms=t0, TFLOPs=(m * m * N) / t0 / 1e9)
    part = torch.empty((64, N), dtype=torch.float64, device=dev)
    out = torch.empty((N,), dtype=torch.float64, device=dev)
    med, mn = time_ms(lambda: _abi.check(lib.tal_colsumsq_dense(_ptr(B), N, m, N, _ptr(part), 16, _ptr(out),
                                                                _stream())), warm=2, reps=8)
    emit(kernel="colsumsq_dense", m=m, N=N, ms=med, GBs=m * N * 8 / med / 1e6)


def bench_condinc(N, m):
    lib = _abi.load()
    dev = torch.device("cuda")
    V = torch.randn((m, N), dtype=torch.float64, device=dev) * 0.01
    pcol = V[:, 777].clone()
    kbuf = torch.ones((N,), dtype=torch.float64, device=dev)
    wbuf = torch.randn((N,), dtype=torch.float64, device=dev) * 0.1
    rho = torch.full((N,), 0.1, dtype=torch.float64, device=dev)
    coef = torch.empty((int(lib.tal_cond_coef_bytes(m)) // 8,), dtype=torch.float64, device=dev)
    cs = torch.empty((N,), dtype=torch.float64, device=dev)
    chunk = torch.empty((int(lib.tal_cond_chunk_scratch_bytes(N)) // 8,), dtype=torch.float64, device=dev)
    for ncols in (N, N // 2, N // 8):
        for chunks in (1, 2, 4, 8, 16, 32):
            med, mn = time_ms(lambda: _abi.check(lib.tal_cond_update(
                _ptr(V), N, m, ncols, ncols, 0, _ptr(pcol), _ptr(kbuf), _ptr(wbuf), _ptr(rho), None, 777,
                _ptr(coef), _ptr(cs), chunks, _ptr(chunk), _abi.F64, _stream())), warm=2, reps=6)
            emit(kernel="cond_update(+coeffs)", m=m, ncols=ncols, chunks=chunks, ms=med, ms_min=mn,
                 GBs_2pass=2 * m * ncols * 8 / med / 1e6)


def bench_probe():
    lib = _abi.load()
    out = torch.zeros((1,), dtype=torch.float64, device="cuda")
    for kind, name, per in ((0, "DMMA.8x8x4", 8 * 512), (1, "DFMA", 256 * 2)):
        for blocks in (148 * 4, 148 * 8):
            iters = 4000
            med, mn = time_ms(lambda: _abi.check(lib.tal_probe_fp64(kind, blocks, iters, _ptr(out), _stream())),
                              warm=2, reps=5)
            emit(kernel="probe/" + name, blocks=blocks, ms=med, TFLOPs=blocks * iters * 16 * per / mn / 1e9)


def bench_gram(N):
    gp = DeviceGP(N, 1, dtype=torch.float64)
    g = np.arange(16) / 16.0
    dom = torch.as_tensor(np.stack(np.meshgrid(g, g, g, g, indexing="ij"), -1).reshape(-1, 4)[:N], device="cuda")
    for kind, name in ((_abi.KERNEL_MATERN52, "matern52"), (_abi.KERNEL_GAUSSIAN, "gaussian")):
        med, mn = time_ms(lambda: _abi.check(gp.lib.tal_gram_stationary(
            kind, _ptr(dom), N, _ptr(dom), N, 4, 1.0, 0.25, _ptr(gp.cov[0]), gp.ld, 0, N, gp.dt, _stream())),
            warm=1, reps=4)
        emit(kernel="gram/" + name, N=N, d=4, ms=med, GBs_write=N * gp.ld * 8 / med / 1e6)
    return gp


if __name__ == "__main__":
    what = sys.argv[1:] or ["probe", "rank1_f64", "rank1_f32", "score", "condvar", "gram"]
    torch.cuda.set_device(0)
    emit(event="start", gpu=torch.cuda.get_device_name(0), what=what, t=time.time())
    N = int(os.environ.get("TAL_N", 65536))
    m = int(os.environ.get("TAL_M", 4096))
    if "probe" in what:
        bench_probe()
    if "rank1_f64" in what:
        gp = bench_rank1(N, torch.float64)
        if "score" in what:
            bench_scoring(gp, m)
        del gp
        torch.cuda.empty_cache()
    if "rank1_f32" in what:
        gp = bench_rank1(N, torch.float32)
        del gp
        torch.cuda.empty_cache()
    if "condvar" in what:
        bench_condvar(N, m)
        torch.cuda.empty_cache()
    if "condinc" in what:
        bench_condinc(N, m)
        torch.cuda.empty_cache()
    if "gram" in what:
        bench_gram(N)
    emit(event="done")
