"""Deferred-update pass (csrc/rankb.cu) at BASELINE configs[2] size on one B200: every kernel variant
(TAL_RANKB_VARIANT, read once per process, so one process per variant), both update arithmetics,
p = 4 / 8 / 16 updates per pass.  JSON lines to gpurun_out/microbench_rankb.jsonl.

    python tools/microbench_rankb.py            # spawns one child per variant
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "microbench_rankb.jsonl")


def child():
    import ctypes as C
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "transductive-active-learning_b200"))
    from talgp import _abi
    from talgp.engine import DeviceGP, _ptr, _stream
    peak = 6535.1
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    variant = int(os.environ.get("TAL_RANKB_VARIANT", "0"))
    N = int(os.environ.get("MB_N", "65536"))
    only_dt, only_ar, only_p = os.environ.get("MB_DTYPE"), os.environ.get("MB_ARITH"), os.environ.get("MB_P")
    for dtype in (torch.float64, torch.float32):
        if only_dt and only_dt != ("f64" if dtype == torch.float64 else "f32"):
            continue
        for arith in ("two_rounding", "fma"):
            if only_ar and only_ar != arith:
                continue
            gp = DeviceGP(N, 1, dtype=dtype, storage="lower", arith=arith)
            gp.set_update_block(32)
            gp._cov.normal_()
            gp._Kp.normal_(); gp._Wp.normal_().mul_(1e-3)
            bytes_ = gp.update_bytes()
            for p in ((int(only_p),) if only_p else (32, 16, 8, 4)):
                def run():
                    _abi.check(gp.lib.tal_rankb_update(_ptr(gp._cov), gp.cov_stride_q, gp.ld, 0, N, N, 1, _ptr(gp._Kp),
                                                       _ptr(gp._Wp), gp.ld, p, 1, gp.dt_upd, _stream()))
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                ts = []
                for _ in range(6):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); run(); b.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
                ms = float(np.median(ts))
                line = dict(kernel="rankb", variant=variant, dtype=str(dtype), arith=arith, p=p, N=N, ms=ms,
                            ms_min=float(np.min(ts)), GBs=bytes_ / ms / 1e6, frac=bytes_ / ms / 1e6 / peak)
                print(json.dumps(line), flush=True)
                with open(OUT, "a") as fh:
                    fh.write(json.dumps(line) + "\n")
            del gp
            torch.cuda.empty_cache()


if __name__ == "__main__":
    if os.environ.get("MB_CHILD"):
        child()
    else:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        for v in [int(x) for x in os.environ.get("MB_VARIANTS", "0,1,2,3,4,5,6,7,8").split(",")]:
            env = dict(os.environ, MB_CHILD="1", TAL_RANKB_VARIANT=str(v))
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
