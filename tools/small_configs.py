"""BASELINE configs[0] (1-D safe BO, ITL-PM, N = 2,000) and configs[1] (2-D safe BO, objective + 1
constraint, VTL-PM, N = 10,000) end to end through the drop-in API on one B200, next to the NumPy
oracle (the reference's algorithm as written) on the host cores at the SAME size.  These models are
small: a step is latency-bound (kernel launches + two host reads), not HBM-bound.
Writes JSON lines to gpurun_out/small_configs.jsonl."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transductive-active-learning_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import problems  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "small_configs.jsonl")


def run(name, p, steps_gpu, steps_cpu):
    model, alg, f = problems.build_cuda(p)
    for _ in range(3):
        alg.step(f)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps_gpu):
        alg.step(f)
    torch.cuda.synchronize()
    gpu = (time.perf_counter() - t0) / steps_gpu
    mo, ao, fo = problems.build_oracle(p)
    ao.step(fo)
    t0 = time.perf_counter()
    for _ in range(steps_cpu):
        ao.step(fo)
    cpu = (time.perf_counter() - t0) / steps_cpu
    line = dict(config=name, N=int(p["domain"].shape[0]), q=len(p["kernels"]), rule=p["rule"] + "-PM",
                gpu_steps_per_s=1.0 / gpu, gpu_ms_per_step=gpu * 1e3, cpu_oracle_steps_per_s=1.0 / cpu,
                cpu_oracle_ms_per_step=cpu * 1e3, cpu_cores=os.cpu_count(), steps_gpu=steps_gpu, steps_cpu=steps_cpu)
    print(json.dumps(line), flush=True)
    with open(OUT, "a") as fh:
        fh.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    torch.cuda.set_device(0)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    run("configs[0] 1-D safe BO ITL-PM", problems.c1(N=2000, T=0), 60, 5)
    run("configs[1] 2-D safe BO VTL-PM", problems.c2(T=0, step=0.06), 40, 3)
