"""Debug: blocked Cholesky + TRSM on an ill-conditioned RBF Gram matrix with both diagonal-block kernels."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transductive-active-learning_b200"))

def child(tag):
    import torch
    from talgp import _abi
    lib = _abi.load()
    m = 1024
    x = np.linspace(-2, 8, 2000)[:m, None]
    S = np.exp(-0.5 * (x - x.T) ** 2) + 1e-10 * np.eye(m)
    B = np.exp(-0.5 * (x - np.linspace(-2, 8, 2000)[None, :1024]) ** 2)
    Sd = torch.as_tensor(S, device="cuda").contiguous(); Bd = torch.as_tensor(B, device="cuda").contiguous()
    work = torch.zeros((int(lib.tal_potrf_work_bytes(m)) // 8,), dtype=torch.float64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    _abi.check(lib.tal_potrf_lower(p(Sd), m, m, p(work), st))
    _abi.check(lib.tal_trsm_lower(p(Sd), m, m, p(work), p(Bd), 1024, 1024, st))
    torch.cuda.synchronize()
    np.save(os.path.join(ROOT, "gpurun_out", f"dbg_L_{tag}.npy"), np.tril(Sd.cpu().numpy()))
    np.save(os.path.join(ROOT, "gpurun_out", f"dbg_Linv_{tag}.npy"), work.cpu().numpy().reshape(m // 128, 128, 128))
    np.save(os.path.join(ROOT, "gpurun_out", f"dbg_V_{tag}.npy"), Bd.cpu().numpy())

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1]); sys.exit(0)
    for tag, env in (("old", "0"), ("new", "1")):
        subprocess.run([sys.executable, os.path.abspath(__file__), tag], env=dict(os.environ, TAL_POTRF_DIAG=env), check=True)
    import scipy.linalg as sl
    m = 1024
    x = np.linspace(-2, 8, 2000)[:m, None]
    S = np.exp(-0.5 * (x - x.T) ** 2) + 1e-10 * np.eye(m)
    Lo, Ln = np.load("gpurun_out/dbg_L_old.npy"), np.load("gpurun_out/dbg_L_new.npy")
    Io, In = np.load("gpurun_out/dbg_Linv_old.npy"), np.load("gpurun_out/dbg_Linv_new.npy")
    Vo, Vn = np.load("gpurun_out/dbg_V_old.npy"), np.load("gpurun_out/dbg_V_new.npy")
    for tag, L, I, V in (("old", Lo, Io, Vo), ("new", Ln, In, Vn)):
        print(tag, "nan L", int(np.isnan(L).sum()), "nan Linv", int(np.isnan(I).sum()), "nan V", int(np.isnan(V).sum()))
        if np.isnan(L).any():
            r, c = np.argwhere(np.isnan(L))[0]; print("   first NaN in L at", r, c, "block", r // 128, c // 128)
        for j in range(m // 128):
            d = slice(j * 128, (j + 1) * 128)
            Lb = L[d, d]
            print("   block", j, "min diag", np.nanmin(np.diag(Lb)), "|L X - I|", np.nanmax(np.abs(Lb @ I[j] - np.eye(128))),
                  "|X L - I|", np.nanmax(np.abs(I[j] @ Lb - np.eye(128))), "upper(Linv) max", np.abs(np.triu(I[j], 1)).max())
        print("   resid |LL^T - S|", np.nanmax(np.abs(L @ L.T - S)), " cs max", np.nanmax((V * V).sum(0)))
    print("max |L_new - L_old|", np.nanmax(np.abs(Ln - Lo)), "max |V_new - V_old|", np.nanmax(np.abs(Vn - Vo)))
