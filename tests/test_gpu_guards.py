"""Memory-safety checks of the hand-written kernels WITHOUT compute-sanitizer (the tool is closed on the GPU
pool of this build: `tools/sanitize.sh` is refused, profiles/r02_sanitize_refused.log).  Every buffer a kernel
family touches is re-homed inside a larger allocation whose margins hold a NaN sentinel:

  * an out-of-bounds WRITE changes a margin  -> byte-compared after the run;
  * an out-of-bounds or uninitialised READ that reaches a result turns it into NaN, or makes it depend on the
    garbage: the run is repeated with a different sentinel and must give bit-identical, finite results.

Sizes are chosen off every tile boundary (N = 1,501: rows and columns not multiples of 32 / 128 / 256)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
MARGIN = 4096  # bytes on each side


class Guard:
    def __init__(self, pattern: int):
        self.pattern = pattern
        self.items = []

    def wrap(self, t: torch.Tensor) -> torch.Tensor:
        """A tensor of the same shape / dtype / contents inside a padded allocation (margins = sentinel)."""
        assert t.is_contiguous()
        nbytes = t.numel() * t.element_size()
        pad = (-nbytes) % 256
        buf = torch.empty((MARGIN + nbytes + pad + MARGIN,), dtype=torch.uint8, device=t.device)
        buf.view(torch.int64).fill_(self.pattern)
        inner = buf[MARGIN: MARGIN + nbytes].view(t.dtype).view(t.shape)
        inner.copy_(t)
        self.items.append((buf, nbytes))
        return inner

    def check(self):
        for buf, nbytes in self.items:
            lo = buf[:MARGIN].view(torch.int64)
            hi = buf[MARGIN + nbytes + ((-nbytes) % 256):].view(torch.int64)
            assert bool((lo == self.pattern).all()) and bool((hi == self.pattern).all()), "a kernel wrote outside its buffer"
            tail = buf[MARGIN + nbytes: MARGIN + nbytes + ((-nbytes) % 256)]
            if tail.numel() >= 8:
                assert bool((tail[: tail.numel() // 8 * 8].view(torch.int64) == self.pattern).all())


PATTERNS = (0x7FF8DEADBEEF0001, 0x7FF8C0FFEE123457)  # two different quiet-NaN payloads (NaN as fp64 and as 2 x fp32)


def _rehome(gp, g: Guard):
    for name in ("_cov", "mean", "var", "l", "u", "rho", "kbuf", "wbuf", "scal", "_Kp", "_Wp", "_result", "_argmax_scratch"):
        if getattr(gp, name, None) is not None:
            setattr(gp, name, g.wrap(getattr(gp, name)))
    gp._ctx = None


def _rehome_cond(gp, g: Guard):
    st = gp._cond
    for k in ("V", "cs", "pcol", "app", "fz"):
        if k in st:
            st[k] = g.wrap(st[k])
    gp._ctx = None


def _run_fused(cuda, pattern, dtype, storage, block):
    from talgp import _abi
    from talgp.engine import DeviceGP
    N, q, m = 1501, 2, 150
    rng = np.random.default_rng(7)
    dom = torch.as_tensor(rng.random((N, 3)), device=cuda)
    gp = DeviceGP(N, q, dtype=dtype, device=cuda, storage=storage, arith="fma")
    gp.set_update_block(block)
    g = Guard(pattern)
    _rehome(gp, g)
    for i in range(q):
        gp.build_gram(i, _abi.KERNEL_MATERN52, 1.0 + 0.5 * i, 0.3, dom)
    gp.rho.fill_(0.1)
    gp.mean[1:].fill_(4.0)
    gp.init_bounds([2.0] * q)
    roi = torch.as_tensor(np.sort(rng.choice(N, m, replace=False)), device=cuda)
    table = g.wrap(torch.as_tensor(rng.standard_normal((q, N)), device=cuda))
    F = gp.score_itl_directed(roi, m, incremental=True, roi_key=roi)  # gather -> potrf (DMMA diag) -> trsm -> colsumsq
    _rehome_cond(gp, g)
    gp.argmax_candidates(F, None, 1, None)
    sel = []
    for t in range(block + 5):  # one full pass + a partial one at the flush
        gp.ctx_step(table, N, [2.0] * q, 1, None, None)
        sel.append(int(gp.argmax_record()[0].item()))
    gp.flush()
    torch.cuda.synchronize()
    g.check()
    out = {n: getattr(gp, n).clone() for n in ("mean", "var", "l", "u")}
    out["cs"] = gp._cond["cs"][:, :N].clone()
    blk = int(gp.lib.tal_sym_block(gp.dt))
    r = torch.arange(N, device=cuda)
    kept = torch.arange(gp.ld, device=cuda)[None, :] < torch.clamp((r // blk + 1) * blk, max=gp.ld)[:, None]
    out["cov"] = gp._cov[:, kept].clone() if storage == "lower" else gp._cov[:, :, :N].clone()
    for k, v in out.items():
        assert bool(torch.isfinite(v).all()), k
    return sel, out


@pytest.mark.parametrize("dtype,storage,block", [(torch.float64, "lower", 16), (torch.float64, "full", 32),
                                                (torch.float32, "lower", 32), (torch.float64, "lower", 1)])
def test_fused_step_and_passes_stay_inside_their_buffers(cuda, dtype, storage, block):
    a = _run_fused(cuda, PATTERNS[0], dtype, storage, block)
    b = _run_fused(cuda, PATTERNS[1], dtype, storage, block)
    assert a[0] == b[0]
    for k in a[1]:
        assert torch.equal(a[1][k], b[1][k]), k


def _run_dense(cuda, pattern):
    """Cholesky / TRSM / gathers / row reduces / dense-block kernels on guarded buffers."""
    from talgp import _abi
    from talgp.engine import DeviceGP
    lib = _abi.load()
    N, m = 1501, 333
    rng = np.random.default_rng(3)
    dom = torch.as_tensor(rng.random((N, 2)), device=cuda)
    res = {}
    for storage in ("lower", "full"):
        gp = DeviceGP(N, 1, dtype=torch.float64, device=cuda, storage=storage)
        g = Guard(pattern)
        _rehome(gp, g)
        gp.build_gram(0, _abi.KERNEL_MATERN52, 1.0, 0.3, dom)
        gp.rho.fill_(0.1)
        gp.init_bounds([2.0])
        for idx in (5, 700, 1500):
            gp.observe(idx, rng.standard_normal(1), [2.0])
        roi = torch.as_tensor(np.sort(rng.choice(N, m, replace=False)), device=cuda)
        big = torch.as_tensor(np.sort(rng.choice(N, 1200, replace=False)), device=cuda)
        # scratch buffers of the conditioning path, pre-allocated inside guards
        m_pad, ncols = gp._cond_dims(1200)
        for name, shape in (("cv_S", (m_pad, m_pad)), ("cv_B", (m_pad, ncols)), ("blk_S", (m_pad, m_pad)),
                            ("ig_P", (m_pad, m_pad)), ("ig_Vt", (m_pad, m_pad)), ("score_part", (64, N)),
                            ("score_c_full", (N,)), ("score_c", (N,)), ("cv_cs", (ncols,)), ("cv_part", (64, ncols)),
                            ("cv_work", (int(lib.tal_potrf_work_bytes(m_pad)) // 8,))):
            gp._ws[name] = g.wrap(torch.zeros(shape, dtype=torch.float64, device=cuda)).view(-1)
        res[storage, "itl_small"] = gp.score_itl_directed(roi, m).clone()
        res[storage, "itl_big"] = gp.score_itl_directed(big, 1200).clone()
        for rule, kw in ((_abi.RULE_VTL, {}), (_abi.RULE_MMITL, {}), (_abi.RULE_TRUVAR, dict(p0=[1.5], p1=0.01))):
            res[storage, "rows_small", rule] = gp.score_rows(rule, roi, m, **kw).clone()
            res[storage, "rows_big", rule] = gp.score_rows(rule, big, 1200, **kw).clone()  # dense form in lower storage
        res[storage, "logdet"] = gp.block_logdet(0, roi, m, None, 1e-6).clone()
        res[storage, "ig"] = gp.information_gain(0, roi, m, big[:400].contiguous(), 400, None).clone()
        z = torch.as_tensor(rng.standard_normal((3, N)), device=cuda)
        res[storage, "sample"] = gp.sample(0, z).clone()
        torch.cuda.synchronize()
        g.check()
    for k, v in res.items():
        assert bool(torch.isfinite(v).all()), k
    return res


def test_dense_kernels_stay_inside_their_buffers(cuda):
    a, b = _run_dense(cuda, PATTERNS[0]), _run_dense(cuda, PATTERNS[1])
    for k in a:
        assert torch.equal(a[k], b[k]), k
