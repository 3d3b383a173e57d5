"""A NumPy stand-in for the third-party modules the reference imports
(`jax`, `jax.numpy`, `jax.random`, `jax.scipy.stats`, `chex`, `jaxtyping`,
`matplotlib`, `IPython`, `optax`), so that the reference's OWN Python source under
/root/reference/lib can be imported and executed in this image, where jax/jaxlib
0.4.30 are absent and not installable.

TEST INFRASTRUCTURE ONLY: used by tests/golden/make_reference_fixtures.py (run in
the build container, where /root/reference exists) to generate golden vectors
from the reference's code.  It contains no code of the reference; it maps the
jax API surface the reference uses onto NumPy/SciPy fp64:

  jnp.*                    -> numpy.* (arrays are an ndarray subclass with `.at[i].set(v)`)
  jnp.linalg.cholesky      -> LAPACK potrf (lower), jnp.linalg.solve -> LAPACK gesv (LU),
                              the same LAPACK routines XLA:CPU calls
  jax.vmap                 -> a Python loop over the mapped axis + stacking of the
                              outputs (tuples, lists, dicts, dataclasses)
  jax.jit                  -> identity
  jax.random               -> numpy Generator streams derived from the key (NOT
                              threefry: draws differ from real jax; the fixtures only
                              use randomness for tie-breaking, which is recorded)

What this pins: the reference's algorithm as its source states it (operation order,
index conventions, quirks).  What it does not pin: XLA's fusion / Eigen's summation
order (rounding-level differences).
"""
from __future__ import annotations

import dataclasses
import functools
import sys
import types

import numpy as np


# --------------------------------------------------------------------- arrays
class _AtIndexer:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtSetter(self.arr, idx)


class _AtSetter:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _apply(self, fn):
        out = np.array(self.arr, copy=True).view(JArr)
        fn(out)
        return out

    def set(self, v):
        def fn(out):
            out[self.idx] = v
        return self._apply(fn)

    def add(self, v):
        def fn(out):
            np.add.at(out, self.idx, v)
        return self._apply(fn)

    def multiply(self, v):
        def fn(out):
            np.multiply.at(out, self.idx, v)
        return self._apply(fn)

    def max(self, v):
        def fn(out):
            np.maximum.at(out, self.idx, v)
        return self._apply(fn)

    def min(self, v):
        def fn(out):
            np.minimum.at(out, self.idx, v)
        return self._apply(fn)

    def get(self):
        return self.arr[self.idx]


class JArr(np.ndarray):
    """ndarray with jax's functional-update property."""
    __array_priority__ = 100

    @property
    def at(self):
        return _AtIndexer(self)

    def block_until_ready(self):
        return self

    def __hash__(self):  # jax arrays are hashable by identity
        return id(self)


def _wrap_out(x):
    if isinstance(x, np.ndarray) and not isinstance(x, JArr):
        return x.view(JArr)
    if isinstance(x, tuple):
        return tuple(_wrap_out(v) for v in x)
    if isinstance(x, list):
        return [_wrap_out(v) for v in x]
    return x


def _wrap_fn(fn):
    @functools.wraps(fn)
    def inner(*a, **k):
        return _wrap_out(fn(*a, **k))
    return inner


class _NumpyProxy(types.ModuleType):
    """Module whose attributes are numpy's, functions returning JArr."""

    def __init__(self, name, target, overrides=None):
        super().__init__(name)
        self._target = target
        self._over = overrides or {}

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name in self._over:
            v = self._over[name]
        else:
            v = getattr(self._target, name)
            if callable(v) and not isinstance(v, type):
                v = _wrap_fn(v)
        setattr(self, name, v)
        return v


def _jnp_array(x, dtype=None, copy=True):
    if isinstance(x, (list, tuple)) and len(x) and any(dataclasses.is_dataclass(v) for v in x):
        raise TypeError("array of dataclasses")
    return np.array(x, dtype=dtype).view(JArr)


def _jnp_asarray(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(JArr)


def _argwhere(a, size=None, fill_value=None):
    return np.argwhere(a).view(JArr)


def _where(cond, x=None, y=None, size=None, fill_value=None):
    if x is None and y is None:
        return tuple(v.view(JArr) for v in np.where(cond))
    return _wrap_out(np.where(cond, x, y))


def _ix_(*args):
    return np.ix_(*[np.asarray(a) for a in args])


# ----------------------------------------------------------------------- vmap
def _is_leaf(x):
    return not (isinstance(x, (tuple, list, dict)) or dataclasses.is_dataclass(x))


def _tree_stack(items):
    x0 = items[0]
    if dataclasses.is_dataclass(x0) and not isinstance(x0, type):
        return type(x0)(**{f.name: _tree_stack([getattr(it, f.name) for it in items])
                           for f in dataclasses.fields(x0)})
    if isinstance(x0, tuple):
        return tuple(_tree_stack([it[i] for it in items]) for i in range(len(x0)))
    if isinstance(x0, list):
        return [_tree_stack([it[i] for it in items]) for i in range(len(x0))]
    if isinstance(x0, dict):
        return {k: _tree_stack([it[k] for it in items]) for k in x0}
    return np.stack([np.asarray(it) for it in items]).view(JArr)


def _tree_index(x, i, axis):
    if dataclasses.is_dataclass(x) and not isinstance(x, type):
        return type(x)(**{f.name: _tree_index(getattr(x, f.name), i, axis) for f in dataclasses.fields(x)})
    if isinstance(x, tuple):
        return tuple(_tree_index(v, i, axis) for v in x)
    if isinstance(x, list):
        return [_tree_index(v, i, axis) for v in x]
    if isinstance(x, dict):
        return {k: _tree_index(v, i, axis) for k, v in x.items()}
    return _wrap_out(np.take(np.asarray(x), i, axis=axis))


def _tree_size(x, axis):
    if dataclasses.is_dataclass(x) and not isinstance(x, type):
        return _tree_size(getattr(x, dataclasses.fields(x)[0].name), axis)
    if isinstance(x, (tuple, list)):
        return _tree_size(x[0], axis)
    if isinstance(x, dict):
        return _tree_size(next(iter(x.values())), axis)
    return np.asarray(x).shape[axis]


def vmap(fun, in_axes=0, out_axes=0):
    assert out_axes == 0, "shim: only out_axes=0"

    @functools.wraps(fun)
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _tree_size(a, ax)
                break
        assert n is not None and n > 0, "shim vmap: empty or unmapped batch"
        outs = []
        for i in range(n):
            call = [a if ax is None else _tree_index(a, i, ax) for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        return _tree_stack(outs)
    return mapped


def jit(fun=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def grad(fun, argnums=0):  # pragma: no cover - hyperopt only
    raise NotImplementedError("shim: jax.grad is not provided (hyper-parameter optimisation is off the path)")


value_and_grad = grad


# --------------------------------------------------------------------- random
def _rng(key):
    k = np.asarray(key, dtype=np.uint32).reshape(-1)
    return np.random.Generator(np.random.Philox(key=[int(k[0]) << 32 | int(k[-1]), 0x7A1B200]))


def _PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32).view(JArr)


def _split(key, num=2):
    g = _rng(key)
    return g.integers(0, 2 ** 32, size=(num, 2), dtype=np.uint64).astype(np.uint32).view(JArr)


def _shape(shape):
    return tuple(shape) if not isinstance(shape, int) else (shape,)


def _normal(key, shape=(), dtype=None):
    return _wrap_out(np.asarray(_rng(key).standard_normal(_shape(shape))))


def _uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return _wrap_out(np.asarray(minval + (np.asarray(maxval) - minval) * _rng(key).random(_shape(shape))))


def _randint(key, shape, minval, maxval, dtype=None):
    return _wrap_out(np.asarray(_rng(key).integers(minval, maxval, size=_shape(shape))))


def _choice(key, a, shape=(), replace=True, p=None, axis=0):
    a = np.asarray(a)
    if a.ndim == 0:
        a = np.arange(int(a))
    idx = _rng(key).choice(a.shape[axis], size=_shape(shape) if shape != () else None, replace=replace, p=p)
    return _wrap_out(np.take(a, idx, axis=axis))


def _permutation(key, x, axis=0, independent=False):
    if np.ndim(x) == 0:
        x = np.arange(int(x))
    return _wrap_out(_rng(key).permutation(np.asarray(x), axis=axis))


def _multivariate_normal(key, mean, cov, shape=None, dtype=None, method="cholesky"):
    mean, cov = np.asarray(mean), np.asarray(cov)
    shape = () if shape is None else _shape(shape)
    z = _rng(key).standard_normal(shape + mean.shape[-1:])
    if method == "svd":
        u, s, _ = np.linalg.svd(cov)
        factor = u * np.sqrt(s[..., None, :])
    elif method == "eigh":
        w, v = np.linalg.eigh(cov)
        factor = v * np.sqrt(np.clip(w, 0, None)[..., None, :])
    else:
        factor = np.linalg.cholesky(cov)
    return _wrap_out(mean + np.einsum("...ij,...j->...i", factor, z))


# ------------------------------------------------------------------- dataclass
def chex_dataclass(cls=None, **kw):
    def wrap(c):
        c = dataclasses.dataclass(c)
        c.replace = lambda self, **ch: dataclasses.replace(self, **ch)
        return c
    return wrap if cls is None else wrap(cls)


# ------------------------------------------------------------------ jaxtyping
class _Sub:
    def __class_getitem__(cls, item):
        return cls

    def __getitem__(self, item):
        return self


# ---------------------------------------------------------------- permissive stubs
class _Anything(types.ModuleType):
    """Module that yields itself-like objects for any attribute (matplotlib, IPython ...)."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name not in ("__call__",):
            raise AttributeError(name)
        child = _Anything(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return self


class _StubFinder:
    ROOTS = ("matplotlib", "IPython", "optax", "wandb", "trajax", "ipywidgets", "mpl_toolkits", "seaborn")

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Anything(spec.name)

    def exec_module(self, module):
        pass


def install():
    """Insert the stand-in modules into sys.modules (idempotent)."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "__refshim__", False):
        return
    if "jax" in sys.modules:
        raise RuntimeError("a real jax is already imported")
    import scipy.stats

    over = dict(array=_jnp_array, asarray=_jnp_asarray, argwhere=_argwhere, where=_where, ix_=_ix_,
                ndarray=np.ndarray, float64=np.float64, float32=np.float32, int32=np.int32, int64=np.int64,
                bool_=np.bool_, inf=np.inf, nan=np.nan, pi=np.pi, e=np.e, newaxis=None)
    jnp = _NumpyProxy("jax.numpy", np, over)
    jnp.linalg = _NumpyProxy("jax.numpy.linalg", np.linalg)

    jr = types.ModuleType("jax.random")
    jr.PRNGKey = _PRNGKey
    jr.key = _PRNGKey
    jr.split = _split
    jr.normal = _normal
    jr.uniform = _uniform
    jr.randint = _randint
    jr.choice = _choice
    jr.permutation = _permutation
    jr.shuffle = _permutation
    jr.multivariate_normal = _multivariate_normal

    jax = types.ModuleType("jax")
    jax.__refshim__ = True
    jax.__path__ = []
    jax.numpy, jax.random = jnp, jr
    jax.vmap, jax.jit, jax.grad, jax.value_and_grad = vmap, jit, grad, value_and_grad
    jax.Array = JArr

    class _Config:
        def update(self, *a, **k):
            pass
    jax.config = _Config()
    src = types.ModuleType("jax._src")
    src.__path__ = []
    cfg = types.ModuleType("jax._src.config")
    cfg.config = jax.config
    src.config = cfg
    jax._src = src

    jsp = types.ModuleType("jax.scipy")
    jsp.__path__ = []
    jstats = types.ModuleType("jax.scipy.stats")

    class _Norm:
        pdf = staticmethod(_wrap_fn(scipy.stats.norm.pdf))
        cdf = staticmethod(_wrap_fn(scipy.stats.norm.cdf))
        logpdf = staticmethod(_wrap_fn(scipy.stats.norm.logpdf))
        ppf = staticmethod(_wrap_fn(scipy.stats.norm.ppf))
    jstats.norm = _Norm

    class _MVN:
        logpdf = staticmethod(_wrap_fn(scipy.stats.multivariate_normal.logpdf))
        pdf = staticmethod(_wrap_fn(scipy.stats.multivariate_normal.pdf))
    jstats.multivariate_normal = _MVN
    jsp.stats = jstats
    jax.scipy = jsp

    chex = types.ModuleType("chex")
    chex.dataclass = chex_dataclass
    chex.Array = JArr

    jt = types.ModuleType("jaxtyping")
    for n in ("Array", "Bool", "Float", "Int", "PRNGKeyArray", "Key", "Shaped", "Num", "PyTree", "UInt"):
        setattr(jt, n, _Sub)

    sys.modules.update({
        "jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": jnp.linalg, "jax.random": jr,
        "jax._src": src, "jax._src.config": cfg, "jax.scipy": jsp, "jax.scipy.stats": jstats,
        "chex": chex,
    })
    try:  # the real jaxtyping (annotations only) is used when the image has it
        import jaxtyping  # noqa: F401
        from jaxtyping import Array, Float  # noqa: F401
    except Exception:
        sys.modules["jaxtyping"] = jt
    sys.meta_path.append(_StubFinder())
