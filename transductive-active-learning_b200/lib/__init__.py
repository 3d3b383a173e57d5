"""Drop-in mirror of the reference's `lib` package for the discrete-domain GP
hot path (lib.model, lib.algorithms, lib.gp, lib.noise, lib.utils), backed by
hand-written sm_100a CUDA kernels through the C ABI of include/tal_b200.h.

Arrays are device-resident `torch.Tensor`s (DLPack-exportable) where the
reference uses `jax.Array`s.  The loop of the reference's docstring works
unchanged (lib/__init__.py:11-19 of the reference):

    model = MarginalModel(...); alg = ITL(model, roi_constructor)
    for t in range(T):
        alg.step(f)
"""
