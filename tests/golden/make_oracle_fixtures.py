"""Freezes oracle outputs on the small seeded problems of tests/problems.py:

    python tests/golden/make_oracle_fixtures.py

writes tests/golden/oracle_trajectories.npz: per problem and step the masked
objective F, the selected index, the size of the tie set, and the final means /
variances / bounds.  These are outputs of the ORACLE (oracle/ref_numpy.py), not
of the JAX reference (which cannot be imported in this image)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import problems  # noqa: E402


def run(p):
    model, alg, f = problems.build_oracle(p)
    Fs, sel, nt = [], [], []
    for t in range(p["T"]):
        F = alg.masked_F()
        ties, _ = model.objective_tie_set(F)
        idx = int(ties[0])
        Fs.append(F); sel.append(idx); nt.append(len(ties))
        X = model.domain[idx: idx + 1]
        model.step(X, f.observe(X))
    return dict(F=np.array(Fs), sel=np.array(sel), n_ties=np.array(nt), means=model.means,
                variances=model.variances, l=model.l, u=model.u)


if __name__ == "__main__":
    out = {}
    for name, fn in problems.ALL.items():
        res = run(fn())
        for k, v in res.items():
            out[f"{name}/{k}"] = v
        print(name, "selections", res["sel"].tolist())
    np.savez_compressed(os.path.join(HERE, "oracle_trajectories.npz"), **out)
