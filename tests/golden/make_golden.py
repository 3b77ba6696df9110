"""Generates tests/golden/adjrhs_lx4.json from the NUMPY twin of the oracle (oracle/np_oracle.py), an
independent einsum-based restatement of the reference call sites.  These are regression fixtures for the
two oracles and the CUDA path -- they are NOT output of the reference: the reference (Fortran on top of
un-vendored Neko) cannot be built or imported in this image and ships no golden vectors for this path
("parity unpinned").

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import Problem  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402


def main():
    lx, ne, deform, lxd = 4, (2, 2, 1), 0.03, 6
    P = Problem(lx, ne=ne, deform=deform)
    v4 = lambda a: [npo.v4(x, lx, P.nelv) for x in a] if isinstance(a, list) else npo.v4(a, lx, P.nelv)
    chi = npo.ramp(P.rho)
    f = npo.adjoint_rhs(v4(P.v), v4(P.ub), P.D, P.w, v4(P.G), v4(P.B), v4(chi))
    f = [a.reshape(-1) for a in f]
    s = npo.sensitivity(P.ub, P.v).reshape(-1)
    z = [np.zeros((P.nelv, lx, lx, lx)) for _ in range(3)]
    fd = [a.reshape(-1) for a in npo.adjoint_advection_dealias(z, v4(P.v), v4(P.ub), lx, lxd, v4(P.G))]
    idx = list(range(0, P.n, 7))
    out = dict(lx=lx, ne=list(ne), deform=deform, lxd=lxd, idx=idx,
               f=[a[idx].tolist() for a in f], sens=s[idx].tolist(), chi=chi[idx].tolist(),
               f_dealias=[a[idx].tolist() for a in fd], sum_abs_f=float(sum(np.abs(a).sum() for a in f)))
    with open(os.path.join(HERE, "adjrhs_lx4.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote", len(idx), "sample points")


if __name__ == "__main__":
    main()
