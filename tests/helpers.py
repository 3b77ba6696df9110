"""Shared test helpers: small problems built with the product-side host code (sem / workloads) and
evaluated with the oracle.  numpy on the oracle side, torch on the product side."""
import numpy as np
import torch

import neko_top_b200  # noqa: F401  (import shim)
from neko_top_b200 import sem, workloads


class Problem:
    """A small brick with geometry and fields, as flat float64 numpy arrays (Fortran order)."""

    def __init__(self, lx, ne=(2, 2, 2), deform=0.03, seed_shift=0, origin=(0.0, 0.0, 0.0), length=(1.0, 1.0, 1.0)):
        self.brick = workloads.BoxBrick(lx=lx, ne=tuple(ne), deform=deform, origin=origin, length=length)
        b = self.brick
        self.lx, self.nelv, self.n = lx, b.nelv, b.n
        self.space = sem.Space(lx)
        x, y, z = workloads.coords(b)
        self.xyz = (x, y, z)
        self.keys = workloads.node_keys(b)
        G, jac, B = sem.geometric_factors(x, y, z, self.space)
        fl = workloads.make_fields(b, x, y, z, self.keys + seed_shift)
        self.t = dict(G=G, jac=jac, B=B, ub=fl.ub, v=fl.v, rho=fl.rho)
        f = lambda a: a.reshape(-1).numpy().copy()
        self.G = [f(g) for g in G]
        self.jac, self.B = f(jac), f(B)
        self.ub = [f(a) for a in fl.ub]
        self.v = [f(a) for a in fl.v]
        self.rho = f(fl.rho)
        self.D, self.w = self.space.dx, self.space.wx

    def cuda(self, name):
        v = self.t[name]
        if isinstance(v, list):
            return [a.reshape(-1).cuda().contiguous() for a in v]
        return v.reshape(-1).cuda().contiguous()


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)
