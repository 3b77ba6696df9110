"""Synthetic meshes and fields for BASELINE.json's configs.

Everything is generated from the GLOBAL node lattice, so fields are C0 across elements and across
GPU partitions, and any rank can generate exactly its own brick.  torch is the array library (CUDA
in bench.py, CPU in the tests); nothing here is inside a timed region.

Element order inside a brick: e = ex + nex*(ey + ney*ez) (x fastest); points (k, j, i) with i fastest,
i.e. tensors of shape (nelv, lx, lx, lx) are the Fortran arrays x(lx,lx,lx,nelv).
"""
import os
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np
import torch

from . import sem


@dataclass
class BoxBrick:
    """One rank's brick of a structured box mesh of hexahedra."""
    lx: int
    ne: Tuple[int, int, int]                       # local elements (nex, ney, nez)
    ne_global: Tuple[int, int, int] = None         # global elements (defaults to ne)
    offset: Tuple[int, int, int] = (0, 0, 0)       # brick origin in the global element lattice
    origin: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    length: Tuple[float, float, float] = (1.0, 1.0, 1.0)   # global box size
    deform: float = 0.0                            # amplitude of a smooth non-affine warp
    name: str = "box"

    def __post_init__(self):
        if self.ne_global is None:
            self.ne_global = tuple(self.ne)

    @property
    def nelv(self):
        return self.ne[0] * self.ne[1] * self.ne[2]

    @property
    def n(self):
        return self.nelv * self.lx ** 3


# ---- BASELINE.json configs ---------------------------------------------------------------------
def config_duct(lx):
    """configs[0] (lx=6) / configs[2] (lx=8): duct 24x8x8 on [0,10]x[-.5,.5]^2
    (/root/reference/data/thermal_mix_channel/square_pipe.jou:6-14)."""
    return BoxBrick(lx=lx, ne=(24, 8, 8), origin=(0.0, -0.5, -0.5), length=(10.0, 1.0, 1.0), name="duct24x8x8")


def config_box(ne, lx=8, deform=0.0):
    """configs[1]: uniform box [0,1]^3 of ne^3 hexes."""
    return BoxBrick(lx=lx, ne=(ne, ne, ne), deform=deform, name=f"box{ne}^3")


def tile_order(brick, tile=(16, 16)):
    """Processing order for b200_adjrhs_set_element_order: columns of tile[0] x tile[1] elements in (x, y),
    walked along z, x fastest inside a layer.  The z-neighbour of an element is then tile[0]*tile[1]
    positions away and the x/y neighbours 1 / tile[0], so a node class completes while the earlier
    members' right-hand side is still in L2; only faces between columns (1/tile of the x and y faces) meet
    a neighbour that was stored long ago.  Returns int32 element ids (e = ex + nex*(ey + ney*ez))."""
    nex, ney, nez = brick.ne
    tx, ty = min(tile[0], nex), min(tile[1], ney)
    ex, ey, ez = np.meshgrid(np.arange(nex), np.arange(ney), np.arange(nez), indexing="ij")
    ex, ey, ez = ex.ravel(), ey.ravel(), ez.ravel()
    key = ((((ey // ty) * ((nex + tx - 1) // tx) + ex // tx) * nez + ez) * ty + ey % ty) * tx + ex % tx
    order = np.argsort(key, kind="stable")
    return (ex[order] + nex * (ey[order] + ney * ez[order])).astype(np.int32)


def rank_grid(nranks):
    """configs[3]: GPU grids 1, 2x1x1, 2x2x1, 2x2x2."""
    grids = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    if nranks not in grids:
        raise ValueError(f"rank grid for {nranks} ranks not defined (1, 2, 4, 8)")
    return grids[nranks]


def config_weak(rank, nranks, ne_per_gpu=64, lx=8):
    """configs[3]: ne_per_gpu^3 elements per GPU, global box (ne*px) x (ne*py) x (ne*pz)."""
    px, py, pz = rank_grid(nranks)
    rx, ry, rz = rank % px, (rank // px) % py, rank // (px * py)
    ne = (ne_per_gpu,) * 3
    return BoxBrick(lx=lx, ne=ne, ne_global=(ne_per_gpu * px, ne_per_gpu * py, ne_per_gpu * pz),
                    offset=(rx * ne_per_gpu, ry * ne_per_gpu, rz * ne_per_gpu),
                    length=(float(px), float(py), float(pz)), name=f"box{ne_per_gpu}^3/gpu")


def config_sweep(lx, target_dof=1.0e8):
    """configs[4]: E = round(1e8/lx^3) arranged as a near-cubic box."""
    ne = max(1, round((target_dof / lx ** 3) ** (1.0 / 3.0)))
    return BoxBrick(lx=lx, ne=(ne, ne, ne), name=f"sweep lx={lx} box{ne}^3")


# ---- lattice indices, coordinates, keys ------------------------------------------------------------
def _lattice(brick, device):
    """Global lattice index (gi, gj, gk) of every local point, each (nelv, lx, lx, lx) int64."""
    lx = brick.lx
    nex, ney, nez = brick.ne
    e = torch.arange(brick.nelv, device=device, dtype=torch.int64)
    ex = (e % nex) + brick.offset[0]
    ey = ((e // nex) % ney) + brick.offset[1]
    ez = (e // (nex * ney)) + brick.offset[2]
    p = torch.arange(lx, device=device, dtype=torch.int64)
    gi = (ex * (lx - 1)).view(-1, 1, 1, 1) + p.view(1, 1, 1, lx)
    gj = (ey * (lx - 1)).view(-1, 1, 1, 1) + p.view(1, 1, lx, 1)
    gk = (ez * (lx - 1)).view(-1, 1, 1, 1) + p.view(1, lx, 1, 1)
    shape = (brick.nelv, lx, lx, lx)
    return gi.expand(shape), gj.expand(shape), gk.expand(shape)


def node_keys(brick, device="cpu"):
    """int64 global node id of every local dof: two dofs are the same physical GLL node iff their
    keys are equal."""
    lx = brick.lx
    gi, gj, gk = _lattice(brick, device)
    NXg = brick.ne_global[0] * (lx - 1) + 1
    NYg = brick.ne_global[1] * (lx - 1) + 1
    return (gi + NXg * (gj + NYg * gk)).contiguous()


def coords(brick, device="cpu"):
    """Nodal coordinates x, y, z (nelv, lx, lx, lx) float64; shared nodes get bit-identical values."""
    lx = brick.lx
    zg, _ = sem.zwgll(lx)
    t = torch.as_tensor((zg + 1.0) * 0.5, dtype=torch.float64, device=device)   # 0 .. 1 exactly at ends
    nex, ney, nez = brick.ne
    e = torch.arange(brick.nelv, device=device, dtype=torch.int64)
    ex = ((e % nex) + brick.offset[0]).to(torch.float64).view(-1, 1, 1, 1)
    ey = (((e // nex) % ney) + brick.offset[1]).to(torch.float64).view(-1, 1, 1, 1)
    ez = ((e // (nex * ney)) + brick.offset[2]).to(torch.float64).view(-1, 1, 1, 1)
    shape = (brick.nelv, lx, lx, lx)
    # normalised global coordinates in [0,1]
    X = ((ex + t.view(1, 1, 1, lx)) / brick.ne_global[0]).expand(shape)
    Y = ((ey + t.view(1, 1, lx, 1)) / brick.ne_global[1]).expand(shape)
    Z = ((ez + t.view(1, lx, 1, 1)) / brick.ne_global[2]).expand(shape)
    if brick.deform != 0.0:
        a = brick.deform
        two_pi = 2.0 * np.pi
        s = torch.sin(two_pi * X) * torch.sin(two_pi * Y) * torch.sin(two_pi * Z)
        X, Y, Z = X + a * s, Y + 0.7 * a * s, Z - 0.5 * a * s
    x = brick.origin[0] + brick.length[0] * X
    y = brick.origin[1] + brick.length[1] * Y
    z = brick.origin[2] + brick.length[2] * Z
    return x.contiguous(), y.contiguous(), z.contiguous()


# ---- unstructured hexahedral meshes (Neko .nmsh element records) ------------------------------------------
_LEX2NEK = (0, 1, 3, 2, 4, 5, 7, 6)     # Nek-style cyclic vertex order -> lexicographic (r fastest, then s, t)


class HexMesh:
    """Straight-sided hexahedra given by their eight vertices (ids and coordinates) in the order of a Neko
    `.nmsh` element record; element and vertex numbering are the file's own.  Produces what the
    path needs: GLL node coordinates (trilinear map) and a global node key per dof built from the TOPOLOGY
    (vertex ids), so that coincident nodes of elements with different local orientations get the same key:
      vertex node   (0, v)                      edge node  (1, v_lo, v_hi, offset from v_lo)
      face node     (2, sorted 4 vertices -> canonical origin = smallest id, first axis towards the smaller
                     of its two neighbours, (p, q) in that frame)
      interior node (3, element, i, j, k)
    GLL points are symmetric, so an offset m from one end is lx-1-m from the other."""

    def __init__(self, vertex_id, vertex_xyz, lx, name="hexmesh"):
        self.vid = np.asarray(vertex_id, dtype=np.int64)[:, list(_LEX2NEK)]        # (nelv, 8) lexicographic
        self.vxyz = np.asarray(vertex_xyz, dtype=np.float64)[:, list(_LEX2NEK), :]
        self.lx, self.nelv, self.name = int(lx), self.vid.shape[0], name
        self.n = self.nelv * self.lx ** 3
        lo, hi = self.vxyz.reshape(-1, 3).min(0), self.vxyz.reshape(-1, 3).max(0)
        self.origin, self.length = tuple(lo), tuple(hi - lo)

    def coords(self, device="cpu"):
        lx = self.lx
        zg, _ = sem.zwgll(lx)
        t = (zg + 1.0) * 0.5
        w = np.stack([1.0 - t, t])                                   # w[a, i]: weight of corner a at node i
        # N[lv, k, j, i] with lv = a + 2b + 4c
        N = np.einsum("ck,bj,ai->cbakji", w, w, w).reshape(8, lx, lx, lx)
        xyz = np.einsum("vkji,evd->dekji", N, self.vxyz)
        return tuple(torch.as_tensor(np.ascontiguousarray(xyz[d]), dtype=torch.float64, device=device) for d in range(3))

    def node_keys(self, device="cpu"):
        lx, L = self.lx, self.lx - 1
        rows = np.zeros((self.nelv, lx, lx, lx, 7), dtype=np.int64)
        k, j, i = np.meshgrid(np.arange(lx), np.arange(lx), np.arange(lx), indexing="ij")
        idx = np.stack([i, j, k], -1)                                # local (r, s, t) index of every node
        onb = (idx == 0) | (idx == L)
        nb = onb.sum(-1)
        corner = (idx == L).astype(np.int64)                         # corner coordinate where on the boundary
        for e in range(self.nelv):
            v = self.vid[e]
            r = rows[e]
            # interior
            m = nb == 0
            r[m] = np.stack([np.full(m.sum(), 3), np.full(m.sum(), e), i[m], j[m], k[m], 0 * i[m], 0 * i[m]], -1)
            # vertices
            m = nb == 3
            lv = corner[m] @ np.array([1, 2, 4])
            r[m] = np.stack([0 * lv, v[lv], 0 * lv, 0 * lv, 0 * lv, 0 * lv, 0 * lv], -1)
            # edges: one free direction d
            m = nb == 2
            free = np.argmax(~onb[m], -1)
            c = corner[m].copy()
            off = idx[m][np.arange(free.size), free]
            c[np.arange(free.size), free] = 0
            va = v[c @ np.array([1, 2, 4])]
            c[np.arange(free.size), free] = 1
            vb = v[c @ np.array([1, 2, 4])]
            swap = va > vb
            lo_, hi_ = np.where(swap, vb, va), np.where(swap, va, vb)
            r[m] = np.stack([0 * lo_ + 1, lo_, hi_, np.where(swap, L - off, off), 0 * lo_, 0 * lo_, 0 * lo_], -1)
            # faces: one fixed direction d, two free directions (d1 < d2)
            m = nb == 1
            fixed = np.argmax(onb[m], -1)
            nn = fixed.size
            d1 = np.where(fixed == 0, 1, 0)
            d2 = np.where(fixed == 2, 1, 2)
            p, q = idx[m][np.arange(nn), d1], idx[m][np.arange(nn), d2]
            base = corner[m] * 0
            base[np.arange(nn), fixed] = corner[m][np.arange(nn), fixed]
            wgt = np.array([1, 2, 4])
            F = np.zeros((nn, 2, 2), dtype=np.int64)                 # F[a, b]: vertex at (d1 = a, d2 = b)
            for a in (0, 1):
                for b in (0, 1):
                    c = base.copy()
                    c[np.arange(nn), d1] = a
                    c[np.arange(nn), d2] = b
                    F[:, a, b] = v[c @ wgt]
            flat = F.reshape(nn, 4)
            o = np.argmin(flat, -1)
            a0, b0 = o // 2, o % 2
            na = F[np.arange(nn), 1 - a0, b0]                        # neighbour of the origin along d1
            nbv = F[np.arange(nn), a0, 1 - b0]                       # ... along d2
            pp = np.where(a0 == 0, p, L - p)
            qq = np.where(b0 == 0, q, L - q)
            first_is_d1 = na < nbv
            c1, c2 = np.where(first_is_d1, pp, qq), np.where(first_is_d1, qq, pp)
            sv = np.sort(flat, -1)
            r[m] = np.stack([0 * c1 + 2, sv[:, 0], sv[:, 1], sv[:, 2], sv[:, 3], c1, c2], -1)
        _, inv = np.unique(rows.reshape(-1, 7), axis=0, return_inverse=True)
        return torch.as_tensor(inv.reshape(self.nelv, lx, lx, lx).astype(np.int64), device=device)


def load_hex_fixture(path, lx):
    """tests/golden/debugging_pipe_mesh.npz (made from /root/reference/data/debugging_pipe.nmsh)."""
    d = np.load(path)
    return HexMesh(d["vertex_id"], d["vertex_xyz"], lx, name=os.path.basename(path))


# ---- deterministic node-keyed pseudo-random numbers (splitmix64) ---------------------------------------
def _i64(v):
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


_GAMMA, _M1, _M2 = _i64(0x9E3779B97F4A7C15), _i64(0xBF58476D1CE4E5B9), _i64(0x94D049BB133111EB)


def _lsr(x, s):
    """logical shift right on two's-complement int64 tensors"""
    return (x >> s) & ((1 << (64 - s)) - 1)


def hash_uniform(keys, seed):
    """U[0,1) per key, identical on every device/rank (wrap-around int64 arithmetic)."""
    z = keys * _GAMMA + _i64(seed * 0x2545F4914F6CDD1D + 0x1234567)
    z = (z ^ _lsr(z, 30)) * _M1
    z = (z ^ _lsr(z, 27)) * _M2
    z = z ^ _lsr(z, 31)
    return _lsr(z, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def hash_uniform_numpy(keys, seed):
    """numpy twin of hash_uniform (uint64 arithmetic) -- used by the tests to pin the generator."""
    k = np.asarray(keys).astype(np.uint64)
    with np.errstate(over="ignore"):
        z = k * np.uint64(0x9E3779B97F4A7C15) + np.uint64((seed * 0x2545F4914F6CDD1D + 0x1234567) & ((1 << 64) - 1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


@dataclass
class Fields:
    ub: list          # base flow U_b (3)
    v: list           # adjoint velocity u-dagger (3)
    rho: torch.Tensor  # filtered design rho-tilde in [0,1]


def make_fields(brick, x, y, z, keys):
    """configs[1] recipe: rho ~ U(0,1) per global node (seed 1234); smooth analytic base flow; adjoint
    velocity = smooth part + 1e-2 * node-keyed noise (seed 4321)."""
    X = (x - brick.origin[0]) / brick.length[0]
    Y = (y - brick.origin[1]) / brick.length[1]
    Z = (z - brick.origin[2]) / brick.length[2]
    pi = np.pi
    r = torch.sqrt(X * X + Y * Y + Z * Z)
    amp = 1.0 + 0.1 * r
    ub = [torch.sin(pi * X) * torch.cos(pi * Y) * torch.cos(pi * Z) * amp,
          -torch.cos(pi * X) * torch.sin(pi * Y) * torch.cos(pi * Z) * amp,
          0.3 * torch.sin(pi * X) * torch.sin(pi * Y) * torch.cos(pi * Z) * amp]
    v = [torch.cos(2 * pi * X) * torch.sin(pi * Y) + 0.5 * Z + 1e-2 * (hash_uniform(keys, 4321) - 0.5),
         torch.sin(pi * X) * torch.cos(2 * pi * Z) - 0.25 * Y + 1e-2 * (hash_uniform(keys, 4322) - 0.5),
         torch.cos(pi * Y) * torch.sin(2 * pi * X) + 0.1 * X * Z + 1e-2 * (hash_uniform(keys, 4323) - 0.5)]
    rho = hash_uniform(keys, 1234)
    return Fields([a.contiguous() for a in ub], [a.contiguous() for a in v], rho.contiguous())


def brinkman_zone_chi(x, y, z, box=((4.75, 5.25), (-0.5, 0.5), (-0.5, 0.0)), perm=1000.0):
    """configs[0]/[2]: chi = 1000 inside the `lowperm` box, 0 elsewhere
    (/root/reference/examples/permeability_block/permeability_3.case:73-90)."""
    inside = ((x >= box[0][0]) & (x <= box[0][1]) & (y >= box[1][0]) & (y <= box[1][1])
              & (z >= box[2][0]) & (z <= box[2][1]))
    return torch.where(inside, torch.full_like(x, perm), torch.zeros_like(x))


def interface_candidates(brick, device="cpu"):
    """Bool mask (nelv, lx, lx, lx): dofs on a brick face that is an interior face of the global box,
    i.e. the only dofs that can be shared with another rank."""
    lx = brick.lx
    gi, gj, gk = _lattice(brick, device)
    m = torch.zeros((brick.nelv, lx, lx, lx), dtype=torch.bool, device=device)
    for g, d in ((gi, 0), (gj, 1), (gk, 2)):
        lo = brick.offset[d] * (lx - 1)
        hi = (brick.offset[d] + brick.ne[d]) * (lx - 1)
        gmax = brick.ne_global[d] * (lx - 1)
        if lo > 0:
            m |= (g == lo)
        if hi < gmax:
            m |= (g == hi)
    return m
