"""CPU tests of the host-side workload generators (product-side host logic)."""
import numpy as np
import torch

import neko_top_b200  # noqa: F401
from neko_top_b200 import sem, workloads


def test_hash_uniform_matches_numpy_twin():
    k = torch.arange(0, 100000, 37, dtype=torch.int64) * 7919 + 13
    for seed in (1234, 4321):
        a = workloads.hash_uniform(k, seed).numpy()
        b = workloads.hash_uniform_numpy(k.numpy(), seed)
        assert np.array_equal(a, b)
        assert 0.0 <= a.min() and a.max() < 1.0 and abs(a.mean() - 0.5) < 0.02


def test_fields_are_c0_and_partition_independent():
    """Fields are functions of the global node: a rank's brick reproduces the global values exactly."""
    lx = 5
    whole = workloads.BoxBrick(lx=lx, ne=(4, 2, 2), ne_global=(4, 2, 2), length=(2.0, 1.0, 1.0))
    xw, yw, zw = workloads.coords(whole)
    kw = workloads.node_keys(whole)
    fw = workloads.make_fields(whole, xw, yw, zw, kw)
    # C0: equal keys -> equal values
    flat_k = kw.reshape(-1).numpy()
    order = np.argsort(flat_k, kind="stable")
    same = flat_k[order][1:] == flat_k[order][:-1]
    for fld in [fw.rho] + fw.v + fw.ub + [xw, yw, zw]:
        a = fld.reshape(-1).numpy()[order]
        assert np.array_equal(a[1:][same], a[:-1][same])
    for rank in range(2):
        part = workloads.config_weak(rank, 2, ne_per_gpu=2, lx=lx)
        assert part.ne_global == (4, 2, 2)
        xp, yp, zp = workloads.coords(part)
        kp = workloads.node_keys(part)
        fp = workloads.make_fields(part, xp, yp, zp, kp)
        # local element e=(ex,ey,ez) of rank r is global element (ex+2r, ey, ez)
        e = np.arange(part.nelv)
        ex, ey, ez = e % 2, (e // 2) % 2, e // 4
        ge = (ex + 2 * rank) + 4 * (ey + 2 * ez)
        assert np.array_equal(kp.numpy(), kw.numpy()[ge])
        assert np.array_equal(fp.rho.numpy(), fw.rho.numpy()[ge])
        assert np.array_equal(fp.v[1].numpy(), fw.v[1].numpy()[ge])
        assert np.array_equal(xp.numpy(), xw.numpy()[ge])


def test_configs():
    d = workloads.config_duct(6)
    assert d.nelv == 1536 and d.n == 331776                      # BASELINE configs[0]
    assert workloads.config_duct(8).n == 786432                  # configs[2]
    assert workloads.config_box(32).n == 16777216                # configs[1]
    assert workloads.config_weak(0, 8).n == 134217728            # configs[3]
    assert workloads.config_sweep(8).ne == (58, 58, 58)          # configs[4]
    assert abs(sem.algorithmic_bytes_per_dof(8) - 200.4) < 0.05
    assert sem.algorithmic_flops_per_dof(8) == 424


def test_brinkman_zone():
    d = workloads.config_duct(4)
    x, y, z = workloads.coords(d)
    chi = workloads.brinkman_zone_chi(x, y, z)
    assert set(np.unique(chi.numpy())) == {0.0, 1000.0}
    inside = chi > 0
    assert float(x[inside].min()) >= 4.75 and float(x[inside].max()) <= 5.25 and float(z[inside].max()) <= 0.0


def test_reference_pipe_mesh_fixture():
    """tests/golden/debugging_pipe_mesh.npz = the hexahedra of /root/reference/data/debugging_pipe.nmsh
    (160 elements, 275 vertices, 10x4x4 on [0,10]x[-.5,.5]^2, the file's own numbering).  The topological
    node keys (vertex / edge / face / interior) must induce exactly the partition given by coincident
    coordinates, the element Jacobians must be positive and the mass matrix must sum to the volume."""
    import os
    import numpy as np
    import torch
    from neko_top_b200 import sem, workloads
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "debugging_pipe_mesh.npz")
    raw = np.load(path)
    assert raw["vertex_id"].shape == (160, 8) and np.unique(raw["vertex_id"]).size == 275
    if os.path.exists("/root/reference/data/debugging_pipe.nmsh"):      # the fixture is the reference's mesh
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            "mk", os.path.join(os.path.dirname(path), "make_pipe_fixture.py"))
        mk = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mk)
        vid, xyz = mk.read_nmsh_hexes("/root/reference/data/debugging_pipe.nmsh")
        assert np.array_equal(vid, raw["vertex_id"]) and np.array_equal(xyz, raw["vertex_xyz"])
    for lx in (4, 7):
        m = workloads.load_hex_fixture(path, lx)
        x, y, z = m.coords()
        keys = m.node_keys().reshape(-1).numpy()
        X = np.stack([a.reshape(-1).numpy() for a in (x, y, z)], 1)
        _, by_coord = np.unique(np.round(X * 1e9).astype(np.int64), axis=0, return_inverse=True)
        nnode = (10 * (lx - 1) + 1) * (4 * (lx - 1) + 1) ** 2
        assert np.unique(keys).size == nnode == np.unique(by_coord).size
        assert np.unique(np.stack([keys, by_coord.reshape(-1)], 1), axis=0).shape[0] == nnode
        _, jac, B = sem.geometric_factors(x, y, z, sem.Space(lx))
        assert float(jac.min()) > 0.0 and abs(float(B.sum()) - 10.0) < 1e-12


def test_tile_order_is_a_permutation_with_short_reach():
    """workloads.tile_order: a permutation of the elements; inside a tile column the z-neighbour of an element is
    exactly tx*ty positions later and the x-neighbour is adjacent."""
    import numpy as np
    from neko_top_b200 import workloads
    b = workloads.BoxBrick(lx=8, ne=(12, 8, 5))
    o = workloads.tile_order(b, (4, 4))
    assert sorted(o.tolist()) == list(range(b.nelv))
    pos = np.empty(b.nelv, dtype=np.int64)
    pos[o] = np.arange(b.nelv)
    e = lambda x, y, z: x + 12 * (y + 8 * z)
    assert pos[e(1, 0, 0)] - pos[e(0, 0, 0)] == 1
    assert pos[e(0, 1, 0)] - pos[e(0, 0, 0)] == 4
    assert pos[e(0, 0, 1)] - pos[e(0, 0, 0)] == 16
    assert pos[e(4, 0, 0)] - pos[e(0, 0, 0)] == 16 * 5          # next tile column


def test_composite_dealias_matrix_identities():
    """What the tensor-core dealiased kernel (csrc/advop_mma_kernel.cuh) relies on: DJ = D_fine J differentiates AND
    interpolates along an axis in one product.  Checked on polynomials up to the GLL degree (exact for them): DJ u =
    u'(fine points); D_fine J = J D_gll (both are the derivative of the GLL interpolant at the fine points); the
    rows of J sum to one.  Every order the library instantiates."""
    from neko_top_b200 import sem
    for lx in range(4, 11):
        ds = sem.DealiasSpace(lx)
        zg, _ = sem.zwgll(lx)
        J, Dd = ds.interp, ds.dxd
        DJ = Dd @ J
        assert J.shape == (ds.lxd, lx) and np.allclose(J.sum(axis=1), 1.0, atol=1e-13)
        assert np.allclose(DJ, J @ sem.dgll(zg), atol=1e-11 * lx)
        for deg in range(lx):
            u = zg ** deg
            du = deg * ds.zd ** (deg - 1) if deg > 0 else np.zeros(ds.lxd)
            assert np.allclose(J @ u, ds.zd ** deg, atol=1e-12)
            assert np.allclose(DJ @ u, du, atol=1e-10)
