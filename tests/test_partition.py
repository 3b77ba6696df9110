"""Multi-process CPU tests (gloo, world_size 2 and 4) of the shared-node discovery that prepares the
NCCL exchange behind gs_op (neko-top_b200/partition.py).  The exchange itself is emulated here in numpy
from the lists each rank produced -- test infrastructure, not a product fallback -- and compared with
the oracle's gather-scatter on the undivided global mesh."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import neko_top_b200  # noqa: F401
from neko_top_b200 import workloads
from _dist_worker import worker


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(nranks, ne, lx, tmp_path):
    mp.spawn(worker, args=(nranks, _free_port(), ne, lx, str(tmp_path)), nprocs=nranks, join=True)
    return [dict(np.load(os.path.join(tmp_path, f"rank{r}.npz"))) for r in range(nranks)]


@pytest.mark.parametrize("nranks", [2, 4])
def test_shared_node_lists(oracle, nranks, tmp_path):
    lx, ne = 4, 2
    R = _run(nranks, ne, lx, tmp_path)
    N = lx ** 3
    for r in range(nranks):
        a = R[r]
        assert np.all(np.diff(a["shared_key"]) > 0)
        assert np.array_equal(a["keys"][a["shared_dof"]], a["shared_key"])
        # smallest local dof carrying the key
        for s in range(0, a["shared_key"].size, 5):
            assert a["shared_dof"][s] == np.nonzero(a["keys"] == a["shared_key"][s])[0][0]
        # brute force: shared with q  <=>  key present on both ranks
        expect_neigh = []
        for q in range(nranks):
            if q == r:
                continue
            common = np.intersect1d(a["keys"], R[q]["keys"])
            if common.size:
                expect_neigh.append(q)
                j = list(a["neigh_rank"]).index(q)
                mine = a["shared_key"][a["neigh_idx"][a["neigh_off"][j]:a["neigh_off"][j + 1]]]
                assert np.array_equal(mine, common)            # ascending key: both sides agree
                jq = list(R[q]["neigh_rank"]).index(r)
                theirs = R[q]["shared_key"][R[q]["neigh_idx"][R[q]["neigh_off"][jq]:R[q]["neigh_off"][jq + 1]]]
                assert np.array_equal(mine, theirs)
        assert list(a["neigh_rank"]) == expect_neigh
        # boundary elements = elements holding a shared key
        is_sh = np.isin(a["keys"], a["shared_key"])
        assert np.array_equal(a["bnd_elem"], np.unique(np.nonzero(is_sh)[0] // N))

    # ---- emulate local gs + exchange + ascending-rank summation, compare with the global oracle ----
    px, py, pz = workloads.rank_grid(nranks)
    whole = workloads.BoxBrick(lx=lx, ne=(ne * px, ne * py, ne * pz))
    gkeys = workloads.node_keys(whole).reshape(-1).numpy()
    rng = np.random.default_rng(0)
    gval = rng.standard_normal(gkeys.size)
    cid, nc = oracle.gs_classes(gkeys)
    gref = oracle.gs_add(gval, cid, nc)
    ref_by_key = dict(zip(gkeys.tolist(), gref.tolist()))
    val_by_place = {}
    # a rank's dof values: take them from the global field by matching (element, point) through keys is
    # ambiguous for duplicated nodes, so build rank fields directly from the global element numbering
    loc = []
    for r in range(nranks):
        b = workloads.config_weak(r, nranks, ne_per_gpu=ne, lx=lx)
        e = np.arange(b.nelv)
        ex, ey, ez = e % ne + b.offset[0], (e // ne) % ne + b.offset[1], e // (ne * ne) + b.offset[2]
        ge = ex + whole.ne[0] * (ey + whole.ne[1] * ez)
        f = gval.reshape(whole.nelv, N)[ge].reshape(-1).copy()
        lc, lnc = oracle.gs_classes(R[r]["keys"])
        loc.append(oracle.gs_add(f, lc, lnc))                       # local direct-stiffness sum
    for r in range(nranks):
        a = R[r]
        total = {int(k): [(r, loc[r][d])] for k, d in zip(a["shared_key"], a["shared_dof"])}
        for j, q in enumerate(a["neigh_rank"]):
            jq = list(R[q]["neigh_rank"]).index(r)
            send_idx = R[q]["neigh_idx"][R[q]["neigh_off"][jq]:R[q]["neigh_off"][jq + 1]]
            recv = loc[q][R[q]["shared_dof"][send_idx]]            # what q packs for r
            mine = a["shared_key"][a["neigh_idx"][a["neigh_off"][j]:a["neigh_off"][j + 1]]]
            for k, x in zip(mine, recv):
                total[int(k)].append((int(q), x))
        out = loc[r].copy()
        for k, contrib in total.items():
            s = 0.0
            for _, x in sorted(contrib):                           # ascending rank, own value in place
                s += x
            out[a["keys"] == k] = s
        expect = np.array([ref_by_key[int(k)] for k in a["keys"]])
        assert np.allclose(out, expect, rtol=1e-13, atol=1e-13)
