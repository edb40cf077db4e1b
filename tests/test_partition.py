"""Host logic of the row-partitioned path: local views of a partitioned discretisation
(owned + ghost numbering, owned-row sparsity, ghost import plans) against the global discretisation,
and the ghost import itself run by two real processes over gloo.  No GPU involved: the device side of
the same plans is covered by tests/test_gpu_multi.py on a multi-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import nsxlib as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def entry_value(rows, cols):
    """deterministic matrix entry from its global (row, col): every rank can fill its share alone"""
    return np.sin(rows * 12.9898 + cols * 78.233) + 0.25


def global_csr(d, name, shape):
    rp, col = d.pattern(name)
    rows = np.repeat(np.arange(shape[0]), np.diff(rp))
    return sp.csr_matrix((entry_value(rows, col.astype(np.int64)), col, rp), shape=shape)


def local_csr(l, name, row_l2g, col_l2g):
    rp, col = l.pattern(name)
    nrows = len(rp) - 1
    rows = np.repeat(np.arange(nrows), np.diff(rp))
    return sp.csr_matrix((entry_value(row_l2g[rows], col_l2g[col]), col, rp), shape=(nrows, len(col_l2g)))


CASES = [(12, 6, False, 2), (14, 6, False, 3), (16, 7, True, 2), (20, 8, True, 4)]


@pytest.mark.parametrize("nx,ny,tri,nranks", CASES)
def test_local_views_tile_the_global_problem(nx, ny, tri, nranks):
    g = N.Disc.generate(nx, ny, triangles=tri, nranks=nranks)
    assert g.nranks == nranks and not g.is_local
    ou, op = g.array("OWNED_U"), g.array("OWNED_P")
    gcd = g.array("CELL_DOFS").reshape(g.ncells, g.dofs_per_cell).astype(np.int64)
    seen_u, seen_p = np.zeros(g.n_u, int), np.zeros(g.n_p, int)
    locs = [g.local(r) for r in range(nranks)]
    for r, l in enumerate(locs):
        assert l.is_local and l.rank == r and l.job_ranks == nranks
        l2gu, l2gp = l.array("L2G_U"), l.array("L2G_P")
        assert l.n_u_owned == ou[r + 1] - ou[r] and l.n_p_owned == op[r + 1] - op[r]
        assert np.array_equal(l2gu[: l.n_u_owned], np.arange(ou[r], ou[r + 1]))
        assert np.array_equal(l2gp[: l.n_p_owned], np.arange(op[r], op[r + 1]))
        # ghosts: ascending global ids, none owned by this rank
        gu, gp = l2gu[l.n_u_owned:], l2gp[l.n_p_owned:]
        assert np.all(np.diff(gu) > 0) and np.all((gu < ou[r]) | (gu >= ou[r + 1]))
        assert np.all(np.diff(gp) > 0) and np.all((gp < op[r]) | (gp >= op[r + 1]))
        seen_u[l2gu[: l.n_u_owned]] += 1
        seen_p[l2gp[: l.n_p_owned]] += 1
        # the local cell table is the global one, renumbered; local cells = every cell touching an owned dof
        cg = l.array("CELL_GLOBAL")
        lcd = l.array("CELL_DOFS").reshape(l.ncells, l.dofs_per_cell).astype(np.int64)
        back = np.where(lcd < l.n_u, l2gu[np.minimum(lcd, l.n_u - 1)], g.n_u + l2gp[np.maximum(lcd - l.n_u, 0)])
        assert np.array_equal(back, gcd[cg])
        isu = gcd < g.n_u
        touches = (((gcd >= ou[r]) & (gcd < ou[r + 1]) & isu) | ((gcd - g.n_u >= op[r]) & (gcd - g.n_u < op[r + 1]) & ~isu)).any(axis=1)
        assert np.array_equal(np.flatnonzero(touches), cg)
        assert np.array_equal(l.array("CELL_OWNED").astype(bool), g.array("CELL_RANK")[cg] == r)
        assert np.array_equal(l.array("CELL_VERTICES").reshape(l.ncells, -1), g.array("CELL_VERTICES").reshape(g.ncells, -1)[cg])
        # owned rows of every block: same entries as the global pattern
        shapes = {"F": (g.n_u, g.n_u), "BT": (g.n_u, g.n_p), "B": (g.n_p, g.n_u), "MP": (g.n_p, g.n_p)}
        for name, (gr, gc) in shapes.items():
            row_l2g = l2gu if name in ("F", "BT") else l2gp
            col_l2g = l2gu if name in ("F", "B") else l2gp
            A = global_csr(g, name, (gr, gc))
            L = local_csr(l, name, row_l2g, col_l2g)
            nown = l.n_u_owned if name in ("F", "BT") else l.n_p_owned
            assert L.shape[0] == nown
            rp, col = l.pattern(name)
            for i in range(nown):   # columns ascending in LOCAL ids (the kernels rely on it)
                assert np.all(np.diff(col[rp[i]:rp[i + 1]]) > 0)
            # same matrix: compare through a product with a global vector
            x = np.cos(np.arange(gc) * 0.37)
            assert np.allclose(L @ x[col_l2g], (A @ x)[row_l2g[:nown]], rtol=0, atol=1e-12)
            assert L.nnz == A[row_l2g[:nown]].nnz
        # Dirichlet list: the owned part of the global one
        gbc = g.array("BC_DOF").astype(np.int64)
        mine = (gbc >= ou[r]) & (gbc < ou[r + 1])
        assert np.array_equal(l.array("BC_DOF").astype(np.int64) + ou[r], gbc[mine])
        assert np.array_equal(l.inlet_values(0.1), g.inlet_values(0.1)[mine])
    assert np.all(seen_u == 1) and np.all(seen_p == 1)
    # halo plans: what a sends to b is exactly what b expects from a, in the same order
    for blk, l2g_name, nown_name in (("U", "L2G_U", "n_u_owned"), ("P", "L2G_P", "n_p_owned")):
        for a, la in enumerate(locs):
            nbr = la.array(f"HALO_{blk}_NBR")
            sp_, si = la.array(f"HALO_{blk}_SEND_PTR"), la.array(f"HALO_{blk}_SEND_IDX")
            rp_ = la.array(f"HALO_{blk}_RECV_PTR")
            assert rp_[-1] == getattr(la, "n_u" if blk == "U" else "n_p") - getattr(la, nown_name)
            for k, b in enumerate(nbr):
                lb = locs[b]
                sent = la.array(l2g_name)[si[sp_[k]:sp_[k + 1]]]
                nb_nbr = list(lb.array(f"HALO_{blk}_NBR"))
                assert a in nb_nbr
                kb = nb_nbr.index(a)
                rb = lb.array(f"HALO_{blk}_RECV_PTR")
                expect = lb.array(l2g_name)[getattr(lb, nown_name) + rb[kb]: getattr(lb, nown_name) + rb[kb + 1]]
                assert np.array_equal(sent, expect)


def test_local_view_of_the_gmsh_mesh():
    g = N.Disc.from_gmsh(N.golden_mesh_path(), nranks=4)
    tot_u = tot_p = 0
    for r in range(4):
        l = g.local(r)
        tot_u += l.n_u_owned
        tot_p += l.n_p_owned
        assert 0 < l.ncells < g.ncells
        assert len(l.array("HALO_U_NBR")) >= 1
    assert tot_u == g.n_u and tot_p == g.n_p
    # lift/drag faces are split over the ranks without loss
    assert sum(len(g.local(r).array("CYL_CELL")) for r in range(4)) == len(g.array("CYL_CELL"))


def test_rank_out_of_range():
    g = N.Disc.generate(8, 4, nranks=2)
    with pytest.raises(RuntimeError):
        g.local(2)
    with pytest.raises(RuntimeError):
        g.local(0).local(0)


WORKER = r'''
import os, sys
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.environ["NSX_ROOT"], "tests")); sys.path.insert(0, os.environ["NSX_ROOT"])
import nsxlib as N
from test_partition import global_csr, local_csr
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = N.Disc.generate(14, 6, nranks=world)
l = g.local(rank)
rng = np.random.default_rng(5)
xg = rng.uniform(-1, 1, g.n_u + g.n_p)                       # the same global vector on every rank
l2gu, l2gp = l.array("L2G_U"), l.array("L2G_P")
x_u = np.zeros(l.n_u); x_u[: l.n_u_owned] = xg[l2gu[: l.n_u_owned]]      # owned entries only; ghosts arrive by the plan
x_p = np.zeros(l.n_p); x_p[: l.n_p_owned] = xg[g.n_u + l2gp[: l.n_p_owned]]
def halo(blk, x, nown):
    nbr = l.array(f"HALO_{blk}_NBR"); sp_ = l.array(f"HALO_{blk}_SEND_PTR"); si = l.array(f"HALO_{blk}_SEND_IDX"); rp_ = l.array(f"HALO_{blk}_RECV_PTR")
    reqs, bufs = [], []
    for k, b in enumerate(nbr):
        s = torch.from_numpy(np.ascontiguousarray(x[si[sp_[k]:sp_[k + 1]]]))
        r = torch.empty(int(rp_[k + 1] - rp_[k]), dtype=torch.float64)
        reqs.append(dist.isend(s, int(b))); reqs.append(dist.irecv(r, int(b))); bufs.append((k, r, s))
    for q in reqs: q.wait()
    for k, r, _ in bufs: x[nown + rp_[k]: nown + rp_[k + 1]] = r.numpy()
halo("U", x_u, l.n_u_owned); halo("P", x_p, l.n_p_owned)
assert np.array_equal(x_u, xg[l2gu]) and np.array_equal(x_p, xg[g.n_u + l2gp])
# block product on the owned rows: y_u = F x_u + Bt x_p, y_p = B x_u -- against the global product
F = local_csr(l, "F", l2gu, l2gu); Bt = local_csr(l, "BT", l2gu, l2gp); B = local_csr(l, "B", l2gp, l2gu)
y_u, y_p = F @ x_u + Bt @ x_p, B @ x_u
GF = global_csr(g, "F", (g.n_u, g.n_u)); GBt = global_csr(g, "BT", (g.n_u, g.n_p)); GB = global_csr(g, "B", (g.n_p, g.n_u))
yg_u, yg_p = GF @ xg[: g.n_u] + GBt @ xg[g.n_u:], GB @ xg[: g.n_u]
assert np.allclose(y_u, yg_u[l2gu[: l.n_u_owned]], rtol=0, atol=1e-12)
assert np.allclose(y_p, yg_p[l2gp[: l.n_p_owned]], rtol=0, atol=1e-12)
# dot product = allreduce of the owned partial sums
t = torch.tensor([float(x_u[: l.n_u_owned] @ x_u[: l.n_u_owned] + x_p[: l.n_p_owned] @ x_p[: l.n_p_owned])], dtype=torch.float64)
dist.all_reduce(t)
assert abs(t.item() - xg @ xg) < 1e-10
dist.barrier()
if rank == 0: print("PARTITION_WORKER_OK")
'''


def test_ghost_import_between_two_processes_gloo(tmp_path):
    """world_size 2 over gloo: each process builds its own local view, imports its ghosts by the
    send / receive plan, multiplies by its owned rows -- the result is the global product."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, NSX_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PARTITION_WORKER_OK" in r.stdout
