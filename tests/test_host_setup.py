"""Host-side set-up stand-in (mesh rules, DoF numbering, block sparsity, boundary lists) against the
only integer known answers the reference holds -- the 154 244-DoF count of the 100x70 mesh
(performance_analysis.ipynb, cell 2) -- and against the size table SURVEY.md section 8 derived
independently from the reference's rules (NSSolverStationary.cpp:12-63, 264-301)."""
import numpy as np
import pytest
import scipy.sparse as sp

import nsxlib as N


def test_dof_count_of_the_published_mesh():
    d = N.Disc.generate(100, 70)
    assert (d.ncells, d.n_u, d.n_p, d.n) == (6942, 126096, 28148, 154244)
    nnz = {k: int(d.pattern(k)[0][-1]) for k in ("F", "BT", "B", "MP")}
    assert nnz == {"F": 6259200, "BT": 1684144, "B": 1684144, "MP": 445808}
    assert nnz["F"] + nnz["BT"] + nnz["B"] == 9627488


def test_readme_mesh_sizes():
    d = N.Disc.generate(300, 100)
    assert (d.ncells, d.n_u, d.n_p) == (29738, 537912, 119828)
    assert int(d.pattern("F")[0][-1]) == 26790480
    assert int(d.pattern("BT")[0][-1]) == 7206232
    assert int(d.pattern("MP")[0][-1]) == 1906736


def test_file_mesh_sizes():
    d = N.Disc.from_gmsh(N.golden_mesh_path())
    assert d.elem == 1 and (d.ncells, d.n_u, d.n_p) == (25619, 104066, 13207)
    assert int(d.pattern("F")[0][-1]) == 2369668
    assert int(d.pattern("BT")[0][-1]) == 490736 == int(d.pattern("B")[0][-1])
    assert int(d.pattern("MP")[0][-1]) == 90859
    bf = d.array("BFACES").reshape(-1, 3)
    counts = {b: int((bf[:, 2] == b).sum()) for b in (6, 7, 8, 10)}
    assert counts == {6: 298, 7: 99, 8: 99, 10: 299}       # physical ids of lab_new/mesh/new_mesh.msh
    # P2 nodes on the Dirichlet boundaries: walls 300 + 298, cylinder 299 + 299, inlet 100 + 99, the two
    # inlet corners shared with the walls counted once; two velocity components each
    assert d.nbc == 2 * (598 + 598 + 199 - 2)
    assert len(d.array("CYL_CELL")) == 299 and len(d.array("OUTLET_CELL")) == 99


@pytest.mark.parametrize("tri", [False, True])
def test_patterns_are_the_cell_couplings(tri):
    d = N.Disc.generate(14, 6, triangles=tri)
    cd = d.array("CELL_DOFS").reshape(d.ncells, d.dofs_per_cell).astype(np.int64)
    assert sorted(np.unique(cd).tolist()) == list(range(d.n))
    is_p = cd[0] >= d.n_u
    # every (row, col) of a block is a pair of dofs of one cell, and every such pair is in the block
    def expected(rows_p, cols_p):
        r = cd[:, is_p == rows_p] - (d.n_u if rows_p else 0)
        c = cd[:, is_p == cols_p] - (d.n_u if cols_p else 0)
        rr = np.repeat(r, c.shape[1], axis=1).ravel()
        cc = np.tile(c, (1, r.shape[1])).ravel()
        m = sp.coo_matrix((np.ones(rr.size), (rr, cc)), shape=(d.n_p if rows_p else d.n_u, d.n_p if cols_p else d.n_u)).tocsr()
        m.sum_duplicates(); m.sort_indices()
        return m
    for name, rp_, cp_ in (("F", False, False), ("BT", False, True), ("B", True, False), ("MP", True, True)):
        rp, col = d.pattern(name)
        e = expected(rp_, cp_)
        assert (e.indptr == rp).all() and (e.indices == col).all(), name
        for i in range(len(rp) - 1):   # columns ascending (Epetra stores sorted rows after compress)
            assert (np.diff(col[rp[i]:rp[i + 1]]) > 0).all()


def test_numbering_walk_and_component_wise_blocks():
    """distribute_dofs walks cells x-fastest, vertices then lines then interior; component_wise keeps
    that order inside the velocity block and the pressure block."""
    d = N.Disc.generate(6, 4)
    cd = d.array("CELL_DOFS").reshape(d.ncells, 41).astype(np.int64)
    # first cell: its velocity dofs are 0..31 in walk order, its pressure dofs n_u..n_u+8
    v = cd[0][cd[0] < d.n_u]
    p = cd[0][cd[0] >= d.n_u]
    assert sorted(v.tolist()) == list(range(32)) and sorted((p - d.n_u).tolist()) == list(range(9))
    # vertex 0 of cell 0 carries [u_x, u_y, p] = velocity 0, 1 and pressure 0
    assert cd[0][:3].tolist() == [0, 1, d.n_u]
    # a shared vertex keeps the number the first cell gave it
    assert cd[1][0] == cd[0][3] and cd[1][1] == cd[0][4] and cd[1][2] == cd[0][5]


def test_boundary_ids_and_dirichlet_values():
    d = N.Disc.generate(44, 20)
    bf = d.array("BFACES").reshape(-1, 3)
    assert set(np.unique(bf[:, 2]).tolist()) == {6, 7, 8, 10}
    assert (bf[:, 2] == 7).sum() == 20 and (bf[:, 2] == 8).sum() == 20
    v = d.inlet_values(0.1)
    y = d.array("BC_Y")
    on = d.array("BC_ON_INLET").astype(bool)
    H = 0.41
    np.testing.assert_allclose(v[on], 4 * 0.1 * y[on] * (H - y[on]) / H ** 2, rtol=0, atol=1e-16)
    assert (v[~on] == 0).all() and on.sum() > 0
    assert (np.diff(d.array("BC_DOF").astype(np.int64)) > 0).all() and d.array("BC_DOF").max() < d.n_u
    # inlet corners belong to the walls (last writer of interpolate_boundary_values wins): value 0 there
    assert abs(v[on]).min() > 0


def test_partition_owned_ranges():
    d = N.Disc.generate(24, 8, nranks=4)
    ou, op = d.array("OWNED_U"), d.array("OWNED_P")
    assert d.nranks == 4 and ou[0] == 0 and ou[-1] == d.n_u and op[-1] == d.n_p
    assert (np.diff(ou) > 0).all() and (np.diff(op) > 0).all()
    rank = d.array("CELL_RANK")
    cd = d.array("CELL_DOFS").reshape(d.ncells, 41).astype(np.int64)
    # a dof belongs to the lowest rank whose cells touch it; interior dofs of a rank's cells lie in its range
    owner = np.full(d.n, 99)
    for r in range(4):
        np.minimum.at(owner, cd[rank == r].ravel(), r)
    for r in range(4):
        u = np.nonzero(owner[: d.n_u] == r)[0]
        assert u.min() == ou[r] and u.max() == ou[r + 1] - 1


def test_gmsh_reader_rejects_garbage(tmp_path):
    p = tmp_path / "bad.msh"
    p.write_text("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
    with pytest.raises(RuntimeError):
        N.Disc.from_gmsh(str(p))
    with pytest.raises(RuntimeError):
        N.Disc.from_gmsh(str(tmp_path / "missing.msh"))
