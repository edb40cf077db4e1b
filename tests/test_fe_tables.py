"""Reference-cell tables (product fe.hpp, exported through nsx_disc_array) against closed forms, and
the oracle's independently coded tables against the same identities via assembled matrices."""
import ctypes as C

import numpy as np
import pytest

import nsxlib as N

MAXD, MAXV, MAXP, MAXQ, MAXQF, MAXF = 41, 16, 9, 16, 4, 4


class FET(C.Structure):
    _fields_ = [("elem", C.c_int), ("nvpc", C.c_int), ("nvn", C.c_int), ("npn", C.c_int), ("ndofs", C.c_int), ("nq", C.c_int),
                ("nqf", C.c_int), ("nfaces", C.c_int), ("dof_comp", C.c_int * MAXD), ("dof_node", C.c_int * MAXD),
                ("qp", C.c_double * 2 * MAXQ), ("qw", C.c_double * MAXQ), ("Nv", C.c_double * MAXQ * MAXV),
                ("dNv", C.c_double * 2 * MAXQ * MAXV), ("Np", C.c_double * MAXQ * MAXP), ("qwf", C.c_double * MAXQF),
                ("Nvf", C.c_double * MAXQF * MAXV * MAXF), ("dNvf", C.c_double * 2 * MAXQF * MAXV * MAXF),
                ("Npf", C.c_double * MAXQF * MAXP * MAXF), ("qpf", C.c_double * MAXQF)]


def tables(tri):
    d = N.Disc.generate(6, 4, triangles=tri)
    raw = d.array("FE_TABLES")
    assert raw.size == C.sizeof(FET)
    t = FET.from_buffer_copy(raw.tobytes())
    return d, t


@pytest.mark.parametrize("tri", [False, True])
def test_quadrature_and_partition_of_unity(tri):
    d, t = tables(tri)
    nq, nvn, npn = t.nq, t.nvn, t.npn
    assert (nq, nvn, npn, t.ndofs) == ((7, 6, 3, 15) if tri else (16, 16, 9, 41))
    qw = np.array(t.qw[:nq]); qp = np.array([list(t.qp[q]) for q in range(nq)])
    assert abs(qw.sum() - (0.5 if tri else 1.0)) < 1e-15
    # exactness: QGauss(4) integrates degree 7 per direction, the 7-point simplex rule degree 5
    if tri:
        for a, b in [(1, 0), (0, 1), (2, 1), (3, 2), (5, 0), (1, 4)]:
            from math import factorial
            exact = factorial(a) * factorial(b) / factorial(a + b + 2)
            assert abs((qw * qp[:, 0] ** a * qp[:, 1] ** b).sum() - exact) < 1e-15
    else:
        for a, b in [(7, 0), (3, 6), (7, 7)]:
            assert abs((qw * qp[:, 0] ** a * qp[:, 1] ** b).sum() - 1.0 / ((a + 1) * (b + 1))) < 1e-15
    Nv = np.array([[t.Nv[a][q] for q in range(nq)] for a in range(nvn)])
    dNv = np.array([[[t.dNv[a][q][k] for k in range(2)] for q in range(nq)] for a in range(nvn)])
    Np = np.array([[t.Np[m][q] for q in range(nq)] for m in range(npn)])
    np.testing.assert_allclose(Nv.sum(0), 1.0, atol=1e-14)
    np.testing.assert_allclose(Np.sum(0), 1.0, atol=1e-14)
    np.testing.assert_allclose(dNv.sum(0), 0.0, atol=1e-13)
    # face tables: partition of unity on every face, face weights sum to 1
    assert abs(sum(t.qwf[: t.nqf]) - 1.0) < 1e-15
    for f in range(t.nfaces):
        s = np.array([[t.Nvf[f][a][q] for q in range(t.nqf)] for a in range(nvn)]).sum(0)
        np.testing.assert_allclose(s, 1.0, atol=1e-14)


def test_q3_uses_gauss_lobatto_nodes_and_q2_equidistant():
    """FE_Q(3) support points are Gauss-Lobatto {0, (1-1/sqrt5)/2, (1+1/sqrt5)/2, 1}: the line dofs of the
    inlet boundary sit at those heights (NSX_DA_BC_Y), FE_Q(2) nodes at the midpoints."""
    d = N.Disc.generate(4, 1)   # one cell in y: dy = 0.41
    y = np.unique(np.round(d.array("BC_Y")[d.array("BC_ON_INLET").astype(bool)] / 0.41, 12))
    g = np.array([(1 - 1 / np.sqrt(5)) / 2, (1 + 1 / np.sqrt(5)) / 2])
    np.testing.assert_allclose(y, g, atol=1e-12)   # the two vertices belong to the walls


@pytest.mark.parametrize("tri", [False, True])
def test_local_dof_layout(tri):
    d, t = tables(tri)
    comp = list(t.dof_comp[: t.ndofs]); node = list(t.dof_node[: t.ndofs])
    nv = t.nvpc
    for v in range(nv):   # vertex v -> [u_x, u_y, p]
        assert comp[3 * v: 3 * v + 3] == [0, 1, 2] and node[3 * v: 3 * v + 3] == [v, v, v]
    if tri:
        assert comp[9:] == [0, 1] * 3 and node[9:] == [3, 3, 4, 4, 5, 5]
    else:
        assert comp[12:17] == [0, 0, 1, 1, 2] and node[12:17] == [4, 5, 4, 5, 4]
        assert comp[32:] == [0] * 4 + [1] * 4 + [2] and node[32:] == [12, 13, 14, 15] * 2 + [8]
    assert comp.count(0) == comp.count(1) == t.nvn and comp.count(2) == t.npn
