"""The README configuration at full size (-m 300,100: 29 738 cells, 657 740 DoFs, 41.2 M non-zeros) through
size-independent properties -- the oracle would need minutes here, these identities need neither side:
symmetry of the Stokes Jacobian before the Dirichlet step, B = Bt^T, Mp row sums = fluid area / nu, constant
pressure in the kernel of Bt's interior rows, the block SpMV (TMA-fed kernel with paired columns) against a
scipy product of the downloaded blocks, linearity of the product, SGS by its defining residual identity, and
the row counts of SURVEY.md section 8."""
import numpy as np
import pytest
import scipy.sparse as sp

import nsxlib as N

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    d = N.Disc.generate(300, 100)
    dev = N.Device(d)
    return d, dev


def test_sizes_of_config_1(big):
    d, dev = big
    assert (d.ncells, d.n_u, d.n_p) == (29738, 537912, 119828)
    assert [dev.nnz(b) for b in (N.BLOCK_F, N.BLOCK_BT, N.BLOCK_B, N.BLOCK_MP)] == [26790480, 7206232, 7206232, 1906736]


def test_stokes_structure_at_full_size(big):
    d, dev = big
    nu = 0.1
    dev.vec_set(N.VEC_SOLUTION, 0.0)
    dev.assemble_cells(N.MODE_STOKES, nu)
    F, Bt, B, Mp = (dev.csr(b) for b in (N.BLOCK_F, N.BLOCK_BT, N.BLOCK_B, N.BLOCK_MP))
    assert abs(F - F.T).max() <= 1e-13 * abs(F).max()
    assert abs(B - Bt.T).max() <= 1e-13 * abs(B).max()
    cv = d.array("CELL_VERTICES").reshape(d.ncells, d.nvpc, 2)
    area = np.sum((cv[:, 1, 0] - cv[:, 0, 0]) * (cv[:, 2, 1] - cv[:, 0, 1]))
    assert abs(Mp.sum() * nu - area) <= 1e-11 * area
    flux = Bt @ np.ones(d.n_p)
    interior = np.ones(d.n_u, bool)
    interior[d.array("BC_DOF")] = False
    out = d.array("CELL_DOFS").reshape(d.ncells, -1)[d.array("OUTLET_CELL")].ravel()
    interior[out[out < d.n_u]] = False
    assert np.abs(flux[interior]).max() <= 1e-12


def test_block_spmv_and_sgs_at_full_size(big):
    d, dev = big
    dev.upload(N.VEC_SOLUTION, N.synthetic_state(d, 1234))
    dev.assemble(N.MODE_NEWTON, False, 1 / 90.0)
    F, Bt, B = (dev.csr(b) for b in (N.BLOCK_F, N.BLOCK_BT, N.BLOCK_B))
    J = sp.bmat([[F, Bt], [B, None]], format="csr")
    rng = np.random.default_rng(42)
    x, z = rng.uniform(-1, 1, d.n), rng.uniform(-1, 1, d.n)
    ref = J @ x
    for flag in (3, 2, 1, 0):           # every SpMV kernel: same product up to summation order
        dev.set_option(N.OPT_STREAM_SPMV, flag)
        y = dev.spmv(N.BLOCK_J, x)
        assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max(), flag
        yf = dev.spmv(N.BLOCK_F, x[: d.n_u])
        assert np.abs(yf - F @ x[: d.n_u]).max() <= 1e-13 * np.abs(ref).max(), flag
    dev.set_option(N.OPT_STREAM_SPMV, 3)
    # linearity: J (a x + z) = a J x + J z
    lin = dev.spmv(N.BLOCK_J, 0.37 * x + z)
    assert np.abs(lin - (0.37 * ref + J @ z)).max() <= 1e-12 * np.abs(ref).max()
    # SGS by definition, default elimination order (CTA-local blocks, multicolour inside): y = (D + U)^-1 D (D + L)^-1 x on the
    # block-diagonal part of F in the elimination order  <=>  (D + L) D^-1 (D + U) y = x on the permuted, block-filtered matrix
    xu = x[: d.n_u]
    y = dev.inner_apply(N.BLOCK_F, 0, xu)
    off, perm = dev.sweep_blocks(N.BLOCK_F)
    assert len(off) - 1 == 2 * 148 or len(off) - 1 >= 64      # two CTA-local blocks per SM on a B200
    blk = np.empty(d.n_u, dtype=np.int64)
    for b in range(len(off) - 1):
        blk[perm[off[b]:off[b + 1]]] = b
    Fc = F.tocoo()
    keep = blk[Fc.row] == blk[Fc.col]
    print("couplings kept by the block-local sweeps:", keep.mean())
    Fb = sp.csr_matrix((Fc.data[keep], (Fc.row[keep], Fc.col[keep])), shape=F.shape)
    Fp = Fb[perm][:, perm].tocsr()
    Dg = Fp.diagonal()
    Lo, Up = sp.tril(Fp, -1, format="csr"), sp.triu(Fp, 1, format="csr")
    yp = y[perm]
    t = Dg * yp + Up @ yp                      # (D + U) y
    lhs = t + Lo @ (t / Dg)                    # (D + L) D^-1 t
    assert np.abs(lhs - xu[perm]).max() <= 1e-11 * np.abs(xu).max()
    assert dev.stat("LEVELS_F") <= 64


SLOW_PATH_SCRIPT = r'''
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.join(os.environ["NSX_ROOT"], "tests")); sys.path.insert(0, os.environ["NSX_ROOT"])
import nsxlib as N
d = N.Disc.generate(300, 100)
dev = N.Device(d, ordering=1)   # the colour-phased persistent kernel of elimination order 1
dev.upload(N.VEC_SOLUTION, N.synthetic_state(d, 1234))
dev.assemble(N.MODE_NEWTON, False, 1 / 90.0)
F = dev.csr(N.BLOCK_F)
x = np.random.default_rng(7).uniform(-1, 1, d.n_u)
y = dev.inner_apply(N.BLOCK_F, 0, x)
perm = dev.ordering(N.BLOCK_F)
yp, xp = y[perm], x[perm]
# SGS: (D + L) D^-1 (D + U) y = x on the permuted matrix
Fp = F[perm][:, perm].tocsr()
Dg = Fp.diagonal()
t = Dg * yp + sp.triu(Fp, 1, format="csr") @ yp
lhs = t + sp.tril(Fp, -1, format="csr") @ (t / Dg)
assert np.abs(lhs - xp).max() <= 1e-11 * np.abs(x).max()
print("SLOW_PATH_OK levels", dev.stat("LEVELS_F"))
'''


def test_sweep_rows_beyond_the_grid(tmp_path):
    """With 256 threads per CTA the persistent sweep covers 9 472 rows per phase, fewer than a level of the 300x100 block
    holds (~16 800): the rows beyond the grid take the stage-and-consume path inside the same phase."""
    import os
    import subprocess
    import sys
    script = tmp_path / "slow_path.py"
    script.write_text(SLOW_PATH_SCRIPT)
    env = dict(os.environ, NSX_ROOT=N.ROOT, NSX_SWEEP_THREADS="256")
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "SLOW_PATH_OK" in r.stdout
