"""One rank of a row-partitioned run on the GPUs of one box (launched by torch.distributed.run, one process per
GPU, NCCL): assemble + SpMV + solve + update + lift/drag through the C ABI on this rank's share, checked on rank 0
against the CPU oracle of the global problem with the same rank-local preconditioner blocks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nsxlib as N  # noqa: E402


def gather_global(l, g, local_vec):
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, (l.array("L2G_U")[: l.n_u_owned].copy(), l.array("L2G_P")[: l.n_p_owned].copy(), local_vec))
    out = np.zeros(g.n_u + g.n_p)
    for gu, gp, v in parts:
        out[gu] = v[: len(gu)]
        out[g.n_u + gp] = v[len(gu):]
    return out


def main():
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    elem = sys.argv[1] if len(sys.argv) > 1 else "quad"
    prec = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    defaults = len(sys.argv) > 3 and sys.argv[3] == "defaults"   # the library's default kernels / orders instead of the parity configuration
    g = N.Disc.generate(24, 10, nranks=world) if elem == "quad" else N.Disc.generate(18, 8, triangles=True, nranks=world)
    l = g.local(rank)
    ids = [N.Device.new_comm_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    dev = N.Device(l, device_id=local_rank, comm_id=ids[0], block_rows=128) if defaults else N.Device(l, device_id=local_rank, ordering=0, ortho=0, comm_id=ids[0])
    nu = 0.1
    sol = N.synthetic_state(g, 7, noise=1e-4)
    dev.upload(N.VEC_SOLUTION, l.scatter_owned(sol, g.n_u))
    dev.upload(N.VEC_DELTA, np.zeros(dev.n))
    # ghost import: the ghosts of the solution arrive from their owners
    dev.halo_exchange(N.VEC_SOLUTION)
    gu, gp = dev.download_ghosts(N.VEC_SOLUTION)
    l2gu, l2gp = l.array("L2G_U"), l.array("L2G_P")
    assert np.array_equal(gu, sol[l2gu[l.n_u_owned:]]) and np.array_equal(gp, sol[g.n_u + l2gp[l.n_p_owned:]])
    mode = N.MODE_STOKES if defaults else N.MODE_NEWTON   # defaults: the Stokes branch, whose node view of F is the multi-rank path to cover
    r_d = dev.assemble(mode, True, nu)
    view = dev.view()
    # block product on the assembled matrices
    x = np.random.default_rng(3).uniform(-1, 1, g.n_u + g.n_p)
    dev.upload(N.VEC_TMP0, l.scatter_owned(x, g.n_u))
    dev._ck(N.nsx().nsx_spmv(dev.h, N.BLOCK_J, N.VEC_TMP0, N.VEC_TMP1))
    y_d = gather_global(l, g, dev.download(N.VEC_TMP1))
    res_d = gather_global(l, g, dev.download(N.VEC_RESIDUAL))
    rc_d, it_d, fr_d = dev.solve(N.STATIONARY, 1, prec, 1e-12, 3000)
    x_d = gather_global(l, g, dev.download(N.VEC_DELTA))
    dev.save_eval_point()
    dev.update(1.0)
    r2_d = dev.assemble(mode, False, nu)
    drag_d, lift_d = dev.lift_drag(nu)
    stats = (dev.stat("HALO_EXCHANGES"), dev.stat("ALLREDUCES"))
    if rank == 0:
        o = N.Oracle(g)
        o.vec(0)[:] = sol
        o.vec(2)[:] = 0
        r_o = o.assemble(mode, True, nu)
        assert abs(r_d - r_o) <= 1e-11 * r_o, (r_d, r_o)
        assert np.abs(res_d - o.vec(3)).max() <= 1e-11 * np.abs(o.vec(3)).max()
        y_o = o.spmv(N.BLOCK_J, x)
        assert np.abs(y_d - y_o).max() <= 1e-12 * np.abs(y_o).max()
        rc_o, it_o, fr_o, _ = o.solve(N.STATIONARY, 1, prec, 1e-12, 3000)
        assert rc_o == 0 and rc_d == 0, (rc_o, rc_d)
        x_o = o.vec(2).copy()
        assert np.linalg.norm(x_d - x_o) <= 1e-8 * np.linalg.norm(x_o), np.linalg.norm(x_d - x_o) / np.linalg.norm(x_o)
        o.vec(0)[:] = sol + x_o
        r2_o = o.assemble(mode, False, nu)
        assert abs(r2_d - r2_o) <= 1e-6 * max(r2_o, 1e-6), (r2_d, r2_o)
        drag_o, lift_o = o.lift_drag(nu)
        assert abs(drag_d - drag_o) <= 1e-6 * abs(drag_o) and abs(lift_d - lift_o) <= 1e-6 * max(abs(lift_o), abs(drag_o))
        if defaults:
            assert view == 2, view   # F = K (x) I_2: the node view with ghost pairs and the node-layout ghost import
        print(f"MGPU_WORKER_OK world {world} elem {elem} prec {prec}{' defaults (view ' + str(view) + ')' if defaults else ''}: outer iterations gpu {it_d} / oracle {it_o}, "
              f"halo exchanges {stats[0]}, allreduces {stats[1]}")
    dist.barrier()
    dev.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)   # no destructors: a rank that failed must not wait for its peers inside ncclCommDestroy
