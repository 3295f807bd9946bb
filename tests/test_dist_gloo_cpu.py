"""world_size-2 gloo test (CPU): the N>1 exchange path -- per-rank plan, one message per peer per
exchange through torch.distributed point-to-point, min-reduction of dt -- against the
single-rank multi-block oracle."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import halo_emul as H
import oracle_lib as O
from athenak_b200 import lib as _lib
from athenak_b200.mesh import Mesh

NMESH, NBLOCK, NG = (32, 16, 16), (8, 8, 8), 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M = Mesh(NMESH, NBLOCK, nranks=world)
        g0, n = int(M.gids_start[rank]), int(M.nmb_rank[rank])
        rng = np.random.default_rng(5)                     # same global field on every rank
        n1 = NBLOCK[0] + 2 * NG
        U = rng.standard_normal((M.nmb_total, 5, n1, n1, n1))
        B = [rng.standard_normal((M.nmb_total, n1, n1, n1 + 1)), rng.standard_normal((M.nmb_total, n1, n1 + 1, n1)),
             rng.standard_normal((M.nmb_total, n1 + 1, n1, n1))]
        u = {-1: U[g0:g0 + n].copy()}
        b = {q: B[q][g0:g0 + n].copy() for q in range(3)}
        cfg = _lib.MhdCfg(n, NG, *NBLOCK, 2, 2, 1.6666667, 0, 0, 0, 0, 1e38)
        plan = H.make_plan(cfg, M.rank_tables(rank)[2], rank, g0)
        for arrays, nvar, kp, ku, kl in ((u, 5, 2, 3, 0), (b, 1, 4, 5, 1)):
            ops, recv = [], []
            for ip, (peer, ncc, nfc) in enumerate(H.peers(plan)):
                cnt = 5 * ncc if nvar == 5 else nfc
                sb = np.zeros(cnt)
                H.pack(H.descs(plan, kp, ip), arrays, nvar, sb)
                st, rt = torch.from_numpy(sb), torch.zeros(cnt, dtype=torch.float64)
                recv.append(rt)
                ops += [dist.P2POp(dist.isend, st, peer), dist.P2POp(dist.irecv, rt, peer)]
            reqs = dist.batch_isend_irecv(ops) if ops else []
            H.local(H.descs(plan, kl), arrays)
            for rq in reqs:
                rq.wait()
            for ip in range(len(recv)):
                H.unpack(H.descs(plan, ku, ip), arrays, nvar, recv[ip].numpy())
        # Mesh::NewTimeStep rank-min (mesh.cpp:634-637)
        t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        np.savez(os.path.join(out, f"r{rank}.npz"), u=u[-1], b1=b[0], b2=b[1], b3=b[2], g0=g0, n=n, dtmin=t.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_matches_oracle(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    lloc, ngh, mbs, _ = O.uniform_mesh(NMESH, NBLOCK)
    cfg = O.make_cfg(len(lloc), NG, NBLOCK)
    A = O.Arrays(cfg)
    rng = np.random.default_rng(5)
    A.u0[:] = rng.standard_normal(A.u0.shape)
    for a in A.b0:
        a[:] = rng.standard_normal(a.shape)
    L = O.lib()
    L.oracle_halo_cc(C.byref(cfg), O.p(A.u0), 5, O.ip(ngh))
    L.oracle_halo_fc(C.byref(cfg), *[O.p(a) for a in A.b0], O.ip(ngh))
    seen = 0
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        g0, n = int(d["g0"]), int(d["n"])
        seen += n
        assert np.array_equal(d["u"], A.u0[g0:g0 + n])
        for k, a in zip(("b1", "b2", "b3"), A.b0):
            assert np.array_equal(d[k], a[g0:g0 + n])
        assert float(d["dtmin"][0]) == 0.1
    assert seen == len(lloc)
