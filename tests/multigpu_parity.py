"""Run under torchrun on N GPUs (N = 2, 4, 8): the 16^3 blast on 8 MeshBlocks of 8^3 sharded
over N ranks (NCCL point-to-point halo exchange, all-reduce(min) of dt) must reproduce the
reference's single-rank restart dump bit for bit (decomposition independence).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_parity.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import golden_io as G  # noqa: E402
from athenak_b200.driver import Driver, assemble_mhd_tasks  # noqa: E402
from athenak_b200.mesh import Mesh  # noqa: E402
from athenak_b200.mhd import MHD  # noqa: E402


def orszag_tang_leg(rank, world, dev):
    """Orszag-Tang 16^3 on 8 blocks over the ranks, 3 cycles, C++ host driver (peer-memory U/B and EMF exchange
    when available) against oracle/mhd_oracle.c run on all 8 blocks by every rank.  Unlike the blast, these ICs
    come from a vector potential evaluated on both sides of the periodic boundary, so the two copies of a shared
    face differ by an ulp and the SendE/RecvE correction changes bits: a lost or misplaced EMF
    message fails this leg."""
    import hydro_util as H
    from athenak_b200.sim import Simulation
    deck = open(os.path.join(HERE, "decks", "ot3d.athinput")).read()
    case = H.Case(deck, overrides=[f"mesh/nx{d}=16" for d in (1, 2, 3)] + [f"meshblock/nx{d}=8" for d in (1, 2, 3)])
    ncyc = 3
    full = Mesh(case.nmesh, case.nblock, nranks=1)
    fu, f1, f2, f3 = (t.numpy() for t in full.pgen_orszag_tang(0, case.ng, case.gamma))
    A, t_or, dt_or, n_or, dt0 = H.oracle_run_mhd(case, fu, (f1, f2, f3), ncycles=ncyc)
    mesh = Mesh(case.nmesh, case.nblock, nranks=world)
    g0, n = int(mesh.gids_start[rank]), int(mesh.nmb_rank[rank])
    ok = True
    for fused in (True, False):
        sim = Simulation(mesh, rank, ng=case.ng, recon=case.recon, rsolver=case.rsolver, gamma=case.gamma,
                         integrator=case.integrator, cfl_number=case.cfl, tlim=case.tlim, fused=fused)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(sim.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        sim.comm_init(bytes(idt.cpu().numpy().tobytes()))
        sim.set_state(*mesh.pgen_orszag_tang(rank, case.ng, case.gamma))
        sim.initialize()
        ok &= sim.status()["dt"] == dt0
        sim.execute(ncyc)
        u, x1, x2, x3 = sim.get_state()
        s = slice(case.ng, -case.ng)
        ok &= np.array_equal(u[:, :, s, s, s], A.u0[g0:g0 + n, :, s, s, s])
        for x, y in zip((x1, x2, x3), A.b0):
            ok &= np.array_equal(x, y[g0:g0 + n])       # face fields: ghost zones included
        ok &= sim.status()["time"] == t_or and sim.status()["dt"] == dt_or
        del sim
    return bool(ok)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ng, gamma = 4, 1.6666667
    d = G.load("blast16_ppm4_hlld_rk3")
    mesh = Mesh((16, 16, 16), (8, 8, 8), nranks=world)
    g0, n = int(mesh.gids_start[rank]), int(mesh.nmb_rank[rank])
    mb_size, _, nghbr = mesh.rank_tables(rank)
    u0, b1, b2, b3 = mesh.pgen_blast(rank, ng, gamma)
    ok = True
    for fused in (True, False):
        e = MHD(n, ng, (8, 8, 8), mb_size, nghbr, recon="ppm4", rsolver="hlld", gamma=gamma, my_rank=rank,
                gids_start=g0, device=dev)
        e.load_state(u0.numpy(), b1.numpy(), b2.numpy(), b3.numpy())
        drv = Driver("rk3", 0.3, tlim=0.3)
        tl = assemble_mhd_tasks(e, fused=fused)
        drv.Initialize(e)
        ok &= drv.dt == float(d["dt_c0"])
        drv.Execute(e, tl, ncycles=4)
        u = e.u0.cpu().numpy()
        bb = [t.cpu().numpy() for t in e.b0]
        for m in range(n):
            lx, ly, lz = (int(v) * 8 for v in mesh.lloc[g0 + m, :3])
            ok &= np.array_equal(u[m, :, 4:-4, 4:-4, 4:-4], d["u0_c4"][0][:, lz:lz + 8, ly:ly + 8, lx:lx + 8])
            ok &= np.array_equal(bb[0][m, 4:-4, 4:-4, 4:-4], d["b1f_c4"][0][lz:lz + 8, ly:ly + 8, lx:lx + 9])
            ok &= np.array_equal(bb[1][m, 4:-4, 4:-4, 4:-4], d["b2f_c4"][0][lz:lz + 8, ly:ly + 9, lx:lx + 8])
            ok &= np.array_equal(bb[2][m, 4:-4, 4:-4, 4:-4], d["b3f_c4"][0][lz:lz + 9, ly:ly + 8, lx:lx + 8])
        ok &= drv.time == float(d["time_c4"])
    # the same run through the C++ host driver with its own NCCL communicator
    from athenak_b200.sim import Simulation
    sim = Simulation(mesh, rank, ng=ng, recon="ppm4", rsolver="hlld", gamma=gamma, integrator="rk3",
                     cfl_number=0.3, tlim=0.3, fused=True)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(sim.unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    sim.comm_init(bytes(idt.cpu().numpy().tobytes()))
    sim.set_state(u0, b1, b2, b3)
    sim.initialize()
    ok &= sim.status()["dt"] == float(d["dt_c0"])
    sim.execute(4)
    u, x1, x2, x3 = sim.get_state()
    for m in range(n):
        lx, ly, lz = (int(v) * 8 for v in mesh.lloc[g0 + m, :3])
        ok &= np.array_equal(u[m, :, 4:-4, 4:-4, 4:-4], d["u0_c4"][0][:, lz:lz + 8, ly:ly + 8, lx:lx + 8])
        ok &= np.array_equal(x1[m, 4:-4, 4:-4, 4:-4], d["b1f_c4"][0][lz:lz + 8, ly:ly + 8, lx:lx + 9])
        ok &= np.array_equal(x2[m, 4:-4, 4:-4, 4:-4], d["b2f_c4"][0][lz:lz + 8, ly:ly + 9, lx:lx + 8])
        ok &= np.array_equal(x3[m, 4:-4, 4:-4, 4:-4], d["b3f_c4"][0][lz:lz + 9, ly:ly + 8, lx:lx + 8])
    ok &= sim.status()["time"] == float(d["time_c4"])
    push_used = sim.status()["peer_push"]
    ok &= (sim.status()["peer_exchanges"] > 0) == push_used
    del sim
    ok &= orszag_tang_leg(rank, world, dev)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_PARITY", "OK" if int(t.item()) == 1 else "FAIL", "world", world, "peer_push", push_used)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
