"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, see oracle/build_ref.sh).

Run in the build container:  python tests/gen_golden.py
Fixtures hold only the ACTIVE zones (ghosts are a periodic wrap of them and are rebuilt by
tests/golden_io.py), so each state is a few hundred kB.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as O  # noqa: E402
from athenak_b200.rst import read_rst  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def active(r):
    ng = r.mb_indcs["ng"]
    s = slice(ng, -ng)
    return dict(u0=r.u0[:, :, s, s, s], b1f=r.b1f[:, s, s, ng:-ng], b2f=r.b2f[:, s, ng:-ng, s],
                b3f=r.b3f[:, ng:-ng, s, s])


BC_NAMES = ("ix1_bc", "ox1_bc", "ix2_bc", "ox2_bc", "ix3_bc", "ox3_bc")


def blast(n, cycles, recon="ppm4", rsolver="hlld", integrator="rk3", name=None, bcs=None, extra=()):
    tmp = tempfile.mkdtemp(prefix="akgold_")
    try:
        ov = [f"mesh/nx1={n}", f"mesh/nx2={n}", f"mesh/nx3={n}", f"meshblock/nx1={n}", f"meshblock/nx2={n}",
              f"meshblock/nx3={n}", f"time/nlim={max(cycles)}", f"mhd/reconstruct={recon}", f"mhd/rsolver={rsolver}",
              f"time/integrator={integrator}", *extra]
        if bcs:
            ov += [f"mesh/{k}={v}" for k, v in zip(BC_NAMES, bcs)]
        O.run_reference("athena_blast", os.path.join(HERE, "decks", "blast3d.athinput"), tmp, ov)
        out = {}
        for c in [0] + list(cycles):
            r = read_rst(os.path.join(tmp, "rst", "Blast.%05d.rst" % c))
            assert r.ncycle == c
            for k, v in active(r).items():
                out[f"{k}_c{c}"] = v
            out[f"time_c{c}"] = r.time
            out[f"dt_c{c}"] = r.dt
            if bcs and c == max(cycles):   # ghost-inclusive final state: the BC fill itself is checked
                out.update({f"u0g_c{c}": r.u0, f"b1fg_c{c}": r.b1f, f"b2fg_c{c}": r.b2f, f"b3fg_c{c}": r.b3f})
        out["meta"] = np.array([n, 4, max(cycles)])
        np.savez_compressed(os.path.join(HERE, "golden", name or f"blast{n}_{recon}_{rsolver}_{integrator}.npz"), **out)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def hydro(deck, base, name, overrides=()):
    """Hydro fixtures: ghost-inclusive u0 of the cycle-0 and final restart dumps + the embedded deck."""
    tmp = tempfile.mkdtemp(prefix="akgold_")
    try:
        O.run_reference("athena_builtin", os.path.join(HERE, "decks", deck), tmp, list(overrides))
        r0 = read_rst(os.path.join(tmp, "rst", base + ".00000.rst"), physics="hydro")
        r1 = read_rst(os.path.join(tmp, "rst", base + ".00001.rst"), physics="hydro")
        np.savez_compressed(os.path.join(HERE, "golden", name + ".npz"), deck=np.array(r0.deck), u0_c0=r0.u0, u0_final=r1.u0,
                            dt_c0=r0.dt, time_final=r1.time, dt_final=r1.dt, ncycle_final=r1.ncycle, lloc=r0.lloc)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def mhd_lowd(deck, base, name, overrides=()):
    """1-D / 2-D MHD fixtures: ghost-inclusive u0 and face fields of the cycle-0 and final dumps."""
    tmp = tempfile.mkdtemp(prefix="akgold_")
    try:
        O.run_reference("athena_builtin", os.path.join(HERE, "decks", deck), tmp, list(overrides))
        r0 = read_rst(os.path.join(tmp, "rst", base + ".00000.rst"))
        r1 = read_rst(os.path.join(tmp, "rst", base + ".00001.rst"))
        np.savez_compressed(os.path.join(HERE, "golden", name + ".npz"), deck=np.array(r0.deck), lloc=r0.lloc,
                            u0_c0=r0.u0, b1f_c0=r0.b1f, b2f_c0=r0.b2f, b3f_c0=r0.b3f,
                            u0_final=r1.u0, b1f_final=r1.b1f, b2f_final=r1.b2f, b3f_final=r1.b3f,
                            dt_c0=r0.dt, time_final=r1.time, dt_final=r1.dt, ncycle_final=r1.ncycle)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def iso():
    """Isothermal-EOS fixtures (tests/iso_util.py CASES): ghost-inclusive initial and final dumps."""
    import iso_util as I
    for name, c in I.CASES.items():
        r0, r1 = I.reference(name)
        out = dict(deck=np.array(r0.deck), lloc=r0.lloc, u0_c0=r0.u0, u0_final=r1.u0, time_final=r1.time,
                   dt_final=r1.dt, ncycle_final=r1.ncycle, nfofc=r1.nfofc)
        if c[3] == "mhd":
            out.update(b1f_c0=r0.b1f, b2f_c0=r0.b2f, b3f_c0=r0.b3f, b1f_final=r1.b1f, b2f_final=r1.b2f, b3f_final=r1.b3f)
        np.savez_compressed(os.path.join(HERE, "golden", name + ".npz"), **out)


SMR_TREES = {   # name -> overrides on tests/decks/lwave3d_hydro_smr.athinput (time/nlim=0: only the mesh is built)
    "smr3d_periodic": (),
    "smr3d_outflow_l3": ("mesh/ix1_bc=outflow", "mesh/ox1_bc=outflow", "mesh/ix3_bc=reflect", "mesh/ox3_bc=reflect",
                         "refined_region1/level=3", "refined_region1/x1min=2.9", "refined_region1/x1max=3.0",
                         "refined_region1/x3min=0.0", "refined_region1/x3max=0.05"),
    "smr2d_mixed": ("mesh/nx3=1", "meshblock/nx3=1", "mesh/nx2=32", "meshblock/nx2=8", "mesh/ix2_bc=outflow", "mesh/ox2_bc=reflect",
                    "refined_region2/level=3", "refined_region2/x1min=2.6", "refined_region2/x1max=3.0", "refined_region2/x2min=1.3",
                    "refined_region2/x2max=1.5"),
    "smr1d": ("mesh/nx2=1", "meshblock/nx2=1", "mesh/nx3=1", "meshblock/nx3=1", "mesh/nx1=128", "meshblock/nx1=8",
              "refined_region2/level=3"),
}


def smr_trees():
    """Z-ordered LogicalLocation lists (lx1, lx2, lx3, level per gid) of the unmodified reference for static-mesh-
    refinement decks -> tests/golden/smr_trees.npz (a few kB).  Read from the header of its cycle-0 restart dump."""
    import struct
    out = {}
    for name, ov in SMR_TREES.items():
        tmp = tempfile.mkdtemp(prefix="akgold_")
        try:
            O.run_reference("athena_builtin", os.path.join(HERE, "decks", "lwave3d_hydro_smr.athinput"), tmp, list(ov) + ["time/nlim=0"])
            raw = open(os.path.join(tmp, "rst", "LinWave.00000.rst"), "rb").read()
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        end = raw.index(b"<par_end>\n") + 10
        nmb, root = struct.unpack_from("<ii", raw, end)
        off = end + 8 + 72 + 76 + 76 + 16 + 4          # restart.cpp:216-294 header
        out[name + "_lloc"] = np.frombuffer(raw, "<i4", 4 * nmb, off).reshape(nmb, 4).copy()
        out[name + "_root"] = root
        out[name + "_deck"] = np.array(raw[:end].decode())
    np.savez_compressed(os.path.join(HERE, "golden", "smr_trees.npz"), **out)


def smr():
    """Hydro SMR fixtures (tests/smr_util.py CASES): ghost-inclusive initial and final dumps on refined meshes."""
    import smr_util as S
    for name in S.CASES:
        r0, r1 = S.reference(name)
        ng = r0.mb_indcs["ng"]
        if r0.u0.shape[2] > 1:      # 3-D: active zones only (ghost zones are rebuilt by Driver::Initialize)
            s = slice(ng, -ng)
            r0.u0, r1.u0 = r0.u0[:, :, s, s, s], r1.u0[:, :, s, s, s]
        np.savez_compressed(os.path.join(HERE, "golden", name + ".npz"), deck=np.array(r0.deck), lloc=r0.lloc, u0_c0=r0.u0,
                            u0_final=r1.u0, dt_c0=r0.dt, ng=ng, time_final=r1.time, dt_final=r1.dt, ncycle_final=r1.ncycle)


def smr_mhd():
    """MHD SMR fixtures (tests/smr_util.py MHD_CASES); 3-D ones keep the active zones / faces only."""
    import smr_util as S
    for name in S.MHD_CASES:
        r0, r1 = S.reference_mhd(name)
        ng = r0.mb_indcs["ng"]
        nb = (r0.mb_indcs["nx1"], r0.mb_indcs["nx2"], r0.mb_indcs["nx3"])
        out = dict(deck=np.array(r0.deck), lloc=r0.lloc, dt_c0=r0.dt, time_final=r1.time, dt_final=r1.dt, ncycle_final=r1.ncycle)
        active_only = nb[2] > 1
        for tag, r in (("c0", r0), ("final", r1)):
            u, b = r.u0, (r.b1f, r.b2f, r.b3f)
            if active_only:
                s = slice(ng, -ng)
                u, b = u[:, :, s, s, s], S.active_faces(b, ng, nb)
            out["u0_" + tag] = u
            for k, x in zip(("b1f_", "b2f_", "b3f_"), b):
                out[k + tag] = x
        out["active_only"] = active_only
        np.savez_compressed(os.path.join(HERE, "golden", name + ".npz"), **out)


def amr():
    """AMR host-map fixture (tests/test_amr_cpu.py): block list + ghost-inclusive density of every cycle of the
    unmodified reference on tests/decks/blast2d_amr.athinput -> tests/golden/amr_blast2d_llocs.npz."""
    import test_amr_cpu as T
    seq = T.reference_sequence()
    out = dict(deck=np.array(open(T.DECK).read()), ncycles=len(seq))
    for k, (l, d) in enumerate(seq):
        out[f"lloc_{k}"] = l
        out[f"dens_{k}"] = d
    np.savez_compressed(os.path.join(HERE, "golden", "amr_blast2d_llocs.npz"), **out)


AMR_EVENTS_OVERRIDES = ("mesh/nx1=64", "mesh/nx2=64", "time/nlim=16", "mesh_refinement/refinement_interval=2",
                        "problem/inner_radius=0.2", "problem/outer_radius=0.2")


def amr_events():
    """AMR data-path fixture (tests/test_amr_data_cpu.py): 64^2 blast, 8^2 blocks; block list + density of every cycle,
    full ghost-inclusive states of the cycles around the first refinement and the first derefinement
    -> tests/golden/amr_blast2d_events.npz."""
    import amr_util as AU
    dumps = AU.reference_dumps(AMR_EVENTS_OVERRIDES)
    p = AU.params(AMR_EVENTS_OVERRIDES)
    deck = "".join(f"<{b}>\n" + "".join(f"{k.split('/', 1)[1]} = {v}\n" for k, v in p.items() if k.split("/")[0] == b) + "\n"
                   for b in dict.fromkeys(k.split("/")[0] for k in p))
    out = dict(deck=np.array(deck), ncycles=len(dumps))
    events, seen = [], set()
    for k in range(1, len(dumps) - 1):
        kind = np.sign(dumps[k + 1].nmb - dumps[k].nmb)
        if kind != 0 and kind not in seen:
            seen.add(kind)
            events.append(k)
    for k, r in enumerate(dumps):
        out[f"lloc_{k}"] = r.lloc
        out[f"dens_{k}"] = r.u0[:, 0]
        if k in events or (k - 1) in events:
            out[f"u0_{k}"], out[f"b1f_{k}"], out[f"b2f_{k}"], out[f"b3f_{k}"] = r.u0, r.b1f, r.b2f, r.b3f
            out[f"time_{k}"], out[f"dt_{k}"] = r.time, r.dt
    out["events"] = np.array(events)
    np.savez_compressed(os.path.join(HERE, "golden", "amr_blast2d_events.npz"), **out)


if __name__ == "__main__":
    if "--amr-only" in sys.argv:
        amr()
        amr_events()
        sys.exit(0)
    if "--smr-only" in sys.argv:
        smr_trees()
        smr()
        smr_mhd()
        sys.exit(0)
    if "--smr-mhd-only" in sys.argv:
        smr_mhd()
        sys.exit(0)
    iso()
    smr_trees()
    smr()
    smr_mhd()
    if "--iso-only" in sys.argv:
        sys.exit(0)
    mhd_lowd("bw1d.athinput", "BrioWu", "mhd_bw1d_ppm4_hlld_rk3",
             ("mhd/rsolver=hlld", "mhd/reconstruct=ppm4", "mesh/nghost=3", "time/integrator=rk3", "mesh/nx1=128", "meshblock/nx1=32"))
    mhd_lowd("bw2d.athinput", "BrioWu", "mhd_bw2d_ppm4_hlld_rk3", ("time/tlim=0.03",))
    mhd_lowd("ot2d.athinput", "OrszagTang", "mhd_ot2d_plm_hlle_rk2",
             ("mesh/nx1=32", "mesh/nx2=32", "meshblock/nx1=16", "meshblock/nx2=16", "mhd/reconstruct=plm", "mhd/rsolver=hlle",
              "time/integrator=rk2", "mesh/nghost=2", "time/tlim=0.1"))
    hydro("sod1d.athinput", "Sod", "hydro_sod1d_plm_llf_rk2")
    hydro("sod1d.athinput", "Sod", "hydro_sod1d_wenoz_roe_rk3", ("hydro/rsolver=roe", "hydro/reconstruct=wenoz", "mesh/nghost=3",
                                                                  "time/integrator=rk3", "time/cfl_number=0.4", "mesh/nx1=128", "meshblock/nx1=32"))
    hydro("sod2d.athinput", "Sod", "hydro_sod2d_ppm4_hllc_rk3")
    hydro("lwave3d_hydro.athinput", "LinWave", "hydro_lwave3d_ppmx_hlle_rk2",
          ("mesh/nx1=16", "mesh/nx2=8", "mesh/nx3=8", "meshblock/nx1=8", "meshblock/nx2=8", "meshblock/nx3=8",
           "hydro/reconstruct=ppmx", "hydro/rsolver=hlle", "time/integrator=rk2", "time/tlim=0.25"))
    if "--hydro-only" in sys.argv or "--new-only" in sys.argv:
        sys.exit(0)
    blast(16, (1, 4))
    blast(16, (2,), recon="plm", rsolver="hlle", integrator="rk2")
    blast(16, (2,), recon="wenoz", rsolver="llf", integrator="rk3")
    blast(16, (2,), recon="ppmx", rsolver="hlld", integrator="rk1")
    blast(16, (3,), recon="teno", rsolver="hlle", integrator="rk2")
    # physical BCs (task 26): a blast wide enough to cross every mesh edge
    blast(16, (3,), name="blast16_bcs_mixed.npz", bcs=("reflect", "outflow", "diode", "reflect", "outflow", "diode"),
          extra=("problem/inner_radius=0.42", "problem/outer_radius=0.55"))
    print("golden fixtures written")
