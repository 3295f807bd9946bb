"""Adaptive mesh refinement, data path (dynamic refinement, second part): oracle/amr_oracle.c
(DerefineCC/FC, CopyCC/FC, CopyForRefinementCC/FC, RefineCC/FC, RepairAMRFC) + the AMR driver sequence
(AdaptiveMeshRefinement: criteria -> tree update -> redistribute/refine -> ghost zones + primitives -> RepairAMRFC ->
ghost zones + primitives -> new time step) against UNMODIFIED reference runs that dump a restart file every cycle.
From the reference's state of cycle k the oracle must reproduce the reference's state of cycle k+1 BITWISE -- conserved
variables and face fields with ghost zones, block list, time and dt -- for every cycle of the run, AMR events included.
Live runs need oracle/_ref/athena_blast; the committed fixture holds the two event cycles of a small 2-D run."""
import numpy as np
import pytest

import amr_util as AU
import golden_io as G
import hydro_util as H
import oracle_lib as O

CASES = {
    "2d_2levels": (),
    "2d_3levels": ("mesh_refinement/num_levels=3", "time/nlim=36"),
    "3d": ("mesh/nx1=32", "mesh/nx2=32", "mesh/nx3=32", "meshblock/nx3=8", "time/nlim=14", "mesh_refinement/refinement_interval=2",
           "problem/inner_radius=0.2", "problem/outer_radius=0.2"),
    # refinement and derefinement alternate (interval 1): the 8-sibling restriction in 3-D
    "3d_deref": ("mesh/nx1=32", "mesh/nx2=32", "mesh/nx3=32", "meshblock/nx3=8", "time/nlim=12", "mesh_refinement/refinement_interval=1",
                 "problem/inner_radius=0.2", "problem/outer_radius=0.2", "amr_criterion0/value_max=0.3"),
    "2d_ppm4_rk3_ng4": ("mesh/nghost=4", "mhd/reconstruct=ppm4", "time/integrator=rk3", "time/nlim=24", "mesh/nx1=64", "mesh/nx2=64"),
}


def check_step(S, r, rn, k):
    A, t, dt = S.advance(r, k)
    assert np.array_equal(S.mesh.lloc, rn.lloc), f"block list after cycle {k}"
    assert t == rn.time and dt == rn.dt, f"time / dt after cycle {k}"
    assert np.array_equal(A.u0, rn.u0), f"cycle {k}: max|dU| = {np.abs(A.u0 - rn.u0).max()}"
    for x, y in zip(A.b0, (rn.b1f, rn.b2f, rn.b3f)):
        assert np.array_equal(x, y), f"cycle {k}: face field differs"


@pytest.mark.parametrize("name", sorted(CASES))
def test_amr_cycle_by_cycle_matches_live_reference(name):
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref/athena_blast not built")
    ov = CASES[name]
    dumps = AU.reference_dumps(ov)
    S = AU.Stepper(AU.params(ov))
    for k in range(1, len(dumps) - 1):
        check_step(S, dumps[k], dumps[k + 1], k)
    assert S.nadapt >= 1 and S.ncreated > 0
    if name in ("2d_3levels", "2d_2levels", "3d_deref"):
        assert S.ndeleted > 0


def test_amr_events_match_fixture():
    """Committed fixture (tests/gen_golden.py --amr-only): a 64^2 run with 8^2 blocks; block list + density of every
    cycle (to replay the mesh history) and the full states around one refinement and one derefinement event."""
    d = G.load("amr_blast2d_events")
    p = H.deck_params(str(d["deck"]))
    S = AU.Stepper(p)
    events = [int(k) for k in d["events"]]
    assert len(events) >= 2

    class R:
        pass

    def state(k):
        r = R()
        r.lloc, r.u0, r.b1f, r.b2f, r.b3f = d[f"lloc_{k}"], d[f"u0_{k}"], d[f"b1f_{k}"], d[f"b2f_{k}"], d[f"b3f_{k}"]
        r.time, r.dt = float(d[f"time_{k}"]), float(d[f"dt_{k}"])
        return r

    for k in range(1, int(d["ncycles"]) - 1):
        if k in events:
            check_step(S, state(k), state(k + 1), k)
        else:       # replay the mesh bookkeeping only
            cfg = AU.ocfg(S.mesh, p)
            w = np.ascontiguousarray(d[f"dens_{k}"][:, None])
            flag = AU.criteria_flags(S.L, cfg, p, w, S.tables()[0])
            S.mesh.adapt(k, flag)
        assert np.array_equal(S.mesh.lloc, d[f"lloc_{k + 1}"])
    assert S.ncreated > 0 and S.ndeleted > 0


@pytest.mark.parametrize("name,overrides", [("2d_plm_hllc_rk2", ()), ("2d_ppm4_hlle_rk3_3levels", ("hydro/reconstruct=ppm4", "hydro/rsolver=hlle",
                                            "time/integrator=rk3", "mesh/nghost=4", "mesh_refinement/num_levels=3", "mesh/nx1=64", "mesh/nx2=64"))])
def test_hydro_amr_cycle_by_cycle_matches_live_reference(name, overrides):
    """The hydro twin (tests/decks/blast2d_hydro_amr.athinput; one ghost-zone rebuild after the data movement, no RepairAMRFC)."""
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref/athena_blast not built")
    dumps = AU.reference_dumps(overrides, deck=AU.HYDRO_DECK, physics="hydro")
    S = AU.HydroStepper(AU.params(overrides, deck=AU.HYDRO_DECK))
    for k in range(1, len(dumps) - 1):
        r, rn = dumps[k], dumps[k + 1]
        u, t, dt = S.advance(r, k)
        assert np.array_equal(S.mesh.lloc, rn.lloc) and t == rn.time and dt == rn.dt, f"cycle {k}"
        assert np.array_equal(u, rn.u0), f"cycle {k}: max|dU| = {np.abs(u - rn.u0).max()}"
    assert S.nadapt >= 2 and S.ncreated > 0 and (S.ndeleted > 0 or name != "2d_plm_hllc_rk2")


_FOFC = ("mhd/fofc=true", "problem/prat=100000", "problem/pi_amb=0.0001", "problem/b_amb=3.0", "time/nlim=20", "mesh/nx1=64", "mesh/nx2=64",
         "mesh/nghost=4", "mesh_refinement/refinement_interval=2")


@pytest.mark.parametrize("name,overrides", [("plm_rk2", _FOFC), ("ppm4_rk3", _FOFC + ("mhd/reconstruct=ppm4", "time/integrator=rk3"))])
def test_fofc_on_a_refined_mesh_matches_live_reference(name, overrides):
    """First-order flux correction inside the Fluxes task of the MULTILEVEL stage (mhd_tasks.cpp:208-211, before SendFlux), on a
    strong magnetised blast where the correction fires, with an AMR event on the way: oracle == reference bitwise, cycle by cycle.
    (The engine's multilevel task list does not take fofc yet.)"""
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref/athena_blast not built")
    dumps = AU.reference_dumps(overrides)
    S = AU.Stepper(AU.params(overrides))
    O.lib().oracle_get_nfofc(1)
    for k in range(1, len(dumps) - 1):
        check_step(S, dumps[k], dumps[k + 1], k)
    assert O.lib().oracle_get_nfofc(1) > 100 and S.nadapt >= 1 and len(np.unique(S.mesh.lloc[:, 3])) == 2

