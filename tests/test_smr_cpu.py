"""CPU tests of the hydro SMR restatement (oracle/smr_oracle.c: restriction, multi-level ghost exchange, coarse
buffer fill + slope-limited prolongation, flux correction) against the unmodified reference: committed fixtures
(tests/golden/hydro_smr*.npz, tests/gen_golden.py smr()) and live runs when oracle/_ref is present.  Gate:
u0 INCLUDING the prolongated ghost zones == the reference's dump, bitwise; dt and cycle count equal."""
import numpy as np
import pytest

import golden_io as G
import oracle_lib as O
import smr_util as S


@pytest.mark.parametrize("name", sorted(S.CASES))
def test_smr_oracle_matches_golden(name):
    d = G.load(name)
    mesh, _ = S.mesh_of(str(d["deck"]))
    assert np.array_equal(mesh.lloc, d["lloc"]) and len(np.unique(mesh.lloc[:, 3])) >= 2
    u_init, u_final, act = S.with_ghosts(d)
    # a ghost-inclusive dump's own ghost zones are the result of Driver::Initialize: restriction, exchange, prolongation
    u, w, t, dt, n, dt0 = S.oracle_run(str(d["deck"]), u_init, ncycles=0)
    assert dt0 == float(d["dt_c0"]) and np.array_equal(u[:, :, act[0], act[1], act[2]], d["u0_c0"])
    u, w, t, dt, n, dt0 = S.oracle_run(str(d["deck"]), u_init)
    assert n == int(d["ncycle_final"]) and t == float(d["time_final"]) and dt == float(d["dt_final"])
    u = u[:, :, act[0], act[1], act[2]]
    assert np.array_equal(u, u_final), f"max|dU| = {np.abs(u - u_final).max()}"


@pytest.mark.parametrize("name", sorted(set(S.LIVE_CASES) - set(S.CASES)))
def test_smr_oracle_matches_live_reference(name):
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    r0, r1 = S.reference(name)
    u, w, t, dt, n, dt0 = S.oracle_run(r0.deck, r0.u0)
    assert dt0 == r0.dt and n == r1.ncycle and t == r1.time and dt == r1.dt
    assert np.array_equal(u, r1.u0), f"max|dU| = {np.abs(u - r1.u0).max()}"
