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


@pytest.mark.parametrize("name", sorted(S.MHD_CASES))
def test_mhd_smr_oracle_matches_golden(name):
    """MHD on refined meshes (oracle/smr_fc_oracle.c: RestrictFC, face-field exchange across levels, divergence-
    preserving ProlongateFC, fine/coarse EMF correction) against fixtures made from the unmodified reference."""
    d = G.load(name)
    mesh, p = S.mesh_of(str(d["deck"]))
    assert np.array_equal(mesh.lloc, d["lloc"]) and len(np.unique(mesh.lloc[:, 3])) >= 2
    active_only = bool(d["active_only"])
    b0 = (d["b1f_c0"], d["b2f_c0"], d["b3f_c0"])
    A, t, dt, n, dt0 = S.oracle_run_mhd(str(d["deck"]), d["u0_c0"], b0, active_only=active_only)
    assert dt0 == float(d["dt_c0"]) and n == int(d["ncycle_final"]) and t == float(d["time_final"]) and dt == float(d["dt_final"])
    ng = int(p["mesh/nghost"])
    if active_only:
        act = tuple(slice(ng, -ng) if k > 1 else slice(None) for k in mesh.nblock[::-1])
        u, b = A.u0[(slice(None), slice(None)) + act], S.active_faces(A.b0, ng, mesh.nblock)
    else:
        u, b = A.u0, A.b0
    assert np.array_equal(u, d["u0_final"]), f"max|dU| = {np.abs(u - d['u0_final']).max()}"
    for x, k in zip(b, ("b1f_final", "b2f_final", "b3f_final")):
        assert np.array_equal(x, d[k]), k


@pytest.mark.parametrize("name", sorted(set(S.MHD_LIVE_CASES) - set(S.MHD_CASES)))
def test_mhd_smr_oracle_matches_live_reference(name):
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    r0, r1 = S.reference_mhd(name)
    A, t, dt, n, dt0 = S.oracle_run_mhd(r0.deck, r0.u0, (r0.b1f, r0.b2f, r0.b3f))
    assert dt0 == r0.dt and n == r1.ncycle and t == r1.time and dt == r1.dt
    assert np.array_equal(A.u0, r1.u0), f"max|dU| = {np.abs(A.u0 - r1.u0).max()}"        # ghost zones included
    for x, y in zip(A.b0, (r1.b1f, r1.b2f, r1.b3f)):
        assert np.array_equal(x, y)


def test_mhd_smr_with_passive_scalar_and_bcs_matches_live_reference():
    """Brio-Wu along x2 on a statically refined 2-D mesh with reflect / outflow / diode boundaries and one passive scalar
    (tests/decks/bw2d_smr_scalars.athinput): the multilevel stage with nvar = 6 -- scalar fluxes, flux correction, restriction
    and prolongation of the scalar -- 49 cycles, ghost zones included."""
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    r0, r1 = S.reference_mhd_deck(S.SCALAR_DECK, 6)
    assert len(np.unique(r0.lloc[:, 3])) == 2 and r0.u0.shape[1] == 6
    A, t, dt, n, dt0 = S.oracle_run_mhd(r0.deck, r0.u0, (r0.b1f, r0.b2f, r0.b3f))
    assert dt0 == r0.dt and n == r1.ncycle and t == r1.time and dt == r1.dt
    assert np.array_equal(A.u0, r1.u0), f"max|dU| = {np.abs(A.u0 - r1.u0).max()}"
    for x, y in zip(A.b0, (r1.b1f, r1.b2f, r1.b3f)):
        assert np.array_equal(x, y)
    assert r1.u0[:, 5].max() - r1.u0[:, 5].min() > 0.5          # the scalar is not trivial

