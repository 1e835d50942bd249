"""The CPU oracle: pinned against the only observables the reference offers and
against hand-checkable properties of the task bodies."""
import numpy as np
import pytest

from mpas_regent_b200 import _abi, dynamics, init_jw
from tests.util import build_pair

L = 8


@pytest.fixture(scope="module")
def pair(grid642):
    st, ora, _ = build_pair(grid642, L, _abi.INDEX_CORRECTED, gpu=False)
    ora.atm_compute_solve_diagnostics(False, -1)
    yield st, ora
    ora.close()


def test_output_txt_observable_u_never_changes(grid2562):
    """output.txt prints er[{0,0}].u after every RK stage of 10 steps and it never changes: nothing on
    the shipped path writes u (acoustic u-update commented out :1585-1613, recover not called,
    rk_timestep.rg:460).  With the literal driver dt = step index (main.rg:66)."""
    st, ora, _ = build_pair(grid2562, 5, _abi.INDEX_LITERAL, m5=False, gpu=False, rkarg=_abi.RKARG_SUBSTEP_TRUNC)
    u0 = ora.download_field("u").copy()
    ora.atm_compute_solve_diagnostics(False, -1)
    for j in range(3):
        ora.atm_srk3(float(j))
    assert np.array_equal(ora.download_field("u"), u0)
    assert np.array_equal(ora.download_field("theta_m"), st.f["theta_m"])
    ora.close()


def test_setup_and_finish_are_copies(pair):
    st, ora = pair
    ora.atm_rk_integration_setup()
    for a, b in (("ru_save", "ru"), ("u_2", "u"), ("rw_save", "rw"), ("w_2", "w"), ("rho_zz_old_split", "rho_zz")):
        x, y = ora.download_field(a), ora.download_field(b)
        assert np.array_equal(x[:, :L], y[:, :L])
        assert not x[:, L].any()            # level nVertLevels is outside every hot-path loop


def test_moist_coefficients(pair):
    st, ora = pair
    ora.atm_compute_moist_coefficients()
    cqw = ora.download_field("cqw")
    assert np.all(cqw[:, 1:L] == 1.0) and np.all(cqw[:, 0] == 0.0)
    assert not ora.download_field("qtot").any()


def test_vert_imp_coefs_uses_previous_gamma(pair):
    """alpha_tri(k) reads gamma_tri(k-1) from the PREVIOUS call (:580-591), so two identical calls differ."""
    st, ora = pair
    ora.atm_compute_moist_coefficients()
    ora.upload_field("gamma_tri", np.zeros(ora.field_shape("gamma_tri")))
    ora.atm_compute_vert_imp_coefs(240.0)
    a1, b, g1 = ora.download_field("alpha_tri"), ora.download_field("b_tri"), ora.download_field("gamma_tri")
    assert np.allclose(a1[:, 1:L], 1.0 / b[:, 1:L], rtol=0, atol=0)        # gamma was all zero
    ora.atm_compute_vert_imp_coefs(240.0)
    a2 = ora.download_field("alpha_tri")
    a_tri = ora.download_field("a_tri")
    expect = 1.0 / (b[:, 2:L] - a_tri[:, 2:L] * g1[:, 1:L - 1])
    assert np.array_equal(a2[:, 2:L], expect)
    cofrz = ora.download_field("cofrz")
    assert np.array_equal(cofrz[:L], 0.5 * 240.0 * 1.1 * st.vert["rdzw"][:L])


def test_divergence_is_s_plus_u(pair, grid642):
    """Q6: divergence accumulates s + u, not s * u (:375)."""
    st, ora = pair
    from mpas_regent_b200.mesh import resolve_ids
    ora.atm_compute_solve_diagnostics(False, 0)
    div = ora.download_field("divergence")
    u = ora.download_field("u")
    S = st.static
    eoc = resolve_ids(S["edgesOnCell"], grid642.nEdges, _abi.INDEX_CORRECTED)
    c, k = 17, 3
    acc = 0.0
    for i in range(S["nEdgesOnCell"][c]):
        e = eoc[c, i]
        acc += S["edgesOnCellSign"][c, i] * S["dvEdge"][e] + u[e, k]
    assert div[c, k] == acc * S["invAreaCell"][c]


def test_v_only_on_stage_2_or_init(pair):
    st, ora = pair
    ora.upload_field("v", np.full(ora.field_shape("v"), 7.0))
    ora.atm_compute_solve_diagnostics(False, 0)
    assert np.all(ora.download_field("v") == 7.0)
    ora.atm_compute_solve_diagnostics(False, 2)
    assert not np.all(ora.download_field("v")[:, :L] == 7.0)


def test_srk3_call_sequence_counts(pair):
    """2 / 2 / 3 acoustic iterations per stage (rk_timestep.rg:450, output.txt)."""
    st, ora = pair
    calls = []
    ora.atm_srk3_by_tasks(600.0, hook=calls.append)
    assert calls.count("advance_acoustic_step") == 7 and calls.count("divergence_damping_3d") == 7
    assert calls.count("compute_dyn_tend") == 3 and calls.count("compute_vert_imp_coefs") == 2
    assert calls.count("compute_solve_diagnostics") == 3 and calls[-1] == "rk_dynamics_substep_finish"


def test_by_tasks_equals_driver(grid642):
    outs = []
    for mode in (0, 1):
        st, ora, _ = build_pair(grid642, L, _abi.INDEX_CORRECTED, gpu=False)
        ora.atm_compute_solve_diagnostics(False, -1)
        (ora.atm_srk3 if mode else ora.atm_srk3_by_tasks)(600.0)
        outs.append(ora.download_all()); ora.close()
    for n in outs[0]:
        assert np.array_equal(outs[0][n], outs[1][n], equal_nan=True), n


def test_threads_do_not_change_results(grid642):
    outs = []
    for t in (1, 4):
        st, ora, _ = build_pair(grid642, L, _abi.INDEX_CORRECTED, gpu=False)
        ora.set_threads(t)
        ora.atm_compute_solve_diagnostics(False, -1)
        ora.atm_srk3(600.0)
        outs.append(ora.download_all()); ora.close()
    for n in outs[0]:
        assert np.array_equal(outs[0][n], outs[1][n], equal_nan=True), n


def test_recover_writes_the_garbage_cell(pair):
    st, ora = pair
    ora.atm_recover_large_step_variables(1, 0, 600.0)
    assert np.all(ora.download_pad("rho_zz")[:L] == 1.0)                  # :1792-1794
