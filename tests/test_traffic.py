"""The per-kernel field lists (mpas_regent_b200/traffic.py) regenerate SURVEY.md 8(d)'s table; a
kernel's declared algorithmic bytes may not drop below the contract figures."""
import re

from mpas_regent_b200 import _abi, traffic as T


def test_every_declared_field_exists():
    for k, (r, w) in T.K.items():
        for n in r + w:
            assert n.startswith("scr") or n in _abi.FIELD_ID, (k, n)


def test_task_units_not_below_survey():
    for task, want in T.SURVEY_TASK_UNITS.items():
        assert T.task_units(task) >= want, (task, T.task_units(task), want)
    for task in ("advance_acoustic_step:s0", "advance_acoustic_step"):        # the fused single-kernel form moves exactly the contract
        assert T.task_units(task + ":fused") == T.SURVEY_TASK_UNITS[task]
    # tasks whose kernels move exactly the contract figure
    for task in ("rk_integration_setup", "compute_moist_coefficients", "compute_vert_imp_coefs", "set_smlstep_pert_variables",
                 "divergence_damping_3d", "rk_dynamics_substep_finish"):
        assert T.task_units(task) == T.SURVEY_TASK_UNITS[task], task


def test_step_units():
    assert T.step_units(True, scratch=False) >= T.SURVEY_STEP_UNITS_CANONICAL
    assert T.step_units(False, scratch=False) >= T.SURVEY_STEP_UNITS_LITERAL
    assert len(T.step_launches(True)) == 58 and len(T.step_launches(False)) == 52
    # corrected physics: + 7 edge updates + 3 x 4 recover launches; more bytes than the literal step, never fewer
    assert len(T.step_launches(True, corrected_physics=True)) == 58 + 7 + 12
    assert T.step_units(True, scratch=False, corrected_physics=True) > T.step_units(True, scratch=False)


def test_every_kernel_in_the_library_is_declared():
    src = open(_abi.REPO_ROOT + "/mpas_regent_b200/csrc/mpas_b200.cu").read()
    names = set(re.findall(r"LAUNCH(?:_STAGED)?\((k_\w+(?:<\w+>)?)", src))
    assert names, "no launches found"
    for n in names:
        if re.search(r"_v\d$", n):      # experimental variants behind mpasb200_debug_divdamp (MPASB200_LAB builds only)
            continue
        assert T.lookup(n) is not None, n
    for n in ("k_dt_edge_s<10>", "k_acoustic_gather_s<6>", "k_dt_theta_flux_s<10>"):
        assert T.units(n) == T.units(T.canon(n)) and T.canon(n) != n
