"""Algorithmic HBM traffic of every kernel, derived from its declared field lists.

Convention: each distinct 3-D field counts once per direction per kernel at
8 B x entity multiplicity (cell 1, edge 3, vertex 2) per cell-level; gathers are assumed
perfectly reused; static per-entity data is excluded.  ``units(kernel)`` x 8 B x nCells x
nVertLevels is the bytes one launch must move; bench.py divides it by the kernel's measured
duration for ``roofline.achieved``.  tests/test_traffic.py checks the per-task sums against
the survey's contract table so a kernel cannot silently under-declare.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

from ._abi import CELL, EDGE, FIELD_ENTITY, FIELD_SLOTS, VERTEX

MULT = {CELL: 1, EDGE: 3, VERTEX: 2}

# slots of zb_cell / zb3_cell actually touched per cell (nEdgesOnCell ~ 6)
USED_SLOTS = 6

# kernel -> (reads, writes).  "scr" = library-private scratch (not algorithmic, counted separately).
K: Dict[str, Tuple[List[str], List[str]]] = {
    "k_setup_cell": (["rw", "rtheta_p", "rho_p", "w", "theta_m", "rho_zz"],
                     ["rw_save", "rtheta_p_save", "rho_p_save", "w_2", "theta_m_2", "rho_zz_2", "rho_zz_old_split"]),
    "k_setup_edge": (["ru", "u"], ["ru_save", "u_2"]),
    "k_moist": ([], ["qtot", "cqw"]),
    "k_vert_imp": (["zz", "cqw", "exner", "theta_m", "qtot", "rho_base", "rtheta_base", "rtheta_p", "exner_base", "gamma_tri"],
                   ["cofwr", "cofwz", "coftz", "cofwt", "a_tri", "b_tri", "c_tri", "alpha_tri", "gamma_tri"]),
    "k_diag_vertex": (["u"], ["vorticity", "pv_vertex"]),
    "k_diag_cell": (["u"], ["divergence", "ke"]),
    "k_diag_edge<false>": (["h", "u", "pv_vertex"], ["h_edge", "ke_edge", "pv_edge"]),
    "k_diag_edge<true>": (["h", "u", "pv_vertex"], ["h_edge", "ke_edge", "pv_edge", "v"]),
    "k_diag_ke_vertex": (["ke_edge"], ["ke_vertex"]),
    "k_diag_ke_holl": (["ke", "ke_vertex"], ["ke"]),
    "k_dt_cell0<true>": (["u", "v", "ru", "rw", "tend_rho_physics", "qtot", "rho_base", "rho_p_save"],
                         ["kdiff", "h_divergence", "tend_rho", "dpdz"]),
    "k_dt_cell0<false>": (["ru"], ["h_divergence"]),
    "k_dt_edge_delsq": (["divergence", "vorticity"], ["delsq_u"]),
    "k_dt_vertex_delsq": (["delsq_u"], ["delsq_vorticity"]),
    "k_dt_cell_delsq": (["delsq_u"], ["delsq_divergence"]),
    "k_dt_edge_euler": (["rho_edge", "cqu", "pressure_p", "zz", "dpdz", "zxu", "divergence", "vorticity", "kdiff", "delsq_divergence",
                         "delsq_vorticity"], ["tend_u_euler"]),
    "k_dt_edge": (["u", "rw", "pv_edge", "rho_edge", "ke", "h_divergence", "w", "tend_ru_physics", "tend_u_euler"],
                  ["wduz", "q", "tend_u"]),
    "k_dt_theta_flux": (["ru", "theta_m"], ["scr_e"]),
    "k_dt_cellA": (["ru", "rho_zz", "uReconstructZonal", "uReconstructMeridional", "theta_m", "kdiff", "rho_edge"],
                   ["w", "ru_edge_w", "delsq_theta", "tend_theta_euler"]),
    "k_dt_cellB": (["w", "kdiff", "rho_edge"], ["delsq_w", "tend_w_euler"]),
    "k_dt_cellC<true>": (["w", "tend_w_euler", "delsq_w", "rw", "pressure_p", "dpdz", "cqw", "ru", "scr_e", "theta_m", "theta_m_save",
                          "rw_save", "rho_zz", "rt_diabatic_tend", "tend_theta_euler", "delsq_theta", "tend_rtheta_physics"],
                         ["wdwz", "tend_w_euler", "w", "flux_arr", "wdtz", "tend_rtheta_adv", "rthdynten", "tend_theta_euler",
                          "tend_theta"]),
    "k_dt_cellC<false>": (["ru", "scr_e", "rho_zz", "uReconstructZonal", "uReconstructMeridional", "rw", "tend_w_euler", "theta_m",
                           "ru_save", "theta_m_save", "rw_save", "rt_diabatic_tend", "tend_theta_euler", "tend_rtheta_physics"],
                          ["ru_edge_w", "wdwz", "w", "flux_arr", "wdtz", "tend_rtheta_adv", "rthdynten", "tend_theta"]),
    "k_smlstep": (["u_tend", "zb_cell", "zb3_cell", "zz", "w"], ["w"]),
    "k_acoustic_flux:s0": (["ru_p", "theta_m"], ["rtheta_pp_old", "scr", "scr"]),
    "k_acoustic_flux": (["rtheta_pp", "ru_p", "theta_m"], ["rtheta_pp_old", "scr", "scr"]),
    "k_acoustic_column:s0": (["tend_rho", "theta_m", "w", "coftz", "cofwz", "cofwr", "cofwt", "a_tri", "alpha_tri", "zz",
                              "rw_save", "rw", "dss", "rho_zz", "scr", "scr"],
                             ["rho_pp", "rtheta_pp", "rw_p", "wwAvg"]),
    "k_acoustic_column": (["tend_rho", "theta_m", "w", "coftz", "cofwz", "cofwr", "cofwt", "a_tri", "alpha_tri", "zz",
                           "rw_save", "rw", "dss", "rho_zz", "scr", "scr", "rho_pp", "rtheta_pp", "rw_p", "wwAvg"],
                          ["rho_pp", "rtheta_pp", "rw_p", "wwAvg"]),
    "k_acoustic_tma<true>": (["scr", "scr", "theta_m", "tend_rho", "w", "coftz", "cofwz", "cofwr", "cofwt", "a_tri", "alpha_tri", "zz",
                              "rw_save", "rw", "dss", "rho_zz"], ["rtheta_pp_old", "rho_pp", "rtheta_pp", "rw_p", "wwAvg"]),
    "k_acoustic_tma<false>": (["scr", "scr", "theta_m", "tend_rho", "w", "coftz", "cofwz", "cofwr", "cofwt", "a_tri", "alpha_tri", "zz",
                               "rw_save", "rw", "dss", "rho_zz", "rho_pp", "rtheta_pp", "rw_p", "wwAvg"],
                              ["rtheta_pp_old", "rho_pp", "rtheta_pp", "rw_p", "wwAvg"]),
    "k_acoustic<true>": (["ru_p", "theta_m", "tend_rho", "w", "coftz", "cofwz", "cofwr", "cofwt", "a_tri", "alpha_tri", "zz",
                          "rw_save", "rw", "dss", "rho_zz"], ["rtheta_pp_old", "rho_pp", "rtheta_pp", "rw_p", "wwAvg"]),
    "k_acoustic<false>": (["ru_p", "theta_m", "tend_rho", "w", "coftz", "cofwz", "cofwr", "cofwt", "a_tri", "alpha_tri", "zz",
                           "rw_save", "rw", "dss", "rho_zz", "rho_pp", "rtheta_pp", "rw_p", "wwAvg"],
                          ["rtheta_pp_old", "rho_pp", "rtheta_pp", "rw_p", "wwAvg"]),
    "k_acoustic_gather": (["ru_p", "theta_m"], ["scr", "scr"]),
    # MPASB200_PHYSICS_CORRECTED
    "k_acoustic_u<true>": (["tend_ru"], ["ru_p", "ruAvg"]),
    "k_acoustic_u<false>": (["tend_ru", "rtheta_pp", "zz", "exner", "rho_pp", "cqu", "zxu", "ru_p", "ruAvg"], ["ru_p", "ruAvg"]),
    "k_acoustic_col<true>": (["scr", "scr", "theta_m", "tend_rho", "w", "coftz", "cofwz", "cofwr", "cofwt", "a_tri", "alpha_tri", "gamma_tri",
                              "zz", "rw_save", "rw", "dss", "rho_zz"], ["rtheta_pp_old", "rho_pp", "rtheta_pp", "rw_p", "wwAvg"]),
    "k_acoustic_col<false>": (["scr", "scr", "theta_m", "tend_rho", "w", "coftz", "cofwz", "cofwr", "cofwt", "a_tri", "alpha_tri", "gamma_tri",
                               "zz", "rw_save", "rw", "dss", "rho_zz", "rho_pp", "rtheta_pp", "rw_p", "wwAvg"],
                              ["rtheta_pp_old", "rho_pp", "rtheta_pp", "rw_p", "wwAvg"]),
    "k_divdamp": (["rtheta_pp", "rtheta_pp_old", "theta_m", "ru_p"], ["ru_p"]),
    "k_rec_cell1": (["rho_p_save", "rho_pp", "rho_base", "wwAvg", "rw_save", "rw_p", "zz", "rtheta_base", "rtheta_p_save",
                     "rtheta_pp"], ["rho_p", "rho_zz", "wwAvg", "rw", "w", "rtheta_p", "theta_m"]),
    "k_rec_edge": (["ruAvg", "ru_save", "ru_p", "rho_zz"], ["ruAvg", "ru", "u"]),
    "k_rec_cell2": (["w", "ru", "zb_cell", "zb3_cell", "rho_zz"], ["w"]),
    "k_finish_cell": (["wwAvg", "rho_zz_old_split"], ["wwAvg_split", "wwAvg", "rho_zz"]),
    "k_finish_edge": (["ruAvg"], ["ruAvg_split", "ruAvg"]),
    # atm_advance_scalars (8 scalars; horiz_flux_arr is library scratch: 8 edge fields = 24 units each way)
    "k_icd_cell1": (["rho_zz", "zz"], ["rho_zz"]),
    "k_icd_edge": (["u", "rho_zz"], ["ru"]),
    "k_icd_cell2": (["w", "rho_zz", "zz", "ru", "zb_cell", "zb3_cell", "rho_base", "theta_base", "theta_m"],
                    ["rw", "rho_p", "rtheta_base", "rtheta_p", "exner", "exner_base", "pressure_p", "pressure_base"]),
    "k_reconstruct": (["u"], ["uReconstructX", "uReconstructY", "uReconstructZ", "uReconstructZonal", "uReconstructMeridional"]),
    "k_zb_cell": (["zb", "zb3"], ["zb_cell", "zb3_cell"]),          # atm_compute_signs, 3-D part (one-time)
    "k_damping_coefs": (["zgrid"], ["dss"]),                        # atm_compute_damping_coefs (one-time)
    "k_jw_rw": (["zz", "ru", "rho_zz", "zb"], ["rw", "w"]),         # init_atm_case_jw (one-time)
    "k_setup_scalars": (["scalars"], ["scalars_old"]),
    "k_scalar_flux<NS>": (["ruAvg", "scalars"], ["scr_e"] * 8),
    "k_scalar_update<NS>": (["ruAvg", "scalars", "scalars_old", "wwAvg", "rho_zz", "rho_zz_old_split"] + ["scr_e"] * 8, ["scalars"]),
}
# the exact streaming acoustic kernel moves the same fields as the affine one
K["k_dt_edge_tile"] = K["k_dt_edge"]          # tile-staged form (kernels_tiles.cuh): same fields
K["k_acoustic_lane<true>"] = K["k_acoustic_tma<true>"]
K["k_acoustic_lane<false>"] = K["k_acoustic_tma<false>"]
# array-typed fields whose every slot is touched
FULL_SLOTS = {"scalars": 8, "scalars_old": 8, "zb": 2, "zb3": 2}


def canon(kernel: str) -> str:
    """name a launch is timed under (mpasb200_kernel_time) -> key of K: the staged forms (kernels_staged.cuh, `_s<slots>`)
    move the same fields as the plain kernels."""
    import re
    m = re.match(r"(k_\w+?)_s<\d+>$", kernel)
    return m.group(1) if m else kernel


def lookup(kernel: str):
    k = canon(kernel)
    return k if k in K else None


def _u(name: str) -> float:
    if name == "scr":
        return 1.0
    if name == "scr_e":
        return 3.0
    s = FIELD_SLOTS[name]
    return MULT[FIELD_ENTITY[name]] * (FULL_SLOTS.get(name, USED_SLOTS) if s > 1 else 1)


def units(kernel: str, scratch: bool = True) -> float:
    """8-byte units per cell-level moved by one launch of ``kernel`` (reads + writes)."""
    r, w = K[canon(kernel)]
    return sum(_u(n) for n in r + w if scratch or not n.startswith("scr"))


# One RK3 step, canonical sequence (stage 0 takes the rk_step == 0 branches), rk_timestep.rg:404-481.
def step_launches(canonical: bool = True, corrected_physics: bool = False) -> List[str]:
    seq = ["k_setup_cell", "k_setup_edge", "k_moist", "k_vert_imp"]
    for stage in range(3):
        if stage == 1:
            seq.append("k_vert_imp")
        if stage == 0 and canonical:
            seq += ["k_dt_cell0<true>", "k_dt_edge_delsq", "k_dt_vertex_delsq", "k_dt_cell_delsq", "k_dt_edge_euler", "k_dt_edge",
                    "k_dt_cellA", "k_dt_cellB", "k_dt_theta_flux", "k_dt_cellC<true>"]
        else:
            seq += ["k_dt_cell0<false>", "k_dt_edge", "k_dt_theta_flux", "k_dt_cellC<false>"]
        seq.append("k_smlstep")
        for ss in range((1 if stage < 2 else 2) + 1):
            t = "<true>" if ss == 0 else "<false>"
            if corrected_physics:
                seq += ["k_acoustic_u" + t, "k_acoustic_gather", "k_acoustic_col" + t, "k_divdamp"]
            else:
                seq += ["k_acoustic_gather", "k_acoustic_lane" + t, "k_divdamp"]
        if corrected_physics:
            seq += ["k_rec_pad", "k_rec_cell1", "k_rec_edge", "k_rec_cell2"]
        seq += ["k_diag_vertex", "k_diag_cell", "k_diag_edge<true>" if stage == 2 else "k_diag_edge<false>"]
    seq += ["k_finish_cell", "k_finish_edge"]
    return seq


def step_units(canonical: bool = True, scratch: bool = True, corrected_physics: bool = False) -> float:
    return sum(units(k, scratch) for k in step_launches(canonical, corrected_physics) if k != "k_rec_pad")


# the contract figures (units of 8 B per cell-level) -- per task call and per step.
SURVEY_TASK_UNITS = {
    "rk_integration_setup": 25, "compute_moist_coefficients": 2, "compute_vert_imp_coefs": 19,
    "compute_dyn_tend:rk0": 96, "compute_dyn_tend:rk>0": 53, "set_smlstep_pert_variables": 18,
    "advance_acoustic_step:s0": 22, "advance_acoustic_step": 26, "divergence_damping_3d": 9,
    "compute_solve_diagnostics": 24, "compute_solve_diagnostics:v": 27, "rk_dynamics_substep_finish": 14,
}
SURVEY_STEP_UNITS_CANONICAL = 643      # 5144 B per cell-level per step
SURVEY_STEP_UNITS_LITERAL = 600        # 4800 B

TASK_KERNELS = {
    "rk_integration_setup": ["k_setup_cell", "k_setup_edge"],
    "compute_moist_coefficients": ["k_moist"],
    "compute_vert_imp_coefs": ["k_vert_imp"],
    "compute_dyn_tend:rk0": ["k_dt_cell0<true>", "k_dt_edge_delsq", "k_dt_vertex_delsq", "k_dt_cell_delsq", "k_dt_edge_euler", "k_dt_edge",
                             "k_dt_cellA", "k_dt_cellB", "k_dt_theta_flux", "k_dt_cellC<true>"],
    "compute_dyn_tend:rk>0": ["k_dt_cell0<false>", "k_dt_edge", "k_dt_theta_flux", "k_dt_cellC<false>"],
    "set_smlstep_pert_variables": ["k_smlstep"],
    "advance_acoustic_step:s0": ["k_acoustic_gather", "k_acoustic_lane<true>"],
    "advance_acoustic_step": ["k_acoustic_gather", "k_acoustic_lane<false>"],
    "advance_acoustic_step:s0:affine": ["k_acoustic_gather", "k_acoustic_tma<true>"],
    "advance_acoustic_step:affine": ["k_acoustic_gather", "k_acoustic_tma<false>"],
    "advance_scalars": ["k_scalar_flux<NS>", "k_scalar_update<NS>"],
    "advance_acoustic_step:s0:fused": ["k_acoustic<true>"],
    "advance_acoustic_step:fused": ["k_acoustic<false>"],
    "advance_acoustic_step:s0:exact": ["k_acoustic_flux:s0", "k_acoustic_column:s0"],
    "advance_acoustic_step:exact": ["k_acoustic_flux", "k_acoustic_column"],
    "advance_acoustic_step:s0:corrected_physics": ["k_acoustic_u<true>", "k_acoustic_gather", "k_acoustic_col<true>"],
    "advance_acoustic_step:corrected_physics": ["k_acoustic_u<false>", "k_acoustic_gather", "k_acoustic_col<false>"],
    "recover_large_step_variables": ["k_rec_cell1", "k_rec_edge", "k_rec_cell2"],
    "divergence_damping_3d": ["k_divdamp"],
    "compute_solve_diagnostics": ["k_diag_vertex", "k_diag_cell", "k_diag_edge<false>"],
    "compute_solve_diagnostics:v": ["k_diag_vertex", "k_diag_cell", "k_diag_edge<true>"],
    "rk_dynamics_substep_finish": ["k_finish_cell", "k_finish_edge"],
}


def task_units(task: str, scratch: bool = False) -> float:
    return sum(units(k, scratch) for k in TASK_KERNELS[task])
