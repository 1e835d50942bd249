"""A second, array-at-a-time restatement (numpy, one level or one edge slot at a time) of four task bodies, written from the
reference text and independent of oracle/mpas_oracle.cpp's loop form.  It pins the oracle against transcription slips:
both must agree to rounding (1e-13 relative).  Memory-model rules as in the oracle header (M1-M4, level -1 reads 0)."""
import numpy as np
import pytest

from mpas_regent_b200 import _abi
from tests.util import build_pair

L = 7
TOL = 1e-13


def _idx(ids, n):                       # INDEX_CORRECTED: stored id - 1, 0 -> pad entity n
    ids = np.asarray(ids).astype(np.int64)
    return np.where(ids > 0, ids - 1, n)


def _pad(a):                            # append the zero pad entity
    return np.concatenate([a, np.zeros((1,) + a.shape[1:], a.dtype)])


def _below(a):                          # a[:, k-1] with level -1 = 0
    return np.concatenate([np.zeros_like(a[:, :1]), a[:, :-1]], axis=1)


@pytest.fixture(scope="module")
def warmed(grid642):
    st, ora, _ = build_pair(grid642, L, _abi.INDEX_CORRECTED, gpu=False)
    ora.atm_compute_solve_diagnostics(False, -1)
    ora.atm_srk3(600.0)
    pre = ora.download_all()
    yield st, ora, pre
    ora.close()


def _reset(ora, pre):
    for n, a in pre.items():
        ora.upload_field(n, a)


def _check(ora, want):
    for n, a in want.items():
        got = ora.download_field(n)
        fin = np.isfinite(a)
        assert np.array_equal(fin, np.isfinite(got)), n
        scale = np.abs(a[fin]).max() if fin.any() else 0.0
        err = np.abs(got[fin] - a[fin]).max() if fin.any() else 0.0
        assert err <= TOL * scale, (n, err, scale)


def test_divergence_damping_3d(warmed):                                     # dynamics_tasks.rg:1736-1763
    st, ora, f = warmed
    _reset(ora, f)
    s, cfg = st.static, ora.cfg
    nC = s["nEdgesOnCell"].shape[0]
    dts = 300.0
    c1, c2 = _idx(s["cellsOnEdge"][:, 0], nC), _idx(s["cellsOnEdge"][:, 1], nC)
    coef = 2.0 * cfg.config_smdiv * cfg.config_len_disp * (1.0 / dts)
    rpp, rppo, tm = _pad(f["rtheta_pp"]), _pad(f["rtheta_pp_old"]), _pad(f["theta_m"])
    shared = _pad(np.asarray(s["isShared"]).astype(bool)) if s.get("isShared") is not None else np.zeros(nC + 1, bool)
    on = ~(shared[c1] & shared[c2])
    div1, div2 = -(rpp[c1] - rppo[c1]), -(rpp[c2] - rppo[c2])
    with np.errstate(all="ignore"):
        upd = f["ru_p"] + coef * (div2 - div1) * (1.0 - s["specZoneMaskEdge"][:, None]) / (tm[c1] + tm[c2])
    want = f["ru_p"].copy()
    want[on, :L] = upd[on, :L]
    ora.atm_divergence_damping_3d(dts)
    _check(ora, {"ru_p": want})
    assert not np.array_equal(want, f["ru_p"])


def test_set_smlstep_pert_variables(warmed):                                # dynamics_tasks.rg:1503-1528, every level incl. 0 (Q22)
    st, ora, f = warmed
    _reset(ora, f)
    s, cfg = st.static, ora.cfg
    nE = s["cellsOnEdge"].shape[0]
    fzm, fzp = f["fzm"][:L], f["fzp"][:L]
    ut = _pad(f["u_tend"])
    w = f["w"].copy()
    act = (s["bdyMaskCell"] <= cfg.nRelaxZone)[:, None]
    for i in range(s["edgesOnCell"].shape[1]):
        on = act & (i < s["nEdgesOnCell"])[:, None]
        e = _idx(s["edgesOnCell"][:, i], nE)
        u_k, u_m = ut[e][:, :L], _below(ut[e])[:, :L]
        flux = s["edgesOnCell_sign"][:, i][:, None] * (fzm * u_k + fzp * u_m)
        w[:, :L] = np.where(on, w[:, :L] - (f["zb_cell"][:, :L, i] + np.copysign(1.0, u_k) * f["zb3_cell"][:, :L, i]) * flux, w[:, :L])
    w[:, :L] = np.where(act, w[:, :L] * (fzm * f["zz"][:, :L] + fzp * _below(f["zz"])[:, :L]), w[:, :L])
    ora.atm_set_smlstep_pert_variables()
    _check(ora, {"w": w})


@pytest.mark.parametrize("small_step", [0, 1])
def test_advance_acoustic_step_literal(warmed, small_step):                 # dynamics_tasks.rg:1615-1704, point by point
    st, ora, f = warmed
    _reset(ora, f)
    s, cfg = st.static, ora.cfg
    nC, nE = s["nEdgesOnCell"].shape[0], s["cellsOnEdge"].shape[0]
    dts = 300.0
    epssm = cfg.config_epssm
    resm = (1.0 - epssm) / (1.0 + epssm)
    cofrz, rdzw, fzm, fzp = (f[k] for k in ("cofrz", "rdzw", "fzm", "fzp"))
    o = {k: f[k].copy() for k in ("rw_p", "wwAvg", "rho_pp", "rtheta_pp", "rtheta_pp_old")}
    rw_p, ww, rho_pp, rt_pp = o["rw_p"], o["wwAvg"], o["rho_pp"], o["rtheta_pp"]
    o["rtheta_pp_old"][:, :L] = 0.0 if small_step == 0 else f["rtheta_pp"][:, :L]
    if small_step == 0:
        rw_p[:] = 0; ww[:] = 0
    c1, c2 = np.append(_idx(s["cellsOnEdge"][:, 0], nC), nC), np.append(_idx(s["cellsOnEdge"][:, 1], nC), nC)
    rup, tm = _pad(f["ru_p"]), _pad(f["theta_m"])
    dv = np.append(s["dvEdge"], 0.0)
    zz, w = f["zz"], f["w"]
    assert np.all(s["specZoneMaskCell"] == 0.0)
    for k in range(L):                                      # levels ascend (M4): level k sees the NEW level k-1
        if small_step == 0:
            rho_pp[:, k] = 0; rt_pp[:, k] = 0
        rs, ts = np.zeros(nC), np.zeros(nC)
        for i in range(s["edgesOnCell"].shape[1]):
            on = i < s["nEdgesOnCell"]
            e = _idx(s["edgesOnCell"][:, i], nE)
            flux = s["edgesOnCellSign"][:, i] * dts * dv[e] * rup[e, k] * s["invAreaCell"]
            rs = np.where(on, rs - flux, rs)
            ts = np.where(on, ts - flux * 0.5 * (tm[c2[e], k] + tm[c1[e], k]), ts)
        rs = rho_pp[:, k] + dts * f["tend_rho"][:, k] + rs - cofrz[k] * resm * (rw_p[:, k + 1] - rw_p[:, k])
        ts = rt_pp[:, k] + dts * f["theta_m"][:, k] + ts - resm * rdzw[k] * (f["coftz"][:, k + 1] * rw_p[:, k + 1] - f["coftz"][:, k] * rw_p[:, k])
        if k > 0:
            ww[:, k] += 0.5 * (1.0 - epssm) * rw_p[:, k]
            # rs[k-1] and ts[k-1] were re-zeroed at this point (Q25)
            rw_p[:, k] += (dts * w[:, k] - f["cofwz"][:, k] * ((zz[:, k] * ts - zz[:, k - 1] * 0.0) + resm * (zz[:, k] * rt_pp[:, k] - zz[:, k - 1] * rt_pp[:, k - 1]))
                           - f["cofwr"][:, k] * ((rs + 0.0) + resm * (rho_pp[:, k] + rho_pp[:, k - 1]))
                           + f["cofwt"][:, k] * (ts + resm * rt_pp[:, k])
                           + f["cofwt"][:, k - 1] * (0.0 + resm * rt_pp[:, k - 1]))
            rw_p[:, k] -= f["a_tri"][:, k] * rw_p[:, k - 1]
            rw_p[:, k] *= f["alpha_tri"][:, k]
            d3 = f["rw_save"][:, k] - f["rw"][:, k]
            rw_p[:, k] += d3 - dts * f["dss"][:, k] * (fzm[k] * zz[:, k] + fzp[k] * zz[:, k - 1]) * (fzm[k] * f["rho_zz"][:, k] + fzp[k] * f["rho_zz"][:, k - 1]) * w[:, k]
            rw_p[:, k] /= (1.0 + dts * f["dss"][:, k])
            rw_p[:, k] -= d3
            ww[:, k] += 0.5 * (1.0 + epssm) * rw_p[:, k]
        rho_pp[:, k] = rs - cofrz[k] * (rw_p[:, k + 1] - rw_p[:, k])
        rt_pp[:, k] = ts - rdzw[k] * (f["coftz"][:, k + 1] * rw_p[:, k + 1] - f["coftz"][:, k] * rw_p[:, k])
    ora.atm_advance_acoustic_step(dts, small_step)
    _check(ora, o)
    assert np.abs(o["rw_p"]).max() > 0


@pytest.mark.parametrize("rk_step", [0, 2])
def test_recover_large_step_variables_literal(warmed, rk_step):             # dynamics_tasks.rg:1786-1871, as written
    st, ora, f = warmed
    _reset(ora, f)
    s, cfg = st.static, ora.cfg
    nC, nE = s["nEdgesOnCell"].shape[0], s["cellsOnEdge"].shape[0]
    ns, dt = 2, 600.0
    rgas = cfg.rgas; rcv = rgas / (cfg.cp - rgas); p0 = 100000
    fzm, fzp = f["fzm"][:L], f["fzp"][:L]
    K = slice(0, L)
    o = {k: f[k].copy() for k in ("rho_p", "rho_zz", "w", "wwAvg", "rw", "rtheta_p", "theta_m", "exner", "pressure_p", "ruAvg", "ru", "u")}
    invNs = 1 / float(ns)
    o["rho_p"][:, K] = f["rho_p_save"][:, K] + f["rho_pp"][:, K]
    o["rho_zz"][:, K] = o["rho_p"][:, K] + f["rho_base"][:, K]
    o["wwAvg"][:, K] = f["wwAvg"][:, K] * invNs + f["rw_save"][:, K]
    o["rw"][:, K] = f["rw_save"][:, K] + f["rw_p"][:, K]
    with np.errstate(all="ignore"):
        o["w"][:, K] = o["rw"][:, K] / (fzm * f["zz"][:, K] + fzp * _below(f["zz"])[:, K])
        if rk_step == 2:
            o["rtheta_p"][:, K] = f["rtheta_p_save"][:, K] + f["rtheta_pp"][:, K] - dt * o["rho_zz"][:, K] * f["rt_diabatic_tend"][:, K]
            o["theta_m"][:, K] = (o["rtheta_p"][:, K] + f["rtheta_base"][:, K]) / o["rho_zz"][:, K]
            o["exner"][:, K] = f["zz"][:, K] * (rgas / p0) * np.power(o["rtheta_p"][:, K] + f["rtheta_base"][:, K], rcv)
            o["pressure_p"][:, K] = f["zz"][:, K] * rgas * (o["exner"][:, K] * o["rtheta_p"][:, K] + f["rtheta_base"][:, K] * (o["exner"][:, K] - f["exner_base"][:, K]))
        else:
            o["rtheta_p"][:, K] = f["rtheta_p_save"][:, K] + f["rtheta_pp"][:, K]
            o["theta_m"][:, K] = (o["rtheta_p"][:, K] + f["rtheta_base"][:, K]) / o["rho_zz"][:, K]
        c1, c2 = _idx(s["cellsOnEdge"][:, 0], nC), _idx(s["cellsOnEdge"][:, 1], nC)
        rz = _pad(o["rho_zz"]); rz[nC, :L] = 1.0                                   # the "garbage cell" (:1792-1794)
        o["ruAvg"][:, K] = f["ruAvg"][:, K] * invNs + f["ru_save"][:, K]
        o["ru"][:, K] = f["ru_save"][:, K] * f["ru_p"][:, K]                      # a product, as written (:1840)
        o["u"][:, K] = 2 * o["ru"][:, K] / (rz[c1][:, K] + rz[c2][:, K])
        ru = _pad(o["ru"])
        cf1, cf2, cf3 = f["cf1"][0], f["cf2"][0], f["cf3"][0]
        act = s["bdyMaskCell"] <= cfg.nRelaxZone
        w = o["w"]
        for y in range(L):                                                        # the surface term is added at EVERY level's visit
            for i in range(s["edgesOnCell"].shape[1]):
                on = act & (i < s["nEdgesOnCell"])
                e = _idx(s["edgesOnCell"][:, i], nE)
                sg = s["edgesOnCell_sign"][:, i]
                flux = cf1 * ru[e, 0] + cf2 * ru[e, 1] + cf3 * ru[e, 2]
                w[:, 0] = np.where(on, w[:, 0] + sg * (f["zb_cell"][:, 0, i] + np.copysign(1.0, flux) * f["zb3_cell"][:, 0, i]) * flux, w[:, 0])
                ru_m = ru[e, y - 1] if y > 0 else np.zeros(nC)
                flux2 = fzm[y] * ru[e, y] * (fzp[y] * ru_m)
                w[:, y] = np.where(on, w[:, y] + sg * (f["zb_cell"][:, y, i] + np.copysign(1.0, flux2) * f["zb3_cell"][:, y, i]) * flux2, w[:, y])
        w[:, 0] = np.where(act, w[:, 0] / (cf1 * o["rho_zz"][:, 0] + cf2 * o["rho_zz"][:, 1] + cf3 * o["rho_zz"][:, 2]), w[:, 0])
        for y in range(1, L):
            w[:, y] = np.where(act, w[:, y] / (fzm[y] * o["rho_zz"][:, y] + fzp[y] * o["rho_zz"][:, y - 1]), w[:, y])
    ora.atm_recover_large_step_variables(ns, rk_step, dt)
    _check(ora, o)
    assert np.all(ora.download_pad("rho_zz")[:L] == 1.0)
