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


@pytest.mark.parametrize("twice", [False, True], ids=["once", "again_with_other_dts"])
def test_compute_vert_imp_coefs(warmed, twice):                             # dynamics_tasks.rg:513-592
    st, ora, f = warmed
    _reset(ora, f)
    cfg = ora.cfg
    rgas = cfg.rgas; rcv = rgas / (cfg.cp - rgas); c2 = cfg.cp * rcv; g = cfg.gravity
    cur = {k: f[k].copy() for k in ("cofrz", "cofwr", "cofwz", "coftz", "cofwt", "a_tri", "b_tri", "c_tri", "alpha_tri", "gamma_tri")}
    fzm, fzp, rdzw, rdzu = f["fzm"], f["fzp"], f["rdzw"], f["rdzu"]
    zz, ex, th = f["zz"], f["exner"], f["theta_m"]
    for dts in ((200.0, 300.0) if twice else (200.0,)):
        dtseps = .5 * dts * (1.0 + cfg.config_epssm)
        cur["cofrz"][:L] = dtseps * rdzw[:L]
        cofrz = cur["cofrz"]
        cur["gamma_tri"][:, 0] = 0.0
        k = np.arange(1, L)
        m = k - 1
        zf = fzm[k] * zz[:, k] + fzp[k] * zz[:, m]
        cur["cofwr"][:, k] = .5 * dtseps * g * zf
        cur["coftz"][:, :L] = 0.0
        cur["cofwz"][:, k] = dtseps * c2 * zf * rdzu[k] * f["cqw"][:, k] * (fzm[k] * ex[:, k] + fzp[k] * ex[:, m])
        cur["coftz"][:, k] = dtseps * (fzm[k] * th[:, k] + fzp[k] * th[:, m])
        K = slice(0, L)
        cur["cofwt"][:, K] = (.5 * dtseps * rcv * zz[:, K] * g * f["rho_base"][:, K] / (1.0 + f["qtot"][:, K]) * ex[:, K]
                              / ((f["rtheta_base"][:, K] + f["rtheta_p"][:, K]) * f["exner_base"][:, K]))
        cofwz, coftz, cofwr, cofwt = cur["cofwz"], cur["coftz"], cur["cofwr"], cur["cofwt"]
        cur["a_tri"][:, k] = (-1.0 * cofwz[:, k] * coftz[:, m] * rdzw[m] * zz[:, m] + cofwr[:, k] * cofrz[m]
                              - cofwt[:, m] * coftz[:, m] * rdzw[m])
        cur["b_tri"][:, k] = (1.0 + cofwz[:, k] * (coftz[:, k] * rdzw[k] * zz[:, k] + coftz[:, k] * rdzw[m] * zz[:, m])
                              - coftz[:, k] * (cofwt[:, k] * rdzw[k] - cofwt[:, k] * rdzw[m]) + cofwr[:, k] * ((cofrz[k] - cofrz[m])))
        cur["c_tri"][:, k] = (-1.0 * cofwz[:, k] * coftz[:, k + 1] * rdzw[k] * zz[:, k] - cofwr[:, k] * cofrz[k]
                              + cofwt[:, k] * coftz[:, k + 1] * rdzw[k])
        # alpha uses gamma_tri of level k-1 as it is BEFORE this call's gamma loop (Q10): a separate, earlier loop
        cur["alpha_tri"][:, k] = 1.0 / (cur["b_tri"][:, k] - cur["a_tri"][:, k] * cur["gamma_tri"][:, m])
        cur["gamma_tri"][:, k] = cur["c_tri"][:, k] * cur["alpha_tri"][:, k]
        ora.atm_compute_vert_imp_coefs(dts)
    _check(ora, cur)


@pytest.mark.parametrize("hollingsworth,rk_step", [(False, 0), (False, 2), (True, -1)])
def test_compute_solve_diagnostics(warmed, hollingsworth, rk_step):         # dynamics_tasks.rg:328-454
    st, ora, f = warmed
    _reset(ora, f)
    s = st.static
    nC, nE, nV = s["nEdgesOnCell"].shape[0], s["cellsOnEdge"].shape[0], s["edgesOnVertex"].shape[0]
    K = slice(0, L)
    zero = lambda n: np.zeros((n, L))
    get = lambda k, n: np.zeros(n) if s.get(k) is None else np.asarray(s[k], dtype=np.float64)
    c1, c2 = _idx(s["cellsOnEdge"][:, 0], nC), _idx(s["cellsOnEdge"][:, 1], nC)
    u = _pad(f["u"])[:, K]
    h = _pad(f["h"])[:, K]
    dc, dv = np.append(get("dcEdge", nE), 0.0), np.append(get("dvEdge", nE), 0.0)
    o = {k: f[k].copy() for k in ("h_edge", "ke_edge", "vorticity", "divergence", "ke", "ke_vertex", "v", "pv_vertex", "pv_edge")}
    o["h_edge"][:, K] = 0.5 * (h[c1] + h[c2])
    o["ke_edge"][:, K] = (dc[:nE] * dv[:nE])[:, None] * u[:nE] ** 2
    vort = zero(nV)
    for i in range(s["edgesOnVertex"].shape[1]):
        e = _idx(s["edgesOnVertex"][:, i], nE)
        vort = vort + (get("edgesOnVertexSign", (nV, 3))[:, i] * dc[e])[:, None] * u[e]
    o["vorticity"][:, K] = vort * get("invAreaTriangle", nV)[:, None]
    div, ke = zero(nC), zero(nC)
    kee = _pad(o["ke_edge"])[:, K]
    for i in range(s["edgesOnCell"].shape[1]):
        on = (i < s["nEdgesOnCell"])[:, None]
        e = _idx(s["edgesOnCell"][:, i], nE)
        div = np.where(on, div + ((s["edgesOnCellSign"][:, i] * dv[e])[:, None] + u[e]), div)         # "s + u", as written (:375)
        ke = np.where(on, ke + 0.25 * kee[e], ke)
    inva = get("invAreaCell", nC)[:, None]
    o["divergence"][:, K] = div * inva
    ke = ke * inva
    if hollingsworth:
        ev = [_idx(s["edgesOnVertex"][:, j], nE) for j in range(3)]
        o["ke_vertex"][:, K] = (kee[ev[0]] + kee[ev[1]] + kee[ev[2]]) * (0.25 * get("invAreaTriangle", nV))[:, None]
        ke_fact = 1.0 - 0.375
        ke = ke * ke_fact
        kev = _pad(o["ke_vertex"])[:, K]
        kav = np.concatenate([np.asarray(s["kiteAreasOnVertex"], dtype=np.float64), np.zeros((1, 3))])
        for i in range(s["edgesOnCell"].shape[1]):
            on = (i < s["nEdgesOnCell"])[:, None]
            vtx = _idx(s["verticesOnCell"][:, i], nV)
            j = np.asarray(s["kiteForCell"])[:, i]
            ke = np.where(on, ke + ((1.0 - ke_fact) * kav[vtx, j])[:, None] * kev[vtx] * inva, ke)
    o["ke"][:, K] = ke
    if not (rk_step != -1 and rk_step != 2):
        v = zero(nE)
        for i in range(1, s["edgesOnEdge_ECP"].shape[1]):                   # starts at 1, as written (:433)
            on = (i < s["nEdgesOnEdge"])[:, None]
            eoe = _idx(s["edgesOnEdge_ECP"][:, i], nE)
            v = np.where(on, v + s["weightsOnEdge"][:, i][:, None] * u[eoe], v)
        o["v"][:, K] = v
    o["pv_vertex"][:, K] = get("fVertex", nV)[:, None] + o["vorticity"][:, K]
    pvv = _pad(o["pv_vertex"])[:, K]
    v1, v2 = _idx(s["verticesOnEdge"][:, 0], nV), _idx(s["verticesOnEdge"][:, 1], nV)
    o["pv_edge"][:, K] = 0.5 * (pvv[v1] + pvv[v2])
    ora.atm_compute_solve_diagnostics(hollingsworth, rk_step)
    _check(ora, o)
