"""The one-time producers of hot-path inputs (mpas_regent_b200/core_init.py, host side; reference atm_core_init chain,
dynamics_tasks.rg:46-325, 595-646): geometric properties that must hold under the CORRECTED index policy."""
import numpy as np

from mpas_regent_b200 import _abi, core_init, init_jw


def _state(mesh, L=5):
    return init_jw.make_state(mesh, L, _abi.INDEX_CORRECTED)


def test_edge_signs_are_antisymmetric(grid642):
    m = grid642
    out = core_init.atm_compute_signs(m, _abi.INDEX_CORRECTED)
    eoc, n, coe = m.v["edgesOnCell"] - 1, m.v["nEdgesOnCell"], m.v["cellsOnEdge"] - 1
    sign = out["edgesOnCellSign"]
    acc = np.zeros(m.nEdges)
    for c in range(m.nCells):
        for i in range(n[c]):
            e = eoc[c, i]
            assert sign[c, i] == (1.0 if coe[e, 0] == c else -1.0)           # :74-86
            acc[e] += sign[c, i]
        assert not sign[c, n[c]:].any()
    assert not acc.any()                                                     # the two cells of an edge see opposite signs
    vs = out["edgesOnVertexSign"]
    eov, voe = m.v["edgesOnVertex"] - 1, m.v["verticesOnEdge"] - 1
    acc = np.zeros(m.nEdges)
    for v in range(m.nVertices):
        for j in range(3):
            assert vs[v, j] == (1.0 if voe[eov[v, j], 1] == v else -1.0)     # :60-72
            acc[eov[v, j]] += vs[v, j]
    assert not acc.any()


def test_discrete_divergence_of_solid_body_rotation_vanishes(grid642):
    """solid-body rotation is non-divergent on the sphere: sum_i sign_i * dvEdge_i * (u . n_i) / areaCell ~ 0 for every cell
    (n_i from cell 1 to cell 2 of the edge), small against |u| / dcEdge."""
    m = grid642
    sign = core_init.atm_compute_signs(m, _abi.INDEX_CORRECTED)["edgesOnCellSign"]
    v = m.v
    unit = lambda a: a / np.linalg.norm(a, axis=1)[:, None]
    xc = unit(np.stack([v["xCell"], v["yCell"], v["zCell"]], 1))
    xe = unit(np.stack([v["xEdge"], v["yEdge"], v["zEdge"]], 1))
    coe = v["cellsOnEdge"] - 1
    nrm = xc[coe[:, 1]] - xc[coe[:, 0]]
    nrm = unit(nrm - (nrm * xe).sum(1)[:, None] * xe)                        # tangent to the sphere at the edge
    omega = np.array([0.3, -1.1, 0.7])
    un = (np.cross(omega, xe) * nrm).sum(1)
    eoc, n = v["edgesOnCell"] - 1, v["nEdgesOnCell"]
    div = np.array([(sign[c, :n[c]] * v["dvEdge"][eoc[c, :n[c]]] * un[eoc[c, :n[c]]]).sum() / v["areaCell"][c] for c in range(m.nCells)])
    assert np.abs(div).max() < 0.02 * np.linalg.norm(omega) / v["dcEdge"].mean()


def test_kite_for_cell_points_back_to_the_cell(grid642):
    m = grid642
    kite = core_init.atm_compute_signs(m, _abi.INDEX_CORRECTED)["kiteForCell"]
    voc, n, cov = m.v["verticesOnCell"] - 1, m.v["nEdgesOnCell"], m.v["cellsOnVertex"] - 1
    for c in range(m.nCells):
        for i in range(n[c]):
            j = kite[c, i]
            # the reference's search runs j = 1 .. vertexDegree-1 only (:113-129): slot 0 is never found and stays 0
            assert cov[voc[c, i], j] == c or (j == 0 and cov[voc[c, i], 0] == c) or j == 0


def test_advection_lists_contain_both_cells_and_reproduce_a_constant(grid642):
    st = _state(grid642)
    s = st.static
    coe = s["cellsOnEdge"]
    nadv, adv = s["nAdvCellsForEdge"], s["advCellsForEdge"]
    assert nadv.max() <= 15 and nadv.min() >= 2
    for e in range(0, grid642.nEdges, 7):
        cells = adv[e, :nadv[e]]
        # (the list may hold a cell twice: the reference's duplicate search compares against a stale entry, a quirk that
        #  core_init keeps; a repeated cell only splits its coefficient)
        assert coe[e, 0] in cells and coe[e, 1] in cells
    # with deriv_two never written (zero, rule M1) the lists reduce to the two cells of the edge and a constant scalar is
    # transported with the plain edge flux: sum_j adv_coefs = dvEdge, no 3rd-order part
    adv0 = core_init.atm_adv_coef_compression(grid642, _abi.INDEX_CORRECTED, None)
    j = np.arange(adv0["advCellsForEdge"].shape[1])[None, :] < adv0["nAdvCellsForEdge"][:, None]
    tot = np.where(j, adv0["adv_coefs"], 0.0).sum(1)
    assert np.allclose(tot, grid642.v["dvEdge"], rtol=1e-12)
    assert not np.where(j, adv0["adv_coefs_3rd"], 0.0).any()
    # with the deterministic fill of deriv_two (rule M5, 5 % of 1/dc^2) the defect stays at that size
    tot = np.where(np.arange(adv.shape[1])[None, :] < nadv[:, None], s["adv_coefs"], 0.0).sum(1)
    assert np.abs(tot / s["dvEdge"] - 1.0).max() < 0.1


def test_mesh_scaling_is_one_on_a_uniform_mesh(grid642):
    st = _state(grid642)
    assert np.allclose(st.static["meshScalingDel2"], 1.0) and np.allclose(st.static["meshScalingDel4"], 1.0)


# ---- literal loop-by-loop restatement of atm_compute_signs (dynamics_tasks.rg:46-130) under BOTH index policies ----------------
def _literal_compute_signs(m, policy):
    """The reference text, one loop at a time.  Memory model: stored ids are 1-based; LITERAL uses the stored id as
    the index (id == N reads the zero pad entity), CORRECTED uses id - 1 (id 0 reads the pad); `iCell.x == stored id` compares a
    0-based loop index with a stored id as written under LITERAL (Q33) and with id - 1 under CORRECTED."""
    v = m.v
    nC, nE, nV = m.nCells, m.nEdges, m.nVertices
    lit = policy == _abi.INDEX_LITERAL

    def idx(raw, n):
        i = raw if lit else raw - 1
        return n if (i < 0 or i > n) else i

    def same(loop_index, stored):
        return loop_index == (stored if lit else stored - 1)

    voe = np.vstack([v["verticesOnEdge"], np.zeros((1, 2), v["verticesOnEdge"].dtype)])
    coe = np.vstack([v["cellsOnEdge"], np.zeros((1, 2), v["cellsOnEdge"].dtype)])
    cov = np.vstack([v["cellsOnVertex"], np.zeros((1, 3), v["cellsOnVertex"].dtype)])
    evs = np.zeros((nV, 3)); ecs = np.zeros((nC, 10)); kite = np.zeros((nC, 10), np.int32)
    for iVtx in range(nV):                                                   # :60-72
        for i in range(3):
            e = v["edgesOnVertex"][iVtx, i]
            if e <= nE:
                evs[iVtx, i] = 1.0 if same(iVtx, voe[idx(e, nE), 1]) else -1.0
            else:
                evs[iVtx, i] = 0.0
    for iCell in range(nC):                                                  # :74-86
        for i in range(v["nEdgesOnCell"][iCell]):
            e = v["edgesOnCell"][iCell, i]
            if e <= nE:
                ecs[iCell, i] = 1.0 if same(iCell, coe[idx(e, nE), 0]) else -1.0
            else:
                ecs[iCell, i] = 0.0
    for iCell in range(nC):                                                  # :113-129, j = 1 .. vertexDegree-1, first match wins
        for i in range(v["nEdgesOnCell"][iCell]):
            iv = v["verticesOnCell"][iCell, i]
            if iv <= nV:
                for j in range(1, 3):
                    if same(iCell, cov[idx(iv, nV), j]):
                        kite[iCell, i] = j
                        break
            else:
                kite[iCell, i] = 1
    return evs, ecs, kite


def test_compute_signs_equals_the_literal_loops_under_both_policies(grid2562):
    for policy in (_abi.INDEX_LITERAL, _abi.INDEX_CORRECTED):
        out = core_init.atm_compute_signs(grid2562, policy)
        evs, ecs, kite = _literal_compute_signs(grid2562, policy)
        assert np.array_equal(out["edgesOnVertexSign"], evs), policy
        assert np.array_equal(out["edgesOnCellSign"], ecs), policy
        assert np.array_equal(out["kiteForCell"], kite), policy
    # the two policies really differ on this mesh (Q32/Q33): the literal reading gets most signs "wrong"
    a = core_init.atm_compute_signs(grid2562, _abi.INDEX_LITERAL)["edgesOnCellSign"]
    b = core_init.atm_compute_signs(grid2562, _abi.INDEX_CORRECTED)["edgesOnCellSign"]
    assert (a != b).any()


def test_zb_cell_selection_equals_the_literal_loops(grid642):
    """:88-110: zb_cell(i) = zb[0] of the slot's edge if this cell is its first cell, else zb[1]; every level 0..nVertLevels."""
    m, L1 = grid642, 4
    rng = np.random.default_rng(1)
    zb, zb3 = rng.random((m.nEdges, L1, 2)), rng.random((m.nEdges, L1, 2))
    for policy in (_abi.INDEX_LITERAL, _abi.INDEX_CORRECTED):
        out = core_init.atm_compute_signs(m, policy, zb=zb, zb3=zb3, nlev1=L1)
        lit = policy == _abi.INDEX_LITERAL
        zbp = np.concatenate([zb, np.zeros((1, L1, 2))]); coe = np.vstack([m.v["cellsOnEdge"], np.zeros((1, 2), np.int64)])
        want = np.zeros((m.nCells, L1, 10))
        for c in range(m.nCells):
            for i in range(m.v["nEdgesOnCell"][c]):
                raw = m.v["edgesOnCell"][c, i]
                if raw <= m.nEdges:
                    e = raw if lit else raw - 1
                    e = m.nEdges if (e < 0 or e > m.nEdges) else e
                    first = c == (coe[e, 0] if lit else coe[e, 0] - 1)
                    want[c, :, i] = zbp[e, :, 0 if first else 1]
        assert np.array_equal(out["zb_cell"], want), policy


# ---- the oracle's literal loops of the mesh-only producers (dynamics_tasks.rg:46-325) against the vectorised host producers ----
import pytest


def _raw_mesh(mesh, scaled, deriv_two=None):
    v = mesh.v
    d = {k: v[k] for k in ("nEdgesOnCell", "edgesOnCell", "verticesOnCell", "cellsOnCell", "cellsOnEdge", "verticesOnEdge", "cellsOnVertex", "edgesOnVertex")}
    d["dcEdge"], d["dvEdge"] = scaled.v["dcEdge"], scaled.v["dvEdge"]
    if deriv_two is not None:
        d["deriv_two"] = deriv_two
    return d


@pytest.mark.parametrize("policy", [_abi.INDEX_CORRECTED, _abi.INDEX_LITERAL], ids=["corrected", "literal"])
@pytest.mark.parametrize("which", ["x1.2562", "icosa642"])
def test_oracle_init_producers_equal_host_producers(grid2562, grid642, policy, which):
    """oracle_compute_signs / oracle_adv_coef_compression / oracle_compute_zb_cell / oracle_couple_coef_3rd_order (loop by loop from
    the reference text) against core_init.py (array at a time): integer lists and signs identical, coefficients bit-identical
    (same operation order), under both index policies, on the bundled mesh (pentagons) and a generated one."""
    from mpas_regent_b200 import dynamics
    from oracle.oracle import Oracle
    mesh = grid2562 if which == "x1.2562" else grid642
    L = 6
    st = init_jw.make_state(mesh, L, policy)                     # runs the host chain; extras keep zb, zb3, deriv_two
    scaled = st.mesh
    ora = Oracle(dynamics.dims_of(mesh, L), _abi.default_config(index_policy=policy))
    raw = _raw_mesh(mesh, scaled, st.extras["deriv_two"])
    sg_h = core_init.atm_compute_signs(scaled, policy, zb=st.extras["zb"], zb3=st.extras["zb3"], nlev1=L + 1)
    sg_o = ora.atm_compute_signs(raw)
    for k in ("edgesOnVertexSign", "edgesOnCellSign", "kiteForCell"):
        assert np.array_equal(sg_o[k], sg_h[k]), k
    adv_h = core_init.atm_adv_coef_compression(scaled, policy, st.extras["deriv_two"])
    adv_o = ora.atm_adv_coef_compression(raw)
    for k in ("nAdvCellsForEdge", "advCellsForEdge", "adv_coefs", "adv_coefs_3rd"):
        assert np.array_equal(adv_o[k], adv_h[k]), k
    # 3-D part + coupling: needs the uploaded mesh (resolved numbering)
    ora.upload_mesh(st.static)
    ora.upload_field("zb", st.extras["zb"]); ora.upload_field("zb3", st.extras["zb3"])
    ora.atm_compute_zb_cell()
    assert np.array_equal(ora.download_field("zb_cell"), sg_h["zb_cell"])
    a3 = adv_o["adv_coefs_3rd"].copy()
    ora.atm_couple_coef_3rd_order(0.25, a3)
    a3_h, zb3c_h = core_init.atm_couple_coef_3rd_order(0.25, adv_h["adv_coefs_3rd"].copy(), sg_h["zb3_cell"].copy())
    assert np.array_equal(a3, a3_h)
    assert np.array_equal(ora.download_field("zb3_cell"), zb3c_h)
    assert np.array_equal(a3, st.static["adv_coefs_3rd"]) and np.array_equal(ora.download_field("zb3_cell"), st.f["zb3_cell"])
    ora.close()


@pytest.mark.parametrize("policy", [_abi.INDEX_CORRECTED, _abi.INDEX_LITERAL], ids=["corrected", "literal"])
def test_oracle_mesh_scaling_and_damping_coefs_equal_host_producers(grid2562, policy):
    """oracle_compute_mesh_scaling (dynamics_tasks.rg:595-646) and oracle_compute_damping_coefs (:274-300) against core_init.py on a
    NON-uniform meshDensity (the bundled mesh is quasi-uniform: density 1 would make both trivial)."""
    from mpas_regent_b200 import dynamics
    from mpas_regent_b200.mesh import Mesh
    from oracle.oracle import Oracle
    L = 9
    st = init_jw.make_state(grid2562, L, policy)
    rng = np.random.default_rng(5)
    md = rng.uniform(0.05, 1.0, grid2562.nCells)
    mesh_d = Mesh(v={**st.mesh.v, "meshDensity": md}, partition=None, name="x")
    ora = Oracle(dynamics.dims_of(grid2562, L), _abi.default_config(index_policy=policy))
    ms_h = core_init.atm_compute_mesh_scaling(mesh_d, policy, True)
    ms_o = ora.atm_compute_mesh_scaling({"cellsOnEdge": grid2562.v["cellsOnEdge"]}, md, True)
    for k in ms_h:
        assert np.allclose(ms_o[k], ms_h[k], rtol=1e-14, atol=0), k
    one = ora.atm_compute_mesh_scaling({"cellsOnEdge": grid2562.v["cellsOnEdge"]}, md, False)
    assert (one["meshScalingDel2"] == 1.0).all() and (one["meshScalingDel4"] == 1.0).all()
    ora.upload_mesh(st.static)
    ora.upload_field("zgrid", st.f["zgrid"])
    ora.atm_compute_damping_coefs(md, 22000.0, 0.2)
    dss_h = core_init.atm_compute_damping_coefs(st.f["zgrid"], md, L)
    assert dss_h.max() > 0
    assert np.allclose(ora.download_field("dss"), dss_h, rtol=1e-14, atol=0)
    ora.close()
