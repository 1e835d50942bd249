"""Offline L1 model for the gather kernels (host only).  Every load of a kernel is one 448-byte column strip (level pairs of
one entity); an item is (field, column).  'cold' = one tile of 8 entities per block, nothing cached at the start (what
ships); 'chunked(cap)' = a block walks consecutive tiles and keeps the last `cap` items (LRU).  Compare 'cold' with the
L1 hit rates ncu reports in r1_ncu_top_kernels.md.   python profiles/l1_locality_kernels.py [nCells]"""
import os
import sys
from collections import OrderedDict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpas_regent_b200 import _abi, core_init, icosa, partition  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40962
mesh = icosa.make_icosahedral_mesh(n)
v = mesh.v
nC, nE, nV = mesh.nCells, mesh.nEdges, mesh.nVertices
ck = partition.hilbert_keys(v["xCell"], v["yCell"], v["zCell"])
cNew = np.empty(nC, np.int64); cNew[np.argsort(ck, kind="stable")] = np.arange(nC)
coe = v["cellsOnEdge"] - 1
a, b = cNew[coe[:, 0]], cNew[coe[:, 1]]
eOrder = np.argsort((np.minimum(a, b) << 32) | np.maximum(a, b), kind="stable")
eNew = np.empty(nE, np.int64); eNew[eOrder] = np.arange(nE)
cOrder = np.argsort(cNew)
# connectivity in the internal numbering, rows in internal order
c1c2 = cNew[coe][eOrder]                                                # [nE, 2]
eoe = np.where(np.arange(v["edgesOnEdge"].shape[1])[None, :] < v["nEdgesOnEdge"][:, None], eNew[np.clip(v["edgesOnEdge"] - 1, 0, nE - 1)], -1)[eOrder]
eoc = np.where(np.arange(v["edgesOnCell"].shape[1])[None, :] < v["nEdgesOnCell"][:, None], eNew[np.clip(v["edgesOnCell"] - 1, 0, nE - 1)], -1)[cOrder]
adv = core_init.atm_adv_coef_compression(mesh, _abi.INDEX_CORRECTED, None)
# the advection list of an edge = cells around its two cells (use cellsOnCell of both cells: 10 distinct cells, like nAdvCellsForEdge)
coc = np.where(np.arange(v["cellsOnCell"].shape[1])[None, :] < v["nEdgesOnCell"][:, None], cNew[np.clip(v["cellsOnCell"] - 1, 0, nC - 1)], -1)
advc = np.concatenate([coc[coe[:, 0]], coc[coe[:, 1]]], axis=1)[eOrder]
c1c2_of_cell_edges = np.where(eoc[:, :, None] >= 0, c1c2[np.clip(eoc, 0, nE - 1)], -1).reshape(nC, -1)     # cells at both ends of a cell's edges


def items(rows, groups):
    """list of item ids touched by entity rows [r0, r1): groups = [(tag, index array [n, w], n_fields)]"""
    out = []
    for tag, idx, nf in groups:
        for f in range(nf):
            x = idx[rows[0]:rows[1]].ravel()
            out.append(((tag * 16 + f) << 40) + x[x >= 0])
    return np.concatenate(out)


def model(name, n_ent, own_loads, groups, tiles=3000, caps=(64, 128, 256)):
    t0 = (n_ent // 8) // 3
    cold_h = cold_t = 0
    for t in range(t0, t0 + tiles):
        it = items((t * 8, t * 8 + 8), groups)
        cold_t += len(it) + own_loads * 8
        cold_h += len(it) - len(np.unique(it))
    res = [f"{name:22s} cold {cold_h / cold_t * 100:5.1f}%"]
    for cap in caps:
        lru = OrderedDict(); h = tot = 0
        for t in range(t0, t0 + tiles):
            it = items((t * 8, t * 8 + 8), groups)
            tot += len(it) + own_loads * 8
            for x in it.tolist():
                if x in lru:
                    h += 1; lru.move_to_end(x)
                else:
                    lru[x] = 1
                    if len(lru) > cap:
                        lru.popitem(last=False)
        res.append(f"chunked({cap}) {h / tot * 100:5.1f}%")
    print("   ".join(res))


print(f"x1.{n}: hit rate over all loads of the kernel (own-column loads always miss)")
E, C = 1, 2
model("k_dt_edge", nE, 7, [(E, eoe, 2), (C, c1c2, 5)])                      # pv_edge,u at edgesOnEdge; rw,ke,h_div,w,w(+1) at the two cells
model("k_dt_edge_euler", nE, 4, [(C, c1c2, 6)])                             # pressure_p,zz,dpdz,kdiff,delsq_divergence,divergence at the two cells (vertices left out)
model("k_dt_theta_flux", nE, 1, [(C, advc, 1)])                             # theta_m at the advection cells
model("k_divdamp", nE, 1, [(C, c1c2, 3)])                                   # rtheta_pp, rtheta_pp_old, theta_m
model("k_acoustic_gather", nC, 0, [(E, eoc, 1), (C, c1c2_of_cell_edges, 1)])  # ru_p at the cell's edges, theta_m at their cells
model("k_dt_cellC<false>", nC, 12, [(E, eoc, 3), (C, c1c2_of_cell_edges, 1)])  # ru, flux, ru_save at the edges; theta_m_save at their cells
