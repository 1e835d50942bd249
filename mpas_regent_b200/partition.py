"""Owned / ghost dependent partitioning of the cell region and the per-rank local meshes.

Mirrors ``partition_regions`` (reference: mesh_loading/mesh_loading.rg:399-483;
data_structures.rg:577-584; README.md:121-141): from a cell colouring ``p`` it derives, per colour,
``ghost_1`` (cells one edge away), ``ghost_2`` (one and two edges away), ``shared_1`` (owned cells
next to ghost_1), ``private_1 = p - shared_1``, ``shared_2`` (shared_1 plus owned cells next to it)
and ``private_2 = private_1 - shared_2``, with the same image / preimage algebra through
``edgesOnCell[0..9]`` and ``cellOne`` / ``cellTwo``.  All sets are sets of CELLS (the reference's
are (cell, level) points: multiply by nVertLevels to get the volumes it prints, :473-478).

The colouring comes from a METIS ``.part.N`` file when there is one (``Mesh.partition``), otherwise
from contiguous chunks of a Hilbert space-filling-curve order (compact halos, deterministic).

``build_local`` turns one colour into the mesh a rank hands to libmpas_b200: cells ordered
[owned | ghost ring 1 | ghost ring 2], every edge and vertex of those cells, 1-based local ids with
0 = "not held here" (resolved to the zero pad entity under the CORRECTED index policy), edges owned
by the owner of cellsOnEdge[0], vertices by the owner of cellsOnVertex[0], and per-peer send / recv
index lists that are bit-for-bit the same on both sides.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from .mesh import CORRECTED, LITERAL, MAX_EDGES, Mesh, resolve_ids


# ------------------------------------------------------------------------------------------------
# colouring
def _spread21(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def hilbert_keys(x: np.ndarray, y: np.ndarray, z: np.ndarray) -> np.ndarray:
    """63-bit Hilbert keys of points on the sphere (Skilling's transpose algorithm, 21 bits per axis);
    the same curve libmpas_b200 uses for its internal renumbering."""
    r = np.sqrt(x * x + y * y + z * z)
    r = np.where(r > 0, r, 1.0)
    q = lambda t: np.clip(((t / r + 1.0) * 0.5), 0.0, 1.0)
    X = [(q(t) * 2097151.0).astype(np.uint32) for t in (x, y, z)]
    M = np.uint32(1 << 20)
    Q = M
    while Q > 1:
        P = np.uint32(Q - 1)
        for i in range(3):
            hit = (X[i] & Q) != 0
            t = (X[0] ^ X[i]) & P
            X0_new = np.where(hit, X[0] ^ P, X[0] ^ t)
            Xi_new = np.where(hit, X[i], X[i] ^ t)
            if i == 0:
                X[0] = np.where(hit, X[0] ^ P, X[0])          # t == 0 when i == 0
            else:
                X[0], X[i] = X0_new, Xi_new
        Q = np.uint32(Q >> 1)
    X[1] ^= X[0]
    X[2] ^= X[1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ np.uint32(Q - 1), t)
        Q = np.uint32(Q >> 1)
    for i in range(3):
        X[i] ^= t
    return (_spread21(X[0]) << np.uint64(2)) | (_spread21(X[1]) << np.uint64(1)) | _spread21(X[2])


def sfc_colouring(mesh: Mesh, n_parts: int) -> np.ndarray:
    """contiguous, equally sized chunks of the Hilbert order of the cell centres."""
    v = mesh.v
    order = np.argsort(hilbert_keys(v["xCell"], v["yCell"], v["zCell"]), kind="stable")
    col = np.empty(mesh.nCells, dtype=np.int32)
    bounds = (np.arange(n_parts + 1, dtype=np.int64) * mesh.nCells) // n_parts
    for c in range(n_parts):
        col[order[bounds[c]:bounds[c + 1]]] = c
    return col


# ------------------------------------------------------------------------------------------------
# partition_regions
@dataclass
class CellPartition:
    """cell_partition_fs (data_structures.rg:577-584) for every colour: lists of sorted cell indices."""
    n_parts: int
    p: List[np.ndarray]
    private_1: List[np.ndarray]
    shared_1: List[np.ndarray]
    ghost_1: List[np.ndarray]
    private_2: List[np.ndarray]
    shared_2: List[np.ndarray]
    ghost_2: List[np.ndarray]


def partition_regions(mesh: Mesh, colours: np.ndarray, n_parts: int, policy: int) -> CellPartition:
    """mesh_loading.rg:399-483 on cell sets.  ``image`` keeps only targets inside the region (ids that
    resolve to the pad are dropped, as Legion's image does); all ten edgesOnCell slots take part (:409-419)."""
    v = mesh.v
    nC, nE = mesh.nCells, mesh.nEdges
    eoc = resolve_ids(v["edgesOnCell"], nE, policy)           # [nC, 10], pad = nE
    c_one = resolve_ids(v["cellsOnEdge"][:, 0], nC, policy)   # cellOne / cellTwo (:433-436)
    c_two = resolve_ids(v["cellsOnEdge"][:, 1], nC, policy)
    out = CellPartition(n_parts, [], [], [], [], [], [], [])

    def cells_mask(idx):
        m = np.zeros(nC + 1, dtype=bool); m[idx] = True; m[nC] = False
        return m

    def edge_image(cmask):          # e = e0 | ... | e9  (:409-419)
        m = np.zeros(nE + 1, dtype=bool)
        m[eoc[cmask[:nC]].ravel()] = True
        m[nE] = False
        return m

    def reach(cmask):               # image(cellTwo) of preimage(cellOne) | image(cellOne) of preimage(cellTwo)
        m = np.zeros(nC + 1, dtype=bool)
        m[c_two[cmask[c_one]]] = True
        m[c_one[cmask[c_two]]] = True
        m[nC] = False
        return m

    for colour in range(n_parts):
        p = cells_mask(np.nonzero(colours == colour)[0])
        e = edge_image(p)
        cp = np.zeros(nC + 1, dtype=bool)                                   # cp_one | cp_two (:442-446)
        cp[c_one[e[:nE]]] = True; cp[c_two[e[:nE]]] = True; cp[nC] = False
        ghost_1 = cp & ~p                                                   # :448
        ghost_2 = reach(cp) & ~p                                            # :451-455
        shared_1 = p & reach(ghost_1)                                       # :458-462
        private_1 = p & ~shared_1                                           # :463
        shared_2 = shared_1 | (private_1 & reach(shared_1))                 # :466-470
        private_2 = private_1 & ~shared_2                                   # :471
        for name, m in (("p", p), ("ghost_1", ghost_1), ("ghost_2", ghost_2), ("shared_1", shared_1),
                        ("private_1", private_1), ("shared_2", shared_2), ("private_2", private_2)):
            getattr(out, name).append(np.nonzero(m[:nC])[0].astype(np.int32))
    return out


def is_shared(part: CellPartition, n_cells: int) -> np.ndarray:
    """fill(isShared,false); mark_shared_cells(shared_1[i]); mark_shared_cells(shared_2[i])  (main.rg:48-52)."""
    out = np.zeros(n_cells, dtype=np.uint8)
    for i in range(part.n_parts):
        out[part.shared_1[i]] = 1
        out[part.shared_2[i]] = 1
    return out


# ------------------------------------------------------------------------------------------------
# local meshes + halo lists
@dataclass
class LocalMesh:
    rank: int
    n_owned: Tuple[int, int, int]                     # owned cells / edges / vertices (they come first)
    cells: np.ndarray                                 # global cell index of every local cell
    edges: np.ndarray
    vertices: np.ndarray
    owner: Dict[str, np.ndarray]                      # per entity type: owning rank of every local entity
    send: Dict[str, Dict[int, np.ndarray]] = field(default_factory=dict)   # entity -> peer -> local indices (owned here)
    recv: Dict[str, Dict[int, np.ndarray]] = field(default_factory=dict)   # entity -> peer -> local indices (ghost here)
    interior_cells: np.ndarray = None                 # private_2: owned cells whose 2-ring is owned (overlap window)


_CELL_ROWS_1D = ("nEdgesOnCell", "latCell", "invAreaCell", "bdyMaskCell", "specZoneMaskCell", "isShared", "inCpr",
                 "xCell", "yCell", "zCell", "lonCell")
_CELL_ROWS_ME = ("kiteForCell", "edgesOnCellSign", "edgesOnCell_sign", "defc_a", "defc_b", "coeffs_reconstruct")
_EDGE_ROWS = ("nEdgesOnEdge", "weightsOnEdge", "dcEdge", "dvEdge", "invDcEdge", "invDvEdge", "angleEdge", "latEdge",
              "nAdvCellsForEdge", "adv_coefs", "adv_coefs_3rd", "meshScalingDel2", "meshScalingDel4", "specZoneMaskEdge")
_VERTEX_ROWS = ("edgesOnVertexSign", "edgesOnVertex_sign", "kiteAreasOnVertex", "fVertex", "invAreaTriangle")


def _remap(raw: np.ndarray, n_global: int, g2l: np.ndarray) -> np.ndarray:
    """1-based global ids -> 1-based local ids, 0 where the target is not held locally (CORRECTED policy)."""
    idx = resolve_ids(raw, n_global, CORRECTED)                # 0-based, pad = n_global
    return g2l[idx].astype(np.int32)                           # g2l[pad] = 0


def build_local(static: Dict[str, np.ndarray], n_global: Tuple[int, int, int], colours: np.ndarray, part: CellPartition,
                rank: int, cells_on_vertex: np.ndarray) -> Tuple[LocalMesh, Dict[str, np.ndarray]]:
    """The mesh (MpasMeshPtrs members, 1-based LOCAL ids) of one rank under the CORRECTED index policy.

    ``static`` is the global level-0 data as handed to ``upload_mesh`` (raw 1-based global ids)."""
    nC, nE, nV = n_global
    owned = part.p[rank]
    g1 = part.ghost_1[rank]
    g2 = np.setdiff1d(part.ghost_2[rank], g1, assume_unique=True)
    cells = np.concatenate([owned, g1, g2]).astype(np.int64)
    eoc = resolve_ids(static["edgesOnCell"], nE, CORRECTED)
    voc = resolve_ids(static["verticesOnCell"], nV, CORRECTED)
    slot = np.arange(eoc.shape[1])[None, :] < static["nEdgesOnCell"][:, None]
    e_all = np.unique(eoc[cells][slot[cells]]); e_all = e_all[e_all < nE]
    v_all = np.unique(voc[cells][slot[cells]]); v_all = v_all[v_all < nV]
    c1g = resolve_ids(static["cellsOnEdge"][:, 0], nC, CORRECTED)
    cvg = resolve_ids(cells_on_vertex[:, 0], nC, CORRECTED)
    col_p = np.concatenate([colours, [-1]])
    e_owner = col_p[c1g[e_all]]
    v_owner = col_p[cvg[v_all]]
    # owned first, then ghosts grouped by owner (keeps the per-peer recv runs contiguous)
    e_ord = np.lexsort((e_all, np.where(e_owner == rank, -1, e_owner)))
    v_ord = np.lexsort((v_all, np.where(v_owner == rank, -1, v_owner)))
    edges, e_owner = e_all[e_ord], e_owner[e_ord]
    vertices, v_owner = v_all[v_ord], v_owner[v_ord]
    c_owner = colours[cells]
    lm = LocalMesh(rank=rank, n_owned=(len(owned), int((e_owner == rank).sum()), int((v_owner == rank).sum())),
                   cells=cells, edges=edges, vertices=vertices, owner=dict(cell=c_owner, edge=e_owner, vertex=v_owner))
    g2l_c = np.zeros(nC + 1, dtype=np.int64); g2l_c[cells] = np.arange(1, len(cells) + 1)
    g2l_e = np.zeros(nE + 1, dtype=np.int64); g2l_e[edges] = np.arange(1, len(edges) + 1)
    g2l_v = np.zeros(nV + 1, dtype=np.int64); g2l_v[vertices] = np.arange(1, len(vertices) + 1)
    loc: Dict[str, np.ndarray] = {}
    for k in _CELL_ROWS_1D + _CELL_ROWS_ME:
        if static.get(k) is not None:
            loc[k] = np.ascontiguousarray(static[k][cells])
    for k in _EDGE_ROWS:
        if static.get(k) is not None:
            loc[k] = np.ascontiguousarray(static[k][edges])
    for k in _VERTEX_ROWS:
        if static.get(k) is not None:
            loc[k] = np.ascontiguousarray(static[k][vertices])
    loc["edgesOnCell"] = _remap(static["edgesOnCell"][cells], nE, g2l_e)
    loc["verticesOnCell"] = _remap(static["verticesOnCell"][cells], nV, g2l_v)
    loc["cellsOnEdge"] = _remap(static["cellsOnEdge"][edges], nC, g2l_c)
    loc["verticesOnEdge"] = _remap(static["verticesOnEdge"][edges], nV, g2l_v)
    loc["edgesOnEdge_ECP"] = _remap(static["edgesOnEdge_ECP"][edges], nE, g2l_e)
    if static.get("edgesOnEdge") is not None:
        loc["edgesOnEdge"] = _remap(static["edgesOnEdge"][edges], nE, g2l_e)
    loc["advCellsForEdge"] = _remap(static["advCellsForEdge"][edges], nC, g2l_c)
    loc["edgesOnVertex"] = _remap(static["edgesOnVertex"][vertices], nE, g2l_e)
    # slots beyond nEdgesOnCell / nEdgesOnEdge / nAdvCellsForEdge are never read by the kernels; zero them
    ne = loc["nEdgesOnCell"]
    for k in ("edgesOnCell", "verticesOnCell"):
        loc[k][np.arange(loc[k].shape[1])[None, :] >= ne[:, None]] = 0
    lm.interior_cells = (g2l_c[part.private_2[rank]] - 1).astype(np.int32)
    return lm, loc


def build_halo_lists(locals_: List[LocalMesh]) -> None:
    """Fill send / recv lists of every rank.  For a pair (owner o, holder h) and an entity type, both sides
    list the same global entities in ascending global order."""
    for ent, attr in (("cell", "cells"), ("edge", "edges"), ("vertex", "vertices")):
        glob = [getattr(lm, attr) for lm in locals_]
        pos = []
        for lm, g in zip(locals_, glob):
            order = np.argsort(g, kind="stable")
            pos.append((g[order], order))
        for lm in locals_:
            lm.send.setdefault(ent, {}); lm.recv.setdefault(ent, {})
        for h, lm in enumerate(locals_):
            own = lm.owner[ent]
            for o in np.unique(own):
                o = int(o)
                if o == h or o < 0:
                    continue
                ghost_local = np.nonzero(own == o)[0]
                g_ids = glob[h][ghost_local]
                srt = np.argsort(g_ids, kind="stable")
                g_ids, ghost_local = g_ids[srt], ghost_local[srt]
                sg, so = pos[o]
                where = np.searchsorted(sg, g_ids)
                assert np.array_equal(sg[where], g_ids), "ghost entity not held by its owner"
                lm.recv[ent][o] = ghost_local.astype(np.int32)
                locals_[o].send[ent][h] = so[where].astype(np.int32)
