"""Mesh container + grid-file reader for the RK3 dynamics hot path.

Mirrors what ``load_mesh`` puts into the level-0 fields of the cell / edge / vertex
regions (reference: mesh_loading/mesh_loading.rg:27-390): the 38 NetCDF variables of an
MPAS ``grid.nc`` file, ids kept exactly as stored (1-based, unit sphere), plus the METIS
colouring read by ``read_file`` (mesh_loading.rg:11-22).

Nothing here touches the GPU.  The index policy (LITERAL / CORRECTED, SURVEY.md 8c rule
M2) is applied by :func:`resolve_ids` and, on the device, by ``mpasb200_upload_mesh``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

LITERAL = 0
CORRECTED = 1

MAX_EDGES = 10      # constants.rg:21
MAX_EDGES2 = 20     # constants.rg:22
VERTEX_DEGREE = 3   # constants.rg:25
FIFTEEN = 15        # constants.rg:24

#: the variables load_mesh reads (mesh_loading.rg:123-161)
GRID_VARS = (
    "latCell", "lonCell", "meshDensity", "xCell", "yCell", "zCell", "indexToCellID",
    "latEdge", "lonEdge", "xEdge", "yEdge", "zEdge", "indexToEdgeID",
    "latVertex", "lonVertex", "xVertex", "yVertex", "zVertex", "indexToVertexID",
    "cellsOnEdge", "nEdgesOnCell", "nEdgesOnEdge", "edgesOnCell", "edgesOnEdge",
    "weightsOnEdge", "dvEdge", "dv1Edge", "dv2Edge", "dcEdge", "angleEdge", "areaCell",
    "areaTriangle", "cellsOnCell", "verticesOnCell", "verticesOnEdge", "edgesOnVertex",
    "cellsOnVertex", "kiteAreasOnVertex",
)
_INT_VARS = {
    "indexToCellID", "indexToEdgeID", "indexToVertexID", "cellsOnEdge", "nEdgesOnCell",
    "nEdgesOnEdge", "edgesOnCell", "edgesOnEdge", "cellsOnCell", "verticesOnCell",
    "verticesOnEdge", "edgesOnVertex", "cellsOnVertex",
}


@dataclass
class Mesh:
    """Raw grid-file content (ids 1-based as stored, lengths on the unit sphere)."""

    v: Dict[str, np.ndarray]
    partition: Optional[np.ndarray] = None   # cell colouring, one int per cell
    name: str = "mesh"
    extras: Dict[str, np.ndarray] = field(default_factory=dict)

    @property
    def nCells(self) -> int:
        return int(self.v["nEdgesOnCell"].shape[0])

    @property
    def nEdges(self) -> int:
        return int(self.v["cellsOnEdge"].shape[0])

    @property
    def nVertices(self) -> int:
        return int(self.v["edgesOnVertex"].shape[0])

    def __getitem__(self, k: str) -> np.ndarray:
        return self.v[k]

    def validate(self) -> None:
        nC, nE, nV = self.nCells, self.nEdges, self.nVertices
        assert self.v["edgesOnCell"].shape == (nC, MAX_EDGES)
        assert self.v["edgesOnEdge"].shape == (nE, MAX_EDGES2)
        assert self.v["cellsOnVertex"].shape == (nV, VERTEX_DEGREE)
        assert nE == 3 * nC - 6 and nV == 2 * nC - 4, "not a closed spherical Voronoi mesh"
        for k in GRID_VARS:
            a = self.v[k]
            want = np.int32 if k in _INT_VARS else np.float64
            assert a.dtype == want, (k, a.dtype)
            assert a.flags["C_CONTIGUOUS"], k


def _canon(name: str, a: np.ndarray) -> np.ndarray:
    dt = np.int32 if name in _INT_VARS else np.float64
    return np.ascontiguousarray(np.asarray(a).astype(dt, copy=False))


def read_grid_netcdf(path: str, graph_path: Optional[str] = None, name: Optional[str] = None) -> Mesh:
    """Read an MPAS grid.nc (NetCDF-3 classic; the reference links libnetcdf, main.rg:13).

    scipy's pure-python NetCDF-3 reader is enough for the classic format of the bundled
    ``x1.2562.grid.nc``; no libnetcdf is needed.
    """
    from scipy.io import netcdf_file

    with netcdf_file(path, "r", mmap=False) as f:
        v = {k: _canon(k, f.variables[k].data) for k in GRID_VARS}
    part = read_graph_partition(graph_path, v["nEdgesOnCell"].shape[0]) if graph_path else None
    m = Mesh(v=v, partition=part, name=name or path)
    m.validate()
    return m


def read_graph_partition(path: str, n_cells: int) -> np.ndarray:
    """``read_file`` (mesh_loading.rg:11-22): one colour per line, atoi of each line."""
    out = np.zeros(n_cells, dtype=np.int32)
    with open(path, "r") as fh:
        for i, line in enumerate(fh):
            if i >= n_cells:
                break
            s = line.strip()
            out[i] = int(s) if s else 0
    return out


def save_npz(mesh: Mesh, path: str) -> None:
    d = dict(mesh.v)
    if mesh.partition is not None:
        d["__partition__"] = mesh.partition.astype(np.int32)
    np.savez_compressed(path, **d)


def load_npz(path: str, name: Optional[str] = None) -> Mesh:
    z = np.load(path)
    v = {k: _canon(k, z[k]) for k in GRID_VARS}
    part = np.ascontiguousarray(z["__partition__"].astype(np.int32)) if "__partition__" in z.files else None
    m = Mesh(v=v, partition=part, name=name or path)
    m.validate()
    return m


def resolve_ids(raw: np.ndarray, n: int, policy: int) -> np.ndarray:
    """Stored id -> 0-based array index with a pad entity at index ``n`` (rule M2).

    LITERAL: the stored 1-based id is used as the index (what the reference's dynamics
    does, e.g. dynamics_tasks.rg:347-350); id == n hits the pad.  CORRECTED: id-1, and
    id 0 (no neighbour) hits the pad.  Anything else out of range also goes to the pad.
    """
    raw = np.asarray(raw, dtype=np.int64)
    idx = raw.copy() if policy == LITERAL else raw - 1
    idx[(idx < 0) | (idx > n)] = n
    return idx.astype(np.int32)
