"""Host-side output and restart: what lies after the hot path in main.rg:70-72.

* ``atm_compute_output_diagnostics``  dynamics_tasks.rg:729-744  (rho, pressure; the theta line is commented out, :739-740)
* ``write_output_plotting``           mesh_loading.rg:810-1125   NetCDF-3 file with level 0 of the eight plotted fields and
  the mesh variables plotting/mpas_patches.py reads (nEdgesOnCell, verticesOnCell, latVertex, lonVertex, ...)
* ``save_checkpoint`` / ``load_checkpoint``  a raw structure-of-arrays image of every field of the device mirror, so a
  long run can be stopped and resumed bit-identically (the reference has no restart path).

Everything here goes through ``TaskAPI.download_field`` / ``upload_field`` only; nothing is on the timed path.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, Optional

import numpy as np

from . import _abi
from .dynamics import TaskAPI
from .mesh import Mesh

#: the eight fields write_output_plotting stores (mesh_loading.rg:946-955), with the entity whose level 0 is written
PLOTTED = (("u", "nEdges"), ("v", "nEdges"), ("w", "nCells"), ("pressure", "nCells"), ("pressure_p", "nCells"),
           ("rho", "nCells"), ("theta", "nCells"), ("surface_pressure", "nCells"))


def atm_compute_output_diagnostics(f: Dict[str, np.ndarray], pressure_base: Optional[np.ndarray] = None,
                                   theta: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """dynamics_tasks.rg:729-744 on levels 0..nVertLevels-1: ``rho = rho_zz * zz``, ``pressure = pressure_base + pressure_p``.
    ``theta`` is declared written but its assignment is commented out (:739-740), so it keeps its previous value (zero for
    a never-written field, memory-model rule M1).  ``pressure_base`` is not a hot-path field; None means never written (0)."""
    L = f["rho_zz"].shape[1] - 1
    rho = np.zeros_like(f["rho_zz"])
    pressure = np.zeros_like(f["pressure_p"])
    pb = np.zeros_like(pressure) if pressure_base is None else pressure_base
    rho[:, :L] = f["rho_zz"][:, :L] * f["zz"][:, :L]
    pressure[:, :L] = pb[:, :L] + f["pressure_p"][:, :L]
    return {"rho": rho, "pressure": pressure, "theta": np.zeros_like(rho) if theta is None else theta}


def write_output_plotting(path: str, mesh: Mesh, fields: Dict[str, np.ndarray]):
    """NetCDF-3 (classic) file in the layout of mesh_loading.rg:810-1125: the mesh variables and LEVEL 0 of the eight
    plotted fields as 1-D variables over nCells / nEdges (``cell_region[{i, 0}].pressure`` etc., :1024-1053)."""
    from scipy.io import netcdf_file
    v = mesh.v
    nc = netcdf_file(path, "w", version=1)
    try:
        dims = {"nCells": mesh.nCells, "nEdges": mesh.nEdges, "nVertices": mesh.nVertices,
                "maxEdges": v["edgesOnCell"].shape[1], "maxEdges2": v["edgesOnEdge"].shape[1], "TWO": 2,
                "vertexDegree": v["edgesOnVertex"].shape[1]}
        for k, n in dims.items():
            nc.createDimension(k, int(n))
        shapes = {(dims["nCells"],): ("nCells",), (dims["nEdges"],): ("nEdges",), (dims["nVertices"],): ("nVertices",),
                  (dims["nCells"], dims["maxEdges"]): ("nCells", "maxEdges"), (dims["nEdges"], 2): ("nEdges", "TWO"),
                  (dims["nEdges"], dims["maxEdges2"]): ("nEdges", "maxEdges2"),
                  (dims["nVertices"], dims["vertexDegree"]): ("nVertices", "vertexDegree")}
        mesh_vars = ("latCell", "lonCell", "xCell", "yCell", "zCell", "indexToCellID", "nEdgesOnCell", "areaCell", "edgesOnCell",
                     "verticesOnCell", "cellsOnCell", "latEdge", "lonEdge", "xEdge", "yEdge", "zEdge", "indexToEdgeID",
                     "nEdgesOnEdge", "dvEdge", "dcEdge", "angleEdge", "cellsOnEdge", "verticesOnEdge", "edgesOnEdge",
                     "weightsOnEdge", "latVertex", "lonVertex", "xVertex", "yVertex", "zVertex", "indexToVertexID",
                     "areaTriangle", "edgesOnVertex", "cellsOnVertex", "kiteAreasOnVertex")
        for name in mesh_vars:
            if name not in v:
                continue
            a = np.asarray(v[name])
            dn = shapes.get(a.shape)
            if dn is None:
                continue
            integer = np.issubdtype(a.dtype, np.integer)
            var = nc.createVariable(name, "i" if integer else "d", dn)
            var[:] = a.astype(np.int32 if integer else np.float64)
        for name, dim in PLOTTED:
            a = fields.get(name)
            var = nc.createVariable(name, "d", (dim,))
            var[:] = np.zeros(dims[dim]) if a is None else np.asarray(a, dtype=np.float64).reshape(dims[dim], -1)[:, 0]
    finally:
        nc.close()


# ---- raw SoA checkpoint ----------------------------------------------------------------------------------------------
def save_checkpoint(backend: TaskAPI, path: str, step: int = 0, names: Optional[Iterable[str]] = None):
    """every field of the mirror (include/mpas_b200_fields.def), one .npy each, plus dims / config / step in meta.json."""
    os.makedirs(path, exist_ok=True)
    names = list(names) if names is not None else [n for (n, _, _) in _abi.FIELDS]
    for n in names:
        np.save(os.path.join(path, n + ".npy"), backend.download_field(n))
    d, c = backend.dims, backend.cfg
    meta = {"step": int(step), "fields": names,
            "dims": {k: int(getattr(d, k)) for k, _ in d._fields_},
            "config": {k: (float(getattr(c, k)) if t is _abi.C.c_double else int(getattr(c, k))) for k, t in c._fields_}}
    with open(os.path.join(path, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


def load_checkpoint(backend: TaskAPI, path: str) -> int:
    """restore a checkpoint into a backend that already has its mesh uploaded; returns the step counter."""
    with open(os.path.join(path, "meta.json")) as fh:
        meta = json.load(fh)
    d = backend.dims
    for k, want in meta["dims"].items():
        if int(getattr(d, k)) != want:
            raise ValueError(f"checkpoint {k} = {want}, backend has {getattr(d, k)}")
    for n in meta["fields"]:
        backend.upload_field(n, np.load(os.path.join(path, n + ".npy")))
    return int(meta["step"])
