"""Regenerates tests/golden/x1.2562.grid.npz from the reference's bundled grid + METIS file.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py
The fixture is DATA (the 38 NetCDF variables load_mesh reads, mesh_loading.rg:123-201, plus the
16-way colouring of x1.2562.graph.info.part.16), not reference source.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mpas_regent_b200 import mesh as M  # noqa: E402

REF = "/root/reference/mesh_loading"
m = M.read_grid_netcdf(os.path.join(REF, "x1.2562.grid.nc"), os.path.join(REF, "x1.2562.graph.info.part.16"), name="x1.2562")
out = os.path.join(HERE, "x1.2562.grid.npz")
M.save_npz(m, out)
print(out, os.path.getsize(out), "bytes")
