"""Output diagnostics, the NetCDF-3 plotting file and the raw checkpoint -- host side, CPU only."""
import numpy as np

from mpas_regent_b200 import _abi, output
from tests.util import build_pair

L = 6


def test_checkpoint_restart_is_bit_identical(grid642, tmp_path):
    st, a, _ = build_pair(grid642, L, _abi.INDEX_CORRECTED, gpu=False)
    a.atm_compute_solve_diagnostics(False, -1)
    for _ in range(2):
        a.atm_srk3(600.0)
    output.save_checkpoint(a, str(tmp_path / "ck"), step=2)
    for _ in range(2):
        a.atm_srk3(600.0)
    st, b, _ = build_pair(grid642, L, _abi.INDEX_CORRECTED, gpu=False)     # fresh backend, same mesh
    assert output.load_checkpoint(b, str(tmp_path / "ck")) == 2
    for _ in range(2):
        b.atm_srk3(600.0)
    for (n, _, _) in _abi.FIELDS:
        assert np.array_equal(a.download_field(n), b.download_field(n), equal_nan=True), n
    a.close(); b.close()


def test_checkpoint_rejects_other_dimensions(grid642, tmp_path):
    import pytest
    st, a, _ = build_pair(grid642, L, _abi.INDEX_CORRECTED, gpu=False)
    output.save_checkpoint(a, str(tmp_path / "ck"), names=["u", "w"])
    st, b, _ = build_pair(grid642, L + 1, _abi.INDEX_CORRECTED, gpu=False)
    with pytest.raises(ValueError):
        output.load_checkpoint(b, str(tmp_path / "ck"))
    a.close(); b.close()


def test_output_diagnostics_and_plotting_file(grid642, tmp_path):
    from scipy.io import netcdf_file
    st, a, _ = build_pair(grid642, L, _abi.INDEX_CORRECTED, gpu=False)
    a.atm_compute_solve_diagnostics(False, -1)
    a.atm_srk3(600.0)
    f = {n: a.download_field(n) for n in ("rho_zz", "zz", "pressure_p", "u", "v", "w")}
    pb = np.full_like(f["pressure_p"], 1.0e5)
    d = output.atm_compute_output_diagnostics(f, pressure_base=pb)
    assert np.array_equal(d["rho"][:, :L], f["rho_zz"][:, :L] * f["zz"][:, :L]) and np.all(d["rho"][:, L] == 0)
    assert np.array_equal(d["pressure"][:, :L], pb[:, :L] + f["pressure_p"][:, :L])
    assert not d["theta"].any()                                  # the assignment is commented out in the reference
    path = str(tmp_path / "out.nc")
    output.write_output_plotting(path, st.mesh, {**f, **d})
    nc = netcdf_file(path, "r", mmap=False)
    assert nc.dimensions["nCells"] == grid642.nCells and nc.dimensions["nEdges"] == grid642.nEdges
    for name, dim in output.PLOTTED:
        assert nc.variables[name].shape == (nc.dimensions[dim],)
    assert np.array_equal(nc.variables["u"][:], f["u"][:, 0]) and np.array_equal(nc.variables["rho"][:], d["rho"][:, 0])
    assert np.array_equal(nc.variables["verticesOnCell"][:], st.mesh.v["verticesOnCell"])     # what plotting/mpas_patches.py reads
    assert np.all(nc.variables["surface_pressure"][:] == 0)
    nc.close()
    a.close()
