"""SURVEY.md 8f rank 4 against the CUDA library (tests/test_output.py runs the same host code on the oracle backend): the
checkpoint / restart path and the output diagnostics through ``Dynamics`` -- download_field / upload_field of the device mirror."""
import numpy as np
import pytest

from mpas_regent_b200 import _abi, output
from tests.util import build_pair

pytestmark = pytest.mark.gpu
L = 6


@pytest.mark.parametrize("physics", [_abi.PHYSICS_LITERAL, _abi.PHYSICS_CORRECTED], ids=["literal", "corrected_physics"])
def test_checkpoint_restart_is_bit_identical_on_the_device(grid642, tmp_path, physics):
    st, ora, a = build_pair(grid642, L, _abi.INDEX_CORRECTED, physics_mode=physics, config_scalar_advection=1)
    ora.close()
    a.upload_field("scalars", 1e-3 * (1.0 + np.random.default_rng(4).random((grid642.nCells, L + 1, 8))))
    a.atm_compute_solve_diagnostics(False, -1)
    a.atm_srk3(600.0)
    output.save_checkpoint(a, str(tmp_path / "ck"), step=1)
    a.atm_srk3(600.0)
    st, ora, b = build_pair(grid642, L, _abi.INDEX_CORRECTED, physics_mode=physics, config_scalar_advection=1)   # fresh handle, same mesh
    ora.close()
    assert output.load_checkpoint(b, str(tmp_path / "ck")) == 1
    b.atm_srk3(600.0)
    for (n, _, _) in _abi.FIELDS:
        assert np.array_equal(a.download_field(n), b.download_field(n), equal_nan=True), n
    a.close(); b.close()


def test_output_diagnostics_and_plotting_file_from_the_device(grid642, tmp_path):
    from scipy.io import netcdf_file
    st, ora, g = build_pair(grid642, L, _abi.INDEX_CORRECTED)
    for b in (ora, g):
        b.atm_compute_solve_diagnostics(False, -1)
        b.atm_srk3(600.0)
    names = ("rho_zz", "zz", "pressure_p", "u", "v", "w")
    f = {n: g.download_field(n) for n in names}
    fo = {n: ora.download_field(n) for n in names}
    d, do = output.atm_compute_output_diagnostics(f), output.atm_compute_output_diagnostics(fo)      # dynamics_tasks.rg:729-744
    for k in d:
        assert np.array_equal(d[k], do[k]), k                     # every input field is bit-identical to the oracle's
    path = str(tmp_path / "out.nc")
    output.write_output_plotting(path, st.mesh, {**f, **d})
    nc = netcdf_file(path, "r", mmap=False)
    assert np.array_equal(nc.variables["u"][:], fo["u"][:, 0]) and np.array_equal(nc.variables["rho"][:], do["rho"][:, 0])
    nc.close()
    g.close(); ora.close()
