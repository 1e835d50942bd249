"""The C-ABI library loads and exports every symbol include/mpas_b200.h declares; host-only
entry points behave; with no GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from mpas_regent_b200 import _abi, dynamics

HEADER = os.path.join(_abi.INCLUDE_DIR, "mpas_b200.h")


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mpasb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = dynamics.load_library()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n


def test_field_table_matches_def_file():
    lib = dynamics.load_library()
    for i, (name, ent, slots) in enumerate(_abi.FIELDS):
        e, s, nm = C.c_int(), C.c_int(), C.c_char_p()
        assert lib.mpasb200_field_info(i, C.byref(e), C.byref(s), C.byref(nm)) == 0
        assert (nm.value.decode(), e.value, s.value) == (name, ent, slots)
        assert lib.mpasb200_field_by_name(name.encode()) == i
    assert lib.mpasb200_field_info(len(_abi.FIELDS), None, None, None) == _abi.E_INVAL
    assert lib.mpasb200_field_by_name(b"no_such_field") == -1


def test_default_config_is_constants_rg():
    lib = dynamics.load_library()
    c = _abi.MpasConfig()
    lib.mpasb200_default_config(C.byref(c))
    py = _abi.default_config()
    for n, _ in _abi.MpasConfig._fields_:
        assert getattr(c, n) == getattr(py, n), n
    assert c.cp == 1004.5 and c.cv == 717.5 and c.config_len_disp == 120000.0 and c.number_of_sub_steps == 2


def test_create_rejects_bad_arguments():
    lib = dynamics.load_library()
    h = C.c_void_p()
    assert lib.mpasb200_create(None, None, C.byref(h)) == _abi.E_INVAL
    d = _abi.make_dims(10, 24, 16, 2)      # nVertLevels too small
    c = _abi.default_config()
    assert lib.mpasb200_create(C.byref(d), C.byref(c), C.byref(h)) == _abi.E_INVAL
    assert b"nVertLevels" in lib.mpasb200_last_error(None)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = _abi.make_dims(12, 30, 20, 5)
    with pytest.raises(dynamics.MpasB200Error, match="no CUDA device"):
        dynamics.Dynamics(d)


def test_product_package_never_imports_the_oracle():
    root = os.path.join(_abi.REPO_ROOT, "mpas_regent_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libmpas_oracle" not in txt, f


def test_dynamics_restrict_maps_a_launch_class_to_its_range():
    """Dynamics.restrict (what the overlapped multi-GPU schedule calls) = set_range(class_range(cls)), cached; None = everything."""
    from mpas_regent_b200 import dynamics
    d = object.__new__(dynamics.Dynamics)
    d._class_cache = {}
    calls, lookups = [], []
    d.class_range = lambda ent, cls: (lookups.append((ent, cls)), (10 * cls, 10 * cls + 5))[1]
    d.set_range = lambda ent, begin=-1, end=-1: calls.append((ent, begin, end))
    d.restrict(_abi.CELL, 1); d.restrict(_abi.CELL, 1); d.restrict(_abi.EDGE, 0); d.restrict(_abi.CELL); d.restrict(_abi.EDGE, None)
    assert calls == [(_abi.CELL, 10, 15), (_abi.CELL, 10, 15), (_abi.EDGE, 0, 5), (_abi.CELL, -1, -1), (_abi.EDGE, -1, -1)]
    assert lookups == [(_abi.CELL, 1), (_abi.EDGE, 0)]
