'C-ABI library: loads on a CPU-only box and exports every symbol include/olfnav_b200.h declares.'
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, 'include', 'olfnav_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(olfnav_[a-z_0-9]+)\s*\(', text)))


def test_header_declares_the_operation_set():
    names = _declared()
    for required in ['olfnav_model_create', 'olfnav_belief_step', 'olfnav_evaluate', 'olfnav_infotaxis_choose',
                     'olfnav_backup_gamma', 'olfnav_backup_select', 'olfnav_backup_assemble', 'olfnav_vi_sweep',
                     'olfnav_env_step', 'olfnav_copy_rows']:
        assert required in names


def test_library_exports_every_declared_symbol():
    from olfactory_navigation_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
        assert name in _lib.SIGNATURES, f'{name} has no ctypes signature'
    assert set(_lib.SIGNATURES) == set(_declared())


def test_no_compute_free_entry_points():
    from olfactory_navigation_b200 import _lib
    assert _lib.lib.olfnav_abi_version() == 1
    assert _lib.lib.olfnav_padded_width(371) == 384
    assert _lib.lib.olfnav_padded_states(83, 371) == 83 * 384          # wide grids: rows padded to 64 floats (whole cache lines)
    assert _lib.lib.olfnav_padded_width_for(371, 4) == 372 and _lib.lib.olfnav_padded_width_for(21, 64) == 64
    assert _lib.lib.olfnav_padded_width(50) == 52


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'olfactory-navigation_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.sh')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+(oracle|pomdp_toolkit)\b', src, flags=re.M), f'{f} imports the oracle'
                assert '/root/reference' not in src, f'{f} reads the reference tree'
