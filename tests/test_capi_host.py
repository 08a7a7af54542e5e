"""CPU: the C-ABI library loads here (no GPU), exports every symbol of include/mdb200.h, and its pure-host
entry points (letterbox geometry) agree with the oracle.  No compute calls."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from megadetector_b200 import _capi
from oracle import letterbox as olb


def _declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'mdb200.h')).read()
    return sorted(set(re.findall(r'MDB_API [^;(]*?\b(mdb_[a-z0-9_]+)\(', hdr)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load_library()
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), 'libmdb200.so does not export {}'.format(n)
    assert sorted(_capi.EXPORTS) == names
    assert b'sm_100a' in lib.mdb_version()


def test_config_struct_layout_matches_header():
    assert C.sizeof(_capi.MdbArch) == 4 * (6 + 5 + 1 + 1) + 4 * 24
    assert C.sizeof(_capi.MdbConfig) == 4 * 11 + C.sizeof(_capi.MdbArch)  # + compat_mode, use_ceil_for_resize


@pytest.mark.parametrize('hw', [(1080, 1920), (450, 600), (450, 800), (600, 800), (1281, 1919), (333, 77), (64, 64),
                                (1280, 1280), (1280, 960), (2160, 3840), (3000, 4000), (1, 1), (5000, 3), (719, 1279)])
def test_letterbox_geometry_matches_oracle(hw):
    cfg = _capi.make_config()
    g = _capi.letterbox_geometry(cfg, hw[0], hw[1])
    o = olb.letterbox_params(hw[0], hw[1])
    assert tuple(g['new_unpad']) == tuple(o['new_unpad'])
    assert (g['top'], g['bottom'], g['left'], g['right']) == (o['top'], o['bottom'], o['left'], o['right'])
    assert tuple(g['out_hw']) == tuple(o['out_hw'])
    assert g['ratio'][0] == o['ratio'][0] and tuple(g['pad']) == tuple(float(x) for x in o['pad'])


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError, match='no CUDA device'):
        _capi.Engine(_capi.make_config())
    from megadetector_b200.detector import B200Detector
    with pytest.raises(Exception):
        B200Detector('synthetic://tiny?seed=0', {'max_batch': 1})
