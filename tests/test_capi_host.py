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


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_halo_main_loop_addressing_tables(mode):
    """K2's 3x3 halo loop (csrc/conv_host.cu::halo_layout): model the shared-memory image the TMA boxes of one chunk leave
    behind and check that every MMA row of every filter tap -- descriptor start + group * SBO + row * 128 bytes -- reads
    exactly the input pixel the convolution needs (stride 1: in[y + kh - 1][x + kw - 1]; stride 2: in[2y + kh - 1][2x + kw - 1]
    through the parity views)."""
    from megadetector_b200 import _capi
    L = _capi.halo_layout(mode)
    stride = 2 if mode == 2 else 1
    tw, th = 8, 16                       # output tile: 8 px x 16 rows = the 128 MMA rows of one CTA
    ox0, oy0 = 5 * tw, 3 * th            # some tile origin (in output = view coordinates)
    smem = {}                            # byte offset of a 128-byte pixel row -> (parity view, view row, view col)
    total = 0
    for b in range(L['n_boxes']):
        assert L['box_off'][b] % 1024 == 0           # every box starts on a swizzle period
        nbytes = L['box_w'][b] * L['box_h'][b] * 128
        total += nbytes
        assert L['box_off'][b] + nbytes <= L['a_stage_bytes']
        for j in range(L['box_h'][b]):
            for i in range(L['box_w'][b]):
                addr = L['box_off'][b] + (j * L['box_w'][b] + i) * 128   # TMA writes the box densely, pixel-major rows
                assert addr not in smem                                 # boxes do not overlap
                smem[addr] = (L['box_par'][b], oy0 + L['box_dy'][b] + j, ox0 + L['box_dx'][b] + i)
    assert total == L['a_tx'] and L['a_stage_bytes'] % 1024 == 0
    for t in range(9):
        kh, kw = divmod(t, 3)
        assert L['tap_off'][t] % 128 == 0 and L['tap_sbo'][t] % 16 == 0
        for r in range(th):
            for x in range(tw):
                m = r * tw + x                                           # MMA row: group m // 8 = image row r, row in group = x
                addr = L['tap_off'][t] + (m // 8) * L['tap_sbo'][t] + (m % 8) * 128
                iy, ix = stride * (oy0 + r) + kh - 1, stride * (ox0 + x) + kw - 1
                if stride == 1:
                    want = (-1, iy, ix)
                else:
                    py, px = iy & 1, ix & 1
                    want = (py * 2 + px, (iy - py) // 2, (ix - px) // 2)
                assert smem.get(addr) == want, (mode, kh, kw, r, x)
    with pytest.raises(ValueError):
        _capi.halo_layout(3)
