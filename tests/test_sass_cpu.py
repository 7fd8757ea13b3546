"""CPU: static check of the built library's machine code (cuobjdump -sass, tools/sass_histogram.py) - the hot kernels are written
for Blackwell's tensor cores and copy engines, not recompiled legacy paths: tcgen05 MMAs (UTCHMMA, the 2-CTA form in the pair
kernel), tensor-memory loads (LDTM), bulk copies (UBLKCP), multicast commits, no mma.sync (HMMA) anywhere; and the fused
kernel's MMA issue block holds all eight MMAs of a staged k-block in one straight-line run (the v3 issue loop)."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))


@pytest.fixture(scope='module')
def table(built_lib):
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not on PATH')
    import sass_histogram as sh
    lib = os.path.join(ROOT, 'diffdock_b200', 'libdiffdock_b200.so')
    return {name: (dict(zip(sh.COLS, counts)), block, var) for name, _, counts, block, var in sh.rows(lib)}


def test_no_legacy_tensor_core_instructions(table):
    assert table and all(c['HMMA'] == 0 for c, _, _ in table.values())


def test_fused_kernel_is_tcgen05_native(table):
    pair = next(v for k, v in table.items() if 'fused_conv_kernel<2, 0>' in k)
    single = next(v for k, v in table.items() if 'fused_conv_kernel<1, 0>' in k)
    for counts, block, var in (pair, single):
        assert counts['UTCHMMA'] >= 8 and counts['LDTM'] > 0 and counts['UBLKCP'] > 0 and counts['SYNCS'] > 0 and counts['REDG'] > 0
        assert block == 8, block                      # one issue block per staged k-block
    assert 'UTCHMMA.2CTA' in pair[2] and 'UTCBAR.2CTA.MULTICAST' in pair[2]


def test_streaming_and_gemm_kernels_use_bulk_copies(table):
    tp = next(v for k, v in table.items() if 'tpconv_accumulate_kernel' in k)
    assert tp[0]['UBLKCP'] > 0 and tp[0]['REDG'] > 0 and tp[0]['UTCHMMA'] == 0      # HBM-bound: no tensor cores by design
    gemm = [v for k, v in table.items() if 'radial_gemm_kernel' in k]
    assert gemm and all(c['UTCHMMA'] > 0 and c['LDTM'] > 0 for c, _, _ in gemm)
