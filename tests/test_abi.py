"""CPU tests: the C-ABI library loads, exports every symbol include/obe_b200.h declares, and
refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'obe_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(obe_[a-z_0-9]+)\s*\(', text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ['obe_update', 'obe_resample_systematic', 'obe_utility', 'obe_draw', 'obe_cdf', 'obe_search',
                 'obe_gather_jitter', 'obe_pick', 'obe_model_builtin', 'obe_model_compile', 'obe_refresh']:
        assert must in syms


def test_library_exports_every_declared_symbol(obe_lib):
    from optbayesexpt_b200 import _lib
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f'{name} declared in include/obe_b200.h but not exported'
    # and the ctypes table binds exactly that surface
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_constants_agree(obe_lib):
    from optbayesexpt_b200 import _lib
    text = open(os.path.join(ROOT, 'include', 'obe_b200.h')).read()
    assert int(re.search(r'#define OBE_TILE_SIZE (\d+)', text).group(1)) == _lib.TILE
    assert int(re.search(r'#define OBE_STATS_DOUBLES (\d+)', text).group(1)) == _lib.STATS_LEN
    assert obe_lib.obe_num_tiles(1) == 1
    assert obe_lib.obe_num_tiles(_lib.TILE) == 1
    assert obe_lib.obe_num_tiles(_lib.TILE + 1) == 2
    assert obe_lib.obe_num_tiles(10 ** 8) == 48829
    assert obe_lib.obe_scratch_bytes(10 ** 8) > 0


def test_model_handles_without_gpu(obe_lib):
    from optbayesexpt_b200 import _lib
    h = C.c_void_p()
    assert obe_lib.obe_model_builtin(b'lorentzian_hwhm', 3, C.byref(h)) == 0
    vals = [C.c_int() for _ in range(5)]
    assert obe_lib.obe_model_info(h, *[C.byref(v) for v in vals]) == 0
    assert [v.value for v in vals] == [1, 3, 1, 1, 3]
    obe_lib.obe_model_free(h)
    assert obe_lib.obe_model_builtin(b'no_such_model', 3, C.byref(h)) != 0
    assert b'unknown built-in' in obe_lib.obe_last_error()
    assert obe_lib.obe_model_builtin(b'line', 7, C.byref(h)) != 0


def test_no_cpu_fallback(obe_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('this check is for the GPU-less build container')
    import optbayesexpt_b200 as obe
    from optbayesexpt_b200._lib import ObeError
    with pytest.raises(ObeError):
        obe.ParticlePDF([[0.0, 1.0, 2.0, 3.0], [1.0, 3.0, 2.0, 4.0]])
    with pytest.raises(TypeError):
        obe.OptBayesExpt(lambda s, p, c: 0.0, (), [[0.0, 1.0]], ())


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'optbayesexpt_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_cycle_struct_layout_matches_the_header(obe_lib, tmp_path):
    """obe_cycle_t is plain data shared between C and ctypes: sizes and offsets must agree with the C compiler's."""
    import subprocess
    from optbayesexpt_b200 import _lib
    fields = [name for name, _ in _lib.Cycle._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "obe_b200.h"', 'int main(void) {',
            '  printf("%zu\\n", sizeof(obe_cycle_t));']
    prog += [f'  printf("%zu\\n", offsetof(obe_cycle_t, {f}));' for f in fields]
    prog += ['  printf("%zu\\n", sizeof(obe_cloud_t));', '  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(prog))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), '-o', str(exe), str(src)])
    nums = [int(v) for v in subprocess.check_output([str(exe)]).decode().split()]
    assert nums[0] == C.sizeof(_lib.Cycle)
    assert nums[1:-1] == [getattr(_lib.Cycle, f).offset for f in fields]
    assert nums[-1] == C.sizeof(_lib.Cloud)
