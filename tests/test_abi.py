"""The C-ABI library loads and exports every symbol include/anqs_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, 'include', 'anqs_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(anqs_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported():
    from anqs_quantum_chemistry_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), 'libanqs_b200.so not built: run __graft_entry__.build()'
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for name in names:
        assert hasattr(handle, name), f'{name} declared in include/anqs_b200.h but not exported'


def test_python_binding_covers_header():
    from anqs_quantum_chemistry_b200 import _lib
    assert set(_declared()) == set(_lib.declared_symbols())
    lib = _lib.lib()
    assert lib.anqs_abi_version() == 1
    assert lib.anqs_hash_capacity(1000) == 2048
    assert lib.anqs_scan_workspace(10) >= 8


def test_no_cpu_fallback():
    import pytest
    import torch
    from anqs_quantum_chemistry_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.dptr(torch.zeros(4, dtype=torch.int64))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            _lib.require_cuda('cuda:0')
    with pytest.raises(RuntimeError):
        _lib.require_cuda('cpu')


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, 'anqs_quantum_chemistry_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'anqs_oracle' not in text, f


def test_header_is_plain_c_and_struct_layouts_match_the_python_mirrors(tmp_path):
    """include/anqs_b200.h compiles as C (the boundary a cgo / JNI / ctypes binding sees) and every struct that crosses it has
    the size and field offsets of its ctypes mirror in _lib.py."""
    import shutil
    import subprocess
    import pytest
    from anqs_quantum_chemistry_b200 import _lib
    if shutil.which('gcc') is None:
        pytest.skip('no C compiler')
    mirrors = {'anqs_made_desc_t': _lib.MadeDesc, 'anqs_nade_desc_t': _lib.NadeDesc,
               'anqs_transformer_desc_t': _lib.TransformerDesc, 'anqs_brg_problem_t': _lib.BrgProblem}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "anqs_b200.h"', 'int main(void) {']
    for cname, mirror in mirrors.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in mirror._fields_:
            lines.append(f'printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    lines += ['return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, mirror in mirrors.items():
        assert int(got[cname]) == ctypes.sizeof(mirror), cname
        for field, _ in mirror._fields_:
            assert int(got[f'{cname}.{field}']) == getattr(mirror, field).offset, f'{cname}.{field}'
