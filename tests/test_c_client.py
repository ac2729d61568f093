"""A plain-C program against include/ce2e.h + libce2e.so (tests/c_abi/kat.c): the boundary of
SURVEY 8b has no Python, torch or C++ types in it.  CPU: the header is valid C99 and the client
links.  GPU: the client's hand-derivable known answers hold."""
import os
import subprocess

import pytest

from env_build_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'tests', 'c_abi', 'kat.c')
CUDA = os.environ.get('CUDA_HOME', '/usr/local/cuda')


def _build(tmp_path):
    if _lib.needs_build():
        _lib.build()
    exe = str(tmp_path / 'kat')
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'),
                    '-isystem', os.path.join(CUDA, 'include'), SRC, '-o', exe, '-L', libdir,
                    '-l:' + os.path.basename(_lib.LIB_PATH), '-L', os.path.join(CUDA, 'lib64'), '-lcudart', '-lm',
                    '-Wl,-rpath,' + libdir, '-Wl,-rpath,' + os.path.join(CUDA, 'lib64')], check=True)
    return exe


def test_header_is_plain_c_and_client_links(tmp_path):
    probe = tmp_path / 'probe.c'
    probe.write_text('#include "ce2e.h"\nint main(void) { return sizeof(ce2e_turn_classes) == CE2E_MAX_VEH ? 0 : 1; }\n')
    subprocess.run(['gcc', '-std=c89', '-pedantic', '-Wall', '-Werror', '-fsyntax-only',
                    '-I', os.path.join(ROOT, 'include'), str(probe)], check=True)
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_plain_c_client_known_answers(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
    out = r.stdout.decode()
    assert r.returncode == 0 and 'ALL OK' in out, out
    assert 'launches ' in out and int(out.split('launches ')[1].split()[0]) >= 6
