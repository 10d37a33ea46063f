"""The C-ABI shared library loads without a GPU and exports every symbol include/ssb.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from helpers import REPO

from semiseg_b200 import _lib

HEADER = os.path.join(REPO, "include", "ssb.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_bound_symbols():
    decl = declared_symbols()
    assert len(decl) >= 25
    assert sorted(_lib.SIGNATURES) == decl, (set(decl) ^ set(_lib.SIGNATURES))


def test_library_loads_and_exports_all_symbols():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.ssb_version() == 1
    assert isinstance(lib.ssb_launch_count(), int)


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors == the C structs of include/ssb.h, as the C compiler lays them out (sizes and key offsets)."""
    import shutil
    import subprocess
    assert ctypes.sizeof(_lib.Geom) == 16
    assert ctypes.sizeof(_lib.StepParams) == 64
    assert ctypes.sizeof(_lib.BN) == 128
    assert _lib.StepParams.conf_thresh.offset == 36 and _lib.StepParams.grad_scale.offset == 32
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler to probe the header with")
    src = tmp_path / "probe.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
typedef void* ssb_stream_t_probe;
#include "ssb.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ssb_geom), sizeof(ssb_step_params), sizeof(ssb_bn), sizeof(ssb_aug_op),
         offsetof(ssb_bn, count_mul), offsetof(ssb_bn, sync_peers), offsetof(ssb_bn, sync_slot), offsetof(ssb_bn, sync_bwd_off));
  return 0;
}
''')
    exe = tmp_path / "probe"
    subprocess.run([cc, "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.Geom), ctypes.sizeof(_lib.StepParams), ctypes.sizeof(_lib.BN), ctypes.sizeof(_lib.AugOp),
            _lib.BN.count_mul.offset, _lib.BN.sync_peers.offset, _lib.BN.sync_slot.offset, _lib.BN.sync_bwd_off.offset]
    assert got == want, (got, want)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "nope.so"))


def test_no_cpu_fallback_in_product_path():
    """The product package never imports the oracle."""
    src_root = os.path.join(REPO, "semi-seg-ecg_b200", "src")
    for dirpath, _, files in os.walk(src_root):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)
