"""The C-ABI shared library builds, loads and exports every symbol that
include/tnf.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import subprocess

from torch_nf_b200 import _build, _lib


def test_library_builds_and_exports_header_symbols():
    path = _build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    declared = _lib.header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(handle, name), "missing C-ABI symbol %s" % name
    assert set(declared) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with include/tnf.h"
    handle.tnf_abi_version.restype = ctypes.c_int
    assert handle.tnf_abi_version() == 3


def test_no_torch_types_in_abi():
    """Signatures are plain C: the header compiles as C99 on its own."""
    hdr = os.path.join(os.path.dirname(_build.HERE), "include", "tnf.h")
    src = '#include "%s"\nint main(void){return tnf_abi_version==0;}\n' % hdr
    res = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", "-"], input=src, text=True,
                         capture_output=True)
    assert res.returncode == 0, res.stderr


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", _build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_argument_errors_without_gpu():
    lib = _lib.lib()
    rc = lib.tnf_coupling(None, None, None, None, 0, 1, 1, 4, 15, 2, 1, 0, 0, 0, None)
    assert rc < 0 and b"null pointer" in lib.tnf_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    rc = lib.tnf_coupling(p, p, p, p, 0, 1, 1, 1, 15, 2, 1, 0, 0, 0, None)
    assert rc < 0 and b"D=1" in lib.tnf_last_error()
    rc = lib.tnf_coupling(p, p, p, p, 0, 1, 1, 4, 15, 9, 1, 0, 0, 0, None)
    assert rc < 0 and b"L=9" in lib.tnf_last_error()
    rc = lib.tnf_affine(p, p, p, p, 0, 1, 1, 4, 0, 7, None)
    assert rc < 0 and b"dtype" in lib.tnf_last_error()


def test_build_digest_is_path_independent(tmp_path, monkeypatch):
    """The in-tree library travels with the repository to other roots (GPU boxes run the snapshot from a scratch
    path): whether it is current must depend on the source CONTENTS only, or every rank would rebuild it."""
    import shutil
    from torch_nf_b200 import _build
    ref = _build.source_digest()
    root = tmp_path / "elsewhere"
    shutil.copytree(_build.CSRC, root / "torch_nf_b200" / "csrc")
    (root / "include").mkdir()
    shutil.copy(os.path.join(os.path.dirname(_build.HERE), "include", "tnf.h"), root / "include" / "tnf.h")
    monkeypatch.setattr(_build, "HERE", str(root / "torch_nf_b200"))
    monkeypatch.setattr(_build, "CSRC", str(root / "torch_nf_b200" / "csrc"))
    assert _build.source_digest() == ref
