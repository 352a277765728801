"""The C ABI: every function include/rtb.h declares is exported by the in-tree library that the
binding expects it in; ctypes struct layouts match the header (checked against gcc)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import rendering_b200 as rb
from rendering_b200 import _ffi

HEADER = os.path.join(rb.REPO_ROOT, "include", "rtb.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb_[a-z_0-9]+)\s*\(", text)))


def test_header_functions_are_all_bound():
    assert declared_functions() == sorted(_ffi.HOST_SYMBOLS + _ffi.CUDA_SYMBOLS)


def test_host_library_exports():
    lib = C.CDLL(_ffi.HOST_LIB_PATH)
    for name in _ffi.HOST_SYMBOLS:
        assert hasattr(lib, name), name


def test_cuda_library_loads_and_exports():
    # loads without a GPU (cudart is linked statically); no compute call is made here
    lib = C.CDLL(_ffi.CUDA_LIB_PATH)
    for name in _ffi.CUDA_SYMBOLS:
        assert hasattr(lib, name), name
    lib.rtb_abi_version.restype = C.c_int
    assert lib.rtb_abi_version() == 2
    lib.rtb_strip_rows_owned.restype = C.c_int
    assert lib.rtb_strip_rows_owned(1080, 32, 0, 8) + lib.rtb_strip_rows_owned(1080, 32, 7, 8) > 0


def test_struct_layouts_match_gcc():
    names = ["RtbCamera", "RtbObject", "RtbLight", "RtbNode", "RtbImage", "RtbMesh", "RtbScene", "RtbStats"]
    src = '#include <stdio.h>\n#include "rtb.h"\nint main(){' + "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.dirname(HEADER), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    for n, s in zip(names, sizes):
        assert C.sizeof(getattr(_ffi, n)) == s, n


def test_cuda_backend_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    sc = rb.Scene(rb.scene_path("cfg1_simple_shapes_256"))
    try:
        rb.Renderer(sc)
    except rb.RtbError as e:
        assert e.code == _ffi.RTB_ERR_CUDA
    else:
        raise AssertionError("Renderer must not come up without a GPU: there is no CPU fallback")
