"""The drop-in boundary without a GPU: libb200sr.so loads, exports every entry point include/b200sr.h declares,
the ctypes binding lists exactly those, and the epilogue struct has the layout a C compiler gives the header.
No kernel is launched."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200sr.h")


@pytest.fixture(scope="module")
def lib():
    from b200sr import _lib

    if not os.path.exists(_lib.LIB_PATH):  # fresh checkout: nvcc cross-compiles without a GPU
        sys.path.insert(0, ROOT)
        import __graft_entry__

        __graft_entry__.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(b200sr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ("b200sr_gemm_bf16", "b200sr_conv3x3_bf16", "b200sr_attention_d64", "b200sr_group_norm_nhwc",
                 "b200sr_layer_norm", "b200sr_sampler_post", "b200sr_tile_accumulate", "b200sr_rel_l1_similarity",
                 "b200sr_sr3_update"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared in include/b200sr.h but not exported: {missing}"


def test_binding_covers_exactly_the_header(lib):
    from b200sr import _lib

    declared = set(declared_symbols())
    bound = set(_lib.SIGNATURES) | {"b200sr_abi_version", "b200sr_num_sms"}
    assert declared - bound == set(), f"no ctypes signature for {sorted(declared - bound)}"
    assert bound - declared == set(), f"bound but not declared in the header: {sorted(bound - declared)}"
    lib.b200sr_abi_version.restype = ctypes.c_int
    assert lib.b200sr_abi_version() == _lib.ABI_VERSION


def test_epilogue_struct_layout_matches_a_c_compiler(tmp_path):
    from b200sr import _lib

    fields = [f[0] for f in _lib.Epilogue._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "b200sr.h"\nint main(void) {\n'
    prog += '  printf("%zu\\n", sizeof(b200sr_epilogue));\n'
    for f in fields:
        prog += f'  printf("%zu\\n", offsetof(b200sr_epilogue, {f}));\n'
    prog += "  return 0;\n}\n"
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text(prog)
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_lib.Epilogue)
    for name, off in zip(fields, out[1:]):
        assert getattr(_lib.Epilogue, name).offset == off, name


def test_integration_doc_embeds_the_current_binding():
    """INTEGRATION.md section 1 shows the struct declarations a maintainer pastes: they must be the binding's own
    (generated) text, field for field — a stale snippet mis-lays the struct silently."""
    from b200sr import _lib

    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert _lib.binding_snippet() in doc, "INTEGRATION.md struct snippet differs from b200sr._lib (regenerate it)"
    assert f"`b200sr_abi_version()` ({_lib.ABI_VERSION})" in doc


def test_copy_struct_layout_matches_a_c_compiler(tmp_path):
    from b200sr import _lib

    prog = ('#include <stdio.h>\n#include <stddef.h>\n#include "b200sr.h"\nint main(void) {\n'
            '  printf("%zu %zu %zu %zu\\n", sizeof(b200sr_copy), offsetof(b200sr_copy, src), offsetof(b200sr_copy, dst), '
            'offsetof(b200sr_copy, bytes));\n  return 0;\n}\n')
    src, exe = tmp_path / "copy.c", tmp_path / "copy"
    src.write_text(prog)
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out == [ctypes.sizeof(_lib.Copy), _lib.Copy.src.offset, _lib.Copy.dst.offset, _lib.Copy.bytes.offset]


def test_size_queries_work_without_a_device(lib):
    lib.b200sr_group_norm_workspace_bytes.restype = ctypes.c_size_t
    lib.b200sr_attention_d64_workspace_bytes.restype = ctypes.c_size_t
    assert lib.b200sr_group_norm_workspace_bytes(2, 1024, 1280, 32) > 1024
    assert lib.b200sr_attention_d64_workspace_bytes(2, 20, 1024, 77) == 0  # one key block: nothing to split
    assert lib.b200sr_attention_d64_workspace_bytes(0, 20, 1024, 1024) == 0
