"""CPU-side checks of the C-ABI boundary: the library builds, loads without a GPU, exports every
symbol include/nvr_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import REPO


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from instant_nvr_b200 import cabi
    return cabi.load()


def declared_functions():
    text = open(os.path.join(REPO, "include", "nvr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nvr_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from instant_nvr_b200 import cabi
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nvr_b200.h but not exported"
        assert n in cabi.SYMBOLS, f"{n} has no ctypes prototype"
    assert sorted(cabi.SYMBOLS) == names


def test_abi_version(lib):
    from instant_nvr_b200 import cabi
    header = open(os.path.join(REPO, "include", "nvr_b200.h")).read()
    declared = int(re.search(r"#define NVR_ABI_VERSION (\d+)", header).group(1))
    assert lib.nvr_abi_version() == cabi.ABI_VERSION == declared


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    from instant_nvr_b200 import cabi
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.engine import Engine
    conf = cabi.NvrConfig(cabi.ABI_VERSION, 0, 0.05, 0)
    h = C.c_void_p()
    assert lib.nvr_create(C.byref(conf), C.byref(h)) != 0 and not h.value
    with pytest.raises(RuntimeError):
        Engine(PathConfig.inb_377(log2_T_cap=10))


def test_bad_abi_version_rejected(lib):
    from instant_nvr_b200 import cabi
    conf = cabi.NvrConfig(99, 0, 0.05, 0)
    h = C.c_void_p()
    assert lib.nvr_create(C.byref(conf), C.byref(h)) == 3


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under instant_nvr_b200/ may reference it."""
    pkg = os.path.join(REPO, "instant_nvr_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "nvr_oracle" not in text and "import oracle" not in text, os.path.join(root, f)


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors in instant_nvr_b200/cabi.py against the C compiler's view of include/nvr_b200.h: sizeof every struct
    and offsetof every field (a drifted mirror would pass garbage pointers across the boundary without any error)."""
    import subprocess
    from instant_nvr_b200 import cabi
    header = open(os.path.join(REPO, "include", "nvr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    structs = {}
    for m in re.finditer(r"typedef struct (Nvr\w+)\s*\{(.*?)\}\s*\1\s*;", text, flags=re.S):
        name, body = m.group(1), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):                      # `int64_t a, b;` style declarations
                fm = re.search(r"(\w+)\s*(\[[^\]]*\]\s*)*$", part.strip())
                assert fm, (name, decl)
                fields.append(fm.group(1))
        structs[name] = fields
    assert len(structs) >= 12 and len(structs["NvrFrame"]) >= 15 and "n_passes" in structs["NvrCounters"]
    mirrors = {n: getattr(cabi, n) for n in structs if hasattr(cabi, n)}
    assert set(mirrors) == set(structs), sorted(set(structs) - set(mirrors))
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nvr_b200.h"', 'int main(void) {']
    for name, fields in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for f in fields:
            lines.append(f'  printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), "-o", str(exe), str(src)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, fields in structs.items():
        cls = mirrors[name]
        assert int(out[name]) == C.sizeof(cls), (name, out[name], C.sizeof(cls))
        assert [f for f, *_ in cls._fields_] == fields, (name, [f for f, *_ in cls._fields_], fields)
        for f in fields:
            assert int(out[f"{name}.{f}"]) == getattr(cls, f).offset, (name, f)
