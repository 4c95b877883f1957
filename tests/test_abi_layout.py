"""The three statements of the C ABI's data layout agree: include/swfr.h as gcc lays it out (sizeof / offsetof printed by
a generated C program), the ctypes mirrors in swf_renderer_b200/capi.py, and the #[repr(C)] structs of the Rust shim
(rs-shim/src/ffi.rs, laid out here by the C rules - there is no Rust toolchain in the image).  Host only."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "swfr.h")
FFI_RS = os.path.join(ROOT, "rs-shim", "src", "ffi.rs")

# header struct -> ctypes class name in capi.py
CTYPES_NAMES = {
    "swfr_rgba8": "Rgba8", "swfr_swf_matrix": "SwfMatrix", "swfr_color_stop": "ColorStop", "swfr_gradient": "Gradient",
    "swfr_fill_style": "FillStyle", "swfr_line_style": "LineStyle", "swfr_styles": "Styles",
    "swfr_shape_record": "ShapeRecord", "swfr_define_shape": "DefineShape", "swfr_color_transform": "ColorTransform",
    "swfr_display_primitive": "DisplayPrimitive", "swfr_stage": "Stage", "swfr_display_object": "DisplayObject",
    "swfr_display_stage": "DisplayStage", "swfr_frames_export": "FramesExport", "swfr_stats": "Stats",
}


def header_structs():
    """[(struct name, [field name, ...])] of every non-opaque struct typedef of the header."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = []
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*\1\s*;", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "const struct swfr_display_object *children", "uint8_t r, g, b, a", "float matrix[6]"
            names = decl.split(",")
            first = re.search(r"(\w+)\s*(\[\w+\])?\s*$", names[0])
            fields.append(first.group(1))
            for more in names[1:]:
                fields.append(re.search(r"(\w+)\s*(\[\w+\])?\s*$", more).group(1))
        out.append((m.group(1), fields))
    return out


@pytest.fixture(scope="module")
def c_layout(tmp_path_factory):
    """{struct: (size, [offset, ...])} as gcc lays the header out."""
    structs = header_structs()
    assert len(structs) >= 16
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "swfr.h"', "int main(void) {"]
    for name, fields in structs:
        lines.append(f'  printf("{name} %zu", sizeof({name}));')
        for f in fields:
            lines.append(f'  printf(" %zu", offsetof({name}, {f}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    d = tmp_path_factory.mktemp("abi")
    (d / "layout.c").write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c11", "-I", os.path.dirname(HEADER), "-o", str(d / "layout"), str(d / "layout.c")], check=True)
    out = subprocess.run([str(d / "layout")], check=True, capture_output=True, text=True).stdout
    lay = {}
    for line in out.splitlines():
        parts = line.split()
        lay[parts[0]] = (int(parts[1]), [int(v) for v in parts[2:]])
    return lay


def test_ctypes_mirrors_have_the_headers_layout(c_layout):
    from swf_renderer_b200 import capi

    assert set(CTYPES_NAMES) == set(c_layout), "a struct of the header has no ctypes mirror (or the other way round)"
    for cname, pyname in CTYPES_NAMES.items():
        cls = getattr(capi, pyname)
        size, offsets = c_layout[cname]
        assert C.sizeof(cls) == size, cname
        assert [getattr(cls, f[0]).offset for f in cls._fields_] == offsets, cname


# ---- the Rust shim: #[repr(C)] laid out by the C rules ----------------------------------------------------

PRIM = {"u8": (1, 1), "i8": (1, 1), "u16": (2, 2), "i16": (2, 2), "u32": (4, 4), "i32": (4, 4), "f32": (4, 4), "u64": (8, 8),
        "i64": (8, 8), "f64": (8, 8), "usize": (8, 8), "isize": (8, 8), "c_int": (4, 4), "c_uint": (4, 4), "c_float": (4, 4)}


def rust_structs():
    src = open(FFI_RS).read()
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*(?:#\[[^\]]*\]\s*)*pub\s+struct\s+(\w+)\s*\{(.*?)\n\}", src, flags=re.S):
        fields = []
        for f in re.finditer(r"pub\s+(\w+)\s*:\s*([^,\n]+(?:\[[^\]]*\])?)\s*,", m.group(2)):
            fields.append((f.group(1), f.group(2).strip()))
        out[m.group(1)] = fields
    return out


def rust_layout(name, structs, memo):
    """(size, align, [offset, ...]) of a #[repr(C)] struct."""
    if name in memo:
        return memo[name]

    def type_layout(t):
        t = t.strip()
        if t.startswith("*const") or t.startswith("*mut"):
            return 8, 8
        arr = re.match(r"\[\s*(.+?)\s*;\s*(\d+)\s*\]$", t)
        if arr:
            s, a = type_layout(arr.group(1))
            return s * int(arr.group(2)), a
        t = t.split("::")[-1]
        if t in PRIM:
            return PRIM[t]
        s, a, _ = rust_layout(t, structs, memo)
        return s, a

    off, align, offsets = 0, 1, []
    for _, t in structs[name]:
        s, a = type_layout(t)
        off = (off + a - 1) // a * a
        offsets.append(off)
        off += s
        align = max(align, a)
    memo[name] = ((off + align - 1) // align * align, align, offsets)
    return memo[name]


def test_rust_shim_structs_have_the_headers_layout(c_layout):
    structs = rust_structs()
    memo = {}
    checked = 0
    for cname, (size, offsets) in c_layout.items():
        if cname not in structs:
            continue
        rsize, _, roffsets = rust_layout(cname, structs, memo)
        assert rsize == size, cname
        assert roffsets == offsets, cname
        checked += 1
    # everything the shim passes across the boundary: tags, styles, records, stages, primitives
    for must in ("swfr_define_shape", "swfr_shape_record", "swfr_fill_style", "swfr_line_style", "swfr_styles",
                 "swfr_gradient", "swfr_color_stop", "swfr_display_primitive", "swfr_stage", "swfr_color_transform"):
        assert must in structs, must
    assert checked >= 10


def test_rust_shim_declares_the_current_abi_version(built_library):
    m = re.search(r"SWFR_ABI_VERSION:\s*u32\s*=\s*(\d+)", open(FFI_RS).read())
    assert int(m.group(1)) == built_library.swfr_abi_version()


def test_ts_addon_type_checks_against_the_header():
    """ts-shim/src/addon.cc (the N-API binding a maintainer adds to the reference's TypeScript package) cannot be built
    here - no Node toolchain - but its use of include/swfr.h can be type-checked: g++ -fsyntax-only against a
    declarations-only stand-in for <napi.h> (tests/mock_napi)."""
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "tests", "mock_napi"), "-I",
                        os.path.dirname(HEADER), os.path.join(ROOT, "ts-shim", "src", "addon.cc")], capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr
