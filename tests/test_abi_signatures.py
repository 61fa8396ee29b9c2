"""Every prototype of include/fnst.h against its ctypes binding in _lib.EXPORTS: same arity, same scalar widths, pointers where
the header has pointers.  A mismatch here is a silent stack / register corruption at call time, so it is checked on CPU."""
import ctypes as C
import os
import re

from fast_neural_style_transfer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _prototypes():
    text = open(os.path.join(ROOT, "include", "fnst.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)                       # drop comments
    for m in re.finditer(r"^(int64_t|int|const char\*)\s+(fnst_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.M | re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        yield name, ret, params


def _kind(param: str) -> str:
    if "*" in param:
        return "ptr"
    base = param.rsplit(" ", 1)[0].replace("const ", "").strip() if " " in param else param
    return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "float": "f32", "double": "f64"}[base]


def _ckind(t) -> str:
    if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
        return "ptr"
    return {C.c_int: "i32", C.c_int32: "i32", C.c_int64: "i64", C.c_float: "f32", C.c_double: "f64"}[t]


def test_header_prototypes_match_ctypes_bindings():
    protos = {name: (ret, params) for name, ret, params in _prototypes()}
    assert set(protos) == set(_lib.EXPORTS), set(protos) ^ set(_lib.EXPORTS)
    for name, (ret, params) in protos.items():
        restype, argtypes = _lib.EXPORTS[name]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for i, (p, t) in enumerate(zip(params, argtypes)):
            assert _kind(p) == _ckind(t), (name, i, p, t)
        want = {"int": C.c_int, "int64_t": C.c_int64, "const char*": C.c_char_p}[ret]
        assert restype is want, (name, ret, restype)


def test_conv_descriptor_layout_matches_the_c_struct(tmp_path):
    """sizeof / offsetof of fnst_conv_desc as the C compiler lays it out == the ctypes.Structure mirror in _lib.ConvDesc."""
    import subprocess
    fields = [f[0] for f in _lib.ConvDesc._fields_]
    src = tmp_path / "layout.c"
    lines = "\n".join(f'  printf("{f} %zu\\n", offsetof(fnst_conv_desc, {f}));' for f in fields)
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "fnst.h"\nint main(void) {\n'
                   '  printf("sizeof %zu\\n", sizeof(fnst_conv_desc));\n' + lines + "\n  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out["sizeof"]) == C.sizeof(_lib.ConvDesc)
    for f in fields:
        assert int(out[f]) == getattr(_lib.ConvDesc, f).offset, f
    assert _lib.MAX_TAPS == 96 and "#define FNST_MAX_TAPS 96" in open(os.path.join(ROOT, "include", "fnst.h")).read()


def test_enum_constants_match_the_header():
    text = open(os.path.join(ROOT, "include", "fnst.h")).read()
    vals = {k: int(v) for k, v in re.findall(r"\b(FNST_[A-Z0-9_]+)\s*=\s*(\d+)", text)}
    vals.update({k: int(v) for k, v in re.findall(r"#define\s+(FNST_[A-Z0-9_]+)\s+(\d+)", text)})
    mirror = {"FNST_F32": _lib.F32, "FNST_F16": _lib.F16, "FNST_BF16": _lib.BF16, "FNST_EPI_NHWC": _lib.EPI_NHWC,
              "FNST_EPI_D2S": _lib.EPI_D2S, "FNST_EPI_NCHW_F32": _lib.EPI_NCHW_F32, "FNST_EPI_ROWSUM9": _lib.EPI_ROWSUM9,
              "FNST_PAD_NONE": _lib.PAD_NONE, "FNST_PAD_REFLECT": _lib.PAD_REFLECT, "FNST_PAD_ZERO": _lib.PAD_ZERO,
              "FNST_DESC_PREZEROED": _lib.DESC_PREZEROED, "FNST_DESC_LINEAR": _lib.DESC_LINEAR, "FNST_MAX_TAPS": _lib.MAX_TAPS}
    for name, value in mirror.items():
        assert vals[name] == value, (name, vals.get(name), value)
