"""The C-ABI boundary: libalq.so loads on a CPU-only box and exports exactly the entry points
include/alq.h declares; the ctypes binding mirrors them.  No compute calls here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "alq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(alq_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    from active_learning_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in alq.h but not exported"


def test_ctypes_binding_mirrors_header():
    from active_learning_b200 import _lib
    assert set(_lib.SIGNATURES) == _header_functions()
    lib = _lib.load()
    assert lib.alq_version() == _lib.ABI_VERSION


def test_greedy_desc_layout_matches_header_field_order():
    from active_learning_b200 import _lib
    src = open(os.path.join(ROOT, "include", "alq.h")).read()
    body = src[src.index("typedef struct alq_greedy_desc {"):src.index("} alq_greedy_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(\w+)\s*;", body)
    assert fields == [f[0] for f in _lib.GreedyDesc._fields_]


def test_create_without_gpu_fails_cleanly():
    import torch
    if torch.cuda.is_available():
        return
    from active_learning_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.alq_create(ctypes.byref(h), 0) != 0 and not h.value


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "alq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int64_t|int|void|const char\*)\s+(alq_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        args = [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"]
        protos[name] = (ret, args)
    return protos


def test_ctypes_argument_lists_match_the_header_prototypes():
    """Every parameter of every prototype: pointer vs 32-bit vs 64-bit integer must agree with the ctypes argtypes
    (a drifted binding would pass garbage without any error)."""
    from active_learning_b200 import _lib
    protos = _header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)

    def kind_of_c(decl):
        if "*" in decl:
            return "ptr"
        t = decl.split()[-2] if len(decl.split()) > 1 else decl
        return {"int64_t": "i64", "int32_t": "i32", "int": "i32", "size_t": "i64"}[t]

    def kind_of_ctypes(t):
        if t in (ctypes.c_int64, ctypes.c_size_t):
            return "i64"
        if t in (ctypes.c_int32, ctypes.c_int):
            return "i32"
        return "ptr"                                   # c_void_p, c_char_p, POINTER(...)

    for name, (ret, args) in protos.items():
        restype, argtypes = _lib.SIGNATURES[name]
        assert [kind_of_c(a) for a in args] == [kind_of_ctypes(t) for t in argtypes], name
        assert (restype is None) == (ret == "void"), name
