"""The Rust FFI crate (rust/rest_tensors_b200) cannot be compiled in this image (no rustc / cargo), so its declarations
are checked against the C header instead: both files are parsed independently (the Rust side by the parser below, not by
the generator's own code) and every symbol must agree in name, arity, and per argument in pointer depth, constness of
the pointee and scalar width.  Reference call shapes: /root/reference/src/external_libs/ffi_restmatr.rs:4-62 (the seven
Fortran-ABI symbols) and src/external_libs/mod.rs:6-190 (usize -> i32 casts at the boundary)."""
from __future__ import annotations

import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import gen_rust_ffi  # noqa: E402

RUST_DIR = os.path.join(ROOT, "rust", "rest_tensors_b200", "src")

# canonical scalar classes: (kind, bits)
C_SCALAR = {"int": ("int", 32), "double": ("float", 64), "char": ("char", 8), "void": ("void", 0), "int64_t": ("int", 64),
            "uint64_t": ("uint", 64), "unsigned char": ("uint", 8), "unsigned": ("uint", 32), "unsigned short": ("uint", 16),
            "rb_ctx": ("opaque", 0)}
RUST_SCALAR = {"c_int": ("int", 32), "i32": ("int", 32), "c_double": ("float", 64), "f64": ("float", 64), "c_char": ("char", 8),
               "c_void": ("void", 0), "i64": ("int", 64), "u64": ("uint", 64), "u8": ("uint", 8), "u32": ("uint", 32), "u16": ("uint", 16),
               "RbCtx": ("opaque", 0)}


def canon_c(ctype: str):
    toks = ctype.replace("*", " * ").split()
    base, consts, i, base_const = [], [], 0, False
    while i < len(toks) and toks[i] != "*":
        if toks[i] == "const":
            base_const = True
        else:
            base.append(toks[i])
        i += 1
    pointee_const = base_const
    while i < len(toks):
        consts.append(pointee_const)
        i += 1
        pointee_const = False
        if i < len(toks) and toks[i] == "const":
            pointee_const = True
            i += 1
    # consts[0] = constness of the innermost pointee ... ; order from the base outwards
    return C_SCALAR[" ".join(base)], tuple(consts)


def canon_rust(rtype: str):
    toks = rtype.split()
    consts = []
    i = 0
    while toks[i] in ("*const", "*mut"):
        consts.append(toks[i] == "*const")
        i += 1
    base = toks[i].split("::")[-1]
    # Rust writes the outermost pointer first; the C canonical form lists from the base outwards
    return RUST_SCALAR[base], tuple(reversed(consts))


def parse_rust_externs(path: str):
    """-> {name: (ret_type_or_None, [arg types])} for every `pub fn` inside an extern "C" block."""
    src = open(path).read()
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for block in re.finditer(r'extern\s+"C"\s*\{(.*?)\n\s*\}', src, flags=re.S):
        for m in re.finditer(r"pub fn\s+([a-z_0-9]+)\s*\((.*?)\)\s*(?:->\s*([^;]+?))?\s*;", block.group(1), flags=re.S):
            name, args, ret = m.group(1), " ".join(m.group(2).split()), m.group(3)
            types = []
            if args:
                for a in args.split(","):
                    a = a.strip()
                    if not a:
                        continue
                    types.append(a.split(":", 1)[1].strip())
            assert name not in out, f"{name} declared twice in {path}"
            out[name] = (ret.strip() if ret else None, types)
    return out


def test_ffi_rs_is_current_with_the_generator():
    assert open(gen_rust_ffi.OUT).read() == gen_rust_ffi.emit(gen_rust_ffi.parse_header()), \
        "rust/.../ffi.rs is stale: run python tools/gen_rust_ffi.py"


def test_every_header_symbol_is_declared_with_matching_types():
    protos = gen_rust_ffi.parse_header()
    assert len(protos) >= 100
    rust = parse_rust_externs(os.path.join(RUST_DIR, "ffi.rs"))
    names = [p[1] for p in protos]
    assert sorted(names) == sorted(rust), (set(names) ^ set(rust))
    for ret, name, args in protos:
        rret, rargs = rust[name]
        assert len(args) == len(rargs), f"{name}: arity {len(args)} (C) vs {len(rargs)} (Rust)"
        if ret == "void":
            assert rret is None, name
        else:
            assert rret is not None and canon_c(ret) == canon_rust(rret), f"{name}: return {ret} vs {rret}"
        for k, ((ctype, aname), rtype) in enumerate(zip(args, rargs)):
            assert canon_c(ctype) == canon_rust(rtype), f"{name} arg {k} ({aname}): `{ctype}` (C) vs `{rtype}` (Rust)"


def test_compat_symbols_have_the_reference_fortran_abi():
    """ffi_restmatr.rs:4-62: every argument is a pointer, ints are 32-bit, doubles 64-bit, op chars are *const c_char, the
    functions return nothing; arities 6 / 25 / 18 / 12 / 15 / 15 / 17."""
    rust = parse_rust_externs(os.path.join(RUST_DIR, "ffi.rs"))
    arity = {"ri_ao2mo_f_": 6, "general_dgemm_f_": 25, "special_dgemm_f_01_": 18, "copy_mm_": 12, "copy_mr_": 15,
             "copy_rm_": 15, "copy_rr_": 17}
    for name, n in arity.items():
        ret, args = rust[name]
        assert ret is None and len(args) == n, name
        for t in args:
            scalar, consts = canon_rust(t)
            assert len(consts) == 1, f"{name}: `{t}` is not a plain pointer"
            assert scalar in (("int", 32), ("float", 64), ("char", 8)), f"{name}: {t}"


def test_library_exports_every_rust_declared_symbol():
    so = os.path.join(ROOT, "rest_tensors_b200", "librest_b200.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    syms = {line.split()[-1] for line in subprocess.check_output(["nm", "-D", "--defined-only", so], text=True).splitlines()}
    rust = parse_rust_externs(os.path.join(RUST_DIR, "ffi.rs"))
    missing = sorted(set(rust) - syms)
    assert not missing, f"declared in ffi.rs but not exported by librest_b200.so: {missing}"


def test_safe_wrappers_only_call_declared_symbols():
    """Every rb_* / compat symbol used by the safe modules of the crate exists in ffi.rs with the arity it is called with
    (a cheap stand-in for the type check rustc would do)."""
    rust = parse_rust_externs(os.path.join(RUST_DIR, "ffi.rs"))
    for fn in sorted(os.listdir(RUST_DIR)):
        if fn == "ffi.rs" or not fn.endswith(".rs"):
            continue
        src = re.sub(r"//[^\n]*", "", open(os.path.join(RUST_DIR, fn)).read())
        for m in re.finditer(r"\b(rb_[a-z_0-9]+|[a-z_0-9]+_f(?:_01)?_|copy_(?:mm|mr|rm|rr)_)\s*\(", src):
            name = m.group(1)
            assert name in rust, f"{fn}: calls undeclared symbol {name}"
            # count top-level commas of the call
            i, depth, commas, empty = m.end(), 1, 0, True
            while depth:
                ch = src[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                elif ch == "," and depth == 1:
                    commas += 1
                if depth and not ch.isspace():
                    empty = False
                i += 1
            argc = 0 if empty else commas + 1
            assert argc == len(rust[name][1]), f"{fn}: {name} called with {argc} args, declared with {len(rust[name][1])}"
