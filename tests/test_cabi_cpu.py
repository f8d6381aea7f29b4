"""CPU-side checks: the C-ABI library loads, exports every symbol include/vpm_b200.h declares, the ctypes
table matches the header, compute entry points fail loudly without a GPU, and the oracle reproduces the
committed golden fixtures (regression pin of the checker)."""
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vpm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vpm_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import vpm_b200
    return vpm_b200


def test_library_exports_every_declared_symbol(built):
    import ctypes
    lib = ctypes.CDLL(built.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vpm_b200.h but not exported"
    assert sorted(built._cabi.SIGNATURES) == syms, "ctypes table and header disagree"
    assert lib.vpm_version() == 100


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.VpmError) as e:
        built.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vlasovparticlemethods.jl_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
            txt = open(path, errors="ignore").read()
            assert "oracle" not in txt.lower(), f"{path} references the oracle"


def test_sm100a_only(built):
    out = os.popen(f"cuobjdump -lelf {built.LIB_PATH} 2>/dev/null").read()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("name", ["vp_k4_n16", "vp_k3_n16_cfg1", "vp_k5_n11_chi"])
def test_oracle_matches_golden_vp(oracle, name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    xs = oracle.XSpace(0.0, float(g["L"]), int(g["K"]), int(g["nh"]))
    rhs = xs.deposit(g["x"], g["w"])
    np.testing.assert_allclose(rhs, g["rhs"], rtol=0, atol=1e-15 * np.abs(g["rhs"]).max() * 10)
    np.testing.assert_allclose(xs.poisson_solve(rhs), g["phi"], rtol=0, atol=1e-13 * np.abs(g["phi"]).max())
    x1, v1, diag, _ = xs.strang_selfconsistent(g["x"], g["v"], g["w"], float(g["dt"]), int(g["nsteps"]), chi=float(g["chi"]))
    np.testing.assert_allclose(x1, g["x1"], rtol=1e-13)
    np.testing.assert_allclose(v1, g["v1"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(diag, g["diag"], rtol=1e-11)


@pytest.mark.parametrize("name", ["lb_k4_n41", "lb_k5_n12"])
def test_oracle_matches_golden_lb(oracle, name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    vs = oracle.VSpace(-10.0, 10.0, int(g["nknots"]), int(g["K"]))
    np.testing.assert_allclose(vs.mass(), g["mass"], atol=1e-15)
    coef = vs.project(g["v"], g["w"])
    np.testing.assert_allclose(coef, g["coef"], atol=1e-14)
    np.testing.assert_allclose(vs.lb_rhs(g["v"], g["w"], 0.7, True)[0], g["vdot_clb"], atol=1e-13)
    v2, d = vs.rk438(g["v"], g["w"], 0.7, float(g["dt"]), int(g["nsteps"]), conservative=True)
    np.testing.assert_allclose(v2, g["v_clb"], atol=1e-12)


def _split_top(s):
    """split a comma-separated list at nesting depth 0"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_shim_binds_the_declared_abi():
    """julia/VPMB200.jl cannot run here (no Julia toolchain), so check statically what a `ccall` would trip over:
    every bound symbol is declared in include/vpm_b200.h, with the same number of arguments, pointer arguments
    where the header has pointers and 64-bit integers where the header has int64_t / uint64_t."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vpm_b200.h")).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*|void)\s+(vpm_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = [a for a in _split_top(m.group(2)) if a and a != "void"]
        protos[m.group(1)] = args
    jl = open(os.path.join(ROOT, "julia", "VPMB200.jl")).read()
    calls = list(re.finditer(r"ccall\(\(:(vpm_[a-z0-9_]+),\s*libvpm\),\s*(\w+),\s*\(", jl))
    assert len(calls) >= 30
    for m in calls:
        name = m.group(1)
        assert name in protos, f"julia shim binds {name}, which include/vpm_b200.h does not declare"
        # the argument-type tuple starts at the '(' that ends the match
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(jl[i], 0)
            i += 1
        jtypes = _split_top(jl[m.end():i - 1])
        cargs = protos[name]
        assert len(jtypes) == len(cargs), f"{name}: julia passes {len(jtypes)} arguments, header declares {len(cargs)}"
        for jt, ca in zip(jtypes, cargs):
            is_ptr_c = "*" in ca
            is_ptr_j = jt.startswith(("Ptr{", "Ref{")) or jt == "Cstring"
            assert is_ptr_c == is_ptr_j, f"{name}: '{ca}' bound as {jt}"
            if not is_ptr_c:
                if "int64_t" in ca:
                    assert jt in ("Int64", "UInt64"), f"{name}: '{ca}' bound as {jt}"
                elif re.match(r"\s*(const\s+)?double\b", ca):
                    assert jt == "Float64", f"{name}: '{ca}' bound as {jt}"
                else:
                    assert jt == "Cint", f"{name}: '{ca}' bound as {jt}"


def test_julia_shim_blocks_and_brackets_balance():
    """No Julia parser is available; catch gross syntax slips statically: with comments and strings removed, every
    (), [], {} closes in order, and at bracket depth 0 every block opener (function, struct, if, for, while, let, begin, try,
    do, module, quote, macro) has its `end`."""
    for name in ("VPMB200.jl", "parity_check.jl"):
        src = open(os.path.join(ROOT, "julia", name)).read()
        src = re.sub(r'"""(?:.|\n)*?"""', '""', src)
        src = re.sub(r'"(?:\\.|[^"\\\n])*"', '""', src)
        src = re.sub(r"#=.*?=#", "", src, flags=re.S)
        src = re.sub(r"#[^\n]*", "", src)
        src = re.sub(r"'(?:\\.|[^'\\\n])'", "' '", src)
        stack, blocks = [], []
        pairs = {")": "(", "]": "[", "}": "{"}
        openers = {"function", "struct", "if", "for", "while", "let", "begin", "try", "do", "module", "quote", "macro"}
        for m in re.finditer(r"[^\W\d]\w*!?|[()\[\]{}]", src):
            tok = m.group(0)
            if tok in "([{":
                stack.append(tok)
            elif tok in pairs:
                assert stack and stack[-1] == pairs[tok], f"{name}: unbalanced '{tok}' at offset {m.start()}"
                stack.pop()
            elif not stack:
                prev = src[max(0, m.start() - 1):m.start()]
                if prev == ":" or prev == ".":          # :end / :function symbols, field access
                    continue
                if tok in openers:
                    if tok == "struct" and blocks and blocks[-1][0] == "mutable":
                        blocks.pop()
                    blocks.append((tok, m.start()))
                elif tok == "mutable":
                    blocks.append((tok, m.start()))
                elif tok == "end":
                    assert blocks, f"{name}: 'end' without an open block at offset {m.start()}"
                    blocks.pop()
        assert not stack, f"{name}: unclosed bracket {stack[-1]}"
        assert not blocks, f"{name}: unclosed block {blocks[-1]}"
