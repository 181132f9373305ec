"""The Rust binding (rust/online-phase/src/b200) cannot be compiled in this image (no cargo/rustc); what CAN be checked is
that the generated `extern "C"` block is exactly what the header declares, that the hand-written modules only call functions
that exist there with the right number of arguments, and that no body is elided."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200 = os.path.join(ROOT, "rust", "online-phase", "src", "b200")


def test_sys_rs_is_generated_from_the_header():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"]) == 0, \
        "rust/online-phase/src/b200/sys.rs is stale: run tools/gen_rust_sys.py"


def test_sys_rs_binds_every_header_symbol():
    from tests.test_abi import declared_symbols

    text = open(os.path.join(B200, "sys.rs")).read()
    bound = set(re.findall(r"pub fn (arkmpc_\w+)\(", text))
    assert bound == set(declared_symbols())


def _arity(sys_text):
    out = {}
    for m in re.finditer(r"pub fn (arkmpc_\w+)\((.*?)\)(?: -> [^;]+)?;", sys_text):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if not args else args.count(":")
    return out


def _split_args(s):
    depth, cur, out = 0, "", []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_wrappers_call_existing_functions_with_the_right_arity():
    arity = _arity(open(os.path.join(B200, "sys.rs")).read())
    for name in ("mod.rs", "batch.rs"):
        text = open(os.path.join(B200, name)).read()
        text = re.sub(r"//.*", "", text)
        for m in re.finditer(r"sys::(arkmpc_\w+)\s*\(", text):
            fn = m.group(1)
            assert fn in arity, f"{name}: calls {fn}, which the header does not declare"
            # balanced-parenthesis scan of the argument list
            i, depth = m.end(), 1
            while depth:
                depth += {"(": 1, ")": -1}.get(text[i], 0)
                i += 1
            n_args = len(_split_args(text[m.end():i - 1]))
            assert n_args == arity[fn], f"{name}: {fn} called with {n_args} arguments, the ABI takes {arity[fn]}"


def test_no_elided_bodies():
    for name in ("mod.rs", "batch.rs", "carrier.rs"):
        text = open(os.path.join(B200, name)).read()
        assert "todo!" not in text and "unimplemented!" not in text and "/* ... */" not in text and "/* … */" not in text
        assert text.count("{") == text.count("}") and text.count("(") == text.count(")")
