"""Julia is not installable in this image, so julia/iLQRB200.jl cannot be executed here.  The only guard the binding
can have is static: every `ccall((:name, lib), Ret, (Args...), ...)` tuple is parsed and checked — name, arity, return
type and every argument type — against the prototypes of include/ilqr_b200.h, and the `Problem` struct against
`struct ilqr_problem` field by field."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "julia", "iLQRB200.jl")
HDR = os.path.join(ROOT, "include", "ilqr_b200.h")


def _strip_c_comments(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def c_prototypes():
    """name -> (return type, [argument types]) with types normalised ('const' dropped, spaces removed)."""
    src = _strip_c_comments(open(HDR).read())
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(ilqr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        if "typedef" in ret:
            continue

        def norm(t):
            t = re.sub(r"\bconst\b", "", t)
            return re.sub(r"\s+", "", t)

        argl = []
        if args.strip() not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)     # drop the parameter name
                argl.append(norm(mm.group(1)))
        out[name] = (norm(ret), argl)
    return out


# which Julia ccall types may stand for a C type
SCALARS = {"int32_t": {"Int32", "Cint"}, "int64_t": {"Int64", "Clonglong"}, "uint64_t": {"UInt64", "Csize_t"},
           "double": {"Float64", "Cdouble"}, "void": {"Cvoid", "Nothing"}}
POINTERS = {"double*": {"Ptr{Float64}", "Ptr{Cvoid}", "Ref{Float64}"}, "int32_t*": {"Ptr{Int32}", "Ptr{Cvoid}", "Ref{Int32}"},
            "int64_t*": {"Ptr{Int64}", "Ref{Int64}"}, "float*": {"Ptr{Float32}", "Ref{Float32}"}, "void*": {"Ptr{Cvoid}"},
            "void**": {"Ptr{Ptr{Cvoid}}", "Ref{Ptr{Cvoid}}"}, "char*": {"Cstring", "Ptr{UInt8}"},
            "ilqr_problem*": {"Ptr{Problem}", "Ref{Problem}"},
            "ilqr_handle*": {"Ptr{Cvoid}"}, "ilqr_pool*": {"Ptr{Cvoid}"}, "ilqr_streamer*": {"Ptr{Cvoid}"},
            "ilqr_handle**": {"Ptr{Ptr{Cvoid}}", "Ref{Ptr{Cvoid}}"}, "ilqr_pool**": {"Ptr{Ptr{Cvoid}}", "Ref{Ptr{Cvoid}}"},
            "ilqr_streamer**": {"Ptr{Ptr{Cvoid}}", "Ref{Ptr{Cvoid}}"}}


def allowed(ctype):
    return SCALARS.get(ctype) or POINTERS.get(ctype) or set()


def _split_top(s):
    """split on commas that are not inside (), {} or []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    """[(name, ret, [argtypes], n_values)] for every ccall in the binding."""
    src = open(JL).read()
    src = re.sub(r"#[^\n]*", "", src)
    calls = []
    for m in re.finditer(r"ccall\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        parts = _split_top(src[m.end():i - 1])
        name = re.match(r"\(:([a-z0-9_]+),\s*lib\)", parts[0]).group(1)
        ret = parts[1]
        assert parts[2].startswith("(") and parts[2].endswith(")"), parts[2]
        argt = _split_top(parts[2][1:-1])
        calls.append((name, ret, argt, len(parts) - 3))
    return calls


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 40
    for name, ret, argt, nvals in calls:
        assert name in protos, "ccall to %s: not declared in include/ilqr_b200.h" % name
        cret, cargs = protos[name]
        assert ret in allowed(cret), "%s: return type %s does not match C %s" % (name, ret, cret)
        assert len(argt) == len(cargs), "%s: %d argument types, header has %d" % (name, len(argt), len(cargs))
        assert nvals == len(argt), "%s: %d values for %d argument types" % (name, nvals, len(argt))
        for k, (jt, ct) in enumerate(zip(argt, cargs)):
            assert jt in allowed(ct), "%s: argument %d is %s, C has %s" % (name, k, jt, ct)


def test_binding_covers_the_host_control_surface():
    """north_star: Julia owns problem setup, convergence and regularisation control — every entry point that serves
    that has a stub (set_reg / set_active / upload_x0 / upload_gains / pool / stream_solve_device were missing)."""
    bound = {c[0] for c in julia_ccalls()}
    need = {"ilqr_create", "ilqr_destroy", "ilqr_last_error", "ilqr_upload", "ilqr_upload_x0", "ilqr_upload_gains",
            "ilqr_backward_pass", "ilqr_forward_pass", "ilqr_commit", "ilqr_iterate", "ilqr_fit", "ilqr_set_reg", "ilqr_set_active",
            "ilqr_download", "ilqr_solve", "ilqr_stream_solve_device", "ilqr_mpc_start", "ilqr_mpc_step",
            "ilqr_pool_create", "ilqr_pool_destroy", "ilqr_pool_submit", "ilqr_pool_submit_device", "ilqr_pool_wait", "ilqr_pool_wait_all",
            "ilqr_streamer_create", "ilqr_streamer_destroy", "ilqr_streamer_submit", "ilqr_streamer_submit_device",
            "ilqr_streamer_submit_x0", "ilqr_streamer_submit_x0_device", "ilqr_streamer_wait", "ilqr_streamer_wait_all",
            "ilqr_problem_two_link", "ilqr_problem_serial_chain", "ilqr_problem_custom", "ilqr_host_alloc", "ilqr_host_free"}
    assert need <= bound, sorted(need - bound)


def test_problem_struct_mirrors_the_header_field_by_field():
    hdr = _strip_c_comments(open(HDR).read())
    body = re.search(r"typedef struct ilqr_problem \{(.*?)\} ilqr_problem;", hdr, flags=re.S).group(1)
    consts = {k: int(v) for k, v in re.findall(r"#define (ILQR_[A-Z_]+) (\d+)", hdr)}
    cfields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        mm = re.match(r"(const char\*|int32_t|double)\s+(.*)$", decl, flags=re.S)
        ctype, names = mm.group(1), mm.group(2)
        for nm in names.split(","):
            nm = nm.strip()
            arr = re.match(r"([a-z_0-9]+)\[(.*)\]$", nm)
            if arr:
                expr = arr.group(2)
                for k, v in consts.items():
                    expr = expr.replace(k, str(v))
                cfields.append((arr.group(1), ctype, int(eval(expr))))
            else:
                cfields.append((nm, ctype, 0))
    jl = open(JL).read()
    sbody = re.search(r"struct Problem\n(.*?)\nend", jl, flags=re.S).group(1)
    jfields = []
    for line in sbody.splitlines():
        line = re.sub(r"#.*", "", line).strip()
        if not line:
            continue
        nm, jt = line.split("::")
        arr = re.match(r"NTuple\{(\d+),Float64\}", jt)
        if arr:
            jfields.append((nm, "double", int(arr.group(1))))
        else:
            jfields.append((nm, {"Int32": "int32_t", "Float64": "double", "Cstring": "const char*"}[jt], 0))
    assert jfields == cfields
