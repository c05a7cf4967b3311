"""julia/FundStabB200.jl cannot be executed in this image (no Julia), so its `ccall`s are checked statically against the
prototypes of include/fsb.h: every called symbol is declared, has the declared number of arguments, and every argument /
return type is of the declared class (Int32 <-> int32_t, Ptr{Float64} <-> double *, ...); the two `struct`s the glue passes
by reference mirror fsb_params / fsb_config field by field.  A mismatch here is a crash or silent corruption on the first
call from Julia."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_c(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def _c_class(t):
    t = " ".join(t.replace("const", " ").split())
    if "*" in t or "[" in t:
        base = t.replace("*", " ").split("[")[0].split()
        base = base[0] if base else ""
        depth = t.count("*") + t.count("[")
        return ("ptr", base, depth)
    return ("val", t, 0)


def _c_prototypes():
    src = _strip_c(open(os.path.join(ROOT, "include", "fsb.h")).read())
    protos = {}
    for ret, name, args in re.findall(r"([A-Za-z_][\w\s\*]*?)\b(fsb_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", src):
        params = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"(.*?)(\b\w+)(\s*\[\d*\])?$", a)
                typ = (m.group(1) + (m.group(3) or "")).strip()
                params.append(_c_class(typ))
        protos[name] = (_c_class(ret.strip()), params)
    return protos


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _julia_ccalls():
    src = open(os.path.join(ROOT, "julia", "FundStabB200.jl")).read()
    src = re.sub(r"#.*", "", src)
    calls = []
    for m in re.finditer(r"ccall\(\(:(fsb_\w+),\s*LIB\),\s*", src):
        rest = src[m.end():]
        ret = re.match(r"(\w+(?:\{[^}]*\})?)\s*,\s*\(", rest)
        assert ret, m.group(1)
        i, depth = ret.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(rest[i], 0)
            i += 1
        calls.append((m.group(1), ret.group(1), _split_top(rest[ret.end():i - 1])))
    return calls


JL_VAL = {"Int32": "int32_t", "Int64": "int64_t", "UInt64": "uint64_t", "UInt32": "uint32_t", "Float64": "double", "Cdouble": "double"}
JL_PTR = {"Float64": "double", "Int64": "int64_t", "Int32": "int32_t", "UInt64": "uint64_t", "Cvoid": "fsb_engine",
          "FsbParams": "fsb_params", "FsbConfig": "fsb_config"}


def _matches(jl, c):
    kind, base, depth = c
    if jl in JL_VAL:
        return kind == "val" and base == JL_VAL[jl]
    if jl == "Cstring":
        return kind == "ptr" and base == "char"
    m = re.match(r"(?:Ptr|Ref)\{(.*)\}$", jl)
    if not m or kind != "ptr":
        return False
    inner = m.group(1)
    m2 = re.match(r"Ptr\{(.*)\}$", inner)
    if m2:  # pointer to pointer
        return depth == 2 and JL_PTR.get(m2.group(1)) == base
    return depth == 1 and JL_PTR.get(inner) == base


def test_every_ccall_matches_its_prototype():
    protos = _c_prototypes()
    calls = _julia_ccalls()
    assert len(calls) >= 15
    assert {"fsb_create", "fsb_upload_state", "fsb_run", "fsb_download_state", "fsb_download_log", "fsb_stylised_facts",
            "fsb_comm_init_all", "fsb_stats_all"} <= {c[0] for c in calls}
    for name, ret, args in calls:
        assert name in protos, name
        cret, cargs = protos[name]
        assert _matches(ret, cret), (name, "return", ret, cret)
        assert len(args) == len(cargs), (name, args, cargs)
        for k, (ja, ca) in enumerate(zip(args, cargs)):
            assert _matches(ja, ca), (name, k, ja, ca)


def _c_struct(name):
    src = _strip_c(open(os.path.join(ROOT, "include", "fsb.h")).read())
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        typ, names = re.match(r"((?:const\s+)?\w+)\s+(.*)$", decl, flags=re.S).groups()
        for n in names.split(","):
            n = n.strip()
            fields.append((n.lstrip("* "), _c_class(typ + (" *" if n.startswith("*") else ""))))
    return fields


def _julia_struct(name):
    src = open(os.path.join(ROOT, "julia", "FundStabB200.jl")).read()
    body = re.search(r"struct %s\n(.*?)\nend" % name, src, flags=re.S).group(1)
    return [tuple(f.strip().split("::")) for line in body.splitlines() for f in line.split(";") if "::" in f]


def test_julia_structs_mirror_the_pods():
    for jl, c in (("FsbParams", "fsb_params"), ("FsbConfig", "fsb_config")):
        jf, cf = _julia_struct(jl), _c_struct(c)
        assert [n for n, _ in jf] == [n for n, _ in cf], (jl, jf, cf)
        for (n, jt), (_, ct) in zip(jf, cf):
            assert _matches(jt, ct), (jl, n, jt, ct)
