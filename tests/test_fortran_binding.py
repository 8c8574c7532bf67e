"""The Fortran side of the boundary (fortran/kestrel_gpu.f90) against the C header, mechanically.

No Fortran compiler exists in this image, so the ISO_C_BINDING module cannot be compiled here.  What can be
checked is what breaks silently at link time: every public entry point of include/kestrel_gpu.h must have an
interface with the same binding name, the same number of arguments, the same by-value / by-reference
passing and interoperable types; the bind(C) derived types must list the structs' members in the header's
order with interoperable types; the named constants must carry the header's enum values."""
import os
import re

from common import ROOT

HDR = os.path.join(ROOT, "include", "kestrel_gpu.h")
F90 = os.path.join(ROOT, "fortran", "kestrel_gpu.f90")
HOST = os.path.join(ROOT, "fortran", "IntegrateTo_gpu.f90")

SCALARS = {"int": "integer(c_int)", "int32_t": "integer(c_int32_t)", "int64_t": "integer(c_int64_t)", "double": "real(c_double)"}


def c_source():
    txt = open(HDR).read()
    return re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def c_prototypes():
    """name -> (return type, [(type, name)])"""
    out = {}
    for m in re.finditer(r"^([A-Za-z_][\w \*]*?)\b(kgpu_\w+)\s*\(([^;{}]*?)\)\s*;", c_source(), flags=re.M | re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if ret.startswith("typedef"):
            continue
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                params.append((mm.group(1).strip(), mm.group(2)))
        out[name] = (ret, params)
    return out


def c_structs():
    out = {}
    for m in re.finditer(r"typedef struct (\w+)\s*\{(.*?)\}\s*\w+\s*;", c_source(), flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            mm = re.match(r"((?:const )?[\w]+(?: \*)?)\s*(.*)$", decl)
            ctype, names = mm.group(1), mm.group(2)
            for nm in names.split(","):
                nm = nm.strip()
                ptr = ctype.endswith("*") or nm.startswith("*")
                fields.append((("ptr" if ptr else ctype), nm.lstrip("*").strip()))
        out[m.group(1)] = fields
    return out


def c_enums():
    vals = {}
    for m in re.finditer(r"enum\s*\{(.*?)\}\s*;", c_source(), flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = [s.strip() for s in item.split("=")]
                nxt = int(v)
            else:
                k = item
            vals[k] = nxt
            nxt += 1
    return vals


def f90_lines(path):
    """Source lines with comments stripped and continuations joined."""
    out, cur = [], ""
    for raw in open(path):
        line = raw.split("!")[0].rstrip()
        if not line.strip():
            continue
        line = line.strip()
        if line.startswith("&"):
            line = line[1:].lstrip()
        if line.endswith("&"):
            cur += line[:-1].rstrip() + " "
            continue
        out.append(cur + line)
        cur = ""
    return out


def f90_interfaces():
    """binding name -> (result decl, [(decl, has_value)] in dummy order)"""
    lines = f90_lines(F90)
    out = {}
    i = 0
    while i < len(lines):
        m = re.match(r"function (\w+)\s*\(([^)]*)\)\s*bind\(C, name=\"(\w+)\"\)\s*result\((\w+)\)", lines[i])
        if not m:
            i += 1
            continue
        fname, dummies, cname, res = m.group(1), [d.strip() for d in m.group(2).split(",") if d.strip()], m.group(3), m.group(4)
        assert fname == cname, (fname, cname)
        decls = {}
        i += 1
        while not lines[i].startswith("end function"):
            mm = re.match(r"(.*?)::\s*(.*)$", lines[i])
            if mm and not lines[i].startswith("import"):
                spec = mm.group(1).strip()
                for nm in mm.group(2).split(","):
                    nm = re.sub(r"\(.*\)", "", nm).strip()
                    decls[nm] = spec
            i += 1
        args = []
        for d in dummies:
            spec = decls[d]
            base = spec.split(",")[0].strip()
            args.append((base, "value" in [s.strip() for s in spec.split(",")[1:]]))
        out[cname] = (decls[res].split(",")[0].strip(), args)
    return out


def f90_types():
    lines = f90_lines(F90)
    out, cur = {}, None
    for l in lines:
        m = re.match(r"type, bind\(C\) :: (\w+)", l)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        if cur and l.startswith("end type"):
            cur = None
            continue
        if cur:
            spec, names = [s.strip() for s in l.split("::")]
            for nm in names.split(","):
                out[cur].append((spec, nm.strip()))
    return out


def f90_constants():
    vals = {}
    for l in f90_lines(F90):
        m = re.match(r"integer\(c_int(?:32_t)?\), parameter :: (.*)$", l)
        if m:
            for item in m.group(1).split(","):
                k, v = [s.strip() for s in item.split("=")]
                vals[k] = int(v)
    return vals


def test_every_public_entry_point_has_a_matching_interface():
    protos, ifaces = c_prototypes(), f90_interfaces()
    assert len(protos) >= 20
    assert sorted(protos) == sorted(ifaces), (sorted(set(protos) - set(ifaces)), sorted(set(ifaces) - set(protos)))
    for name, (ret, params) in protos.items():
        fres, fargs = ifaces[name]
        # result
        if "*" in ret:
            assert fres == "type(c_ptr)", (name, ret, fres)
        else:
            assert fres == SCALARS[ret.replace("const ", "")], (name, ret, fres)
        assert len(params) == len(fargs), (name, len(params), len(fargs))
        for (ctype, cname), (ftype, by_value) in zip(params, fargs):
            base = ctype.replace("const ", "").strip()
            if "*" not in base:                       # scalar by value
                assert by_value and ftype == SCALARS[base], (name, cname, ctype, ftype, by_value)
            elif by_value:                            # a pointer passed as an address
                assert ftype == "type(c_ptr)", (name, cname, ctype, ftype)
            else:                                     # a pointer expressed as a by-reference dummy
                pointee = base.replace("*", "").strip()
                if base.count("*") == 2 or pointee == "void":
                    assert ftype == "type(c_ptr)", (name, cname, ctype, ftype)
                elif pointee in SCALARS:
                    assert ftype == SCALARS[pointee], (name, cname, ctype, ftype)
                else:
                    assert ftype == f"type({pointee})", (name, cname, ctype, ftype)


def test_derived_types_mirror_the_structs():
    cs, ft = c_structs(), f90_types()
    for sname in ("kgpu_source", "kgpu_params", "kgpu_step_info"):
        cf, ff = cs[sname], ft[sname]
        assert [n.lstrip("_") for _, n in cf] == [n for _, n in ff], sname
        for (ctype, cname), (ftype, _) in zip(cf, ff):
            ctype = ctype.replace("const ", "")
            if ctype == "ptr" or ctype == "void":
                assert ftype == "type(c_ptr)", (sname, cname, ftype)
            elif ctype == "kgpu_heights_fn":
                assert ftype == "type(c_funptr)", (sname, cname, ftype)
            else:
                assert ftype == SCALARS[ctype], (sname, cname, ctype, ftype)


def test_named_constants_carry_the_enum_values():
    ce, fc = c_enums(), f90_constants()
    assert len(ce) > 50
    for k, v in ce.items():
        assert fc.get(k) == v, (k, v, fc.get(k))


def test_host_module_uses_only_declared_entry_points():
    """IntegrateTo_gpu.f90 calls nothing the interface module does not declare, and covers the call sequence
    create -> upload_tile -> integrate_to -> active_tiles -> download_tile -> destroy."""
    txt = "\n".join(f90_lines(HOST))
    used = set(re.findall(r"\b(kgpu_[a-z_0-9]+)\s*\(", txt))
    declared = set(f90_interfaces()) | {"kgpu_c_string"}
    assert used <= declared, used - declared
    for need in ("kgpu_create", "kgpu_upload_tile", "kgpu_integrate_to", "kgpu_active_tiles", "kgpu_download_tile", "kgpu_destroy"):
        assert need in used, need
    # every field of kgpu_params is assigned by MakeParams
    fields = [n for _, n in f90_types()["kgpu_params"]]
    missing = [f for f in fields if not re.search(r"\bp%" + f + r"\s*=", txt)]
    assert not missing, missing
    # every enum constant the host module selects exists
    for k in set(re.findall(r"\b(KGPU_[A-Z0-9_]+)\b", txt)):
        assert k in f90_constants(), k
