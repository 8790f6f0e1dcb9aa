"""Structural check of julia/RimuB200.jl against include/rimu_b200.h (Julia itself is not available in this image):

* every `ccall((:rimu_xxx, LIB), ...)` names a function the header declares, with the header's argument count, and the
  argument-type tuple has as many entries as values are passed;
* the three mirrored structs (HamDesc, StepParams, StepStats) have the C structs' fields, in order, with matching widths;
* no `...` placeholder bodies: every helper the module calls is defined in it (or is a known Rimu / Base name);
* the methods the FCIQMC driver needs from an AbstractDVec (Interfaces/dictvectors.jl:22-140, the list in SURVEY 8b) exist.
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "julia", "RimuB200.jl")
HEADER = os.path.join(ROOT, "include", "rimu_b200.h")


def _strip_comments_jl(src):
    out = []
    for line in src.splitlines():
        # drop '#' comments (no '#' occurs inside string literals of this file except in interpolation-free strings)
        in_str, res = False, []
        for ch in line:
            if ch == '"':
                in_str = not in_str
            if ch == "#" and not in_str:
                break
            res.append(ch)
        out.append("".join(res))
    return "\n".join(out)


def _balanced(s, start):
    """index just past the parenthesis group that opens at s[start] == '('"""
    depth = 0
    for i in range(start, len(s)):
        if s[i] in "([{":
            depth += 1
        elif s[i] in ")]}":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced")


def _split_top(s):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    last = "".join(cur).strip()
    if last:
        parts.append(last)
    return parts


def header_functions():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    fns = {}
    for m in re.finditer(r"\b(?:int|void|uint64_t|const char \*)\s*\**(rimu_\w+)\s*\(([^;{]*?)\)\s*;", src):
        name, args = m.group(1), m.group(2).strip()
        fns[name] = 0 if args in ("", "void") else len(_split_top(args))
    return fns


def shim_ccalls():
    src = _strip_comments_jl(open(SHIM).read())
    calls = []
    for m in re.finditer(r"ccall\(", src):
        end = _balanced(src, m.end() - 1)
        parts = _split_top(src[m.end():end - 1])
        sym = re.match(r"\(:(\w+),\s*LIB\)", parts[0]).group(1)
        argtypes = parts[2]
        assert argtypes.startswith("(") and argtypes.endswith(")"), (sym, argtypes)
        types = _split_top(argtypes[1:-1])
        calls.append((sym, parts[1], types, parts[3:]))
    return calls


def test_every_ccall_matches_the_header():
    fns = header_functions()
    calls = shim_ccalls()
    assert len(calls) >= 35
    for sym, ret, types, values in calls:
        assert sym in fns, f"{sym} is not declared in include/rimu_b200.h"
        assert len(types) == fns[sym], f"{sym}: header takes {fns[sym]} arguments, ccall declares {len(types)}"
        assert len(values) == len(types), f"{sym}: {len(types)} argument types but {len(values)} values"
        assert ret in ("Cint", "Cstring", "Cvoid", "UInt64"), (sym, ret)
    # the entry points of the path itself are all bound
    bound = {c[0] for c in calls}
    for must in ("rimu_ctx_create", "rimu_ctx_destroy", "rimu_ham_create", "rimu_ham_destroy", "rimu_vec_create", "rimu_vec_destroy",
                 "rimu_vec_upload", "rimu_vec_download", "rimu_vec_length", "rimu_vec_get", "rimu_vec_copy", "rimu_vec_clear",
                 "rimu_vec_scale", "rimu_vec_axpby", "rimu_vec_dot", "rimu_vec_dot_sparse", "rimu_vec_norm", "rimu_step",
                 "rimu_comm_unique_id", "rimu_comm_init", "rimu_comm_detach", "rimu_sizeof_ham_desc", "rimu_sizeof_step_params",
                 "rimu_sizeof_step_stats", "rimu_ham_diagonal", "rimu_ham_num_offdiagonals", "rimu_ham_offdiagonals"):
        assert must in bound, must


C_WIDTH = {"int32_t": 4, "int64_t": 8, "uint64_t": 8, "double": 8, "float": 4}
JL_WIDTH = {"Int32": 4, "Int64": 8, "UInt64": 8, "Float64": 8, "Float32": 4}


def c_struct_fields(name):
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    end = re.search(r"\}\s*" + name + r"\s*;", src).start()
    body = src[src.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
    consts = {"RIMU_MAX_MODES": 128, "RIMU_MAX_TABLE_MODES": 64, "RIMU_MAX_COMPONENTS": 4}
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, rest = decl.split(None, 1)
        for item in rest.split(","):
            item = item.strip()
            m = re.match(r"(\w+)(?:\[(.+)\])?$", item)
            count = 1
            if m.group(2):
                expr = m.group(2)
                for k, v in consts.items():
                    expr = expr.replace(k, str(v))
                count = eval(expr)
            fields.append((m.group(1).rstrip("_"), C_WIDTH[ctype], count))
    return fields


def jl_struct_fields(name):
    src = _strip_comments_jl(open(SHIM).read())
    body = re.search(r"\bstruct " + name + r"\b(.*?)\nend", src, flags=re.S).group(1)
    fields = []
    for line in body.splitlines():
        line = line.strip()
        if "::" not in line:
            continue
        fname, ftype = line.split("::")
        m = re.match(r"NTuple\{(\d+),\s*(\w+)\}", ftype)
        if m:
            fields.append((fname, JL_WIDTH[m.group(2)], int(m.group(1))))
        else:
            fields.append((fname, JL_WIDTH[ftype], 1))
    return fields


def test_mirrored_structs_have_the_c_layout():
    for cname, jname in (("rimu_ham_desc", "HamDesc"), ("rimu_step_params", "StepParams"), ("rimu_step_stats", "StepStats"),
                         ("rimu_shift_params", "ShiftParams")):
        cf, jf = c_struct_fields(cname), jl_struct_fields(jname)
        assert [f[0] for f in cf] == [f[0] for f in jf], (cname, [f[0] for f in cf], [f[0] for f in jf])
        assert [(w, n) for _, w, n in cf] == [(w, n) for _, w, n in jf], cname
    # and the sizes agree with what the compiled library reports (natural alignment: no padding surprises)
    import rimu_b200 as R
    lib = R._lib.lib()
    for cname, fn in (("rimu_ham_desc", lib.rimu_sizeof_ham_desc), ("rimu_step_params", lib.rimu_sizeof_step_params),
                      ("rimu_step_stats", lib.rimu_sizeof_step_stats), ("rimu_shift_params", lib.rimu_sizeof_shift_params)):
        size, off = 0, 0
        for _, w, n in c_struct_fields(cname):
            off = (off + w - 1) // w * w
            off += w * n
        size = (off + 7) // 8 * 8
        assert size == fn(), (cname, size, fn())


KNOWN = set("""
ccall check get joinpath unsafe_string throw print zeros cld all map sum pointer enumerate onr set_bit! get_bit ntuple fieldtypes fieldtype Tuple Int
Ref finalizer max min length collect first last reduce vcat view isempty num_modes num_particles Int32 Int64 UInt64 UInt8 Float64
Float32 Cint pad pad3 size vec permutedims get! rand typeof eltype zero iterate similar copy copy! sizeof fieldcount Dict IdDict
Pair Symbol ArgumentError RimuB200Error Context GPUHam GPUDVec GPUWorkingMemory HamDesc StepParams StepStats FrozenDVec DVec
IsDeterministic IsStochasticInteger NonInitiator to_key from_key words desc base_desc addr_kind particles components gpu_ham
make_current table_slots resize_table! grow_exchange! create_vec val_type upload! download similar_empty global_length
style_params initiator_params compression_threshold step_stats_tuple default_style apply_operator! mul! dot norm pairs
working_memory last_error num_offdiagonals cos StochasticStyle zerovector! isa Returns something showerror Vector
""".split())


def test_no_placeholders_and_every_helper_is_defined():
    raw = open(SHIM).read()
    src = _strip_comments_jl(raw)
    # "..." is legal Julia only as splatting/slurping after an identifier, a call or inside a signature; a bare placeholder is not
    for m in re.finditer(r"\.\.\.", src):
        before = src[max(0, m.start() - 1)]
        assert before.isalnum() or before in ")_]}", f"placeholder '...' at offset {m.start()}: {src[max(0, m.start() - 40):m.end() + 10]!r}"
    defined = set(re.findall(r"^\s*(?:function\s+)?(?:Base\.|Rimu\.)?([A-Za-z_]\w*!?)\s*(?:\{[^}]*\})?\(", src, flags=re.M))
    defined |= set(re.findall(r"^\s*(?:mutable\s+)?struct\s+(\w+)", src, flags=re.M))
    called = set(re.findall(r"(?<![\w.:@])([a-z_]\w*!?)\(", src))
    unknown = {c for c in called if c not in defined and c not in KNOWN}
    assert not unknown, f"helpers used but never defined: {sorted(unknown)}"


def test_driver_facing_methods_exist():
    src = _strip_comments_jl(open(SHIM).read())
    for sig in ("function apply_operator!(wm::GPUWorkingMemory, target::GPUDVec, source::GPUDVec, op, boost=1)",
                "working_memory(v::GPUDVec", "zerovector(v::GPUDVec)", "zerovector!(v::GPUDVec)", "walkernumber(v::GPUDVec)",
                "walkernumber_and_length(v::GPUDVec)", "Base.length(v::GPUDVec)", "StochasticStyle(v::GPUDVec)", "localpart(v::GPUDVec)",
                "freeze(v::GPUDVec)", "function dot(f::FrozenDVec, v::GPUDVec", "function scale!(v::GPUDVec", "function add!(y::GPUDVec",
                "function dot(x::GPUDVec, y::GPUDVec)", "function norm(v::GPUDVec", "function mul!(y::GPUDVec", "Base.copy!(dst::GPUDVec",
                "Base.deepcopy(v::GPUDVec)", "Base.pairs(v::GPUDVec)", "Base.getindex(v::GPUDVec"):
        assert sig in src, sig
    # the statistics names the styles report (styles.jl:14-20, 94-96, 203-209; compression.jl:16)
    assert "(:spawn_attempts, :spawns, :deaths, :clones, :zombies)" in src
    assert "(:exact_steps, :inexact_steps, :spawn_attempts, :spawns)" in src and ":len_before" in src
