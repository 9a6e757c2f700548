"""The C-ABI library loads on a CPU-only box and exports every symbol include/mmb200.h declares;
the host-side parameter helpers (no GPU involved) agree with the oracle's."""
import ctypes as C
import os
import re

import numpy as np

import helpers as H
import materialmodels_jl_b200._abi as abi
import materialmodels_jl_b200._lib as L


def declared_symbols():
    hdr = open(os.path.join(H.ROOT, "include", "mmb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mmh?_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_all_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 20
    lib = C.CDLL(L.SO_PATH)
    for n in names:
        assert hasattr(lib, n), "libmmb200.so does not export %s" % n
    # the ctypes mirror covers exactly the declared set
    assert sorted(abi.SIGNATURES) == names


def test_struct_sizes_match_header():
    assert C.sizeof(abi.MMLinearElastic) == 16 and C.sizeof(abi.MMPlastic) == 56 and C.sizeof(abi.MMHyper) == 32
    assert C.sizeof(abi.MMXuNeedleman) == 40 and C.sizeof(abi.MMNewtonOpts) == 16
    assert C.sizeof(abi.MMCrystal) == 80 + 8 + 12 * 9 * 8 + 12 * 12 * 8


def test_version_and_no_device_is_reported_not_fatal():
    lib = L.load()
    assert b"sm_100a" in lib.mm_version()
    assert lib.mm_device_count() >= 0


def test_bad_arguments_return_error_codes():
    lib = L.load()
    p = abi.MMLinearElastic(200e3, 0.3)
    assert lib.mm_linear_elastic(C.byref(p), None, None, None, 10, 0, None) == abi.MM_ERR_ARG
    assert b"bad argument" in lib.mm_last_error()
    assert lib.mm_linear_elastic(C.byref(p), None, None, None, 0, 0, None) == 0          # empty batch is a no-op
    x = abi.MMXuNeedleman(1, 1, 1, 1, 2)
    buf = (C.c_double * 8)()
    assert lib.mm_xu_needleman(C.byref(x), buf, buf, buf, 4, 0, 0, None) == abi.MM_ERR_ARG  # dim must be 2 or 3
    c = abi.MMCrystal()
    c.nslip = 7
    assert lib.mm_crystal(C.byref(c), buf, 1.0, buf, buf, buf, buf, None, None, None, None, 1, 0, None) == abi.MM_ERR_ARG


def test_elastic_tangent_helper_matches_oracle(oracle, mm):
    assert np.array_equal(mm.elastic_tangent_3D(200e3, 0.3), oracle.elastic_tangent_3d(200e3, 0.3))


def test_slipsystems_and_crystal_setup_match_oracle(oracle, mm):
    for structure, code in ((mm.FCC(), 0), (mm.BCC(), 1)):
        ss = mm.slipsystems(structure, (-0.665267, 0.0203875, 1.08633))
        pl, di = oracle.slipsystems(code, (-0.665267, 0.0203875, 1.08633))
        assert np.allclose(np.array([s[0] for s in ss]), pl, rtol=0, atol=2e-16)
        assert np.allclose(np.array([s[1] for s in ss]), di, rtol=0, atol=2e-16)
        m = mm.CrystalViscoPlastic(slipsystems=ss, **mm.synth.CRYSTAL_PARAMS)
        ref = H.crystal_c(oracle, structure=code)
        assert np.allclose(np.array(m._c.MS), np.array(ref.MS), rtol=0, atol=3e-16)
        assert np.array_equal(np.array(m._c.H), np.array(ref.H))


def test_xu_setup_matches_reference_formulas(mm):
    Φn = mm.xu_needleman_Φₙ(400.0, 0.002)
    Φt = mm.xu_needleman_Φₜ(200.0, 0.002)
    assert np.isclose(mm.xu_needleman_σₘₐₓ(Φn, 0.002), 400.0) and np.isclose(mm.xu_needleman_τₘₐₓ(Φt, 0.002), 200.0)
    m = mm.XuNeedleman(σₘₐₓ=400.0, τₘₐₓ=200.0, Φₙ=Φn, Φₜ=Φt, Δₙˢ=0.01)
    out = abi.MMXuNeedleman()
    assert L.load().mm_xu_needleman_setup(400.0, 200.0, Φn, Φt, 0.01, C.byref(out)) == 0
    assert np.isclose(out.delta_n, 0.002) and np.isclose(out.delta_t, 0.002)
    assert out.delta_n == m.δₙ and out.delta_t == m.δₜ


def test_julia_binding_ccalls_match_the_header():
    """julia/MaterialModelsB200 cannot be executed here (no Julia in the image), so at least every `ccall` in it must name an
    exported symbol and pass the argument kinds (int / int64 / double / pointer) the C declaration has, in order."""
    import ctypes as C
    import os
    import re

    import materialmodels_jl_b200._abi as abi

    src = open(os.path.join(os.path.dirname(__file__), "..", "julia", "MaterialModelsB200", "src", "MaterialModelsB200.jl"), encoding="utf-8").read()
    calls = re.findall(r"ccall\(\(:(\w+),\s*libmmb200\),\s*(\w+),\s*\(([^()]*)\)", src)
    assert len(calls) >= 12

    def jl_kind(t):
        t = t.strip()
        if t in ("Cint", "Int32"):
            return "i32"
        if t == "Int64":
            return "i64"
        if t == "Cdouble":
            return "f64"
        if re.match(r"^(CuPtr|Ptr|Ref)\{", t):
            return "ptr"
        raise AssertionError("unknown Julia argument type %r" % t)

    def c_kind(t):
        if t is C.c_int:
            return "i32"
        if t is C.c_int64:
            return "i64"
        if t is C.c_double:
            return "f64"
        return "ptr"

    for sym, ret, args in calls:
        assert sym in abi.SIGNATURES, "Julia binding calls %s, which include/mmb200.h does not declare" % sym
        restype, argtypes = abi.SIGNATURES[sym]
        jl = [jl_kind(t) for t in args.split(",") if t.strip()]
        cc = [c_kind(t) for t in argtypes]
        assert jl == cc, "%s: Julia passes %s, C expects %s" % (sym, jl, cc)
        assert ret == ("Cint" if restype is C.c_int else ret)


def test_julia_binding_structs_match_the_header():
    """the Julia `struct MM*` mirrors are passed by reference to the C side: field ORDER, field types and total size must equal the
    header's (checked against the ctypes mirror, which test_header_and_ctypes... pins to include/mmb200.h)"""
    import ctypes as C
    import os
    import re

    import materialmodels_jl_b200._abi as abi

    src = open(os.path.join(os.path.dirname(__file__), "..", "julia", "MaterialModelsB200", "src", "MaterialModelsB200.jl"), encoding="utf-8").read()
    size = {"Cdouble": 8, "Int32": 4, "Cint": 4, "Int64": 8}
    found = {}
    for name, body in re.findall(r"^struct (MM\w+)\b;?(.*?)\bend$", src, re.M | re.S):
        fields = []
        for fname, ftype in re.findall(r"(\w+)::([\w{},]+)", body):
            m = re.match(r"NTuple\{(\d+),(\w+)\}", ftype)
            fields.append((fname, ftype, int(m.group(1)) * size[m.group(2)] if m else size[ftype]))
        found[name] = fields
    assert set(found) >= {"MMPlastic", "MMLinearElastic", "MMHyper", "MMXuNeedleman", "MMNewtonOpts", "MMTransverselyIsotropic", "MMCrystal"}
    alias = {"lambda": "lambda_", "pad": "pad_"}
    for name, fields in found.items():
        cs = getattr(abi, name)
        cfields = [(f[0], C.sizeof(f[1])) for f in cs._fields_]
        assert [(alias.get(n, n), sz) for n, _, sz in fields] == cfields, "%s: Julia fields %s, C fields %s" % (name, fields, cfields)
        assert sum(sz for _, _, sz in fields) == C.sizeof(cs)           # no padding holes on either side (all 8-byte aligned runs)


def test_header_structs_match_the_ctypes_mirror():
    """include/mmb200.h is the contract; _abi.py (and through it the Julia mirror above) must list the same fields in the same order"""
    import os
    import re

    import materialmodels_jl_b200._abi as abi

    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "mmb200.h"), encoding="utf-8").read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for name, body in re.findall(r"typedef struct (MM\w+)\s*\{(.*?)\}\s*\1;", hdr, re.S):
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ctype, rest = decl.split(None, 1)
            for part in rest.split(","):
                names.append(re.match(r"\s*(\w+)", part).group(1))
        cs = getattr(abi, name)
        alias = {"lambda": "lambda_"}
        assert [alias.get(n, n) for n in names] == [f[0] for f in cs._fields_], name


def test_tuning_switches_are_known_by_name():
    """mm_set_tuning (host-side A/B switches; results never depend on them): every documented name is accepted and restored, an unknown
    name is refused with -1 — no GPU needed."""
    names = [b"crystal_force_general", b"crystal_thresh", b"crystal_chunk", b"crystal_fixup_per_sm", b"crystal_variant",
             b"wrapped_j2_variant", b"wrapped_j2_th_outer", b"wrapped_j2_th_final", b"wrapped_j2_chunk",
             b"wrapped_crystal_variant", b"wrapped_crystal_slab", b"xu_soa_tiled", b"mmh_plastic_via_factored"]
    lib = L.load()
    for n in names:
        old = lib.mm_set_tuning(n, 7)
        assert lib.mm_set_tuning(n, old) == 7, n
    assert lib.mm_set_tuning(b"no_such_switch", 1) == -1
    hdr = open(os.path.join(H.ROOT, "include", "mmb200.h")).read()
    for n in names:
        assert n.decode() in hdr, "undocumented tuning switch %s" % n.decode()
