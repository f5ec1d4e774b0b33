"""hs/ -- the Haskell binding layer of INTEGRATION.md -- has never met a compiler (no GHC in this image).  What can be
checked without one, with the parser of oracle/hs_eval.py extended to whole modules (tests/hs_lint.py):
the sources parse, every name they use is in scope, foreign imports and class methods are applied to as many arguments
as their signatures have, instances are complete, export lists only name things that exist.  Not a type check."""
import glob
import os

import pytest

from oracle import hs_eval as H
from tests import hs_lint as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = sorted(glob.glob(os.path.join(ROOT, "hs", "Numerical", "BLAS", "*.hs")) + glob.glob(os.path.join(ROOT, "hs", "Numerical", "BLAS", "Device", "*.hs"))) + \
    sorted(glob.glob(os.path.join(ROOT, "hs", "test", "*.hs")))
REF = "/root/reference"
# constructors and fields behind `GivensRot(..)`, `ModGivensRot(..)` (src/Numerical/BLAS/Types.hs:8,15-20 of the reference)
REFERENCE_TYPES = {"GivensRot": ["GIVENSROT"],
                   "ModGivensRot": ["FLAGNEG2", "FLAGNEG1", "FLAG0", "FLAG1", "d1", "d2", "x1", "h11", "h12", "h21", "h22"]}


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is only present in the build container")
def test_the_parser_accepts_every_module_of_the_reference():
    """GHC compiles these; a parser that is to vouch for hs/ must at least take them"""
    files = sorted(glob.glob(REF + "/src/Numerical/*.hs") + glob.glob(REF + "/src/Numerical/BLAS/*.hs") + glob.glob(REF + "/test/*.hs"))
    assert len(files) == 8
    for f in files:
        p, decls, datas = L.parse_file(f)
    p, _, _ = L.parse_file(REF + "/test/Foreign.hs")
    assert len(p.foreign) == 50 and p.foreign["sasum_foreign"] == {"entity": "sasum_", "arity": 3}          # Foreign.hs:32-81


@pytest.mark.parametrize("src,why", [
    ("f x = (x + 1", "unbalanced parenthesis"),
    ("f x = do\n    y <- g x\n", "a do block must end in an expression"),
    ("f x = case x\n    1 -> 2\n", "case without of"),
    ("f x = let y = 1\n", "let without in"),
    ("f = if a then b\n", "if without else"),
    ("f x = x +\n", "dangling operator"),
])
def test_the_parser_rejects_broken_sources(src, why):
    with pytest.raises(H.HsError):
        p = L.FullParser(H.layout(H.lex("module T where\n" + src)), why)
        p.parse_module()
        if not p.is_("special", "}") and not p.is_("eof"):
            raise H.HsError("trailing tokens")


def test_shim_modules_exist():
    assert [os.path.basename(f) for f in SHIM] == ["Device.hs", "Types.hs", "DeviceMain.hs"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is only present in the build container")
def test_member_lists_of_the_references_modules_are_its_export_lists():
    """hs_lint.MEMBERS states what `import Numerical.BLAS.Single` / `import Gen` bring into scope: hold that to the reference"""
    p, _, _ = L.parse_file(REF + "/src/Numerical/BLAS/Single.hs")
    assert sorted(n for n, _ in p.exports) == sorted(L.MEMBERS["Numerical.BLAS.Single"])
    g, gd, _ = L.parse_file(REF + "/test/Gen.hs")
    defined = set(L.decl_names(gd))
    assert set(L.MEMBERS["Gen"]) <= defined, set(L.MEMBERS["Gen"]) - defined


def test_device_suite_applies_the_device_routines_to_as_many_arguments_as_they_take():
    """hs/test/DeviceMain.hs calls Numerical.BLAS.Device qualified as D: every D.f is applied to nothing (passed on) or to exactly
    the number of arguments of f's signature in Device.hs"""
    dev, _, _ = L.parse_file(os.path.join(ROOT, "hs", "Numerical", "BLAS", "Device.hs"))
    p, decls, _ = L.parse_file(os.path.join(ROOT, "hs", "test", "DeviceMain.hs"))
    bad, seen = [], set()

    def visit(name, nargs):
        if name.startswith("D."):
            f = name[2:]
            seen.add(f)
            if f in dev.sigs and nargs not in (0, dev.sigs[f]):
                bad.append((name, nargs, dev.sigs[f]))
    for d in decls:
        L.applications(d[3] if d[0] == "fun" else d[2], visit)
        L.applications(d[4] if d[0] == "fun" else d[3], visit)
    assert not bad, bad
    assert {"toDevice", "fromDevice", "dot", "iamax", "scal", "copy", "swap", "rot", "axpy", "withContext"} <= seen, seen


def module_scope(path):
    """(parser, declarations, names defined at top level, names visible through imports)"""
    p, decls, datas = L.parse_file(path)
    defined = set(L.decl_names(decls)) | set(p.foreign) | set(datas) | set(p.types)
    for con, info in datas.items():
        defined |= set(info["fields"] or [])
    for c in p.classes.values():
        defined |= set(c["methods"])
    defined |= set(p.classes)
    visible, aliases = set(), {}
    for imp in p.imports:
        mod = imp["module"]
        members = None
        own = os.path.join(ROOT, "hs", *mod.split(".")) + ".hs"
        if os.path.exists(own):                                   # a module of this repository: its export list
            q, d2, dd2 = L.parse_file(own)
            members = set()
            for name, sub in (q.exports or []):
                members.add(name)
                if sub:
                    members |= set(dd2) | {f for info in dd2.values() for f in (info["fields"] or [])}
        elif mod in L.MEMBERS:
            members = set(L.MEMBERS[mod])
        if imp["names"] is not None and not imp["hiding"]:
            got = set()
            for name, sub in imp["names"]:
                got.add(name)
                if sub:
                    got |= set(REFERENCE_TYPES.get(name, []))
            members = got if members is None else got & (members | got)
        if members is None:
            raise AssertionError(f"{path}: import of {mod} without an import list and no member list in hs_lint.MEMBERS")
        aliases.setdefault(imp["alias"], set()).update(members)
        if not imp["qualified"]:
            visible |= members
    return p, decls, defined, visible, aliases


@pytest.mark.parametrize("path", SHIM, ids=[os.path.relpath(f, ROOT) for f in SHIM])
def test_every_name_is_in_scope(path):
    p, decls, defined, visible, aliases = module_scope(path)
    used = set()
    L.decls_used(decls, set(), used)
    for inst in p.instances:
        L.decls_used(inst["decls"], set(), used)
    unknown = []
    for name in sorted(used):
        if name.startswith("("):                               # operators: Prelude / base combinators
            continue
        if "." in name and name[0].isupper() and not name.endswith("."):
            alias, member = name.rsplit(".", 1)
            if alias not in aliases or member not in aliases[alias]:
                unknown.append(name)
            continue
        if name not in defined and name not in visible and name not in L.PRELUDE:
            unknown.append(name)
    assert not unknown, f"not in scope in {os.path.basename(path)}: {unknown}"


@pytest.mark.parametrize("path", SHIM, ids=[os.path.relpath(f, ROOT) for f in SHIM])
def test_foreign_imports_and_methods_are_fully_applied(path):
    p, decls, defined, visible, aliases = module_scope(path)
    arity = {n: f["arity"] for n, f in p.foreign.items()}
    for c in p.classes.values():
        arity.update(c["methods"])
    bad = []

    def visit(name, nargs):
        if name in arity and nargs not in (0, arity[name]):
            bad.append((name, nargs, arity[name]))
    # (locally bound names that shadow a foreign import would confuse this; the shim has none: checked by naming convention)
    for d in decls:
        L.applications(d[3] if d[0] == "fun" else d[2], visit)
        L.applications(d[4] if d[0] == "fun" else d[3], visit)
    assert not bad, f"(name, arguments given, arguments in the signature): {bad}"


def test_instances_define_every_method_of_their_class():
    for path in SHIM:
        p, *_ = L.parse_file(path)
        for inst in p.instances:
            cls = p.classes.get(inst["class"])
            if cls is None:
                continue                                        # a class from elsewhere (Show, Exception, NFData ...)
            have = L.decl_names(inst["decls"])
            missing = set(cls["methods"]) - have - cls["defaults"]
            extra = have - set(cls["methods"])
            assert not missing and not extra, (inst["class"], inst["head"], sorted(missing), sorted(extra))
            for d in inst["decls"]:                             # a method bound to a foreign import of a different arity cannot type-check
                if d[0] == "fun" and not d[2] and len(d[3]) == 1 and d[3][0][1][0] == "var":
                    target = d[3][0][1][1]
                    if target in p.foreign:
                        assert p.foreign[target]["arity"] == cls["methods"][d[1]], (d[1], target)


@pytest.mark.parametrize("path", SHIM, ids=[os.path.relpath(f, ROOT) for f in SHIM])
def test_export_lists_name_things_that_exist(path):
    p, decls, defined, visible, aliases = module_scope(path)
    missing = [n for n, _ in (p.exports or []) if n not in defined and n not in visible and n != ".."]
    assert not missing, missing
    # every top-level function has a type signature (the reference builds with -Wall -Werror, lambda-blas.cabal:48)
    unsigned = [d[1] for d in decls if d[0] == "fun" and d[1] not in p.sigs]
    assert not unsigned, unsigned


@pytest.mark.skipif(not os.path.isdir(REF) or not __import__("shutil").which("patch"), reason="needs /root/reference and patch(1)")
def test_patches_apply_to_the_reference_as_it_lies_there(tmp_path):
    """hs/patches/*.patch are what INTEGRATION.md tells a maintainer to apply: they must apply cleanly (no fuzz, no offset)"""
    import shutil
    import subprocess
    (tmp_path / "src" / "Numerical" / "BLAS").mkdir(parents=True)
    shutil.copy(REF + "/lambda-blas.cabal", tmp_path / "lambda-blas.cabal")
    shutil.copy(REF + "/src/Numerical/BLAS/Types.hs", tmp_path / "src" / "Numerical" / "BLAS" / "Types.hs")
    for name in ("lambda-blas.cabal.patch", "Types.hs.patch"):
        r = subprocess.run(["patch", "-p1", "--dry-run", "-i", os.path.join(ROOT, "hs", "patches", name)], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0 and "fuzz" not in r.stdout and "offset" not in r.stdout, (name, r.stdout, r.stderr)
