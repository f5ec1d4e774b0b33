"""What the built library is made of, checked on the CPU with cuobjdump (no device needed).

The path is HBM-bound Level-1 work: the kernels must be sm_100a code that moves 16- and 32-byte packs
(LDG/STG .128 / .256), combines with warp shuffles, overlaps kernel boundaries with programmatic dependent launch
(PREEXIT = griddepcontrol.launch_dependents, ACQBULK = griddepcontrol.wait) -- and uses neither tensor cores nor
TMA, because nothing on the path is a contraction or reuses a tile."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lambda_blas_b200", "liblbk.so")


@pytest.fixture(scope="module")
def sass():
    if not shutil.which("cuobjdump") or not os.path.exists(LIB):
        pytest.skip("cuobjdump or liblbk.so not available")
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur].append(m.group(1))
    return out, kernels


def test_built_for_sm_100a(sass):
    out, kernels = sass
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs
    assert len(kernels) > 300          # map / reduce x routines x shapes x {float, double}


def test_streaming_kernels_use_wide_accesses_and_pdl(sass):
    _, kernels = sass
    unit = {k: v for k, v in kernels.items() if "map_unit_kernel" in k or "reduce_unit_kernel" in k}
    assert unit
    for name, ins in unit.items():
        wide = [i for i in ins if re.match(r"(LDG|STG)\.E.*\.(128|256)", i)]
        scalar_shape = re.search(r"ELi1ELi(4|8|16)E", name) is not None      # the any-alignment shapes (one element per access)
        assert wide or scalar_shape, name
        assert "PREEXIT" in ins and "ACQBULK" in ins, name           # releases its dependents, waits for its predecessors
    reduces = [v for k, v in kernels.items() if "reduce_unit_kernel" in k]
    assert all(any(i.startswith("SHFL.BFLY") for i in ins) and any(i.startswith("ATOMG") for i in ins) for ins in reduces)


def test_no_tensor_cores_no_tma(sass):
    _, kernels = sass
    banned = re.compile(r"^(HMMA|IMMA|DMMA|QGMMA|HGMMA|UTC\w*MMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM)")
    for name, ins in kernels.items():
        hit = [i for i in ins if banned.match(i)]
        assert not hit, (name, hit[:3])


def test_default_shifted_load_shapes_do_not_spill():
    """The shifted-load kernels (lbk_kernels.cuh, shift_pack) have no local arrays, so a stack frame means register
    spills -- which cost these shapes 4-20 points of bandwidth (profiles/r02_misaligned.md).  The shapes the library
    picks by default (maps: 16-byte packs x 4; dot / sdsdot: x 4 for Float, x 2 for Double) must have none."""
    if not shutil.which("cuobjdump") or not os.path.exists(LIB):
        pytest.skip("cuobjdump or liblbk.so not available")
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    seen = 0
    for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", out):
        name, usage = m.group(1), m.group(2)
        k = re.search(r"(map|reduce)_unit_kernelINS_\d+(\w+?)(I[fd]E)?ELi(\d)ELi(\d)ELi\dELb0ELi([123])E", name)
        if not k:
            continue
        kind, vec, unroll = k.group(1), int(k.group(4)), int(k.group(5))
        default = unroll == 4 if (kind == "map" or vec == 4) else unroll == 2
        if default:
            seen += 1
            assert re.search(r"STACK:0\b", usage), (name, usage)
    assert seen >= 3 * 4 + 3 + 1 + 3          # axpy / copy / scal x (3 Float shifts + 1 Double), dot Float x 3 + Double, sdsdot x 3


def test_default_kernels_do_not_spill():
    """Every shape kDefaultCfg (lbk_lib.cu) picks for the unit-stride routines, Float and Double: no stack frame.  (A change to
    the kernel templates once cost idamax 4 % through a 12-byte spill that no functional test could see.)"""
    if not shutil.which("cuobjdump") or not os.path.exists(LIB):
        pytest.skip("cuobjdump or liblbk.so not available")
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {m.group(1): m.group(2) for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", out)}
    # (kernel, functor / op, float packs, unroll, policy) as in kDefaultCfg; Double instances halve the pack
    defaults = [("reduce", "5DotOp", 8, 4, 1), ("reduce", "6AsumOp", 8, 2, 1), ("reduce", "6Nrm2Op", 8, 2, 1), ("reduce", "7IamaxOp", 8, 2, 1),
                ("map", "5AxpyF", 8, 2, 6), ("map", "5ScalF", 4, 2, 6), ("map", "5CopyF", 4, 2, 6), ("map", "5SwapF", 4, 1, 6),
                ("map", "4RotF", 4, 1, 6), ("map", "6Rotm0F", 4, 1, 6), ("map", "9RotmNeg1F", 4, 1, 6), ("map", "6Rotm1F", 4, 1, 6),
                ("map", "5FillF", 8, 4, 1)]
    seen = 0
    for kind, who, vec, unroll, policy in defaults:
        for ty, v in (("f", vec), ("d", vec // 2)):
            pat = f"{kind}_unit_kernelINS_{who}I{ty}EELi{v}ELi{unroll}ELi{policy}ELb0ELi0E"
            hits = [n for n in usage if pat in n]
            assert len(hits) == 1, (pat, hits)
            assert re.search(r"STACK:0\b", usage[hits[0]]), (hits[0], usage[hits[0]])
            seen += 1
    hits = [n for n in usage if "reduce_unit_kernelINS_7DsdotOpELi8ELi2ELi1ELb0ELi0E" in n]
    assert len(hits) == 1 and re.search(r"STACK:0\b", usage[hits[0]])
    assert seen == 26


def test_shifted_load_kernels_read_aligned_packs_and_shuffle(sass):
    """lbk_kernels.cuh, shift_pack: 16-byte loads of x plus lane shuffles -- down for a shift of at most half a pack, up beyond"""
    _, kernels = sass
    shifted = {k: v for k, v in kernels.items() if re.search(r"_unit_kernelI.*ELb0ELi[123]E", k)}
    assert len(shifted) >= 38
    for name, ins in shifted.items():
        assert any(re.match(r"LDG\.E.*\.128", i) for i in ins), name
        shift = int(re.search(r"ELb0ELi([123])E", name).group(1))
        vec = int(re.search(r"ELi(\d)ELi\dELi\dELb0", name).group(1))
        want = "SHFL.DOWN" if 2 * shift <= vec else "SHFL.UP"
        assert any(i.startswith(want) for i in ins), (name, want)
