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
