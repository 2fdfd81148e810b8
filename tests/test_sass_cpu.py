"""Static evidence that the hot kernel is what DESIGN.md says it is (no GPU needed): the built
library's SASS for both shapes of the update kernel contains the FP64 tensor-core instruction
(DMMA.8x8x4), the TMA-engine bulk copy (UBLKCP) with mbarrier completion (SYNCS), no local-memory
spills, and register/occupancy figures that allow 2 CTAs per SM for the update shape."""
import os
import re
import shutil
import subprocess

import pytest

from dense_linear_app_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")


@pytest.fixture(scope="module")
def sass():
    out = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and "/*" in line:
            funcs[cur].append(line)
    return funcs


@pytest.fixture(scope="module")
def res_usage():
    out = subprocess.run([CUOBJDUMP, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            res[cur] = {k: int(v) for k, v in re.findall(r"([A-Z]+)(?:\[0\])?:(\d+)", line)}
    return res


def gemm_kernels(d):
    return {k: v for k, v in d.items() if "gemm_nt_dmma_kernel" in k}


def test_update_kernel_uses_fp64_tensor_cores_and_tma(sass):
    ks = gemm_kernels(sass)
    assert len(ks) == 2                                  # GemmCfg<64,4,2> and GemmCfg<128,6,1>
    for name, lines in ks.items():
        text = "\n".join(lines)
        assert text.count("DMMA.8x8x4") == 128, name     # 4 k4-steps x 32 accumulators, fully unrolled slab
        assert "UBLKCP" in text, name                    # cp.async.bulk on the TMA engine
        assert "SYNCS.ARRIVE.TRANS64" in text and "TRYWAIT" in text, name   # mbarrier expect_tx / try_wait
        assert "LDS.128" in text, name
        assert not re.search(r"\b(LDL|STL)\b", text), f"{name}: local-memory traffic in the hot kernel"


def test_update_kernel_resources_allow_two_ctas_per_sm(res_usage):
    ks = gemm_kernels(res_usage)
    pair = next(v for k, v in ks.items() if "Li64ELi4ELi2E" in k)
    wide = next(v for k, v in ks.items() if "Li128ELi6ELi1E" in k)
    for r in (pair, wide):
        assert r["STACK"] == 0 and r["LOCAL"] == 0
    assert pair["REG"] * 160 * 2 <= 65536                # 5 warps x 2 CTAs fit the register file
    assert wide["REG"] * 288 <= 65536


def test_diag_kernel_has_bare_shuffles(sass):
    """The warp-0 branch of the diagonal-block kernel is convergent for the compiler: shuffles are not
    wrapped in WARPSYNC/ENDCOLLECTIVE (1056 pairs and 30k SASS lines before that fix)."""
    lines = next(v for k, v in sass.items() if "potrf_diag_kernel" in k)
    text = "\n".join(lines)
    assert text.count("SHFL") >= 1000
    assert text.count("WARPSYNC") <= 4 and len(lines) < 20000
