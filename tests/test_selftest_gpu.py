"""Runs the standalone GEMM self-test (vln-bevbert_b200/csrc/selftest_gemm.cu, built by `make` / __graft_entry__.build())
inside the GPU test tier: every operand-major combination, tails, split-K, the attention-shaped batched products and the
FULL-SIZE bench shapes (14112 x 2304 x 768, 14112 x 3072 x 768 with GELU + second output, 8192^3, 768 x 3072 x 14112
split-K ...) are compared element by element with an fp32 reference kernel, through the C ABI (bb_gemm_bf16)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(os.path.dirname(HERE), "vln-bevbert_b200", "csrc", "build", "selftest_gemm")


@pytest.mark.gpu
def test_gemm_selftest_all_cases_pass():
    if not os.path.exists(BIN):
        pytest.skip("selftest_gemm not built (run `make -C vln-bevbert_b200/csrc`)")
    out = subprocess.run([BIN, "all"], capture_output=True, text=True, timeout=600)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("CASE")]
    assert len(lines) >= 26, out.stdout[-2000:] + out.stderr[-2000:]
    bad = [ln for ln in lines if " PASS " not in ln]
    assert not bad and out.returncode == 0, "\n".join(bad) + out.stderr[-2000:]
