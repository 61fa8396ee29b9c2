"""Properties of the built library that can be read without a GPU (B200_PROFILING.md "What proves a Blackwell-native kernel"):
the SASS of libfnst.so contains tcgen05 MMAs (UTCHMMA, incl. the 2-CTA form), TMEM loads (LDTM) and TMA tensor loads (UTMALDG),
and no legacy mma.sync path (HMMA); ptxas reports no register spills for the tensor-core kernels and the register budget the
two-blocks-per-SM InstanceNorm backward kernel is built for."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fast_neural_style_transfer_b200")


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_sass_shows_tcgen05_tmem_and_tma():
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(PKG, "libfnst.so")], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "EF_CUDA_SM100" in sass.upper()
    count = lambda pat: len(re.findall(pat, sass))
    assert count(r"\bUTCHMMA\b") >= 10                 # tcgen05.mma.kind::f16 (conv, wgrad / Gram, final_conv kernels)
    assert count(r"\bUTCHMMA\.2CTA\b") >= 1            # cta_group::2 CTA-pair form
    assert count(r"\bLDTM\b") >= 4                     # tcgen05.ld (TMEM -> registers in the epilogues)
    assert count(r"\bUTMALDG\.4D\b") >= 4              # TMA 4-D boxes = implicit im2col
    assert count(r"\bHMMA\b") == 0                     # no mma.sync / wmma path in the library
    assert count(r"\bHGMMA\b") == 0


def test_ptxas_reports():
    logs = glob.glob(os.path.join(PKG, "csrc", "*.ptxas.log"))
    if len(logs) < 5:
        pytest.skip("ptxas logs not present (run the build first)")
    text = "\n".join(open(f).read() for f in logs)
    entries = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'.*?(\d+) bytes spill stores, (\d+) bytes spill loads.*?Used (\d+) registers",
                         text, flags=re.S)
    by_name = {n: (int(st), int(ld), int(r)) for n, st, ld, r in entries}
    assert len(by_name) > 100
    for name, (st, ld, regs) in by_name.items():
        if any(k in name for k in ("conv_tc_kernel", "wgrad_tc_kernel", "finalconv_tc_kernel", "inorm_apply_kernel", "mt_adam", "mt_sqnorm",
                                   "mt_scale", "resize_to_tensor")):
            assert st == 0 and ld == 0, (name, st, ld)
    two_block = [v for n, v in by_name.items() if "inorm_bwd_reduce_kernelI6__half13__nv_bfloat16Li2E" in n]
    assert two_block and two_block[0][2] <= 128        # fits two 256-thread blocks per SM (65536 registers)
