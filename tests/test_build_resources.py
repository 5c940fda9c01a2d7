"""Static resource budget of the kernels that are meant to SHARE an SM: registers are allocated
per SM sub-partition (16384 each, warps dealt round-robin), so whether a CTA of the concurrent stream fits beside a
recurrence CTA is decided by (warps per sub-partition) x (registers per thread) of both kernels.  Checked on the built
library with cuobjdump -- no GPU needed; a compiler or source change that silently grows a kernel past its budget would
otherwise only show up as a slower step."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "uncertainty-aware-multimodal-emotion-recognition_b200", "csrc", "libdeer_b200.so")
SUBPART_REGS = 16384


def _usage():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run([exe, "--dump-resource-usage", LIB], capture_output=True, text=True, timeout=300).stdout
    res = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            res[name] = int(m.group(1))
            name = None
    assert res, "no kernels found in the library"
    return res


def _alloc(regs):          # registers are allocated per warp in units of 8 per thread
    return (regs + 7) // 8 * 8 * 32


def test_recurrence_kernels_leave_room_for_a_concurrent_cta():
    use = _usage()
    bwd = {k: v for k, v in use.items() if "lstm_bwd_cluster_kernel" in k}
    assert bwd
    for k, r in bwd.items():
        # BPTT: 8 warps = 2 per sub-partition; >= 5120 registers must stay free there (two warps of an 80-register kernel)
        assert r <= 160, (k, r)
        assert SUBPART_REGS - 2 * _alloc(r) >= 2 * _alloc(80), (k, r)
    # training forward (one 16-column tile per CTA, 9 warps = 3 in the fullest sub-partition), with and without the
    # in-kernel input projection, whole-tile (check-free) instantiation: two warps of a 64-register kernel still fit
    fwd = {k: v for k, v in use.items() if "lstm_fwd_cluster_kernelILi16ELb1ELi1ELi1ELb1" in k}
    assert len(fwd) == 2, sorted(fwd)
    for k, r in fwd.items():
        assert SUBPART_REGS - 3 * _alloc(r) >= 2 * _alloc(64), (k, r)


def test_small_tf32_gemm_fits_beside_the_recurrence():
    use = _usage()
    small = {k: v for k, v in use.items() if "gemm_tf32_kernel" in k}
    assert small
    for k, r in small.items():      # 6 warps: two in the fullest sub-partition
        assert r <= 80, (k, r)
