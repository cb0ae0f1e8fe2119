"""CPU tier: the sm_100a objects really contain what DESIGN.md says they do (cuobjdump -sass of the in-tree build): FP64 tensor
instructions fed by TMA in the GEMM core, 32-byte LDG/STG in the layout kernels, bulk-tensor loads AND stores in the opt-in layout
path, and no library or fallback code path (no tcgen05 FP64 kind exists: the MMA must be DMMA)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "rest_tensors_b200", "csrc", "build")


def _sass(obj):
    if shutil.which("cuobjdump") is None:
        pytest.skip("no cuobjdump")
    path = os.path.join(BUILD, obj)
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    return subprocess.check_output(["cuobjdump", "-sass", path], text=True)


def _count(sass, pattern):
    return len(re.findall(pattern, sass))


def test_gemm_core_is_dmma_fed_by_tma():
    s = _sass("rb_gemm.o")
    assert "sm_100a" in s
    assert _count(s, r"\bDMMA\.8x8x4\b") > 5000          # six instantiations of the unrolled k loop
    assert _count(s, r"\bUTMALDG\.3D\b") >= 48           # cp.async.bulk.tensor.3d loads
    assert _count(s, r"\bSYNCS\b") > 100                 # mbarrier pipeline
    assert _count(s, r"USETMAXREG") >= 12                # setmaxnreg hand-over producer -> consumers
    assert _count(s, r"\bUTC[A-Z]*MMA\b") == 0           # tcgen05 has no FP64 kind
    assert _count(s, r"\bHMMA\b|\bIMMA\b") == 0


def test_layout_kernels_use_32_byte_accesses():
    s = _sass("rb_layout.o")
    assert _count(s, r"\bLDG\.E\.ENL2\.256\b") >= 50 and _count(s, r"\bSTG\.E\.ENL2\.256\b") >= 40


def test_opt_in_layout_path_is_bulk_tensor_both_ways():
    s = _sass("rb_layout_tma.o")
    assert _count(s, r"\bUTMALDG\.3D\b") >= 2 and _count(s, r"\bUTMASTG\.3D\b") >= 2


def test_library_links_no_math_library():
    so = os.path.join(ROOT, "rest_tensors_b200", "librest_b200.so")
    out = subprocess.check_output(["ldd", so], text=True)
    for lib in ("cublas", "cusolver", "cutlass", "nccl", "openblas", "lapack"):
        assert lib not in out.lower(), f"librest_b200.so must not link {lib} (NCCL is bound at run time with dlopen)"


def test_single_pass_dp_j_kernel_is_what_the_design_says():
    """rb_dpj.o: 16-byte LDGSTS (global -> shared without registers) tracked by commit groups, GPU-scope (L1-bypassing) loads / stores for
    the sentinel-valued exchange, a non-blocking named-barrier arrive for the hand-off to the exchange warps, and no atomics."""
    s = _sass("rb_dpj.o")
    assert _count(s, r"\bLDGSTS\.E\.BYPASS\.128\b") >= 100
    assert _count(s, r"\bLDGDEPBAR\b") >= 24 and _count(s, r"\bDEPBAR\.LE\b") >= 24
    assert _count(s, r"\bLDG\.E\.64\.STRONG\.GPU\b") >= 24 and _count(s, r"\bSTG\.E\.64\.STRONG\.GPU\b") >= 24
    assert _count(s, r"\bBAR\.ARV\b") >= 24
    assert _count(s, r"\bATOM[A-Z.]*\b|\bRED\.[A-Z.]*\b") == 0
