"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares;
the hot kernels carry no fused multiply-add (FP contract).  No compute call is made here."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "icp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(icp_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    for must in ("icp_get_lms", "icp_get_reps", "icp_rbc_construct", "icp_rbc_search", "icp_weights", "icp_mean",
                 "icp_mean_weighted", "icp_devs", "icp_sij", "icp_power_method", "icp_svd_solve", "icp_transform_quaternion",
                 "icp_transform_matrix", "icp_step_run", "icp_run", "icp_batch_register"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from icp_b200 import capi
    L = capi.lib()            # raises if the .so is missing: the product has no fallback
    for s in header_symbols():
        assert hasattr(L, s), f"libicp_b200.so does not export {s}"
        assert s in capi.SIGNATURES, f"capi.py has no signature for {s}"
    assert b"sm_100a" in L.icp_version()


def test_no_cuda_device_fails_loudly():
    """Without a GPU the library must refuse to create a context (no CPU fallback)."""
    import ctypes as C
    from icp_b200 import capi
    L = capi.lib()
    h = C.c_void_p()
    rc = L.icp_ctx_create(0, None, C.byref(h))
    if rc == 0:               # a GPU is present (GPU box): nothing to check here
        L.icp_ctx_destroy(h)
        pytest.skip("CUDA device present")
    assert rc == capi.ICP_ERR_CUDA
    assert b"no CPU fallback" in L.icp_last_error() or b"CUDA" in L.icp_last_error()


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_distance_kernels_are_not_fma_contracted():
    """The RBC distance must stay an exactly ordered sequence of rounded mul/add (north_star): the search
    kernels may contain FMUL/FADD but no FFMA (packed or scalar)."""
    so = os.path.join(ROOT, "icp_b200", "libicp_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    cur = None
    counts = {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = {"FFMA": 0, "FMUL": 0, "FADD": 0}
            continue
        if cur:
            for op in ("FFMA", "FMUL", "FADD"):
                if re.search(r"\b%s2?\b" % op, line):
                    counts[cur][op] += 1
    # k_search additionally holds one IEEE division (the weight 100/(100+d)), whose Newton steps are FFMA by design
    hot = [k for k in counts if re.search(r"k_assign|k_nearest_rep|k_rbc_stage2", k)]
    assert len(hot) >= 3, "hot kernels not found in the SASS dump"
    for k in hot:
        assert counts[k]["FMUL"] > 0 and counts[k]["FADD"] > 0, (k, counts[k])
        if re.search(r"k_assign_triILb1ELb[01]ELb1", k):
            # the SETTLE instantiations of the pruned kernel A also hold the temporal pruning of stage 1: its bound
            # arithmetic uses IEEE sqrt with directed rounding (__fsqrt_ru / __fsqrt_rd), whose Newton steps are FFMA by design
            # (every one of them next to a MUFU: test_search_kernels_only_fuse_inside_division_and_sqrt_sequences).  The distance
            # code is the same inlined source as in the instantiations without it, which must hold none at all.
            assert counts[k]["FFMA"] <= 48, (k, counts[k])
            continue
        assert counts[k]["FFMA"] == 0, (k, counts[k])
    assert any("k_assign_triILb0" in k for k in hot), "build flavour of the pruned kernel A not found"
    assert any(re.search(r"k_assign_triILb1ELb0ELb0", k) for k in hot), "latency-mode search flavour of the pruned kernel A not found"


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_search_kernels_only_fuse_inside_division_and_sqrt_sequences():
    """The dominant kernels (stage-2 list scans k_search_sorted / k_search_grouped / k_search<L>, and the pruned stage-1 kernel
    k_assign_tri) legitimately hold a few FFMA: the Newton steps of the IEEE division (weight 100 / (100 + d)) and of the
    directed-rounding sqrt of the temporal-pruning bounds.  Those sequences all start from a MUFU (RCP / RSQ) instruction;
    the distance arithmetic has none.  So: every FFMA must lie within 30 instructions of a MUFU -- an FFMA anywhere else
    would be a contracted distance term."""
    so = os.path.join(ROOT, "icp_b200", "libicp_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    cur, ins = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            ins[cur] = []
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if cur and m:
            ins[cur].append(m.group(1))
    hot = [k for k in ins if re.search(r"k_search_sortedILb0|k_search_grouped|k_searchILi|k_assign_triILb1", k)]
    assert len(hot) >= 5, hot
    for k in hot:
        v = ins[k]
        mufu = [i for i, x in enumerate(v) if "MUFU" in x]
        ffma = [i for i, x in enumerate(v) if re.search(r"\bFFMA2?\b", x)]
        fmul = sum(1 for x in v if re.search(r"\bFMUL\b", x))
        assert fmul > 40, (k, fmul)
        assert len(ffma) <= 48, (k, len(ffma))
        for i in ffma:
            assert mufu and min(abs(i - j) for j in mufu) <= 30, f"{k}: FFMA at instruction {i} is not part of a division / sqrt sequence"
