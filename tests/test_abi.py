"""CPU tier: the C-ABI library builds for sm_100a, loads without a GPU, and
exports every symbol include/upflow_b200.h declares with the arity the ctypes
binding uses.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "upflow_b200.h")


def _declarations():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|long long|const char\*)\s+(upf_\w+)\s*\(([^)]*)\)\s*;", src):
        name, args = m.group(1), m.group(2).strip()
        decls[name] = 0 if args in ("", "void") else len(args.split(","))
    return decls


@pytest.fixture(scope="module")
def lib_path():
    from upflow_pytorch_b200 import build
    return build.build()


def test_header_declares_the_hot_path():
    d = _declarations()
    for name in ("upf_corr_lrelu_fwd", "upf_corr_lrelu_bwd", "upf_warp_fwd", "upf_warp_bwd", "upf_featnorm_stats",
                 "upf_featnorm_apply", "upf_resize_bilinear", "upf_sgu_blend", "upf_conv2d_fwd",
                 "upf_conv_tc_pack_weights", "upf_last_error"):
        assert name in d


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _declarations():
        assert hasattr(lib, name), name


def test_ctypes_signatures_match_header(lib_path):
    from upflow_pytorch_b200 import _ext
    d = _declarations()
    assert set(d) == set(_ext.SIGNATURES)
    for name, n in d.items():
        assert len(_ext.SIGNATURES[name][1]) == n, name
    lib = _ext.load()
    assert lib.upf_abi_version() == _ext.ABI_VERSION == 2
    assert lib.upf_launch_count() == 0


def test_bad_arguments_return_codes_not_crashes(lib_path):
    """argument validation happens on the host before any launch (works without a GPU)"""
    from upflow_pytorch_b200 import _ext
    lib = _ext.load()
    rc = lib.upf_corr_lrelu_fwd(None, 32, None, 32, None, 81, 1, 8, 8, 32, 4, None, None, 0, 0.1, 0, None)
    assert rc == -1 and b"null" in lib.upf_last_error()
    rc = lib.upf_conv2d_fwd(ctypes.c_void_p(16), 32, ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 32,
                            None, 0, 1, 8, 8, 32, 32, 5, 1, 1, 0.1, 0, None)
    assert rc == -1 and b"kernel size" in lib.upf_last_error()
    # the training-side entry points validate the same way
    P = ctypes.c_void_p(64)
    rc = lib.upf_robust_loss_fwd(P, 2, P, 2, None, 0, P, P, 100, 2, 7, 0.4, None)
    assert rc == -1 and b"kind" in lib.upf_last_error()
    rc = lib.upf_robust_loss_bwd(P, 2, P, 2, None, 0, P, P, None, 2, None, 2, 100, 2, 0, 0.4, None)
    assert rc == -1 and b"null" in lib.upf_last_error()          # neither gradient requested
    rc = lib.upf_edge_smooth1_fwd(P, 3, 3, P, 2, 2, P, P, 1, 1, 8, None)
    assert rc == -1 and b"H, W >= 2" in lib.upf_last_error()
    rc = lib.upf_census_loss_fwd(P, 3, P, 3, None, 0, P, P, P, P, 1, 8, 8, 9, 0.4, None)
    assert rc == -1 and b"radius" in lib.upf_last_error()
    rc = lib.upf_boundary_warp_fwd(P, 3, 3, 8, 8, P, 1, P, P, 3, 1, 4, 4, None)
    assert rc == -1 and b"bad shape" in lib.upf_last_error()     # a flow needs two channels
    rc = lib.upf_repack_conv_weight_tc(P, P, 8, 8, 5, 0, None)
    assert rc == -1 and b"repack" in lib.upf_last_error()
    rc = lib.upf_conv2d_wgrad_tc_planar(None, 8, 0, P, 8, P, P, P, 1, 8, 8, 8, 8, 3, 1, None)
    assert rc == -1 and b"null" in lib.upf_last_error()
    assert lib.upf_loss_workspace_elems() >= 2 * 148
    # round 2: the planar (NCHW) correlation and the fp16 / bf16 storage variants
    L3 = ctypes.c_longlong * 3
    dense = lambda c: L3(8, 64, c * 64)
    rc = lib.upf_corr_lrelu_fwd_planar(P, dense(32), P, dense(32), P, dense(81), 1, 8, 8, 32, 5, 0, 0.1, 0, None)
    assert rc == -2 and b"max_disp > 4" in lib.upf_last_error()            # UPF_ENOTSUP: the pixel-major kernel serves d = 5, 6
    rc = lib.upf_corr_lrelu_fwd_planar(P, L3(6, 48, 1536), P, L3(6, 48, 1536), P, L3(8, 64, 5184), 1, 8, 6, 32, 4, 0, 0.1, 0, None)
    assert rc == -2 and b"multiples of 4" in lib.upf_last_error()           # TMA needs 16-byte pitches
    rc = lib.upf_corr_lrelu_fwd_planar(P, dense(32), P, dense(32), P, dense(81), 1, 8, 8, 32, 4, 3, 0.1, 0, None)
    assert rc == -1 and b"batch shift" in lib.upf_last_error()
    rc = lib.upf_corr_lrelu_fwd_lp(P, 32, P, 32, P, 81, 7, 1, 8, 8, 32, 4, None, None, 0, 0.1, None)
    assert rc == -1 and b"dtype" in lib.upf_last_error()
    rc = lib.upf_warp_fwd_lp(P, 32, P, 2, P, 32, 0, 1, 8, 8, 32, 0, 1.0, 0, None, None)
    assert rc == -1 and b"dtype" in lib.upf_last_error()
    assert lib.upf_wgrad_tc_planar_elems(2, 8, 8, 16, 3, 1) > 16 * lib.upf_wgrad_tc_planar_pitch(2, 8, 8, 3, 1)   # k padding + slack
    assert lib.upf_wgrad_tc_planar_pitch(2, 8, 8, 3, 1) % 32 == 0


def test_sass_has_blackwell_tensor_and_tma_instructions(lib_path):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper()
    assert re.search(r"UTC\w*MMA", sass), "tcgen05.mma missing"
    assert "UTMALDG" in sass, "TMA loads missing"
    assert "LDTM" in sass, "tcgen05.ld missing"
