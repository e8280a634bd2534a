"""CPU suite: the C-ABI library builds/loads here (nvcc cross-compiles, no GPU) and exports every symbol that
include/stp.h declares; the ctypes binding covers exactly that set.  No compute calls."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "stp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from segmentation_training_pipeline_b200 import build, lib
    build.build()
    l = lib.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(l, n), n
    assert set(names) == set(lib.SIGNATURES.keys()), set(names) ^ set(lib.SIGNATURES.keys())
    assert l.stp_version() == 100


def test_no_cuda_call_needed_for_host_helpers():
    from segmentation_training_pipeline_b200 import lib
    l = lib.load()
    assert l.stp_bn_nblk(16 * 128 * 128, 64) >= 1
    assert l.stp_bn_nblk(16 * 512 * 512, 3) <= 1024
    assert l.stp_loss_partial_floats() > 0
    assert l.stp_last_error() is not None


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "segmentation_training_pipeline_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                s = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M), os.path.join(dp, f)


def test_kernel_selection_options_are_known_and_unknown_names_fail():
    """every A/B knob include/stp.h documents is accepted by stp_set_option (no GPU involved), anything else is an error that names it"""
    from segmentation_training_pipeline_b200 import lib
    l = lib.load()
    hdr = open(os.path.join(ROOT, "include", "stp.h")).read()
    doc = hdr[hdr.index("debugging / A-B knobs"):hdr.index("int stp_set_option")]
    names = set(re.findall(r'"([a-z0-9_]+)"', doc))
    assert {"tc3", "gemm1x1", "tc2_1x1", "tc2_up2", "wgrad1x1", "nconv", "g1_bn", "bnb_fuse"} <= names
    for n in sorted(names):
        assert l.stp_set_option(n.encode(), 0) == 0, n
    assert l.stp_set_option(b"no_such_option", 1) != 0
    assert b"no_such_option" in l.stp_last_error()
