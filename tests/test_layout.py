"""Repository rules that are cheap to check mechanically."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "relativisticraytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "oracle/" not in txt.replace("the oracle", ""), f"{f} references oracle/"
                assert "/root/reference" not in txt, f


def test_oracle_says_it_is_test_infrastructure():
    for f in ("oracle_abi.h", "rrt_oracle.c", "ref_harness.cpp", "tex_emul.h", "__init__.py", "ref_cuda_harness.cu",
              "ref_cuda_planes.cu", "ref_cuda_planes_prelude.h", "ref_stb.c"):
        assert "TEST INFRASTRUCTURE ONLY" in open(os.path.join(ROOT, "oracle", f)).read(), f


def test_runtime_files_do_not_read_the_reference_tree():
    """/root/reference does not exist on the GPU box: bench.py, smoke() and the gpu tests must not need it"""
    for f in ("bench.py", "__graft_entry__.py"):
        txt = open(os.path.join(ROOT, f)).read()
        for line in txt.splitlines():
            if "/root/reference" in line:
                assert "RRT_REFERENCE_TREE" in line or "isdir" in line or line.strip().startswith(("#", '"')), line
