"""CPU-side tests: the C-ABI library loads and exports every symbol the header declares, and the
product's host-side operator algebra produces the same ordered connection tables as the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers as H
from oracle import estimators as OE
from oracle.models import lindblad_ising_1d, tfim_1d, tfim_2d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(nq):
    hdr = open(os.path.join(ROOT, "include", "nqcuda.h")).read()
    declared = sorted(set(re.findall(r"\b(nq_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 50
    lib = ctypes.CDLL(nq.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libnqcuda.so does not export %s" % name
    assert sorted(nq.EXPORTS) == declared, "ctypes prototypes out of sync with the header"
    assert lib.nq_version() == 100


def test_reference_side_bindings_cover_the_header():
    """Every entry point the header declares is bound in the Julia shim (a `ccall((:name, lib), ...)`) and mapped to the
    reference interface it replaces in INTEGRATION.md."""
    hdr = open(os.path.join(ROOT, "include", "nqcuda.h")).read()
    declared = sorted(set(re.findall(r"\b(nq_[a-z0-9_]+)\s*\(", hdr)))
    shim = open(os.path.join(ROOT, "julia", "NQCuda.jl")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    bound = set(re.findall(r"ccall\(\(:(nq_[a-z0-9_]+), lib\)", shim))
    assert [s for s in declared if s not in bound] == []
    assert [s for s in bound if s not in declared] == [], "the shim binds a symbol the header does not declare"
    assert [s for s in declared if s not in doc] == []


def test_no_cpu_fallback(nq):
    """Without a CUDA device the product refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nq.NQError):
        nq.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "neuralquantum.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def _digits(hilb_kind, sig):
    return ((sig + 1) / 2).astype(int) if hilb_kind == "spin" else sig.astype(int)


@pytest.mark.parametrize("N", [3, 6])
def test_ket_tables_match_oracle_order(nq, N):
    oh, oH = tfim_1d(N, 0.7, 1.3)
    ph, pH = H.p_tfim_1d(nq, N, 0.7, 1.3)
    tb = pH.tables()
    S = H.rand_states("spin", N, 20, 5)
    for b in range(S.shape[1]):
        ref = OE.connection_list_ket(oH, S[:, b])
        got = H.enumerate_tables(tb, _digits("spin", S[:, b]))
        assert [(m, f) for m, f, _ in got] == ref


def test_ket_tables_2d(nq):
    oh, oH = tfim_2d(3)
    ph, pH = H.p_tfim_2d(nq, 3)
    tb = pH.tables()
    S = H.rand_states("spin", 9, 10, 6)
    for b in range(S.shape[1]):
        assert [(m, f) for m, f, _ in H.enumerate_tables(tb, _digits("spin", S[:, b]))] == \
            OE.connection_list_ket(oH, S[:, b])


@pytest.mark.parametrize("fock", [True, False])
def test_liouvillian_tables_match_oracle_order(nq, fock):
    N = 4
    oh, oHm, oj, ol = lindblad_ising_1d(N, 0.4, 2.0, fock=fock)
    ph, pHm, pj, pl = H.p_lindblad_ising_1d(nq, N, 0.4, 2.0, fock=fock)
    tb = pl.tables()
    kind = "fock" if fock else "spin"
    R = H.rand_states(kind, N, 30, 7)
    Cc = H.rand_states(kind, N, 30, 8)
    for b in range(R.shape[1]):
        ref = OE.connection_list_super(ol, R[:, b], Cc[:, b])
        got = H.enumerate_tables(tb, _digits(kind, R[:, b]), _digits(kind, Cc[:, b]))
        assert got == ref


def test_operator_algebra_edge_cases(nq):
    h = nq.HomogeneousSpin(3)
    z = nq.LocalOperator(h)
    assert z.tables()["n_terms"] == 0
    a = nq.sigmax(h, 1) + nq.sigmax(h, 1)            # same-site merge keeps one term
    assert len(a.terms) == 1 and a.terms[0].rows[0][1][0] == 2
    b = nq.sigmax(h, 2) - nq.sigmax(h, 2)            # merged to zero mel, entry kept (skipped on device)
    assert b.terms[0].rows[0] == [[0j, 0], [0j, 1]]
    with pytest.raises(NotImplementedError):
        nq.HomogeneousFock(3, 3)
