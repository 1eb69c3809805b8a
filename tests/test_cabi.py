"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol declared in
include/tlsq_b200.h, fails loudly (no CPU fallback) and the host mirror validates arguments like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

import tlsq_b200 as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "tlsq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(tlsq_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(T._cabi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tlsq_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(T._cabi.SIGNATURES) == syms


def test_abi_version_and_device_count():
    lib = T.load()
    assert lib.tlsq_abi_version() == 1
    assert lib.tlsq_device_count() >= 0


def _no_gpu():
    return T.load().tlsq_device_count() == 0


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(T.TlsqError) as ei:
        T.rpca(np.random.default_rng(0).random((8, 4)))
    assert ei.value.code == T._cabi.TLSQ_ERR_NO_DEVICE
    with pytest.raises(T.TlsqError):
        T.rpca_ga(np.random.default_rng(0).random((8, 4)), 2)
    with pytest.raises(T.TlsqError):
        T.lowrankfilter(np.sin(np.arange(100.0)), 5)


def test_host_argument_validation():
    y = np.sin(0.1 * np.arange(100.0))
    with pytest.raises(AssertionError):           # @assert L <= N/2   (src/robustPCA.jl:79)
        T.lowrankfilter(y, 60)
    with pytest.raises(AssertionError):           # @assert lag <= L   (:80)
        T.lowrankfilter(y, 4, lag=5)
    with pytest.raises(AssertionError):
        T.hankel(y, 60)
    with pytest.raises(TypeError):                # Float64 only, no fallback
        T.rpca(np.zeros((4, 4), dtype=np.float32))
    with pytest.raises(NotImplementedError):      # custom svd cannot cross the C ABI
        T.rpca(np.zeros((4, 4)), svd=lambda Z, k: None)
    with pytest.raises(NotImplementedError):
        T.rpca_ga(np.zeros((4, 4)), 2, mu=lambda s, w, U: s)


def test_shard_rows_partition():
    for M, n, align in [(1_000_000, 8, 15625), (1_000_000, 4, 15625), (1003, 3, 1), (10, 4, 1), (2_000_000, 8, 15625)]:
        ranges = [T.synth.shard_rows(M, n, r, align) for r in range(n)]
        assert ranges[0][0] == 0 and ranges[-1][1] == M
        for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
            assert a1 == b0 and a0 <= a1
        assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= align


def test_synthetic_generators_are_deterministic():
    a = T.synth.lowrank_sparse_np(50, 8, 2, 0.1, seed=7)
    b = T.synth.lowrank_sparse_np(50, 8, 2, 0.1, seed=7)
    assert np.array_equal(a, b) and a.flags.f_contiguous
