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
    with pytest.raises(TypeError):                # complex element types are not accelerated; no fallback
        T.rpca(np.zeros((4, 4), dtype=np.complex128))
    with pytest.raises(NotImplementedError):      # arbitrary mu callables cannot cross the C ABI (two built-ins can)
        T.rpca_ga(np.zeros((4, 4)), 2, mu=lambda s, w, U: s)


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_float32_and_callables_still_need_the_gpu():
    """Float32 inputs are promoted and svd / opnorm callables cross the C ABI as function pointers -- both still end
    in the CUDA library (no CPU fallback)."""
    with pytest.raises(T.TlsqError) as ei:
        T.rpca(np.ones((6, 4), dtype=np.float32))
    assert ei.value.code == T._cabi.TLSQ_ERR_NO_DEVICE
    with pytest.raises(T.TlsqError) as ei:
        T.rpca(np.ones((6, 4)), svd=lambda Z, k: np.linalg.svd(Z, full_matrices=False), opnorm=lambda Z: 1.0)
    assert ei.value.code == T._cabi.TLSQ_ERR_NO_DEVICE


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


def _plan(nranks, votes, env=-1):
    lib = T.load()
    f, w, ip = ctypes.c_int(-1), ctypes.c_int(-1), ctypes.c_int(-1)
    v = (ctypes.c_double * 4)(*[float(x) for x in votes])
    assert lib.tlsq_plan_pipeline(nranks, v, env, ctypes.byref(f), ctypes.byref(w), ctypes.byref(ip)) == 0
    return f.value, w.value, ip.value


def test_pipeline_choice_is_a_function_of_the_rank_summed_votes():
    """Sharded solves: the per-iteration pipelines issue different collectives, so the choice must only depend on the
    votes summed over the ranks (a shard with an odd row count cannot run the two-kernel pipeline: it must pull every
    rank onto the one-pass kernel, not pick it alone)."""
    # votes_sum = {can two-kernel, can one-pass, wants one-pass (memory), wants Y in place}
    assert _plan(8, [8, 8, 0, 0]) == (0, 1, 0)            # everything fits everywhere: two-kernel pipeline
    assert _plan(8, [7, 8, 0, 0]) == (1, 0, 0)            # one shard cannot run it -> one-pass kernel on ALL ranks
    assert _plan(8, [8, 8, 1, 0]) == (1, 0, 0)            # one rank short of memory -> one-pass kernel on all
    assert _plan(8, [8, 8, 1, 1]) == (1, 0, 1)            # ... and Y in place on all
    assert _plan(8, [7, 7, 0, 0]) == (0, 0, 0)            # neither available everywhere -> generic kernels on all
    assert _plan(8, [8, 3, 8, 8]) == (0, 1, 1)            # one-pass not available everywhere: stay on two-kernel
    assert _plan(1, [1, 1, 0, 0], env=1) == (1, 0, 0) and _plan(1, [1, 1, 1, 0], env=0) == (0, 1, 0)
    assert _plan(2, [2, 1, 0, 0], env=1) == (0, 1, 0)     # forcing cannot override eligibility


@pytest.mark.parametrize("K,P", [(49_999_745, 8), (49_999_745, 2), (24_576, 2), (12_000, 8), (461, 4), (7, 8), (100, 1)])
def test_hankel_row_shards_cover_the_rows_once(K, P):
    lib = T.load()
    nxt = 0
    sizes = []
    for r in range(P):
        r0, kl = ctypes.c_int64(-1), ctypes.c_int64(-1)
        assert lib.tlsq_plan_hankel_shard(K, P, r, ctypes.byref(r0), ctypes.byref(kl)) == 0
        assert r0.value == nxt and kl.value >= 0
        nxt += kl.value
        sizes.append(kl.value)
    assert nxt == K
    if K >= 64 * P:
        assert all(s % 32 == 0 for s in sizes[:-1])       # tile-aligned shards, remainder on the last rank
        assert max(sizes) - min(sizes) < 32 * P + 32


def test_one_pass_kernel_strip_table_covers_the_upper_triangle_once():
    """fused.cu: the 36 upper 32 x 32 tiles of the 256 x 256 Gram = 144 strips of 32 x 8, 72 per CTA of the cluster pair,
    9 per warp, every strip exactly once, at most two tile rows per warp (two A-fragment sets), and CTA 0 only touches
    columns < 192 (it receives columns 128..191 from its peer, nothing else)."""
    rows = np.zeros((2, 8, 9), dtype=np.uint8)
    cols = np.zeros((2, 8, 9), dtype=np.uint8)
    assert T.load().tlsq_plan_fused_strips(rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p)) == 0
    seen = set()
    for r in range(2):
        for w in range(8):
            assert len(set(rows[r, w].tolist())) <= 2
            assert sorted(rows[r, w].tolist()) == rows[r, w].tolist()          # first n1 strips share tile row A
            for s in range(9):
                a, c = int(rows[r, w, s]), int(cols[r, w, s])
                assert c // 4 >= a                                              # upper triangle (tile column >= tile row)
                assert (a, c) not in seen
                seen.add((a, c))
                if r == 0:
                    assert a <= 3 and c < 24
    assert len(seen) == 144 and seen == {(a, c) for a in range(8) for c in range(4 * a, 32)}


def test_hierarchical_jacobi_tournament_visits_every_pair_once_per_sweep():
    """eig.cu jacobi_cluster_block_kernel, modelled on the host: C CTAs x two half-blocks of spc columns; per outer round
    the spc x spc cross pairs of a CTA (inner round r: T[w] with B[(w + r) % spc]), in outer round 0 also the pairs
    inside the half-blocks (circle method), then the half-blocks move (top[0] fixed, top[c] -> top[c+1],
    top[C-1] -> bot[C-1], bot[c] -> bot[c-1], bot[0] -> top[1])."""
    for C, spc in [(16, 8), (8, 8), (4, 5), (2, 3), (1, 8), (1, 1)]:
        ncol = 2 * C * spc
        top = [list(range(c * 2 * spc, c * 2 * spc + spc)) for c in range(C)]
        bot = [list(range(c * 2 * spc + spc, (c + 1) * 2 * spc)) for c in range(C)]
        pairs = []
        spe = spc + (spc & 1)
        for orow in range(2 * C - 1):
            for c in range(C):
                if orow == 0 and spc > 1:
                    for half in (top[c], bot[c]):
                        for r in range(spe - 1):
                            for pi in range(spe // 2):
                                p, q = (spe - 1, r) if pi == 0 else ((r + pi) % (spe - 1), (r - pi + spe - 1) % (spe - 1))
                                if p < spc and q < spc:
                                    pairs.append(frozenset((half[p], half[q])))
                for r in range(spc):
                    for w in range(spc):
                        pairs.append(frozenset((top[c][w], bot[c][(w + r) % spc])))
            if C > 1:
                ntop, nbot = [None] * C, [None] * C
                for c in range(C):
                    if c == 0:
                        ntop[0] = top[0]
                    elif c == C - 1:
                        nbot[C - 1] = top[c]
                    else:
                        ntop[c + 1] = top[c]
                    if c == 0:
                        ntop[1] = bot[0]
                    else:
                        nbot[c - 1] = bot[c]
                top, bot = ntop, nbot
        assert len(pairs) == len(set(pairs)) == ncol * (ncol - 1) // 2, (C, spc)
