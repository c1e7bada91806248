"""CPU: host-side helpers of the C ABI that need no GPU -- compact transports (2-bit packed reads, one bit per base),
the read partition for multi-GPU runs, sub-batch planning and the longest-first order."""
import numpy as np
import pytest

import spumoni_b200 as sp


def _ragged(seed, n=3000, hi=700):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, hi, size=n)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    return rng, lens, off


@pytest.mark.parametrize("up", [False, True])
def test_pack_reads_round_trip(up):
    rng, lens, off = _ragged(1)
    alphabet = np.frombuffer(b"ACGTacgtNn*\x01\xff", dtype=np.uint8)
    reads = rng.choice(alphabet, size=int(off[-1]), p=[.235] * 4 + [.01] * 4 + [.004] * 5)
    packed, pos, val = sp.pack_reads_2bit(reads, toupper=up)
    assert packed.size == int(sp.load_library().spb_packed_bytes(reads.size))
    i = np.arange(reads.size)
    un = np.frombuffer(b"ACGT", dtype=np.uint8)[(packed[i // 4] >> (2 * (i % 4))) & 3].copy()
    assert np.all(np.diff(pos.astype(np.int64)) > 0)
    un[pos.astype(np.int64)] = val
    want = reads.copy()
    if up:  # lower-case acgt are packed as their upper-case codes: the query upper-cases them anyway
        low = np.isin(want, np.frombuffer(b"acgt", dtype=np.uint8))
        want[low] -= 32
    assert np.array_equal(un, want)
    # nothing to pack / only exceptions
    p0, e0, v0 = sp.pack_reads_2bit(np.zeros(0, np.uint8))
    assert e0.size == 0
    _, e1, v1 = sp.pack_reads_2bit(np.frombuffer(b"NNNNN", dtype=np.uint8))
    assert e1.tolist() == [0, 1, 2, 3, 4] and bytes(v1) == b"NNNNN"


def test_bits_expand_and_widen():
    rng, lens, off = _ragged(2, n=500, hi=900)
    total = int(off[-1])
    m = rng.random(total) < 0.97
    L = np.zeros(total, dtype=np.uint32)
    for r in range(len(lens)):
        l = 0
        for j in range(int(off[r + 1]) - 1, int(off[r]) - 1, -1):
            l = l + 1 if m[j] else 0
            L[j] = l
    assert L.max() > 255
    bits = np.zeros(sp.bits_bytes(total) // 8, dtype=np.uint64)
    idx = np.nonzero(m)[0]
    np.bitwise_or.at(bits, idx // 64, np.uint64(1) << (idx % 64).astype(np.uint64))
    for w in (4, 2, 1):
        assert np.array_equal(sp.expand_bits(bits, off, w), np.minimum(L, (1 << (8 * w)) - 1))
    assert np.array_equal(sp.widen_lengths(np.minimum(L, 255).astype(np.uint8), off), L)
    assert np.array_equal(sp.widen_lengths(np.minimum(L, 65535).astype(np.uint16), off), L)
    assert sp.bits_bytes(0) == 0 and sp.bits_bytes(1) == 32 and sp.bits_bytes(256) == 32 and sp.bits_bytes(257) == 64


def test_shard_by_bases_tiles_and_balances():
    rng, lens, off = _ragged(3, n=5000, hi=12000)
    for n_shards in (1, 2, 3, 8):
        cuts = sp.shard_by_bases(off, n_shards)
        assert cuts[0] == 0 and cuts[-1] == len(lens) and np.all(np.diff(cuts.astype(np.int64)) >= 0)
        per = np.diff(off[cuts.astype(np.int64)].astype(np.int64))
        assert per.sum() == int(off[-1]) and per.max() - per.min() <= 2 * lens.max()
    # more shards than reads, empty input
    tiny = np.array([0, 5, 9], dtype=np.uint64)
    c = sp.shard_by_bases(tiny, 8)
    assert c[0] == 0 and c[-1] == 2 and np.all(np.diff(c.astype(np.int64)) >= 0)
    assert sp.shard_by_bases(np.zeros(1, np.uint64), 4).tolist() == [0] * 5


def test_plan_chunks_sizes_by_reads_not_bases():
    # 1 kb reads: a sub-batch closes at want_reads reads once it has min_bases bases
    off = np.arange(0, 200_001, dtype=np.uint64) * 1000
    c = sp.plan_chunks(off, want_reads=43000, min_bases=8 << 20, max_bases=256 << 20)
    assert np.all(np.diff(c.astype(np.int64))[:-1] == 43000) and c[-1] == 200_000
    # 10 kb reads: 43 000 reads would be 430 M bases -> the device-buffer bound closes the sub-batch, and it still holds
    # tens of thousands of reads (the old 32 Mi-base rule gave 3 355)
    off = np.arange(0, 100_001, dtype=np.uint64) * 10_000
    c = sp.plan_chunks(off, want_reads=43000, min_bases=8 << 20, max_bases=256 << 20)
    d = np.diff(c.astype(np.int64))
    assert d[0] == (256 << 20) // 10_000 and d.sum() == 100_000
    # 150 bp reads: enough reads long before min_bases -> min_bases decides
    off = np.arange(0, 1_000_001, dtype=np.uint64) * 150
    d = np.diff(sp.plan_chunks(off, 43000, 8 << 20, 256 << 20).astype(np.int64))
    assert d[0] == -(-(8 << 20) // 150)
    # a few long reads among short ones: the sub-batch grows to 32 768 x the longest read, so that the long chains overlap
    # with the rest instead of being the whole kernel
    lens = np.full(1_000_000, 150, dtype=np.int64)
    lens[::1000] = 5000
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    c = sp.plan_chunks(off, 43000, 8 << 20, 256 << 20)
    b = np.diff(off[c.astype(np.int64)].astype(np.int64))
    assert np.all(b[:-1] >= 5000 * 32768) and np.all(b <= (256 << 20))
    # one read longer than max_bases gets a sub-batch of its own; zero-length reads are legal
    off = np.array([0, 10, 10, 5000, 5003], dtype=np.uint64)
    assert sp.plan_chunks(off, 2, 1, 100).tolist() == [0, 2, 3, 4]
    with pytest.raises(sp.SpumoniError):
        sp.plan_chunks(np.array([0, 10, 5], dtype=np.uint64), 2, 1, 100)


def test_longest_first_order():
    rng, lens, off = _ragged(4, n=4000, hi=20000)
    order = sp.plan_order(off)
    assert sorted(order.tolist()) == list(range(len(lens)))
    ll = lens[order].astype(np.float64)
    # quarter-octave classes, longest class first: a later read is never more than one class (x 1.25) longer
    assert np.all(ll[1:][ll[:-1] >= 4] <= 1.26 * ll[:-1][ll[:-1] >= 4])
    assert ll[0] >= 0.8 * lens.max()
    # uniform lengths: no order (the kernels pull reads in input order)
    assert sp.plan_order(np.arange(0, 1001, dtype=np.uint64) * 1000) is None
    assert sp.plan_order(np.array([0, 150, 310, 452], dtype=np.uint64)) is None


# ---- property tests (hypothesis): the planning / transport helpers on arbitrary ragged inputs
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

# deterministic (the driver re-runs this suite: no run-to-run variation) and no health-check verdicts on generation cost
_PROP = dict(deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))

_lens = st.lists(st.integers(min_value=0, max_value=3000), min_size=1, max_size=300)


def _offsets(lens):
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    return off


@settings(max_examples=120, **_PROP)
@given(_lens, st.integers(1, 9))
def test_prop_shards_tile_the_reads(lens, n_shards):
    off = _offsets(lens)
    cuts = sp.shard_by_bases(off, n_shards).astype(np.int64)
    assert cuts[0] == 0 and cuts[-1] == len(lens) and np.all(np.diff(cuts) >= 0)
    total = int(off[-1])
    for s in range(1, n_shards):  # a cut never lands before the read that holds the s-th fraction of the bases
        assert int(off[cuts[s]]) >= total * s // n_shards or cuts[s] == len(lens)


@settings(max_examples=120, **_PROP)
@given(_lens, st.integers(1, 50), st.integers(1, 5000), st.integers(1, 20000))
def test_prop_chunks_tile_the_reads_and_respect_the_bounds(lens, want_reads, min_bases, max_extra):
    off = _offsets(lens)
    max_bases = min_bases + max_extra
    c = sp.plan_chunks(off, want_reads, min_bases, max_bases).astype(np.int64)
    assert c[0] == 0 and c[-1] == len(lens) and np.all(np.diff(c) >= 1)
    for a, b in zip(c[:-1], c[1:]):
        bases = int(off[b] - off[a])
        assert bases <= max_bases or b - a == 1          # only a single over-long read may exceed the device-buffer bound
        if b != len(lens):                                # a sub-batch that was closed early had a reason to be
            longest = int(np.max(np.diff(off[a:b + 1].astype(np.int64))))
            enough = (b - a) >= want_reads and bases >= max(min_bases, longest * 32768)
            nxt = int(off[b + 1] - off[a])
            assert enough or nxt > max_bases


@settings(max_examples=80, **_PROP)
@given(st.binary(min_size=0, max_size=2000), st.booleans())
def test_prop_pack_round_trip(data, up):
    reads = np.frombuffer(data, dtype=np.uint8)
    packed, pos, val = sp.pack_reads_2bit(reads, toupper=up)
    i = np.arange(reads.size)
    un = np.frombuffer(b"ACGT", dtype=np.uint8)[(packed[i // 4] >> (2 * (i % 4))) & 3].copy() if reads.size else np.zeros(0, np.uint8)
    un[pos.astype(np.int64)] = val
    want = reads.copy()
    if up:
        low = np.isin(want, np.frombuffer(b"acgt", dtype=np.uint8))
        want[low] -= 32
    assert np.array_equal(un, want)
    assert np.all(np.diff(pos.astype(np.int64)) > 0)


@settings(max_examples=80, **_PROP)
@given(_lens, st.integers(0, 2**32 - 1))
def test_prop_bits_and_clamped_lengths_are_lossless(lens, seed):
    off = _offsets(lens)
    total = int(off[-1])
    m = np.random.RandomState(seed).random_sample(total) < 0.93  # (one drawn seed: a draw per base exhausts hypothesis' entropy budget)
    L = np.zeros(total, dtype=np.uint32)
    for r in range(len(lens)):
        l = 0
        for j in range(int(off[r + 1]) - 1, int(off[r]) - 1, -1):
            l = l + 1 if m[j] else 0
            L[j] = l
    bits = np.zeros(max(sp.bits_bytes(total) // 8, 1), dtype=np.uint64)
    idx = np.nonzero(m)[0]
    np.bitwise_or.at(bits, idx // 64, np.uint64(1) << (idx % 64).astype(np.uint64))
    assert np.array_equal(sp.expand_bits(bits, off, 4), L)
    assert np.array_equal(sp.widen_lengths(np.minimum(L, 255).astype(np.uint8), off), L)


@settings(max_examples=80, **_PROP)
@given(_lens)
def test_prop_longest_first_order_is_a_permutation(lens):
    off = _offsets(lens)
    order = sp.plan_order(off)
    if order is None:
        return
    assert sorted(order.tolist()) == list(range(len(lens)))
    ll = np.asarray(lens)[order]
    cls = lambda v: 0 if v < 4 else 4 * (int(v).bit_length() - 1) + ((int(v) >> (int(v).bit_length() - 3)) & 3)  # noqa: E731
    assert all(cls(ll[i]) >= cls(ll[i + 1]) for i in range(len(ll) - 1))
