"""Pins the CPU oracle against every golden vector the reference ships
(SURVEY Appendix C): 730 RTL-simulation pairs, 598 ssearch36 scores,
16 swalign scores (gap_open=-8), 1 CAPI end-to-end result."""
import random

import pytest

from oracle.smith_waterman import gotoh, localalignment


def _fasta(golden, name):
    return dict((n, s) for n, s in golden["fasta"][name])


def _query(golden, name):
    recs = golden["fasta"][name]
    assert len(recs) == 1
    return recs[0][1]


def test_counts(golden):
    assert sum(len(s["rows"]) for s in golden["rtl"]) == 730
    assert sum(len(s["rows"]) for s in golden["ssearch"]) == 598
    assert len(golden["swalign"]["rows"]) == 16
    assert golden["capi"]["result"] == 102 and golden["capi"]["biased"] == 2150


@pytest.mark.parametrize("width", [0, 12])
def test_rtl_outputs_c_oracle(golden, oracle_mod, width):
    o = oracle_mod.Oracle(score_width=width)
    n = 0
    for s in golden["rtl"]:
        q = _query(golden, s["query"])
        db = _fasta(golden, s["db"])
        for name, score, _t in s["rows"]:
            assert o.score(q, db[name]) == score, (s["file"], name)
            n += 1
    assert n == 730


def test_ssearch_scores_c_oracle(golden, oracle_mod):
    o = oracle_mod.Oracle()
    n = 0
    for s in golden["ssearch"]:
        q = _query(golden, s["query"])
        db = _fasta(golden, s["db"])
        for name, score in s["rows"]:
            assert o.score(q, db[name]) == score, (s["file"], name)
            n += 1
    assert n == 598


def test_swalign_scores_alt_params(golden, oracle_mod):
    sw = golden["swalign"]
    o = oracle_mod.Oracle(**sw["params"])
    q = _query(golden, sw["query"])
    db = _fasta(golden, sw["db"])
    for name, score in sw["rows"]:
        assert o.score(q, db[name]) == score, name


@pytest.mark.parametrize("v03", [0, 1])
def test_capi_end_to_end(golden, oracle_mod, v03):
    c = golden["capi"]
    o = oracle_mod.Oracle(score_width=12, first_col_v03=v03)
    assert o.score(c["query"], c["library"]) == c["result"]
    assert o.score(c["query"], c["library"]) + 2048 == c["biased"]


def test_python_localalignment_config1(golden):
    """BASELINE config 1: data1.fa x query1.fa through the completed
    data/smith-waterman.py, checked against data1.fa_query1.fa_out.txt by name."""
    s = [x for x in golden["rtl"] if x["file"] == "data1.fa_query1.fa_out.txt"][0]
    q = _query(golden, s["query"])
    db = _fasta(golden, s["db"])
    assert len(s["rows"]) == 20
    for name, score, _t in s["rows"]:
        assert localalignment(q, db[name]) == score, name


def test_python_localalignment_sample_of_data500(golden):
    s = [x for x in golden["ssearch"] if x["file"] == "score500.txt"][0]
    q = _query(golden, s["query"])
    db = _fasta(golden, s["db"])
    for name, score in s["rows"][:25]:
        assert localalignment(q, db[name]) == score, name


def _rand(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def _mutate(rng, s, psub=0.1, pindel=0.05):
    out = []
    for ch in s:
        r = rng.random()
        if r < pindel / 2:
            continue
        if r < pindel:
            out.append(rng.choice("ACGT"))
        out.append(rng.choice("ACGT") if rng.random() < psub else ch)
    return "".join(out)


def test_c_oracle_equals_python(oracle_mod):
    rng = random.Random(1234)
    for params in [(5, -4, -12, -4), (5, -4, -8, -4), (5, -4, -2, -1), (1, -3, -5, -2), (2, -5, 0, -2)]:
        o = oracle_mod.Oracle(*params)
        for _ in range(40):
            a = _rand(rng, rng.randint(1, 60))
            b = _mutate(rng, a) if rng.random() < 0.7 else _rand(rng, rng.randint(1, 60))
            if not b:
                b = "A"
            assert o.score(a, b) == localalignment(a, b, *params)


def test_score_is_symmetric(oracle_mod):
    """The combined-I recurrence treats 'up' and 'left' alike, so swapping query
    and subject cannot change the score (used by the engine to pick which
    sequence lives in registers)."""
    rng = random.Random(7)
    o = oracle_mod.Oracle()
    for _ in range(100):
        a = _rand(rng, rng.randint(1, 80))
        b = _mutate(rng, a)
        if not b:
            continue
        assert o.score(a, b) == o.score(b, a)


def test_pe_vs_gotoh(oracle_mod):
    """identical to Gotoh for the shipped penalties, higher for cheap gaps."""
    rng = random.Random(99)
    same = oracle_mod.Oracle(5, -4, -12, -4)
    diff = 0
    cheap = oracle_mod.Oracle(5, -4, -2, -1)
    for _ in range(300):
        a = _rand(rng, rng.randint(5, 60))
        b = _mutate(rng, a, 0.15, 0.15) or "A"
        assert same.score(a, b) == gotoh(a, b, 5, -4, -12, -4)
        pe, gt = cheap.score(a, b), gotoh(a, b, 5, -4, -2, -1)
        assert pe >= gt
        diff += pe != gt
    assert diff > 0


def test_width12_wrap_then_clamp(oracle_mod):
    """identical sequences of length L score 5*L while 5*L <= 2047;
    beyond that the diagonal run wraps to 0 and the running max keeps the last
    in-range value (2045 on the pure diagonal)."""
    rng = random.Random(1)
    w12 = oracle_mod.Oracle(score_width=12)
    wide = oracle_mod.Oracle()
    for L in (1, 100, 400, 409):
        s = _rand(rng, L)
        assert w12.score(s, s) == 5 * L == wide.score(s, s)
    for L in (410, 411, 500, 900):
        s = _rand(rng, L)
        assert wide.score(s, s) == 5 * L
        assert 2045 <= w12.score(s, s) <= 2047


def test_edge_cases(oracle_mod):
    o = oracle_mod.Oracle()
    assert o.score("A" * 50, "T" * 50) == 0
    assert o.score("A", "A") == 5
    assert o.score("A", "C") == 0
    assert o.score("ACGT", "acgt") == 20
    assert o.score("ACGTNACGT", "ACGTTACGT") == 45   # unknown letters pack as T (aligner_Header.c:38-39)


def test_first_col_variant_coincides_on_domain(oracle_mod):
    """the PE v0.3 first-column form equals the general form whenever
    match + gap_open <= 0."""
    rng = random.Random(5)
    a = oracle_mod.Oracle(5, -4, -12, -4, 0, 0)
    b = oracle_mod.Oracle(5, -4, -12, -4, 0, 1)
    for _ in range(200):
        x = _rand(rng, rng.randint(1, 40))
        y = _mutate(rng, x, 0.2, 0.2) or "G"
        assert a.score(x, y) == b.score(x, y)


def test_batch_packed_matches_single(oracle_mod):
    import numpy as np
    import ctypes as C
    rng = random.Random(3)
    o = oracle_mod.Oracle()
    qs = [_rand(rng, rng.randint(1, 70)) for _ in range(3)]
    ts = [_rand(rng, rng.randint(1, 90)) for _ in range(17)]

    def pack(seqs):
        bufs, lens, offs, off = [], [], [], 0
        for s in seqs:
            b = (C.c_uint8 * ((len(s) + 3) // 4))()
            o.lib.swo_pack_2bit(s.encode(), C.c_size_t(len(s)), b)
            bufs.append(bytes(b)); lens.append(len(s)); offs.append(off); off += len(b)
        return (np.frombuffer(b"".join(bufs), dtype=np.uint8), np.array(lens, np.uint32),
                np.array(offs, np.uint64))
    qp, ql, qo = pack(qs)
    tp, tl, to = pack(ts)
    out, used = o.score_batch_packed(qp, ql, qo, tp, tl, to)
    assert used >= 1
    for i, q in enumerate(qs):
        for j, t in enumerate(ts):
            assert out[i, j] == o.score(q, t)
