"""GPU parity tests: the CUDA path, called through the C ABI (libsw_b200.so via ctypes),
against the golden vectors of the reference and against the CPU oracle on seeded inputs.
Bit-exact is the bar (integer work)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fasta(golden, name):
    return golden["fasta"][name]


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


def _oracle_matrix(oracle_mod, pkg, queries, subjects, **params):
    o = oracle_mod.Oracle(**params)
    qp, ql, qo = pkg.pack_sequences(queries)
    tp, tl, to = pkg.pack_sequences(subjects)
    out, _ = o.score_batch_packed(qp, ql, qo, tp, tl, to)
    return out


# every strip-kernel variant the library instantiates (kept in sync by
# tests/test_host_abi.py::test_variant_list_matches_library) + automatic choice + 32-bit fallback
STRIP_VARIANTS = [
    "strip_s16x2_R30x1_G1", "strip_s16x2_R38x1_G1", "strip_s16x2_R75x1_G1", "strip_s16x2_R32x1_G1",
    "strip_s16x2_R50x1_G1", "strip_s16x2_R25x2_G1", "strip_s16x2_R19x2_G1", "strip_s16x2_R15x3_G1",
    "strip_s16x2_R30x2_G1", "strip_s16x2_R64x1_G1", "strip_s16x2_R32x2_G1", "strip_s16x2_R25x3_G1",
    "strip_s16x2_R38x2_G1", "strip_s16x2_R25x4_G1", "strip_s16x2_R25x1_G2", "strip_s16x2_R75x1_G2",
    "strip_s16x2_R25x3_G2", "strip_s16x2_R38x1_G4", "strip_s16x2_R19x2_G4", "strip_s16x2_R32x1_G4",
    "strip_s16x2_R16x1_G32", "strip_s16x2_R8x2_G32",
    # small-R warp-wide variants of the latency path (P ~ query length)
    "strip_s16x2_R1x1_G32", "strip_s16x2_R2x1_G32", "strip_s16x2_R4x1_G32", "strip_s16x2_R8x1_G32",
    "strip_s16x2_R8x1_G16", "strip_s16x2_R16x1_G8", "strip_s16x2_R2x2_G32",
    # experimental: 8 columns per trip of the step loop (A/B against the 4-column instances)
    "strip_s16x2_R25x2_G1_U8", "strip_s16x2_R25x3_G1_U8", "strip_s16x2_R38x2_G1_U8",
    # interior (predicate-free) trips over the columns every lane of a warp has
    "strip_s16x2_R25x2_G1_U4_F31",
]
# variants that also exist as DIRECT instances (column codes formed on the fly: the small-batch path)
DIRECT_VARIANTS = ["strip_s16x2_R16x1_G32", "strip_s16x2_R1x1_G32", "strip_s16x2_R2x1_G32", "strip_s16x2_R4x1_G32",
                   "strip_s16x2_R8x1_G32", "strip_s16x2_R8x1_G16", "strip_s16x2_R16x1_G8", "strip_s16x2_R2x2_G32"]
S16_VARIANTS = [v for v in STRIP_VARIANTS if "s16x2" in v]
VARIANTS = ["auto", "generic32"] + STRIP_VARIANTS


def _choose(e, variant):
    """auto = the library's own choice (small batches then take the latency path with its DIRECT
    instances); a named variant is run through the regular path (code stream in HBM) -- the DIRECT
    instances have their own tests in test_gpu_round2.py."""
    if variant == "auto":
        return
    e.set_small_batch_path(False)
    if variant == "generic32":
        e.set_kernel_choice(0, 0, True, -1)
    else:
        e.set_kernel_name(variant)


def test_config2_data500_query100_bit_exact(golden, pkg):
    """BASELINE config 2: data500.fa x query100.fa on one B200, bit-exact against both
    data500.fa_query100.fa_out.txt (RTL) and score500.txt (ssearch36)."""
    q = _fasta(golden, "query100.fa")[0][1]
    db = _fasta(golden, "data500.fa")
    names = [n for n, _ in db]
    with pkg.Engine() as e:
        sc = e.score([q], [s for _, s in db])
        assert e.kernel_launches >= 1 and "strip_s16x2" in e.last_kernel_name
    got = dict(zip(names, sc[0].tolist()))
    rtl = [s for s in golden["rtl"] if s["file"] == "data500.fa_query100.fa_out.txt"][0]
    assert len(rtl["rows"]) == 499
    for name, score, _t in rtl["rows"]:
        assert got[name] == score, name
    ss = [s for s in golden["ssearch"] if s["file"] == "score500.txt"][0]
    for name, score in ss["rows"]:
        assert got[name] == score, name


@pytest.mark.parametrize("variant", VARIANTS)
def test_all_golden_sets_every_variant(golden, pkg, variant):
    """All 730 RTL pairs + 598 ssearch36 scores, through every kernel variant."""
    n = 0
    with pkg.Engine() as e:
        _choose(e, variant)
        for s in golden["rtl"] + golden["ssearch"]:
            q = _fasta(golden, s["query"])[0][1]
            db = _fasta(golden, s["db"])
            sc = e.score([q], [x for _, x in db])
            got = dict(zip([nm for nm, _ in db], sc[0].tolist()))
            for row in s["rows"]:
                assert got[row[0]] == row[1], (s["file"], row[0], e.last_kernel_name)
                n += 1
    assert n == 730 + 598


def test_swalign_alt_params_and_capi(golden, pkg):
    sw = golden["swalign"]
    q = _fasta(golden, sw["query"])[0][1]
    db = dict(_fasta(golden, sw["db"]))
    with pkg.Engine(**sw["params"]) as e:
        sc = e.score([q], [db[n] for n, _ in sw["rows"]])
    assert sc[0].tolist() == [s for _, s in sw["rows"]]
    c = golden["capi"]
    with pkg.Engine(score_width=12) as e:
        assert int(e.score([c["query"]], [c["library"]])[0, 0]) == c["result"]


@pytest.mark.parametrize("variant", VARIANTS)
def test_random_mixed_lengths_vs_oracle(oracle_mod, pkg, variant):
    rng = random.Random(1000 + VARIANTS.index(variant))
    queries = [_rand(rng, n) for n in (1, 37, 150, 151, 203)]
    subjects = []
    for _ in range(300):
        base = rng.choice(queries)
        if rng.random() < 0.5:
            s = _mutate(rng, base, 0.08, 0.06)
            a = rng.randint(0, max(0, len(s) - 1))
            s = _rand(rng, rng.randint(0, 40)) + s[a:] + _rand(rng, rng.randint(0, 40))
        else:
            s = _rand(rng, rng.randint(1, 260))
        subjects.append(s or "A")
    subjects += ["", "A", "C", "ACGT" * 60, "T" * 150]
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    with pkg.Engine() as e:
        _choose(e, variant)
        got = e.score(queries, subjects)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("params", [(5, -4, -12, -4), (5, -4, -8, -4), (5, -4, -2, -1), (1, -3, -5, -2),
                                    (2, -5, 0, -2), (5, -9, -3, -4), (3, 0, -1, 0)])
def test_parameter_sets_vs_oracle(oracle_mod, pkg, params):
    """Run-time loadable penalties (ScoreBank_v2.v:34,161), including the cheap-gap sets
    where the PE recurrence differs from Gotoh."""
    rng = random.Random(hash(params) & 0xFFFF)
    queries = [_rand(rng, 90), _rand(rng, 150)]
    subjects = [(_mutate(rng, rng.choice(queries), 0.15, 0.15) or "A") for _ in range(200)]
    keys = dict(zip(("match", "mismatch", "gap_open", "gap_extend"), params))
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects, **keys)
    for choice in ["auto", "strip_s16x2_R38x1_G4", "strip_s16x2_R25x2_G1", "strip_s16x2_R25x3_G1", "generic32"]:
        with pkg.Engine(*params) as e:
            _choose(e, choice)
            got = e.score(queries, subjects)
        np.testing.assert_array_equal(got, want)


def test_score_width_12_wrap_then_clamp(oracle_mod, pkg):
    """identical sequences of length 400..900; 12-bit mode must
    reproduce the RTL's wrap-to-zero, wide mode the true 5*L."""
    rng = random.Random(1)
    seqs = [_rand(rng, L) for L in (400, 409, 410, 411, 500, 900)]
    noise = [_mutate(rng, s, 0.02, 0.01) for s in seqs]
    for width in (0, 12):
        o = oracle_mod.Oracle(score_width=width)
        for choice in ["auto", "strip_s16x2_R50x1_G1", "strip_s16x2_R25x2_G1", "strip_s16x2_R8x2_G32", "generic32"]:
            with pkg.Engine(score_width=width) as e:
                _choose(e, choice)
                got = e.score(seqs, seqs + noise)
            for i, q in enumerate(seqs):
                for j, t in enumerate(seqs + noise):
                    assert got[i, j] == o.score(q, t), (width, choice, len(q), len(t))
    with pkg.Engine() as e:
        got = e.score(seqs, seqs)
    assert [int(got[i, i]) for i in range(len(seqs))] == [5 * len(s) for s in seqs]


def test_long_query_multi_pass_and_chunks(oracle_mod, pkg):
    """Queries longer than one pass (R*G rows) and longer than one shared-memory profile
    chunk; subjects long enough to exercise the pass-boundary scratch."""
    rng = random.Random(77)
    q1 = _rand(rng, 2500)
    q2 = _rand(rng, 1030)
    subjects = []
    for _ in range(70):
        a = rng.randint(0, 1800)
        subjects.append(_mutate(rng, q1[a:a + rng.randint(50, 700)], 0.05, 0.03) or "A")
    subjects += [_rand(rng, rng.randint(1, 900)) for _ in range(30)]
    want = _oracle_matrix(oracle_mod, pkg, [q1, q2], subjects)
    for choice in ["auto", "generic32"] + S16_VARIANTS:
        with pkg.Engine() as e:
            _choose(e, choice)
            got = e.score([q1, q2], subjects)
        np.testing.assert_array_equal(got, want, err_msg=str(choice))


def test_scores_beyond_int16_are_recomputed_in_32bit(oracle_mod, pkg):
    """Pairs whose score leaves the 16-bit range are flagged by the packed kernel and recomputed
    by the 32-bit kernel; everything else in the batch stays on the fast path."""
    rng = random.Random(5)
    s = _rand(rng, 6700)                      # identical pair scores 33500 > 32767
    t = _mutate(rng, s, 0.004, 0.002)         # around the threshold
    u = _mutate(rng, s, 0.10, 0.05)           # well inside 16 bit
    others = [_rand(rng, rng.randint(100, 3000)) for _ in range(12)]
    subjects = [s, t, u, "ACGT", s[:6560], s[100:6652] + "A"] + others
    o = oracle_mod.Oracle()
    want = [o.score(s, x) for x in subjects]
    assert want[0] == 33500 and 32700 < want[4] == 32800 and max(want[6:]) < 2000
    for choice in ["auto", "strip_s16x2_R38x2_G1", "strip_s16x2_R16x1_G32", "generic32"]:
        sub = subjects[:5] if choice == "generic32" else subjects      # the scalar kernel is slow
        with pkg.Engine() as e:
            _choose(e, choice)
            got = e.score([s], sub)
            assert (e.last_kernel_name == "generic32") == (choice == "generic32")
        assert got[0].tolist() == want[:len(sub)], choice


def test_edge_cases(pkg):
    with pkg.Engine() as e:
        got = e.score(["A" * 50, "A", "ACGT"], ["T" * 50, "A", "C", "", "acgt", "ACGTNACGT"])
    assert got[0].tolist() == [0, 5, 0, 0, 5, 5]
    assert got[1].tolist() == [0, 5, 0, 0, 5, 5]
    assert got[2].tolist() == [5, 5, 5, 0, 20, 20]
    with pkg.Engine() as e:          # empty database, empty query set
        e.set_queries(["ACGT"])
        e.score_batch([])
        assert e.fetch().shape == (1, 0)
        e.set_queries([])
        e.score_batch(["ACGT"])
        assert e.fetch().shape == (0, 1)


def test_large_synthetic_sample_vs_oracle(oracle_mod, pkg):
    """Config 3 shape at reduced count: 150-nt reads, 8 queries, 40k subjects; the whole
    matrix is compared (oracle needs a few seconds)."""
    packed, ln, off = pkg.random_packed_db(40000, 150, seed=20160912)
    qp, ql, qo = pkg.random_packed_db(8, 150, seed=7)
    # plant homologs so that gaps are exercised
    rng = np.random.default_rng(3)
    nb = 38
    db2 = packed[:40000 * nb].reshape(40000, nb).copy()
    for k in range(0, 40000, 100):
        db2[k] = qp[(k // 100 % 8) * nb:(k // 100 % 8 + 1) * nb]
        db2[k, rng.integers(0, nb)] ^= np.uint8(rng.integers(1, 255))
    db2[:, -1] &= np.uint8(0x0F)
    flat = np.concatenate([db2.reshape(-1), np.zeros(16, np.uint8)])
    o = oracle_mod.Oracle()
    want, _ = o.score_batch_packed(qp, ql, qo, flat, ln, off)
    with pkg.Engine() as e:
        got = e.score((qp, ql, qo), (flat, ln, off))
        assert e.last_cells == 8 * 150 * 40000 * 150
    np.testing.assert_array_equal(got, want)
    assert want.max() > 600          # the planted homologs are really there


def test_resident_db_best_hit_and_ids(oracle_mod, pkg):
    rng = random.Random(11)
    queries = [_rand(rng, 120) for _ in range(3)]
    subjects = [_rand(rng, rng.randint(30, 200)) for _ in range(500)]
    subjects[137] = queries[1]
    subjects[400] = queries[1]
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    with pkg.Engine() as e:
        e.load_db(subjects, ids=np.arange(500, dtype=np.uint64) + 1000)
        e.set_queries(queries)
        e.score_db()
        e.wait()
        assert e.last_kernel_ms > 0
        np.testing.assert_array_equal(e.fetch_db(), want)
        bs, bi = e.fetch_best()
        assert bs.tolist() == want.max(axis=1).tolist()
        assert bi.tolist() == want.argmax(axis=1).tolist()
        assert int(bi[1]) == 137 and int(bs[1]) == 600
        assert e.fetch_ids()[5] == 1005
        # new query set against the same resident database
        e.set_queries(queries[:1])
        e.score_db()
        np.testing.assert_array_equal(e.fetch_db(), want[:1])


def test_state_errors_and_timeout(pkg):
    with pkg.Engine() as e:
        with pytest.raises(pkg.SwError) as ei:
            e.fetch()
        assert ei.value.code == pkg.SW_ESTATE
        e.set_queries(["ACGT" * 30])
        e.score_batch(["ACGT" * 30] * 10)
        e.score_batch(["ACGT" * 30] * 3)               # second buffer
        assert e.batches_in_flight == 2
        with pytest.raises(pkg.SwError) as ei:
            e.score_batch(["ACGT"])
        assert ei.value.code == pkg.SW_EAGAIN          # both buffers busy: the bank's `full`
        out = e.fetch(timeout_ms=10000)
        assert out[0].tolist() == [600] * 10
        assert e.fetch(timeout_ms=10000)[0].tolist() == [600] * 3
        assert e.batches_in_flight == 0
    with pytest.raises(pkg.SwError) as ei:
        pkg.Engine(gap_extend=3)
    assert ei.value.code == pkg.SW_EINVAL
    with pytest.raises(pkg.SwError) as ei:
        pkg.Engine(gpu_ids=[99])
    assert ei.value.code == pkg.SW_ENODEV


def test_multi_gpu_handle_if_available(oracle_mod, pkg):
    n = pkg.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    rng = random.Random(4)
    queries = [_rand(rng, 150) for _ in range(4)]
    subjects = [_rand(rng, rng.randint(1, 300)) for _ in range(3000)]
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    with pkg.Engine(gpu_ids=list(range(n))) as e:
        got = e.score(queries, subjects)
    np.testing.assert_array_equal(got, want)


def test_config4_long_query_sample_vs_oracle(oracle_mod, pkg):
    """BASELINE config 4 shape at reduced count: a 10 kb query against 1 kb subjects (scores up to
    5000 => packed 16-bit exact), automatic variant choice and the warp-wide systolic variants."""
    rng = random.Random(4)
    q = _rand(rng, 10000)
    subjects = []
    for k in range(240):
        if k % 3 == 0:
            a = rng.randint(0, 9000)
            subjects.append((_mutate(rng, q[a:a + 1000], 0.04, 0.02) + _rand(rng, 1000))[:1000])
        else:
            subjects.append(_rand(rng, 1000))
    want = _oracle_matrix(oracle_mod, pkg, [q], subjects)
    assert want.max() > 3000
    for choice in ["auto", "strip_s16x2_R16x1_G32", "strip_s16x2_R8x2_G32", "strip_s16x2_R25x3_G1"]:
        with pkg.Engine() as e:
            _choose(e, choice)
            got = e.score([q], subjects)
            assert "s16x2" in e.last_kernel_name
        np.testing.assert_array_equal(got, want, err_msg=choice)


def test_config5_mixed_length_sweep_vs_oracle(oracle_mod, pkg):
    """BASELINE config 5: lengths log-uniform 32..4096, length-bucketed batching, plus the hand
    cases (all-mismatch, single base, gap-only-profitable)."""
    rng = random.Random(55)
    nrng = np.random.default_rng(55)
    lens = np.exp(nrng.uniform(np.log(32), np.log(4096), size=700)).astype(int)
    queries = [_rand(rng, 32), _rand(rng, 150), _rand(rng, 1000), _rand(rng, 4096)]
    subjects = []
    for L in lens:
        if rng.random() < 0.4:
            src = rng.choice(queries)
            a = rng.randint(0, max(0, len(src) - 1))
            s = _mutate(rng, src[a:a + int(L)], 0.06, 0.04)
            s = (s + _rand(rng, int(L)))[:int(L)]
        else:
            s = _rand(rng, int(L))
        subjects.append(s)
    subjects += ["A" * 500, "T" * 500, "G", "ACGT" * 200 + "TTTT" + "ACGT" * 200]
    queries.append("ACGT" * 200 + "ACGT" * 200)          # the subject above needs a 4-base gap
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    with pkg.Engine() as e:
        got = e.score(queries, subjects)
    np.testing.assert_array_equal(got, want)
    assert 7960 <= want[4, -1] < 8000                       # 1600 matches minus one gap of ~4 residues


def test_streaming_double_buffered_batches(oracle_mod, pkg):
    """Streaming database mode: batches are submitted ahead of the fetch of the
    previous one; results come back in submission order and equal the one-shot result."""
    rng = random.Random(31)
    queries = [_rand(rng, 150) for _ in range(3)]
    batches = [[_rand(rng, rng.randint(1, 220)) for _ in range(n)] for n in (700, 1, 350, 0, 512)]
    want = [_oracle_matrix(oracle_mod, pkg, queries, b) if b else np.zeros((3, 0), np.int32) for b in batches]
    with pkg.Engine() as e:
        e.set_queries(queries)
        got = []
        e.score_batch(batches[0], ids=np.arange(700, dtype=np.uint64) + 7)
        for k in range(1, len(batches)):
            e.score_batch(batches[k])                  # batch k is enqueued ...
            assert e.batches_in_flight == 2
            got.append(e.fetch())                      # ... before batch k-1 is fetched
        got.append(e.fetch())
        assert e.batches_in_flight == 0
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)


def test_both_strands(oracle_mod, pkg):
    rng = random.Random(32)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    queries = [_rand(rng, n) for n in (31, 150, 77)]
    rc = ["".join(comp[c] for c in reversed(q)) for q in queries]
    subjects = [_rand(rng, rng.randint(20, 200)) for _ in range(200)]
    subjects += [rc[1][10:120], queries[2][5:70], rc[0]]
    want = _oracle_matrix(oracle_mod, pkg, queries + rc, subjects)
    with pkg.Engine() as e:
        e.set_strands(True)
        got = e.score(queries, subjects)
        assert got.shape == (6, len(subjects))
    np.testing.assert_array_equal(got, want)
    assert got[4, 200] == 5 * 110 and got[3, 202] == 5 * 31


def test_cli_regenerates_golden_files(golden, tmp_path):
    """File-level drop-in: the CLI, the counterpart of `main_test -q -l -t`, reads the
    FASTA pair and writes both text layouts; scores equal data500.fa_query100.fa_out.txt / score500.txt."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "bin", "sw_b200_cli")
    assert os.path.exists(cli), "bin/sw_b200_cli missing: run __graft_entry__.build()"
    qf, lf = tmp_path / "query100.fa", tmp_path / "data500.fa"
    qf.write_text(">query\n" + golden["fasta"]["query100.fa"][0][1] + "\n")
    lf.write_text("".join(f">{n}\n{s}\n" for n, s in golden["fasta"]["data500.fa"]))
    out_txt, score_txt = tmp_path / "out.txt", tmp_path / "score.txt"
    r = subprocess.run([cli, "-q", str(qf), "-l", str(lf), "-t", "30", "-o", str(out_txt), "-R", str(score_txt)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    rtl = dict((n, s) for n, s, _t in [x for x in golden["rtl"] if x["file"] == "data500.fa_query100.fa_out.txt"][0]["rows"])
    got = {}
    for line in out_txt.read_text().splitlines():
        m = re.match(r"@\s*(\d+)ns:\s+>(\S+) score:\s+(-?\d+)", line)
        assert m, line
        got[m.group(2)] = int(m.group(3))
    assert got == rtl
    ss = dict([x for x in golden["ssearch"] if x["file"] == "score500.txt"][0]["rows"])
    rows = [l.split() for l in score_txt.read_text().splitlines() if not l.startswith(("#", ">"))]
    assert {r_[0]: int(r_[5]) for r_ in rows} == ss
    # the one-pair form prints like main_test.c:528
    c = golden["capi"]
    (tmp_path / "query").write_text(c["query"] + "\n")
    (tmp_path / "library").write_text(c["library"] + "\n")
    r = subprocess.run([cli, "-q", str(tmp_path / "query"), "-l", str(tmp_path / "library"), "-w", "12"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "result: 102, biased: 2150(0x0866)" in r.stdout, r.stdout


def test_very_long_subjects_short_queries(oracle_mod, pkg):
    """Long column loops (a 60 kb 'genome' as subject), few pairs: the warp-wide variants."""
    rng = random.Random(60)
    queries = [_rand(rng, 150), _rand(rng, 90)]
    genome = _rand(rng, 60000)
    genome = genome[:30000] + _mutate(rng, queries[0], 0.03, 0.02) + genome[30000:]
    subjects = [genome, _rand(rng, 20000), _rand(rng, 20000), _rand(rng, 777)]
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    assert want[0, 0] > 500
    for choice in ["auto", "strip_s16x2_R16x1_G32", "strip_s16x2_R25x2_G1", "strip_s16x2_R38x1_G4"]:
        with pkg.Engine() as e:
            _choose(e, choice)
            got = e.score(queries, subjects)
        np.testing.assert_array_equal(got, want, err_msg=choice)


def test_fetch_timeout_then_success(pkg):
    """The host's -t option (main_test.c:422-477): a fetch that times out leaves the batch in
    flight; a later fetch returns it."""
    db = pkg.random_packed_db(1500000, 150, seed=9)
    q = pkg.random_packed_db(60, 150, seed=10)
    with pkg.Engine() as e:
        e.set_queries(q)
        e.score_batch(db)
        with pytest.raises(pkg.SwError) as ei:
            e.fetch(timeout_ms=1)
        assert ei.value.code == pkg.SW_ETIMEOUT and e.batches_in_flight == 1
        out = e.fetch(timeout_ms=-1)
        assert out.shape == (60, 1500000) and e.batches_in_flight == 0
        assert 20 <= int(out.max()) <= 750 and int(out.min()) >= 0


@pytest.mark.parametrize("fixed", [True, False])
def test_fixed_and_runtime_penalty_instances_agree(golden, oracle_mod, pkg, fixed):
    """The main variants exist twice: gap penalties as immediates (default set) and as run-time
    operands.  Both must give the golden / oracle scores for the default penalties."""
    rng = random.Random(71)
    queries = [_rand(rng, 150), _rand(rng, 64)]
    subjects = [(_mutate(rng, rng.choice(queries), 0.1, 0.08) or "A") for _ in range(400)]
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    q = _fasta(golden, "query100.fa")[0][1]
    db = _fasta(golden, "data500.fa")
    ss = dict([x for x in golden["ssearch"] if x["file"] == "score500.txt"][0]["rows"])
    # second compiled-in set (gap_open -8): the swalign golden vectors
    sw = golden["swalign"]
    sq = _fasta(golden, sw["query"])[0][1]
    sdb = dict(_fasta(golden, sw["db"]))
    pkg.set_fixed_penalty_kernels(fixed)
    try:
        for name in ["strip_s16x2_R25x2_G1", "strip_s16x2_R38x2_G1", "strip_s16x2_R50x1_G1"]:
            with pkg.Engine(**sw["params"]) as e:
                e.set_kernel_name(name)
                got = e.score([sq], [sdb[n] for n, _ in sw["rows"]])
            assert got[0].tolist() == [x for _, x in sw["rows"]], name
        for name in ["strip_s16x2_R25x2_G1", "strip_s16x2_R38x2_G1", "strip_s16x2_R50x1_G1", "strip_s16x2_R38x1_G4",
                     "strip_s16x2_R16x1_G32"]:
            with pkg.Engine() as e:
                e.set_kernel_name(name)
                np.testing.assert_array_equal(e.score(queries, subjects), want, err_msg=name)
                sc = e.score([q], [s for _, s in db])
                assert dict(zip([n for n, _ in db], sc[0].tolist())) == ss, name
    finally:
        pkg.set_fixed_penalty_kernels(True)


def test_every_shipped_fasta_file_vs_oracle(golden, oracle_mod, pkg):
    """All 18 FASTA files of the reference's data/ directory (including the ones that ship
    without an expected-output file: data.fa, data2.fa, data50/80/200/300/400.fa, score_test.fa,
    datal.fa and the 23 unpinned records of data40.fa) against both shipped queries.
    Expectations here are ORACLE-GENERATED (the pinned subsets are covered by the golden tests)."""
    queries = [golden["fasta"]["query1.fa"][0][1], golden["fasta"]["query100.fa"][0][1]]
    n_files = 0
    with pkg.Engine() as e:
        for fn, recs in sorted(golden["fasta"].items()):
            subjects = [s for _n, s in recs]
            if not subjects:
                continue
            want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
            got = e.score(queries, subjects)
            np.testing.assert_array_equal(got, want, err_msg=fn)
            n_files += 1
    assert n_files == 18


def test_randomised_stress_vs_oracle(oracle_mod, pkg):
    """Seeded property test: random penalties inside the supported domain, random score width,
    random kernel variant, ragged lengths (including 0 and 1), a few planted homologs."""
    rng = random.Random(20160912)
    variants = ["auto", "generic32"] + STRIP_VARIANTS
    for it in range(40):
        match = rng.randint(1, 9)
        mismatch = -rng.randint(0, 9)
        gap_extend = -rng.randint(0, 6)
        gap_open = -rng.randint(0, 14)
        if gap_open + gap_extend > 0:
            gap_open = -gap_extend
        width = rng.choice([0, 0, 0, 9, 10, 12, 15])
        if width and (match >= (1 << (width - 1)) or gap_open + 2 * gap_extend < -(1 << (width - 1))):
            width = 0
        nq = rng.randint(1, 4)
        queries = [_rand(rng, rng.choice([1, 2, 31, 50, 75, 76, 77, 150, 300, rng.randint(1, 400)])) for _ in range(nq)]
        subjects = []
        for _ in range(rng.randint(1, 120)):
            if rng.random() < 0.3:
                src = rng.choice(queries)
                s = _mutate(rng, src, 0.1, 0.1)
                s = _rand(rng, rng.randint(0, 20)) + s + _rand(rng, rng.randint(0, 20))
            else:
                s = _rand(rng, rng.choice([0, 1, 3, rng.randint(1, 350)]))
            subjects.append(s)
        params = dict(match=match, mismatch=mismatch, gap_open=gap_open, gap_extend=gap_extend, score_width=width)
        want = _oracle_matrix(oracle_mod, pkg, queries, subjects, **params)
        choice = rng.choice(variants)
        with pkg.Engine(**params) as e:
            _choose(e, choice)
            got = e.score(queries, subjects)
        np.testing.assert_array_equal(got, want, err_msg=f"iteration {it}: {params} {choice} {e.last_kernel_name}")
