"""GPU parity tests of the round-2 paths, all through the C ABI: the small-batch latency path (DIRECT
instances), int16 output + overflow side list, fused top-k, query groups, the list-driven 32-bit
fix-up, run-time specialised penalties, virtual multi-shard handles, and the bounds-check build."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from tests.test_gpu_parity import DIRECT_VARIANTS, _mutate, _oracle_matrix, _rand

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _topk_want(mat, k):
    """Oracle top-k: score descending, ties by ascending index; short rows padded with (-1, 2^64-1)."""
    nq, ns = mat.shape
    sc = np.full((nq, k), -1, dtype=np.int32)
    ix = np.full((nq, k), np.iinfo(np.uint64).max, dtype=np.uint64)
    for q in range(nq):
        order = np.lexsort((np.arange(ns), -mat[q].astype(np.int64)))[:k]
        sc[q, :len(order)] = mat[q, order]
        ix[q, :len(order)] = order
    return sc, ix


@pytest.mark.parametrize("variant", ["auto"] + DIRECT_VARIANTS)
def test_small_path_direct_variants_vs_oracle(oracle_mod, pkg, variant):
    """Latency path: one staging copy, one DIRECT kernel, mapped output.  Ragged lengths, empty
    subjects, odd counts, queries shorter / longer than one pass of the variant."""
    rng = random.Random(500 + len(variant))
    rows_per_pass = {"auto": 512, "strip_s16x2_R16x1_G32": 512, "strip_s16x2_R1x1_G32": 32, "strip_s16x2_R2x1_G32": 64,
                     "strip_s16x2_R4x1_G32": 128, "strip_s16x2_R8x1_G32": 256, "strip_s16x2_R8x1_G16": 128,
                     "strip_s16x2_R16x1_G8": 128, "strip_s16x2_R2x2_G32": 128}[variant]
    for qlens, nsub in (((1, 31, 32, 17), 61), ((128,), 499), ((150, 64), 300), ((600, 257), 40), ((1300, 513), 12), ((32,), 1)):
        if variant not in ("auto", "strip_s16x2_R16x1_G32") and max(qlens) > rows_per_pass:
            continue            # small-P DIRECT instances are single-pass by construction (the host checks)
        queries = [_rand(rng, n) for n in qlens]
        subjects = []
        for _ in range(nsub):
            if rng.random() < 0.4:
                s = _mutate(rng, rng.choice(queries), 0.08, 0.06)
                s = _rand(rng, rng.randint(0, 30)) + s + _rand(rng, rng.randint(0, 30))
            else:
                s = _rand(rng, rng.choice([0, 1, 2, 5, rng.randint(1, 300)]))
            subjects.append(s[:1024])           # the latency path takes subjects of up to 1024 bases
        if nsub == 1:
            subjects = [_rand(rng, 128)]
        want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
        with pkg.Engine() as e:
            if variant != "auto":
                e.set_kernel_name(variant)
            got = e.score(queries, subjects)
            assert e.last_kernel_name.endswith("+direct"), e.last_kernel_name
            assert e.device_error_bits == 0
        np.testing.assert_array_equal(got, want, err_msg=f"{variant} {qlens}")


def test_small_path_golden_and_streaming(golden, oracle_mod, pkg):
    """config 2 through the latency path == the reference's golden files; two small batches in
    flight; the regular path gives the same matrix."""
    q = golden["fasta"]["query100.fa"][0][1]
    db = golden["fasta"]["data500.fa"]
    ss = dict([x for x in golden["ssearch"] if x["file"] == "score500.txt"][0]["rows"])
    rng = random.Random(8)
    b2 = [_rand(rng, rng.randint(1, 200)) for _ in range(77)]
    want2 = _oracle_matrix(oracle_mod, pkg, [q], b2)
    with pkg.Engine() as e:
        e.set_queries([q])
        e.score_batch([s for _, s in db], ids=np.arange(len(db), dtype=np.uint64) + 100)
        assert e.last_kernel_name.endswith("+direct")
        e.score_batch(b2)
        assert e.batches_in_flight == 2
        got = e.fetch()
        assert e.fetch_ids()[3] == 103
        got2 = e.fetch()
        assert e.last_kernel_ms > 0
        e.set_small_batch_path(False)
        ref = e.score([q], [s for _, s in db])
        assert "+direct" not in e.last_kernel_name
    assert dict(zip([n for n, _ in db], got[0].tolist())) == ss
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(got2, want2)


def test_small_path_completion_protocols_and_reuse(oracle_mod, pkg, monkeypatch):
    """Latency path: completion read from the sentinel-filled result words (default) or from the flag
    (SW_B200_SMALL_SENTINEL=0, and whenever a query is empty); staging read in place (default) or
    copied (SW_B200_SMALL_ZEROCOPY=0); many batches through the same two slots, alternating with
    regular-path batches that reuse the slots; both strands."""
    rng = random.Random(91)
    queries = [_rand(rng, 100), _rand(rng, 37)]
    batches = [[_rand(rng, rng.choice([0, 1, 3, 64, 128, rng.randint(1, 400)])) for _ in range(rng.randint(1, 90))] for _ in range(12)]
    big = [_rand(rng, rng.randint(50, 300)) for _ in range(9000)]             # too many subjects: regular path
    wants = [_oracle_matrix(oracle_mod, pkg, queries, b) for b in batches]
    want_big = _oracle_matrix(oracle_mod, pkg, queries, big)
    for sentinel, zero_copy in (("1", "1"), ("0", "1"), ("1", "0"), ("0", "0")):
        monkeypatch.setenv("SW_B200_SMALL_SENTINEL", sentinel)
        monkeypatch.setenv("SW_B200_SMALL_ZEROCOPY", zero_copy)
        with pkg.Engine() as e:
            e.set_queries(queries)
            for k, b in enumerate(batches):
                e.score_batch(b)
                assert e.last_kernel_name.endswith("+direct")
                if k % 2:                                  # two in flight, fetched in order
                    np.testing.assert_array_equal(e.fetch(), wants[k - 1])
                    np.testing.assert_array_equal(e.fetch(), wants[k])
                if k == 5:                                 # a regular-path batch reuses the slots
                    np.testing.assert_array_equal(e.score(queries, big), want_big)
                    assert "+direct" not in e.last_kernel_name
                    e.set_queries(queries)
            assert e.device_error_bits == 0
    # an empty query (flag protocol by construction) and both strands
    with pkg.Engine() as e:
        got = e.score([queries[0], "", queries[1]], batches[3])
        assert e.last_kernel_name.endswith("+direct")
        np.testing.assert_array_equal(got[[0, 2]], wants[3])
        assert not got[1].any()
    with pkg.Engine() as e:
        e.set_strands(True)
        got = e.score(queries, batches[4])
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        rc = ["".join(comp[c] for c in reversed(q)) for q in queries]
        want = _oracle_matrix(oracle_mod, pkg, queries + rc, batches[4])
        assert got.shape == want.shape
        np.testing.assert_array_equal(got, want)


def _overflow_case(rng):
    s = _rand(rng, 6700)                      # identical pair scores 33500 > 32767
    t = _mutate(rng, s, 0.004, 0.002)         # around the threshold
    u = _mutate(rng, s, 0.10, 0.05)           # well inside 16 bit
    others = [_rand(rng, rng.randint(100, 2000)) for _ in range(9)]
    return s, [s, t, u, "ACGT", s[:6560], s[100:6652] + "A"] + others


def test_output_i16_and_overflow_side_list(oracle_mod, pkg):
    """int16 matrix: half the bytes; scores above 32767 are -1 in the matrix and exact in the side list."""
    rng = random.Random(5)
    s, subjects = _overflow_case(rng)
    queries = [s, _rand(rng, 150)]
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    assert want[0, 0] == 33500 and (want > 32767).sum() >= 2
    for resident in (False, True):
        with pkg.Engine() as e:
            e.set_output(pkg.SW_OUTPUT_I16)
            e.set_queries(queries)
            if resident:
                e.load_db(subjects)
                e.score_db()
                got = e.fetch_db()
            else:
                e.score_batch(subjects)
                got = e.fetch()
            idx, sc = e.fetch_overflow()
            assert e.device_error_bits == 0
        assert got.dtype == np.int16
        big = want > 32767
        assert np.array_equal(got[~big].astype(np.int32), want[~big])
        assert np.all(got[big] == -1)
        flat = sorted(zip(idx.tolist(), sc.tolist()))
        assert flat == sorted((int(i), int(want.reshape(-1)[i])) for i in np.nonzero(big.reshape(-1))[0])
    # int32 output of the same batch is exact everywhere (list-driven 32-bit fix-up)
    with pkg.Engine() as e:
        np.testing.assert_array_equal(e.score(queries, subjects), want)
    # plain case: no overflow possible, int16 == int32
    q2 = [_rand(rng, 150) for _ in range(3)]
    s2 = [_rand(rng, rng.randint(0, 300)) for _ in range(5000)]
    w2 = _oracle_matrix(oracle_mod, pkg, q2, s2)
    with pkg.Engine() as e:
        e.set_output(pkg.SW_OUTPUT_I16)
        g2 = e.score(q2, s2)
        assert e.fetch_overflow()[0].size == 0
    np.testing.assert_array_equal(g2.astype(np.int32), w2)


@pytest.mark.parametrize("choice", ["auto", "strip_s16x2_R25x2_G1", "strip_s16x2_R16x1_G32", "strip_s16x2_R38x1_G4"])
def test_topk_vs_oracle_argsort(oracle_mod, pkg, choice):
    """Fused per-query top-k in the strip epilogue (ScoreBank_v2.v:42-43 max / vld_max, generalised):
    equals the oracle's (score desc, index asc) ranking; no matrix is fetched.  Duplicated subjects
    create ties; empty subjects score 0."""
    rng = random.Random(1234)
    queries = [_rand(rng, n) for n in (150, 40, 151, 300, 75)]
    subjects = []
    for i in range(3000):
        r = rng.random()
        if r < 0.15:
            subjects.append(_mutate(rng, rng.choice(queries), 0.1, 0.05) or "A")
        elif r < 0.25 and subjects:
            subjects.append(rng.choice(subjects))            # exact duplicate: tie on score
        elif r < 0.28:
            subjects.append("")
        else:
            subjects.append(_rand(rng, rng.randint(1, 260)))
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    for k in (1, 5, 16, 32):
        wsc, wix = _topk_want(want, k)
        with pkg.Engine() as e:
            if choice != "auto":
                e.set_kernel_name(choice)
            e.set_topk(k)
            e.set_queries(queries)
            e.load_db(subjects)
            e.score_db()
            sc, ix = e.fetch_db_topk()
            assert e.device_error_bits == 0
            with pytest.raises(pkg.SwError):
                e.fetch_db()                                  # there is no matrix in top-k mode
        np.testing.assert_array_equal(sc, wsc, err_msg=f"k={k}")
        np.testing.assert_array_equal(ix, wix, err_msg=f"k={k}")


def test_topk_streaming_small_and_sharded(oracle_mod, pkg):
    """Top-k for streaming batches (two in flight), with fewer subjects than k, all-empty batches, and
    one handle whose database is split in three shards (gather + merge across shards)."""
    rng = random.Random(77)
    queries = [_rand(rng, 120), _rand(rng, 33)]
    batches = [[_rand(rng, rng.randint(0, 200)) for _ in range(n)] for n in (900, 3, 0, 450)]
    batches.append(["", "", ""])
    k = 8
    for gpu_ids in ([0], [0, 0, 0]):
        with pkg.Engine(gpu_ids=gpu_ids) as e:
            e.set_topk(k)
            e.set_queries(queries)
            got = []
            e.score_batch(batches[0])
            for b in batches[1:]:
                e.score_batch(b)
                got.append(e.fetch_topk())
            got.append(e.fetch_topk())
        for b, (sc, ix) in zip(batches, got):
            want = _oracle_matrix(oracle_mod, pkg, queries, b) if b else np.zeros((2, 0), np.int32)
            wsc, wix = _topk_want(want, k)
            np.testing.assert_array_equal(sc, wsc, err_msg=str((gpu_ids, len(b))))
            np.testing.assert_array_equal(ix, wix, err_msg=str((gpu_ids, len(b))))


def test_topk_with_scores_beyond_int16(oracle_mod, pkg):
    rng = random.Random(5)
    s, subjects = _overflow_case(rng)
    want = _oracle_matrix(oracle_mod, pkg, [s], subjects)
    wsc, wix = _topk_want(want, 4)
    with pkg.Engine() as e:
        e.set_topk(4)
        e.set_queries([s])
        e.score_batch(subjects)
        sc, ix = e.fetch_topk()
    assert sc[0, 0] == 33500
    np.testing.assert_array_equal(sc, wsc)
    np.testing.assert_array_equal(ix, wix)


def test_virtual_multi_shard_handle(oracle_mod, pkg):
    """One handle, database sharded over several per-GPU contexts with a host-side gather (SURVEY 8e;
    ScoreBank_v2.v:78-139,162 at the GPU level).  The same GPU listed several times gives every
    shard its own streams and buffers, so the shard plan, the per-shard sort and the column-range
    gather run through the CUDA path on a one-GPU box as well; with more GPUs they are used too."""
    rng = random.Random(4)
    queries = [_rand(rng, 150) for _ in range(4)]
    subjects = [_rand(rng, rng.randint(0, 300)) for _ in range(3001)]
    subjects[1700] = queries[2]
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    n = pkg.device_count()
    for gpu_ids in ([0, 0], [0, 0, 0, 0, 0], list(range(n)) + [0]):
        with pkg.Engine(gpu_ids=gpu_ids) as e:
            e.set_small_batch_path(False)
            got = e.score(queries, subjects)
            np.testing.assert_array_equal(got, want, err_msg=str(gpu_ids))
            e.set_queries(queries)
            e.load_db(subjects, ids=np.arange(len(subjects), dtype=np.uint64) * 3)
            e.score_db()
            e.wait()
            np.testing.assert_array_equal(e.fetch_db(), want)
            bs, bi = e.fetch_best()
            assert bs.tolist() == want.max(axis=1).tolist() and bi.tolist() == want.argmax(axis=1).tolist()
            assert int(bi[2]) == 1700 and e.fetch_ids()[1700] == 5100
            st = e.stats()
            assert st["kernel_ms_max"] >= st["kernel_ms_min"] > 0
            e.set_output(pkg.SW_OUTPUT_I16)
            e.score_db()
            np.testing.assert_array_equal(e.fetch_db().astype(np.int32), want)


def test_query_groups_get_their_own_variant(oracle_mod, pkg):
    """Mixed query lengths (config 5): queries are grouped by the variant whose pass height fits them
    and every group is launched with its own variant; the matrix stays in input order."""
    rng = random.Random(55)
    nrng = np.random.default_rng(55)
    queries = [_rand(rng, n) for n in (32, 4096, 150, 1000, 31, 2049, 150)]
    ns = 120000
    lens = nrng.integers(200, 900, size=ns).astype(np.uint32)
    nbytes = (lens.astype(np.uint64) + 3) // 4
    off = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.uint64)
    packed = nrng.integers(0, 256, size=int(nbytes.sum()) + 16, dtype=np.uint8)
    tail = (lens % 4).astype(np.int64)
    last = (off + nbytes - 1).astype(np.int64)
    m = tail > 0
    packed[last[m]] &= ((1 << (2 * tail[m])) - 1).astype(np.uint8)
    # plant homologs of the queries (byte-aligned copies of query stretches)
    qp = pkg.pack_sequences(queries)
    for k in range(0, ns, 997):
        qi = (k // 997) % len(queries)
        qb = int((qp[1][qi] + 3) // 4)
        n = min(int(nbytes[k]) - 1, qb - 1)
        if n > 4:
            packed[int(off[k]): int(off[k]) + n] = qp[0][int(qp[2][qi]): int(qp[2][qi]) + n]
    db = (packed, lens, off)
    with pkg.Engine() as e:
        e.set_launch_plan(1, True)           # one launch per (length group, query group), concurrent streams
        got = e.score(qp, db)
        name = e.last_kernel_name
        assert e.device_error_bits == 0
    assert "+groups" in name, name
    # the oracle checks a sample of the columns (the full matrix would take minutes on the CPU)
    cols = np.array(sorted(set(range(0, ns, 997)) | set(rng.sample(range(ns), 250))))
    sub = (np.concatenate([packed[int(off[c]): int(off[c]) + int(nbytes[c])] for c in cols] + [np.zeros(16, np.uint8)]),
           lens[cols], np.concatenate([[0], np.cumsum(nbytes[cols])[:-1]]).astype(np.uint64))
    want, _ = oracle_mod.Oracle().score_batch_packed(qp[0], qp[1], qp[2], sub[0], sub[1], sub[2])
    np.testing.assert_array_equal(got[:, cols], want)
    assert want.max() > 1000
    with pkg.Engine() as e:
        e.set_kernel_name("strip_s16x2_R25x3_G1")
        one = e.score(qp, db)
        e.set_kernel_name("")
        e.set_launch_plan(1, False)
        two = e.score(qp, db)
        assert "+groups" in e.last_kernel_name
        e.set_launch_plan(2, False)          # the default: one launch here
        three = e.score(qp, db)
        assert "+groups" not in e.last_kernel_name
    np.testing.assert_array_equal(got, one)
    np.testing.assert_array_equal(got, two)
    np.testing.assert_array_equal(got, three)


def test_load_db_right_after_async_score_db(oracle_mod, pkg):
    """sw_score_db is asynchronous: a following sw_load_db / sw_score_batch must not overwrite the
    slot's buffers under the running kernels (uploads wait for the slot's scoring to finish)."""
    rng = random.Random(9)
    q = pkg.random_packed_db(24, 150, seed=3)
    db1 = pkg.random_packed_db(400000, 150, seed=4)
    small = [_rand(rng, rng.randint(1, 200)) for _ in range(9000)]
    with pkg.Engine() as e:
        e.set_queries(q)
        e.load_db(db1)
        e.score_db()                      # ~30 ms of kernels in flight
        e.load_db(small)                  # same slot, different shape
        e.score_db()
        got = e.fetch_db()
        e.load_db(db1)
        e.score_db()
        e.score_batch(small)              # slot 0 again (nothing in flight in the fifo)
        got2 = e.fetch()
        assert e.device_error_bits == 0
    o = oracle_mod.Oracle()
    sp = pkg.pack_sequences(small)
    want_small, _ = o.score_batch_packed(q[0], q[1], q[2], sp[0], sp[1], sp[2])
    np.testing.assert_array_equal(got, want_small)
    np.testing.assert_array_equal(got2, want_small)


def test_long_subject_overflow_fixup_needs_little_scratch(oracle_mod, pkg):
    """A 300 kb subject next to a 7 kb query whose copy it contains: the score leaves 16 bits, the
    pair goes to the overflow list and the list-driven 32-bit kernel (shorter sequence as columns,
    bounded scratch) recomputes it -- no allocation proportional to the longest subject x all threads."""
    rng = random.Random(12)
    q = _rand(rng, 7000)
    genome = _rand(rng, 150000) + q + _rand(rng, 143000)
    subjects = [genome, _rand(rng, 5000), _mutate(rng, q, 0.2, 0.1), "ACGT"]
    want = _oracle_matrix(oracle_mod, pkg, [q], subjects)
    assert want[0, 0] == 35000
    with pkg.Engine() as e:
        got = e.score([q], subjects)
    np.testing.assert_array_equal(got, want)


def test_runtime_specialised_penalties(oracle_mod, pkg):
    """Any penalty set gets immediates: the variant a job uses is compiled at run time with the
    handle's penalties (ld_penalties is a run-time bus in the reference, ScoreBank_v2.v:34,161)."""
    if not pkg.load_library().sw_jit_is_available():
        pytest.skip("NVRTC not loadable on this box")
    params = dict(match=5, mismatch=-4, gap_open=-10, gap_extend=-3)
    ok, msg = pkg.jit_compile_check("strip_s16x2_R25x2_G1", -10, -3)
    assert ok == 1, msg
    rng = random.Random(3)
    queries = [_rand(rng, 150), _rand(rng, 90)]
    subjects = [(_mutate(rng, rng.choice(queries), 0.15, 0.15) or "A") for _ in range(20000)]
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects, **params)
    with pkg.Engine(**params) as e:
        e.set_jit(2)
        got = e.score(queries, subjects)
        assert "+jit" in e.last_kernel_name, e.last_kernel_name
        e.set_jit(0)
        plain = e.score(queries, subjects)
        assert "+jit" not in e.last_kernel_name
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(plain, want)


def test_mixed_one_very_long_subject_with_many_short(oracle_mod, pkg):
    """One 60 kb subject next to 10^5 short ones and a multi-pass query: the launch plan must not size
    the pass-boundary scratch of every resident block by the longest subject."""
    rng = random.Random(60)
    query = _rand(rng, 400)
    genome = _rand(rng, 30000) + _mutate(rng, query, 0.03, 0.02) + _rand(rng, 30000)
    db = pkg.random_packed_db(100000, 150, seed=61)
    gp = pkg.pack_sequences([genome])
    nb = 38
    packed = np.concatenate([db[0][:100000 * nb], gp[0]])
    ln = np.concatenate([db[1], gp[1]]).astype(np.uint32)
    off = np.concatenate([db[2], np.array([100000 * nb], dtype=np.uint64)])
    qp = pkg.pack_sequences([query])
    with pkg.Engine() as e:
        got = e.score(qp, (packed, ln, off))
        assert e.device_error_bits == 0
    o = oracle_mod.Oracle()
    idx = np.arange(0, 100000, 41)
    sub = (np.concatenate([packed[i * nb:(i + 1) * nb] for i in idx] + [np.zeros(16, np.uint8)]),
           ln[idx], (np.arange(len(idx), dtype=np.uint64) * nb))
    want, _ = o.score_batch_packed(qp[0], qp[1], qp[2], sub[0], sub[1], sub[2])
    np.testing.assert_array_equal(got[:, idx], want)
    assert int(got[0, 100000]) == o.score(query, genome) > 1500


@pytest.mark.parametrize("sticky", ["-1", "0", "2", "1000000"])
@pytest.mark.parametrize("nq,ns", [(3, 20000), (40, 8000), (300, 3000)])
def test_work_order_switch_vs_oracle(oracle_mod, pkg, monkeypatch, sticky, nq, ns):
    """The work order of a strip launch (one queue per query with a drift bound, or super-blocks) only
    changes who scores what when: every order gives the oracle's matrix.  3 and 40 equally long queries
    take the per-query queues (with / without the drift check, which covers at most 32 queries), 300
    queries exceed the per-launch limit of the sticky order; ragged subjects so that the warps run
    general and interior trips; SW_B200_STICKY: -1 automatic bound, 0 super-block order, 2 a tight
    bound (blocks keep moving to the slowest queue), 10^6 no bound."""
    monkeypatch.setenv("SW_B200_STICKY", sticky)
    monkeypatch.setenv("SW_B200_SUPERBLOCK_MB", "0.05" if sticky == "0" else "24")
    rng = np.random.default_rng(700 + nq)
    qp, ql, qo = pkg.random_packed_db(nq, 150, seed=70 + nq)
    packed, ln, off = pkg.random_packed_db(ns, 150, seed=71 + nq)
    nb = 38
    rows = packed[:ns * nb].reshape(ns, nb).copy()
    for k in range(0, ns, 37):                      # planted near-copies of a query: gaps and high scores
        rows[k] = qp[(k % nq) * nb:(k % nq + 1) * nb]
        rows[k, rng.integers(0, nb)] ^= np.uint8(rng.integers(1, 255))
    ln = rng.integers(1, 151, size=ns).astype(np.uint32)
    ln[::5] = 150
    flat = np.concatenate([rows.reshape(-1), np.zeros(16, np.uint8)])
    o = oracle_mod.Oracle()
    want, _ = o.score_batch_packed(qp, ql, qo, flat, ln, off)
    with pkg.Engine() as e:
        e.set_small_batch_path(False)
        got = e.score((qp, ql, qo), (flat, ln, off))
        assert e.device_error_bits == 0
        e.set_output(pkg.SW_OUTPUT_I16)
        got16 = e.score((qp, ql, qo), (flat, ln, off))
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got16.astype(np.int32), want)
    assert want.max() > 500


@pytest.mark.parametrize("variant", ["strip_s16x2_R25x2_G1", "strip_s16x2_R25x2_G1_U4_F31", "strip_s16x2_R38x2_G1", "strip_s16x2_R32x2_G1",
                                     "strip_s16x2_R25x3_G2", "strip_s16x2_R19x2_G4", "strip_s16x2_R16x1_G32"])
@pytest.mark.parametrize("parts", [3, 64])
def test_pass_split_forced_vs_oracle(oracle_mod, pkg, variant, parts):
    """Pass split (sw_set_pass_split): the passes of a work item are handed out as separate items whose
    boundary row and running maximum wait in per-chain scratch.  Forced here for multi-chunk queries of
    DIFFERENT lengths (one of them single-pass: its later parts are empty and only carry the maximum),
    ragged subjects with planted homologs, in every output mode; 64 parts = one profile chunk per part."""
    rng = random.Random(4100 + parts)
    queries = [_rand(rng, n) for n in (1500, 700, 64, 1501, 2304)]
    subjects = []
    for i in range(2500):
        r = rng.random()
        if r < 0.2:
            q = rng.choice(queries)
            a = rng.randint(0, max(0, len(q) - 80))
            subjects.append(_mutate(rng, q[a:a + rng.randint(40, 500)], 0.06, 0.04) or "A")
        elif r < 0.22:
            subjects.append("")
        else:
            subjects.append(_rand(rng, rng.randint(1, 420)))
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    assert want.max() > 1200
    with pkg.Engine() as e:
        e.set_small_batch_path(False)
        e.set_wave_mode(0)
        e.set_kernel_name(variant)
        e.set_pass_split(parts)
        got = e.score(queries, subjects)
        assert e.last_pass_parts >= 2, e.last_kernel_name
        assert e.device_error_bits == 0
        e.set_output(pkg.SW_OUTPUT_I16)
        got16 = e.score(queries, subjects)
        e.set_output(pkg.SW_OUTPUT_I32)
        e.set_topk(7)
        e.score_batch(subjects)
        tsc, tix = e.fetch_topk()
        e.set_topk(0)
        e.set_pass_split(0)
        plain = e.score(queries, subjects)
        assert e.last_pass_parts == 1
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got16.astype(np.int32), want)
    np.testing.assert_array_equal(plain, want)
    wsc, wix = _topk_want(want, 7)
    np.testing.assert_array_equal(tsc, wsc)
    np.testing.assert_array_equal(tix, wix)


def test_pass_split_automatic_on_few_rounds_of_long_items(oracle_mod, pkg):
    """The automatic rule: one 3 000-nt query against 200 000 short subjects is a few rounds of multi-chunk
    work items -> split; the same database against 150-nt queries (single pass) is not.  Scores of a
    sample of subjects equal the oracle's, and the split and unsplit matrices are identical."""
    rng = random.Random(77)
    query = _rand(rng, 3000)
    db = pkg.random_packed_db(200000, 100, seed=78)
    qp = pkg.pack_sequences([query])
    with pkg.Engine() as e:
        e.set_wave_mode(0)
        got = e.score(qp, db)
        parts = e.last_pass_parts
        assert e.device_error_bits == 0
        e.set_pass_split(0)
        plain = e.score(qp, db)
        assert e.last_pass_parts == 1
        e.set_pass_split(-1)
        short = e.score(pkg.random_packed_db(4, 150, seed=79), db)
        assert e.last_pass_parts == 1 and short.shape == (4, 200000)
    assert parts >= 2
    np.testing.assert_array_equal(got, plain)
    nb = 25
    idx = np.arange(0, 200000, 100)
    sub = (np.concatenate([db[0][i * nb:(i + 1) * nb] for i in idx] + [np.zeros(16, np.uint8)]),
           db[1][idx], (np.arange(len(idx), dtype=np.uint64) * nb))
    o = oracle_mod.Oracle()
    want, _ = o.score_batch_packed(qp[0], qp[1], qp[2], sub[0], sub[1], sub[2])
    np.testing.assert_array_equal(got[:, idx], want)


def test_bounds_check_build_runs_clean(pkg):
    """Stand-in for compute-sanitizer (closed on this pool): the same sources built with
    -DSW_BOUNDS_CHECK (device-side index checks on tp / bnd / profile / out, canaries around every
    device buffer) run the ragged parity tests; any violation turns into SW_EDEVICE."""
    lib = os.path.join(ROOT, "smith-waterman-fpga-module_b200", "libsw_b200_check.so")
    if not os.path.exists(lib):
        pytest.skip("libsw_b200_check.so not built (make -C csrc CHECK=1)")
    if pkg.is_check_build():
        pytest.skip("already running the check build")
    env = dict(os.environ, SW_B200_LIB=lib)
    sel = ("test_random_mixed_lengths_vs_oracle or test_long_query_multi_pass_and_chunks or test_edge_cases or "
           "test_small_path_direct_variants_vs_oracle or test_topk_vs_oracle_argsort or test_output_i16 or "
           "test_very_long_subjects_short_queries or test_query_groups or test_is_check_build or test_wave_kernel or "
           "test_randomised_modes_stress or test_virtual_multi_shard or test_overflow_list_scored_by_32bit_bands or "
           "test_overflow_32bit_bands_sharded or test_small_path_completion_protocols or test_overflow_list_more_entries or "
           "test_topk_of_few_long_pairs or test_pass_split or test_work_order_switch")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-x", "-q", "-m", "gpu", "-k", sel,
                        "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout


def test_is_check_build_flag(pkg):
    """In the check build every fetch also verifies canaries and the device error word."""
    want = os.environ.get("SW_B200_LIB", "").endswith("_check.so")
    assert pkg.is_check_build() == want
    with pkg.Engine() as e:
        e.score(["ACGTACGT"], ["ACGTTCGT", "", "A"])
        assert e.device_error_bits == 0


def test_wave_kernel_few_long_pairs_vs_oracle(oracle_mod, pkg):
    """Band-pipelined kernel (sw_wave.cuh): the 512-row bands of a long query run concurrently on
    different warps, each consuming the bottom row of the band above as it is produced
    (ScoringModule_v1.1.v:36-39,49-54: the module-chaining ports left "for future use")."""
    rng = random.Random(404)
    q = _rand(rng, 3000)                                    # 6 bands
    q2 = _rand(rng, 1100)                                   # 3 bands, the last one almost empty
    subjects = []
    for k in range(61):                                     # odd count: one pair has a single member
        L = rng.choice([1, 5, 31, 32, 33, 64, 100, 511, 700, 1000, rng.randint(1, 1200)])
        if k % 3 == 0:
            a = rng.randint(0, 2000)
            s = (_mutate(rng, q[a:a + L], 0.05, 0.03) + _rand(rng, L))[:L]
        else:
            s = _rand(rng, L)
        subjects.append(s or "A")
    subjects += ["", q[500:2500], _mutate(rng, q, 0.02, 0.01)]
    want = _oracle_matrix(oracle_mod, pkg, [q, q2, q[:600]], subjects)
    assert want.max() > 9000
    for mode, out_mode in ((2, pkg.SW_OUTPUT_I32), (2, pkg.SW_OUTPUT_I16), (1, pkg.SW_OUTPUT_I32)):
        with pkg.Engine() as e:
            e.set_small_batch_path(False)
            e.set_wave_mode(mode)
            e.set_output(out_mode)
            got = e.score([q, q2, q[:600]], subjects)
            assert "wave" in e.last_kernel_name, e.last_kernel_name
            assert e.device_error_bits == 0
        np.testing.assert_array_equal(got.astype(np.int32), want, err_msg=str((mode, out_mode)))
    # the strip kernel alone gives the same matrix
    with pkg.Engine() as e:
        e.set_small_batch_path(False)
        e.set_wave_mode(0)
        np.testing.assert_array_equal(e.score([q, q2, q[:600]], subjects), want)
        assert "wave" not in e.last_kernel_name


@pytest.mark.parametrize("instance", [0, 1, 2, 3])
def test_wave_kernel_every_instance_vs_oracle(oracle_mod, pkg, monkeypatch, instance):
    """Each instance of the band-pipelined kernel (512-row bands, one column per step; 256-row bands
    with four / two columns per step and one pair per block; 512-row bands with two columns per
    step) forced in turn over ragged lengths:
    subject lengths around the 4- and 8-column trip, the 32-column boundary block and the 128-column
    code-word block; query lengths that leave the last band almost empty or exactly full."""
    monkeypatch.setenv("SW_B200_WAVE_INSTANCE", str(instance))
    rng = random.Random(4100 + instance)
    q = _rand(rng, 2049)
    queries = [q, q[:1024], q[:769], _rand(rng, 513)]
    subjects = []
    for L in (1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 500, 777, 1031):
        a = rng.randint(0, 1000)
        s = (_mutate(rng, q[a:a + L], 0.05, 0.03) + _rand(rng, L))[:L] if L % 2 else _rand(rng, L)
        subjects.append(s)
    subjects += ["", q[100:1900], _mutate(rng, q, 0.02, 0.01)]
    want = _oracle_matrix(oracle_mod, pkg, queries, subjects)
    with pkg.Engine() as e:
        e.set_small_batch_path(False)
        e.set_wave_mode(2)
        got = e.score(queries, subjects)
        assert "wave" in e.last_kernel_name, e.last_kernel_name
        assert e.device_error_bits == 0
        np.testing.assert_array_equal(got, want)
        # a second call on the same handle: the boundary tags of the first launches must not match
        got = e.score(queries[::-1], subjects)
        np.testing.assert_array_equal(got, want[::-1])


def test_topk_of_few_long_pairs_uses_the_band_pipelined_kernel(oracle_mod, pkg):
    """Top-k for a handful of long subjects: the long queries run on the band-pipelined kernel (a
    scratch score row per query, its k best joined to the key lists), the short query on the strip
    kernel, in the same call; one pair's score is beyond 16 bits (overflow list -> 32-bit bands);
    empty subjects; a sharded handle merges the lists of its shards."""
    rng = random.Random(406)
    q1 = _rand(rng, 7000)
    q2 = _rand(rng, 2600)
    q3 = _rand(rng, 90)
    subjects = [q1[100:6900], _mutate(rng, q1, 0.05, 0.02), _rand(rng, 3000), "", _mutate(rng, q2, 0.1, 0.05), q2[500:2100],
                _rand(rng, 40), q3 + _rand(rng, 200), "", _rand(rng, 1500), _mutate(rng, q1[2000:5000], 0.02, 0.01)]
    subjects += [_rand(rng, rng.randint(1, 900)) for _ in range(20)]
    want = _oracle_matrix(oracle_mod, pkg, [q1, q2, q3], subjects)
    assert want[0, 0] == 34000
    for k in (1, 5, 32):
        wsc, wix = _topk_want(want, k)
        for gpu_ids in (None, [0, 0]):
            with pkg.Engine(gpu_ids=gpu_ids) as e:
                e.set_small_batch_path(False)
                e.set_topk(k)
                e.set_queries([q1, q2, q3])
                e.score_batch(subjects)
                sc, ix = e.fetch_topk()
                assert "wave" in e.last_kernel_name, e.last_kernel_name
                assert e.device_error_bits == 0
            np.testing.assert_array_equal(sc, wsc, err_msg=str((k, gpu_ids)))
            np.testing.assert_array_equal(ix, wix, err_msg=str((k, gpu_ids)))


def test_wave_kernel_single_long_pair_and_overflow(oracle_mod, pkg):
    """One long pair spread over many warps; and a pair whose score leaves the 16-bit range inside
    the wave kernel (flagged per band, recomputed in 32 bit from the overflow list)."""
    rng = random.Random(405)
    a = _rand(rng, 6000)
    b = _mutate(rng, a, 0.03, 0.02)                         # score ~ 25 000: still 16-bit
    c = _rand(rng, 5000)
    o = oracle_mod.Oracle()
    with pkg.Engine() as e:
        got = e.score([a], [b])
        assert "wave" in e.last_kernel_name
        assert int(got[0, 0]) == o.score(a, b) > 20000
        got = e.score([a], [c])
        assert int(got[0, 0]) == o.score(a, c)
        big = _rand(rng, 7000)
        got = e.score([big], [big, c, big[:6900]])
        assert got[0].tolist() == [35000, o.score(big, c), 34500]
        assert e.device_error_bits == 0


def test_overflow_list_scored_by_32bit_bands(oracle_mod, pkg):
    """Scores beyond 16 bits: the long entries of the overflow list go to the band-pipelined 32-bit
    scorer (sw_wave32_kernel: 256-row bands on many warps, list-driven, slots of boundary rows reused
    by later entries), short ones to the one-thread scorer -- same numbers either way, in the int32
    matrix, the int16 side list and the top-k lists."""
    rng = random.Random(77)
    big = _rand(rng, 7000)
    # 150 exact substrings of the query, 6 600 .. 7 000 nt: score = 5 x length, every one overflows;
    # more entries than boundary-row slots (64), so the slot chain is exercised
    subs, want = [], []
    for k in range(150):
        L = rng.randint(6600, 7000)
        a0 = rng.randint(0, 7000 - L)
        subs.append(big[a0:a0 + L])
        want.append(5 * L)
    subs += [_rand(rng, 3000), "", _mutate(rng, big, 0.3, 0.1)]
    o = oracle_mod.Oracle()
    want += [o.score(big, subs[-3]), 0, o.score(big, subs[-1])]
    want = np.array([want], dtype=np.int32)
    assert (want > 32767).sum() == 150
    for enable in (True, False):
        with pkg.Engine() as e:
            e.set_overflow_wave(enable)
            got = e.score([big], subs)
            assert e.device_error_bits == 0
            np.testing.assert_array_equal(got, want, err_msg=f"wave32={enable}")
    with pkg.Engine() as e:                                   # int16 matrix + side list
        e.set_output(pkg.SW_OUTPUT_I16)
        got = e.score([big], subs)
        idx, sc = e.fetch_overflow()
        assert np.all(got[0, :150] == -1) and np.array_equal(got[0, 150:].astype(np.int32), want[0, 150:])
        assert sorted(zip(idx.tolist(), sc.tolist())) == [(i, int(want[0, i])) for i in range(150)]
    with pkg.Engine() as e:                                   # top-k sees the recomputed scores
        e.set_topk(5)
        e.set_queries([big])
        e.score_batch(subs)
        sc, ix = e.fetch_topk()
        order = sorted(range(len(subs)), key=lambda i: (-int(want[0, i]), i))[:5]
        assert ix[0].tolist() == order and sc[0].tolist() == [int(want[0, i]) for i in order]


def test_overflow_32bit_bands_sharded_and_streaming(oracle_mod, pkg):
    """The 32-bit band scorer of the overflow list on a handle with three (virtual) shards, and on
    streaming batches with two in flight (the scorer's state and boundary rows are per GPU, shared
    by the slots)."""
    rng = random.Random(79)
    big = _rand(rng, 7000)
    subs, want = [], []
    for k in range(40):
        if k % 3 == 0:
            L = rng.randint(6600, 7000)
            a0 = rng.randint(0, 7000 - L)
            subs.append(big[a0:a0 + L]); want.append(5 * L)
        else:
            subs.append(_rand(rng, rng.randint(1, 400))); want.append(None)
    o = oracle_mod.Oracle()
    want = np.array([[w if w is not None else o.score(big, s) for w, s in zip(want, subs)]], dtype=np.int32)
    with pkg.Engine(gpu_ids=[0, 0, 0]) as e:
        np.testing.assert_array_equal(e.score([big], subs), want)
        assert e.device_error_bits == 0
    with pkg.Engine() as e:
        e.set_small_batch_path(False)
        e.set_queries([big])
        e.score_batch(subs[:21])
        e.score_batch(subs[21:])
        np.testing.assert_array_equal(e.fetch(), want[:, :21])
        np.testing.assert_array_equal(e.fetch(), want[:, 21:])
        assert e.device_error_bits == 0


def test_overflow_list_mixed_long_and_short_entries(oracle_mod, pkg):
    """match = 100: a 330-nt identity already leaves 16 bits, so the list mixes entries below and
    above the cell threshold of the 32-bit band scorer; both scorers share one list and one call."""
    rng = random.Random(78)
    q1 = _rand(rng, 2100)
    q2 = _rand(rng, 400)
    subs = []
    for k in range(120):
        src = q1 if k % 2 else q2
        L = rng.randint(340, len(src))
        a0 = rng.randint(0, len(src) - L)
        subs.append(src[a0:a0 + L])
    subs += [_rand(rng, 500), _mutate(rng, q1, 0.05, 0.02)]
    with pkg.Engine(match=100, mismatch=-40, gap_open=-120, gap_extend=-20) as e:
        want = _oracle_matrix(oracle_mod, pkg, [q1, q2], subs, match=100, mismatch=-40, gap_open=-120, gap_extend=-20)
        assert (want > 32767).sum() >= 120
        for enable, cells in ((True, 1000000), (True, 1), (False, 1000000)):
            e.set_overflow_wave(enable, cells)
            got = e.score([q1, q2], subs)
            assert e.device_error_bits == 0
            np.testing.assert_array_equal(got, want, err_msg=str((enable, cells)))


def test_overflow_list_more_entries_than_one_launch_of_the_band_scorer(pkg):
    """5 000 overflow entries: the 32-bit band scorer takes 4 096 list entries per launch (12 tag bits)
    and runs a few launches per call.  match = 100, subjects = exact substrings of the query."""
    rng = random.Random(80)
    q = _rand(rng, 700)
    subs, want = [], []
    for k in range(5000):
        L = rng.randint(340, 700)
        a0 = rng.randint(0, 700 - L)
        subs.append(q[a0:a0 + L]); want.append(100 * L)
    want = np.array([want], dtype=np.int32)
    with pkg.Engine(match=100, mismatch=-40, gap_open=-120, gap_extend=-20) as e:
        e.set_small_batch_path(False)
        for cells in (1, 1000000000):                       # all entries by bands / all by one thread each
            e.set_overflow_wave(True, cells)
            got = e.score([q], subs)
            assert e.device_error_bits == 0
            np.testing.assert_array_equal(got, want, err_msg=str(cells))


def test_randomised_modes_stress_vs_oracle(oracle_mod, pkg):
    """Seeded property test over the round-2 switches: output mode (int32 / int16 / top-k), launch plan
    (one launch / length groups / query groups), band-pipelined kernel on or forced, latency path on
    or off, one or three shards, resident or streaming -- against the oracle."""
    rng = random.Random(20161018)
    for it in range(36):
        match = rng.randint(1, 7)
        mismatch = -rng.randint(0, 6)
        gap_extend = -rng.randint(0, 4)
        gap_open = -rng.randint(0, 12)
        params = dict(match=match, mismatch=mismatch, gap_open=gap_open, gap_extend=gap_extend)
        nq = rng.randint(1, 4)
        queries = [_rand(rng, rng.choice([1, 17, 32, 150, 513, 700, rng.randint(1, 1400)])) for _ in range(nq)]
        ns = rng.choice([1, 2, 7, 60, 300, 900])
        subjects = []
        for _ in range(ns):
            if rng.random() < 0.3:
                src = rng.choice(queries)
                s = _rand(rng, rng.randint(0, 20)) + _mutate(rng, src, 0.1, 0.1) + _rand(rng, rng.randint(0, 20))
            else:
                s = _rand(rng, rng.choice([0, 1, 3, rng.randint(1, 400), rng.randint(1, 1500)]))
            subjects.append(s)
        want = _oracle_matrix(oracle_mod, pkg, queries, subjects, **params)
        mode = rng.choice(["i32", "i32", "i16", "topk"])
        k = rng.choice([1, 3, 16, 32])
        shards = rng.choice([[0], [0], [0, 0, 0]])
        resident = rng.random() < 0.5
        desc = f"it {it}: {params} nq={nq} ns={ns} mode={mode} k={k} shards={len(shards)} resident={resident}"
        with pkg.Engine(gpu_ids=shards, **params) as e:
            e.set_small_batch_path(rng.random() < 0.6)
            e.set_wave_mode(rng.choice([0, 1, 2]))
            e.set_launch_plan(rng.choice([0, 1, 2]), rng.random() < 0.5)
            if mode == "i16":
                e.set_output(pkg.SW_OUTPUT_I16)
            if mode == "topk":
                e.set_topk(k)
            e.set_queries(queries)
            if resident:
                e.load_db(subjects)
                e.score_db()
                got = e.fetch_db_topk() if mode == "topk" else e.fetch_db()
            else:
                e.score_batch(subjects)
                got = e.fetch_topk() if mode == "topk" else e.fetch()
            desc += " " + e.last_kernel_name
            assert e.device_error_bits == 0, desc
        if mode == "topk":
            wsc, wix = _topk_want(want, k)
            np.testing.assert_array_equal(got[0], wsc, err_msg=desc)
            np.testing.assert_array_equal(got[1], wix, err_msg=desc)
        else:
            np.testing.assert_array_equal(got.astype(np.int32), want, err_msg=desc)
