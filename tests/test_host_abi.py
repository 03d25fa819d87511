"""CPU-side tests: the C-ABI library loads and exports every symbol include/sw_b200.h declares,
the pure-host helpers (packing, FASTA, text writers) behave like the reference's, and the
product refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "sw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(sw_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libsw_b200.so does not export {n}"


def test_variant_list_matches_library(pkg):
    from tests.test_gpu_parity import STRIP_VARIANTS
    assert pkg.kernel_variants() == STRIP_VARIANTS


def test_no_cpu_fallback(pkg):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(pkg.SwError) as ei:
        pkg.Engine()
    assert ei.value.code == pkg.SW_ENODEV


def test_product_never_references_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or load it."""
    pkgdir = os.path.join(ROOT, "smith-waterman-fpga-module_b200")
    for dp, _dn, fns in os.walk(pkgdir):
        for fn in fns:
            if fn.endswith((".py", ".c", ".cu", ".h", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "sw_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, fn
    out = subprocess.run(["ldd", os.path.join(pkgdir, "libsw_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_pack_matches_reference_check_values(pkg):
    lib = pkg.load_library()
    buf = (C.c_uint8 * 1)()
    lib.sw_pack_2bit(b"AGGG", 4, buf)
    assert buf[0] == 0xFE                      # build/main_test_output.txt: "AGGG" -> 0xfe
    for seq in ["A", "ACGTACGTAC", "ttgacN", "GATTACA" * 9]:
        n = len(seq)
        b = (C.c_uint8 * ((n + 3) // 4))()
        lib.sw_pack_2bit(seq.encode(), n, b)
        packed, ln, off = pkg.pack_sequences([seq])
        assert bytes(b) == packed[: (n + 3) // 4].tobytes()
        back = C.create_string_buffer(n + 1)
        lib.sw_unpack_2bit(b, n, back)
        assert back.value.decode() == seq.upper().replace("N", "T")


class SeqSet(C.Structure):
    _fields_ = [("n", C.c_size_t), ("packed", C.POINTER(C.c_uint8)), ("len", C.POINTER(C.c_uint32)),
                ("off", C.POINTER(C.c_uint64)), ("name", C.POINTER(C.c_char_p)), ("packed_bytes", C.c_size_t)]


def _read_fasta(lib, path):
    lib.sw_read_fasta.argtypes = [C.c_char_p, C.POINTER(C.POINTER(SeqSet))]
    lib.sw_seqset_free.argtypes = [C.POINTER(SeqSet)]
    lib.sw_seqset_free.restype = None
    p = C.POINTER(SeqSet)()
    rc = lib.sw_read_fasta(path.encode(), C.byref(p))
    return rc, p


def test_read_fasta_like_the_testbench(pkg, golden, tmp_path):
    lib = pkg.load_library()
    recs = golden["fasta"]["data1.fa"]
    f = tmp_path / "data1.fa"
    f.write_text("".join(f">{n}\n{s}\n" for n, s in recs))
    rc, p = _read_fasta(lib, str(f))
    assert rc == 0 and p.contents.n == len(recs) == 20
    for i, (n, s) in enumerate(recs):
        assert p.contents.name[i].decode() == n
        assert p.contents.len[i] == len(s)
        back = C.create_string_buffer(len(s) + 1)
        lib.sw_unpack_2bit(C.cast(C.addressof(p.contents.packed.contents) + p.contents.off[i], C.c_void_p), len(s), back)
        assert back.value.decode() == s.upper()
    lib.sw_seqset_free(p)
    # wrapped lines are concatenated; header-less files yield the first token (main_test.c:304)
    g = tmp_path / "wrapped.fa"
    g.write_text(">x some description\nACGT\nTTGA\n\n>y\nGG\n")
    rc, p = _read_fasta(lib, str(g))
    assert rc == 0 and p.contents.n == 2 and p.contents.len[0] == 8 and p.contents.name[0] == b"x"
    lib.sw_seqset_free(p)
    hless = tmp_path / "query"
    hless.write_text(golden["capi"]["query"] + "\n")
    rc, p = _read_fasta(lib, str(hless))
    assert rc == 0 and p.contents.n == 1 and p.contents.len[0] == 32
    lib.sw_seqset_free(p)
    rc, p = _read_fasta(lib, str(tmp_path / "missing.fa"))
    assert rc == pkg.SW_EIO


def test_out_txt_layout_matches_golden_file(pkg, golden, tmp_path):
    """sw_write_out_txt reproduces data1.fa_query1.fa_out.txt byte for byte when given the RTL's
    completion times (ScoreBank_v1_tb.sv:280-281)."""
    lib = pkg.load_library()
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    lib.sw_write_out_txt.argtypes = [C.c_void_p, C.POINTER(SeqSet), C.c_void_p, C.c_void_p]
    rtl = [s for s in golden["rtl"] if s["file"] == "data1.fa_query1.fa_out.txt"][0]
    seqs = dict(golden["fasta"]["data1.fa"])
    fa = tmp_path / "ordered.fa"
    fa.write_text("".join(f">{n}\n{seqs[n]}\n" for n, _s, _t in rtl["rows"]))
    rc, p = _read_fasta(lib, str(fa))
    assert rc == 0
    scores = np.array([s for _n, s, _t in rtl["rows"]], dtype=np.int32)
    times = np.array([t for _n, _s, t in rtl["rows"]], dtype=np.uint64)
    outp = tmp_path / "out.txt"
    fh = libc.fopen(str(outp).encode(), b"w")
    assert lib.sw_write_out_txt(fh, p, scores.ctypes.data, times.ctypes.data) == 0
    libc.fclose(fh)
    lines = outp.read_text().splitlines()
    assert lines[:3] == golden["format_samples"]["out_txt_first_lines"]
    lib.sw_seqset_free(p)


def test_ssearch_R_rows_have_score_in_sixth_field(pkg, golden, tmp_path):
    lib = pkg.load_library()
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    lib.sw_write_ssearch_R.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(SeqSet), C.POINTER(SeqSet), C.c_void_p]
    ss = [s for s in golden["ssearch"] if s["file"] == "score500.txt"][0]
    db = golden["fasta"]["data500.fa"]
    fa = tmp_path / "db.fa"
    fa.write_text("".join(f">{n}\n{s}\n" for n, s in db))
    qf = tmp_path / "q.fa"
    qf.write_text(">query\n" + golden["fasta"]["query100.fa"][0][1] + "\n")
    rc, pdb = _read_fasta(lib, str(fa))
    rc2, pq = _read_fasta(lib, str(qf))
    assert rc == 0 and rc2 == 0
    want = dict((n, s) for n, s in ss["rows"])
    scores = np.array([want[n] for n, _ in db], dtype=np.int32)
    outp = tmp_path / "score.txt"
    fh = libc.fopen(str(outp).encode(), b"w")
    assert lib.sw_write_ssearch_R(fh, b"query100.fa", b"data500.fa", pq, pdb, scores.ctypes.data) == 0
    libc.fclose(fh)
    rows = [l.split() for l in outp.read_text().splitlines() if not l.startswith(("#", ">"))]
    assert len(rows) == 499
    for r in rows:
        assert int(r[5]) == want[r[0]] and int(r[1]) == 128
    # same parser as tests/golden/make_golden.py accepts the reference's own file layout
    ref_line = golden["format_samples"]["ssearch_R_first_lines"][2].split()
    assert ref_line[0] == rows[0][0] and ref_line[1:5] == rows[0][1:5] and int(ref_line[5]) == want["db1"]
    lib.sw_seqset_free(pdb)
    lib.sw_seqset_free(pq)


def test_random_packed_db_layout(pkg):
    packed, ln, off = pkg.random_packed_db(10, 150, seed=1)
    assert ln.tolist() == [150] * 10 and off.tolist() == [38 * i for i in range(10)]
    assert (packed[37::38][:10] & 0xF0).sum() == 0           # unused tail bits are zero
    p2, _, _ = pkg.random_packed_db(10, 150, seed=1)
    assert np.array_equal(packed, p2)


def test_cli_fails_loudly_without_inputs_or_gpu(pkg, golden, tmp_path):
    """The CLI (counterpart of main_test -q -l -t): missing inputs print the reference's message;
    without a GPU it stops at sw_init instead of falling back to anything."""
    cli = os.path.join(ROOT, "bin", "sw_b200_cli")
    assert os.path.exists(cli)
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode != 0 and "Input files missing" in r.stdout          # main_test.c:281-285
    r = subprocess.run([cli, "-q", str(tmp_path / "nope.fa"), "-l", str(tmp_path / "nope2.fa")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Query file error!" in r.stdout            # main_test.c:246-250
    if pkg.device_count() == 0:
        q = tmp_path / "q.fa"
        q.write_text(">query\nACGTACGT\n")
        r = subprocess.run([cli, "-q", str(q), "-l", str(q)], capture_output=True, text=True)
        assert r.returncode != 0 and "no usable CUDA device" in r.stdout
