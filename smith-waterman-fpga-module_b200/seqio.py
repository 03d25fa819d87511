"""Host-side sequence plumbing for tests and the bench: 2-bit packing in numpy
(same layout as sw_pack_2bit / aligner_Header.c:14-47) and synthetic databases
(same model as data/generate.py:6-23: iid uniform A/C/G/T)."""
import numpy as np

_CODE = np.zeros(256, dtype=np.uint8)
for _ch, _c in (("A", 2), ("C", 1), ("G", 3), ("T", 0)):
    _CODE[ord(_ch)] = _c
    _CODE[ord(_ch.lower())] = _c
LETTERS = np.frombuffer(b"TCAG", dtype=np.uint8)     # code -> letter


def pack_codes(codes):
    """codes: 1-D uint8 array of 2-bit codes -> packed bytes (LSB-first, 4 per byte)."""
    n = len(codes)
    pad = (-n) % 4
    c = np.concatenate([codes.astype(np.uint8), np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)


def pack_sequences(seqs):
    """list of str -> (packed uint8, len uint32, off uint64); every record byte-aligned."""
    bufs, lens, offs, off = [], [], [], 0
    for s in seqs:
        b = s.encode() if isinstance(s, str) else bytes(s)
        codes = _CODE[np.frombuffer(b, dtype=np.uint8)] if len(b) else np.zeros(0, np.uint8)
        p = pack_codes(codes)
        bufs.append(p)
        lens.append(len(b))
        offs.append(off)
        off += len(p)
    packed = np.concatenate(bufs) if bufs else np.zeros(0, np.uint8)
    packed = np.concatenate([packed, np.zeros(16, np.uint8)])      # slack for vector loads
    return packed, np.array(lens, dtype=np.uint32), np.array(offs, dtype=np.uint64)


def random_packed_db(n, length, seed):
    """n iid-uniform sequences of fixed length, generated directly in packed form.
    Returns (packed, len, off).  Tail bits of the last byte of a record are zero."""
    rng = np.random.default_rng(seed)
    nbytes = (length + 3) // 4
    packed = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
    tail = length % 4
    if tail:
        packed[:, -1] &= np.uint8((1 << (2 * tail)) - 1)
    ln = np.full(n, length, dtype=np.uint32)
    off = (np.arange(n, dtype=np.uint64) * np.uint64(nbytes))
    flat = np.concatenate([packed.reshape(-1), np.zeros(16, np.uint8)])
    return flat, ln, off


def unpack_codes(packed2d, length):
    """[n, nbytes] packed rows -> [n, length] 2-bit codes."""
    idx = np.arange(length)
    return (packed2d[:, idx >> 2] >> ((idx & 3) * 2).astype(np.uint8)) & 3


def plant_homologs(db, queries, frac, seed, psub=0.05, pindel=0.02):
    """Replaces a fraction of the fixed-length subjects of `db` by mutated copies of the queries
    (substitutions + indels), in place on the packed form, so that the
    gap path of the recurrence carries real alignments.  Every round(1/frac)-th subject is planted."""
    packed, ln, off = db
    qpacked, qln, qoff = queries
    n, length = len(ln), int(ln[0])
    nb = (length + 3) // 4
    qlen = int(qln[0])
    assert np.all(ln == length) and np.all(qln == qlen) and qlen >= length
    rng = np.random.default_rng(seed)
    step = max(1, int(round(1.0 / frac)))
    rows = np.arange(0, n, step)
    qcodes = unpack_codes(qpacked[: len(qln) * ((qlen + 3) // 4)].reshape(len(qln), -1), qlen)
    src = qcodes[rng.integers(0, len(qln), size=len(rows))][:, :length].copy()        # [k, length]
    sub = rng.random(src.shape) < psub
    src[sub] = (src[sub] + rng.integers(1, 4, size=int(sub.sum()))) & 3
    # indels: delete one base and shift left (deletion) or shift right and insert (insertion)
    nind = rng.binomial(length, pindel, size=len(rows))
    for r in np.nonzero(nind)[0]:
        for _ in range(int(nind[r])):
            p = int(rng.integers(0, length))
            if rng.random() < 0.5:
                src[r, p:-1] = src[r, p + 1:]
                src[r, -1] = rng.integers(0, 4)
            else:
                src[r, p + 1:] = src[r, p:-1]
                src[r, p] = rng.integers(0, 4)
    pad = (-length) % 4
    c = np.concatenate([src, np.zeros((len(rows), pad), src.dtype)], axis=1).reshape(len(rows), -1, 4).astype(np.uint8)
    rows_packed = c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)
    view = packed[: n * nb].reshape(n, nb)
    view[rows] = rows_packed
    return len(rows)


def unpack_to_str(packed, length, off=0):
    idx = np.arange(length)
    codes = (packed[off + (idx >> 2)] >> ((idx & 3) * 2)) & 3
    return LETTERS[codes].tobytes().decode()
