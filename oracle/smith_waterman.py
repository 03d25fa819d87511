"""Completed Python form of the reference's ``data/smith-waterman.py``.

TEST INFRASTRUCTURE ONLY -- part of the oracle, never used by the product.

The reference file stops after allocating the two matrices ``M`` and ``I``
(data/smith-waterman.py:5-16).  This module keeps its signature
``localalignment(P, T)``, its constants (:7-10) and its two-matrix layout
(``M[i][j]``, ``I[i][j]``, (len(P)+1) x (len(T)+1), zero initialised) and fills
in the recurrence the RTL implements (ScoreBank/SW_ProcessingElement_v1.0.v:
119-129, 287-291, 411-420).
"""


def localalignment(P, T, match=5, mismatch=-4, gap_open=-12, gap_extend=-4):
    M = []
    I = []
    for i in range(len(P) + 1):  # initialise matrices (smith-waterman.py:13-16)
        M.append([0] * (len(T) + 1))
        I.append([0] * (len(T) + 1))
    best = 0
    P = P.upper()
    T = T.upper()
    for j in range(1, len(T) + 1):          # time step = target base
        for i in range(1, len(P) + 1):      # PE index = query base
            lut = match if P[i - 1] == T[j - 1] else mismatch            # v1.0.v:119
            m = lut + max(M[i - 1][j - 1], I[i - 1][j - 1])              # v1.0.v:123,287
            M[i][j] = m if m > 0 else 0                                  # v1.0.v:288
            m_open = max(M[i - 1][j], M[i][j - 1]) + gap_open + gap_extend   # v1.0.v:127-128
            i_ext = max(I[i - 1][j], I[i][j - 1]) + gap_extend               # v1.0.v:126,129
            I[i][j] = max(m_open, i_ext)                                 # v1.0.v:291
            best = max(best, M[i][j], I[i][j])                           # v1.0.v:411-420
    return best


def gotoh(P, T, match=5, mismatch=-4, gap_open=-12, gap_extend=-4):
    """Textbook three-state affine local alignment (first gap residue costs
    gap_open + gap_extend, the FASTA/ssearch36 convention).  Only used by tests
    that document where the PE recurrence differs from Gotoh."""
    NEG = -10 ** 9
    n, m = len(T), len(P)
    H = [[0] * (n + 1) for _ in range(m + 1)]
    E = [[NEG] * (n + 1) for _ in range(m + 1)]
    F = [[NEG] * (n + 1) for _ in range(m + 1)]
    best = 0
    for i in range(1, m + 1):
        for j in range(1, n + 1):
            E[i][j] = max(E[i][j - 1] + gap_extend, H[i][j - 1] + gap_open + gap_extend)
            F[i][j] = max(F[i - 1][j] + gap_extend, H[i - 1][j] + gap_open + gap_extend)
            s = match if P[i - 1].upper() == T[j - 1].upper() else mismatch
            H[i][j] = max(0, H[i - 1][j - 1] + s, E[i][j], F[i][j])
            best = max(best, H[i][j])
    return best
