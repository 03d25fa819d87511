/*
 * sw_oracle.c -- scalar CPU restatement of the reference's processing element.
 *
 * TEST INFRASTRUCTURE ONLY (see sw_oracle.h).  Never linked into the product.
 *
 * Geometry: query position i = one PE (ScoringModule_v1.1.v:155-235), target
 * position j = one time step.  PE i sees, while target base j is on its input:
 *   M_in/I_in   = output of PE i-1 for the same target base  -> cell (i-1, j)
 *   M_out/I_out = its own previous output                    -> cell (i,   j-1)
 *   M_diag/I_diag = M_in/I_in latched one step earlier       -> cell (i-1, j-1)
 *                   (SW_ProcessingElement_v1.0.v:184-185)
 * PE 0 has M_in = I_in = High_in = ZERO (ScoringModule_v1.1.v:176-179) and an idle
 * PE drives ZERO (SW_ProcessingElement_v1.0.v:192-198), so every boundary is 0.
 */
#include "sw_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static inline uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

/* Exact-integer form.  Column-major sweep like the hardware:
 * outer loop = time (target base), inner loop = PE index. */
static int32_t score_exact(const uint8_t *q, int m, const uint8_t *t, int n,
                           const swo_params_t *p, int32_t *buf)
{
    /* buf holds M and I of the previous column, index 0 = boundary row. */
    int32_t *Mp = buf, *Ip = buf + (m + 1);
    const int32_t goe = p->gap_open + p->gap_extend;   /* v1.0.v:128 */
    const int32_t ge = p->gap_extend;                  /* v1.0.v:129 */
    const int32_t ma = p->match, mi = p->mismatch;
    int32_t high = 0;
    for (int i = 0; i <= m; ++i) { Mp[i] = 0; Ip[i] = 0; }
    for (int j = 0; j < n; ++j) {
        const uint8_t tj = t[j];
        int32_t m_up = 0, i_up = 0;       /* cell (i-1, j)   : M_in, I_in   */
        int32_t m_dg = 0, i_dg = 0;       /* cell (i-1, j-1) : M_diag, I_diag */
        const int first = (j == 0) && p->first_col_v03;
        for (int i = 1; i <= m; ++i) {
            const int32_t m_lf = Mp[i], i_lf = Ip[i];          /* cell (i, j-1): M_out, I_out */
            const int32_t lut = (q[i - 1] == tj) ? ma : mi;     /* v1.0.v:119 */
            const int32_t diag_max = imax(m_dg, i_dg);          /* v1.0.v:123 */
            int32_t m_open, i_ext;
            if (first) {                                        /* v0.3.v:153-156 */
                m_open = goe;
                i_ext = ge;
            } else {
                m_open = imax(m_up, m_lf) + goe;                /* v1.0.v:127-128 */
                i_ext = imax(i_up, i_lf) + ge;                  /* v1.0.v:126,129 */
            }
            const int32_t m_sc = lut + diag_max;                /* v1.0.v:287 */
            const int32_t mm = m_sc > 0 ? m_sc : 0;             /* v1.0.v:288 */
            const int32_t ii = imax(m_open, i_ext);             /* v1.0.v:291 */
            high = imax(high, imax(mm, ii));                    /* v1.0.v:411-420 */
            m_dg = m_lf; i_dg = i_lf;
            Mp[i] = mm;  Ip[i] = ii;
            m_up = mm;   i_up = ii;
        }
    }
    return high;
}

/* Bit-accurate W-bit machine: values are W-bit unsigned,
 * biased by ZERO = 2^(W-1); adds wrap mod 2^W; compares are unsigned (`MAX macro,
 * v1.0.v:11); the local-alignment clamp is "MSB clear => ZERO" (v1.0.v:288). */
static int32_t score_wbit(const uint8_t *q, int m, const uint8_t *t, int n,
                          const swo_params_t *p, int32_t *buf)
{
    const int W = p->score_width;
    const uint32_t mask = (W >= 32) ? 0xFFFFFFFFu : ((1u << W) - 1u);
    const uint32_t ZERO = 1u << (W - 1);
    const uint32_t ma = (uint32_t)p->match & mask, mi = (uint32_t)p->mismatch & mask;
    const uint32_t go = (uint32_t)p->gap_open & mask, ge = (uint32_t)p->gap_extend & mask;
    uint32_t *Mp = (uint32_t *)buf, *Ip = (uint32_t *)buf + (m + 1);
    uint32_t high = ZERO;
    for (int i = 0; i <= m; ++i) { Mp[i] = ZERO; Ip[i] = ZERO; }
    for (int j = 0; j < n; ++j) {
        const uint8_t tj = t[j];
        uint32_t m_up = ZERO, i_up = ZERO, m_dg = ZERO, i_dg = ZERO;
        const int first = (j == 0) && p->first_col_v03;
        for (int i = 1; i <= m; ++i) {
            const uint32_t m_lf = Mp[i], i_lf = Ip[i];
            const uint32_t lut = (q[i - 1] == tj) ? ma : mi;
            const uint32_t diag_max = umax(m_dg, i_dg);
            uint32_t m_open, i_ext;
            if (first) {
                m_open = (ZERO + go + ge) & mask;
                i_ext = (ZERO + ge) & mask;
            } else {
                m_open = (umax(m_up, m_lf) + go + ge) & mask;
                i_ext = (umax(i_up, i_lf) + ge) & mask;
            }
            const uint32_t m_sc = (lut + diag_max) & mask;
            const uint32_t mm = (m_sc & ZERO) ? m_sc : ZERO;
            const uint32_t ii = umax(m_open, i_ext);
            high = umax(high, umax(mm, ii));
            m_dg = m_lf; i_dg = i_lf;
            Mp[i] = mm;  Ip[i] = ii;
            m_up = mm;   i_up = ii;
        }
    }
    return (int32_t)high - (int32_t)ZERO;   /* ScoreBank_v1_tb.sv:280, main_test.c:528 */
}

static int32_t score_with_buf(const uint8_t *q, int m, const uint8_t *t, int n,
                              const swo_params_t *p, int32_t *buf)
{
    if (m <= 0 || n <= 0) return 0;   /* no cell is ever computed: High stays ZERO */
    return p->score_width > 0 ? score_wbit(q, m, t, n, p, buf) : score_exact(q, m, t, n, p, buf);
}

int32_t swo_score_codes(const uint8_t *q, int m, const uint8_t *t, int n, const swo_params_t *p)
{
    if (m <= 0 || n <= 0) return 0;
    int32_t *buf = (int32_t *)malloc(sizeof(int32_t) * 2 * ((size_t)m + 1));
    if (!buf) return -1;
    int32_t r = score_with_buf(q, m, t, n, p, buf);
    free(buf);
    return r;
}

/* aligner_Header.c:34-39 : A=10 C=01 G=11 T=00, everything else 00 */
static inline uint8_t code_of(char c)
{
    switch (c) {
        case 'a': case 'A': return 2;
        case 'c': case 'C': return 1;
        case 'g': case 'G': return 3;
        default: return 0;
    }
}

int32_t swo_score_ascii(const char *q, int m, const char *t, int n, const swo_params_t *p)
{
    if (m <= 0 || n <= 0) return 0;
    uint8_t *cq = (uint8_t *)malloc((size_t)m + (size_t)n);
    if (!cq) return -1;
    uint8_t *ct = cq + m;
    for (int i = 0; i < m; ++i) cq[i] = code_of(q[i]);
    for (int j = 0; j < n; ++j) ct[j] = code_of(t[j]);
    int32_t r = swo_score_codes(cq, m, ct, n, p);
    free(cq);
    return r;
}

void swo_pack_2bit(const char *seq, size_t len, uint8_t *out)
{
    memset(out, 0, (len + 3) / 4);
    for (size_t i = 0; i < len; ++i)
        out[i >> 2] |= (uint8_t)(code_of(seq[i]) << ((i * 2) & 7));   /* aligner_Header.c:27,34 */
}

static void unpack_2bit(const uint8_t *packed, uint32_t len, uint8_t *codes)
{
    for (uint32_t i = 0; i < len; ++i) codes[i] = (packed[i >> 2] >> ((i * 2) & 7)) & 3;
}

int swo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int swo_score_batch_packed(const uint8_t *qpacked, const uint32_t *qlen, const uint64_t *qoff, int nq,
                           const uint8_t *tpacked, const uint32_t *tlen, const uint64_t *toff, size_t ns,
                           const swo_params_t *p, int32_t *out, int nthreads)
{
    uint32_t maxq = 0;
    for (int i = 0; i < nq; ++i) if (qlen[i] > maxq) maxq = qlen[i];
    uint32_t maxt = 0;
    for (size_t s = 0; s < ns; ++s) if (tlen[s] > maxt) maxt = tlen[s];

    /* unpack all queries once */
    uint8_t **qc = (uint8_t **)calloc((size_t)(nq > 0 ? nq : 1), sizeof(uint8_t *));
    for (int i = 0; i < nq; ++i) {
        qc[i] = (uint8_t *)malloc(qlen[i] ? qlen[i] : 1);
        unpack_2bit(qpacked + qoff[i], qlen[i], qc[i]);
    }
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        uint8_t *tc = (uint8_t *)malloc(maxt ? maxt : 1);
        int32_t *buf = (int32_t *)malloc(sizeof(int32_t) * 2 * ((size_t)maxq + 1));
#ifdef _OPENMP
#pragma omp single
        used = omp_get_num_threads();
#pragma omp for schedule(dynamic, 16)
#endif
        for (long long s = 0; s < (long long)ns; ++s) {
            unpack_2bit(tpacked + toff[s], tlen[s], tc);
            for (int i = 0; i < nq; ++i)
                out[(size_t)i * ns + (size_t)s] =
                    score_with_buf(qc[i], (int)qlen[i], tc, (int)tlen[s], p, buf);
        }
        free(tc);
        free(buf);
    }
    for (int i = 0; i < nq; ++i) free(qc[i]);
    free(qc);
    return used;
}
