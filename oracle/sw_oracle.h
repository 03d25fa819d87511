/*
 * sw_oracle.h -- CPU oracle for the score-only Smith-Waterman hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The shipped engine
 * (libsw_b200.so) never links, loads or calls it and has no CPU fallback.
 *
 * The oracle restates, in scalar C, the arithmetic of the reference's
 * systolic processing element:
 *   ScoreBank/SW_ProcessingElement_v1.0.v:119-129  (stage 1: LUT, diag_max, M_open, I_extend)
 *   ScoreBank/SW_ProcessingElement_v1.0.v:287-291  (stage 2: M clamp, I = max(M_open, I_extend))
 *   ScoreBank/SW_ProcessingElement_v1.0.v:411-420  (stage 3: running high score)
 *   ScoreBank/ScoringModule_v1.1.v:107,176-179     (boundary row = ZERO, result at PE[qlen-1])
 * Parity status: PINNED -- checked against every golden vector the reference
 * ships (730 RTL-simulation pairs, 598 ssearch36 scores, 16 swalign scores,
 * 1 CAPI end-to-end result); see tests/test_oracle_golden.py.
 */
#ifndef SW_ORACLE_H_
#define SW_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int match;        /* ScoreBank_v1_tb.sv:16  default  5 */
    int mismatch;     /* ScoreBank_v1_tb.sv:17  default -4 */
    int gap_open;     /* ScoreBank_v1_tb.sv:18  default -12 */
    int gap_extend;   /* ScoreBank_v1_tb.sv:19  default -4 */
    int score_width;  /* 0 = exact integers; W>0 = bit-accurate W-bit biased machine
                         (SW_ProcessingElement_v1.0.v:15,20 -- the RTL uses W = 12) */
    int first_col_v03;/* 1 = PE v0.3 first-column variant
                         (capi_sample_aligner/hdl-verliog/SW_ProcessingElement_v0.3.v:145-158) */
} swo_params_t;

/* One pair, sequences given as one 2-bit code per byte (only equality matters,
 * SW_ProcessingElement_v1.0.v:119).  Returns the unbiased score (result - ZERO). */
int32_t swo_score_codes(const uint8_t *q, int m, const uint8_t *t, int n,
                        const swo_params_t *p);

/* One pair, ASCII input (A/C/G/T any case; anything else packs as code 0 like
 * aligner_Header.c:38-39). */
int32_t swo_score_ascii(const char *q, int m, const char *t, int n,
                        const swo_params_t *p);

/* Whole score matrix on 2-bit packed input (LSB-first, aligner_Header.c:27-37):
 * out[iq * ns + is].  nthreads <= 0 -> all OpenMP threads.  Returns the number
 * of threads actually used. */
int swo_score_batch_packed(const uint8_t *qpacked, const uint32_t *qlen, const uint64_t *qoff, int nq,
                           const uint8_t *tpacked, const uint32_t *tlen, const uint64_t *toff, size_t ns,
                           const swo_params_t *p, int32_t *out, int nthreads);

/* ASCII -> 2-bit LSB-first packing used by the batch entry (aligner_Header.c:14-47). */
void swo_pack_2bit(const char *seq, size_t len, uint8_t *out);

int swo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
