/*
 * sw_kernels.h -- host-side launch interface of the CUDA kernels (internal to
 * libsw_b200.so; the public boundary is include/sw_b200.h).
 */
#ifndef SW_KERNELS_H_
#define SW_KERNELS_H_

#include <cuda_runtime.h>
#include <stdint.h>

#define SW_NO_SUBJECT 0xFFFFFFFFu
#define SW_OVERFLOW_SENTINEL (-1)   /* strip kernel: 16-bit range possibly exceeded, recompute in 32 bit */

/* Database shard resident in HBM.  Layout (DESIGN.md "data layout"):
 *   raw/off/len   : the caller's 2-bit packed records, as uploaded
 *   pair_subj     : [2*npairs] shard-local subject index of the low / high 16-bit
 *                   lane of pair p (SW_NO_SUBJECT = lane unused)
 *   pair_len      : [npairs] columns of the pair (both members have this length)
 *   tp            : column codes, one byte per column (0..15 = t_lo | t_hi << 2; 16..19 = the
 *                   shorter member has ended; 20 = no column), four columns per 32-bit
 *                   word, tiles of 32 pairs, word k of the 32 pairs of a tile contiguous:
 *                   tp[tile_woff[tile] + k*32 + slot]
 */
struct SwDevDb {
    const uint8_t  *raw;
    const uint64_t *off;
    const uint32_t *len;
    uint32_t        ns;
    const uint32_t *pair_subj;
    const uint32_t *pair_len;
    const uint64_t *tile_woff;
    uint32_t       *tp;
    uint32_t        npairs;
    uint32_t        max_len;
};

struct SwDevQueries {
    const uint8_t  *packed;   /* 2-bit packed, each query byte aligned */
    const uint32_t *off;      /* byte offset of each query             */
    const uint32_t *len;
    int             nq;
    uint32_t        max_len;
};

struct SwScoring {
    int match, mismatch, goe /* gap_open + gap_extend */, ge;
    int limit;                /* 0 = exact; else 2^(W-1)-1: M above it restarts at 0 */
};

/* One strip-kernel variant = (rows per lane R = RS*S, lanes per pair G). */
struct SwStripVariant {
    int R, G;
    int block_threads;
    int S;            /* independent sub-strips per lane (R = RS * S) */
    int min_blocks;   /* resident blocks per SM the kernel was compiled for */
    const char *name;
};

int sw_strip_variant_count(void);
/* testing hook: true = never use the instances with compile-time gap penalties */
void sw_strip_disable_fixed(bool off);
const SwStripVariant *sw_strip_variant(int idx);

/* Dynamic shared memory of variant idx when the query profile of chunk_passes passes is resident. */
size_t sw_strip_smem_bytes(int idx, int chunk_passes);

/* Resident blocks per SM of variant idx with smem_bytes of dynamic shared memory. */
cudaError_t sw_strip_occupancy(int idx, size_t smem_bytes, int *blocks_per_sm);

/* Scores queries [q0,q1) against all pairs of db.  out[(q)*out_stride + subj].
 * bnd: scratch for pass boundaries, grid * bnd_cols * (block_threads/G) uint2.
 * counter: zeroed device word (work queue).  chunk_passes: passes (of R*G rows) whose
 * query profile is held in shared memory at once. */
cudaError_t sw_launch_strip(int idx, cudaStream_t st, const SwDevDb &db, const SwDevQueries &q,
                            int q0, int q1, const SwScoring &sc, int32_t *out, size_t out_stride,
                            uint2 *bnd, uint32_t bnd_cols, unsigned *counter, int grid,
                            int chunk_passes);

/* 32-bit kernel: any length, any score range.  scratch: 2 * (db.max_len) * threads int32.
 * fix_only: recompute only the entries the strip kernel marked SW_OVERFLOW_SENTINEL. */
cudaError_t sw_launch_generic32(cudaStream_t st, const SwDevDb &db, const SwDevQueries &q,
                                int q0, int q1, const SwScoring &sc, int32_t *out, size_t out_stride,
                                int32_t *scratch, int threads_total, bool fix_only);

cudaError_t sw_launch_build_tp(cudaStream_t st, const SwDevDb &db);

/* Per-query arg-max over subjects (first index reaching the max). */
cudaError_t sw_launch_best(cudaStream_t st, const int32_t *scores, size_t stride, uint32_t ns,
                           int nq, int32_t *best_score, uint32_t *best_index);

#endif
