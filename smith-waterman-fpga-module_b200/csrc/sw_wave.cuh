/*
 * sw_wave.cuh -- band-pipelined ("wavefront over warps") strip kernel for few, long pairs.
 *
 * The strip kernel gives one warp (G = 32 lanes x R rows = P rows per pass) to a pair and walks the
 * passes of a long query one after the other, so with few pairs most of the GPU idles, and a single
 * pair runs on a single warp.  Here the passes ("bands") of one pair are DIFFERENT work items: band
 * b of a pair can start as soon as band b-1 has produced the first columns of its bottom row, so
 * the bands of one pair run concurrently on different warps / SMs, staggered by a few dozen
 * columns -- the module-chaining ports the reference left "for future use"
 * (ScoringModule_v1.1.v:36-39, 49-54: M_in / I_in / High_in of one module fed by the outputs of
 * another), with L2 as the wire and a tag inside every boundary element as the valid signal.
 *
 *   work item  = (band, block of 4 pairs); items are claimed from an atomic counter in band-major
 *                order, so the item a band waits for (same pairs, band - 1) was always claimed
 *                earlier by a block that is resident and running: no deadlock by construction
 *                (plus a watchdog that raises SW_DEVERR_SPIN instead of hanging the GPU).
 *   boundary   = bottom row (H, G) of a band, per pair and band parity: bnd[pair][band & 1][column],
 *                one 16-byte element {tag:H, tag:G} per column: two 64-bit words, each stored
 *                atomically, each carrying tag = (launch epoch, band).  The consumer loads 32
 *                columns at a time (coalesced ld.cg) and simply retries until every element carries
 *                the tag it expects -- no fence, no separate progress counter on the producer's path.
 *                Band b+2 reuses the slots of band b, but only behind band b+1's read position
 *                (band b+2 lags band b+1 by at least 32 columns plus the 32 S - 1 steps of the
 *                systolic skew, and band b+1 has staged a column before it uses it).
 *   result     = max over the bands: atomicMax per pair; the band that finishes last writes the score.
 * Arithmetic and the one-step-ahead code pipeline are those of the strip kernel (sw_strip.cuh); the
 * profile uses 24 code slots per row pair instead of 32 (48 KB per block: four resident blocks per
 * SM).  Exact arithmetic only (the W-bit mode keeps the strip kernel).
 */
#ifndef SW_WAVE_CUH_
#define SW_WAVE_CUH_

#include "sw_strip.cuh"

namespace swk {

constexpr int kWaveBlock = 32;       // default number of columns staged per block by the consuming band
constexpr int kWaveCodes = 24;       // profile slots per row pair (codes 0 .. 20 are used)
constexpr int kWaveSlack = 8;        // boundary elements of slack on either side of a row (batched stores)
constexpr int kWavePrefetch = 8;     // steps between the prefetch of a boundary block and its use
constexpr int kWaveBotSlots = 8;     // shared-memory slots for the bottom-row values of one loop trip (LS)

struct WaveArgs {
    const uint32_t *tp;
    const uint64_t *tile_woff;
    const uint32_t *pair_len;
    const uint32_t *pair_subj;
    uint32_t npairs, npb;      // npb = ceil(npairs / pairs per block)
    const uint8_t *qpacked;
    const uint32_t *qoff;
    const uint32_t *qlen;
    int q;                     // the query of this launch
    int out_row;               // row of `out` its scores go to (q, or the row of a scratch matrix)
    int npass;                 // bands of that query (of R * 32 rows each), <= 4095
    void *out;
    size_t out_stride;
    int out_mode;              // SW_OUT_I32 / SW_OUT_I16
    ulonglong2 *bnd;           // [pair][2][cols_stride] tagged boundary elements
    uint32_t cols_stride;
    uint32_t epoch;            // launch number (tags of earlier launches never match), 1 .. 2^20 - 1
    int *best;                 // [pair][2]: running maximum of the two members, INT_MAX = 16-bit overflow
    unsigned *done;            // [pair]: bands finished
    unsigned *counter;
    int match, mismatch, goe, ge;
    uint32_t goe2, ge2;
    int ovf_limit;
    uint32_t zero;
    unsigned *ovf_count;
    uint2 *ovf_list;
    unsigned ovf_cap;
    unsigned *dev_err;
    unsigned spin_limit;       // polls of a boundary block before the watchdog gives up
    uint64_t tp_words, bnd_elems, out_elems;   // buffer sizes (only read by the -DSW_BOUNDS_CHECK build)
};

// One band of one pair (one warp).  Returns the band's running maximum (K representation).
template <int RS, int S, class AR, int BLK, bool HAS_TOP, bool HAS_BOTTOM>
__device__ __forceinline__ uint32_t wave_band(const WaveArgs &a, const uint2 *prof_lane, uint2 *s_top, const uint32_t *tpp,
                                              int ncols, const ulonglong2 *top, ulonglong2 *bot, uint32_t tag_top, uint32_t tag_bot,
                                              uint32_t goe2, uint32_t ge2, uint32_t h0, uint32_t gb2, uint32_t zero)
{
    constexpr int G = 32, RP = (RS + 1) / 2, VPE = G * S, U = 4;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const bool head = lane == 0;
    uint32_t best = h0;
    uint32_t H[S][RS], Gl[S][RS];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int r = 0; r < RS; ++r) { H[s][r] = h0; Gl[s][r] = gb2; }
    // code words (4 columns each): lane i holds word 32 B + i of the current / next block of 128
    // columns, loaded 128 columns ahead; the step loop takes one word per four steps by shuffle
    auto bulk = [&](int blk) -> uint32_t {
        const int w = blk * 32 + lane;
        if (4 * w < ncols) SW_CHECK((unsigned long long)(tpp - a.tp) + (unsigned long long)w * 32 < a.tp_words, SW_DEVERR_TP, a);
        return 4 * w < ncols ? __ldg(tpp + (size_t)w * 32) : kPadCode * 0x01010101u;
    };
    uint32_t cw_cur = bulk(0), cw_nxt = bulk(1);
    uint32_t wcur = __shfl_sync(FULL, cw_cur, 0), wnext = __shfl_sync(FULL, cw_cur, 1);
    uint32_t pub_h[S], pub_g[S], pub_t[S], hd_top[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { pub_h[s] = h0; pub_g[s] = gb2; pub_t[s] = kPadCode; hd_top[s] = h0; }
    if (head && ncols > 0) pub_t[0] = wcur & 255u;
    uint2 sv[S][RP];
    load_scores<RS, S, G, kWaveCodes>(sv, prof_lane, pub_t);
    uint2 bcur = make_uint2(h0, gb2);
    ulonglong2 pre = make_ulonglong2(0ull, 0ull);
    uint32_t sto_h[U], sto_g[U];
    const unsigned long long tag_hi = (unsigned long long)tag_bot << 32;
    const int nsteps = ncols > 0 ? (ncols + (VPE - 1) + U - 1) / U * U : 0;

#pragma unroll 1
    for (int t2 = 0; t2 < nsteps; t2 += U) {
        if constexpr (HAS_TOP) {
            if ((t2 & (BLK - 1)) == 0 && t2 < ncols) {
                // stage columns t2 .. t2 + BLK - 1 of the bottom row of the band above: every lane
                // takes one element (prefetched eight steps ago, see below) and retries the load
                // until both of its words carry the producer's tag
                const int c = t2 + lane;
                const bool mine = lane < BLK && c < ncols;
                ulonglong2 e = pre;
                bool have = t2 != 0;
                unsigned spins = 0;
                for (;;) {
                    if (mine) SW_CHECK((unsigned long long)(top + c - a.bnd) < a.bnd_elems, SW_DEVERR_BND, a);
                    if (!have && mine) e = __ldcg(top + c);
                    const bool ok = !mine || ((uint32_t)(e.x >> 32) == tag_top && (uint32_t)(e.y >> 32) == tag_top);
                    if (__all_sync(FULL, ok)) break;
                    have = false;
                    __nanosleep(64);
                    if (++spins > a.spin_limit) {              // watchdog: never hang the GPU
                        if (head && a.dev_err) atomicOr(a.dev_err, SW_DEVERR_SPIN);
                        break;
                    }
                }
                const uint2 v = mine ? make_uint2((uint32_t)e.x, (uint32_t)e.y) : make_uint2(h0, gb2);
                __syncwarp();                                  // the head lane is done with the previous block
                if (lane < BLK) s_top[lane] = v;
                __syncwarp();
            }
            if ((t2 & (BLK - 1)) == BLK - kWavePrefetch && t2 + kWavePrefetch < ncols) {
                // the next block's elements, a few steps before they are needed: the L2 round trip
                // overlaps the steps in between; a stale element fails the tag check above and is re-read
                const int c = t2 + kWavePrefetch + lane;
                if (lane < BLK && c < ncols) pre = __ldcg(top + c);
            }
        }
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {
            const int t = t2 + uu;
            const int u = uu & 3;
            uint32_t in_h[S], in_g[S], in_t[S];
            in_h[0] = __shfl_up_sync(FULL, pub_h[S - 1], 1, G);
            in_g[0] = __shfl_up_sync(FULL, pub_g[S - 1], 1, G);
            in_t[0] = __shfl_up_sync(FULL, pub_t[S - 1], 1, G);
            if constexpr (HAS_TOP) bcur = s_top[t & (BLK - 1)];
            const uint32_t wsel = (u < 3) ? wcur : wnext;
            const uint32_t lead_t = (t + 1 < ncols) ? ((wsel >> (8 * ((u + 1) & 3))) & 255u) : (uint32_t)kPadCode;
            in_h[0] = head ? bcur.x : in_h[0];
            in_g[0] = head ? bcur.y : in_g[0];
            in_t[0] = head ? lead_t : in_t[0];
            if (u == 3) {
                wcur = wnext;
                const int k = (t >> 2) + 2;
                if ((k & 31) == 0) { cw_cur = cw_nxt; cw_nxt = bulk((k >> 5) + 1); }
                wnext = __shfl_sync(FULL, cw_cur, k & 31);
            }
#pragma unroll
            for (int s = 1; s < S; ++s) { in_h[s] = pub_h[s - 1]; in_g[s] = pub_g[s - 1]; in_t[s] = pub_t[s - 1]; }
            SW_CHECK(in_t[0] <= (uint32_t)kPadCode, SW_DEVERR_PROF, a);
            uint2 sv_next[S][RP];
            load_scores<RS, S, G, kWaveCodes>(sv_next, prof_lane, in_t);
            column_step_multi<RS, S, G, AR, false>(H, Gl, best, hd_top, in_g, sv, goe2, ge2, zero, 0u);
#pragma unroll
            for (int s = 0; s < S; ++s) {
#pragma unroll
                for (int k = 0; k < RP; ++k) sv[s][k] = sv_next[s][k];
                hd_top[s] = in_h[s];
                pub_h[s] = H[s][RS - 1]; pub_g[s] = Gl[s][RS - 1]; pub_t[s] = in_t[s];
            }
            sto_h[uu] = pub_h[S - 1];
            sto_g[uu] = pub_g[S - 1];
        }
        if constexpr (HAS_BOTTOM) {
            // the last virtual PE finished columns cl0 .. cl0 + 3 in these four steps: one branch,
            // four 16-byte stores.  Columns outside 0 .. ncols - 1 land in the row's slack
            // (kWaveSlack elements on either side) and are never read.
            const int cl0 = t2 - (VPE - 1);
            if (lane == G - 1 && cl0 + (U - 1) >= 0 && cl0 < ncols) {
#pragma unroll
                for (int uu = 0; uu < U; ++uu) {
                    SW_CHECK((unsigned long long)(bot + cl0 + uu - a.bnd) < a.bnd_elems, SW_DEVERR_BND, a);
                    __stcg(bot + cl0 + uu, make_ulonglong2(tag_hi | sto_h[uu], tag_hi | sto_g[uu]));
                }
            }
        }
    }
    return best;
}

// The same band with C columns per systolic step: a lane finishes the C x RS tile of its (S x RS)
// rows before it hands (H, G) of its bottom row -- C columns at once -- to the next lane, so the
// lanes are skewed by C columns, a band needs ncols / C + 31 steps instead of ncols + 31, and the
// fixed cost of a step (shuffles, selects, the profile address, loop control) is paid once per C
// columns.  Inside the tile column j + 1 of row r only needs column j of row r and column j + 1
// of row r - 1: C interleaved dependency chains, which is what a lone warp on its scheduler needs
// (one long pair = one warp per band: latency-bound, not throughput-bound).
// Codes travel as one word of C bytes, one step ahead of H / G, exactly as in wave_band.
// Code sources of wave_band_c: word(w) = the codes of columns 4 w .. 4 w + 3, one byte each.
struct TiledCodes {                  // a pair's column of the tiled code stream (build_tp_kernel)
    static constexpr int kCodes = kWaveCodes, kPad = kPadCode;
    const uint32_t *tpp;
    __device__ __forceinline__ uint32_t word(int w, int) const { return __ldg(tpp + (size_t)w * 32); }
};
struct RawCodes {                    // one subject's packed 2-bit record: code = nucleotide, 4 = past its end
    static constexpr int kCodes = 8, kPad = 4;
    const uint8_t *raw;
    __device__ __forceinline__ uint32_t word(int w, int ncols) const {
        const uint32_t b = __ldg(raw + w);
        uint32_t x = (b & 3u) | ((b & 12u) << 6) | ((b & 48u) << 12) | ((b & 192u) << 18);
        const int rem = ncols - 4 * w;                         // >= 1 (the caller checks 4 w < ncols)
        if (rem < 4) x = (x & ((1u << (8 * rem)) - 1u)) | ((kPad * 0x01010101u) << (8 * rem));
        return x;
    }
};

template <int RS, int S, int C, class AR, int BLK, bool HAS_TOP, bool HAS_BOTTOM, bool LS, class ARGS, class SRC>
__device__ __forceinline__ uint32_t wave_band_c(const ARGS &a, const uint2 *prof_lane, uint2 *s_top, const SRC src,
                                                int ncols, const ulonglong2 *top, ulonglong2 *bot, uint32_t tag_top, uint32_t tag_bot,
                                                uint32_t goe2, uint32_t ge2, uint32_t h0, uint32_t gb2, uint32_t zero)
{
    static_assert(C == 2 || C == 4, "2 or 4 columns per step");
    constexpr int G = 32, RP = (RS + 1) / 2, VPE = G * S, TC = 8, U = TC / C;   // a loop trip = 8 columns = U steps
    constexpr int PF = 16;                                            // columns between prefetch and use of a boundary block
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr uint32_t PADW = SRC::kPad * 0x01010101u;
    constexpr uint32_t CMASK = C == 4 ? 0xFFFFFFFFu : 0xFFFFu;
    static_assert(BLK % TC == 0 && BLK == 32, "boundary blocks of 32 columns");
    const int lane = threadIdx.x & 31;
    const bool head = lane == 0;
    uint32_t best = h0;
    uint32_t H[S][RS], Gl[S][RS];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int r = 0; r < RS; ++r) { H[s][r] = h0; Gl[s][r] = gb2; }

    // Code words (4 columns each): the warp keeps 2 x 32 of them in one register per lane -- lane i
    // holds word 32 B + i of the current / next block of 128 columns, loaded 128 columns before they
    // are needed (the stream of a lone pair comes from DRAM: a load issued one trip ahead would
    // bound the trip by the memory latency) -- and a trip takes its two words by shuffle.
    auto bulk = [&](int blk) -> uint32_t {
        const int w = blk * 32 + lane;
        return 4 * w < ncols ? src.word(w, ncols) : PADW;
    };
    uint32_t cw_cur = bulk(0), cw_nxt = bulk(1);
    uint32_t tc0 = __shfl_sync(FULL, cw_cur, 0), tc1 = __shfl_sync(FULL, cw_cur, 1);
    uint32_t tn0 = __shfl_sync(FULL, cw_cur, 2), tn1 = __shfl_sync(FULL, cw_cur, 3);

    // what each sub-strip hands to the next virtual PE (the next sub-strip of the lane, or for
    // s = S-1 the next lane): bottom H and G of its C columns, and the code word it uses NEXT step
    uint32_t pub_h[S][C], pub_g[S][C], pub_t[S], hd_carry[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        pub_t[s] = PADW & CMASK;
        hd_carry[s] = h0;
#pragma unroll
        for (int j = 0; j < C; ++j) { pub_h[s][j] = h0; pub_g[s][j] = gb2; }
    }
    if (head) pub_t[0] = tc0 & CMASK;
    auto load_sv = [&](uint2 (&sv)[S][C][RP], const uint32_t (&tw)[S]) {
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const uint2 *prow = prof_lane + (s * RP * SRC::kCodes + ((tw[s] >> (8 * j)) & 255u)) * G;
#pragma unroll
                for (int k = 0; k < RP; ++k) sv[s][j][k] = prow[k * SRC::kCodes * G];
            }
    };
    uint2 sv[S][C][RP];
    load_sv(sv, pub_t);
    ulonglong2 pre = make_ulonglong2(0ull, 0ull);
    uint32_t sto_h[TC], sto_g[TC];
    uint2 *s_bot = s_top + BLK;            // LS: the last lane's bottom-row values of this trip
    const unsigned long long tag_hi = (unsigned long long)tag_bot << 32;
    const int nsteps = ncols > 0 ? ((ncols + C - 1) / C + (VPE - 1) + U - 1) / U * U : 0;

#pragma unroll 1
    for (int k0 = 0; k0 < nsteps; k0 += U) {
        const int c0 = k0 * C;                                 // the head lane's first column of this trip
        const int wq = (c0 >> 2) + 4;                          // the words of the trip after next
        if ((wq & 31) == 0) { cw_cur = cw_nxt; cw_nxt = bulk((wq >> 5) + 1); }
        const uint32_t tnn0 = __shfl_sync(FULL, cw_cur, wq & 31), tnn1 = __shfl_sync(FULL, cw_cur, (wq & 31) + 1);
        if constexpr (HAS_TOP) {
            if ((c0 & (BLK - 1)) == 0 && c0 < ncols) {
                const int c = c0 + lane;
                const bool mine = c < ncols;
                ulonglong2 e = pre;
                bool have = c0 != 0;
                unsigned spins = 0;
                for (;;) {
                    if (mine) SW_CHECK((unsigned long long)(top + c - a.bnd) < a.bnd_elems, SW_DEVERR_BND, a);
                    if (!have && mine) e = __ldcg(top + c);
                    const bool ok = !mine || ((uint32_t)(e.x >> 32) == tag_top && (uint32_t)(e.y >> 32) == tag_top);
                    if (__all_sync(FULL, ok)) break;
                    have = false;
                    __nanosleep(64);
                    if (++spins > a.spin_limit) {              // watchdog: never hang the GPU
                        if (head && a.dev_err) atomicOr(a.dev_err, SW_DEVERR_SPIN);
                        break;
                    }
                }
                const uint2 v = mine ? make_uint2((uint32_t)e.x, (uint32_t)e.y) : make_uint2(h0, gb2);
                __syncwarp();                                  // the head lane is done with the previous block
                s_top[lane] = v;
                __syncwarp();
            }
            if ((c0 & (BLK - 1)) == BLK - PF && c0 + PF < ncols) {
                const int c = c0 + PF + lane;
                if (c < ncols) pre = __ldcg(top + c);
            }
        }
        const uint2 *stp = s_top + (c0 & (BLK - 1));
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {
            uint32_t in_h[S][C], in_g[S][C], in_t[S];
#pragma unroll
            for (int j = 0; j < C; ++j) {
                in_h[0][j] = __shfl_up_sync(FULL, pub_h[S - 1][j], 1, G);
                in_g[0][j] = __shfl_up_sync(FULL, pub_g[S - 1][j], 1, G);
            }
            in_t[0] = __shfl_up_sync(FULL, pub_t[S - 1], 1, G);
            // head lane: the band above (or the matrix edge) and the codes of its NEXT step's columns
            const int nb = C * (uu + 1);                       // first byte of those codes inside this trip's 8
            uint32_t lead;
            if (nb >= TC) lead = tn0;
            else if (C == 4) lead = tc1;
            else lead = nb == 2 ? (tc0 >> 16) : nb == 4 ? tc1 : (tc1 >> 16);
#pragma unroll
            for (int j = 0; j < C; ++j) {
                uint2 b = make_uint2(h0, gb2);
                if constexpr (HAS_TOP) b = stp[uu * C + j];
                in_h[0][j] = head ? b.x : in_h[0][j];
                in_g[0][j] = head ? b.y : in_g[0][j];
            }
            in_t[0] = head ? (lead & CMASK) : in_t[0];
            // sub-strip s > 0 of the lane is fed by sub-strip s - 1 as it stood after the previous step
#pragma unroll
            for (int s = 1; s < S; ++s) {
                in_t[s] = pub_t[s - 1];
#pragma unroll
                for (int j = 0; j < C; ++j) { in_h[s][j] = pub_h[s - 1][j]; in_g[s][j] = pub_g[s - 1][j]; }
            }
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int j = 0; j < C; ++j) SW_CHECK(((in_t[s] >> (8 * j)) & 255u) <= (uint32_t)SRC::kPad, SW_DEVERR_PROF, a);
            uint2 sv_next[S][C][RP];
            load_sv(sv_next, in_t);
#pragma unroll
            for (int j = 0; j < C; ++j) {
                uint32_t hd[S], gt[S];
                uint2 svj[S][RP];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    hd[s] = j == 0 ? hd_carry[s] : in_h[s][j - 1];
                    gt[s] = in_g[s][j];
#pragma unroll
                    for (int k = 0; k < RP; ++k) svj[s][k] = sv[s][j][k];
                }
                column_step_multi<RS, S, G, AR, false>(H, Gl, best, hd, gt, svj, goe2, ge2, zero, 0u);
#pragma unroll
                for (int s = 0; s < S; ++s) { pub_h[s][j] = H[s][RS - 1]; pub_g[s][j] = Gl[s][RS - 1]; }
                if constexpr (HAS_BOTTOM && LS) {
                    if (lane == G - 1) s_bot[uu * C + j] = make_uint2(pub_h[S - 1][j], pub_g[S - 1][j]);
                } else {
                    sto_h[uu * C + j] = pub_h[S - 1][j];
                    sto_g[uu * C + j] = pub_g[S - 1][j];
                }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
                hd_carry[s] = in_h[s][C - 1];
                pub_t[s] = in_t[s];
#pragma unroll
                for (int j = 0; j < C; ++j)
#pragma unroll
                    for (int k = 0; k < RP; ++k) sv[s][j][k] = sv_next[s][j][k];
            }
        }
        tc0 = tn0; tc1 = tn1; tn0 = tnn0; tn1 = tnn1;
        if constexpr (HAS_BOTTOM) {
            // the last virtual PE finished columns cl0 .. cl0 + 7 in this trip; columns outside
            // 0 .. ncols - 1 land in the row's slack and are never read
            const int cl0 = c0 - C * (VPE - 1);
            if constexpr (LS) {
                // LS (latency-bound instances: one warp per scheduler): stored by TC lanes, one element
                // each, through shared memory.  A single lane storing TC elements reuses one register
                // quad for the 16-byte stores and waits on every store for the previous one to have
                // read it -- 9 % of a lone warp's time; with several warps per scheduler that wait is
                // hidden and the extra shared-memory round trip costs 1-7 % instead.
                __syncwarp();
                if (lane < TC && cl0 + (TC - 1) >= 0 && cl0 < ncols) {
                    const uint2 v = s_bot[lane];
                    SW_CHECK((unsigned long long)(bot + cl0 + lane - a.bnd) < a.bnd_elems, SW_DEVERR_BND, a);
                    __stcg(bot + cl0 + lane, make_ulonglong2(tag_hi | v.x, tag_hi | v.y));
                }
                __syncwarp();
            } else if (lane == G - 1 && cl0 + (TC - 1) >= 0 && cl0 < ncols) {
#pragma unroll
                for (int i = 0; i < TC; ++i) {
                    SW_CHECK((unsigned long long)(bot + cl0 + i - a.bnd) < a.bnd_elems, SW_DEVERR_BND, a);
                    __stcg(bot + cl0 + i, make_ulonglong2(tag_hi | sto_h[i], tag_hi | sto_g[i]));
                }
            }
        }
    }
    return best;
}

template <int RS, int S, class AR, int BT, int MINB, int CGOE = 0, int CGE = 0, int BLK = kWaveBlock, int C = 1, bool LS = false>
__global__ void __launch_bounds__(BT, MINB) sw_wave_kernel(const WaveArgs a)
{
    extern __shared__ uint2 s_prof[];
    __shared__ unsigned s_work;
    constexpr int G = 32, R = RS * S, P = R * G, RP = (RS + 1) / 2, VPE = G * S;
    constexpr int PASS_ENTRIES = VPE * RP * kWaveCodes;
    constexpr int PPB = BT / G;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    __shared__ __align__(16) uint2 s_top[PPB][BLK + kWaveBotSlots];
    __shared__ uint8_t s_qb[P / 4 + 4];                // packed query bytes of the band

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t zero = a.zero;
    const uint32_t goe2 = CGOE ? ((uint32_t)(CGOE & 0xFFFF) * 0x10001u) : a.goe2;
    const uint32_t ge2 = CGOE ? ((uint32_t)(CGE & 0xFFFF) * 0x10001u) : a.ge2;
    const uint32_t gb2 = 0u;               // clamped form: boundary gap value 0
    const uint32_t h0 = goe2;              // "H = 0" in the K = H + goe representation
    const int q = a.q;
    const int m = (int)a.qlen[q];
    const uint8_t *qp = a.qpacked + a.qoff[q];
    const unsigned nitems = (unsigned)a.npass * a.npb;

    int prof_pass = -1;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_work = atomicAdd(a.counter, 1u);
        __syncthreads();
        const unsigned work = s_work;
        if (work >= nitems) break;
        // band-major, longest pairs first inside a band
        const int pass = (int)(work / a.npb);
        const unsigned pb = a.npb - 1u - (work % a.npb);
        const unsigned pair = pb * PPB + (unsigned)warp;
        const bool valid = pair < a.npairs;
        const int ncols = valid ? (int)a.pair_len[2 * pair] : 0;
        const uint32_t *tpp = a.tp;
        if (valid) tpp += a.tile_woff[pair >> 5] + (pair & 31);
        if (valid && ncols > 0)
            SW_CHECK((unsigned long long)(tpp - a.tp) + (unsigned long long)((ncols - 1) >> 2) * 32 < a.tp_words, SW_DEVERR_TP, a);

        if (prof_pass != pass) {
            prof_pass = pass;
            __syncthreads();
            const int qb0 = (pass * P) >> 2;
            const int qnb = ((min((pass + 1) * P, m) + 3) >> 2) - qb0;
            for (int i = threadIdx.x; i < qnb; i += BT) s_qb[i] = qp[qb0 + i];
            __syncthreads();
            for (int idx = threadIdx.x; idx < PASS_ENTRIES; idx += BT) {
                // layout: ((s * RP + rp) * kWaveCodes + code) * G + lane
                const int lg = idx % G;
                const int code = (idx / G) % kWaveCodes;
                if (code > kPadCode) continue;
                const int rp = (idx / (G * kWaveCodes)) % RP;
                const int ss = (idx / (G * kWaveCodes * RP)) % S;
                const int vpe = lg * S + ss;
                uint32_t e[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int rr = 2 * rp + k;
                    const int i = pass * P + vpe * RS + rr;
                    int lo = AR::kPad, hi = AR::kPad;
                    if (rr < RS && i < m && code < kPadCode) {
                        const int qi = (s_qb[(i >> 2) - qb0] >> ((i & 3) * 2)) & 3;
                        lo = (qi == (code & 3)) ? a.match : a.mismatch;   // v1.0.v:119
                        if (code < kHiEndedCode) hi = (qi == (code >> 2)) ? a.match : a.mismatch;
                    }
                    e[k] = AR::pack_score(lo, hi);
                }
                s_prof[idx] = make_uint2(e[0], e[1]);
            }
            __syncthreads();
        }
        const uint2 *prof_lane = s_prof + lane;
        const bool has_top = pass > 0, has_bottom = pass + 1 < a.npass;
        const size_t prow = (size_t)pair * 2;
        const ulonglong2 *top = a.bnd + (prow + (size_t)((pass - 1) & 1)) * a.cols_stride + kWaveSlack;
        ulonglong2 *bot = a.bnd + (prow + (size_t)(pass & 1)) * a.cols_stride + kWaveSlack;
        const uint32_t tag_bot = (a.epoch << 12) | (uint32_t)pass, tag_top = tag_bot - 1u;

        uint32_t best;
#define SW_WAVE_BAND(T, B)                                                                                                   \
        do {                                                                                                                 \
            if constexpr (C > 1)                                                                                             \
                best = wave_band_c<RS, S, C, AR, BLK, T, B, LS>(a, prof_lane, s_top[warp], TiledCodes{tpp}, ncols, top, bot, tag_top, tag_bot, goe2, ge2, h0, gb2, zero); \
            else                                                                                                             \
                best = wave_band<RS, S, AR, BLK, T, B>(a, prof_lane, s_top[warp], tpp, ncols, top, bot, tag_top, tag_bot, goe2, ge2, h0, gb2, zero); \
        } while (0)
        if (has_top) { if (has_bottom) SW_WAVE_BAND(true, true); else SW_WAVE_BAND(true, false); }
        else { if (has_bottom) SW_WAVE_BAND(false, true); else SW_WAVE_BAND(false, false); }
#undef SW_WAVE_BAND

#pragma unroll
        for (int o = G / 2; o >= 1; o >>= 1) best = AR::max2(best, __shfl_xor_sync(FULL, best, o));
        if (lane == 0 && valid) {
            const int b0 = AR::extract(best, 0), b1 = AR::extract(best, 1);
            // a running maximum above the threshold means a 16-bit wrap may have happened in this band
            atomicMax(a.best + 2 * pair, b0 > a.ovf_limit ? 0x7FFFFFFF : b0 - a.goe);
            atomicMax(a.best + 2 * pair + 1, b1 > a.ovf_limit ? 0x7FFFFFFF : b1 - a.goe);
            __threadfence();
            if (atomicAdd(a.done + pair, 1u) == (unsigned)a.npass - 1u) {
                // last band of this pair: the maximum over all bands is final
                __threadfence();
                const int f0 = atomicMax(a.best + 2 * pair, 0), f1 = atomicMax(a.best + 2 * pair + 1, 0);
                const uint32_t subj_lo = a.pair_subj[2 * pair], subj_hi = a.pair_subj[2 * pair + 1];
                const bool ov0 = f0 == 0x7FFFFFFF, ov1 = f1 == 0x7FFFFFFF, has1 = subj_hi != SW_NO_SUBJECT;
                if (a.ovf_list) {
                    if (ov0) { const unsigned p = atomicAdd(a.ovf_count, 1u); if (p < a.ovf_cap) a.ovf_list[p] = make_uint2((unsigned)q, subj_lo); }
                    if (has1 && ov1) { const unsigned p = atomicAdd(a.ovf_count, 1u); if (p < a.ovf_cap) a.ovf_list[p] = make_uint2((unsigned)q, subj_hi); }
                }
                SW_CHECK((unsigned long long)a.out_row * a.out_stride + subj_lo < a.out_elems && (!has1 || (unsigned long long)a.out_row * a.out_stride + subj_hi < a.out_elems), SW_DEVERR_OUT, a);
                if (a.out_mode == SW_OUT_I16) {
                    int16_t *orow = (int16_t *)a.out + (size_t)a.out_row * a.out_stride;
                    orow[subj_lo] = (int16_t)(ov0 ? SW_OVERFLOW_SENTINEL : f0);
                    if (has1) orow[subj_hi] = (int16_t)(ov1 ? SW_OVERFLOW_SENTINEL : f1);
                } else {
                    int32_t *orow = (int32_t *)a.out + (size_t)a.out_row * a.out_stride;
                    orow[subj_lo] = ov0 ? SW_OVERFLOW_SENTINEL : f0;
                    if (has1) orow[subj_hi] = ov1 ? SW_OVERFLOW_SENTINEL : f1;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// 32-bit band-pipelined scorer of the overflow list.  The packed 16-bit kernels flag the (query,
// subject) entries whose score may leave the 16-bit range; with the default scoring those are
// pairs of at least 6 400 x 6 400 nt, and one thread of score32_kernel per entry takes seconds
// to minutes for them.  Here the bands (256 rows) of an entry are work items exactly as in
// sw_wave_kernel -- same band function, ArithS32 instead of ArithS16, codes straight from the
// subject's 2-bit record -- and the kernel is list-driven: the entry count is read on the device,
// so the launch stays asynchronous.
//   item      = (entry e, band b), index e * maxb + b, claimed in increasing order: the band an
//               item waits for was claimed earlier by a running block (no deadlock by construction)
//   boundary  = bnd[e % nslots][b & 1][column], tag = (epoch : 8, e : 12, b : 12); entry e may only
//               start when entry e - nslots has finished (flag[e - nslots], set by its last band, or
//               by band 0 of an entry this kernel leaves to score32_kernel)
//   entries   = kWave32MaxEntries list entries per launch (12 tag bits), a few launches per call:
//               the first entry_limit entries of the list that satisfy wave32_takes();
//               score32_kernel (mode 2) skips exactly those.
// ------------------------------------------------------------------------------------------
constexpr unsigned kWave32MaxEntries = 4096;
constexpr int kWave32Rows = 256;

__host__ __device__ inline bool wave32_takes(unsigned e, uint32_t m, uint32_t n, unsigned long long min_cells, unsigned limit)
{
    return e < limit && (unsigned long long)m * n >= min_cells && n > 0 &&
           (m + kWave32Rows - 1) / kWave32Rows <= 4095u;
}

struct Wave32Args {
    const uint8_t *raw;            // subjects: packed 2-bit records
    const uint64_t *off;
    const uint32_t *len;
    const uint8_t *qpacked;
    const uint32_t *qoff;
    const uint32_t *qlen;
    const unsigned *list_count;
    const uint2 *list;             // (query, subject)
    unsigned list_cap;
    unsigned entry_base, entry_limit;   // this launch: list entries entry_base .. entry_base + kWave32MaxEntries - 1; all launches: < entry_limit
    int32_t *list_score;
    void *out;                     // score matrix (may be null), as for score32_kernel
    size_t out_stride;
    int out_mode;
    ulonglong2 *bnd;               // [nslots][2][cols_stride]
    uint32_t cols_stride, nslots, epoch;
    int *best;                     // [kWave32MaxEntries], zeroed
    unsigned *done;                // [kWave32MaxEntries], zeroed: bands finished
    unsigned *flag;                // [kWave32MaxEntries], zeroed: entry finished (or not taken)
    unsigned *counter;
    uint32_t maxb;                 // bands of the longest query
    unsigned long long min_cells;
    int match, mismatch, goe, ge;
    unsigned *dev_err;
    unsigned spin_limit;
    uint64_t bnd_elems, out_elems;
};

template <int RS, int C, int MINB>
__global__ void __launch_bounds__(32, MINB) sw_wave32_kernel(const Wave32Args a)
{
    constexpr int G = 32, P = RS * G, RP = (RS + 1) / 2, CODES = RawCodes::kCodes, BLK = 32;
    static_assert(P == kWave32Rows, "band height");
    constexpr unsigned FULL = 0xFFFFFFFFu;
    __shared__ uint2 s_prof[G * RP * CODES];
    __shared__ __align__(16) uint2 s_top[BLK + kWaveBotSlots];
    __shared__ uint8_t s_qb[P / 4 + 4];
    const int lane = threadIdx.x;
    const unsigned total = min(*a.list_count, a.list_cap);
    const unsigned listed = total > a.entry_base ? min(total - a.entry_base, kWave32MaxEntries) : 0u;
    const uint32_t goe = (uint32_t)a.goe, ge = (uint32_t)a.ge, h0 = goe, gb = 0u;

    for (;;) {
        unsigned work = 0;
        if (lane == 0) work = atomicAdd(a.counter, 1u);
        work = __shfl_sync(FULL, work, 0);
        const unsigned e = work / a.maxb;
        const int b = (int)(work % a.maxb);
        if (e >= listed) break;
        const uint2 ent = a.list[a.entry_base + e];
        const int q = (int)ent.x;
        const uint32_t subj = ent.y;
        const int m = (int)a.qlen[q], n = (int)a.len[subj];
        const int npass = (m + P - 1) / P;
        const bool take = wave32_takes(a.entry_base + e, (uint32_t)m, (uint32_t)n, a.min_cells, a.entry_limit);
        if (e >= a.nslots) {
            // the boundary rows of this slot are free once entry e - nslots has finished (every item
            // of the entry waits here, so that the waits inside a band stay pipeline-short)
            const volatile unsigned *f = a.flag + (e - a.nslots);
            unsigned spins = 0;
            while (*f == 0u) {
                __nanosleep(256);
                if (++spins > (1u << 28)) {                    // ~ a minute: never hang the GPU
                    if (lane == 0 && a.dev_err) atomicOr(a.dev_err, SW_DEVERR_SPIN);
                    break;
                }
            }
            __threadfence();
        }
        if (b == 0 && !take) {
            // left to score32_kernel: the slot chain must not stop here
            if (lane == 0) { __threadfence(); atomicExch(a.flag + e, 1u); }
            continue;
        }
        if (!take || b >= npass) continue;

        // profile of the band: s_prof[(rp * CODES + code) * G + lane] = scores of rows 2 rp, 2 rp + 1
        __syncwarp();
        const uint8_t *qp = a.qpacked + a.qoff[q];
        const int qb0 = (b * P) >> 2;
        const int qnb = ((min((b + 1) * P, m) + 3) >> 2) - qb0;
        for (int i = lane; i < qnb; i += G) s_qb[i] = qp[qb0 + i];
        __syncwarp();
        for (int idx = lane; idx < G * RP * CODES; idx += G) {
            const int code = (idx / G) % CODES;
            const int rp = idx / (G * CODES);
            uint32_t sc[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int rr = 2 * rp + k;
                const int i = b * P + lane * RS + rr;
                int v = ArithS32::kPad;
                if (rr < RS && i < m && code < RawCodes::kPad) {
                    const int qi = (s_qb[(i >> 2) - qb0] >> ((i & 3) * 2)) & 3;
                    v = (qi == code) ? a.match : a.mismatch;       // v1.0.v:119
                }
                sc[k] = (uint32_t)v;
            }
            s_prof[idx] = make_uint2(sc[0], sc[1]);
        }
        __syncwarp();

        const bool has_top = b > 0, has_bottom = b + 1 < npass;
        const size_t slot = (size_t)(e % a.nslots) * 2;
        const ulonglong2 *top = a.bnd + (slot + (size_t)((b - 1) & 1)) * a.cols_stride + kWaveSlack;
        ulonglong2 *bot = a.bnd + (slot + (size_t)(b & 1)) * a.cols_stride + kWaveSlack;
        const uint32_t tag_bot = ((a.epoch & 255u) << 24) | (e << 12) | (uint32_t)b, tag_top = tag_bot - 1u;
        const RawCodes src{a.raw + a.off[subj]};
        const uint2 *prof_lane = s_prof + lane;
        uint32_t best;
#define SW_WAVE32_BAND(T, B) \
        best = wave_band_c<RS, 1, C, ArithS32, BLK, T, B, true>(a, prof_lane, s_top, src, n, top, bot, tag_top, tag_bot, goe, ge, h0, gb, 0u)
        if (has_top) { if (has_bottom) SW_WAVE32_BAND(true, true); else SW_WAVE32_BAND(true, false); }
        else { if (has_bottom) SW_WAVE32_BAND(false, true); else SW_WAVE32_BAND(false, false); }
#undef SW_WAVE32_BAND
#pragma unroll
        for (int o = G / 2; o >= 1; o >>= 1) best = ArithS32::max2(best, __shfl_xor_sync(FULL, best, o));
        if (lane == 0) {
            atomicMax(a.best + e, (int)best - a.goe);
            __threadfence();
            if (atomicAdd(a.done + e, 1u) == (unsigned)npass - 1u) {
                __threadfence();
                const int f = atomicMax(a.best + e, 0);
                if (a.list_score) a.list_score[a.entry_base + e] = f;
                if (a.out) {
                    SW_CHECK((unsigned long long)q * a.out_stride + subj < a.out_elems, SW_DEVERR_OUT, a);
                    // a 16-bit matrix keeps the sentinel for scores it cannot hold (they are on the list)
                    if (a.out_mode == SW_OUT_I16) { if (f <= 32767) ((int16_t *)a.out)[(size_t)q * a.out_stride + subj] = (int16_t)f; }
                    else ((int32_t *)a.out)[(size_t)q * a.out_stride + subj] = f;
                }
                __threadfence();
                atomicExch(a.flag + e, 1u);
            }
        }
    }
}

}  // namespace swk

#endif  /* SW_WAVE_CUH_ */
