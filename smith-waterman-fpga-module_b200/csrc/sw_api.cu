/*
 * sw_api.cu -- host side of libsw_b200.so: the C ABI declared in include/sw_b200.h.
 *
 * Replaces the reference's job path (main_test.c:290-477: build sequence records, attach the
 * AFU, poll the WED) and the bank-level dispatch (ScoreBank_v2.v:142-169 + PrioEncoder.v +
 * SM_Feeder2.v): subjects are length-sorted and paired on the host (the "first free module"
 * arbitration becomes a length-bucketed work queue drained by persistent blocks), sharded
 * over the handle's GPUs as contiguous input ranges, and moved with cudaMemcpyAsync from
 * pinned staging on per-GPU streams.  No CPU scoring path exists in this library.
 */
#include "../../include/sw_b200.h"
#include "sw_kernels.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

namespace {

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes ? bytes : 16;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct QueryChunk { int q0, q1; cudaEvent_t done; };

// Everything that belongs to one database batch on one GPU.  Two slots per GPU: while the
// kernels of batch k run, batch k+1 can be sorted, uploaded and enqueued, and batch k-1 copied
// back (the feeder's double buffering, SM_Feeder2.v:104-205, at batch granularity).
struct Slot {
    size_t s0 = 0, s1 = 0;            // global subject range [s0, s1) of this GPU's shard
    DevBuf d_raw, d_off, d_len, d_pair_subj, d_pair_len, d_tile_woff, d_tp, d_out;
    uint32_t npairs = 0, max_len = 0;
    uint64_t sum_len = 0;
    PinnedBuf h_stage_a, h_stage_b, h_stage_c, h_stage_d;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_upload = nullptr;
    std::vector<QueryChunk> chunks;
    bool scored = false;
};

struct GpuCtx {
    int dev = 0;
    int num_sms = 0;
    cudaStream_t st_compute = nullptr, st_copy = nullptr;
    // queries
    DevBuf d_qpacked, d_qoff, d_qlen;
    Slot slot[2];
    // scratch shared by both slots (kernels of one GPU run in stream order)
    DevBuf d_bnd, d_counters, d_scratch32, d_best_score, d_best_index;
};

// Host-side description of one batch (all GPUs).
struct Batch {
    size_t ns = 0;
    int nq = 0;                       // query rows the scores of this batch have
    bool loaded = false, scored = false;
    std::vector<uint64_t> ids;
    bool have_ids = false;
    uint64_t cells = 0;
};

}  // namespace

struct sw_handle {
    sw_params_t params;
    std::vector<GpuCtx> gpus;
    // host copy of the queries
    std::vector<uint8_t> q_packed;
    std::vector<uint32_t> q_off, q_len;
    uint32_t q_max_len = 0;
    uint64_t q_sum_len = 0;
    bool both_strands = false;
    int nq_user = 0;                  // queries given by the caller (rows = nq_user * strands)
    // database batches: slot 0 doubles as the resident database of sw_load_db
    Batch batch[2];
    int fifo[2] = {0, 0};             // slots in flight, oldest first
    int n_inflight = 0;
    int last_slot = 0;                // slot of the most recent load / fetch (ids, cells)
    // bookkeeping
    int last_cuda = 0;
    std::atomic<uint64_t> launches{0};
    uint64_t last_cells = 0;
    double last_ms = 0.0;
    const char *last_kernel = "none";
    int force_R = 0, force_G = 0, force32 = 0, force_arith = -1;
    int force_variant = -1;
    bool autotune = true;             // time the model's top candidates on a sample of large jobs
    uint64_t tune_key = 0;            // workload signature of the cached decision
    int tune_choice = -1;
};

namespace {

const int kMaxCounters = 4096;

#define SW_CUDA(h, call)                                                        \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) { (h)->last_cuda = (int)e__; return SW_ECUDA; } \
    } while (0)

int validate_params(const sw_params_t *p)
{
    if (p->match <= 0 || p->match > 255) return SW_EINVAL;
    if (p->mismatch > p->match || p->mismatch < -255) return SW_EINVAL;
    if (p->gap_extend > 0 || p->gap_extend < -255) return SW_EINVAL;
    const int goe = (int)p->gap_open + (int)p->gap_extend;
    if (goe > 0 || p->gap_open < -2040) return SW_EINVAL;
    if (p->score_width != 0) {
        if (p->score_width < 6 || p->score_width > 15) return SW_EINVAL;
        const int half = 1 << (p->score_width - 1);
        /* keep every intermediate of the W-bit machine inside its range except the
           M overflow the mode exists to reproduce */
        if (p->match >= half || goe + p->gap_extend < -half || p->mismatch < -half) return SW_EINVAL;
    }
    return SW_OK;
}

void free_gpu(GpuCtx &g)
{
    cudaSetDevice(g.dev);
    for (Slot &b : g.slot) {
        for (auto &c : b.chunks) if (c.done) cudaEventDestroy(c.done);
        b.chunks.clear();
        DevBuf *bufs[] = {&b.d_raw, &b.d_off, &b.d_len, &b.d_pair_subj, &b.d_pair_len, &b.d_tile_woff, &b.d_tp, &b.d_out};
        for (DevBuf *d : bufs) d->release();
        b.h_stage_a.release(); b.h_stage_b.release(); b.h_stage_c.release(); b.h_stage_d.release();
        if (b.ev_start) cudaEventDestroy(b.ev_start);
        if (b.ev_stop) cudaEventDestroy(b.ev_stop);
        if (b.ev_upload) cudaEventDestroy(b.ev_upload);
    }
    DevBuf *bufs[] = {&g.d_qpacked, &g.d_qoff, &g.d_qlen, &g.d_bnd, &g.d_counters, &g.d_scratch32,
                      &g.d_best_score, &g.d_best_index};
    for (DevBuf *d : bufs) d->release();
    if (g.st_compute) cudaStreamDestroy(g.st_compute);
    if (g.st_copy) cudaStreamDestroy(g.st_copy);
}

int upload_queries(sw_handle *h, GpuCtx &g)
{
    SW_CUDA(h, cudaSetDevice(g.dev));
    const size_t nq = h->q_len.size();
    SW_CUDA(h, g.d_qpacked.reserve(h->q_packed.size() + 16));
    SW_CUDA(h, g.d_qoff.reserve(nq * sizeof(uint32_t)));
    SW_CUDA(h, g.d_qlen.reserve(nq * sizeof(uint32_t)));
    // the compute stream orders these copies before any later kernel
    SW_CUDA(h, cudaMemcpyAsync(g.d_qpacked.p, h->q_packed.data(), h->q_packed.size(), cudaMemcpyHostToDevice, g.st_compute));
    SW_CUDA(h, cudaMemcpyAsync(g.d_qoff.p, h->q_off.data(), nq * sizeof(uint32_t), cudaMemcpyHostToDevice, g.st_compute));
    SW_CUDA(h, cudaMemcpyAsync(g.d_qlen.p, h->q_len.data(), nq * sizeof(uint32_t), cudaMemcpyHostToDevice, g.st_compute));
    SW_CUDA(h, cudaStreamSynchronize(g.st_compute));
    return SW_OK;
}

// Length-sorts the shard's subjects, pairs neighbours in length order, lays out 32-pair tiles, uploads.
int load_shard(sw_handle *h, GpuCtx &gc, Slot &g, const uint8_t *packed, const uint32_t *len, const uint64_t *off)
{
    SW_CUDA(h, cudaSetDevice(gc.dev));
    const size_t n = g.s1 - g.s0;
    g.npairs = 0; g.max_len = 0; g.sum_len = 0; g.scored = false;
    if (n == 0) return SW_OK;
    if (n >= 0xFFFFFFF0ull) return SW_EINVAL;
    const uint32_t *ln = len + g.s0;
    const uint64_t *of = off + g.s0;

    uint64_t bmin = ~0ull, bmax = 0;
    uint32_t maxlen = 0;
    uint64_t sum = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint64_t b = (ln[i] + 3ull) >> 2;
        if (ln[i] == 0) continue;
        bmin = std::min(bmin, of[i]);
        bmax = std::max(bmax, of[i] + b);
        maxlen = std::max(maxlen, ln[i]);
        sum += ln[i];
    }
    g.max_len = maxlen; g.sum_len = sum;
    if (maxlen == 0) return SW_OK;

    // --- order by length (counting sort when the range is small, else comparison sort)
    SW_CUDA(h, g.h_stage_a.reserve((n + 64) * 2 * sizeof(uint32_t)));      // pair_subj
    SW_CUDA(h, g.h_stage_b.reserve((n + 128) * sizeof(uint32_t)));         // pair_len (2 per pair)
    std::vector<uint32_t> order;
    order.reserve(n);
    if (maxlen <= (1u << 22)) {
        std::vector<uint32_t> start((size_t)maxlen + 2, 0);
        for (size_t i = 0; i < n; ++i) start[ln[i] + 1]++;
        for (size_t l = 1; l < start.size(); ++l) start[l] += start[l - 1];
        order.resize(n);
        for (size_t i = 0; i < n; ++i) order[start[ln[i]]++] = (uint32_t)i;
    } else {
        order.resize(n);
        for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ln[a] < ln[b]; });
    }
    // --- pair neighbours in length order: the longer member drives the column loop (low lane),
    //     the shorter one (high lane) sees PAD columns once it has ended
    uint32_t *pair_subj = (uint32_t *)g.h_stage_a.p;
    uint32_t *pair_len = (uint32_t *)g.h_stage_b.p;
    size_t np = 0, i = 0;
    while (i < n && ln[order[i]] == 0) ++i;           // empty subjects score 0 (output is pre-zeroed)
    if ((n - i) & 1) {                                // odd count: the shortest one stays single
        pair_subj[0] = order[i]; pair_subj[1] = SW_NO_SUBJECT;
        pair_len[0] = ln[order[i]]; pair_len[1] = 0;
        ++i; ++np;
    }
    for (; i + 1 < n; i += 2, ++np) {
        const uint32_t shorter = order[i], longer = order[i + 1];
        pair_subj[2 * np] = longer; pair_subj[2 * np + 1] = shorter;
        pair_len[2 * np] = ln[longer]; pair_len[2 * np + 1] = ln[shorter];
    }
    const size_t ntiles = (np + 31) / 32;
    for (size_t p = np; p < ntiles * 32; ++p) {
        pair_subj[2 * p] = SW_NO_SUBJECT; pair_subj[2 * p + 1] = SW_NO_SUBJECT;
        pair_len[2 * p] = 0; pair_len[2 * p + 1] = 0;
    }
    SW_CUDA(h, g.h_stage_c.reserve((ntiles + 1) * sizeof(uint64_t)));
    uint64_t *tile_woff = (uint64_t *)g.h_stage_c.p;
    uint64_t w = 0;
    for (size_t t = 0; t < ntiles; ++t) {
        tile_woff[t] = w;
        const size_t last = std::min(np, (t + 1) * 32) - 1;     // lengths ascend: last valid pair is longest
        w += 32ull * ((pair_len[2 * last] + 3) / 4);        // one byte per column, 4 per word
    }
    tile_woff[ntiles] = w;
    g.npairs = (uint32_t)np;

    // --- local byte offsets
    SW_CUDA(h, g.h_stage_d.reserve(n * sizeof(uint64_t)));
    uint64_t *loc_off = (uint64_t *)g.h_stage_d.p;
    for (size_t k = 0; k < n; ++k) loc_off[k] = ln[k] ? of[k] - bmin : 0;

    // --- device buffers + uploads (copy stream), then the code-stream build (compute stream)
    const size_t raw_bytes = bmax - bmin;
    SW_CUDA(h, g.d_raw.reserve(raw_bytes + 16));
    SW_CUDA(h, g.d_off.reserve(n * sizeof(uint64_t)));
    SW_CUDA(h, g.d_len.reserve(n * sizeof(uint32_t)));
    SW_CUDA(h, g.d_pair_subj.reserve(ntiles * 32 * 2 * sizeof(uint32_t)));
    SW_CUDA(h, g.d_pair_len.reserve(ntiles * 32 * 2 * sizeof(uint32_t)));
    SW_CUDA(h, g.d_tile_woff.reserve((ntiles + 1) * sizeof(uint64_t)));
    SW_CUDA(h, g.d_tp.reserve((w + 32) * sizeof(uint32_t)));
    cudaStream_t cs = gc.st_copy;
    SW_CUDA(h, cudaMemcpyAsync(g.d_raw.p, packed + bmin, raw_bytes, cudaMemcpyHostToDevice, cs));
    SW_CUDA(h, cudaMemcpyAsync(g.d_len.p, ln, n * sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
    SW_CUDA(h, cudaMemcpyAsync(g.d_off.p, loc_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, cs));
    SW_CUDA(h, cudaMemcpyAsync(g.d_pair_subj.p, pair_subj, ntiles * 32 * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
    SW_CUDA(h, cudaMemcpyAsync(g.d_pair_len.p, pair_len, ntiles * 32 * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
    SW_CUDA(h, cudaMemcpyAsync(g.d_tile_woff.p, tile_woff, (ntiles + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, cs));
    SW_CUDA(h, cudaEventRecord(g.ev_upload, cs));
    SW_CUDA(h, cudaStreamWaitEvent(gc.st_compute, g.ev_upload, 0));

    SwDevDb db;
    db.raw = g.d_raw.as<uint8_t>(); db.off = g.d_off.as<uint64_t>(); db.len = g.d_len.as<uint32_t>();
    db.ns = (uint32_t)n; db.pair_subj = g.d_pair_subj.as<uint32_t>(); db.pair_len = g.d_pair_len.as<uint32_t>();
    db.tile_woff = g.d_tile_woff.as<uint64_t>(); db.tp = g.d_tp.as<uint32_t>();
    db.npairs = g.npairs; db.max_len = g.max_len;
    SW_CUDA(h, sw_launch_build_tp(gc.st_compute, db));
    h->launches++;
    return SW_OK;
}

SwDevDb dev_db(const Slot &g)
{
    SwDevDb db;
    db.raw = g.d_raw.as<uint8_t>(); db.off = g.d_off.as<uint64_t>(); db.len = g.d_len.as<uint32_t>();
    db.ns = (uint32_t)(g.s1 - g.s0); db.pair_subj = g.d_pair_subj.as<uint32_t>();
    db.pair_len = g.d_pair_len.as<uint32_t>(); db.tile_woff = g.d_tile_woff.as<uint64_t>();
    db.tp = g.d_tp.as<uint32_t>(); db.npairs = g.npairs; db.max_len = g.max_len;
    return db;
}

// Measured steady-state speed of each variant in GCUPS over PADDED cells (150-nt reads,
// profiles/r01_variant_sweep.txt); only the ratios matter for the choice below.
double variant_speed(const SwStripVariant *v)
{
    struct { const char *name; double gcups; } tab[] = {
        {"strip_s16x2_R30x1_G1", 8220}, {"strip_s16x2_R38x1_G1", 8310}, {"strip_s16x2_R75x1_G1", 8040},
        {"strip_s16x2_R32x1_G1", 8230}, {"strip_s16x2_R50x1_G1", 8180}, {"strip_s16x2_R25x2_G1", 8700},
        {"strip_s16x2_R19x2_G1", 7810}, {"strip_s16x2_R15x3_G1", 7590}, {"strip_s16x2_R30x2_G1", 8040},
        {"strip_s16x2_R64x1_G1", 7910}, {"strip_s16x2_R32x2_G1", 8000}, {"strip_s16x2_R25x3_G1", 8620},
        {"strip_s16x2_R38x2_G1", 8600}, {"strip_s16x2_R25x4_G1", 7200}, {"strip_s16x2_R25x1_G2", 7345},
        {"strip_s16x2_R75x1_G2", 7380}, {"strip_s16x2_R25x3_G2", 7480}, {"strip_s16x2_R38x1_G4", 7325},
        {"strip_s16x2_R19x2_G4", 7220}, {"strip_s16x2_R32x1_G4", 7180}, {"strip_s16x2_R16x1_G32", 5980},
        {"strip_s16x2_R8x2_G32", 6320},
    };
    for (auto &t : tab) if (std::strcmp(t.name, v->name) == 0) return t.gcups;
    return 5000.0;
}

// Picks the strip variant: least estimated time = padded rows x (columns + pipeline fill)
// / measured speed / fraction of the GPU the pairs can keep busy.
int choose_variant(const sw_handle *h, const GpuCtx &gc, const Slot &g, uint32_t maxq, std::vector<int> *ranked = nullptr)
{
    const int nv = sw_strip_variant_count();
    if (h->force_variant >= 0) return h->force_variant;
    if (h->force_R || h->force_G) {
        for (int i = 0; i < nv; ++i) {
            const SwStripVariant *v = sw_strip_variant(i);
            if ((!h->force_R || v->R == h->force_R) && (!h->force_G || v->G == h->force_G))
                return i;
        }
        return -1;
    }
    int best = -1;
    double best_cost = 0;
    std::vector<std::pair<double, int>> costs;
    for (int i = 0; i < nv; ++i) {
        const SwStripVariant *v = sw_strip_variant(i);
        const int P = v->R * v->G;
        double rows = 0;                      // padded rows over all queries
        for (uint32_t ql : h->q_len) rows += (double)((ql + P - 1) / P) * P;
        if (rows == 0) rows = (double)((maxq + P - 1) / P) * P;
        const double lanes = (double)g.npairs * v->G;
        const double fill = (double)gc.num_sms * v->min_blocks * v->block_threads;
        const double util = std::min(1.0, lanes / fill);
        const double cols = (double)std::max<uint32_t>(g.max_len, 1);
        // time of the whole job at the variant's measured speed ...
        const double job = rows * (cols + v->G * v->S - 1) / cols * (double)g.sum_len / variant_speed(v) / util;
        // ... plus half the time of the longest work item (block of pairs x longest query) running
        // at its share of an SM: with few, long items the tail of the launch is what matters, and
        // variants with fewer resident blocks per SM finish a single item sooner
        const int ppb = v->block_threads / v->G;
        const double qrows = (double)((maxq + P - 1) / P) * P;
        const double item = qrows * cols * 2.0 * ppb * ((double)gc.num_sms * v->min_blocks) / variant_speed(v);
        const double cost = job + 0.5 * item;
        if (best < 0 || cost < best_cost) { best = i; best_cost = cost; }
        costs.emplace_back(cost, i);
    }
    if (ranked) {
        std::sort(costs.begin(), costs.end());
        for (auto &c : costs) ranked->push_back(c.second);
    }
    return best;
}

// Launch geometry of a strip variant for this shard and query set (also grows the boundary scratch).
int strip_setup(sw_handle *h, GpuCtx &gc, const Slot &g, int vidx, int *grid, int *chunk_passes)
{
    const SwStripVariant *v = sw_strip_variant(vidx);
    const int P = v->R * v->G;
    const int need_passes = (int)((h->q_max_len + P - 1) / P);
    const size_t pass_bytes = sw_strip_smem_bytes(vidx, 1);
    const int budget_passes = std::max<int>(1, (int)((48 * 1024) / pass_bytes));
    *chunk_passes = std::max(1, std::min(need_passes, budget_passes));
    int bps = 0;
    SW_CUDA(h, sw_strip_occupancy(vidx, sw_strip_smem_bytes(vidx, *chunk_passes), &bps));
    if (bps < 1) return SW_ECUDA;
    const int ppb = v->block_threads / v->G;
    const uint32_t npb = (g.npairs + ppb - 1) / ppb;
    *grid = (int)std::min<uint64_t>(npb, (uint64_t)gc.num_sms * bps);
    if (need_passes > 1) {
        // pass-boundary scratch: one (H, G) per column of the longest subject, per pair slot, per
        // resident block.  A very long subject would make that huge, so the grid shrinks to keep
        // the scratch within a budget (correct, slower; splitting the launch by length group is
        // the better answer and is left for later).
        const size_t per_block = (size_t)g.max_len * ppb * sizeof(uint2);
        const size_t budget = (size_t)16 << 30;
        if (per_block > budget) return SW_ENOMEM;
        *grid = (int)std::min<size_t>((size_t)*grid, std::max<size_t>(1, budget / per_block));
        SW_CUDA(h, gc.d_bnd.reserve((size_t)*grid * per_block));
    }
    return SW_OK;
}

// Times the model's best candidates on a window of the pair list (middle of the length order)
// and a few queries, and returns the fastest.  Every variant produces identical scores, so the
// sample launches may write into the real output buffer.
int autotune_variant(sw_handle *h, GpuCtx &gc, Slot &g, const SwDevDb &db, const SwDevQueries &dq,
                     const SwScoring &sc, const std::vector<int> &ranked, int nq)
{
    const int ncand = std::min<int>(3, (int)ranked.size());
    const int nqs = std::min(nq, 8);
    uint64_t qrows = 0;
    for (int q = 0; q < nqs; ++q) qrows += h->q_len[q];
    const double mean_len = (double)g.sum_len / (double)std::max<size_t>(g.s1 - g.s0, 1);
    const double cells_per_pair = 2.0 * mean_len * (double)std::max<uint64_t>(qrows, 1);
    uint64_t pairs_s = (uint64_t)(1.5e11 / cells_per_pair);                      // ~20 ms of work
    pairs_s = std::max<uint64_t>(pairs_s, (uint64_t)gc.num_sms * 4 * 128 * 4);   // >= 4 waves of blocks
    pairs_s = std::min<uint64_t>(pairs_s, g.npairs) & ~31ull;
    if (pairs_s < 1024) return ranked[0];
    const uint64_t p0 = ((g.npairs - pairs_s) / 2) & ~31ull;
    SwDevDb win = db;
    win.pair_subj = db.pair_subj + 2 * p0;
    win.pair_len = db.pair_len + 2 * p0;
    win.tile_woff = db.tile_woff + p0 / 32;
    win.npairs = (uint32_t)pairs_s;
    cudaEvent_t e0, e1;
    SW_CUDA(h, cudaEventCreate(&e0));
    SW_CUDA(h, cudaEventCreate(&e1));
    int best = ranked[0];
    float best_ms = 0.f;
    for (int c = 0; c < ncand; ++c) {
        const int vidx = ranked[c];
        int grid = 0, chunk_passes = 1;
        Slot tmp_geom;                       // geometry of the window (npairs / max_len only)
        tmp_geom.npairs = win.npairs; tmp_geom.max_len = g.max_len;
        int rc = strip_setup(h, gc, tmp_geom, vidx, &grid, &chunk_passes);
        if (rc != SW_OK) { cudaEventDestroy(e0); cudaEventDestroy(e1); return ranked[0]; }
        float ms = 0.f;
        for (int rep = 0; rep < 2; ++rep) {  // first run warms caches and the instruction cache
            SW_CUDA(h, cudaMemsetAsync(gc.d_counters.p, 0, sizeof(unsigned), gc.st_compute));
            SW_CUDA(h, cudaEventRecord(e0, gc.st_compute));
            SW_CUDA(h, sw_launch_strip(vidx, gc.st_compute, win, dq, 0, nqs, sc, g.d_out.as<int32_t>(), g.s1 - g.s0,
                                       gc.d_bnd.as<uint2>(), g.max_len, gc.d_counters.as<unsigned>(), grid, chunk_passes));
            h->launches++;
            SW_CUDA(h, cudaEventRecord(e1, gc.st_compute));
            SW_CUDA(h, cudaEventSynchronize(e1));
            SW_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        }
        if (c == 0 || ms < best_ms) { best = vidx; best_ms = ms; }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

int score_gpu(sw_handle *h, GpuCtx &gc, Slot &g)
{
    SW_CUDA(h, cudaSetDevice(gc.dev));
    const int nq = (int)h->q_len.size();
    const size_t n = g.s1 - g.s0;
    for (auto &c : g.chunks) if (c.done) cudaEventDestroy(c.done);
    g.chunks.clear();
    g.scored = true;
    if (n == 0 || nq == 0) return SW_OK;

    SW_CUDA(h, g.d_out.reserve((size_t)nq * n * sizeof(int32_t)));
    SW_CUDA(h, cudaMemsetAsync(g.d_out.p, 0, (size_t)nq * n * sizeof(int32_t), gc.st_compute));
    SW_CUDA(h, cudaEventRecord(g.ev_start, gc.st_compute));
    if (g.npairs == 0) {
        SW_CUDA(h, cudaEventRecord(g.ev_stop, gc.st_compute));
        QueryChunk c{0, nq, nullptr};
        SW_CUDA(h, cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
        SW_CUDA(h, cudaEventRecord(c.done, gc.st_compute));
        g.chunks.push_back(c);
        return SW_OK;
    }

    SwScoring sc;
    sc.match = h->params.match; sc.mismatch = h->params.mismatch;
    sc.goe = (int)h->params.gap_open + (int)h->params.gap_extend; sc.ge = h->params.gap_extend;
    sc.limit = h->params.score_width ? (1 << (h->params.score_width - 1)) - 1 : 0;

    SwDevDb db = dev_db(g);
    SwDevQueries dq;
    dq.packed = gc.d_qpacked.as<uint8_t>(); dq.off = gc.d_qoff.as<uint32_t>(); dq.len = gc.d_qlen.as<uint32_t>();
    dq.nq = nq; dq.max_len = h->q_max_len;

    // Value range: the packed 16-bit kernel is always used; when match * min(m, n) could exceed
    // the 16-bit range it flags the (rare) pairs whose running maximum got near 32767 and the
    // 32-bit kernel recomputes exactly those.
    const uint64_t smax = (uint64_t)sc.match * std::min<uint64_t>(h->q_max_len, g.max_len);
    const bool may_overflow = !sc.limit && (smax + (uint64_t)sc.match >= 32000ull);
    SW_CUDA(h, gc.d_counters.reserve(kMaxCounters * sizeof(unsigned)));
    int vidx = -1;
    std::vector<int> ranked;
    if (!h->force32) {
        vidx = choose_variant(h, gc, g, h->q_max_len, &ranked);
        if (vidx < 0 && (h->force_R || h->force_G)) return SW_EINVAL;
    }

    // query chunks: a handful of launches so that D2H of finished rows overlaps compute
    // (each launch has its own tail: aim for >= ~50 ms of work per launch, at ~6 TCUPS)
    const double est_ms = (double)g.sum_len * (double)h->q_sum_len / 6.0e9;
    int nchunks = std::max(1, std::min(std::min(nq, 8), (int)(est_ms / 50.0)));
    // the device work counter is 32-bit: (pair blocks) x (queries per launch) must stay below 2^31
    while (nchunks < nq && (uint64_t)(g.npairs / 4 + 1) * (uint64_t)((nq + nchunks - 1) / nchunks) >= (1ull << 31)) nchunks *= 2;
    nchunks = std::min(nchunks, nq);

    // large jobs: let the GPU pick among the model's top candidates (decision cached per workload shape)
    if (vidx >= 0 && h->autotune && h->force_variant < 0 && !h->force_R && !h->force_G && est_ms >= 400.0 && ranked.size() > 1) {
        uint64_t key = 1469598103934665603ull;
        auto log2b = [](uint64_t x) { uint64_t b = 0; while (x >>= 1) ++b; return b; };
        const uint64_t parts[] = {h->q_max_len, h->q_sum_len, (uint64_t)nq, g.max_len, log2b(g.npairs), log2b(g.sum_len),
                                  (uint64_t)(uint16_t)h->params.match, (uint64_t)(uint16_t)h->params.mismatch,
                                  (uint64_t)(uint16_t)h->params.gap_open, (uint64_t)(uint16_t)h->params.gap_extend,
                                  (uint64_t)h->params.score_width};
        for (uint64_t x : parts) { key ^= x; key *= 1099511628211ull; }
        if (h->tune_key != key || h->tune_choice < 0) {
            h->tune_choice = autotune_variant(h, gc, g, db, dq, sc, ranked, nq);
            h->tune_key = key;
        }
        vidx = h->tune_choice;
    }

    SW_CUDA(h, cudaMemsetAsync(gc.d_counters.p, 0, kMaxCounters * sizeof(unsigned), gc.st_compute));
    int grid = 0, chunk_passes = 1;
    if (vidx >= 0) {
        int rc = strip_setup(h, gc, g, vidx, &grid, &chunk_passes);
        if (rc != SW_OK) return rc;
        h->last_kernel = sw_strip_variant(vidx)->name;
    }
    if (vidx < 0 || may_overflow) {
        const int threads_total = gc.num_sms * 2 * 128;
        SW_CUDA(h, gc.d_scratch32.reserve((size_t)2 * std::max<uint32_t>(g.max_len, 1) * threads_total * sizeof(int32_t)));
        if (vidx < 0) h->last_kernel = "generic32";
    }

    for (int c = 0; c < nchunks; ++c) {
        QueryChunk qc;
        qc.q0 = (int)((long long)nq * c / nchunks);
        qc.q1 = (int)((long long)nq * (c + 1) / nchunks);
        qc.done = nullptr;
        if (qc.q1 <= qc.q0) continue;
        if (vidx >= 0) {
            SW_CUDA(h, sw_launch_strip(vidx, gc.st_compute, db, dq, qc.q0, qc.q1, sc, g.d_out.as<int32_t>(), n,
                                       gc.d_bnd.as<uint2>(), g.max_len, gc.d_counters.as<unsigned>() + (c % kMaxCounters),
                                       grid, chunk_passes));
            if (may_overflow) {
                SW_CUDA(h, sw_launch_generic32(gc.st_compute, db, dq, qc.q0, qc.q1, sc, g.d_out.as<int32_t>(), n,
                                               gc.d_scratch32.as<int32_t>(), gc.num_sms * 2 * 128, true));
                h->launches++;
            }
        } else {
            SW_CUDA(h, sw_launch_generic32(gc.st_compute, db, dq, qc.q0, qc.q1, sc, g.d_out.as<int32_t>(), n,
                                           gc.d_scratch32.as<int32_t>(), gc.num_sms * 2 * 128, false));
        }
        h->launches++;
        SW_CUDA(h, cudaEventCreateWithFlags(&qc.done, cudaEventDisableTiming));
        SW_CUDA(h, cudaEventRecord(qc.done, gc.st_compute));
        g.chunks.push_back(qc);
    }
    SW_CUDA(h, cudaEventRecord(g.ev_stop, gc.st_compute));
    return SW_OK;
}

// waits for an event with a deadline; deadline_ms < 0 = forever
int wait_event(sw_handle *h, cudaEvent_t ev, const std::chrono::steady_clock::time_point &t_end, bool forever)
{
    if (forever) { SW_CUDA(h, cudaEventSynchronize(ev)); return SW_OK; }
    for (;;) {
        cudaError_t e = cudaEventQuery(ev);
        if (e == cudaSuccess) return SW_OK;
        if (e != cudaErrorNotReady) { h->last_cuda = (int)e; return SW_ECUDA; }
        if (std::chrono::steady_clock::now() >= t_end) return SW_ETIMEOUT;
        std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
}

int copy_out(sw_handle *h, int si, int32_t *scores, size_t cap, int timeout_ms)
{
    const Batch &bt = h->batch[si];
    const size_t nq = (size_t)bt.nq;
    if (cap < nq * bt.ns) return SW_ECAPACITY;
    const bool forever = timeout_ms < 0;
    const auto t_end = std::chrono::steady_clock::now() + std::chrono::milliseconds(forever ? 0 : timeout_ms);
    size_t maxchunks = 0;
    for (auto &g : h->gpus) maxchunks = std::max(maxchunks, g.slot[si].chunks.size());
    for (size_t c = 0; c < maxchunks; ++c) {
        for (auto &g : h->gpus) {
            Slot &b = g.slot[si];
            if (c >= b.chunks.size()) continue;
            const size_t n = b.s1 - b.s0;
            if (n == 0) continue;
            SW_CUDA(h, cudaSetDevice(g.dev));
            const QueryChunk &qc = b.chunks[c];
            int rc = wait_event(h, qc.done, t_end, forever);
            if (rc != SW_OK) {
                // copies already enqueued must not outlive this call: the caller owns `scores`
                for (auto &gg : h->gpus) { cudaSetDevice(gg.dev); cudaStreamSynchronize(gg.st_copy); }
                return rc;
            }
            SW_CUDA(h, cudaMemcpy2DAsync(scores + (size_t)qc.q0 * bt.ns + b.s0, bt.ns * sizeof(int32_t),
                                         b.d_out.as<int32_t>() + (size_t)qc.q0 * n, n * sizeof(int32_t),
                                         n * sizeof(int32_t), (size_t)(qc.q1 - qc.q0), cudaMemcpyDeviceToHost,
                                         g.st_copy));
        }
    }
    double ms_max = 0.0;
    for (auto &g : h->gpus) {
        Slot &b = g.slot[si];
        SW_CUDA(h, cudaSetDevice(g.dev));
        SW_CUDA(h, cudaStreamSynchronize(g.st_copy));
        if (b.scored && (b.s1 > b.s0) && nq) {
            SW_CUDA(h, cudaEventSynchronize(b.ev_stop));
            float ms = 0.f;
            SW_CUDA(h, cudaEventElapsedTime(&ms, b.ev_start, b.ev_stop));
            ms_max = std::max(ms_max, (double)ms);
        }
    }
    h->last_ms = ms_max;
    h->last_cells = bt.cells;
    h->last_slot = si;
    return SW_OK;
}

int load_batch(sw_handle *h, int si, const uint8_t *packed, const uint32_t *len, const uint64_t *off,
               const uint64_t *ids, size_t ns)
{
    Batch &bt = h->batch[si];
    bt.ns = ns; bt.loaded = false; bt.scored = false; bt.cells = 0;
    bt.have_ids = ids != nullptr;
    if (ids) bt.ids.assign(ids, ids + ns); else bt.ids.clear();
    const size_t ng = h->gpus.size();
    std::vector<uint64_t> starts(ng + 1);
    sw_plan_shards(len, ns, (int)ng, starts.data());
    for (size_t gi = 0; gi < ng; ++gi) { h->gpus[gi].slot[si].s0 = starts[gi]; h->gpus[gi].slot[si].s1 = starts[gi + 1]; }
    // one host worker per GPU: the shards' length sort / pairing / uploads run concurrently
    if (ng == 1) {
        int rc = load_shard(h, h->gpus[0], h->gpus[0].slot[si], packed, len, off);
        if (rc != SW_OK) return rc;
    } else {
        std::vector<int> rcs(ng, SW_OK);
        std::vector<std::thread> workers;
        for (size_t gi = 0; gi < ng; ++gi)
            workers.emplace_back([&, gi]() { rcs[gi] = load_shard(h, h->gpus[gi], h->gpus[gi].slot[si], packed, len, off); });
        for (auto &w : workers) w.join();
        for (int rc : rcs) if (rc != SW_OK) return rc;
    }
    // caller's buffers must be reusable on return
    for (auto &g : h->gpus) {
        SW_CUDA(h, cudaSetDevice(g.dev));
        SW_CUDA(h, cudaStreamSynchronize(g.st_copy));
    }
    bt.loaded = true;
    h->last_slot = si;
    return SW_OK;
}

int score_batch_slot(sw_handle *h, int si)
{
    Batch &bt = h->batch[si];
    uint64_t db_len = 0;
    for (auto &g : h->gpus) db_len += g.slot[si].sum_len;
    bt.cells = db_len * h->q_sum_len;
    bt.nq = (int)h->q_len.size();
    h->last_cells = bt.cells;
    for (auto &g : h->gpus) {
        int rc = score_gpu(h, g, g.slot[si]);
        if (rc != SW_OK) return rc;
    }
    bt.scored = true;
    return SW_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

void sw_default_params(sw_params_t *p)
{
    if (!p) return;
    p->match = 5; p->mismatch = -4; p->gap_open = -12; p->gap_extend = -4; p->score_width = 0;
}

const char *sw_version(void) { return "sw_b200 0.1 (sm_100a)"; }

int sw_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *sw_strerror(int code)
{
    switch (code) {
        case SW_OK: return "ok";
        case SW_EINVAL: return "invalid argument or unsupported parameter set";
        case SW_ENOMEM: return "out of memory";
        case SW_ECUDA: return "CUDA error (see sw_last_cuda_error)";
        case SW_ENODEV: return "no usable CUDA device";
        case SW_ESTATE: return "call out of order";
        case SW_ETIMEOUT: return "timed out";
        case SW_ECAPACITY: return "output buffer too small";
        case SW_EIO: return "I/O error";
        case SW_EAGAIN: return "busy: both batch buffers are in flight (fetch one first)";
        default: return "unknown error";
    }
}

int sw_init(sw_handle_t **out, const sw_params_t *p, const int *gpu_ids, int n_gpus)
{
    if (!out) return SW_EINVAL;
    *out = nullptr;
    sw_params_t prm;
    if (p) prm = *p; else sw_default_params(&prm);
    int rc = validate_params(&prm);
    if (rc != SW_OK) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return SW_ENODEV; }
    std::vector<int> ids;
    if (!gpu_ids || n_gpus <= 0) ids.push_back(0);
    else ids.assign(gpu_ids, gpu_ids + n_gpus);
    for (int id : ids) if (id < 0 || id >= ndev) return SW_ENODEV;

    sw_handle *h = new (std::nothrow) sw_handle();
    if (!h) return SW_ENOMEM;
    h->params = prm;
    if (const char *e = std::getenv("SW_B200_AUTOTUNE")) h->autotune = (e[0] != '0');
    h->gpus.resize(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) {
        GpuCtx &g = h->gpus[i];
        g.dev = ids[i];
        cudaError_t e = cudaSetDevice(g.dev);
        cudaDeviceProp prop;
        if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, g.dev);
        if (e == cudaSuccess) {
            g.num_sms = prop.multiProcessorCount;
            if (prop.major < 10) e = cudaErrorInvalidDevice;      // sm_100a cubin only
        }
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g.st_compute, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g.st_copy, cudaStreamNonBlocking);
        for (Slot &b : g.slot) {
            if (e == cudaSuccess) e = cudaEventCreate(&b.ev_start);
            if (e == cudaSuccess) e = cudaEventCreate(&b.ev_stop);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.ev_upload, cudaEventDisableTiming);
        }
        if (e != cudaSuccess) {
            for (auto &gg : h->gpus) free_gpu(gg);
            delete h;
            cudaGetLastError();
            return e == cudaErrorInvalidDevice ? SW_ENODEV : SW_ECUDA;
        }
    }
    *out = h;
    return SW_OK;
}

void sw_destroy(sw_handle_t *h)
{
    if (!h) return;
    for (auto &g : h->gpus) {
        cudaSetDevice(g.dev);
        cudaDeviceSynchronize();
        free_gpu(g);
    }
    delete h;
}

int sw_set_strands(sw_handle_t *h, int both)
{
    if (!h) return SW_EINVAL;
    if (h->n_inflight) return SW_EAGAIN;
    h->both_strands = both != 0;
    return SW_OK;
}

int sw_set_queries(sw_handle_t *h, const uint8_t *packed, const uint32_t *len, const uint64_t *off, int nq)
{
    if (!h || nq < 0 || (nq > 0 && (!packed || !len || !off))) return SW_EINVAL;
    if (h->n_inflight) return SW_EAGAIN;
    h->q_packed.clear(); h->q_off.clear(); h->q_len.clear();
    h->q_max_len = 0; h->q_sum_len = 0;
    h->nq_user = nq;
    const int strands = h->both_strands ? 2 : 1;
    for (int st = 0; st < strands; ++st) {
        for (int i = 0; i < nq; ++i) {
            const size_t bytes = ((size_t)len[i] + 3) / 4;
            if (h->q_packed.size() + bytes > 0xFFFFFFF0ull) return SW_EINVAL;
            h->q_off.push_back((uint32_t)h->q_packed.size());
            h->q_len.push_back(len[i]);
            const uint8_t *src = packed + off[i];
            if (st == 0) {
                h->q_packed.insert(h->q_packed.end(), src, src + bytes);
                if (len[i] & 3) h->q_packed.back() &= (uint8_t)((1u << (2 * (len[i] & 3))) - 1u);
            } else {
                // reverse complement: A(10) <-> T(00), C(01) <-> G(11)  =  code ^ 2, order reversed
                const size_t base = h->q_packed.size();
                h->q_packed.resize(base + bytes, 0);
                for (uint32_t k = 0; k < len[i]; ++k) {
                    const uint32_t j = len[i] - 1 - k;
                    const uint8_t c = (uint8_t)(((src[j >> 2] >> ((j & 3) * 2)) & 3) ^ 2);
                    h->q_packed[base + (k >> 2)] |= (uint8_t)(c << ((k & 3) * 2));
                }
            }
            h->q_max_len = std::max(h->q_max_len, len[i]);
            h->q_sum_len += len[i];
        }
    }
    h->q_packed.resize(h->q_packed.size() + 16, 0);
    for (auto &g : h->gpus) {
        int rc = upload_queries(h, g);
        if (rc != SW_OK) return rc;
        for (Slot &b : g.slot) b.scored = false;
    }
    for (Batch &bt : h->batch) bt.scored = false;
    return SW_OK;
}

int sw_plan_shards(const uint32_t *len, size_t ns, int n_shards, uint64_t *starts)
{
    if (n_shards <= 0 || !starts || (ns > 0 && !len)) return SW_EINVAL;
    uint64_t total = 0;
    for (size_t s = 0; s < ns; ++s) total += len[s];
    size_t s = 0;
    uint64_t acc = 0;
    for (int gi = 0; gi < n_shards; ++gi) {
        starts[gi] = s;
        const uint64_t target = (total * (uint64_t)(gi + 1)) / (uint64_t)n_shards;
        if (gi + 1 == n_shards) s = ns;
        else while (s < ns && acc + len[s] / 2 < target) { acc += len[s]; ++s; }
    }
    starts[n_shards] = ns;
    return SW_OK;
}

int sw_load_db(sw_handle_t *h, const uint8_t *packed, const uint32_t *len, const uint64_t *off,
               const uint64_t *ids, size_t ns)
{
    if (!h || (ns > 0 && (!packed || !len || !off))) return SW_EINVAL;
    if (h->n_inflight) return SW_EAGAIN;
    return load_batch(h, 0, packed, len, off, ids, ns);
}

int sw_score_db(sw_handle_t *h)
{
    if (!h) return SW_EINVAL;
    if (h->n_inflight) return SW_EAGAIN;
    if (!h->batch[0].loaded) return SW_ESTATE;
    return score_batch_slot(h, 0);
}

int sw_wait(sw_handle_t *h, int timeout_ms)
{
    if (!h) return SW_EINVAL;
    const bool forever = timeout_ms < 0;
    const auto t_end = std::chrono::steady_clock::now() + std::chrono::milliseconds(forever ? 0 : timeout_ms);
    double ms_max = 0.0;
    for (auto &g : h->gpus) {
        Slot &b = g.slot[0];
        if (!b.scored || b.chunks.empty()) continue;
        SW_CUDA(h, cudaSetDevice(g.dev));
        int rc = wait_event(h, b.chunks.back().done, t_end, forever);
        if (rc != SW_OK) return rc;
        SW_CUDA(h, cudaEventSynchronize(b.ev_stop));
        float ms = 0.f;
        SW_CUDA(h, cudaEventElapsedTime(&ms, b.ev_start, b.ev_stop));
        ms_max = std::max(ms_max, (double)ms);
    }
    h->last_ms = ms_max;
    h->last_cells = h->batch[0].cells;
    return SW_OK;
}

int sw_fetch_db(sw_handle_t *h, int32_t *scores, size_t cap)
{
    if (!h || !scores) return SW_EINVAL;
    if (h->n_inflight) return SW_EAGAIN;
    if (!h->batch[0].loaded || !h->batch[0].scored) return SW_ESTATE;
    return copy_out(h, 0, scores, cap, -1);
}

int sw_score_batch(sw_handle_t *h, const uint8_t *packed, const uint32_t *len, const uint64_t *off,
                   const uint64_t *ids, size_t ns)
{
    if (!h || (ns > 0 && (!packed || !len || !off))) return SW_EINVAL;
    if (h->n_inflight >= 2) return SW_EAGAIN;          // both buffers busy: the bank's `full`
    // take the slot that is not in flight; this replaces whatever sw_load_db left there
    const int si = (h->n_inflight == 1) ? (1 - h->fifo[0]) : 0;
    int rc = load_batch(h, si, packed, len, off, ids, ns);
    if (rc != SW_OK) return rc;
    rc = score_batch_slot(h, si);
    if (rc != SW_OK) return rc;
    h->fifo[h->n_inflight++] = si;
    return SW_OK;
}

int sw_fetch(sw_handle_t *h, int32_t *scores, size_t cap, int timeout_ms)
{
    if (!h || (!scores && cap)) return SW_EINVAL;
    if (h->n_inflight == 0) return SW_ESTATE;
    const int si = h->fifo[0];
    int rc = copy_out(h, si, scores, cap, timeout_ms);
    if (rc == SW_ETIMEOUT || rc == SW_ECAPACITY) return rc;   // batch stays in flight; fetch again
    h->fifo[0] = h->fifo[1];
    h->n_inflight--;
    return rc;
}

int sw_batches_in_flight(const sw_handle_t *h) { return h ? h->n_inflight : 0; }

int sw_fetch_ids(sw_handle_t *h, uint64_t *ids, size_t cap)
{
    if (!h || !ids) return SW_EINVAL;
    const Batch &bt = h->batch[h->last_slot];
    if (cap < bt.ns) return SW_ECAPACITY;
    for (size_t s = 0; s < bt.ns; ++s) ids[s] = bt.have_ids ? bt.ids[s] : (uint64_t)s;
    return SW_OK;
}

int sw_fetch_best(sw_handle_t *h, int32_t *best_score, uint64_t *best_index, int nq_cap)
{
    if (!h || !best_score || !best_index) return SW_EINVAL;
    if (h->n_inflight) return SW_EAGAIN;
    const Batch &bt = h->batch[0];
    const int nq = bt.nq;
    if (!bt.loaded || !bt.scored) return SW_ESTATE;
    if (nq_cap < nq) return SW_ECAPACITY;
    for (int q = 0; q < nq; ++q) { best_score[q] = 0; best_index[q] = 0; }
    std::vector<int32_t> hs(nq);
    std::vector<uint32_t> hi(nq);
    bool first = true;
    for (auto &g : h->gpus) {
        Slot &b = g.slot[0];
        const size_t n = b.s1 - b.s0;
        if (n == 0 || nq == 0) continue;
        SW_CUDA(h, cudaSetDevice(g.dev));
        SW_CUDA(h, g.d_best_score.reserve(nq * sizeof(int32_t)));
        SW_CUDA(h, g.d_best_index.reserve(nq * sizeof(uint32_t)));
        SW_CUDA(h, sw_launch_best(g.st_compute, b.d_out.as<int32_t>(), n, (uint32_t)n, nq,
                                  g.d_best_score.as<int32_t>(), g.d_best_index.as<uint32_t>()));
        h->launches++;
        SW_CUDA(h, cudaMemcpyAsync(hs.data(), g.d_best_score.p, nq * sizeof(int32_t), cudaMemcpyDeviceToHost, g.st_compute));
        SW_CUDA(h, cudaMemcpyAsync(hi.data(), g.d_best_index.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, g.st_compute));
        SW_CUDA(h, cudaStreamSynchronize(g.st_compute));
        for (int q = 0; q < nq; ++q) {
            if (first || hs[q] > best_score[q]) { best_score[q] = hs[q]; best_index[q] = b.s0 + hi[q]; }
        }
        first = false;
    }
    return SW_OK;
}

int sw_query_rows(const sw_handle_t *h) { return h ? (int)h->q_len.size() : 0; }

int sw_last_cuda_error(const sw_handle_t *h) { return h ? h->last_cuda : 0; }
const char *sw_last_cuda_error_string(const sw_handle_t *h)
{
    return cudaGetErrorString((cudaError_t)(h ? h->last_cuda : 0));
}
double sw_last_kernel_ms(const sw_handle_t *h) { return h ? h->last_ms : 0.0; }
uint64_t sw_kernel_launches(const sw_handle_t *h) { return h ? h->launches.load() : 0; }
uint64_t sw_last_cells(const sw_handle_t *h) { return h ? h->last_cells : 0; }
const char *sw_last_kernel_name(const sw_handle_t *h) { return h ? h->last_kernel : "none"; }

int sw_set_kernel_choice(sw_handle_t *h, int rows_per_lane, int lanes_per_pair, int force32)
{
    if (!h) return SW_EINVAL;
    h->force_R = rows_per_lane; h->force_G = lanes_per_pair; h->force32 = force32;
    return SW_OK;
}

int sw_kernel_variant_count(void) { return sw_strip_variant_count(); }

const char *sw_kernel_variant_name(int idx)
{
    const SwStripVariant *v = sw_strip_variant(idx);
    return v ? v->name : nullptr;
}

int sw_set_kernel_name(sw_handle_t *h, const char *name)
{
    if (!h) return SW_EINVAL;
    if (!name || !*name) { h->force_variant = -1; return SW_OK; }
    for (int i = 0; i < sw_strip_variant_count(); ++i)
        if (std::strcmp(sw_strip_variant(i)->name, name) == 0) { h->force_variant = i; return SW_OK; }
    return SW_EINVAL;
}

int sw_set_fixed_penalty_kernels(int enable)
{
    sw_strip_disable_fixed(enable == 0);
    return SW_OK;
}

int sw_set_autotune(sw_handle_t *h, int enable)
{
    if (!h) return SW_EINVAL;
    h->autotune = enable != 0;
    h->tune_choice = -1;
    return SW_OK;
}

int sw_set_arith(sw_handle_t *h, int arith)
{
    if (!h || arith < -1 || arith > 0) return SW_EINVAL;     /* only packed s16 is compiled in */
    h->force_arith = arith;
    return SW_OK;
}

}  // extern "C"
