// b200zk.cu -- device context, workspace management and the C ABI of include/b200zk.h.
// One process drives one B200; every entry point is serialised on the context mutex and
// issues its kernels on the caller's stream (the "_dev" variants) or on the context stream.
// There is no CPU fallback anywhere in this file: without a device the calls fail.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "b200zk.h"
#include "ctx.hpp"
#include "field.cuh"
#include "g1.cuh"
#include "msm.cuh"
#include "ntt.cuh"

using namespace b200zk;

namespace {

thread_local std::string t_err;
std::mutex g_mu;
std::atomic<uint64_t> g_launches{0};

int32_t fail(int32_t code, const std::string& msg) {
    t_err = msg;
    return code;
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            char b_[512];                                                                         \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(e_ == cudaErrorMemoryAllocation ? B200ZK_ERR_OOM : B200ZK_ERR_CUDA, b_);  \
        }                                                                                         \
    } while (0)
#define LAUNCH(kern, grid, block, smem, stream, ...)                                              \
    do {                                                                                          \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                 \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                       \
        CU(cudaGetLastError());                                                                   \
    } while (0)
#define TRY(expr)                                                                                 \
    do {                                                                                          \
        int32_t rc_ = (expr);                                                                     \
        if (rc_ != B200ZK_OK) return rc_;                                                         \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int32_t ensure(size_t bytes) {
        if (bytes <= cap) return B200ZK_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        CU(cudaMalloc(&p, want));
        cap = want;
        return B200ZK_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct BaseTable {
    uint32_t* d = nullptr;  // packed Montgomery affine, 24 limbs per point; `rows` rows of n points
    uint64_t n = 0;
    uint32_t rows = 1;      // > 1: window tables, row w = 2^(c*w) * row 0
    uint32_t c = 0;         // window bits the rows were built for
};

struct NttPlan {
    uint32_t log_n = 0, npass = 0;
    uint32_t deg[3] = {0, 0, 0};
    uint32_t* tw_local[3] = {nullptr, nullptr, nullptr};
    uint32_t* tw_pass[3] = {nullptr, nullptr, nullptr};
    uint32_t* ninv = nullptr;  // Montgomery 1/n (only for inverse plans)
};
struct NttKey {
    uint32_t log_n, inverse;
    uint8_t omega[32];
    bool operator<(const NttKey& o) const {
        if (log_n != o.log_n) return log_n < o.log_n;
        if (inverse != o.inverse) return inverse < o.inverse;
        return memcmp(omega, o.omega, 32) < 0;
    }
};
struct CosetKey {
    uint32_t log_n;
    uint8_t shift[32];
    bool operator<(const CosetKey& o) const {
        if (log_n != o.log_n) return log_n < o.log_n;
        return memcmp(shift, o.shift, 32) < 0;
    }
};

struct Ctx {
    bool inited = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    std::map<uint64_t, BaseTable> tables;
    uint64_t next_handle = 1;
    std::map<NttKey, NttPlan> ntt_plans;
    std::map<CosetKey, uint32_t*> coset_tables;
    uint32_t* fixed_table = nullptr;  // 8 x 256 multiples of G for the synthetic-base generator
    uint32_t tune_c = 0, tune_smax = 0, tune_variant = 6, tune_no_tables = 0;
    // phase timing (b200zk_set_profiling): events recorded on the launching stream
    bool profiling = false;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int ev_count = 0;   // events recorded by the last profiled call
    int ev_kind = 0;    // 1 = msm, 2 = ntt
    MsmPlan last_plan{};
    // MSM workspace
    DevBuf scalars, counts, offsets, cursor, ntask, task_off, entries, task_bucket, task_start, task_len, buckets,
        partials, len_hist, len_off, order, heavy, heavy_items, adhoc, buckets2, aff_a, aff_b, aff_pre, redS[2], redA[2], scan_tmp[2], out_mont, out_canon, stage, flag;
    // NTT workspace
    DevBuf ntt_data, ntt_tmp[2], small;
    // second stream + events for host-buffer MSMs that stream their scalars in two halves
    cudaStream_t copy_stream = nullptr, down_stream = nullptr;
    std::vector<cudaEvent_t> pipe_up, pipe_done;
    cudaEvent_t copy_ev[2] = {nullptr, nullptr};
    // cross-stream ordering of the shared workspaces
    cudaEvent_t ws_event = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_used = false;
};
Ctx g;

int32_t prof_mark(int idx, cudaStream_t s) {
    if (!g.profiling) return B200ZK_OK;
    if (!g.ev[idx]) CU(cudaEventCreate(&g.ev[idx]));
    CU(cudaEventRecord(g.ev[idx], s));
    g.ev_count = idx + 1;
    return B200ZK_OK;
}

// The MSM / NTT workspaces are shared by every call of the process.  Calls on one stream are ordered by the
// stream itself; when a call arrives on a different stream than the previous one, that stream is made to wait
// for the previous call's last kernel, so callers may use any stream without racing on the workspaces.
int32_t ws_enter(cudaStream_t s) {
    if (!g.ws_event) CU(cudaEventCreateWithFlags(&g.ws_event, cudaEventDisableTiming));
    if (g.ws_used && s != g.ws_stream) CU(cudaStreamWaitEvent(s, g.ws_event, 0));
    return B200ZK_OK;
}
int32_t ws_leave(cudaStream_t s) {
    CU(cudaEventRecord(g.ws_event, s));
    g.ws_stream = s;
    g.ws_used = true;
    return B200ZK_OK;
}

int32_t need_init() {
    if (!g.inited) return fail(B200ZK_ERR_NOT_INIT, "b200zk_init has not been called (or failed): no CUDA device is bound");
    return B200ZK_OK;
}

// ------------------------------------------------------------------------------------------
// exclusive scan of n u32 values (in -> out), recursive over 2048-element tiles
// ------------------------------------------------------------------------------------------
int32_t scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, int level, cudaStream_t s) {
    uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles <= 1) {
        LAUNCH(scan_tile_kernel, 1, SCAN_THREADS, 0, s, in, out, n, (uint32_t*)nullptr, (const uint32_t*)nullptr);
        return B200ZK_OK;
    }
    if (level >= 2) return fail(B200ZK_ERR_INVALID_ARG, "scan: input too large");
    TRY(g.scan_tmp[level].ensure(2 * tiles * sizeof(uint32_t)));
    uint32_t* sums = g.scan_tmp[level].as<uint32_t>();
    uint32_t* bases = sums + tiles;
    // pass 1: tile totals only (out is rewritten in pass 3)
    LAUNCH(scan_tile_kernel, (unsigned)tiles, SCAN_THREADS, 0, s, in, out, n, sums, (const uint32_t*)nullptr);
    TRY(scan_u32(sums, bases, tiles, level + 1, s));
    LAUNCH(scan_tile_kernel, (unsigned)tiles, SCAN_THREADS, 0, s, in, out, n, (uint32_t*)nullptr, (const uint32_t*)bases);
    return B200ZK_OK;
}

// ------------------------------------------------------------------------------------------
// MSM
// ------------------------------------------------------------------------------------------
// window bits minimising (mixed adds) + (bucket reduction); `shared` = all windows share one bucket set
uint32_t msm_choose_window(uint64_t n, uint32_t batch, bool shared) {
    uint32_t best_c = 4;
    double best_cost = 1e300;
    for (uint32_t c = 4; c <= 24; c++) {
        uint32_t W = (256 + c - 1) / c;
        double nb = (double)(1u << (c - 1));
        double sets = shared ? 1.0 : (double)W;
        // mixed add ~10 Fp mul per point and window; bucket reduce ~2.2 full adds (14 mul) per bucket
        double cost = W * 10.0 * (double)n + sets * 31.0 * nb;
        if ((double)batch * sets * nb * 192.0 > 4.0e9 && c > 8) continue;   // bucket arrays within ~4 GB
        if (cost < best_cost) { best_cost = cost; best_c = c; }
    }
    return best_c;
}

MsmPlan msm_plan(uint64_t n, uint32_t batch, const BaseTable* tab) {
    MsmPlan pl{};
    bool precomp = tab && tab->rows > 1 && n * 16 >= tab->n;   // tiny slices of a big table: plain path
    if (precomp) {
        pl.c = tab->c;
        pl.precomp = 1;
        pl.row_stride = tab->n;
    } else {
        pl.c = msm_choose_window(n, batch, false);
        if (g.tune_c >= 2 && g.tune_c <= 24) pl.c = g.tune_c;
    }
    pl.W = (256 + pl.c - 1) / pl.c;
    pl.nb = 1u << (pl.c - 1);
    double avg = (double)n * (precomp ? pl.W : 1) / (double)pl.nb;
    uint32_t smax = 32;
    while ((double)smax < 2.0 * avg && smax < (1u << 20)) smax <<= 1;
    // a task is one thread's serial chain: never longer than 1024 additions.  (Splitting typical buckets
    // further to "fill the machine" on small problems was measured to cost more in the collapse step
    // than it saved in the accumulate kernel.)
    if (smax > 1024) smax = 1024;
    if (g.tune_smax) smax = g.tune_smax;
    pl.smax = smax;
    return pl;
}

// d_scalars: batch*n Fr (device).  d_bases: n packed Montgomery affine points.
// One chunk of a scalar vector that arrives in pieces (host-buffer MSMs): n_total fixes the plan, i0 is the index of the
// chunk's first point, the first chunk owns g.buckets, later chunks accumulate into g.buckets2 and are folded in, the
// last chunk runs the tail.
struct MsmChunk {
    uint64_t n_total, i0;
    bool first, last;
};
int32_t msm_run(const BaseTable* tab, const uint32_t* d_bases, const uint32_t* d_scalars, uint64_t n, uint32_t batch,
                uint32_t scalar_fmt, uint32_t* d_out_mont, uint32_t* d_out_canon, cudaStream_t s,
                uint32_t* d_out_xyzz = nullptr, const MsmChunk* ck = nullptr) {
    if (n == 0 || batch == 0) {
        if (d_out_mont) CU(cudaMemsetAsync(d_out_mont, 0, 96 * (size_t)std::max(batch, 1u), s));
        if (d_out_canon) CU(cudaMemsetAsync(d_out_canon, 0, 96 * (size_t)std::max(batch, 1u), s));
        return B200ZK_OK;
    }
    if (n >= (1ull << 31)) return fail(B200ZK_ERR_INVALID_ARG, "msm: n must be < 2^31");
    MsmPlan pl = msm_plan(ck ? ck->n_total : n, batch, tab);
    uint64_t nwin = (uint64_t)batch * (pl.precomp ? 1 : pl.W);   // bucket sets
    uint64_t NBt = nwin * pl.nb;
    uint64_t max_entries = (uint64_t)batch * n * pl.W;
    if (max_entries >= (1ull << 32) || NBt >= (1ull << 31) || (pl.precomp && pl.row_stride * pl.W >= (1ull << 31)))
        return fail(B200ZK_ERR_INVALID_ARG, "msm: batch * n * windows exceeds 2^32 entries; split the batch");
    uint64_t max_tasks = std::min(NBt, max_entries) + max_entries / pl.smax + 1;

    TRY(g.counts.ensure((NBt + 1) * 4));
    TRY(g.offsets.ensure((NBt + 1) * 4));
    TRY(g.cursor.ensure((NBt + 1) * 4));
    TRY(g.ntask.ensure((NBt + 1) * 4));
    TRY(g.task_off.ensure((NBt + 1) * 4));
    TRY(g.entries.ensure(max_entries * 4));
    TRY(g.task_bucket.ensure(max_tasks * 4));
    TRY(g.task_start.ensure(max_tasks * 4));
    TRY(g.task_len.ensure(max_tasks * 4));
    TRY(g.len_hist.ensure(((size_t)pl.smax + 2) * 4));
    TRY(g.len_off.ensure(((size_t)pl.smax + 2) * 4));
    TRY(g.order.ensure(max_tasks * 4));
    const uint64_t hmax = max_entries / pl.smax + 2;                       // heavy buckets (split into > 1 task)
    const uint64_t imax = hmax + max_tasks / HEAVY_CHUNK + 2;              // their work items
    TRY(g.heavy.ensure((4 + 3 * hmax + 2 * imax) * 4));
    TRY(g.heavy_items.ensure(imax * 192));
    TRY(g.buckets.ensure(NBt * 192));
    if (ck && !ck->first) TRY(g.buckets2.ensure(NBt * 192));
    TRY(g.partials.ensure(max_tasks * 192));
    uint64_t m1 = (pl.nb + RED_RADIX - 1) / RED_RADIX, m2 = (m1 + RED_RADIX - 1) / RED_RADIX;
    TRY(g.redS[0].ensure(nwin * m1 * 192));
    TRY(g.redA[0].ensure(nwin * m1 * 192));
    TRY(g.redS[1].ensure(nwin * m2 * 192));
    TRY(g.redA[1].ensure(nwin * m2 * 192));

    uint32_t* counts = g.counts.as<uint32_t>();
    uint32_t* offsets = g.offsets.as<uint32_t>();
    uint32_t* cursor = g.cursor.as<uint32_t>();
    uint32_t* ntask = g.ntask.as<uint32_t>();
    uint32_t* task_off = g.task_off.as<uint32_t>();
    uint32_t* entries = g.entries.as<uint32_t>();
    uint32_t* buckets = (ck && !ck->first) ? g.buckets2.as<uint32_t>() : g.buckets.as<uint32_t>();
    const uint64_t i0 = ck ? ck->i0 : 0;
    uint32_t* partials = g.partials.as<uint32_t>();

    TRY(ws_enter(s));
    g.ev_kind = 1;
    TRY(prof_mark(0, s));
    CU(cudaMemsetAsync(counts, 0, (NBt + 1) * 4, s));
    CU(cudaMemsetAsync(ntask, 0, (NBt + 1) * 4, s));
    dim3 dgrid((unsigned)((n + 255) / 256), batch);
    const uint32_t fmt_mont = scalar_fmt == B200ZK_FMT_MONT ? 1u : 0u;
    // (A two-pass sort -- coarse bins of 2048 buckets staged through shared memory, then one CTA per bin --
    // was built and measured at 2^24: 11.0 ms against 7.9 ms for this one-pass histogram + scatter; removed.)
    LAUNCH(msm_digits_kernel<0>, dgrid, 256, 0, s, d_scalars, n, fmt_mont, pl, counts, (uint32_t*)nullptr, 0u, 0xffffffffu, (const uint32_t*)nullptr, 0u, i0);
    TRY(scan_u32(counts, offsets, NBt + 1, 0, s));
    CU(cudaMemcpyAsync(cursor, offsets, (NBt + 1) * 4, cudaMemcpyDeviceToDevice, s));
    {
        // The scatter writes 4-byte entries into bucket lists that are spread over the whole entry array.  When that
        // array is much larger than L2, the 32-byte sector a list is currently filling is evicted half full and read
        // back (DRAM read-modify-write).  Scattering one bucket range at a time keeps the open sectors (32 B per bucket
        // of the range) resident; the scalars are re-read and re-coded once per pass, which is cheap.
        static int passes_env = -1;
        if (passes_env < 0) { const char* v = getenv("B200ZK_SCATTER_PASSES"); passes_env = v ? atoi(v) : 0; }
        uint32_t passes = 1;
        if (passes_env > 0) passes = (uint32_t)passes_env;
        else if (NBt * 32 >= (48ull << 20)) passes = 2;   // measured at 2^24 (2^21 buckets): 1 pass 8.98 ms, 2 passes 7.44, 4 passes 9.22 (each pass re-codes every scalar)
        for (uint32_t ps = 0; ps < passes; ps++) {
            uint32_t lo = (uint32_t)(NBt * ps / passes), hi = (uint32_t)(NBt * (ps + 1) / passes);
            LAUNCH(msm_digits_kernel<1>, dgrid, 256, 0, s, d_scalars, n, fmt_mont, pl, cursor, entries, lo, hi,
                   passes > 1 ? (const uint32_t*)(offsets + NBt) : (const uint32_t*)nullptr, 100u << 20, i0);
        }
    }
    unsigned bgrid = (unsigned)((NBt + 255) / 256);
    LAUNCH(msm_task_count_kernel, bgrid, 256, 0, s, (const uint32_t*)counts, NBt, pl.smax, ntask);
    TRY(scan_u32(ntask, task_off, NBt + 1, 0, s));
    uint32_t* len_hist = g.len_hist.as<uint32_t>();
    uint32_t* len_off = g.len_off.as<uint32_t>();
    uint32_t* order = g.order.as<uint32_t>();
    HeavyArrays hv;
    hv.count = g.heavy.as<uint32_t>();
    hv.bucket = hv.count + 4;
    hv.base = hv.bucket + hmax;
    hv.done = hv.base + hmax;
    hv.item_slot = hv.done + hmax;
    hv.item_chunk = hv.item_slot + imax;
    CU(cudaMemsetAsync(len_hist, 0, ((size_t)pl.smax + 2) * 4, s));
    CU(cudaMemsetAsync(hv.count, 0, 16, s));
    LAUNCH(msm_task_emit_kernel, bgrid, 256, 0, s, (const uint32_t*)counts, (const uint32_t*)offsets,
           (const uint32_t*)task_off, NBt, pl.smax, g.task_bucket.as<uint32_t>(), g.task_start.as<uint32_t>(),
           g.task_len.as<uint32_t>(), len_hist);
    TRY(scan_u32(len_hist, len_off, (uint64_t)pl.smax + 1, 0, s));
    LAUNCH(msm_task_order_kernel, (unsigned)((max_tasks + 255) / 256), 256, 0, s, (const uint32_t*)g.task_len.as<uint32_t>(),
           (const uint32_t*)(task_off + NBt), pl.smax, len_off, order);
    LAUNCH(msm_heavy_list_kernel, bgrid, 256, 0, s, (const uint32_t*)ntask, NBt, hv);
    CU(cudaMemsetAsync(buckets, 0, NBt * 192, s));
    TRY(prof_mark(1, s));
    {
        unsigned agrid = (unsigned)((max_tasks + 127) / 128);
        const uint32_t *tb = g.task_bucket.as<uint32_t>(), *ts = g.task_start.as<uint32_t>(), *tl = g.task_len.as<uint32_t>();
        const uint32_t* ntp = task_off + NBt;
        switch (g.tune_variant) {
            case 1: LAUNCH(msm_accumulate_kernel_v1, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
            case 2: LAUNCH(msm_accumulate_kernel_v2, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
            case 7: LAUNCH(msm_accumulate_kernel_v7, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
            case 6: LAUNCH(msm_accumulate_kernel_v6, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
            case 3: LAUNCH(msm_accumulate_kernel_v3, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
            case 4:
            case 5: {
                size_t slots = max_entries / 2 + 2;
                TRY(g.aff_a.ensure(slots * 96));
                TRY(g.aff_b.ensure(slots * 96));
                TRY(g.aff_pre.ensure(slots * 48));
                if (g.tune_variant == 4)
                    LAUNCH(msm_accumulate_affine_kernel, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials,
                           g.aff_a.as<uint32_t>(), g.aff_b.as<uint32_t>(), g.aff_pre.as<uint32_t>());
                else
                    LAUNCH(msm_accumulate_affine_kernel_r168, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials,
                           g.aff_a.as<uint32_t>(), g.aff_b.as<uint32_t>(), g.aff_pre.as<uint32_t>());
                break;
            }
            default: LAUNCH(msm_accumulate_kernel, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
        }
    }
    TRY(prof_mark(2, s));
    LAUNCH(msm_collapse_kernel, (unsigned)(g.prop.multiProcessorCount * 4), 128, 0, s, (const uint32_t*)ntask,
           (const uint32_t*)task_off, hv, (const uint32_t*)partials, g.heavy_items.as<uint32_t>(), buckets);

    if (ck && !ck->first)
        LAUNCH(msm_bucket_merge_kernel, (unsigned)((NBt + 127) / 128), 128, 0, s, g.buckets.as<uint32_t>(), (const uint32_t*)buckets, NBt);
    if (ck && !ck->last) {
        TRY(ws_leave(s));
        g.last_plan = pl;
        return B200ZK_OK;
    }
    buckets = g.buckets.as<uint32_t>();

    // bucket reduction tree: work-efficient serial radix-16 groups while there are enough of them to fill
    // the machine, then warp-cooperative radix-32 groups (short dependency chains) for the upper levels
    const uint32_t* S_in = buckets;
    const uint32_t* A_in = nullptr;
    uint32_t m = pl.nb, scale_log = 0;
    int pp = 0;
    const uint32_t* win_sums = nullptr;
    do {
        // serial groups while there are >= 4096 of them: at 2^24 the second level (8192 groups) takes 0.5 ms serially
        // against 1.5 ms for 4096 warp-cooperative groups (1.7 waves of a latency-bound kernel)
        bool serial = (uint64_t)((m + RED_RADIX - 1) / RED_RADIX) * nwin >= 4096;
        uint32_t radix = serial ? RED_RADIX : COOP_RADIX;
        uint32_t m_out = (m + radix - 1) / radix;
        uint32_t* S_out = g.redS[pp].as<uint32_t>();
        uint32_t* A_out = g.redA[pp].as<uint32_t>();
        uint64_t groups = (uint64_t)m_out * nwin;
        if (serial)
            LAUNCH(msm_reduce_kernel, (unsigned)((groups + 63) / 64), 64, 0, s, S_in, A_in, S_out, A_out, m, m_out,
                   (uint32_t)nwin, scale_log);
        else if (groups <= 600)
            // top of the tree: one CTA per group, 8 lanes per node, additions in 4 product levels instead of 14 products
            LAUNCH(msm_reduce_coop8_kernel, (unsigned)groups, 256, 0, s, S_in, A_in, S_out, A_out, m, m_out, (uint32_t)nwin, scale_log);
        else
            LAUNCH(msm_reduce_coop_kernel, (unsigned)((groups + 3) / 4), 128, 0, s, S_in, A_in, S_out, A_out, m, m_out,
                   (uint32_t)nwin, scale_log);
        S_in = S_out;
        A_in = A_out;
        win_sums = A_out;
        m = m_out;
        scale_log += serial ? RED_LOG : COOP_LOG;
        pp ^= 1;
    } while (m > 1);
    LAUNCH(msm_combine_kernel, batch, 32, 0, s, win_sums, pl.precomp ? 1u : pl.W, pl.c, d_out_mont, d_out_canon, d_out_xyzz);
    TRY(prof_mark(3, s));
    TRY(ws_leave(s));
    g.last_plan = pl;
    return B200ZK_OK;
}

int32_t lookup_bases(uint64_t handle, uint64_t offset, uint64_t n, const uint32_t** out, const BaseTable** tab = nullptr) {
    auto it = g.tables.find(handle);
    if (it == g.tables.end()) return fail(B200ZK_ERR_BAD_HANDLE, "unknown base-table handle");
    if (offset > it->second.n || n > it->second.n - offset)
        return fail(B200ZK_ERR_INVALID_ARG, "msm: offset + n exceeds the registered table");
    *out = it->second.d + 24 * offset;
    if (tab) *tab = &it->second;
    return B200ZK_OK;
}

int32_t ingest_bases(const uint8_t* d_src, uint64_t n, uint32_t fmt, uint32_t stride, uint32_t* d_dst, cudaStream_t s) {
    TRY(g.flag.ensure(4));
    CU(cudaMemsetAsync(g.flag.p, 0, 4, s));
    LAUNCH(g1_ingest_kernel, (unsigned)((n + 127) / 128), 128, 0, s, d_src, n, stride, fmt == B200ZK_FMT_MONT ? 1u : 0u,
           d_dst, g.flag.as<uint32_t>());
    uint32_t bad = 0;
    CU(cudaMemcpyAsync(&bad, g.flag.p, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (bad) return fail(B200ZK_ERR_BAD_POINT, "a base point is not a canonical point on the curve");
    return B200ZK_OK;
}

// ------------------------------------------------------------------------------------------
// NTT
// ------------------------------------------------------------------------------------------
int32_t upload_fr_mont(const uint8_t v[32], uint32_t* d_dst, cudaStream_t s) {
    CU(cudaMemcpyAsync(d_dst, v, 32, cudaMemcpyHostToDevice, s));
    LAUNCH(fr_convert_kernel, 1, 32, 0, s, d_dst, (uint64_t)1, 1u);
    return B200ZK_OK;
}

int32_t ntt_get_plan(uint32_t log_n, const uint8_t omega[32], bool inverse, cudaStream_t s, NttPlan** out) {
    NttKey key;
    key.log_n = log_n;
    key.inverse = inverse ? 1 : 0;
    memcpy(key.omega, omega, 32);
    auto it = g.ntt_plans.find(key);
    if (it != g.ntt_plans.end()) { *out = &it->second; return B200ZK_OK; }
    NttPlan pl;
    pl.log_n = log_n;
    pl.npass = (log_n + NTT_LOGB - 1) / NTT_LOGB;
    if (pl.npass == 0) pl.npass = 1;
    if (pl.npass > 3) return fail(B200ZK_ERR_INVALID_ARG, "ntt: log_n > 33 is not supported");
    for (uint32_t i = 0; i < pl.npass; i++) pl.deg[i] = log_n / pl.npass + (i < log_n % pl.npass ? 1 : 0);
    uint32_t* d_omega = nullptr;
    CU(cudaMalloc(&d_omega, 32));
    TRY(upload_fr_mont(omega, d_omega, s));
    if (inverse) {
        CU(cudaMalloc(&pl.ninv, 32));
        LAUNCH(fr_inv_pow2_kernel, 1, 1, 0, s, pl.ninv, log_n);
    }
    uint32_t log_s = 0;
    for (uint32_t i = 0; i < pl.npass; i++) {
        uint32_t deg = pl.deg[i];
        uint64_t cnt = deg ? ((uint64_t)1 << (deg - 1)) : 1;
        CU(cudaMalloc(&pl.tw_local[i], cnt * 32));
        LAUNCH(fr_powers_kernel, (unsigned)((cnt + 127) / 128), 128, 0, s, pl.tw_local[i], (const uint32_t*)d_omega,
               (const uint32_t*)nullptr, cnt, (uint64_t)1 << (log_n - deg), 0u, 0u);
        if (i + 1 < pl.npass) {
            uint64_t m = (uint64_t)1 << (log_n - log_s);
            CU(cudaMalloc(&pl.tw_pass[i], m * 32));
            const uint32_t* scale = (i == 0 && inverse) ? pl.ninv : nullptr;
            LAUNCH(fr_powers_kernel, (unsigned)((m + 127) / 128), 128, 0, s, pl.tw_pass[i], (const uint32_t*)d_omega, scale,
                   m, (uint64_t)1 << log_s, 1u, deg);
        }
        log_s += deg;
    }
    CU(cudaStreamSynchronize(s));
    CU(cudaFree(d_omega));
    auto ins = g.ntt_plans.emplace(key, pl);
    *out = &ins.first->second;
    return B200ZK_OK;
}

int32_t ntt_get_coset(uint32_t log_n, const uint8_t shift[32], cudaStream_t s, uint32_t** out) {
    CosetKey key;
    key.log_n = log_n;
    memcpy(key.shift, shift, 32);
    auto it = g.coset_tables.find(key);
    if (it != g.coset_tables.end()) { *out = it->second; return B200ZK_OK; }
    uint64_t n = (uint64_t)1 << log_n;
    uint32_t *d_shift = nullptr, *tab = nullptr;
    CU(cudaMalloc(&d_shift, 32));
    TRY(upload_fr_mont(shift, d_shift, s));
    CU(cudaMalloc(&tab, n * 32));
    LAUNCH(fr_powers_kernel, (unsigned)((n + 127) / 128), 128, 0, s, tab, (const uint32_t*)d_shift, (const uint32_t*)nullptr, n,
           (uint64_t)1, 0u, 0u);
    CU(cudaStreamSynchronize(s));
    CU(cudaFree(d_shift));
    g.coset_tables[key] = tab;
    *out = tab;
    return B200ZK_OK;
}

int32_t ntt_run(uint32_t* d_data, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                const uint8_t* coset_shift, cudaStream_t s) {
    if (log_n > 32) return fail(B200ZK_ERR_INVALID_ARG, "ntt: log_n exceeds the 2-adicity of Fr");
    if ((flags & (B200ZK_NTT_COSET_IN | B200ZK_NTT_COSET_OUT)) && !coset_shift)
        return fail(B200ZK_ERR_INVALID_ARG, "ntt: COSET flag without a shift");
    if ((flags & B200ZK_NTT_COSET_IN) && (flags & B200ZK_NTT_COSET_OUT))
        return fail(B200ZK_ERR_INVALID_ARG, "ntt: COSET_IN and COSET_OUT are mutually exclusive");
    if (batch == 0) return B200ZK_OK;
    bool inverse = (flags & B200ZK_NTT_INVERSE_SCALE) != 0;
    NttPlan* pl = nullptr;
    TRY(ntt_get_plan(log_n, omega, inverse, s, &pl));
    uint32_t* coset = nullptr;
    if (flags & (B200ZK_NTT_COSET_IN | B200ZK_NTT_COSET_OUT)) TRY(ntt_get_coset(log_n, coset_shift, s, &coset));
    uint64_t n = (uint64_t)1 << log_n, total = n * batch;
    uint32_t* bufs[2] = {nullptr, nullptr};
    if (pl->npass > 1) {
        TRY(g.ntt_tmp[0].ensure(total * 32));
        bufs[0] = g.ntt_tmp[0].as<uint32_t>();
        if (pl->npass > 2) {
            TRY(g.ntt_tmp[1].ensure(total * 32));
            bufs[1] = g.ntt_tmp[1].as<uint32_t>();
        }
    }
    static bool attr_set = false;
    static int ntt_variant = 1;   // 1: two CTAs per SM (default); 0: one CTA per SM.  B200ZK_NTT_VARIANT overrides.
    if (!attr_set) {
        CU(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
        CU(cudaFuncSetAttribute(ntt_pass_kernel_occ2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
        CU(cudaFuncSetAttribute(ntt_pass_kernel_call2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
        CU(cudaFuncSetAttribute(ntt_pass_kernel_call3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
        CU(cudaFuncSetAttribute(ntt_pass_kernel_plain2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
        CU(cudaFuncSetAttribute(ntt_pass_kernel_wl2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
        if (const char* v = getenv("B200ZK_NTT_VARIANT")) ntt_variant = atoi(v);
        attr_set = true;
    }
    uint32_t log_s = 0;
    const uint32_t* src = d_data;
    TRY(ws_enter(s));
    g.ev_kind = 2;
    TRY(prof_mark(0, s));
    for (uint32_t i = 0; i < pl->npass; i++) {
        bool last = (i + 1 == pl->npass);
        uint32_t* dst = last ? d_data : bufs[i & 1];
        if (pl->npass == 3 && i == 1) dst = bufs[1];
        NttPassArgs a;
        memset(&a, 0, sizeof a);
        a.in = src;
        a.out = dst;
        a.tw_local = pl->tw_local[i];
        a.tw_pass = pl->tw_pass[i];
        a.in_scale = (i == 0 && (flags & B200ZK_NTT_COSET_IN)) ? coset : nullptr;
        a.out_scale = (last && (flags & B200ZK_NTT_COSET_OUT)) ? coset : nullptr;
        a.scalar = (last && inverse && pl->npass == 1) ? pl->ninv : nullptr;
        a.reduce_in = (i == 0 && !(flags & B200ZK_NTT_MONT)) ? 1 : 0;
        a.log_n = log_n;
        a.deg = pl->deg[i];
        a.log_s = log_s;
        a.log_cols = log_n - pl->deg[i];
        a.total_cols = (uint64_t)batch << a.log_cols;
        uint64_t ctas = (total + NTT_B - 1) / NTT_B;
        if (ntt_variant == 0) LAUNCH(ntt_pass_kernel, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else if (ntt_variant == 2) LAUNCH(ntt_pass_kernel_call2, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else if (ntt_variant == 3) LAUNCH(ntt_pass_kernel_call3, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else if (ntt_variant == 4) LAUNCH(ntt_pass_kernel_plain2, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else if (ntt_variant == 5) LAUNCH(ntt_pass_kernel_wl2, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else LAUNCH(ntt_pass_kernel_occ2, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        TRY(prof_mark((int)i + 1, s));
        src = dst;
        log_s += pl->deg[i];
    }
    TRY(ws_leave(s));
    return B200ZK_OK;
}

// ------------------------------------------------------------------------------------------
// self-test and micro-benchmark kernels
// ------------------------------------------------------------------------------------------
template <class P>
__global__ void selftest_kernel(const uint32_t* a, const uint32_t* b, uint32_t* out, uint64_t count, uint32_t op) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fe<P> x, y, r;
    for (int k = 0; k < P::N; k++) { x.l[k] = a[P::N * i + k]; y.l[k] = b ? b[P::N * i + k] : 0; }
    x = fe_to_mont(x);
    y = fe_to_mont(y);
    if (op == 0) r = fe_mul(x, y);
    else if (op == 1) r = fe_add(x, y);
    else if (op == 2) r = fe_sub(x, y);
    else r = fe_inv(x);
    r = fe_from_mont(r);
    for (int k = 0; k < P::N; k++) out[P::N * i + k] = r.l[k];
}

__global__ void __launch_bounds__(256) mb_imad_wide_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    // plain IMAD.WIDE.U32 (no carry in or out); the multiplicand rotates through the other
    // accumulators so that ptxas cannot hoist the product out of the loop
    uint64_t acc[8];
    uint32_t b = b0 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = a0 + threadIdx.x * 8 + k;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = (uint64_t)(uint32_t)acc[(k + 1) & 7] * b + acc[k];
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) mb_imad_pair_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    uint32_t lo[8], hi[8];
    uint32_t a = a0 + threadIdx.x, b = b0 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { lo[k] = k; hi[k] = k + 1; }
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[k]) : "r"(a), "r"(b));
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= ((uint64_t)hi[k] << 32) | lo[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// carry-chained rows exactly as fe_mul issues them: 4 independent rows of 6 IMAD.WIDE.U32.X
__global__ void __launch_bounds__(256) mb_imad_chain_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    uint32_t acc[4][12];
    uint32_t a = a0 + threadIdx.x, b = b0 + blockIdx.x;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int k = 0; k < 12; k++) acc[r][k] = r * 12 + k;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            mad_wide_cc(acc[r][0], acc[r][1], a, b, acc[r][0], acc[r][1]);
#pragma unroll
            for (int k = 2; k < 12; k += 2) madc_wide_cc(acc[r][k], acc[r][k + 1], a, b, acc[r][k], acc[r][k + 1]);
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int k = 0; k < 12; k++) s = s * 31 + acc[r][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// every wide MAD produces a carry-out but takes no carry-in
__global__ void __launch_bounds__(256) mb_imad_cout_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    uint32_t lo[8], hi[8];
    uint32_t a = a0 + threadIdx.x, b = b0 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { lo[k] = k; hi[k] = k + 1; }
    uint32_t sink = 0;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) mad_wide_cc(lo[k], hi[k], a, b, lo[k], hi[k]);
        }
        sink = addc(sink, 0);
    }
    uint64_t s = sink;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= ((uint64_t)hi[k] << 32) | lo[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) mb_dfma_kernel(uint64_t* out, uint32_t iters, double a0, double b0) {
    double acc[8];
    double a = a0 + threadIdx.x * 1e-9, b = b0 + blockIdx.x * 1e-9;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = k;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[k]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)__double_as_longlong(s);
}
// one LOP3 (ALU pipe) per IMAD.WIDE (FMA pipe): do the two pipes issue side by side?
__global__ void __launch_bounds__(256) mb_imad_alu_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    uint64_t acc[8];
    uint32_t x[8];
    uint32_t b = b0 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = a0 + threadIdx.x * 8 + k; x[k] = k * 3; }
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                acc[k] = (uint64_t)(uint32_t)acc[(k + 1) & 7] * b + acc[k];
                x[k] = (x[k] ^ x[(k + 3) & 7]) & ~x[(k + 5) & 7];
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k] + x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// MODE 0: IMAD.HI.U32 only; 1: 32-bit IMAD only; 2: unfused pair IMAD + IMAD.HI.U32 on the same operands with an
// immediate multiplier (what ptxas emits for the m*p rows when the modulus limb is not in a plain register).
// The multiplicand rotates through the accumulators so that nothing can be hoisted.
template <int MODE>
__global__ void __launch_bounds__(256) mb_imad_parts_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    uint32_t lo[8], hi[8];
    uint32_t b = b0 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { lo[k] = a0 + threadIdx.x * 8 + k; hi[k] = k + 1; }
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (MODE == 0) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[k]) : "r"(hi[(k + 1) & 7]), "r"(b));
                else if (MODE == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(lo[(k + 1) & 7]), "r"(b));
                else {
                    uint32_t m = lo[(k + 1) & 7];
                    asm volatile("mad.lo.cc.u32 %0, %2, 0x53bda402, %0;\n\tmadc.hi.u32 %1, %2, 0x53bda402, %1;"
                                 : "+r"(lo[k]), "+r"(hi[k]) : "r"(m));
                }
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= ((uint64_t)hi[k] << 32) | lo[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class P>
__global__ void __launch_bounds__(256) mb_femul_kernel(uint32_t* out, uint32_t iters) {
    Fe<P> x = fe_one<P>(), y = fe_one<P>();
    x.l[0] += threadIdx.x;
    y.l[1] += blockIdx.x;
    for (uint32_t it = 0; it < iters; it++) {
        x = fe_mul(x, y);
        y = fe_mul(y, x);
    }
    uint32_t s = 0;
    for (int k = 0; k < P::N; k++) s ^= x.l[k] ^ y.l[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(128) mb_madd_kernel(const uint32_t* gen_mont, uint32_t* out, uint32_t iters) {
    G1Affine q = g1a_ldg(gen_mont, 0);
    G1Xyzz acc;
    xyzz_from_affine(acc, q, false);
    xyzz_dbl(acc);
    for (uint32_t k = 0; k < (threadIdx.x & 7); k++) xyzz_dbl(acc);
    for (uint32_t it = 0; it < iters; it++) xyzz_add_mixed(acc, q, false);
    uint32_t s = 0;
    for (int k = 0; k < 12; k++) s ^= acc.x.l[k] ^ acc.zzz.l[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// generator of G1 in canonical wire form (its compressed form is the KAT at
// /root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:125)
const uint8_t G1_GEN_X_BE[48] = {0x17, 0xf1, 0xd3, 0xa7, 0x31, 0x97, 0xd7, 0x94, 0x26, 0x95, 0x63, 0x8c, 0x4f, 0xa9, 0xac, 0x0f,
                                 0xc3, 0x68, 0x8c, 0x4f, 0x97, 0x74, 0xb9, 0x05, 0xa1, 0x4e, 0x3a, 0x3f, 0x17, 0x1b, 0xac, 0x58,
                                 0x6c, 0x55, 0xe8, 0x3f, 0xf9, 0x7a, 0x1a, 0xef, 0xfb, 0x3a, 0xf0, 0x0a, 0xdb, 0x22, 0xc6, 0xbb};
const uint8_t G1_GEN_Y_BE[48] = {0x08, 0xb3, 0xf4, 0x81, 0xe3, 0xaa, 0xa0, 0xf1, 0xa0, 0x9e, 0x30, 0xed, 0x74, 0x1d, 0x8a, 0xe4,
                                 0xfc, 0xf5, 0xe0, 0x95, 0xd5, 0xd0, 0x0a, 0xf6, 0x00, 0xdb, 0x18, 0xcb, 0x2c, 0x04, 0xb3, 0xed,
                                 0xd0, 0x3c, 0xc7, 0x44, 0xa2, 0x88, 0x8a, 0xe4, 0x0c, 0xaa, 0x23, 0x29, 0x46, 0xc5, 0xe7, 0xe1};

// device copy of the generator (Montgomery affine), built on first use
int32_t get_generator_dev(uint32_t** out, cudaStream_t s) {
    static uint32_t* d_gen = nullptr;
    if (!d_gen) {
        uint8_t wire[96];
        for (int i = 0; i < 48; i++) { wire[i] = G1_GEN_X_BE[47 - i]; wire[48 + i] = G1_GEN_Y_BE[47 - i]; }
        uint8_t* d_wire = nullptr;
        CU(cudaMalloc(&d_wire, 96));
        CU(cudaMalloc(&d_gen, 96));
        CU(cudaMemcpyAsync(d_wire, wire, 96, cudaMemcpyHostToDevice, s));
        int32_t rc = ingest_bases(d_wire, 1, B200ZK_FMT_CANONICAL, 96, d_gen, s);
        cudaFree(d_wire);
        if (rc != B200ZK_OK) { cudaFree(d_gen); d_gen = nullptr; return rc; }
    }
    *out = d_gen;
    return B200ZK_OK;
}

}  // namespace

// what b200zk_ext.cu shares with this context (ctx.hpp)
namespace b200zk_ctx {
int32_t fail(int32_t code, const std::string& msg) { return ::fail(code, msg); }
std::mutex& mutex() { return g_mu; }
int32_t need_init() { return ::need_init(); }
cudaStream_t stream() { return g.stream; }
int sm_count() { return g.prop.multiProcessorCount; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int32_t generator_dev(uint32_t** out, cudaStream_t s) { return get_generator_dev(out, s); }
static std::vector<void (*)()> g_hooks;
void on_shutdown(void (*fn)()) { g_hooks.push_back(fn); }
int32_t ws_enter(cudaStream_t s) { return ::ws_enter(s); }
int32_t ws_leave(cudaStream_t s) { return ::ws_leave(s); }
}  // namespace b200zk_ctx

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int32_t b200zk_init(int32_t device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.inited) return B200ZK_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(B200ZK_ERR_NO_DEVICE,
                    std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) return fail(B200ZK_ERR_NO_DEVICE, "device index out of range");
    CU(cudaSetDevice(device));
    CU(cudaGetDeviceProperties(&g.prop, device));
    if (g.prop.major != 10)
        return fail(B200ZK_ERR_NO_DEVICE, std::string("device '") + g.prop.name +
                                              "' is not sm_100: this library ships sm_100a code only and has no fallback");
    CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    g.device = device;
    g.inited = true;
    return B200ZK_OK;
}

int32_t b200zk_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g.inited) return B200ZK_OK;
    cudaSetDevice(g.device);
    cudaDeviceSynchronize();
    for (auto fn : b200zk_ctx::g_hooks) fn();
    for (auto& kv : g.tables) cudaFree(kv.second.d);
    g.tables.clear();
    for (auto& kv : g.ntt_plans) {
        for (int i = 0; i < 3; i++) {
            if (kv.second.tw_local[i]) cudaFree(kv.second.tw_local[i]);
            if (kv.second.tw_pass[i]) cudaFree(kv.second.tw_pass[i]);
        }
        if (kv.second.ninv) cudaFree(kv.second.ninv);
    }
    g.ntt_plans.clear();
    for (auto& kv : g.coset_tables) cudaFree(kv.second);
    g.coset_tables.clear();
    if (g.fixed_table) { cudaFree(g.fixed_table); g.fixed_table = nullptr; }
    DevBuf* all[] = {&g.scalars, &g.counts, &g.offsets, &g.cursor, &g.ntask, &g.task_off, &g.entries, &g.task_bucket,
                     &g.task_start, &g.task_len, &g.buckets, &g.partials, &g.len_hist, &g.len_off, &g.order, &g.heavy, &g.heavy_items, &g.adhoc, &g.buckets2, &g.aff_a, &g.aff_b, &g.aff_pre, &g.redS[0], &g.redS[1], &g.redA[0], &g.redA[1],
                     &g.scan_tmp[0], &g.scan_tmp[1], &g.out_mont, &g.out_canon, &g.stage, &g.flag, &g.ntt_data,
                     &g.ntt_tmp[0], &g.ntt_tmp[1], &g.small};
    for (DevBuf* b : all) b->release();
    if (g.down_stream) { cudaStreamDestroy(g.down_stream); g.down_stream = nullptr; }
    for (cudaEvent_t e : g.pipe_up) cudaEventDestroy(e);
    for (cudaEvent_t e : g.pipe_done) cudaEventDestroy(e);
    g.pipe_up.clear();
    g.pipe_done.clear();
    if (g.copy_stream) {
        cudaStreamDestroy(g.copy_stream);
        cudaEventDestroy(g.copy_ev[0]);
        cudaEventDestroy(g.copy_ev[1]);
        g.copy_stream = nullptr;
    }
    if (g.ws_event) { cudaEventDestroy(g.ws_event); g.ws_event = nullptr; }
    g.ws_used = false;
    if (g.stream) cudaStreamDestroy(g.stream);
    g.stream = nullptr;
    g.inited = false;
    return B200ZK_OK;
}

int32_t b200zk_last_error(char* buf, size_t len) {
    if (!buf || len == 0) return B200ZK_ERR_INVALID_ARG;
    snprintf(buf, len, "%s", t_err.c_str());
    return B200ZK_OK;
}

int32_t b200zk_device_info(char* buf, size_t len) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (!buf || len == 0) return fail(B200ZK_ERR_INVALID_ARG, "null buffer");
    snprintf(buf, len, "%s sm_%d%d %dSM", g.prop.name, g.prop.major, g.prop.minor, g.prop.multiProcessorCount);
    return B200ZK_OK;
}

int32_t b200zk_host_alloc(void** out, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (!out) return fail(B200ZK_ERR_INVALID_ARG, "null out pointer");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return B200ZK_OK;
}
int32_t b200zk_host_free(void* p) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (p) CU(cudaFreeHost(p));
    return B200ZK_OK;
}

static int32_t register_common(const uint8_t* d_src, uint64_t n, uint32_t fmt_flags, uint32_t stride, uint64_t* out_handle) {
    BaseTable t;
    t.n = n;
    uint32_t fmt = fmt_flags & 0xffu;
    // Window tables: W rows of n points.  Built for resident SRS tables (they are registered once and
    // committed against many times); skipped on request, for tiny tables, or when HBM is short.
    bool want_rows = !(fmt_flags & B200ZK_BASES_NO_WINDOW_TABLES) && !g.tune_no_tables && n >= 1;
    uint32_t c = 0, W = 1;
    if (want_rows) {
        c = (g.tune_c >= 2 && g.tune_c <= 24) ? g.tune_c : std::max(msm_choose_window(n, 1, true), 8u);   // c >= 8: at most 32 rows
        W = (256 + c - 1) / c;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if ((double)W * n * 96.0 > 0.45 * (double)free_b || (double)W * n >= 2147483648.0 || W > (uint32_t)MSM_MAX_ROWS) { W = 1; c = 0; }
    }
    CU(cudaMalloc(&t.d, std::max<uint64_t>(n, 1) * 96 * W));
    int32_t rc = n ? ingest_bases(d_src, n, fmt, stride, t.d, g.stream) : B200ZK_OK;
    if (rc != B200ZK_OK) { cudaFree(t.d); return rc; }
    if (W > 1) {
        g1_window_tables_kernel<<<(unsigned)((n + 127) / 128), 128, 0, g.stream>>>(t.d, n, c, W);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
        if (e != cudaSuccess) { cudaFree(t.d); return fail(B200ZK_ERR_CUDA, std::string("window tables: ") + cudaGetErrorString(e)); }
        t.rows = W;
        t.c = c;
    }
    uint64_t h = g.next_handle++;
    g.tables[h] = t;
    *out_handle = h;
    return B200ZK_OK;
}

int32_t b200zk_bases_register(const uint8_t* g1_affine, uint64_t n, uint32_t fmt, uint32_t stride_bytes,
                              uint64_t* out_handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (!out_handle || (!g1_affine && n)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if ((fmt & 0xffu) > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown point format");
    uint32_t stride = stride_bytes ? stride_bytes : 96;
    if (stride < 96 || (stride & 3)) return fail(B200ZK_ERR_INVALID_ARG, "stride must be >= 96 and a multiple of 4");
    size_t bytes = n ? (size_t)(n - 1) * stride + 96 : 0;
    TRY(g.stage.ensure(bytes + 16));
    if (bytes) CU(cudaMemcpyAsync(g.stage.p, g1_affine, bytes, cudaMemcpyHostToDevice, g.stream));
    return register_common(g.stage.as<uint8_t>(), n, fmt, stride, out_handle);
}

int32_t b200zk_bases_register_dev(const void* d_g1_affine, uint64_t n, uint32_t fmt, uint32_t stride_bytes,
                                  uint64_t* out_handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (!out_handle || (!d_g1_affine && n)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if ((fmt & 0xffu) > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown point format");
    uint32_t stride = stride_bytes ? stride_bytes : 96;
    if (stride < 96 || (stride & 3)) return fail(B200ZK_ERR_INVALID_ARG, "stride must be >= 96 and a multiple of 4");
    CU(cudaDeviceSynchronize());  // the source may have been produced on another stream
    return register_common(reinterpret_cast<const uint8_t*>(d_g1_affine), n, fmt, stride, out_handle);
}

int32_t b200zk_bases_release(uint64_t handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    auto it = g.tables.find(handle);
    if (it == g.tables.end()) return fail(B200ZK_ERR_BAD_HANDLE, "unknown base-table handle");
    CU(cudaDeviceSynchronize());
    cudaFree(it->second.d);
    g.tables.erase(it);
    return B200ZK_OK;
}

int32_t b200zk_bases_read(uint64_t handle, uint64_t start, uint64_t n, uint8_t* out_affine) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    const uint32_t* d = nullptr;
    TRY(lookup_bases(handle, start, n, &d));
    if (n == 0) return B200ZK_OK;
    if (!out_affine) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    TRY(g.stage.ensure(n * 96));
    LAUNCH(g1_export_kernel, (unsigned)((n + 127) / 128), 128, 0, g.stream, d, n, g.stage.as<uint32_t>());
    CU(cudaMemcpyAsync(out_affine, g.stage.p, n * 96, cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    return B200ZK_OK;
}

int32_t b200zk_msm_g1_batch(uint64_t bases, uint64_t offset, const uint8_t* scalars, uint64_t n, uint32_t batch,
                            uint32_t scalar_fmt, uint8_t* out_affine) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (!out_affine || (!scalars && n && batch)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (scalar_fmt > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown scalar format");
    if (batch == 0) return B200ZK_OK;
    const uint32_t* d_bases = nullptr;
    const BaseTable* tab = nullptr;
    TRY(lookup_bases(bases, offset, n, &d_bases, &tab));
    size_t bytes = (size_t)n * batch * 32;
    TRY(g.scalars.ensure(bytes + 16));
    TRY(g.out_canon.ensure((size_t)batch * 96));
    static int64_t chunk_min = -1;   // a single MSM of at least this many points streams its scalars in two pieces
    if (chunk_min < 0) { const char* v = getenv("B200ZK_MSM_CHUNK_MIN"); chunk_min = v ? atoll(v) : (1ll << 23); }
    // batches of columns: worth two passes only when the transfer is long (measured: 75 MB at k = 17 loses 0.7 ms, 288 MB at
    // k = 19 gains 3.3 ms)
    static size_t batch_stream_min = 0;
    if (!batch_stream_min) { const char* v = getenv("B200ZK_BATCH_STREAM_MIN_BYTES"); batch_stream_min = v ? (size_t)atoll(v) : ((size_t)128 << 20); if (!batch_stream_min) batch_stream_min = 1; }
    if (batch == 1 && chunk_min > 0 && n >= (uint64_t)chunk_min) {
        // The second piece of the scalars crosses PCIe while the first is sorted and accumulated: the first piece's
        // buckets are g.buckets, the second accumulates into g.buckets2 and is folded in before the tail.
        if (!g.copy_stream) {
            CU(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&g.copy_ev[0], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&g.copy_ev[1], cudaEventDisableTiming));
        }
        // first piece = 1/8 of the points: its copy (1.2 ms at 2^24) is the only exposed transfer, and its sort + accumulate
        // (9 ms) cover the copy of the other 7/8 (8.5 ms).  Measured at 2^24: 84.9 ms unchunked, 83.1 ms with halves.
        // From pageable memory (a plain Rust Vec) the copy runs at ~11 GB/s instead of ~55: the balance point
        // copy(rest) = compute(first) moves from 1/8 to 3/8 (measured at 2^24, pageable: 122.8 ms unchunked, 118.7 with 1/8).
        cudaPointerAttributes pa{};
        bool pinned = cudaPointerGetAttributes(&pa, scalars) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const uint64_t nA = ((pinned ? n / 8 : 3 * (n / 8)) + 255) & ~(uint64_t)255, nB = n - nA;
        uint8_t* d_sc = g.scalars.as<uint8_t>();
        // (the first piece's kernels are enqueued before the second copy is issued: from pageable host memory a copy
        // blocks the calling thread, and the device must already have work by then)
        MsmChunk ca{n, 0, true, false}, cb{n, nA, false, true};
        CU(cudaMemcpyAsync(d_sc, scalars, nA * 32, cudaMemcpyHostToDevice, g.copy_stream));
        CU(cudaEventRecord(g.copy_ev[0], g.copy_stream));
        CU(cudaStreamWaitEvent(g.stream, g.copy_ev[0], 0));
        TRY(msm_run(tab, d_bases, g.scalars.as<uint32_t>(), nA, 1, scalar_fmt, nullptr, nullptr, g.stream, nullptr, &ca));
        CU(cudaMemcpyAsync(d_sc + nA * 32, scalars + nA * 32, nB * 32, cudaMemcpyHostToDevice, g.copy_stream));
        CU(cudaEventRecord(g.copy_ev[1], g.copy_stream));
        CU(cudaStreamWaitEvent(g.stream, g.copy_ev[1], 0));
        TRY(msm_run(tab, d_bases, g.scalars.as<uint32_t>() + 8 * nA, nB, 1, scalar_fmt, nullptr, g.out_canon.as<uint32_t>(), g.stream,
                    nullptr, &cb));
    } else if (batch >= 4 && chunk_min > 0 && bytes >= batch_stream_min) {
        // The prover's pattern (all columns of a phase in one call): the columns are independent, so the first eighth of
        // them goes up and starts computing while the others cross PCIe; no merge is needed.
        if (!g.copy_stream) {
            CU(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&g.copy_ev[0], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&g.copy_ev[1], cudaEventDisableTiming));
        }
        const uint32_t bA = std::max(1u, batch / 8), bB = batch - bA;
        const size_t col = (size_t)n * 32;
        uint8_t* d_sc = g.scalars.as<uint8_t>();
        CU(cudaMemcpyAsync(d_sc, scalars, bA * col, cudaMemcpyHostToDevice, g.copy_stream));
        CU(cudaEventRecord(g.copy_ev[0], g.copy_stream));
        CU(cudaStreamWaitEvent(g.stream, g.copy_ev[0], 0));
        TRY(msm_run(tab, d_bases, g.scalars.as<uint32_t>(), n, bA, scalar_fmt, nullptr, g.out_canon.as<uint32_t>(), g.stream));
        CU(cudaMemcpyAsync(d_sc + bA * col, scalars + bA * col, bB * col, cudaMemcpyHostToDevice, g.copy_stream));
        CU(cudaEventRecord(g.copy_ev[1], g.copy_stream));
        CU(cudaStreamWaitEvent(g.stream, g.copy_ev[1], 0));
        TRY(msm_run(tab, d_bases, g.scalars.as<uint32_t>() + 8 * (size_t)n * bA, n, bB, scalar_fmt, nullptr,
                    g.out_canon.as<uint32_t>() + 24 * (size_t)bA, g.stream));
    } else {
        if (bytes) CU(cudaMemcpyAsync(g.scalars.p, scalars, bytes, cudaMemcpyHostToDevice, g.stream));
        TRY(msm_run(tab, d_bases, g.scalars.as<uint32_t>(), n, batch, scalar_fmt, nullptr, g.out_canon.as<uint32_t>(), g.stream));
    }
    CU(cudaMemcpyAsync(out_affine, g.out_canon.p, (size_t)batch * 96, cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    return B200ZK_OK;
}

int32_t b200zk_msm_g1(uint64_t bases, uint64_t offset, const uint8_t* scalars, uint64_t n, uint32_t scalar_fmt,
                      uint8_t out_affine[96]) {
    return b200zk_msm_g1_batch(bases, offset, scalars, n, 1, scalar_fmt, out_affine);
}

int32_t b200zk_msm_g1_adhoc(const uint8_t* g1_affine, uint32_t point_fmt, const uint8_t* scalars, uint32_t scalar_fmt,
                            uint64_t n, uint8_t out_affine[96]) {
    // One pass on the context stream with cached workspaces: no table registration, no allocation, one
    // synchronisation at the end (the verifier calls this once per proof or per batch).
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (!out_affine || ((!g1_affine || !scalars) && n)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (point_fmt > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown point format");
    if (scalar_fmt > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown scalar format");
    if (n == 0) { memset(out_affine, 0, 96); return B200ZK_OK; }
    cudaStream_t s = g.stream;
    TRY(g.stage.ensure(n * 96 + 16));
    TRY(g.adhoc.ensure(n * 96));
    TRY(g.scalars.ensure(n * 32 + 16));
    TRY(g.out_canon.ensure(96));
    TRY(g.flag.ensure(4));
    CU(cudaMemcpyAsync(g.stage.p, g1_affine, n * 96, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(g.scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(g.flag.p, 0, 4, s));
    LAUNCH(g1_ingest_kernel, (unsigned)((n + 127) / 128), 128, 0, s, (const uint8_t*)g.stage.as<uint8_t>(), n, 96u,
           point_fmt == B200ZK_FMT_MONT ? 1u : 0u, g.adhoc.as<uint32_t>(), g.flag.as<uint32_t>());
    BaseTable t;
    t.d = g.adhoc.as<uint32_t>();
    t.n = n;
    TRY(msm_run(&t, t.d, g.scalars.as<uint32_t>(), n, 1, scalar_fmt, nullptr, g.out_canon.as<uint32_t>(), s));
    uint32_t bad = 0;
    CU(cudaMemcpyAsync(out_affine, g.out_canon.p, 96, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&bad, g.flag.p, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (bad) { memset(out_affine, 0, 96); return fail(B200ZK_ERR_BAD_POINT, "a base point is not a canonical point on the curve"); }
    return B200ZK_OK;
}

int32_t b200zk_msm_g1_dev(uint64_t bases, uint64_t offset, const void* d_scalars, uint64_t n, uint32_t batch,
                          uint32_t scalar_fmt, void* d_out_mont, void* d_out_canon, void* stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if ((!d_scalars && n && batch) || (!d_out_mont && !d_out_canon)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (scalar_fmt > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown scalar format");
    if (((uintptr_t)d_scalars | (uintptr_t)d_out_mont | (uintptr_t)d_out_canon) & 15)
        return fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    const uint32_t* d_bases = nullptr;
    const BaseTable* tab = nullptr;
    TRY(lookup_bases(bases, offset, n, &d_bases, &tab));
    return msm_run(tab, d_bases, reinterpret_cast<const uint32_t*>(d_scalars), n, batch, scalar_fmt,
                   reinterpret_cast<uint32_t*>(d_out_mont), reinterpret_cast<uint32_t*>(d_out_canon),
                   reinterpret_cast<cudaStream_t>(stream));
}

int32_t b200zk_msm_g1_partial_dev(uint64_t bases, uint64_t offset, const void* d_scalars, uint64_t n, uint32_t scalar_fmt,
                                  void* d_out_xyzz, void* stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if ((!d_scalars && n) || !d_out_xyzz) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (scalar_fmt > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown scalar format");
    if (((uintptr_t)d_scalars | (uintptr_t)d_out_xyzz) & 15) return fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    const uint32_t* d_bases = nullptr;
    const BaseTable* tab = nullptr;
    TRY(lookup_bases(bases, offset, n, &d_bases, &tab));
    if (n == 0) { CU(cudaMemsetAsync(d_out_xyzz, 0, 192, reinterpret_cast<cudaStream_t>(stream))); return B200ZK_OK; }
    return msm_run(tab, d_bases, reinterpret_cast<const uint32_t*>(d_scalars), n, 1, scalar_fmt, nullptr, nullptr,
                   reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<uint32_t*>(d_out_xyzz));
}

int32_t b200zk_g1_sum_partials_dev(const void* d_partials_xyzz, uint32_t n, void* d_out_mont, void* d_out_canon, void* stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if ((!d_partials_xyzz && n) || (!d_out_mont && !d_out_canon)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    LAUNCH(g1_sum_xyzz_kernel, 1, 32, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const uint32_t*>(d_partials_xyzz),
           n, reinterpret_cast<uint32_t*>(d_out_mont), reinterpret_cast<uint32_t*>(d_out_canon));
    return B200ZK_OK;
}

int32_t b200zk_g1_sum_dev(const void* d_points_mont, uint32_t n, void* d_out_mont, void* d_out_canon, void* stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if ((!d_points_mont && n) || (!d_out_mont && !d_out_canon)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    LAUNCH(g1_sum_kernel, 1, 32, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const uint32_t*>(d_points_mont), n,
           reinterpret_cast<uint32_t*>(d_out_mont), reinterpret_cast<uint32_t*>(d_out_canon));
    return B200ZK_OK;
}

int32_t b200zk_ntt_fr_dev(void* d_data, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                          const uint8_t coset_shift[32], void* stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (!d_data || !omega) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if ((uintptr_t)d_data & 15) return fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    return ntt_run(reinterpret_cast<uint32_t*>(d_data), batch, log_n, omega, flags, coset_shift,
                   reinterpret_cast<cudaStream_t>(stream));
}

int32_t b200zk_ntt_fr_batch(uint8_t* data, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                            const uint8_t coset_shift[32]) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (!data || !omega) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (log_n > 32) return fail(B200ZK_ERR_INVALID_ARG, "ntt: log_n exceeds the 2-adicity of Fr");
    if (batch == 0) return B200ZK_OK;
    size_t bytes = ((size_t)batch << log_n) * 32;
    TRY(g.ntt_data.ensure(bytes));
    static int64_t pipe_min = -1;   // batches with at least this many bytes are pipelined group by group
    if (pipe_min < 0) { const char* v = getenv("B200ZK_NTT_PIPE_MIN_BYTES"); pipe_min = v ? atoll(v) : (32ll << 20); }
    if (batch >= 4 && pipe_min > 0 && bytes >= (size_t)pipe_min) {
        // A transform is PCIe-bound end to end (2^22: 2.4 ms up, 0.94 ms of kernels, 2.4 ms down), and the link is full
        // duplex: group g+1 goes up and group g-1 comes down while group g is transformed.
        if (!g.copy_stream) {
            CU(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&g.copy_ev[0], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&g.copy_ev[1], cudaEventDisableTiming));
        }
        if (!g.down_stream) CU(cudaStreamCreateWithFlags(&g.down_stream, cudaStreamNonBlocking));
        const uint32_t groups = std::min<uint32_t>(8, batch / 2);
        const size_t poly = ((size_t)1 << log_n) * 32;
        std::vector<cudaEvent_t>& up = g.pipe_up;
        std::vector<cudaEvent_t>& done = g.pipe_done;
        while (up.size() < groups) {
            cudaEvent_t e1, e2;
            CU(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            up.push_back(e1);
            done.push_back(e2);
        }
        uint8_t* d = g.ntt_data.as<uint8_t>();
        for (uint32_t k = 0; k < groups; k++) {
            const uint32_t b0 = (uint32_t)((uint64_t)batch * k / groups), b1 = (uint32_t)((uint64_t)batch * (k + 1) / groups);
            const size_t off = b0 * poly, len = (size_t)(b1 - b0) * poly;
            CU(cudaMemcpyAsync(d + off, data + off, len, cudaMemcpyHostToDevice, g.copy_stream));
            CU(cudaEventRecord(up[k], g.copy_stream));
            CU(cudaStreamWaitEvent(g.stream, up[k], 0));
            TRY(ntt_run(reinterpret_cast<uint32_t*>(d + off), b1 - b0, log_n, omega, flags, coset_shift, g.stream));
            CU(cudaEventRecord(done[k], g.stream));
            CU(cudaStreamWaitEvent(g.down_stream, done[k], 0));
            CU(cudaMemcpyAsync(data + off, d + off, len, cudaMemcpyDeviceToHost, g.down_stream));
        }
        CU(cudaStreamSynchronize(g.down_stream));
        CU(cudaStreamSynchronize(g.stream));
        return B200ZK_OK;
    }
    CU(cudaMemcpyAsync(g.ntt_data.p, data, bytes, cudaMemcpyHostToDevice, g.stream));
    TRY(ntt_run(g.ntt_data.as<uint32_t>(), batch, log_n, omega, flags, coset_shift, g.stream));
    CU(cudaMemcpyAsync(data, g.ntt_data.p, bytes, cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    return B200ZK_OK;
}

int32_t b200zk_ntt_fr(uint8_t* data, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                      const uint8_t coset_shift[32]) {
    return b200zk_ntt_fr_batch(data, 1, log_n, omega, flags, coset_shift);
}

// p as big-endian bytes, for the y > p - y test of the compressed encoding
static const uint8_t FP_P_BE[48] = {0x1a, 0x01, 0x11, 0xea, 0x39, 0x7f, 0xe6, 0x9a, 0x4b, 0x1b, 0xa7, 0xb6, 0x43, 0x4b, 0xac, 0xd7,
                                    0x64, 0x77, 0x4b, 0x84, 0xf3, 0x85, 0x12, 0xbf, 0x67, 0x30, 0xd2, 0xa0, 0xf6, 0xb0, 0xf6, 0x24,
                                    0x1e, 0xab, 0xff, 0xfe, 0xb1, 0x53, 0xff, 0xff, 0xb9, 0xfe, 0xff, 0xff, 0xff, 0xff, 0xaa, 0xab};

int32_t b200zk_g1_compress(const uint8_t affine[96], uint8_t out[48]) {
    if (!affine || !out) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    bool zero = true;
    for (int i = 0; i < 96; i++) if (affine[i]) { zero = false; break; }
    if (zero) { memset(out, 0, 48); out[0] = 0xC0; return B200ZK_OK; }
    uint8_t y_be[48], twoy[49];
    for (int i = 0; i < 48; i++) { out[i] = affine[47 - i]; y_be[i] = affine[48 + 47 - i]; }
    if (out[0] & 0xE0) return fail(B200ZK_ERR_BAD_POINT, "x coordinate out of range");
    // y is "larger" iff y > p - y iff 2y > p
    unsigned carry = 0;
    for (int i = 47; i >= 0; i--) { unsigned v = 2u * y_be[i] + carry; twoy[i + 1] = (uint8_t)v; carry = v >> 8; }
    twoy[0] = (uint8_t)carry;
    bool larger = twoy[0] != 0 || memcmp(twoy + 1, FP_P_BE, 48) > 0;
    out[0] |= 0x80 | (larger ? 0x20 : 0);
    return B200ZK_OK;
}

int32_t b200zk_g1_synth_bases_dev(uint64_t seed, uint64_t start, uint64_t n, void* d_out_mont, void* stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (!d_out_mont && n) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (!g.fixed_table) {
        uint32_t* d_gen = nullptr;
        TRY(get_generator_dev(&d_gen, s));
        CU(cudaMalloc(&g.fixed_table, 8 * 256 * 96));
        LAUNCH(g1_fixed_table_kernel, 16, 128, 0, s, (const uint32_t*)d_gen, g.fixed_table);
    }
    if (n) LAUNCH(g1_synth_bases_kernel, (unsigned)((n + 127) / 128), 128, 0, s, (const uint32_t*)g.fixed_table, seed, start, n,
                  reinterpret_cast<uint32_t*>(d_out_mont));
    return B200ZK_OK;
}

int32_t b200zk_selftest_field(uint32_t field, uint32_t op, const uint8_t* a, const uint8_t* b, uint8_t* out, uint64_t count) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (field > 1 || op > 3 || !a || !out || (op < 3 && !b)) return fail(B200ZK_ERR_INVALID_ARG, "bad selftest arguments");
    if (count == 0) return B200ZK_OK;
    size_t esz = field ? 48 : 32, bytes = esz * count;
    TRY(g.stage.ensure(3 * bytes + 64));
    uint8_t* da = g.stage.as<uint8_t>();
    uint8_t* db = da + bytes;
    uint8_t* dout = db + bytes;
    CU(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, g.stream));
    if (b) CU(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, g.stream));
    unsigned grid = (unsigned)((count + 127) / 128);
    if (field == 0)
        LAUNCH(selftest_kernel<FrParams>, grid, 128, 0, g.stream, (const uint32_t*)da, b ? (const uint32_t*)db : nullptr,
               (uint32_t*)dout, count, op);
    else
        LAUNCH(selftest_kernel<FpParams>, grid, 128, 0, g.stream, (const uint32_t*)da, b ? (const uint32_t*)db : nullptr,
               (uint32_t*)dout, count, op);
    CU(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    return B200ZK_OK;
}

int32_t b200zk_microbench(uint32_t kind, uint32_t iters, double* out_ops_per_s, double* out_ms) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (kind > 11 || !out_ops_per_s) return fail(B200ZK_ERR_INVALID_ARG, "bad microbench arguments");
    int sms = g.prop.multiProcessorCount;
    unsigned threads = (kind == 3) ? 128 : 256;
    unsigned blocks = (unsigned)sms * ((kind == 3) ? 3 : ((kind == 2) ? 4 : 8));
    TRY(g.stage.ensure((size_t)blocks * threads * 8 + 64));
    uint32_t* d_gen = nullptr;
    if (kind == 3) TRY(get_generator_dev(&d_gen, g.stream));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double per_thread = 0;
    for (int rep = 0; rep < 2; rep++) {  // rep 0 warms up
        CU(cudaEventRecord(e0, g.stream));
        switch (kind) {
            case 0:
                LAUNCH(mb_imad_wide_kernel, blocks, threads, 0, g.stream, g.stage.as<uint64_t>(), iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
            case 1:
                LAUNCH(mb_imad_pair_kernel, blocks, threads, 0, g.stream, g.stage.as<uint64_t>(), iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
            case 2:
                LAUNCH(mb_femul_kernel<FpParams>, blocks, threads, 0, g.stream, g.stage.as<uint32_t>(), iters);
                per_thread = 2.0 * iters;
                break;
            case 3:
                LAUNCH(mb_madd_kernel, blocks, threads, 0, g.stream, (const uint32_t*)d_gen, g.stage.as<uint32_t>(), iters);
                per_thread = 1.0 * iters;
                break;
            case 4:
                LAUNCH(mb_femul_kernel<FrParams>, blocks, threads, 0, g.stream, g.stage.as<uint32_t>(), iters);
                per_thread = 2.0 * iters;
                break;
            case 5:
                LAUNCH(mb_imad_chain_kernel, blocks, threads, 0, g.stream, g.stage.as<uint64_t>(), iters, 12345u, 67890u);
                per_thread = 24.0 * iters;
                break;
            case 6:
                LAUNCH(mb_dfma_kernel, blocks, threads, 0, g.stream, g.stage.as<uint64_t>(), iters, 1.000001, 0.999999);
                per_thread = 32.0 * iters;
                break;
            case 7:
                LAUNCH(mb_imad_cout_kernel, blocks, threads, 0, g.stream, g.stage.as<uint64_t>(), iters, 12345u, 67890u);
                per_thread = 24.0 * iters;
                break;
            case 9:
                LAUNCH(mb_imad_parts_kernel<0>, blocks, threads, 0, g.stream, g.stage.as<uint64_t>(), iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
            case 10:
                LAUNCH(mb_imad_parts_kernel<1>, blocks, threads, 0, g.stream, g.stage.as<uint64_t>(), iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
            case 11:
                LAUNCH(mb_imad_parts_kernel<2>, blocks, threads, 0, g.stream, g.stage.as<uint64_t>(), iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
            default:
                LAUNCH(mb_imad_alu_kernel, blocks, threads, 0, g.stream, g.stage.as<uint64_t>(), iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
        }
        CU(cudaEventRecord(e1, g.stream));
        CU(cudaEventSynchronize(e1));
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out_ops_per_s = per_thread * (double)blocks * threads / (ms * 1e-3);
    if (out_ms) *out_ms = ms;
    return B200ZK_OK;
}

uint64_t b200zk_launch_count(void) { return g_launches.load(); }

int32_t b200zk_set_profiling(uint32_t enable) {
    std::lock_guard<std::mutex> lk(g_mu);
    g.profiling = enable != 0;
    g.ev_count = 0;
    return B200ZK_OK;
}

int32_t b200zk_get_profile(uint32_t* kind, double* phase_ms, uint32_t cap, uint32_t* n_phases, uint32_t* msm_window_bits,
                           uint32_t* msm_windows) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(need_init());
    if (!kind || !phase_ms || !n_phases) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    *kind = (uint32_t)g.ev_kind;
    *n_phases = 0;
    if (msm_window_bits) *msm_window_bits = g.last_plan.c;
    if (msm_windows) *msm_windows = g.last_plan.W;
    if (g.ev_count < 2) return B200ZK_OK;
    CU(cudaEventSynchronize(g.ev[g.ev_count - 1]));
    for (int i = 0; i + 1 < g.ev_count && (uint32_t)i < cap; i++) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, g.ev[i], g.ev[i + 1]));
        phase_ms[i] = ms;
        *n_phases = (uint32_t)i + 1;
    }
    return B200ZK_OK;
}

int32_t b200zk_set_msm_tuning(uint32_t window_bits, uint32_t smax) {
    std::lock_guard<std::mutex> lk(g_mu);
    g.tune_c = window_bits & 0xffu;
    g.tune_variant = ((window_bits >> 8) & 0x7fu) ? ((window_bits >> 8) & 0x7fu) - 1 : 6;  // bits 8..14: 1 + accumulate-kernel variant
    g.tune_no_tables = (window_bits >> 15) & 1u;    // bit 15: do not build window tables at registration
    g.tune_smax = smax;
    return B200ZK_OK;
}

}  // extern "C"
