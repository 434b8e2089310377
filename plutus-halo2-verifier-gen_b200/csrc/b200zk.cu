// b200zk.cu -- device contexts, workspaces and the C ABI of include/b200zk.h for the hot path:
// base-table residency, the G1 MSM and the Fr NTT.
//
// One process drives every GPU bound by b200zk_init_devices (the reference prover is a single process:
// /root/reference/examples/simple_mul.rs:39-141 calls commit at :62,72).  Per GPU there is one context with
// NSLOT workspace slots; a slot is one in-flight MSM or NTT with its own stream, so two host threads (rayon
// callers) overlap on one GPU -- the memory-bound sort of one call runs under the integer-bound bucket
// accumulation of the other -- and no lock is held while a call waits for the device.  A base table is
// either replicated on every GPU (small SRS: the columns of a batch are dealt out, no exchange) or sharded by
// point range (large SRS: every GPU reduces its slice to one XYZZ partial, the partials meet on the first GPU
// through peer stores fused into the last kernel of the MSM, msm.cuh).  Host-side fan-out uses one worker
// thread per GPU so that copies from pageable memory and kernel launches of different GPUs do not serialise.
// There is no CPU fallback anywhere in this file: without a device the calls fail.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "b200zk.h"
#include "ctx.hpp"
#include "field.cuh"
#include "g1.cuh"
#include "msm.cuh"
#include "ntt.cuh"

using namespace b200zk;
using b200zk_ctx::DeviceScope;

namespace {

thread_local std::string t_err;
thread_local int t_dev_index = 0;          // b200zk_set_device
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
struct PinBuf {   // pinned host staging for small results: a D2H copy into it never blocks the enqueuing thread
    void* p = nullptr;
    size_t cap = 0;
    int32_t ensure(size_t bytes) {
        if (bytes <= cap) return B200ZK_OK;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = std::max<size_t>(bytes * 2, 4096);
        CU(cudaHostAlloc(&p, want, cudaHostAllocPortable));
        cap = want;
        return B200ZK_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct BaseTable {
    uint32_t* d = nullptr;  // packed Montgomery affine, 24 limbs per point; `rows` rows of n points
    uint64_t n = 0;
    uint32_t rows = 1;      // > 1: window tables, row w = 2^(c*w) * row 0
    uint32_t c = 0;         // window bits the rows were built for
};
// A registered table as the process sees it: replicated (every shard holds all n points) or partitioned by point range.
struct TableShard {
    int dev = 0;            // index into the bound devices
    uint64_t start = 0, n = 0;
    BaseTable t;
};
struct TableSet {
    uint64_t n = 0;
    bool replicated = false;
    std::vector<TableShard> shards;
};

struct NttPlan {
    uint32_t log_n = 0, npass = 0;
    uint32_t logb = NTT_LOGB;  // log2 of the CTA tile: 11, or 12 for the two-pass plans of 2^23 / 2^24
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

// tables of one GPU's share of a multi-GPU transform (n = R * C, GPU g owns columns [g C/G, (g+1) C/G) in the first step)
struct ShardPlan {
    uint32_t log_r = 0, log_c = 0;
    uint32_t* tw_local = nullptr;   // omega_R^e, e < R/2
    uint32_t* tw1 = nullptr;        // [u][k]: (1/n) * omega^((g C/G + u) k)
};
struct ShardKey {
    uint32_t log_n, inverse, parts, part;
    uint8_t omega[32];
    bool operator<(const ShardKey& o) const {
        if (log_n != o.log_n) return log_n < o.log_n;
        if (inverse != o.inverse) return inverse < o.inverse;
        if (parts != o.parts) return parts < o.parts;
        if (part != o.part) return part < o.part;
        return memcmp(omega, o.omega, 32) < 0;
    }
};
struct ShardCosetKey {
    uint32_t log_n, parts, part, out;
    uint8_t shift[32];
    bool operator<(const ShardCosetKey& o) const {
        if (log_n != o.log_n) return log_n < o.log_n;
        if (parts != o.parts) return parts < o.parts;
        if (part != o.part) return part < o.part;
        if (out != o.out) return out < o.out;
        return memcmp(shift, o.shift, 32) < 0;
    }
};

// one in-flight MSM or NTT: stream, scratch, staging
constexpr int NSLOT = 2;
struct Slot {
    int id = 0;
    bool busy = false;
    cudaStream_t stream = nullptr, copy_stream = nullptr, down_stream = nullptr;
    cudaEvent_t copy_ev[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> pipe_up, pipe_done;
    // MSM workspace
    DevBuf scalars, counts, offsets, cursor, ntask, task_off, entries, task_bucket, task_start, task_len, buckets, partials, len_hist,
        len_off, order, heavy, heavy_items, adhoc, buckets2, aff_a, aff_b, aff_pre, redS[2], redA[2], scan_tmp[2], out_mont, out_canon,
        stage, flag;
    // NTT workspace
    DevBuf ntt_data, ntt_tmp[2], ntt_b;
    cudaEvent_t xdev_ev = nullptr;       // "my column pass has stored everything into its peers" (multi-GPU transform)
    // batch pipeline: the short, latency-bound phases of an MSM (sort, tail) run on a high-priority stream, the long
    // integer-bound accumulation on a low-priority one, so that the sort / tail of one half of a batch are dispatched into
    // the accumulation of the other half instead of waiting behind it
    cudaStream_t hi = nullptr, lo = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_sorted = nullptr, ev_acc = nullptr, ev_done = nullptr;
    PinBuf h_out;
    // a "_dev" caller's stream is ordered against the previous user of the slot's scratch
    cudaEvent_t ws_event = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_used = false;
    // phase timing (b200zk_set_profiling): events recorded on the launching stream
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int ev_count = 0, ev_kind = 0;   // kind 1 = msm, 2 = ntt
    MsmPlan last_plan{};
    std::vector<DevBuf*> all() {
        return {&scalars, &counts, &offsets, &cursor, &ntask, &task_off, &entries, &task_bucket, &task_start, &task_len, &buckets,
                &partials, &len_hist, &len_off, &order, &heavy, &heavy_items, &adhoc, &buckets2, &aff_a, &aff_b, &aff_pre, &redS[0],
                &redS[1], &redA[0], &redA[1], &scan_tmp[0], &scan_tmp[1], &out_mont, &out_canon, &stage, &flag, &ntt_data, &ntt_tmp[0],
                &ntt_tmp[1], &ntt_b};
    }
};

// host-side fan-out: one worker thread per bound GPU (its CUDA device is set once)
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    bool stop = false;
    void start(int ordinal) {
        th = std::thread([this, ordinal] {
            cudaSetDevice(ordinal);
            for (;;) {
                std::function<void()> job;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [this] { return stop || !q.empty(); });
                    if (q.empty()) return;
                    job = std::move(q.front());
                    q.pop_front();
                }
                job();
            }
        });
    }
    void post(std::function<void()> fn) {
        {
            std::lock_guard<std::mutex> lk(mu);
            q.push_back(std::move(fn));
        }
        cv.notify_one();
    }
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_one();
        if (th.joinable()) th.join();
    }
};

}  // namespace

namespace b200zk_ctx {
struct Dev {
    int ordinal = -1, index = 0;
    cudaDeviceProp prop{};
    std::mutex mu;                    // caches below, slot bookkeeping
    std::mutex gen_mu;                // d_gen only
    std::condition_variable cv;       // a slot became free
    Slot slots[NSLOT];
    cudaStream_t stream = nullptr;    // registration, table read-back, ext-TU host entry points
    std::map<NttKey, NttPlan> ntt_plans;
    std::map<CosetKey, uint32_t*> coset_tables;
    std::map<ShardKey, ShardPlan> shard_plans;        // multi-GPU transforms: this GPU's tables
    std::map<ShardCosetKey, uint32_t*> shard_cosets;
    uint32_t* fixed_table = nullptr;  // 8 x 256 multiples of G for the synthetic-base generator
    uint32_t* d_gen = nullptr;        // generator of G1, Montgomery affine
    bool ntt_attr_set = false;
    DevBuf reg_stage, reg_flag;
    cudaEvent_t ws_event = nullptr;   // ordering of the ext-TU scratch
    cudaStream_t ws_stream = nullptr;
    bool ws_used = false;
    Worker worker;
};
}  // namespace b200zk_ctx
using Ctx = b200zk_ctx::Dev;

namespace {

std::mutex g_mu;                                   // init / shutdown / tuning / profiling selection
// bound devices, in b200zk_init_devices order.  Heap-allocated and never destroyed: a process may exit without calling
// b200zk_shutdown, and static destruction must not touch worker threads or a CUDA runtime that is already gone.
std::vector<std::unique_ptr<Ctx>>& g_devs = *new std::vector<std::unique_ptr<Ctx>>();
std::mutex g_tab_mu;
std::map<uint64_t, std::shared_ptr<TableSet>>& g_tables = *new std::map<uint64_t, std::shared_ptr<TableSet>>();
uint64_t g_next_handle = 1;
std::vector<void (*)()> g_hooks;
// tuning overrides (b200zk_set_msm_tuning) and environment switches, read once
uint32_t g_tune_c = 0, g_tune_smax = 0, g_tune_variant = 6, g_tune_no_tables = 0;
bool g_profiling = false;
Ctx* g_prof_ctx = nullptr;
Slot* g_prof_slot = nullptr;
int g_ntt_variant = 9;            // 9: two CTAs per SM, compile-time twiddle exponents in the lowest sweep, uncorrected differences into the products (default); 7: with corrected differences; 1: without either;
                                  // 0: one CTA per SM; 2..6 experiments.  B200ZK_NTT_VARIANT overrides.
int64_t g_coop8_max = 0;           // top of the bucket tree: CTA-per-group kernel while there are at most this many groups (0 = one per SM)
int g_red_tp = 0;                  // bucket-tree levels with many groups: throughput build with this many CTAs of 128 threads per SM (0 = off)
int64_t g_red_tp_min = 1ll << 16;  // ... from this many groups
int64_t g_chunk_min = 1ll << 20;  // a single host-buffer MSM of at least this many points streams its scalars in two pieces
                                  // (measured end to end, pinned scalars: 2^21 12.38 -> 11.72 ms, 2^22 23.02 -> 21.33 ms; this is also what
                                  // each GPU of a sharded 2^24 MSM does with its 2^21..2^23-point slice)
size_t g_batch_stream_min = (size_t)128 << 20;
int64_t g_ntt_pipe_min = 32ll << 20;
int g_scatter_passes_env = 0;
uint64_t g_replicate_max_bytes = (uint64_t)2 << 30;   // tables up to this size (with window rows) are replicated on every GPU
uint64_t g_range_split_min = 1ull << 16;              // single MSMs with at least this many points are split by point range over the GPUs

void read_env() {
    if (const char* v = getenv("B200ZK_NTT_VARIANT")) g_ntt_variant = atoi(v);
    if (const char* v = getenv("B200ZK_MSM_CHUNK_MIN")) g_chunk_min = atoll(v);
    if (const char* v = getenv("B200ZK_COOP8_MAX")) g_coop8_max = atoll(v);
    if (const char* v = getenv("B200ZK_RED_TP")) g_red_tp = atoi(v);
    if (const char* v = getenv("B200ZK_RED_TP_MIN")) g_red_tp_min = atoll(v);
    if (const char* v = getenv("B200ZK_BATCH_STREAM_MIN_BYTES")) { g_batch_stream_min = (size_t)atoll(v); if (!g_batch_stream_min) g_batch_stream_min = 1; }
    if (const char* v = getenv("B200ZK_NTT_PIPE_MIN_BYTES")) g_ntt_pipe_min = atoll(v);
    if (const char* v = getenv("B200ZK_SCATTER_PASSES")) g_scatter_passes_env = atoi(v);
    if (const char* v = getenv("B200ZK_REPLICATE_MAX_BYTES")) g_replicate_max_bytes = (uint64_t)atoll(v);
    if (const char* v = getenv("B200ZK_RANGE_SPLIT_MIN")) g_range_split_min = (uint64_t)atoll(v);
}

int32_t need_init() {
    if (g_devs.empty()) return fail(B200ZK_ERR_NOT_INIT, "b200zk_init has not been called (or failed): no CUDA device is bound");
    return B200ZK_OK;
}
int32_t current_ctx(Ctx** out) {
    TRY(need_init());
    int i = t_dev_index;
    if (i < 0 || i >= (int)g_devs.size()) i = 0;
    *out = g_devs[i].get();
    return B200ZK_OK;
}
// the bound device that owns a device pointer
int32_t ctx_for_pointer(const void* p, Ctx** out) {
    TRY(need_init());
    if (g_devs.size() == 1 || !p) return current_ctx(out);
    cudaPointerAttributes pa{};
    if (cudaPointerGetAttributes(&pa, p) != cudaSuccess || (pa.type != cudaMemoryTypeDevice && pa.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return fail(B200ZK_ERR_INVALID_ARG, "not a device pointer");
    }
    for (auto& d : g_devs)
        if (d->ordinal == pa.device) { *out = d.get(); return B200ZK_OK; }
    return fail(B200ZK_ERR_INVALID_ARG, "the pointer lives on a GPU that b200zk_init_devices did not bind");
}

// ---- slots ---------------------------------------------------------------------------------
Slot* acquire_slot(Ctx& c) {
    std::unique_lock<std::mutex> lk(c.mu);
    for (;;) {
        // lowest free slot: a sequential caller keeps reusing slot 0 (warm workspaces, half the memory); the second slot
        // only comes into play when two calls overlap
        for (int k = 0; k < NSLOT; k++) {
            Slot& s = c.slots[k];
            if (!s.busy) {
                s.busy = true;
                return &s;
            }
        }
        c.cv.wait(lk);
    }
}
void release_slot(Ctx& c, Slot* s) {
    {
        std::lock_guard<std::mutex> lk(c.mu);
        s->busy = false;
    }
    c.cv.notify_one();
}
struct SlotLease {
    Ctx* c = nullptr;
    Slot* s = nullptr;
    SlotLease() = default;
    explicit SlotLease(Ctx& ctx) : c(&ctx), s(acquire_slot(ctx)) {}
    SlotLease(SlotLease&& o) noexcept : c(o.c), s(o.s) { o.s = nullptr; }
    ~SlotLease() { if (s) release_slot(*c, s); }
};

int32_t prof_mark(Slot& sl, int idx, cudaStream_t s) {
    if (!g_profiling) return B200ZK_OK;
    if (!sl.ev[idx]) CU(cudaEventCreate(&sl.ev[idx]));
    CU(cudaEventRecord(sl.ev[idx], s));
    sl.ev_count = idx + 1;
    return B200ZK_OK;
}
void prof_select(Ctx& c, Slot& sl, int kind) {
    sl.ev_kind = kind;
    if (!g_profiling) return;
    std::lock_guard<std::mutex> lk(g_mu);
    g_prof_ctx = &c;
    g_prof_slot = &sl;
}

// A slot's scratch is used by one call at a time on the host side (busy flag); on the device side, a call that arrives
// on a different stream than the slot's previous user waits for that user's last kernel.
int32_t ws_enter(Slot& sl, cudaStream_t s) {
    if (!sl.ws_event) CU(cudaEventCreateWithFlags(&sl.ws_event, cudaEventDisableTiming));
    if (sl.ws_used && s != sl.ws_stream) CU(cudaStreamWaitEvent(s, sl.ws_event, 0));
    return B200ZK_OK;
}
int32_t ws_leave(Slot& sl, cudaStream_t s) {
    CU(cudaEventRecord(sl.ws_event, s));
    sl.ws_stream = s;
    sl.ws_used = true;
    return B200ZK_OK;
}
int32_t slot_pipe_streams(Slot& sl) {
    if (sl.hi) return B200ZK_OK;
    int least = 0, greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CU(cudaStreamCreateWithPriority(&sl.hi, cudaStreamNonBlocking, greatest));
    CU(cudaStreamCreateWithPriority(&sl.lo, cudaStreamNonBlocking, least));
    for (cudaEvent_t* e : {&sl.ev_fork, &sl.ev_sorted, &sl.ev_acc, &sl.ev_done}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return B200ZK_OK;
}
Slot* try_acquire_slot(Ctx& c) {
    std::lock_guard<std::mutex> lk(c.mu);
    for (int k = 0; k < NSLOT; k++)
        if (!c.slots[k].busy) { c.slots[k].busy = true; return &c.slots[k]; }
    return nullptr;
}
int32_t slot_streams(Slot& sl) {
    if (!sl.copy_stream) {
        CU(cudaStreamCreateWithFlags(&sl.copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&sl.copy_ev[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&sl.copy_ev[1], cudaEventDisableTiming));
    }
    if (!sl.down_stream) CU(cudaStreamCreateWithFlags(&sl.down_stream, cudaStreamNonBlocking));
    return B200ZK_OK;
}

// ---- multi-device fan-out ------------------------------------------------------------------
struct Job {
    Ctx* c;
    std::function<int32_t()> fn;
};
// Runs every job on its device's worker thread (a single job runs inline with the device made current) and returns the
// first failure, whose text becomes the calling thread's last error.
int32_t run_jobs(std::vector<Job>& jobs) {
    if (jobs.empty()) return B200ZK_OK;
    if (jobs.size() == 1) {
        DeviceScope ds(jobs[0].c->ordinal);
        return jobs[0].fn();
    }
    struct Shared {
        std::mutex mu;
        std::condition_variable cv;
        size_t left;
        std::vector<int32_t> rc;
        std::vector<std::string> err;
    } sh;
    sh.left = jobs.size();
    sh.rc.assign(jobs.size(), B200ZK_OK);
    sh.err.resize(jobs.size());
    for (size_t i = 0; i < jobs.size(); i++) {
        Job* j = &jobs[i];
        jobs[i].c->worker.post([j, i, &sh] {
            int32_t rc = j->fn();
            std::string e = rc != B200ZK_OK ? t_err : std::string();
            std::lock_guard<std::mutex> lk(sh.mu);
            sh.rc[i] = rc;
            sh.err[i] = std::move(e);
            if (--sh.left == 0) sh.cv.notify_all();
        });
    }
    std::unique_lock<std::mutex> lk(sh.mu);
    sh.cv.wait(lk, [&sh] { return sh.left == 0; });
    for (size_t i = 0; i < jobs.size(); i++)
        if (sh.rc[i] != B200ZK_OK) return fail(sh.rc[i], "GPU " + std::to_string(jobs[i].c->ordinal) + ": " + sh.err[i]);
    return B200ZK_OK;
}

// ---- exchange areas ------------------------------------------------------------------------
// The meeting point of the per-GPU partial sums of a point-range sharded MSM (msm_combine_kernel): counters,
// result slots and the gather buffer, in the HBM of a "home" GPU.  In-process areas are reached through peer
// access; areas shared between processes (one process per GPU, dist.py) through a CUDA-IPC mapping.
struct XchgArea {
    uint8_t* base = nullptr;          // device memory on the home GPU (or its IPC mapping)
    uint32_t* host_canon = nullptr;   // mapped pinned host memory (in-process areas): the result lands here directly
    uint32_t n_parts = 0, part = 0;
    bool owner = false, ipc = false, busy = false;
    bool fetch = false;               // host-buffer calls of a multi-process exchange: wait for the result and read it back
    int home_ordinal = -1;
    unsigned long long seq = 0;       // exchanges issued so far (host side)
    DevBuf status;                    // fetch-kernel status word (multi-process)
    static size_t bytes() { return (size_t)XCHG_MAX_COLS * (8 + 8 + 96 + 96 + 192 * (size_t)XCHG_MAX_PARTS); }
    XchgArgs args(uint32_t col0) const {
        XchgArgs a{};
        a.arrive = reinterpret_cast<unsigned long long*>(base);
        a.done = a.arrive + XCHG_MAX_COLS;
        a.res_mont = reinterpret_cast<uint32_t*>(a.done + XCHG_MAX_COLS);
        a.res_canon = a.res_mont + 24 * XCHG_MAX_COLS;
        a.gather = a.res_canon + 24 * XCHG_MAX_COLS;
        a.host_canon = host_canon;
        a.seq = seq;
        a.n_parts = n_parts;
        a.part = part;
        a.col0 = col0;
        return a;
    }
};
XchgArea g_areas[NSLOT];              // in-process: one per concurrent multi-GPU call
std::mutex g_area_mu;
std::condition_variable g_area_cv;
std::map<uint64_t, std::unique_ptr<XchgArea>>& g_xchg = *new std::map<uint64_t, std::unique_ptr<XchgArea>>();   // b200zk_xchg_create / _open handles
uint64_t g_next_xchg = 1;

XchgArea* acquire_area() {
    std::unique_lock<std::mutex> lk(g_area_mu);
    for (;;) {
        for (auto& a : g_areas)
            if (a.base && !a.busy) { a.busy = true; return &a; }
        g_area_cv.wait(lk);
    }
}
void release_area(XchgArea* a) {
    {
        std::lock_guard<std::mutex> lk(g_area_mu);
        a->busy = false;
    }
    g_area_cv.notify_one();
}

// ------------------------------------------------------------------------------------------
// exclusive scan of n u32 values (in -> out), recursive over 2048-element tiles
// ------------------------------------------------------------------------------------------
int32_t scan_u32(Slot& sl, const uint32_t* in, uint32_t* out, uint64_t n, int level, cudaStream_t s) {
    uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles <= 1) {
        LAUNCH(scan_tile_kernel, 1, SCAN_THREADS, 0, s, in, out, n, (uint32_t*)nullptr, (const uint32_t*)nullptr);
        return B200ZK_OK;
    }
    if (level >= 2) return fail(B200ZK_ERR_INVALID_ARG, "scan: input too large");
    TRY(sl.scan_tmp[level].ensure(2 * tiles * sizeof(uint32_t)));
    uint32_t* sums = sl.scan_tmp[level].as<uint32_t>();
    uint32_t* bases = sums + tiles;
    // pass 1: tile totals only (out is rewritten in pass 3)
    LAUNCH(scan_tile_kernel, (unsigned)tiles, SCAN_THREADS, 0, s, in, out, n, sums, (const uint32_t*)nullptr);
    TRY(scan_u32(sl, sums, bases, tiles, level + 1, s));
    LAUNCH(scan_tile_kernel, (unsigned)tiles, SCAN_THREADS, 0, s, in, out, n, (uint32_t*)nullptr, (const uint32_t*)bases);
    return B200ZK_OK;
}

// ------------------------------------------------------------------------------------------
// MSM
// ------------------------------------------------------------------------------------------
// Window bits minimising the modelled time of one MSM; `shared` = all windows share one bucket set (window tables).
// Throughput terms: one mixed addition (~10 Fp products) per point and window, ~2.2 full additions (~31 products) per
// bucket in the reduction tree.  Latency term: the upper levels of the tree are chains of dependent additions on a
// nearly empty machine, one level per 4-5 bits of bucket index, which a throughput model does not see -- without it
// small MSMs (the prover's k = 14..20 columns) get windows whose tail costs more than their accumulation.
uint32_t msm_choose_window(uint64_t n, uint32_t batch, bool shared) {
    uint32_t best_c = 4;
    double best_cost = 1e300;
    const double unit_rate = 2.6e10;         // Fp products per second at the ~85 % pipe utilisation of a full machine
    for (uint32_t c = 4; c <= 24; c++) {
        uint32_t W = (256 + c - 1) / c;
        double nb = (double)(1u << (c - 1));
        double sets = shared ? 1.0 : (double)W;
        double work = (W * 10.0 * (double)n + sets * 31.0 * nb) * batch / unit_rate;
        // latency floor of the accumulation: a bucket's additions are one thread's serial chain (~7 us per addition
        // when the machine is not full)
        double chain = std::min(1024.0, std::max(1.0, (double)n * (shared ? W : 1) / nb)) * 7e-6;
        double levels = (double)((c - 1 + 4) / 5);
        double tail_latency = levels * 1.1e-4 + (shared ? 0.0 : (double)c * W * 1.2e-6);
        double cost = std::max(work, chain) + tail_latency;
        if ((double)batch * sets * nb * 192.0 > 4.0e9 && c > 8) continue;   // bucket arrays within ~4 GB
        if (cost < best_cost) { best_cost = cost; best_c = c; }
    }
    return best_c;
}

// Window bits of a table's rows, fixed when it is registered.  The model above is calibrated on single MSMs; a prover-size
// table (k <= 19) is mostly committed against in batches (18-43 columns per phase), where every column reduces its own
// bucket set and the tail is bound by throughput, not latency.  Measured on B200 (tools/batch_phases.py,
// profiles/README.md "window bits of prover-size tables"): k = 13, 14: c = 16 beats the model's 15 for single columns
// (1.03 against 1.45 ms at k = 14) and ties for batches; k = 15..18: 16 either way; k = 19: 18 beats 19 for an 18-column
// batch (19.3 against 21.9 ms) and is within 0.4 ms for a single column; k = 20: 20 (the model's) wins for single MSMs.
uint32_t table_window_bits(uint64_t n) {
    if (n > (1ull << 12) && n <= (1ull << 18)) return 16;
    if (n > (1ull << 18) && n <= (1ull << 19)) return 18;
    return std::max(msm_choose_window(n, 1, true), 8u);   // c >= 8: at most 32 rows
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
        if (g_tune_c >= 2 && g_tune_c <= 24) pl.c = g_tune_c;
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
    if (g_tune_smax) smax = g_tune_smax;
    pl.smax = smax;
    return pl;
}

// d_scalars: batch*n Fr (device).  d_bases: n packed Montgomery affine points.
// One chunk of a scalar vector that arrives in pieces (host-buffer MSMs): n_total fixes the plan, i0 is the index of the
// chunk's first point, the first chunk owns sl.buckets, later chunks accumulate into sl.buckets2 and are folded in, the
// last chunk runs the tail.
struct MsmChunk {
    uint64_t n_total, i0;
    bool first, last;
};
int32_t msm_run(Ctx& c, Slot& sl, const BaseTable* tab, const uint32_t* d_bases, const uint32_t* d_scalars, uint64_t n, uint32_t batch,
                uint32_t scalar_fmt, uint32_t* d_out_mont, uint32_t* d_out_canon, cudaStream_t s, uint32_t* d_out_xyzz = nullptr,
                const MsmChunk* ck = nullptr, const XchgArgs* xa = nullptr, cudaStream_t s_acc = nullptr) {
    XchgArgs xnone{};
    if (batch == 0) return B200ZK_OK;
    if (n == 0) {
        if (d_out_mont) CU(cudaMemsetAsync(d_out_mont, 0, 96 * (size_t)batch, s));
        if (d_out_canon) CU(cudaMemsetAsync(d_out_canon, 0, 96 * (size_t)batch, s));
        if (d_out_xyzz) CU(cudaMemsetAsync(d_out_xyzz, 0, 192 * (size_t)batch, s));
        // an empty slice still arrives at the exchange (with the identity)
        if (xa) LAUNCH(msm_combine_kernel, batch, 32, 0, s, (const uint32_t*)nullptr, 0u, 0u, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, *xa);
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

    TRY(sl.counts.ensure((NBt + 1) * 4));
    TRY(sl.offsets.ensure((NBt + 1) * 4));
    TRY(sl.cursor.ensure((NBt + 1) * 4));
    TRY(sl.ntask.ensure((NBt + 1) * 4));
    TRY(sl.task_off.ensure((NBt + 1) * 4));
    TRY(sl.entries.ensure(max_entries * 4));
    TRY(sl.task_bucket.ensure(max_tasks * 4));
    TRY(sl.task_start.ensure(max_tasks * 4));
    TRY(sl.task_len.ensure(max_tasks * 4));
    TRY(sl.len_hist.ensure(((size_t)pl.smax + 2) * 4));
    TRY(sl.len_off.ensure(((size_t)pl.smax + 2) * 4));
    TRY(sl.order.ensure(max_tasks * 4));
    const uint64_t hmax = max_entries / pl.smax + 2;                       // heavy buckets (split into > 1 task)
    const uint64_t imax = hmax + max_tasks / HEAVY_CHUNK + 2;              // their work items
    TRY(sl.heavy.ensure((4 + 3 * hmax + 2 * imax) * 4));
    TRY(sl.heavy_items.ensure(imax * 192));
    TRY(sl.buckets.ensure(NBt * 192));
    if (ck && !ck->first) TRY(sl.buckets2.ensure(NBt * 192));
    TRY(sl.partials.ensure(max_tasks * 192));
    uint64_t m1 = (pl.nb + RED_RADIX - 1) / RED_RADIX, m2 = (m1 + RED_RADIX - 1) / RED_RADIX;
    TRY(sl.redS[0].ensure(nwin * m1 * 192));
    TRY(sl.redA[0].ensure(nwin * m1 * 192));
    TRY(sl.redS[1].ensure(nwin * m2 * 192));
    TRY(sl.redA[1].ensure(nwin * m2 * 192));

    uint32_t* counts = sl.counts.as<uint32_t>();
    uint32_t* offsets = sl.offsets.as<uint32_t>();
    uint32_t* cursor = sl.cursor.as<uint32_t>();
    uint32_t* ntask = sl.ntask.as<uint32_t>();
    uint32_t* task_off = sl.task_off.as<uint32_t>();
    uint32_t* entries = sl.entries.as<uint32_t>();
    uint32_t* buckets = (ck && !ck->first) ? sl.buckets2.as<uint32_t>() : sl.buckets.as<uint32_t>();
    const uint64_t i0 = ck ? ck->i0 : 0;
    uint32_t* partials = sl.partials.as<uint32_t>();

    TRY(ws_enter(sl, s));
    prof_select(c, sl, 1);
    TRY(prof_mark(sl, 0, s));
    CU(cudaMemsetAsync(counts, 0, (NBt + 1) * 4, s));
    CU(cudaMemsetAsync(ntask, 0, (NBt + 1) * 4, s));
    dim3 dgrid((unsigned)((n + 255) / 256), batch);
    const uint32_t fmt_mont = scalar_fmt == B200ZK_FMT_MONT ? 1u : 0u;
    // (A two-pass sort -- coarse bins of 2048 buckets staged through shared memory, then one CTA per bin --
    // was built and measured at 2^24: 11.0 ms against 7.9 ms for this one-pass histogram + scatter; removed.)
    LAUNCH(msm_digits_kernel<0>, dgrid, 256, 0, s, d_scalars, n, fmt_mont, pl, counts, (uint32_t*)nullptr, 0u, 0xffffffffu, (const uint32_t*)nullptr, 0u, i0);
    TRY(scan_u32(sl, counts, offsets, NBt + 1, 0, s));
    CU(cudaMemcpyAsync(cursor, offsets, (NBt + 1) * 4, cudaMemcpyDeviceToDevice, s));
    {
        // The scatter writes 4-byte entries into bucket lists that are spread over the whole entry array.  When that
        // array is much larger than L2, the 32-byte sector a list is currently filling is evicted half full and read
        // back (DRAM read-modify-write).  Scattering one bucket range at a time keeps the open sectors (32 B per bucket
        // of the range) resident; the scalars are re-read and re-coded once per pass, which is cheap.
        uint32_t passes = 1;
        if (g_scatter_passes_env > 0) passes = (uint32_t)g_scatter_passes_env;
        else if (NBt * 32 >= (48ull << 20)) passes = 2;   // measured at 2^24 (2^21 buckets): 1 pass 8.98 ms, 2 passes 7.44, 4 passes 9.22 (each pass re-codes every scalar)
        for (uint32_t ps = 0; ps < passes; ps++) {
            uint32_t lo = (uint32_t)(NBt * ps / passes), hi = (uint32_t)(NBt * (ps + 1) / passes);
            LAUNCH(msm_digits_kernel<1>, dgrid, 256, 0, s, d_scalars, n, fmt_mont, pl, cursor, entries, lo, hi,
                   passes > 1 ? (const uint32_t*)(offsets + NBt) : (const uint32_t*)nullptr, 100u << 20, i0);
        }
    }
    unsigned bgrid = (unsigned)((NBt + 255) / 256);
    LAUNCH(msm_task_count_kernel, bgrid, 256, 0, s, (const uint32_t*)counts, NBt, pl.smax, ntask);
    TRY(scan_u32(sl, ntask, task_off, NBt + 1, 0, s));
    uint32_t* len_hist = sl.len_hist.as<uint32_t>();
    uint32_t* len_off = sl.len_off.as<uint32_t>();
    uint32_t* order = sl.order.as<uint32_t>();
    HeavyArrays hv;
    hv.count = sl.heavy.as<uint32_t>();
    hv.bucket = hv.count + 4;
    hv.base = hv.bucket + hmax;
    hv.done = hv.base + hmax;
    hv.item_slot = hv.done + hmax;
    hv.item_chunk = hv.item_slot + imax;
    CU(cudaMemsetAsync(len_hist, 0, ((size_t)pl.smax + 2) * 4, s));
    CU(cudaMemsetAsync(hv.count, 0, 16, s));
    LAUNCH(msm_task_emit_kernel, bgrid, 256, 0, s, (const uint32_t*)counts, (const uint32_t*)offsets,
           (const uint32_t*)task_off, NBt, pl.smax, sl.task_bucket.as<uint32_t>(), sl.task_start.as<uint32_t>(),
           sl.task_len.as<uint32_t>(), len_hist);
    TRY(scan_u32(sl, len_hist, len_off, (uint64_t)pl.smax + 1, 0, s));
    LAUNCH(msm_task_order_kernel, (unsigned)((max_tasks + 255) / 256), 256, 0, s, (const uint32_t*)sl.task_len.as<uint32_t>(),
           (const uint32_t*)(task_off + NBt), pl.smax, len_off, order);
    LAUNCH(msm_heavy_list_kernel, bgrid, 256, 0, s, (const uint32_t*)ntask, NBt, hv);
    CU(cudaMemsetAsync(buckets, 0, NBt * 192, s));
    TRY(prof_mark(sl, 1, s));
    const cudaStream_t s_main = s;
    if (s_acc) {   // the accumulation runs on its own (low-priority) stream between two events
        CU(cudaEventRecord(sl.ev_sorted, s_main));
        CU(cudaStreamWaitEvent(s_acc, sl.ev_sorted, 0));
        s = s_acc;
    }
    {
        unsigned agrid = (unsigned)((max_tasks + 127) / 128);
        const uint32_t *tb = sl.task_bucket.as<uint32_t>(), *ts = sl.task_start.as<uint32_t>(), *tl = sl.task_len.as<uint32_t>();
        const uint32_t* ntp = task_off + NBt;
        switch (g_tune_variant) {
            case 1: LAUNCH(msm_accumulate_kernel_v1, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
            case 2: LAUNCH(msm_accumulate_kernel_v2, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
            case 7: LAUNCH(msm_accumulate_kernel_v7, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
            case 6: LAUNCH(msm_accumulate_kernel_v6, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
            case 3: LAUNCH(msm_accumulate_kernel_v3, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
            case 4:
            case 5: {
                size_t slots = max_entries / 2 + 2;
                TRY(sl.aff_a.ensure(slots * 96));
                TRY(sl.aff_b.ensure(slots * 96));
                TRY(sl.aff_pre.ensure(slots * 48));
                if (g_tune_variant == 4)
                    LAUNCH(msm_accumulate_affine_kernel, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials,
                           sl.aff_a.as<uint32_t>(), sl.aff_b.as<uint32_t>(), sl.aff_pre.as<uint32_t>());
                else
                    LAUNCH(msm_accumulate_affine_kernel_r168, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials,
                           sl.aff_a.as<uint32_t>(), sl.aff_b.as<uint32_t>(), sl.aff_pre.as<uint32_t>());
                break;
            }
            default: LAUNCH(msm_accumulate_kernel, agrid, 128, 0, s, d_bases, (const uint32_t*)entries, tb, ts, tl, (const uint32_t*)order, ntp, buckets, partials); break;
        }
    }
    if (s_acc) {
        CU(cudaEventRecord(sl.ev_acc, s_acc));
        s = s_main;
        CU(cudaStreamWaitEvent(s, sl.ev_acc, 0));
    }
    TRY(prof_mark(sl, 2, s));
    LAUNCH(msm_collapse_kernel, (unsigned)(c.prop.multiProcessorCount * 4), 128, 0, s, (const uint32_t*)ntask,
           (const uint32_t*)task_off, hv, (const uint32_t*)partials, sl.heavy_items.as<uint32_t>(), buckets);

    if (ck && !ck->first)
        LAUNCH(msm_bucket_merge_kernel, (unsigned)((NBt + 127) / 128), 128, 0, s, sl.buckets.as<uint32_t>(), (const uint32_t*)buckets, NBt);
    if (ck && !ck->last) {
        TRY(ws_leave(sl, s));
        sl.last_plan = pl;
        return B200ZK_OK;
    }
    buckets = sl.buckets.as<uint32_t>();

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
        uint32_t* S_out = sl.redS[pp].as<uint32_t>();
        uint32_t* A_out = sl.redA[pp].as<uint32_t>();
        uint64_t groups = (uint64_t)m_out * nwin;
        if (serial && g_red_tp > 0 && groups >= (uint64_t)g_red_tp_min) {
            // enough groups to fill the machine: the throughput build of the same level (msm.cuh)
            if (g_red_tp == 3)
                LAUNCH(msm_reduce_tp_kernel<3>, (unsigned)((groups + 127) / 128), 128, 0, s, S_in, A_in, S_out, A_out, m, m_out, (uint32_t)nwin, scale_log);
            else
                LAUNCH(msm_reduce_tp_kernel<4>, (unsigned)((groups + 127) / 128), 128, 0, s, S_in, A_in, S_out, A_out, m, m_out, (uint32_t)nwin, scale_log);
        } else if (serial)
            LAUNCH(msm_reduce_kernel, (unsigned)((groups + 63) / 64), 64, 0, s, S_in, A_in, S_out, A_out, m, m_out,
                   (uint32_t)nwin, scale_log);
        else if (groups <= (uint64_t)(g_coop8_max > 0 ? g_coop8_max : c.prop.multiProcessorCount))
            // top of the tree: one CTA per group, 8 lanes per node, additions in 4 product levels instead of 14 products
            // (only while the groups fit the machine at once -- the kernel holds one CTA per SM: at 1 024 groups this kernel takes
            // 1.2 ms against 0.38 ms for one warp per group, and the 512-group level of a 2^18-bucket tree cost a single 2^19
            // MSM 0.38 ms of its 1.75 ms tail while the limit stood at 600 groups)
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
    LAUNCH(msm_combine_kernel, batch, 32, 0, s, win_sums, pl.precomp ? 1u : pl.W, pl.c, d_out_mont, d_out_canon, d_out_xyzz, xa ? *xa : xnone);
    TRY(prof_mark(sl, 3, s));
    TRY(ws_leave(sl, s));
    sl.last_plan = pl;
    return B200ZK_OK;
}

// A batch of columns on caller stream s.  With a second free slot and enough work the batch is cut in two halves that run as
// two pipelines (sort, tail on each slot's high-priority stream; accumulate on its low-priority stream): the second half's
// sort is dispatched into the first half's accumulation and the first half's tail into the second half's accumulation.
// Measured (B200, 18 columns of 2^17 prover-like scalars / 18 of 2^19): see profiles/README.md.
int32_t msm_run_batch(Ctx& c, Slot& sl, const BaseTable* tab, const uint32_t* d_bases, const uint32_t* d_scalars, uint64_t n, uint32_t batch,
                      uint32_t scalar_fmt, uint32_t* d_out_mont, uint32_t* d_out_canon, cudaStream_t s, const XchgArgs* xa) {
    static int64_t pipe_min = -1;
    // off by default: measured on B200 (profiles/r02_batch_pipeline.jsonl) the two pipelines do not beat one launch sequence over
    // the whole batch (18 x 2^17: 7.60 ms in one sequence, 8.96 ms piped; 18 x 2^19: 28.7 / 29.1 ms)
    if (pipe_min < 0) { const char* v = getenv("B200ZK_BATCH_PIPE_MIN_POINTS"); pipe_min = v ? atoll(v) : 0; }
    Slot* sl2 = nullptr;
    if (batch >= 2 && pipe_min > 0 && (int64_t)(n * batch) >= pipe_min) sl2 = try_acquire_slot(c);
    if (!sl2) return msm_run(c, sl, tab, d_bases, d_scalars, n, batch, scalar_fmt, d_out_mont, d_out_canon, s, nullptr, nullptr, xa);
    struct Release { Ctx& c; Slot* s; ~Release() { release_slot(c, s); } } rel{c, sl2};
    Slot* half_slot[2] = {&sl, sl2};
    TRY(slot_pipe_streams(sl));
    TRY(slot_pipe_streams(*sl2));
    const uint32_t bA = (batch + 1) / 2;
    CU(cudaEventRecord(sl.ev_fork, s));
    for (int h = 0; h < 2; h++) {
        Slot& hs = *half_slot[h];
        const uint32_t b0 = h ? bA : 0, cnt = h ? batch - bA : bA;
        CU(cudaStreamWaitEvent(hs.hi, sl.ev_fork, 0));
        XchgArgs xh{};
        if (xa) { xh = *xa; xh.col0 = xa->col0 + b0; }
        TRY(msm_run(c, hs, tab, d_bases, d_scalars + 8 * (size_t)n * b0, n, cnt, scalar_fmt, d_out_mont ? d_out_mont + 24 * (size_t)b0 : nullptr,
                    d_out_canon ? d_out_canon + 24 * (size_t)b0 : nullptr, hs.hi, nullptr, nullptr, xa ? &xh : nullptr, hs.lo));
        CU(cudaEventRecord(hs.ev_done, hs.hi));
    }
    for (int h = 0; h < 2; h++) CU(cudaStreamWaitEvent(s, half_slot[h]->ev_done, 0));
    // later users of either slot's scratch order themselves behind s
    TRY(ws_leave(sl, s));
    TRY(ws_leave(*sl2, s));
    return B200ZK_OK;
}

std::shared_ptr<TableSet> find_table(uint64_t handle) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto it = g_tables.find(handle);
    return it == g_tables.end() ? nullptr : it->second;
}
int32_t get_table(uint64_t handle, uint64_t offset, uint64_t n, std::shared_ptr<TableSet>* out) {
    auto ts = find_table(handle);
    if (!ts) return fail(B200ZK_ERR_BAD_HANDLE, "unknown base-table handle");
    if (offset > ts->n || n > ts->n - offset) return fail(B200ZK_ERR_INVALID_ARG, "msm: offset + n exceeds the registered table");
    *out = ts;
    return B200ZK_OK;
}
// the shard of `ts` on device `dev` that holds all of [offset, offset + n), or null
const TableShard* shard_on(const TableSet& ts, int dev, uint64_t offset, uint64_t n) {
    for (const TableShard& sh : ts.shards)
        if (sh.dev == dev && offset >= sh.start && offset + n <= sh.start + sh.n) return &sh;
    return nullptr;
}

int32_t ingest_bases(Ctx& c, const uint8_t* d_src, uint64_t n, uint32_t fmt, uint32_t stride, uint32_t* d_dst, cudaStream_t s) {
    TRY(c.reg_flag.ensure(4));
    CU(cudaMemsetAsync(c.reg_flag.p, 0, 4, s));
    LAUNCH(g1_ingest_kernel, (unsigned)((n + 127) / 128), 128, 0, s, d_src, n, stride, fmt == B200ZK_FMT_MONT ? 1u : 0u,
           d_dst, c.reg_flag.as<uint32_t>());
    uint32_t bad = 0;
    CU(cudaMemcpyAsync(&bad, c.reg_flag.p, 4, cudaMemcpyDeviceToHost, s));
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

// plans and coset tables are per device and built once (under the device mutex, on the utility stream)
int32_t ntt_get_plan(Ctx& c, uint32_t log_n, const uint8_t omega[32], bool inverse, NttPlan** out) {
    std::lock_guard<std::mutex> lk(c.mu);
    NttKey key;
    key.log_n = log_n;
    key.inverse = inverse ? 1 : 0;
    memcpy(key.omega, omega, 32);
    auto it = c.ntt_plans.find(key);
    if (it != c.ntt_plans.end()) { *out = &it->second; return B200ZK_OK; }
    cudaStream_t s = c.stream;
    NttPlan pl;
    pl.log_n = log_n;
    // experiment (B200ZK_NTT_VARIANT=8): 2^23 and 2^24 as two passes over a 4096-element tile instead of three over a
    // 2048-element one.  Measured at 2^24: 4.13 ms against 3.67 ms -- one 512-thread CTA per SM loses more at its barriers and
    // load / store phases (0.18 ms per butterfly stage against 0.164) than the saved pass and twiddle product give back.
    if (log_n > 2 * NTT_LOGB && log_n <= 2 * NTT_LOGB12 && g_ntt_variant == 8) pl.logb = NTT_LOGB12;
    pl.npass = (log_n + pl.logb - 1) / pl.logb;
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
    auto ins = c.ntt_plans.emplace(key, pl);
    *out = &ins.first->second;
    return B200ZK_OK;
}

int32_t ntt_get_coset(Ctx& c, uint32_t log_n, const uint8_t shift[32], uint32_t** out) {
    std::lock_guard<std::mutex> lk(c.mu);
    CosetKey key;
    key.log_n = log_n;
    memcpy(key.shift, shift, 32);
    auto it = c.coset_tables.find(key);
    if (it != c.coset_tables.end()) { *out = it->second; return B200ZK_OK; }
    cudaStream_t s = c.stream;
    uint64_t n = (uint64_t)1 << log_n;
    uint32_t *d_shift = nullptr, *tab = nullptr;
    CU(cudaMalloc(&d_shift, 32));
    TRY(upload_fr_mont(shift, d_shift, s));
    CU(cudaMalloc(&tab, n * 32));
    LAUNCH(fr_powers_kernel, (unsigned)((n + 127) / 128), 128, 0, s, tab, (const uint32_t*)d_shift, (const uint32_t*)nullptr, n,
           (uint64_t)1, 0u, 0u);
    CU(cudaStreamSynchronize(s));
    CU(cudaFree(d_shift));
    c.coset_tables[key] = tab;
    *out = tab;
    return B200ZK_OK;
}

int32_t ntt_set_attrs(Ctx& c) {
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.ntt_attr_set) return B200ZK_OK;
    CU(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
    CU(cudaFuncSetAttribute(ntt_pass_kernel_occ2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
    CU(cudaFuncSetAttribute(ntt_pass_kernel_call2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
    CU(cudaFuncSetAttribute(ntt_pass_kernel_call3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
    CU(cudaFuncSetAttribute(ntt_pass_kernel_plain2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
    CU(cudaFuncSetAttribute(ntt_pass_kernel_wl2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
    CU(cudaFuncSetAttribute(ntt_pass_kernel_tw2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM_TW));
    CU(cudaFuncSetAttribute(ntt_pass_kernel_lb0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
    CU(cudaFuncSetAttribute(ntt_pass_kernel_lb0_lazy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM));
    CU(cudaFuncSetAttribute(ntt_pass_kernel_lb0_t12, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM12));
    c.ntt_attr_set = true;
    return B200ZK_OK;
}

int32_t ntt_check_args(uint32_t log_n, uint32_t flags, const uint8_t* coset_shift) {
    if (log_n > 32) return fail(B200ZK_ERR_INVALID_ARG, "ntt: log_n exceeds the 2-adicity of Fr");
    if ((flags & (B200ZK_NTT_COSET_IN | B200ZK_NTT_COSET_OUT)) && !coset_shift)
        return fail(B200ZK_ERR_INVALID_ARG, "ntt: COSET flag without a shift");
    if ((flags & B200ZK_NTT_COSET_IN) && (flags & B200ZK_NTT_COSET_OUT))
        return fail(B200ZK_ERR_INVALID_ARG, "ntt: COSET_IN and COSET_OUT are mutually exclusive");
    return B200ZK_OK;
}

struct NttExtra {
    uint32_t* out_final = nullptr;       // the last pass writes here instead of d_data
    uint32_t t_out = 0;                  // 1 + log2(batch): ... transposed, element (poly, i) at (i << log2 batch) + poly
    const uint32_t* out_scale = nullptr; // multiplier table indexed like the (transposed) output
};
int32_t ntt_run(Ctx& c, Slot& sl, uint32_t* d_data, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                const uint8_t* coset_shift, cudaStream_t s, const NttExtra* ex = nullptr) {
    TRY(ntt_check_args(log_n, flags, coset_shift));
    if (batch == 0) return B200ZK_OK;
    bool inverse = (flags & B200ZK_NTT_INVERSE_SCALE) != 0;
    NttPlan* pl = nullptr;
    TRY(ntt_get_plan(c, log_n, omega, inverse, &pl));
    uint32_t* coset = nullptr;
    if (flags & (B200ZK_NTT_COSET_IN | B200ZK_NTT_COSET_OUT)) TRY(ntt_get_coset(c, log_n, coset_shift, &coset));
    uint64_t n = (uint64_t)1 << log_n, total = n * batch;
    uint32_t* bufs[2] = {nullptr, nullptr};
    if (pl->npass > 1) {
        TRY(sl.ntt_tmp[0].ensure(total * 32));
        bufs[0] = sl.ntt_tmp[0].as<uint32_t>();
        if (pl->npass > 2) {
            TRY(sl.ntt_tmp[1].ensure(total * 32));
            bufs[1] = sl.ntt_tmp[1].as<uint32_t>();
        }
    }
    TRY(ntt_set_attrs(c));
    const int ntt_variant = g_ntt_variant;
    uint32_t log_s = 0;
    const uint32_t* src = d_data;
    TRY(ws_enter(sl, s));
    prof_select(c, sl, 2);
    TRY(prof_mark(sl, 0, s));
    for (uint32_t i = 0; i < pl->npass; i++) {
        bool last = (i + 1 == pl->npass);
        uint32_t* dst = last ? ((ex && ex->out_final) ? ex->out_final : d_data) : bufs[i & 1];
        if (pl->npass == 3 && i == 1) dst = bufs[1];
        NttPassArgs a;
        memset(&a, 0, sizeof a);
        a.in = src;
        a.out = dst;
        a.tw_local = pl->tw_local[i];
        a.tw_pass = pl->tw_pass[i];
        a.in_scale = (i == 0 && (flags & B200ZK_NTT_COSET_IN)) ? coset : nullptr;
        a.out_scale = (last && (flags & B200ZK_NTT_COSET_OUT)) ? coset : nullptr;
        if (last && ex) {
            if (ex->out_scale) a.out_scale = ex->out_scale;
            a.t_out = ex->t_out;
        }
        a.scalar = (last && inverse && pl->npass == 1) ? pl->ninv : nullptr;
        a.reduce_in = (i == 0 && !(flags & B200ZK_NTT_MONT)) ? 1 : 0;
        a.log_n = log_n;
        a.deg = pl->deg[i];
        a.log_s = log_s;
        a.log_cols = log_n - pl->deg[i];
        a.total_cols = (uint64_t)batch << a.log_cols;
        uint64_t ctas = (total + ((uint64_t)1 << pl->logb) - 1) >> pl->logb;
        if (pl->logb == NTT_LOGB12) LAUNCH(ntt_pass_kernel_lb0_t12, (unsigned)ctas, NTT_THREADS12, NTT_SMEM12, s, a);
        else if (ntt_variant == 0) LAUNCH(ntt_pass_kernel, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else if (ntt_variant == 2) LAUNCH(ntt_pass_kernel_call2, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else if (ntt_variant == 3) LAUNCH(ntt_pass_kernel_call3, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else if (ntt_variant == 4) LAUNCH(ntt_pass_kernel_plain2, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else if (ntt_variant == 5) LAUNCH(ntt_pass_kernel_wl2, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else if (ntt_variant == 6) LAUNCH(ntt_pass_kernel_tw2, (unsigned)ctas, NTT_THREADS, NTT_SMEM_TW, s, a);
        else if (ntt_variant == 1) LAUNCH(ntt_pass_kernel_occ2, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else if (ntt_variant == 9) LAUNCH(ntt_pass_kernel_lb0_lazy, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        else LAUNCH(ntt_pass_kernel_lb0, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
        TRY(prof_mark(sl, (int)i + 1, s));
        src = dst;
        log_s += pl->deg[i];
    }
    TRY(ws_leave(sl, s));
    return B200ZK_OK;
}

// generator of G1 in canonical wire form (its compressed form is the KAT at
// /root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:125)
const uint8_t G1_GEN_X_BE[48] = {0x17, 0xf1, 0xd3, 0xa7, 0x31, 0x97, 0xd7, 0x94, 0x26, 0x95, 0x63, 0x8c, 0x4f, 0xa9, 0xac, 0x0f,
                                 0xc3, 0x68, 0x8c, 0x4f, 0x97, 0x74, 0xb9, 0x05, 0xa1, 0x4e, 0x3a, 0x3f, 0x17, 0x1b, 0xac, 0x58,
                                 0x6c, 0x55, 0xe8, 0x3f, 0xf9, 0x7a, 0x1a, 0xef, 0xfb, 0x3a, 0xf0, 0x0a, 0xdb, 0x22, 0xc6, 0xbb};
const uint8_t G1_GEN_Y_BE[48] = {0x08, 0xb3, 0xf4, 0x81, 0xe3, 0xaa, 0xa0, 0xf1, 0xa0, 0x9e, 0x30, 0xed, 0x74, 0x1d, 0x8a, 0xe4,
                                 0xfc, 0xf5, 0xe0, 0x95, 0xd5, 0xd0, 0x0a, 0xf6, 0x00, 0xdb, 0x18, 0xcb, 0x2c, 0x04, 0xb3, 0xed,
                                 0xd0, 0x3c, 0xc7, 0x44, 0xa2, 0x88, 0x8a, 0xe4, 0x0c, 0xaa, 0x23, 0x29, 0x46, 0xc5, 0xe7, 0xe1};

// the device's copy of the generator (Montgomery affine), built on first use.  Guarded by its own mutex and using its own
// scratch: entry points of the other translation units call this while they hold the device mutex.
int32_t get_generator_dev(Ctx& c, uint32_t** out, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(c.gen_mu);
    if (!c.d_gen) {
        uint8_t wire[96];
        for (int i = 0; i < 48; i++) { wire[i] = G1_GEN_X_BE[47 - i]; wire[48 + i] = G1_GEN_Y_BE[47 - i]; }
        uint8_t* d_wire = nullptr;
        uint32_t* d_flag = nullptr;
        CU(cudaMalloc(&d_wire, 96 + 16));
        CU(cudaMalloc(&c.d_gen, 96));
        d_flag = reinterpret_cast<uint32_t*>(d_wire + 96);
        CU(cudaMemcpyAsync(d_wire, wire, 96, cudaMemcpyHostToDevice, s));
        CU(cudaMemsetAsync(d_flag, 0, 4, s));
        LAUNCH(g1_ingest_kernel, 1, 128, 0, s, (const uint8_t*)d_wire, (uint64_t)1, 96u, 0u, c.d_gen, d_flag);
        CU(cudaStreamSynchronize(s));
        cudaFree(d_wire);
    }
    *out = c.d_gen;
    return B200ZK_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer MSM on one device
// ------------------------------------------------------------------------------------------
// where the scalar columns of a call live on the host: an array of pointers (separate Vecs) or one block
struct ColSrc {
    const uint8_t* const* ptrs = nullptr;
    const uint8_t* base = nullptr;
    size_t stride = 0;   // bytes between consecutive columns of `base`
    // resident scalars of a sharded call: slice k lives in the HBM of the GPU that holds shard k of the table (one column)
    const void* const* dev_slices = nullptr;
    const uint8_t* col(uint32_t j) const { return ptrs ? ptrs[j] : base + (size_t)j * stride; }
};

// copies scalars [i0, i0 + n) of columns [col0, col0 + ncols) to d_dst (column-major, n per column)
int32_t upload_columns(const ColSrc& src, uint32_t col0, uint32_t ncols, uint64_t i0, uint64_t n, uint8_t* d_dst, cudaStream_t s) {
    const size_t colb = (size_t)n * 32;
    if (!src.ptrs && src.stride == colb) {
        CU(cudaMemcpyAsync(d_dst, src.base + (size_t)col0 * src.stride + i0 * 32, colb * ncols, cudaMemcpyHostToDevice, s));
        return B200ZK_OK;
    }
    for (uint32_t j = 0; j < ncols; j++)
        CU(cudaMemcpyAsync(d_dst + (size_t)j * colb, src.col(col0 + j) + i0 * 32, colb, cudaMemcpyHostToDevice, s));
    return B200ZK_OK;
}

// Columns [col0, col0 + ncols) of a call, points [i0, i0 + n) of each, against tab (whose point 0 pairs with scalar
// index i0 when bases_off is the matching offset).  Either writes the finished commitments to `out` (96 bytes per
// column) or, with an exchange, leaves its XYZZ partials at the meeting point.  Enqueues on the slot's stream and
// waits for it.
int32_t msm_host_device(Ctx& c, Slot& sl, const BaseTable* tab, uint64_t bases_off, const ColSrc& src, uint32_t col0, uint32_t ncols,
                        uint64_t i0, uint64_t n, uint32_t scalar_fmt, const XchgArea* area, uint8_t* out) {
    cudaStream_t s = sl.stream;
    const uint32_t* d_bases = tab->d + 24 * bases_off;
    const size_t bytes = (size_t)n * ncols * 32;
    if (src.dev_slices) {
        // scalars already in this GPU's HBM: straight to the kernels, the partial goes to the meeting point
        XchgArgs xa = area->args(0);
        TRY(msm_run(c, sl, tab, d_bases, reinterpret_cast<const uint32_t*>(src.dev_slices[area->part]), n, 1, scalar_fmt, nullptr, nullptr, s,
                    nullptr, nullptr, &xa));
        CU(cudaStreamSynchronize(s));
        return B200ZK_OK;
    }
    TRY(sl.scalars.ensure(bytes + 16));
    TRY(sl.out_canon.ensure((size_t)ncols * 96));
    TRY(sl.h_out.ensure((size_t)ncols * 96 + 16));
    uint8_t* d_sc = sl.scalars.as<uint8_t>();
    uint32_t* d_out = area ? nullptr : sl.out_canon.as<uint32_t>();
    {
        const uint32_t gcols = ncols;                          // (an exchange has at most XCHG_MAX_COLS columns: checked by the caller)
        XchgArgs xa{};
        if (area) xa = area->args(0);
        const XchgArgs* xp = area ? &xa : nullptr;
        uint32_t* d_sc32 = reinterpret_cast<uint32_t*>(d_sc);
        uint64_t nA = 0;
        if (gcols == 1 && g_chunk_min > 0 && n >= (uint64_t)g_chunk_min) {
            // first piece = 1/8 of the points: its copy (1.2 ms at 2^24) is the only exposed transfer, and its sort + accumulate
            // (9 ms) cover the copy of the other 7/8 (8.5 ms).  Measured at 2^24: 84.9 ms unchunked, 83.1 ms with halves.
            // From pageable memory (a plain Rust Vec) the copy runs at ~11 GB/s instead of ~55: the balance point
            // copy(rest) = compute(first) moves from 1/8 to 3/8 (measured at 2^24, pageable: 122.8 ms unchunked, 118.7 with 1/8).
            cudaPointerAttributes pa{};
            bool pinned = cudaPointerGetAttributes(&pa, src.col(col0)) == cudaSuccess && pa.type == cudaMemoryTypeHost;
            cudaGetLastError();
            nA = ((pinned ? n / 8 : 3 * (n / 8)) + 255) & ~(uint64_t)255;
            if (nA >= n) nA = 0;                               // (tiny n with a lowered threshold: no chunking)
        }
        if (nA) {
            // The second piece of the scalars crosses PCIe while the first is sorted and accumulated: the first piece's
            // buckets are sl.buckets, the second accumulates into sl.buckets2 and is folded in before the tail.
            // (the first piece's kernels are enqueued before the second copy is issued: from pageable host memory a copy
            // blocks the calling thread, and the device must already have work by then)
            TRY(slot_streams(sl));
            const uint64_t nB = n - nA;
            const uint8_t* h = src.col(col0) + i0 * 32;
            MsmChunk ca{n, 0, true, false}, cb{n, nA, false, true};
            CU(cudaMemcpyAsync(d_sc32, h, nA * 32, cudaMemcpyHostToDevice, sl.copy_stream));
            CU(cudaEventRecord(sl.copy_ev[0], sl.copy_stream));
            CU(cudaStreamWaitEvent(s, sl.copy_ev[0], 0));
            TRY(msm_run(c, sl, tab, d_bases, d_sc32, nA, 1, scalar_fmt, nullptr, nullptr, s, nullptr, &ca));
            CU(cudaMemcpyAsync(d_sc32 + 8 * nA, h + nA * 32, nB * 32, cudaMemcpyHostToDevice, sl.copy_stream));
            CU(cudaEventRecord(sl.copy_ev[1], sl.copy_stream));
            CU(cudaStreamWaitEvent(s, sl.copy_ev[1], 0));
            TRY(msm_run(c, sl, tab, d_bases, d_sc32 + 8 * nA, nB, 1, scalar_fmt, nullptr, d_out, s, nullptr, &cb, xp));
        } else if (gcols >= 4 && g_chunk_min > 0 && bytes >= g_batch_stream_min) {
            // The prover's pattern (all columns of a phase in one call): the columns are independent, so the first eighth of
            // them goes up and starts computing while the others cross PCIe; no merge is needed.
            TRY(slot_streams(sl));
            const uint32_t bA = std::max(1u, gcols / 8), bB = gcols - bA;
            XchgArgs xb = xa;
            xb.col0 = bA;
            TRY(upload_columns(src, col0, bA, i0, n, reinterpret_cast<uint8_t*>(d_sc32), sl.copy_stream));
            CU(cudaEventRecord(sl.copy_ev[0], sl.copy_stream));
            CU(cudaStreamWaitEvent(s, sl.copy_ev[0], 0));
            TRY(msm_run(c, sl, tab, d_bases, d_sc32, n, bA, scalar_fmt, nullptr, d_out, s, nullptr, nullptr, xp));
            TRY(upload_columns(src, col0 + bA, bB, i0, n, reinterpret_cast<uint8_t*>(d_sc32 + 8 * (size_t)n * bA), sl.copy_stream));
            CU(cudaEventRecord(sl.copy_ev[1], sl.copy_stream));
            CU(cudaStreamWaitEvent(s, sl.copy_ev[1], 0));
            TRY(msm_run_batch(c, sl, tab, d_bases, d_sc32 + 8 * (size_t)n * bA, n, bB, scalar_fmt, nullptr, d_out ? d_out + 24 * (size_t)bA : nullptr,
                              s, area ? &xb : nullptr));
        } else {
            TRY(upload_columns(src, col0, gcols, i0, n, reinterpret_cast<uint8_t*>(d_sc32), s));
            if (gcols >= 2) TRY(msm_run_batch(c, sl, tab, d_bases, d_sc32, n, gcols, scalar_fmt, nullptr, d_out, s, xp));
            else TRY(msm_run(c, sl, tab, d_bases, d_sc32, n, gcols, scalar_fmt, nullptr, d_out, s, nullptr, nullptr, xp));
        }
    }
    const bool readback = !area || area->fetch;
    if (area && area->fetch) {
        XchgArgs xa = area->args(0);
        TRY(sl.flag.ensure(4));
        CU(cudaMemsetAsync(sl.flag.p, 0, 4, s));
        LAUNCH(xchg_fetch_kernel, ncols, 32, 0, s, (const unsigned long long*)xa.done, xa.seq, (const uint32_t*)xa.res_mont,
               (const uint32_t*)xa.res_canon, (uint32_t*)nullptr, sl.out_canon.as<uint32_t>(), sl.flag.as<uint32_t>(), 8000000000ll);
        CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(sl.h_out.p) + (size_t)ncols * 96, sl.flag.p, 4, cudaMemcpyDeviceToHost, s));
    }
    if (readback) CU(cudaMemcpyAsync(sl.h_out.p, sl.out_canon.p, (size_t)ncols * 96, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (area && area->fetch) {
        uint32_t timed_out = 0;
        memcpy(&timed_out, reinterpret_cast<uint8_t*>(sl.h_out.p) + (size_t)ncols * 96, 4);
        if (timed_out) return fail(B200ZK_ERR_CUDA, "exchange: a peer never delivered its partial sum (timed out)");
    }
    if (readback) memcpy(out, sl.h_out.p, (size_t)ncols * 96);
    return B200ZK_OK;
}

// The plan of one host-buffer MSM call over the bound devices.
struct MsmPiece {
    Ctx* c;
    const TableShard* sh;
    uint32_t col0, ncols;     // columns of the call this device computes
    uint64_t i0, n;           // scalar index range of those columns
    uint64_t bases_off;       // offset of point i0 inside the shard's table
};

int32_t msm_host(uint64_t bases, uint64_t offset, const ColSrc& src, uint64_t n, uint32_t batch, uint32_t scalar_fmt, uint8_t* out_affine) {
    TRY(need_init());
    if (!out_affine || (!src.ptrs && !src.base && !src.dev_slices && n && batch)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (scalar_fmt > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown scalar format");
    if (batch == 0) return B200ZK_OK;
    if (src.ptrs && n)
        for (uint32_t j = 0; j < batch; j++)
            if (!src.ptrs[j]) return fail(B200ZK_ERR_INVALID_ARG, "null scalar column");
    std::shared_ptr<TableSet> ts;
    TRY(get_table(bases, offset, n, &ts));
    if (n == 0) { memset(out_affine, 0, (size_t)batch * 96); return B200ZK_OK; }
    const int D = (int)g_devs.size();
    std::vector<MsmPiece> pieces;
    bool exchange = false;
    if (src.dev_slices && (ts->shards.size() == 1 || ts->replicated || offset != 0 || n != ts->n || batch != 1))
        return fail(B200ZK_ERR_INVALID_ARG, "sharded_dev: needs a table partitioned over several GPUs and one full-length scalar vector sliced like it");
    if (ts->shards.size() == 1) {
        const TableShard& sh = ts->shards[0];
        pieces.push_back({g_devs[sh.dev].get(), &sh, 0, batch, 0, n, offset - sh.start});
    } else if (ts->replicated && (batch >= 2 || n < g_range_split_min)) {
        // independent columns: deal them out in contiguous blocks, no exchange (a lone small MSM stays on one GPU)
        int used = (int)std::min<uint32_t>(batch, (uint32_t)D);
        static std::atomic<unsigned> rr{0};
        unsigned first = batch == 1 ? rr.fetch_add(1) % (unsigned)D : 0;
        for (int k = 0; k < used; k++) {
            uint32_t b0 = (uint32_t)((uint64_t)batch * k / used), b1 = (uint32_t)((uint64_t)batch * (k + 1) / used);
            const TableShard& sh = ts->shards[(first + k) % D];
            pieces.push_back({g_devs[sh.dev].get(), &sh, b0, b1 - b0, 0, n, offset});
        }
    } else {
        // point-range split with the partials meeting on the first GPU
        exchange = true;
        if (batch > XCHG_MAX_COLS) return fail(B200ZK_ERR_INVALID_ARG, "msm: more than 1024 columns in one sharded call; split the batch");
        for (int k = 0; k < (int)ts->shards.size(); k++) {
            const TableShard& sh = ts->shards[k];
            uint64_t lo, hi;
            if (ts->replicated) {
                lo = offset + n * k / D;
                hi = offset + n * (k + 1) / D;
            } else {
                lo = std::max(offset, sh.start);
                hi = std::min(offset + n, sh.start + sh.n);
                if (hi < lo) hi = lo;
            }
            // every device takes part in the exchange, an empty slice contributes the identity
            pieces.push_back({g_devs[sh.dev].get(), &sh, 0, batch, lo - offset, hi - lo, (hi > lo ? lo : sh.start) - sh.start});
        }
    }
    // slots in device order (ordered acquisition: concurrent multi-GPU calls cannot deadlock), then the meeting point
    std::sort(pieces.begin(), pieces.end(), [](const MsmPiece& a, const MsmPiece& b) { return a.c->index < b.c->index; });
    std::vector<SlotLease> leases;
    for (auto& p : pieces) leases.emplace_back(*p.c);
    XchgArea* area = nullptr;
    if (exchange) {
        area = acquire_area();
        area->seq++;
        area->n_parts = (uint32_t)pieces.size();
    }
    std::vector<Job> jobs;
    std::vector<XchgArea> views(pieces.size());
    for (size_t k = 0; k < pieces.size(); k++) {
        MsmPiece* p = &pieces[k];
        Slot* sl = leases[k].s;
        XchgArea* view = nullptr;
        if (area) {
            views[k].base = area->base;
            views[k].host_canon = area->host_canon;
            views[k].n_parts = area->n_parts;
            views[k].part = (uint32_t)k;
            views[k].seq = area->seq;
            view = &views[k];
        }
        uint8_t* out = out_affine + (size_t)p->col0 * 96;
        jobs.push_back({p->c, [p, sl, &src, scalar_fmt, view, out]() -> int32_t {
                            return msm_host_device(*p->c, *sl, &p->sh->t, p->bases_off, src, p->col0, p->ncols, p->i0, p->n, scalar_fmt, view, out);
                        }});
    }
    int32_t rc = run_jobs(jobs);
    if (area) {
        if (rc == B200ZK_OK) memcpy(out_affine, area->host_canon, (size_t)batch * 96);   // every stream has been waited for
        release_area(area);
    }
    return rc;
}

// ------------------------------------------------------------------------------------------
// host-buffer NTT on one device: polynomials [p0, p0 + cnt) of the call
// ------------------------------------------------------------------------------------------
struct PolySrc {
    uint8_t* const* ptrs = nullptr;
    uint8_t* base = nullptr;
};
int32_t ntt_host_device(Ctx& c, Slot& sl, const PolySrc& src, uint32_t p0, uint32_t cnt, uint32_t log_n, const uint8_t omega[32],
                        uint32_t flags, const uint8_t* coset_shift) {
    if (cnt == 0) return B200ZK_OK;
    const size_t poly = ((size_t)1 << log_n) * 32;
    const size_t bytes = poly * cnt;
    TRY(sl.ntt_data.ensure(bytes));
    uint8_t* d = sl.ntt_data.as<uint8_t>();
    cudaStream_t s = sl.stream;
    auto copy_group = [&](uint32_t b0, uint32_t b1, bool up, cudaStream_t st) -> int32_t {
        if (!src.ptrs) {
            uint8_t* h = src.base + (size_t)(p0 + b0) * poly;
            if (up) CU(cudaMemcpyAsync(d + b0 * poly, h, (size_t)(b1 - b0) * poly, cudaMemcpyHostToDevice, st));
            else CU(cudaMemcpyAsync(h, d + b0 * poly, (size_t)(b1 - b0) * poly, cudaMemcpyDeviceToHost, st));
            return B200ZK_OK;
        }
        for (uint32_t j = b0; j < b1; j++) {
            if (up) CU(cudaMemcpyAsync(d + j * poly, src.ptrs[p0 + j], poly, cudaMemcpyHostToDevice, st));
            else CU(cudaMemcpyAsync(src.ptrs[p0 + j], d + j * poly, poly, cudaMemcpyDeviceToHost, st));
        }
        return B200ZK_OK;
    };
    if (cnt >= 4 && g_ntt_pipe_min > 0 && bytes >= (size_t)g_ntt_pipe_min) {
        // A transform is PCIe-bound end to end (2^22: 2.4 ms up, 0.94 ms of kernels, 2.4 ms down), and the link is full
        // duplex: group g+1 goes up and group g-1 comes down while group g is transformed.
        TRY(slot_streams(sl));
        const uint32_t groups = std::min<uint32_t>(8, cnt / 2);
        while (sl.pipe_up.size() < groups) {
            cudaEvent_t e1, e2;
            CU(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            sl.pipe_up.push_back(e1);
            sl.pipe_done.push_back(e2);
        }
        for (uint32_t k = 0; k < groups; k++) {
            const uint32_t b0 = (uint32_t)((uint64_t)cnt * k / groups), b1 = (uint32_t)((uint64_t)cnt * (k + 1) / groups);
            TRY(copy_group(b0, b1, true, sl.copy_stream));
            CU(cudaEventRecord(sl.pipe_up[k], sl.copy_stream));
            CU(cudaStreamWaitEvent(s, sl.pipe_up[k], 0));
            TRY(ntt_run(c, sl, reinterpret_cast<uint32_t*>(d + b0 * poly), b1 - b0, log_n, omega, flags, coset_shift, s));
            CU(cudaEventRecord(sl.pipe_done[k], s));
            CU(cudaStreamWaitEvent(sl.down_stream, sl.pipe_done[k], 0));
            TRY(copy_group(b0, b1, false, sl.down_stream));
        }
        CU(cudaStreamSynchronize(sl.down_stream));
        CU(cudaStreamSynchronize(s));
        return B200ZK_OK;
    }
    TRY(copy_group(0, cnt, true, s));
    TRY(ntt_run(c, sl, sl.ntt_data.as<uint32_t>(), cnt, log_n, omega, flags, coset_shift, s));
    TRY(copy_group(0, cnt, false, s));
    CU(cudaStreamSynchronize(s));
    return B200ZK_OK;
}

// ------------------------------------------------------------------------------------------
// one large transform over all bound GPUs (worth it at 2^23..2^24 because the transform is bound by the integer
// pipe): four-step decomposition n = R * C,
//   X[k1 + R k2] = sum_{i2} (w^R)^(i2 k2) * w^(i2 k1) * sum_{i1} (w^C)^(i1 k1) x[i1 C + i2].
// GPU g takes the columns i2 in [g C/G, (g+1) C/G) -- from host memory that is a strided copy, the DMA does the first
// transposition -- runs the R-point column transforms in ONE pass (R <= 2^11) whose store multiplies by w^(i2 k1) and writes
// row k1 straight into the HBM of the GPU that owns it (the single exchange, fused into the kernel as peer stores); after a
// cross-GPU event every GPU runs its R/G row transforms of length C, the last pass storing transposed so that natural
// order is again one strided copy away.  Layouts of the resident form: in [R][C/G] per GPU, out [C][R/G] per GPU.
// ------------------------------------------------------------------------------------------
struct HostBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n, waiting = 0, gen = 0;
    bool failed = false;
    explicit HostBarrier(int parties) : n(parties) {}
    // returns false if any party reported a failure (everybody then leaves)
    bool arrive(bool ok) {
        std::unique_lock<std::mutex> lk(mu);
        if (!ok) failed = true;
        int g = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
        return !failed;
    }
};

void ntt_shard_shape(uint32_t log_n, int parts, uint32_t* log_r, uint32_t* log_c) {
    uint32_t lg = 0;
    while ((1 << lg) < parts) lg++;
    uint32_t r = std::min<uint32_t>(NTT_LOGB, log_n - lg);   // C >= parts
    *log_r = r;
    *log_c = log_n - r;
}

int32_t ntt_get_shard_plan(Ctx& c, uint32_t log_n, const uint8_t omega[32], bool inverse, int parts, int part, ShardPlan** out) {
    std::lock_guard<std::mutex> lk(c.mu);
    ShardKey key;
    key.log_n = log_n; key.inverse = inverse ? 1 : 0; key.parts = (uint32_t)parts; key.part = (uint32_t)part;
    memcpy(key.omega, omega, 32);
    auto it = c.shard_plans.find(key);
    if (it != c.shard_plans.end()) { *out = &it->second; return B200ZK_OK; }
    cudaStream_t s = c.stream;
    ShardPlan pl;
    ntt_shard_shape(log_n, parts, &pl.log_r, &pl.log_c);
    const uint64_t R = 1ull << pl.log_r, C = 1ull << pl.log_c, Cl = C / parts;
    uint32_t *d_omega = nullptr, *d_ninv = nullptr;
    CU(cudaMalloc(&d_omega, 64));
    d_ninv = d_omega + 8;
    TRY(upload_fr_mont(omega, d_omega, s));
    if (inverse) LAUNCH(fr_inv_pow2_kernel, 1, 1, 0, s, d_ninv, log_n);
    uint64_t cnt = pl.log_r ? (R >> 1) : 1;
    CU(cudaMalloc(&pl.tw_local, cnt * 32));
    LAUNCH(fr_powers_kernel, (unsigned)((cnt + 127) / 128), 128, 0, s, pl.tw_local, (const uint32_t*)d_omega, (const uint32_t*)nullptr, cnt, C, 0u, 0u);
    CU(cudaMalloc(&pl.tw1, Cl * R * 32));
    LAUNCH(fr_power_block_kernel, (unsigned)((Cl * R + 127) / 128), 128, 0, s, pl.tw1, (const uint32_t*)d_omega,
           inverse ? (const uint32_t*)d_ninv : (const uint32_t*)nullptr, Cl * R, pl.log_r, (uint64_t)part * Cl);
    CU(cudaStreamSynchronize(s));
    CU(cudaFree(d_omega));
    auto ins = c.shard_plans.emplace(key, pl);
    *out = &ins.first->second;
    return B200ZK_OK;
}

int32_t ntt_get_shard_coset(Ctx& c, uint32_t log_n, const uint8_t shift[32], int parts, int part, bool out_side, uint32_t** out) {
    std::lock_guard<std::mutex> lk(c.mu);
    ShardCosetKey key;
    key.log_n = log_n; key.parts = (uint32_t)parts; key.part = (uint32_t)part; key.out = out_side ? 1 : 0;
    memcpy(key.shift, shift, 32);
    auto it = c.shard_cosets.find(key);
    if (it != c.shard_cosets.end()) { *out = it->second; return B200ZK_OK; }
    cudaStream_t s = c.stream;
    uint32_t log_r, log_c;
    ntt_shard_shape(log_n, parts, &log_r, &log_c);
    const uint64_t R = 1ull << log_r, C = 1ull << log_c, Cl = C / parts, Rl = R / parts;
    uint32_t *d_shift = nullptr, *tab = nullptr;
    CU(cudaMalloc(&d_shift, 32));
    TRY(upload_fr_mont(shift, d_shift, s));
    CU(cudaMalloc(&tab, (R * C / parts) * 32));
    if (!out_side)   // like the input slice [i1][u]: shift^(i1 C + part Cl + u)
        LAUNCH(fr_power_grid_kernel, (unsigned)((R * Cl + 127) / 128), 128, 0, s, tab, (const uint32_t*)d_shift, R, Cl, C, (uint64_t)1, (uint64_t)part * Cl);
    else             // like the output slice [k2][k1l]: shift^(k2 R + part Rl + k1l)
        LAUNCH(fr_power_grid_kernel, (unsigned)((C * Rl + 127) / 128), 128, 0, s, tab, (const uint32_t*)d_shift, C, Rl, R, (uint64_t)1, (uint64_t)part * Rl);
    CU(cudaStreamSynchronize(s));
    CU(cudaFree(d_shift));
    c.shard_cosets[key] = tab;
    *out = tab;
    return B200ZK_OK;
}

// omega^(2^k) as canonical bytes, on the host with the kernels' field code
void fr_pow2k_host(const uint8_t in[32], uint32_t k, uint8_t out[32]) {
    Fr v;
    memcpy(v.l, in, 32);
    fe_reduce_loose(v);
    v = fe_to_mont(v);
    for (uint32_t i = 0; i < k; i++) v = fe_sqr(v);
    v = fe_from_mont(v);
    memcpy(out, v.l, 32);
}

struct ShardIO {
    uint8_t* host = nullptr;                 // natural-order host buffer (in place), or
    void* const* d_in = nullptr;             // per GPU: [R][C/G] column block ...
    void* const* d_out = nullptr;            // ... and [C][R/G] result block (may alias d_in)
};

int32_t ntt_sharded(const ShardIO& io, uint32_t log_n, const uint8_t omega[32], uint32_t flags, const uint8_t* coset_shift) {
    TRY(ntt_check_args(log_n, flags, coset_shift));
    const int D = (int)g_devs.size();
    if (D < 2 || (D & (D - 1))) return fail(B200ZK_ERR_INVALID_ARG, "ntt: a sharded transform needs 2, 4, 8 or 16 bound GPUs");
    uint32_t log_r, log_c;
    ntt_shard_shape(log_n, D, &log_r, &log_c);
    const uint64_t R = 1ull << log_r, C = 1ull << log_c;
    if (R < (uint64_t)D || C < (uint64_t)D) return fail(B200ZK_ERR_INVALID_ARG, "ntt: transform too small to shard over the bound GPUs");
    const uint64_t Cl = C / D, Rl = R / D;
    const bool inverse = (flags & B200ZK_NTT_INVERSE_SCALE) != 0;
    uint8_t omega2[32];
    fr_pow2k_host(omega, log_r, omega2);     // w^R generates the row transforms
    std::vector<SlotLease> leases;
    for (auto& c : g_devs) leases.emplace_back(*c);
    std::vector<uint32_t*> peerB(D, nullptr);
    HostBarrier bar(D);
    std::vector<Job> jobs;
    for (int g = 0; g < D; g++) {
        Ctx* c = g_devs[g].get();
        Slot* sl = leases[g].s;
        jobs.push_back({c, [=, &peerB, &bar, &leases, &io]() -> int32_t {
            cudaStream_t s = sl->stream;
            const size_t slice = (size_t)(R * Cl) * 32;          // = n / D elements
            ShardPlan* pl = nullptr;
            uint32_t *cos_in = nullptr, *cos_out = nullptr;
            // ---- phase 0: buffers and tables; publish where the peers must store
            auto phase0 = [&]() -> int32_t {
                TRY(sl->ntt_b.ensure(slice));
                if (io.host) TRY(sl->ntt_data.ensure(slice));
                TRY(ntt_get_shard_plan(*c, log_n, omega, inverse, D, g, &pl));
                if (flags & B200ZK_NTT_COSET_IN) TRY(ntt_get_shard_coset(*c, log_n, coset_shift, D, g, false, &cos_in));
                if (flags & B200ZK_NTT_COSET_OUT) TRY(ntt_get_shard_coset(*c, log_n, coset_shift, D, g, true, &cos_out));
                if (!sl->xdev_ev) CU(cudaEventCreateWithFlags(&sl->xdev_ev, cudaEventDisableTiming));
                TRY(ntt_set_attrs(*c));
                peerB[g] = sl->ntt_b.as<uint32_t>();
                return B200ZK_OK;
            };
            int32_t rc = phase0();
            if (!bar.arrive(rc == B200ZK_OK)) return rc != B200ZK_OK ? rc : fail(B200ZK_ERR_CUDA, "ntt: another GPU failed to set up");
            // ---- phase 1: my columns up (strided), column pass with the exchange fused into its store
            uint32_t* A = io.host ? sl->ntt_data.as<uint32_t>() : reinterpret_cast<uint32_t*>(io.d_in[g]);
            uint32_t* OUT = io.host ? A : reinterpret_cast<uint32_t*>(io.d_out[g]);
            auto phase1 = [&]() -> int32_t {
                TRY(ws_enter(*sl, s));
                if (io.host)
                    CU(cudaMemcpy2DAsync(A, Cl * 32, io.host + (size_t)g * Cl * 32, C * 32, Cl * 32, R, cudaMemcpyHostToDevice, s));
                NttPassArgs a;
                memset(&a, 0, sizeof a);
                a.in = A;
                a.tw_local = pl->tw_local;
                a.tw_pass = pl->tw1;
                a.in_scale = cos_in;
                a.reduce_in = (flags & B200ZK_NTT_MONT) ? 0 : 1;
                a.log_n = log_r + (uint32_t)(log_c - (uint32_t)__builtin_ctz((unsigned)D));
                a.deg = log_r;
                a.log_s = 0;
                a.log_cols = a.log_n - log_r;
                a.total_cols = Cl;
                a.scatter = 1;
                a.log_rl = log_r - (uint32_t)__builtin_ctz((unsigned)D);
                a.log_c = log_c;
                a.col0 = (uint32_t)(g * Cl);
                for (int h = 0; h < D; h++) a.peer[h] = peerB[h];
                uint64_t ctas = (R * Cl + NTT_B - 1) / NTT_B;
                prof_select(*c, *sl, 2);
                TRY(prof_mark(*sl, 0, s));
                if (g_ntt_variant == 1) LAUNCH(ntt_pass_kernel_occ2, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
                else if (g_ntt_variant == 9) LAUNCH(ntt_pass_kernel_lb0_lazy, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
                else LAUNCH(ntt_pass_kernel_lb0, (unsigned)ctas, NTT_THREADS, NTT_SMEM, s, a);
                TRY(prof_mark(*sl, 1, s));
                CU(cudaEventRecord(sl->xdev_ev, s));
                return B200ZK_OK;
            };
            rc = phase1();
            if (!bar.arrive(rc == B200ZK_OK)) { cudaStreamSynchronize(s); return rc != B200ZK_OK ? rc : fail(B200ZK_ERR_CUDA, "ntt: another GPU failed in the column pass"); }
            // ---- phase 2: every peer has stored its part of my rows; row transforms, transposed store, strided copy down
            auto phase2 = [&]() -> int32_t {
                for (int h = 0; h < D; h++)
                    if (h != g) CU(cudaStreamWaitEvent(s, leases[h].s->xdev_ev, 0));
                NttExtra ex;
                ex.out_final = OUT;
                ex.t_out = 1 + (log_r - (uint32_t)__builtin_ctz((unsigned)D));
                ex.out_scale = cos_out;
                TRY(ntt_run(*c, *sl, sl->ntt_b.as<uint32_t>(), (uint32_t)Rl, log_c, omega2, B200ZK_NTT_MONT, nullptr, s, &ex));
                if (io.host)
                    CU(cudaMemcpy2DAsync(io.host + (size_t)g * Rl * 32, R * 32, OUT, Rl * 32, Rl * 32, C, cudaMemcpyDeviceToHost, s));
                return B200ZK_OK;
            };
            rc = phase2();
            cudaError_t e = cudaStreamSynchronize(s);
            if (rc == B200ZK_OK && e != cudaSuccess) rc = fail(B200ZK_ERR_CUDA, std::string("ntt: ") + cudaGetErrorString(e));
            // nobody reuses (or frees) a buffer its peers may still be writing or reading
            bar.arrive(rc == B200ZK_OK);
            return rc;
        }});
    }
    return run_jobs(jobs);
}

int32_t ntt_host(const PolySrc& src, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags, const uint8_t* coset_shift) {
    TRY(need_init());
    if ((!src.ptrs && !src.base) || !omega) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    TRY(ntt_check_args(log_n, flags, coset_shift));
    if (batch == 0) return B200ZK_OK;
    if (src.ptrs)
        for (uint32_t j = 0; j < batch; j++)
            if (!src.ptrs[j]) return fail(B200ZK_ERR_INVALID_ARG, "null polynomial");
    const int D = (int)g_devs.size();
    static int shard_min_log = -1;
    if (shard_min_log < 0) { const char* v = getenv("B200ZK_NTT_SHARD_MIN_LOG"); shard_min_log = v ? atoi(v) : 23; }
    if (batch == 1 && D > 1 && !(D & (D - 1)) && (int)log_n >= shard_min_log) {
        ShardIO io;
        io.host = src.ptrs ? src.ptrs[0] : src.base;
        return ntt_sharded(io, log_n, omega, flags, coset_shift);
    }
    // independent polynomials: contiguous blocks per GPU, no exchange; each GPU's PCIe link carries its own block
    int used = (int)std::min<uint32_t>(batch, (uint32_t)D);
    Ctx* first = nullptr;
    TRY(current_ctx(&first));
    struct Piece { Ctx* c; uint32_t p0, cnt; };
    std::vector<Piece> pieces;
    for (int k = 0; k < used; k++) {
        uint32_t b0 = (uint32_t)((uint64_t)batch * k / used), b1 = (uint32_t)((uint64_t)batch * (k + 1) / used);
        pieces.push_back({used == 1 ? first : g_devs[k].get(), b0, b1 - b0});
    }
    std::vector<SlotLease> leases;
    for (auto& p : pieces) leases.emplace_back(*p.c);
    std::vector<Job> jobs;
    for (size_t k = 0; k < pieces.size(); k++) {
        Piece* p = &pieces[k];
        Slot* sl = leases[k].s;
        jobs.push_back({p->c, [p, sl, &src, log_n, omega, flags, coset_shift]() -> int32_t {
                            return ntt_host_device(*p->c, *sl, src, p->p0, p->cnt, log_n, omega, flags, coset_shift);
                        }});
    }
    return run_jobs(jobs);
}

// ------------------------------------------------------------------------------------------
// registration of one shard on one device
// ------------------------------------------------------------------------------------------
// h_src: host source of the whole table (or null); d_src / src_ordinal: device source of the whole table (or null)
int32_t register_shard(Ctx& c, TableShard& sh, const uint8_t* h_src, const uint8_t* d_src, int src_ordinal, uint32_t fmt_flags,
                       uint32_t stride) {
    BaseTable& t = sh.t;
    const uint64_t n = sh.n;
    t.n = n;
    uint32_t fmt = fmt_flags & 0xffu;
    cudaStream_t s = c.stream;
    // Window tables: W rows of n points.  Built for resident SRS tables of >= 1024 points (they are registered once and
    // committed against many times); skipped on request, for small tables, or when HBM is short.
    bool want_rows = !(fmt_flags & B200ZK_BASES_NO_WINDOW_TABLES) && !g_tune_no_tables && n >= 1024;
    uint32_t cbits = 0, W = 1;
    if (want_rows) {
        cbits = (g_tune_c >= 2 && g_tune_c <= 24) ? g_tune_c : table_window_bits(n);
        W = (256 + cbits - 1) / cbits;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if ((double)W * n * 96.0 > 0.45 * (double)free_b || (double)W * n >= 2147483648.0 || W > (uint32_t)MSM_MAX_ROWS) { W = 1; cbits = 0; }
    }
    const uint8_t* d_in = nullptr;
    if (n) {
        size_t bytes = (size_t)(n - 1) * stride + 96;
        const size_t off = (size_t)sh.start * stride;
        if (h_src) {
            TRY(c.reg_stage.ensure(bytes + 16));
            CU(cudaMemcpyAsync(c.reg_stage.p, h_src + off, bytes, cudaMemcpyHostToDevice, s));
            d_in = c.reg_stage.as<uint8_t>();
        } else if (src_ordinal == c.ordinal) {
            d_in = d_src + off;
        } else {
            TRY(c.reg_stage.ensure(bytes + 16));
            CU(cudaMemcpyPeerAsync(c.reg_stage.p, c.ordinal, d_src + off, src_ordinal, bytes, s));
            d_in = c.reg_stage.as<uint8_t>();
        }
    }
    CU(cudaMalloc(&t.d, std::max<uint64_t>(n, 1) * 96 * W));
    int32_t rc = n ? ingest_bases(c, d_in, n, fmt, stride, t.d, s) : B200ZK_OK;
    if (rc != B200ZK_OK) { cudaFree(t.d); t.d = nullptr; return rc; }
    if (W > 1) {
        g1_window_tables_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(t.d, n, cbits, W);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { cudaFree(t.d); t.d = nullptr; return fail(B200ZK_ERR_CUDA, std::string("window tables: ") + cudaGetErrorString(e)); }
        t.rows = W;
        t.c = cbits;
    }
    return B200ZK_OK;
}

int32_t register_common(const uint8_t* h_src, const uint8_t* d_src, uint64_t n, uint32_t fmt_flags, uint32_t stride_bytes,
                        uint64_t* out_handle) {
    TRY(need_init());
    if (!out_handle || (!h_src && !d_src && n)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if ((fmt_flags & 0xffu) > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown point format");
    if ((fmt_flags & B200ZK_BASES_SHARD) && (fmt_flags & B200ZK_BASES_REPLICATE))
        return fail(B200ZK_ERR_INVALID_ARG, "B200ZK_BASES_SHARD and B200ZK_BASES_REPLICATE are mutually exclusive");
    uint32_t stride = stride_bytes ? stride_bytes : 96;
    if (stride < 96 || (stride & 3)) return fail(B200ZK_ERR_INVALID_ARG, "stride must be >= 96 and a multiple of 4");
    const int D = (int)g_devs.size();
    int src_ordinal = -1;
    Ctx* home = nullptr;
    if (d_src) {
        TRY(ctx_for_pointer(d_src, &home));
        src_ordinal = home->ordinal;
        DeviceScope ds(home->ordinal);
        CU(cudaDeviceSynchronize());   // the source may have been produced on another stream
    } else {
        TRY(current_ctx(&home));
    }
    auto ts = std::make_shared<TableSet>();
    ts->n = n;
    if (D == 1 || n < (uint64_t)D) {
        ts->shards.resize(1);
        ts->shards[0].dev = home->index;
        ts->shards[0].start = 0;
        ts->shards[0].n = n;
    } else {
        // small tables are replicated (a batch of columns is dealt out over the GPUs), large ones partitioned by point range
        bool replicate = (fmt_flags & B200ZK_BASES_REPLICATE) != 0;
        if (!(fmt_flags & (B200ZK_BASES_REPLICATE | B200ZK_BASES_SHARD))) {
            uint32_t W = (fmt_flags & B200ZK_BASES_NO_WINDOW_TABLES) ? 1 : 16;
            replicate = (double)n * 96.0 * W <= (double)g_replicate_max_bytes;
        }
        ts->replicated = replicate;
        ts->shards.resize(D);
        for (int k = 0; k < D; k++) {
            ts->shards[k].dev = k;
            ts->shards[k].start = replicate ? 0 : n * k / D;
            ts->shards[k].n = replicate ? n : n * (k + 1) / D - n * k / D;
        }
    }
    std::vector<Job> jobs;
    for (auto& shr : ts->shards) {
        TableShard* sh = &shr;
        Ctx* c = g_devs[sh->dev].get();
        jobs.push_back({c, [c, sh, h_src, d_src, src_ordinal, fmt_flags, stride]() -> int32_t {
                            std::lock_guard<std::mutex> lk(c->mu);   // c.reg_stage / c.reg_flag / c.stream
                            return register_shard(*c, *sh, h_src, d_src, src_ordinal, fmt_flags, stride);
                        }});
    }
    int32_t rc = run_jobs(jobs);
    if (rc != B200ZK_OK) {
        for (auto& sh : ts->shards)
            if (sh.t.d) { DeviceScope ds(g_devs[sh.dev]->ordinal); cudaFree(sh.t.d); }
        return rc;
    }
    std::lock_guard<std::mutex> lk(g_tab_mu);
    uint64_t h = g_next_handle++;
    g_tables[h] = ts;
    *out_handle = h;
    return B200ZK_OK;
}

int32_t init_devices_locked(const int32_t* ids, int32_t n) {
    if (!g_devs.empty()) return B200ZK_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(B200ZK_ERR_NO_DEVICE,
                    std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    }
    if (n < 1 || n > (int32_t)XCHG_MAX_PARTS) return fail(B200ZK_ERR_INVALID_ARG, "between 1 and 16 devices can be bound");
    std::vector<int> ords;
    for (int i = 0; i < n; i++) {
        int d = ids ? ids[i] : i;
        if (d < 0) {
            if (cudaGetDevice(&d) != cudaSuccess) d = 0;
        }
        if (d >= count) return fail(B200ZK_ERR_NO_DEVICE, "device index out of range");
        if (std::find(ords.begin(), ords.end(), d) != ords.end()) return fail(B200ZK_ERR_INVALID_ARG, "a device is listed twice");
        ords.push_back(d);
    }
    int prev = -1;
    cudaGetDevice(&prev);
    read_env();
    std::vector<std::unique_ptr<Ctx>> devs;
    auto undo = [&](int32_t rc) {
        for (auto& c : devs) {
            cudaSetDevice(c->ordinal);
            for (Slot& sl : c->slots)
                if (sl.stream) cudaStreamDestroy(sl.stream);
            if (c->stream) cudaStreamDestroy(c->stream);
        }
        if (prev >= 0) cudaSetDevice(prev);
        return rc;
    };
    for (int i = 0; i < n; i++) {
        auto c = std::make_unique<Ctx>();
        c->ordinal = ords[i];
        c->index = i;
        cudaError_t e2 = cudaSetDevice(c->ordinal);
        if (e2 == cudaSuccess) e2 = cudaGetDeviceProperties(&c->prop, c->ordinal);
        if (e2 != cudaSuccess) return undo(fail(B200ZK_ERR_CUDA, std::string("cannot bind device: ") + cudaGetErrorString(e2)));
        if (c->prop.major != 10)
            return undo(fail(B200ZK_ERR_NO_DEVICE, std::string("device '") + c->prop.name +
                                                       "' is not sm_100: this library ships sm_100a code only and has no fallback"));
        e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        for (int k = 0; k < NSLOT && e2 == cudaSuccess; k++) {
            c->slots[k].id = k;
            e2 = cudaStreamCreateWithFlags(&c->slots[k].stream, cudaStreamNonBlocking);
        }
        devs.push_back(std::move(c));
        if (e2 != cudaSuccess) return undo(fail(B200ZK_ERR_CUDA, std::string("stream creation failed: ") + cudaGetErrorString(e2)));
    }
    if (n > 1) {
        // every pair of bound GPUs can reach each other's HBM (NVLink / NVSwitch): the exchange of partial sums and the
        // transposes of the multi-GPU NTT are plain stores into a peer's memory
        for (int i = 0; i < n; i++) {
            cudaSetDevice(ords[i]);
            for (int j = 0; j < n; j++) {
                if (i == j) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, ords[i], ords[j]);
                if (!can) return undo(fail(B200ZK_ERR_NO_DEVICE, "GPUs " + std::to_string(ords[i]) + " and " + std::to_string(ords[j]) +
                                                                     " have no peer access: the multi-GPU paths need NVLink / PCIe P2P"));
                cudaError_t e3 = cudaDeviceEnablePeerAccess(ords[j], 0);
                if (e3 == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e3 != cudaSuccess) return undo(fail(B200ZK_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e3)));
            }
        }
        cudaSetDevice(ords[0]);
        for (auto& a : g_areas) {
            cudaError_t e3 = cudaMalloc(&a.base, XchgArea::bytes());
            if (e3 == cudaSuccess) e3 = cudaMemset(a.base, 0, XchgArea::bytes());
            if (e3 == cudaSuccess) e3 = cudaHostAlloc((void**)&a.host_canon, 96 * (size_t)XCHG_MAX_COLS, cudaHostAllocPortable | cudaHostAllocMapped);
            if (e3 != cudaSuccess) return undo(fail(B200ZK_ERR_CUDA, std::string("exchange area: ") + cudaGetErrorString(e3)));
            a.owner = true;
            a.busy = false;
            a.seq = 0;
            a.home_ordinal = ords[0];
        }
    }
    g_devs = std::move(devs);
    if (g_devs.size() > 1)
        for (auto& c : g_devs) c->worker.start(c->ordinal);
    if (prev >= 0) cudaSetDevice(prev);
    return B200ZK_OK;
}

}  // namespace

// what the other translation units share with these contexts (ctx.hpp)
namespace b200zk_ctx {
int32_t fail(int32_t code, const std::string& msg) { return ::fail(code, msg); }
int32_t current(Dev** out) { return ::current_ctx(out); }
int32_t for_pointer(const void* p, Dev** out) { return ::ctx_for_pointer(p, out); }
int32_t by_index(int index, Dev** out) {
    XTRY(::need_init());
    if (index < 0 || index >= (int)g_devs.size()) return ::fail(B200ZK_ERR_BAD_HANDLE, "handle names a device that is not bound");
    *out = g_devs[index].get();
    return B200ZK_OK;
}
int device_count() { return (int)g_devs.size(); }
int ordinal(Dev* d) { return d->ordinal; }
int index(Dev* d) { return d->index; }
std::mutex& mutex(Dev* d) { return d->mu; }
cudaStream_t stream(Dev* d) { return d->stream; }
int sm_count(Dev* d) { return d->prop.multiProcessorCount; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int32_t generator_dev(Dev* d, uint32_t** out, cudaStream_t s) { return get_generator_dev(*d, out, s); }
void on_shutdown(void (*fn)()) {
    static std::mutex mu;   // the other translation units register from under their own locks
    std::lock_guard<std::mutex> lk(mu);
    g_hooks.push_back(fn);
}
int32_t ws_enter(Dev* d, cudaStream_t s) {
    if (!d->ws_event) XCU(cudaEventCreateWithFlags(&d->ws_event, cudaEventDisableTiming));
    if (d->ws_used && s != d->ws_stream) XCU(cudaStreamWaitEvent(s, d->ws_event, 0));
    return B200ZK_OK;
}
int32_t ws_leave(Dev* d, cudaStream_t s) {
    XCU(cudaEventRecord(d->ws_event, s));
    d->ws_stream = s;
    d->ws_used = true;
    return B200ZK_OK;
}
}  // namespace b200zk_ctx

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int32_t b200zk_init_devices(const int32_t* device_ids, int32_t n_devices) {
    std::lock_guard<std::mutex> lk(g_mu);
    return init_devices_locked(device_ids, n_devices);
}

int32_t b200zk_init(int32_t device) {
    std::lock_guard<std::mutex> lk(g_mu);
    return init_devices_locked(&device, 1);
}

int32_t b200zk_device_count(void) { return (int32_t)g_devs.size(); }

int32_t b200zk_set_device(int32_t index) {
    TRY(need_init());
    if (index < 0 || index >= (int32_t)g_devs.size()) return fail(B200ZK_ERR_INVALID_ARG, "device index outside the bound devices");
    t_dev_index = index;
    return B200ZK_OK;
}

int32_t b200zk_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_devs.empty()) return B200ZK_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    for (auto& c : g_devs) c->worker.shutdown();
    for (auto& c : g_devs) {
        cudaSetDevice(c->ordinal);
        cudaDeviceSynchronize();
    }
    for (auto fn : g_hooks) fn();
    {
        std::lock_guard<std::mutex> lt(g_tab_mu);
        for (auto& kv : g_tables)
            for (auto& sh : kv.second->shards) {
                cudaSetDevice(g_devs[sh.dev]->ordinal);
                cudaFree(sh.t.d);
            }
        g_tables.clear();
    }
    {
        std::lock_guard<std::mutex> la(g_area_mu);
        for (auto& kv : g_xchg) {
            XchgArea& a = *kv.second;
            cudaSetDevice(a.home_ordinal);
            if (a.ipc && a.base) cudaIpcCloseMemHandle(a.base);
            else if (a.base) cudaFree(a.base);
            a.status.release();
        }
        g_xchg.clear();
        for (auto& a : g_areas) {
            if (a.base) { cudaSetDevice(a.home_ordinal); cudaFree(a.base); }
            if (a.host_canon) cudaFreeHost(a.host_canon);
            a = XchgArea();
        }
    }
    for (auto& cp : g_devs) {
        Ctx& c = *cp;
        cudaSetDevice(c.ordinal);
        for (auto& kv : c.ntt_plans) {
            for (int i = 0; i < 3; i++) {
                if (kv.second.tw_local[i]) cudaFree(kv.second.tw_local[i]);
                if (kv.second.tw_pass[i]) cudaFree(kv.second.tw_pass[i]);
            }
            if (kv.second.ninv) cudaFree(kv.second.ninv);
        }
        for (auto& kv : c.coset_tables) cudaFree(kv.second);
        for (auto& kv : c.shard_plans) { cudaFree(kv.second.tw_local); cudaFree(kv.second.tw1); }
        for (auto& kv : c.shard_cosets) cudaFree(kv.second);
        if (c.fixed_table) cudaFree(c.fixed_table);
        if (c.d_gen) cudaFree(c.d_gen);
        c.reg_stage.release();
        c.reg_flag.release();
        for (Slot& sl : c.slots) {
            for (DevBuf* b : sl.all()) b->release();
            sl.h_out.release();
            for (cudaEvent_t e : sl.pipe_up) cudaEventDestroy(e);
            for (cudaEvent_t e : sl.pipe_done) cudaEventDestroy(e);
            for (cudaEvent_t e : sl.ev)
                if (e) cudaEventDestroy(e);
            if (sl.copy_stream) {
                cudaStreamDestroy(sl.copy_stream);
                cudaEventDestroy(sl.copy_ev[0]);
                cudaEventDestroy(sl.copy_ev[1]);
            }
            if (sl.down_stream) cudaStreamDestroy(sl.down_stream);
            if (sl.ws_event) cudaEventDestroy(sl.ws_event);
            if (sl.xdev_ev) cudaEventDestroy(sl.xdev_ev);
            for (cudaEvent_t e : {sl.ev_fork, sl.ev_sorted, sl.ev_acc, sl.ev_done})
                if (e) cudaEventDestroy(e);
            if (sl.hi) cudaStreamDestroy(sl.hi);
            if (sl.lo) cudaStreamDestroy(sl.lo);
            if (sl.stream) cudaStreamDestroy(sl.stream);
        }
        if (c.ws_event) cudaEventDestroy(c.ws_event);
        if (c.stream) cudaStreamDestroy(c.stream);
    }
    g_devs.clear();
    g_prof_ctx = nullptr;
    g_prof_slot = nullptr;
    if (prev >= 0) cudaSetDevice(prev);
    return B200ZK_OK;
}

int32_t b200zk_last_error(char* buf, size_t len) {
    if (!buf || len == 0) return B200ZK_ERR_INVALID_ARG;
    snprintf(buf, len, "%s", t_err.c_str());
    return B200ZK_OK;
}

int32_t b200zk_device_info(char* buf, size_t len) {
    Ctx* c = nullptr;
    TRY(current_ctx(&c));
    if (!buf || len == 0) return fail(B200ZK_ERR_INVALID_ARG, "null buffer");
    snprintf(buf, len, "%s sm_%d%d %dSM", c->prop.name, c->prop.major, c->prop.minor, c->prop.multiProcessorCount);
    return B200ZK_OK;
}

int32_t b200zk_host_alloc(void** out, size_t bytes) {
    TRY(need_init());
    if (!out) return fail(B200ZK_ERR_INVALID_ARG, "null out pointer");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));   // pinned for every bound GPU
    return B200ZK_OK;
}
int32_t b200zk_host_free(void* p) {
    TRY(need_init());
    if (p) CU(cudaFreeHost(p));
    return B200ZK_OK;
}

int32_t b200zk_bases_register(const uint8_t* g1_affine, uint64_t n, uint32_t fmt, uint32_t stride_bytes, uint64_t* out_handle) {
    return register_common(g1_affine, nullptr, n, fmt, stride_bytes, out_handle);
}

int32_t b200zk_bases_register_dev(const void* d_g1_affine, uint64_t n, uint32_t fmt, uint32_t stride_bytes, uint64_t* out_handle) {
    return register_common(nullptr, reinterpret_cast<const uint8_t*>(d_g1_affine), n, fmt, stride_bytes, out_handle);
}

int32_t b200zk_bases_release(uint64_t handle) {
    TRY(need_init());
    std::shared_ptr<TableSet> ts;
    {
        std::lock_guard<std::mutex> lk(g_tab_mu);
        auto it = g_tables.find(handle);
        if (it == g_tables.end()) return fail(B200ZK_ERR_BAD_HANDLE, "unknown base-table handle");
        ts = it->second;
        g_tables.erase(it);
    }
    for (auto& sh : ts->shards) {
        DeviceScope ds(g_devs[sh.dev]->ordinal);
        CU(cudaDeviceSynchronize());
        cudaFree(sh.t.d);
        sh.t.d = nullptr;
    }
    return B200ZK_OK;
}

int32_t b200zk_bases_read(uint64_t handle, uint64_t start, uint64_t n, uint8_t* out_affine) {
    TRY(need_init());
    std::shared_ptr<TableSet> ts;
    TRY(get_table(handle, start, n, &ts));
    if (n == 0) return B200ZK_OK;
    if (!out_affine) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    for (const TableShard& sh : ts->shards) {
        uint64_t lo = std::max(start, sh.start), hi = std::min(start + n, sh.start + sh.n);
        if (hi <= lo) continue;
        Ctx& c = *g_devs[sh.dev];
        DeviceScope ds(c.ordinal);
        std::lock_guard<std::mutex> lk(c.mu);
        uint64_t cnt = hi - lo;
        TRY(c.reg_stage.ensure(cnt * 96));
        LAUNCH(g1_export_kernel, (unsigned)((cnt + 127) / 128), 128, 0, c.stream, (const uint32_t*)(sh.t.d + 24 * (lo - sh.start)), cnt,
               c.reg_stage.as<uint32_t>());
        CU(cudaMemcpyAsync(out_affine + (lo - start) * 96, c.reg_stage.p, cnt * 96, cudaMemcpyDeviceToHost, c.stream));
        CU(cudaStreamSynchronize(c.stream));
        if (ts->replicated) break;
    }
    return B200ZK_OK;
}

int32_t b200zk_msm_g1_batch(uint64_t bases, uint64_t offset, const uint8_t* scalars, uint64_t n, uint32_t batch,
                            uint32_t scalar_fmt, uint8_t* out_affine) {
    ColSrc src;
    src.base = scalars;
    src.stride = (size_t)n * 32;
    return msm_host(bases, offset, src, n, batch, scalar_fmt, out_affine);
}

int32_t b200zk_msm_g1_batch_ptrs(uint64_t bases, uint64_t offset, const uint8_t* const* scalars, uint64_t n, uint32_t batch,
                                 uint32_t scalar_fmt, uint8_t* out_affine) {
    if (!scalars && batch) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    ColSrc src;
    src.ptrs = scalars;
    return msm_host(bases, offset, src, n, batch, scalar_fmt, out_affine);
}

int32_t b200zk_msm_g1(uint64_t bases, uint64_t offset, const uint8_t* scalars, uint64_t n, uint32_t scalar_fmt,
                      uint8_t out_affine[96]) {
    return b200zk_msm_g1_batch(bases, offset, scalars, n, 1, scalar_fmt, out_affine);
}

int32_t b200zk_bases_layout(uint64_t bases, uint32_t cap, uint32_t* out_n_shards, int32_t* out_device, uint64_t* out_start, uint64_t* out_n,
                            uint32_t* out_replicated) {
    TRY(need_init());
    if (!out_n_shards) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    auto ts = find_table(bases);
    if (!ts) return fail(B200ZK_ERR_BAD_HANDLE, "unknown base-table handle");
    *out_n_shards = (uint32_t)ts->shards.size();
    if (out_replicated) *out_replicated = ts->replicated ? 1u : 0u;
    for (uint32_t k = 0; k < ts->shards.size() && k < cap; k++) {
        if (out_device) out_device[k] = ts->shards[k].dev;
        if (out_start) out_start[k] = ts->shards[k].start;
        if (out_n) out_n[k] = ts->shards[k].n;
    }
    return B200ZK_OK;
}

int32_t b200zk_msm_g1_sharded_dev(uint64_t bases, const void* const* d_scalar_slices, uint32_t n_slices, uint32_t scalar_fmt,
                                  uint8_t out_affine[96]) {
    TRY(need_init());
    if (!d_scalar_slices || !out_affine) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    auto ts = find_table(bases);
    if (!ts) return fail(B200ZK_ERR_BAD_HANDLE, "unknown base-table handle");
    if (n_slices != ts->shards.size()) return fail(B200ZK_ERR_INVALID_ARG, "sharded_dev: one scalar slice per shard of the table");
    for (uint32_t k = 0; k < n_slices; k++)
        if ((!d_scalar_slices[k] && ts->shards[k].n) || ((uintptr_t)d_scalar_slices[k] & 15))
            return fail(B200ZK_ERR_INVALID_ARG, "sharded_dev: null or misaligned slice");
    ColSrc src;
    src.dev_slices = d_scalar_slices;
    return msm_host(bases, 0, src, ts->n, 1, scalar_fmt, out_affine);
}

// points [i0, i0 + n) of an ad-hoc sum on one device: upload, on-curve check, MSM; the result (or, with an exchange, the
// partial sum) leaves through the usual paths.  *bad = 1 if a point was rejected.
static int32_t adhoc_device(Ctx& c, Slot& sl, const uint8_t* g1_affine, uint32_t point_fmt, const uint8_t* scalars, uint32_t scalar_fmt,
                            uint64_t i0, uint64_t n, const XchgArea* area, uint8_t* out, uint32_t* bad) {
    cudaStream_t s = sl.stream;
    TRY(sl.stage.ensure(n * 96 + 16));
    TRY(sl.adhoc.ensure(n * 96 + 16));
    TRY(sl.scalars.ensure(n * 32 + 16));
    TRY(sl.out_canon.ensure(96));
    TRY(sl.flag.ensure(4));
    TRY(sl.h_out.ensure(128));
    TRY(ws_enter(sl, s));
    if (n) {
        CU(cudaMemcpyAsync(sl.stage.p, g1_affine + 96 * i0, n * 96, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(sl.scalars.p, scalars + 32 * i0, n * 32, cudaMemcpyHostToDevice, s));
    }
    CU(cudaMemsetAsync(sl.flag.p, 0, 4, s));
    if (n)
        LAUNCH(g1_ingest_kernel, (unsigned)((n + 127) / 128), 128, 0, s, (const uint8_t*)sl.stage.as<uint8_t>(), n, 96u,
               point_fmt == B200ZK_FMT_MONT ? 1u : 0u, sl.adhoc.as<uint32_t>(), sl.flag.as<uint32_t>());
    BaseTable t;
    t.d = sl.adhoc.as<uint32_t>();
    t.n = n;
    XchgArgs xa{};
    if (area) xa = area->args(0);
    TRY(msm_run(c, sl, &t, t.d, sl.scalars.as<uint32_t>(), n, 1, scalar_fmt, nullptr, area ? nullptr : sl.out_canon.as<uint32_t>(), s, nullptr,
                nullptr, area ? &xa : nullptr));
    uint8_t* h = reinterpret_cast<uint8_t*>(sl.h_out.p);
    if (!area) CU(cudaMemcpyAsync(h, sl.out_canon.p, 96, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h + 96, sl.flag.p, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    memcpy(bad, h + 96, 4);
    if (!area) memcpy(out, h, 96);
    return B200ZK_OK;
}

int32_t b200zk_msm_g1_adhoc(const uint8_t* g1_affine, uint32_t point_fmt, const uint8_t* scalars, uint32_t scalar_fmt,
                            uint64_t n, uint8_t out_affine[96]) {
    // One pass per GPU on a slot's stream with cached workspaces: no table registration, no allocation, one synchronisation
    // at the end (the verifier calls this once per proof or per batch).  A long sum (the 1024-proof batch of
    // /root/reference/src/circuits/schnorr_circuit.rs:224-229) is split by point range over the bound GPUs, whose partial
    // sums meet like those of a sharded commitment.
    Ctx* cp = nullptr;
    TRY(current_ctx(&cp));
    if (!out_affine || ((!g1_affine || !scalars) && n)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (point_fmt > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown point format");
    if (scalar_fmt > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown scalar format");
    if (n == 0) { memset(out_affine, 0, 96); return B200ZK_OK; }
    const int D = (int)g_devs.size();
    static uint64_t split_min = 0;
    if (!split_min) { const char* v = getenv("B200ZK_ADHOC_SPLIT_MIN"); split_min = v ? (uint64_t)atoll(v) : 8192; if (!split_min) split_min = 1; }
    int32_t rc;
    uint32_t any_bad = 0;
    if (D == 1 || n < split_min) {
        DeviceScope ds(cp->ordinal);
        SlotLease lease(*cp);
        rc = adhoc_device(*cp, *lease.s, g1_affine, point_fmt, scalars, scalar_fmt, 0, n, nullptr, out_affine, &any_bad);
    } else {
        std::vector<SlotLease> leases;
        for (auto& c : g_devs) leases.emplace_back(*c);
        XchgArea* area = acquire_area();
        area->seq++;
        area->n_parts = (uint32_t)D;
        std::vector<XchgArea> views(D);
        std::vector<uint32_t> bad(D, 0);
        std::vector<Job> jobs;
        for (int k = 0; k < D; k++) {
            views[k].base = area->base;
            views[k].host_canon = area->host_canon;
            views[k].n_parts = area->n_parts;
            views[k].part = (uint32_t)k;
            views[k].seq = area->seq;
            uint64_t lo = n * k / D, hi = n * (k + 1) / D;
            Ctx* c = g_devs[k].get();
            Slot* sl = leases[k].s;
            XchgArea* view = &views[k];
            uint32_t* b = &bad[k];
            jobs.push_back({c, [=]() -> int32_t {
                                return adhoc_device(*c, *sl, g1_affine, point_fmt, scalars, scalar_fmt, lo, hi - lo, view, nullptr, b);
                            }});
        }
        rc = run_jobs(jobs);
        if (rc == B200ZK_OK) memcpy(out_affine, area->host_canon, 96);
        release_area(area);
        for (uint32_t b : bad) any_bad |= b;
    }
    if (rc != B200ZK_OK) return rc;
    if (any_bad) { memset(out_affine, 0, 96); return fail(B200ZK_ERR_BAD_POINT, "a base point is not a canonical point on the curve"); }
    return B200ZK_OK;
}

// the device a "_dev" MSM runs on is the one that owns the scalars; the table must have the slice resident there
static int32_t msm_dev_common(uint64_t bases, uint64_t offset, const void* d_scalars, uint64_t n, uint32_t batch, uint32_t scalar_fmt,
                              void* d_out_mont, void* d_out_canon, void* d_out_xyzz, void* stream, const XchgArgs* xa) {
    if (scalar_fmt > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown scalar format");
    if (((uintptr_t)d_scalars | (uintptr_t)d_out_mont | (uintptr_t)d_out_canon | (uintptr_t)d_out_xyzz) & 15)
        return fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    Ctx* cp = nullptr;
    const void* hint = d_scalars ? d_scalars : (d_out_mont ? d_out_mont : (d_out_canon ? d_out_canon : d_out_xyzz));
    TRY(ctx_for_pointer(hint, &cp));
    std::shared_ptr<TableSet> ts;
    TRY(get_table(bases, offset, n, &ts));
    const TableShard* sh = shard_on(*ts, cp->index, offset, n);
    if (!sh) return fail(B200ZK_ERR_INVALID_ARG, "msm: this slice of the table is not resident on the GPU that owns the scalars");
    DeviceScope ds(cp->ordinal);
    SlotLease lease(*cp);
    if (!d_out_xyzz && batch >= 2)
        return msm_run_batch(*cp, *lease.s, &sh->t, sh->t.d + 24 * (offset - sh->start), reinterpret_cast<const uint32_t*>(d_scalars), n,
                             batch, scalar_fmt, reinterpret_cast<uint32_t*>(d_out_mont), reinterpret_cast<uint32_t*>(d_out_canon),
                             reinterpret_cast<cudaStream_t>(stream), xa);
    return msm_run(*cp, *lease.s, &sh->t, sh->t.d + 24 * (offset - sh->start), reinterpret_cast<const uint32_t*>(d_scalars), n, batch,
                   scalar_fmt, reinterpret_cast<uint32_t*>(d_out_mont), reinterpret_cast<uint32_t*>(d_out_canon),
                   reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<uint32_t*>(d_out_xyzz), nullptr, xa);
}

int32_t b200zk_msm_g1_dev(uint64_t bases, uint64_t offset, const void* d_scalars, uint64_t n, uint32_t batch,
                          uint32_t scalar_fmt, void* d_out_mont, void* d_out_canon, void* stream) {
    TRY(need_init());
    if ((!d_scalars && n && batch) || (!d_out_mont && !d_out_canon)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    return msm_dev_common(bases, offset, d_scalars, n, batch, scalar_fmt, d_out_mont, d_out_canon, nullptr, stream, nullptr);
}

int32_t b200zk_msm_g1_partial_dev(uint64_t bases, uint64_t offset, const void* d_scalars, uint64_t n, uint32_t scalar_fmt,
                                  void* d_out_xyzz, void* stream) {
    TRY(need_init());
    if ((!d_scalars && n) || !d_out_xyzz) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    return msm_dev_common(bases, offset, d_scalars, n, 1, scalar_fmt, nullptr, nullptr, d_out_xyzz, stream, nullptr);
}

int32_t b200zk_g1_sum_partials_dev(const void* d_partials_xyzz, uint32_t n, void* d_out_mont, void* d_out_canon, void* stream) {
    Ctx* c = nullptr;
    TRY(ctx_for_pointer(d_partials_xyzz, &c));
    if ((!d_partials_xyzz && n) || (!d_out_mont && !d_out_canon)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    DeviceScope ds(c->ordinal);
    LAUNCH(g1_sum_xyzz_kernel, 1, 32, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const uint32_t*>(d_partials_xyzz),
           n, reinterpret_cast<uint32_t*>(d_out_mont), reinterpret_cast<uint32_t*>(d_out_canon));
    return B200ZK_OK;
}

int32_t b200zk_g1_sum_dev(const void* d_points_mont, uint32_t n, void* d_out_mont, void* d_out_canon, void* stream) {
    Ctx* c = nullptr;
    TRY(ctx_for_pointer(d_points_mont, &c));
    if ((!d_points_mont && n) || (!d_out_mont && !d_out_canon)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    DeviceScope ds(c->ordinal);
    LAUNCH(g1_sum_kernel, 1, 32, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const uint32_t*>(d_points_mont), n,
           reinterpret_cast<uint32_t*>(d_out_mont), reinterpret_cast<uint32_t*>(d_out_canon));
    return B200ZK_OK;
}

// ---- exchange between processes (one process per GPU) ------------------------------------------
int32_t b200zk_xchg_create(uint32_t n_parts, uint8_t out_ipc_handle[64], uint64_t* out_xchg) {
    Ctx* c = nullptr;
    TRY(current_ctx(&c));
    if (!out_ipc_handle || !out_xchg) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (n_parts < 1 || n_parts > XCHG_MAX_PARTS) return fail(B200ZK_ERR_INVALID_ARG, "an exchange has 1..16 parts");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceScope ds(c->ordinal);
    auto a = std::make_unique<XchgArea>();
    CU(cudaMalloc(&a->base, XchgArea::bytes()));
    CU(cudaMemset(a->base, 0, XchgArea::bytes()));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, a->base));
    memcpy(out_ipc_handle, &h, 64);
    a->n_parts = n_parts;
    a->part = 0;
    a->owner = true;
    a->home_ordinal = c->ordinal;
    std::lock_guard<std::mutex> lk(g_area_mu);
    *out_xchg = g_next_xchg++;
    g_xchg[*out_xchg] = std::move(a);
    return B200ZK_OK;
}

int32_t b200zk_xchg_open(const uint8_t ipc_handle[64], uint32_t n_parts, uint32_t part, uint64_t* out_xchg) {
    Ctx* c = nullptr;
    TRY(current_ctx(&c));
    if (!ipc_handle || !out_xchg) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (n_parts < 1 || n_parts > XCHG_MAX_PARTS || part >= n_parts) return fail(B200ZK_ERR_INVALID_ARG, "bad part index");
    DeviceScope ds(c->ordinal);
    auto a = std::make_unique<XchgArea>();
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    CU(cudaIpcOpenMemHandle((void**)&a->base, h, cudaIpcMemLazyEnablePeerAccess));
    a->n_parts = n_parts;
    a->part = part;
    a->ipc = true;
    a->home_ordinal = c->ordinal;
    std::lock_guard<std::mutex> lk(g_area_mu);
    *out_xchg = g_next_xchg++;
    g_xchg[*out_xchg] = std::move(a);
    return B200ZK_OK;
}

int32_t b200zk_xchg_close(uint64_t xchg) {
    TRY(need_init());
    std::unique_ptr<XchgArea> a;
    {
        std::lock_guard<std::mutex> lk(g_area_mu);
        auto it = g_xchg.find(xchg);
        if (it == g_xchg.end()) return fail(B200ZK_ERR_BAD_HANDLE, "unknown exchange handle");
        a = std::move(it->second);
        g_xchg.erase(it);
    }
    DeviceScope ds(a->home_ordinal);
    CU(cudaDeviceSynchronize());
    if (a->ipc) CU(cudaIpcCloseMemHandle(a->base));
    else CU(cudaFree(a->base));
    a->status.release();
    return B200ZK_OK;
}

int32_t b200zk_msm_g1_xchg_dev(uint64_t bases, uint64_t offset, const void* d_scalars, uint64_t n, uint32_t scalar_fmt, uint64_t xchg,
                               void* d_out_mont, void* d_out_canon, void* stream) {
    TRY(need_init());
    if ((!d_scalars && n) || (!d_out_mont && !d_out_canon)) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    XchgArea* a = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_area_mu);
        auto it = g_xchg.find(xchg);
        if (it == g_xchg.end()) return fail(B200ZK_ERR_BAD_HANDLE, "unknown exchange handle");
        a = it->second.get();
    }
    // every part calls this the same number of times, so the sequence numbers agree without communication
    a->seq++;
    XchgArgs xa = a->args(0);
    Ctx* cp = nullptr;
    TRY(ctx_for_pointer(d_out_mont ? d_out_mont : d_out_canon, &cp));
    DeviceScope ds(cp->ordinal);
    TRY(a->status.ensure(16));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (a->seq == 1) CU(cudaMemsetAsync(a->status.p, 0, 16, s));
    if (n == 0) {
        LAUNCH(msm_combine_kernel, 1, 32, 0, s, (const uint32_t*)nullptr, 0u, 0u, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, xa);
    } else {
        TRY(msm_dev_common(bases, offset, d_scalars, n, 1, scalar_fmt, nullptr, nullptr, nullptr, stream, &xa));
    }
    // wait (bounded: ~4 s of SM clock) for the last arriver's result and bring it into local HBM
    LAUNCH(xchg_fetch_kernel, 1, 32, 0, s, (const unsigned long long*)xa.done, xa.seq, (const uint32_t*)xa.res_mont, (const uint32_t*)xa.res_canon,
           reinterpret_cast<uint32_t*>(d_out_mont), reinterpret_cast<uint32_t*>(d_out_canon), a->status.as<uint32_t>(), 8000000000ll);
    return B200ZK_OK;
}

int32_t b200zk_msm_g1_xchg(uint64_t bases, uint64_t offset, const uint8_t* scalars, uint64_t n, uint32_t scalar_fmt, uint64_t xchg,
                           uint8_t out_affine[96]) {
    Ctx* cp = nullptr;
    TRY(current_ctx(&cp));
    if ((!scalars && n) || !out_affine) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (scalar_fmt > B200ZK_FMT_MONT) return fail(B200ZK_ERR_INVALID_ARG, "unknown scalar format");
    XchgArea* a = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_area_mu);
        auto it = g_xchg.find(xchg);
        if (it == g_xchg.end()) return fail(B200ZK_ERR_BAD_HANDLE, "unknown exchange handle");
        a = it->second.get();
    }
    std::shared_ptr<TableSet> ts;
    TRY(get_table(bases, offset, n, &ts));
    const TableShard* sh = shard_on(*ts, cp->index, offset, n);
    if (!sh) return fail(B200ZK_ERR_INVALID_ARG, "msm: this slice of the table is not resident on the selected GPU");
    a->seq++;
    XchgArea view = *a;
    view.status = DevBuf();
    view.fetch = true;
    ColSrc src;
    src.base = scalars;
    src.stride = (size_t)n * 32;
    DeviceScope ds(cp->ordinal);
    SlotLease lease(*cp);
    // (the scalars stream up in two pieces under the first piece's accumulation, like the single-GPU call)
    return msm_host_device(*cp, *lease.s, &sh->t, offset - sh->start, src, 0, 1, 0, n, scalar_fmt, &view, out_affine);
}

int32_t b200zk_xchg_status(uint64_t xchg, uint32_t* out_timed_out) {
    TRY(need_init());
    if (!out_timed_out) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    XchgArea* a = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_area_mu);
        auto it = g_xchg.find(xchg);
        if (it == g_xchg.end()) return fail(B200ZK_ERR_BAD_HANDLE, "unknown exchange handle");
        a = it->second.get();
    }
    *out_timed_out = 0;
    if (!a->status.p) return B200ZK_OK;
    DeviceScope ds(a->home_ordinal);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out_timed_out, a->status.p, 4, cudaMemcpyDeviceToHost));
    return B200ZK_OK;
}

int32_t b200zk_ntt_fr_dev(void* d_data, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                          const uint8_t coset_shift[32], void* stream) {
    TRY(need_init());
    if (!d_data || !omega) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if ((uintptr_t)d_data & 15) return fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    Ctx* c = nullptr;
    TRY(ctx_for_pointer(d_data, &c));
    DeviceScope ds(c->ordinal);
    SlotLease lease(*c);
    return ntt_run(*c, *lease.s, reinterpret_cast<uint32_t*>(d_data), batch, log_n, omega, flags, coset_shift,
                   reinterpret_cast<cudaStream_t>(stream));
}

int32_t b200zk_ntt_fr_batch(uint8_t* data, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                            const uint8_t coset_shift[32]) {
    PolySrc src;
    src.base = data;
    return ntt_host(src, batch, log_n, omega, flags, coset_shift);
}

int32_t b200zk_ntt_fr_batch_ptrs(uint8_t* const* data, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                                 const uint8_t coset_shift[32]) {
    if (!data && batch) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    PolySrc src;
    src.ptrs = data;
    return ntt_host(src, batch, log_n, omega, flags, coset_shift);
}

int32_t b200zk_ntt_sharded_layout(uint32_t log_n, uint32_t* out_log_r, uint32_t* out_log_c) {
    TRY(need_init());
    if (!out_log_r || !out_log_c) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (log_n > 32) return fail(B200ZK_ERR_INVALID_ARG, "ntt: log_n exceeds the 2-adicity of Fr");
    ntt_shard_shape(log_n, (int)g_devs.size(), out_log_r, out_log_c);
    return B200ZK_OK;
}

int32_t b200zk_ntt_fr_sharded_dev(void* const* d_in, void* const* d_out, uint32_t n_parts, uint32_t log_n, const uint8_t omega[32],
                                  uint32_t flags, const uint8_t coset_shift[32]) {
    TRY(need_init());
    if (!d_in || !d_out || !omega) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (n_parts != g_devs.size()) return fail(B200ZK_ERR_INVALID_ARG, "ntt: one block per bound GPU");
    for (uint32_t k = 0; k < n_parts; k++)
        if (!d_in[k] || !d_out[k] || (((uintptr_t)d_in[k] | (uintptr_t)d_out[k]) & 15)) return fail(B200ZK_ERR_INVALID_ARG, "ntt: null or misaligned block");
    ShardIO io;
    io.d_in = d_in;
    io.d_out = d_out;
    return ntt_sharded(io, log_n, omega, flags, coset_shift);
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
    Ctx* cp = nullptr;
    TRY(ctx_for_pointer(d_out_mont, &cp));
    if (!d_out_mont && n) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    Ctx& c = *cp;
    DeviceScope ds(c.ordinal);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    uint32_t* d_gen = nullptr;
    TRY(get_generator_dev(c, &d_gen, s));
    {
        std::lock_guard<std::mutex> lk(c.mu);
        if (!c.fixed_table) {
            CU(cudaMalloc(&c.fixed_table, 8 * 256 * 96));
            LAUNCH(g1_fixed_table_kernel, 16, 128, 0, c.stream, (const uint32_t*)d_gen, c.fixed_table);
            CU(cudaStreamSynchronize(c.stream));
        }
    }
    if (n) LAUNCH(g1_synth_bases_kernel, (unsigned)((n + 127) / 128), 128, 0, s, (const uint32_t*)c.fixed_table, seed, start, n,
                  reinterpret_cast<uint32_t*>(d_out_mont));
    return B200ZK_OK;
}

uint64_t b200zk_launch_count(void) { return g_launches.load(); }

int32_t b200zk_set_profiling(uint32_t enable) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_profiling = enable != 0;
    g_prof_ctx = nullptr;
    g_prof_slot = nullptr;
    return B200ZK_OK;
}

int32_t b200zk_get_profile(uint32_t* kind, double* phase_ms, uint32_t cap, uint32_t* n_phases, uint32_t* msm_window_bits,
                           uint32_t* msm_windows) {
    TRY(need_init());
    if (!kind || !phase_ms || !n_phases) return fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    Ctx* c;
    Slot* sl;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        c = g_prof_ctx;
        sl = g_prof_slot;
    }
    *kind = 0;
    *n_phases = 0;
    if (msm_window_bits) *msm_window_bits = 0;
    if (msm_windows) *msm_windows = 0;
    if (!c || !sl) return B200ZK_OK;
    *kind = (uint32_t)sl->ev_kind;
    if (msm_window_bits) *msm_window_bits = sl->last_plan.c;
    if (msm_windows) *msm_windows = sl->last_plan.W;
    if (sl->ev_count < 2) return B200ZK_OK;
    DeviceScope ds(c->ordinal);
    CU(cudaEventSynchronize(sl->ev[sl->ev_count - 1]));
    for (int i = 0; i + 1 < sl->ev_count && (uint32_t)i < cap; i++) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, sl->ev[i], sl->ev[i + 1]));
        phase_ms[i] = ms;
        *n_phases = (uint32_t)i + 1;
    }
    return B200ZK_OK;
}

int32_t b200zk_set_msm_tuning(uint32_t window_bits, uint32_t smax) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_tune_c = window_bits & 0xffu;
    g_tune_variant = ((window_bits >> 8) & 0x7fu) ? ((window_bits >> 8) & 0x7fu) - 1 : 6;  // bits 8..14: 1 + accumulate-kernel variant
    g_tune_no_tables = (window_bits >> 15) & 1u;    // bit 15: do not build window tables at registration
    g_tune_smax = smax;
    return B200ZK_OK;
}

}  // extern "C"
