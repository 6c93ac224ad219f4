// ac_api.cu -- C ABI of libauriclass_cuda.so (include/auriclass_cuda.h) and the
// host-side orchestration of the kernels: input arena, segment pipeline on a
// copy stream + a compute stream, threshold speculation / verification.
//
// Boundary being replaced: the `mash sketch` / `mash dist` subprocess calls of
// auriclass/classes.py:92-104, :576-596, :696-713.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <set>
#include <vector>

#include "../../include/auriclass_cuda.h"
#include "ac_common.cuh"
#include "ac_xchg.cuh"

using namespace ac;

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(e_ == cudaErrorMemoryAllocation ? AC_E_NOMEM : AC_E_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                      \
    } while (0)

// AC_TRACE=1: host wall-clock of the phases of ac_sketch_end on stderr (tuning aid)
#include <chrono>
static bool trace_on() { static int v = -1; if (v < 0) { const char *e = getenv("AC_TRACE"); v = (e && *e && *e != '0') ? 1 : 0; } return v == 1; }
struct Trace {
    std::chrono::steady_clock::time_point t0;
    Trace() : t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!trace_on()) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[ac trace] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ------------------------------------------------------------------ types
struct Slab {
    uint8_t *ptr;
    size_t cap, used;
};

struct Piece {                 // contiguous raw text of one file, in stream order
    const uint8_t *ptr;
    size_t len;                // may shrink at EOF (trailing blank lines are dropped)
    int file_idx, fmt;
    bool first_of_file, last_of_file, append_nl;
    cudaEvent_t ready;         // H2D copy of the piece's latest bytes (nullptr for device feeds)
};

// Sketch buffers are recycled through their creating context: cudaMalloc / cudaFree of the last sub-allocation of a
// driver block maps / unmaps the block and synchronises the device (milliseconds per isolate otherwise).
struct SketchBuf { unsigned long long *h; uint32_t *c; uint64_t cap; };

struct ac_ctx {
    int dev;
    cudaStream_t s_copy, s_comp;
    cudaStream_t s_own;        // the compute stream this context created (s_comp unless the caller injected one)
    bool stream_injected;      // s_comp belongs to the caller (ac_set_stream): device-side outputs are stream ordered, not synced
    uint64_t launches;
    float last_ms[4];
    bool profile;
    size_t cand_cap;           // entries in the candidate / sort buffers
    size_t seg_bytes;          // raw text per segment (SEG_BYTES, or AC_SEG_MB for tests)
    size_t stream_seg_bytes;   // bytes that must be pending before a still-open file is processed (streamed feeds:
                               // smaller than seg_bytes, so little is left to do when the last chunk arrives)
    SelectBufs sb;
    bool sb_ok;
    CountTable ct;             // count table of speculative passes (key/cnt/qhist owned here; qlist aliases sb.keys_b)
    int ct_log2;               // log2(slots), AC_TABLE_LOG2
    uint8_t *d_Z[2];           // base stream of the segment in flight, ping-pong by segment parity
    unsigned long long *d_st1, *d_st2;  // K1 (FASTA) look-back status words (state, output offset), one per tile
    uint32_t seg_epoch;        // serial of the last segment launched on this context (tags the status words)
    size_t max_tiles;
    JobDev *d_jd, *h_jd;
    std::vector<Slab> slabs;
    size_t cur_slab;
    size_t sk_cap;             // entries in sb.sk_hash / sb.sk_cnt
    ac_job *active;
    // dist scratch
    double *d_dist, *d_p, *h_dist, *h_p;
    unsigned long long *d_common, *d_denom, *h_common, *h_denom;
    unsigned long long *d_qmeta, *h_qmeta;  // [3][max_q]: pointer, n, length
    void *d_dist_parts;        // per-slice counts of the split dist kernels (DIST_SPLIT_MAX_PAIRS * 64 bytes)
    size_t dist_cap, q_cap;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<SketchBuf> sk_pool;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof[3];
    cudaEvent_t ev_job0, ev_job1;
    cudaEvent_t ev_t0, ev_t1;  // ac_timer_start / ac_timer_stop
};

struct ac_job {
    ac_ctx *ctx;
    int k, s, m, mode, use64;
    uint32_t seed;
    std::vector<Piece> pieces;
    // file being fed
    int cur_file;
    bool file_open;
    int cur_fmt;               // -1 until the first non-blank byte is seen
    uint8_t tail[64];
    size_t tail_len;
    // pass state
    bool pass_started;
    unsigned long long tau0;
    bool safe;
    bool table;                       // the pass counts its candidates in the count table (speculative passes), not by sorting
    int passes;
    size_t proc_piece, proc_off;      // first unprocessed byte
    size_t pending;                   // fed but unprocessed bytes
    bool first_of_stream;
    double cand_bound;                // host-side bound on entries in the candidate buffer
    uint32_t n_table_host;
    unsigned long long tau_host;
    size_t bytes_done, next_mark;
    uint32_t seg_serial, ticket_base; // segments / K1 tiles launched in the current pass
    bool ended_local;
    bool table_pending;               // ended_local in count-table mode: the exchange table is cut straight from the count
                                      // table, range by range, once tau_cut is known (ac_job_table_ranges)
};

struct ac_sketch {
    ac_ctx *ctx;               // creator (may already be destroyed when the sketch is freed: use `dev`, not ctx->dev)
    int dev;
    unsigned long long *d_hash;
    uint32_t *d_cnt;
    uint64_t n;
    uint64_t cap;              // entries allocated (>= n): decides whether a pooled buffer pair can be reused
    int k, s, hash_bits;
    uint64_t length;
};

// A sketch may outlive its context, so before its buffers go back to the context's pool liveness is checked here.
static std::mutex g_ctx_mu;
static std::set<ac_ctx *> g_live_ctx;

struct ac_refs {
    ac_ctx *ctx;
    int dev;
    unsigned long long *d_hash, *d_off, *d_len;
    int n_refs, k, s, hash_bits;
    uint64_t max_n;
};

static const size_t SLAB_BYTES = 256u << 20;
static const size_t HOLD_BACK = 64;

// ------------------------------------------------------------------ library
extern "C" int ac_abi_version(void) { return AC_ABI_VERSION; }

extern "C" const char *ac_last_error(void) { return g_err; }

extern "C" int ac_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int ac_probe(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(AC_E_CUDA, "no CUDA device visible (%s)", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    }
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    if (p.major != 10) return fail(AC_E_CUDA, "device 0 is sm_%d%d; libauriclass_cuda.so is built for sm_100a only", p.major, p.minor);
    return AC_OK;
}

extern "C" void *ac_alloc_pinned(size_t n) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, n ? n : 1, cudaHostAllocDefault) != cudaSuccess) {
        fail(AC_E_NOMEM, "cudaHostAlloc(%zu) failed: %s", n, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
extern "C" void ac_free_pinned(void *p) { if (p) cudaFreeHost(p); }

static void free_select(ac_ctx *c) {
    SelectBufs &b = c->sb;
    cudaFree(b.keys_a); cudaFree(b.keys_b); cudaFree(b.vals_a); cudaFree(b.vals_b);
    cudaFree(b.tab_hash); cudaFree(b.tab_cnt); cudaFree(b.start_sum); cudaFree(b.hist); cudaFree(b.tile_tmp);
    cudaFree(b.sk_hash); cudaFree(b.sk_cnt);
    memset(&b, 0, sizeof(b));
    cudaFree(c->ct.key); cudaFree(c->ct.cnt); cudaFree(c->ct.qhist);
    memset(&c->ct, 0, sizeof(c->ct));
    c->sb_ok = false;
}

// device buffers for a sketch of n entries: a pooled pair when one is big enough, else fresh allocations
static cudaError_t sketch_buf_get(ac_ctx *c, uint64_t n, SketchBuf *out) {
    const uint64_t want = n ? n : 1;
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        for (size_t i = 0; i < c->sk_pool.size(); i++)
            if (c->sk_pool[i].cap >= want) { *out = c->sk_pool[i]; c->sk_pool.erase(c->sk_pool.begin() + i); return cudaSuccess; }
    }
    out->h = nullptr; out->c = nullptr; out->cap = want;
    cudaError_t e = cudaMalloc(&out->h, want * 8);
    if (e == cudaSuccess) e = cudaMalloc(&out->c, want * 4);
    if (e != cudaSuccess) { cudaFree(out->h); cudaFree(out->c); out->h = nullptr; out->c = nullptr; }
    return e;
}
static void sketch_buf_put(ac_ctx *c, int dev, const SketchBuf &b) {
    if (!b.h && !b.c) return;
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        if (c && g_live_ctx.count(c) && c->sk_pool.size() < 64) { c->sk_pool.push_back(b); return; }
    }
    cudaSetDevice(dev);
    cudaFree(b.h); cudaFree(b.c);
}

static int alloc_select(ac_ctx *c, size_t sk_cap) {
    SelectBufs &b = c->sb;
    if (c->sb_ok && sk_cap > c->sk_cap) {
        cudaFree(b.sk_hash); cudaFree(b.sk_cnt);
        b.sk_hash = nullptr; b.sk_cnt = nullptr; c->sk_cap = 0;
        CK(cudaMalloc(&b.sk_hash, sk_cap * 8)); CK(cudaMalloc(&b.sk_cnt, sk_cap * 4));
        c->sk_cap = sk_cap;
    }
    if (c->sb_ok) return AC_OK;
    memset(&b, 0, sizeof(b));
    const size_t cap = c->cand_cap;
    b.cap = cap; b.tab_cap = cap;
    CK(cudaMalloc(&b.keys_a, cap * 8)); CK(cudaMalloc(&b.keys_b, cap * 8));
    CK(cudaMalloc(&b.vals_a, cap * 4)); CK(cudaMalloc(&b.vals_b, cap * 4));
    CK(cudaMalloc(&b.tab_hash, (cap + 1) * 8)); CK(cudaMalloc(&b.tab_cnt, (cap + 1) * 4));
    CK(cudaMalloc(&b.start_sum, (cap + 2) * 4));
    const size_t rs_tiles = cap / 4096 + 2;
    CK(cudaMalloc(&b.hist, (256 * rs_tiles + 256) * 4));
    CK(cudaMalloc(&b.tile_tmp, (2 * (cap / 2048 + 2)) * 4));
    CK(cudaMalloc(&b.sk_hash, sk_cap * 8)); CK(cudaMalloc(&b.sk_cnt, sk_cap * 4));
    c->sk_cap = sk_cap;
    {
        const size_t slots = (size_t)1 << c->ct_log2;
        memset(&c->ct, 0, sizeof(c->ct));
        CK(cudaMalloc(&c->ct.key, slots * 8)); CK(cudaMalloc(&c->ct.cnt, slots * 4));
        CK(cudaMalloc(&c->ct.qhist, ((size_t)1 << CT_BUCKET_BITS) * 4));
        c->ct.mask = (uint32_t)(slots - 1);
        c->ct.qlist = b.keys_b;   // the sort scratch is idle while a table pass runs
        c->ct.qcap = (uint32_t)std::min<size_t>(cap, 0xffffffffu);
    }
    c->sb_ok = true;
    return AC_OK;
}

extern "C" void ac_sketch_abort(ac_job *j);

static int create_impl(ac_ctx *c, int device);

extern "C" int ac_create(int device, ac_ctx **out) {
    if (!out) return fail(AC_E_PARAM, "ac_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(AC_E_CUDA, "no CUDA device visible: libauriclass_cuda.so has no CPU path"); }
    if (device < 0 || device >= n) return fail(AC_E_PARAM, "device %d out of range (0..%d)", device, n - 1);
    CK(cudaSetDevice(device));
    ac_ctx *c = new ac_ctx();   // value-initialised: every handle starts out null, so a partial context can be destroyed
    c->dev = device;
    const int rc = create_impl(c, device);
    if (rc) {   // streams, events and buffers made so far go the way of a finished context (g_err keeps the cause)
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        ac_destroy(c);
        cudaGetLastError();
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    { std::lock_guard<std::mutex> lk(g_ctx_mu); g_live_ctx.insert(c); }
    *out = c;
    return AC_OK;
}

static int create_impl(ac_ctx *c, int device) {
    c->dev = device; c->launches = 0; c->profile = false; c->sb_ok = false; c->active = nullptr;
    c->cur_slab = 0; c->sk_cap = 0;
    memset(c->last_ms, 0, sizeof(c->last_ms));
    memset(&c->sb, 0, sizeof(c->sb));
    c->d_dist = c->d_p = c->h_dist = c->h_p = nullptr;
    c->d_common = c->d_denom = c->h_common = c->h_denom = nullptr;
    c->d_qmeta = c->h_qmeta = nullptr; c->dist_cap = c->q_cap = 0;
    const char *cap_env = getenv("AC_CAND_CAP");
    c->cand_cap = cap_env ? (size_t)strtoull(cap_env, nullptr, 10) : ((size_t)48 << 20);
    if (c->cand_cap < (1u << 16)) c->cand_cap = 1u << 16;
    const char *tab_env = getenv("AC_TABLE_LOG2");
    c->ct_log2 = tab_env ? atoi(tab_env) : 23;   // 8 Mi slots (96 MB): C. auris at 200x leaves it under 40 % full
    if (c->ct_log2 < 10) c->ct_log2 = 10;
    if (c->ct_log2 > 30) c->ct_log2 = 30;
    memset(&c->ct, 0, sizeof(c->ct));
    const char *seg_env = getenv("AC_SEG_MB");
    c->seg_bytes = seg_env ? ((size_t)strtoull(seg_env, nullptr, 10) << 20) : SEG_BYTES;
    if (c->seg_bytes < ((size_t)1 << 20)) c->seg_bytes = (size_t)1 << 20;
    if (c->seg_bytes > ((size_t)1 << 30)) c->seg_bytes = (size_t)1 << 30;
    c->stream_seg_bytes = std::min(c->seg_bytes, SEG_SAFE_BYTES);
    CK(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
    c->s_own = c->s_comp; c->stream_injected = false;
    for (int i = 0; i < 2; i++) CK(cudaMalloc(&c->d_Z[i], Z_CARRY + c->seg_bytes + K1_TILE + K2_TILE + 4096));
    c->max_tiles = c->seg_bytes / K1_TILE + 4;
    CK(cudaMalloc(&c->d_st1, c->max_tiles * 8)); CK(cudaMalloc(&c->d_st2, c->max_tiles * 8));
    CK(cudaMemset(c->d_st1, 0, c->max_tiles * 8)); CK(cudaMemset(c->d_st2, 0, c->max_tiles * 8));
    c->seg_epoch = 0;
    CK(cudaMalloc(&c->d_jd, sizeof(JobDev)));
    CK(cudaHostAlloc(&c->h_jd, sizeof(JobDev), cudaHostAllocDefault));
    CK(cudaEventCreate(&c->ev_job0)); CK(cudaEventCreate(&c->ev_job1));
    CK(cudaEventCreate(&c->ev_t0)); CK(cudaEventCreate(&c->ev_t1));
    return AC_OK;
}

extern "C" void ac_destroy(ac_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->dev);
    if (c->active) ac_sketch_abort(c->active);  // a job cannot outlive its context
    { std::lock_guard<std::mutex> lk(g_ctx_mu); g_live_ctx.erase(c); }
    cudaDeviceSynchronize();
    for (auto &b : c->sk_pool) { cudaFree(b.h); cudaFree(b.c); }
    c->sk_pool.clear();
    free_select(c);
    cudaFree(c->d_Z[0]); cudaFree(c->d_Z[1]); cudaFree(c->d_st1); cudaFree(c->d_st2); cudaFree(c->d_jd); cudaFreeHost(c->h_jd);
    for (auto &s : c->slabs) cudaFree(s.ptr);
    cudaFree(c->d_dist); cudaFree(c->d_p); cudaFree(c->d_common); cudaFree(c->d_denom); cudaFree(c->d_qmeta); cudaFree(c->d_dist_parts);
    cudaFreeHost(c->h_dist); cudaFreeHost(c->h_p); cudaFreeHost(c->h_common); cudaFreeHost(c->h_denom); cudaFreeHost(c->h_qmeta);
    for (auto ev : c->ev_pool) cudaEventDestroy(ev);
    for (int i = 0; i < 3; i++) for (auto &p : c->prof[i]) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (cudaEvent_t ev : {c->ev_job0, c->ev_job1, c->ev_t0, c->ev_t1}) if (ev) cudaEventDestroy(ev);
    for (cudaStream_t st : {c->s_copy, c->s_own}) if (st) cudaStreamDestroy(st);
    cudaGetLastError();
    delete c;
}

extern "C" int ac_sync(ac_ctx *c) {
    if (!c) return fail(AC_E_PARAM, "ac_sync: ctx is NULL");
    CK(cudaSetDevice(c->dev));
    CK(cudaStreamSynchronize(c->s_copy));
    CK(cudaStreamSynchronize(c->s_comp));
    return AC_OK;
}

extern "C" int ac_set_stream(ac_ctx *c, void *stream, int injected) {
    if (!c) return fail(AC_E_PARAM, "ac_set_stream: ctx is NULL");
    if (c->active) return fail(AC_E_STATE, "ac_set_stream while a sketch job is active");
    CK(cudaSetDevice(c->dev));
    CK(cudaStreamSynchronize(c->s_comp));
    if (injected) { c->s_comp = (cudaStream_t)stream; c->stream_injected = true; }
    else { c->s_comp = c->s_own; c->stream_injected = false; }
    return AC_OK;
}

extern "C" uint64_t ac_launch_count(const ac_ctx *c) { return c ? c->launches : 0; }

// CUDA-event stopwatch on the compute stream: brackets whatever the caller enqueues in between (a sketch job and the
// dist call that follows it, collectives a caller put on an injected stream included)
extern "C" int ac_timer_start(ac_ctx *c) {
    if (!c) return fail(AC_E_PARAM, "ac_timer_start: ctx is NULL");
    CK(cudaSetDevice(c->dev));
    CK(cudaEventRecord(c->ev_t0, c->s_comp));
    return AC_OK;
}
extern "C" int ac_timer_stop(ac_ctx *c, float *ms) {
    if (!c || !ms) return fail(AC_E_PARAM, "ac_timer_stop: NULL argument");
    CK(cudaSetDevice(c->dev));
    CK(cudaEventRecord(c->ev_t1, c->s_comp));
    CK(cudaEventSynchronize(c->ev_t1));
    CK(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return AC_OK;
}

extern "C" int ac_set_profile(ac_ctx *c, int on) {
    if (!c) return fail(AC_E_PARAM, "ctx is NULL");
    c->profile = on != 0;
    return AC_OK;
}

extern "C" int ac_last_job_times(const ac_ctx *c, float ms[4]) {
    if (!c || !ms) return fail(AC_E_PARAM, "ac_last_job_times: NULL argument");
    memcpy(ms, c->last_ms, sizeof(c->last_ms));
    return AC_OK;
}

extern "C" int ac_int_peak(ac_ctx *c, double *instr_per_s) {
    if (!c || !instr_per_s) return fail(AC_E_PARAM, "ac_int_peak: NULL argument");
    CK(cudaSetDevice(c->dev));
    LaunchCtx lc{c->s_comp, &c->launches};
    *instr_per_s = run_int_peak(lc);
    CK(cudaGetLastError());
    return AC_OK;
}

extern "C" int ac_pipe_rates(ac_ctx *c, double rates[3]) {
    if (!c || !rates) return fail(AC_E_PARAM, "ac_pipe_rates: NULL argument");
    CK(cudaSetDevice(c->dev));
    LaunchCtx lc{c->s_comp, &c->launches};
    run_pipe_rates(lc, rates);
    CK(cudaGetLastError());
    return AC_OK;
}

extern "C" int ac_alu_rates(ac_ctx *c, double rates[5]) {
    if (!c || !rates) return fail(AC_E_PARAM, "ac_alu_rates: NULL argument");
    CK(cudaSetDevice(c->dev));
    LaunchCtx lc{c->s_comp, &c->launches};
    run_alu_rates(lc, rates);
    CK(cudaGetLastError());
    return AC_OK;
}

// ------------------------------------------------------------------ profiling helpers
struct ProfScope {
    ac_ctx *c; int slot; cudaEvent_t a, b; bool on;
    ProfScope(ac_ctx *c_, int slot_) : c(c_), slot(slot_), a(nullptr), b(nullptr), on(c_->profile) {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, c->s_comp); }
    }
    ~ProfScope() { if (on) { cudaEventRecord(b, c->s_comp); c->prof[slot].push_back({a, b}); } }
};

static void prof_collect(ac_ctx *c) {
    for (int i = 0; i < 3; i++) {
        float tot = 0;
        for (auto &p : c->prof[i]) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) tot += ms;
            cudaEventDestroy(p.first); cudaEventDestroy(p.second);
        }
        c->prof[i].clear();
        c->last_ms[i] = tot;
    }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev_job0, c->ev_job1) == cudaSuccess) c->last_ms[3] = ms;
    cudaGetLastError();
}

// ------------------------------------------------------------------ sketch job
static bool is_blank(uint8_t ch) { return ch == '\n' || ch == '\r' || ch == ' ' || ch == '\t'; }

extern "C" int ac_sketch_begin(ac_ctx *c, int k, int sketch_size, uint32_t seed, int min_copies, int mode, ac_job **out) {
    if (!c || !out) return fail(AC_E_PARAM, "ac_sketch_begin: NULL argument");
    *out = nullptr;
    if (k < 1 || k > 32) return fail(AC_E_PARAM, "k = %d outside 1..32", k);
    if (sketch_size < 1 || sketch_size > (1 << 24)) return fail(AC_E_PARAM, "sketch size %d outside 1..2^24", sketch_size);
    if (min_copies < 1) return fail(AC_E_PARAM, "min_copies = %d < 1", min_copies);
    if (mode != AC_MODE_READS && mode != AC_MODE_FASTA) return fail(AC_E_PARAM, "unknown mode %d", mode);
    if (c->active) return fail(AC_E_STATE, "context already has an active sketch job");
    CK(cudaSetDevice(c->dev));
    cudaGetLastError();  // a stale error of an unrelated earlier call must not be blamed on this job's launches
    {
        int rc = alloc_select(c, (size_t)sketch_size);
        if (rc) { free_select(c); return rc; }
    }
    ac_job *j = new ac_job();
    j->ctx = c; j->k = k; j->s = sketch_size; j->m = (mode == AC_MODE_FASTA) ? 1 : min_copies; j->mode = mode;
    j->use64 = k > 16 ? 1 : 0; j->seed = seed;
    j->cur_file = -1; j->file_open = false; j->cur_fmt = -1; j->tail_len = 0;
    j->pass_started = false; j->tau0 = 0; j->safe = false; j->table = false; j->passes = 0;
    j->proc_piece = 0; j->proc_off = 0; j->pending = 0; j->first_of_stream = true;
    j->cand_bound = 0; j->n_table_host = 0; j->tau_host = 0; j->bytes_done = 0; j->next_mark = 0; j->ended_local = false; j->table_pending = false;
    for (auto &s : c->slabs) s.used = 0;
    c->cur_slab = 0;
    c->active = j;
    *out = j;
    return AC_OK;
}

static void job_free(ac_job *j) {
    if (!j) return;
    ac_ctx *c = j->ctx;
    cudaStreamSynchronize(c->s_copy);
    cudaStreamSynchronize(c->s_comp);
    for (auto &p : j->pieces) if (p.ready) c->ev_pool.push_back(p.ready);
    if (c->active == j) c->active = nullptr;
    delete j;
}

extern "C" void ac_sketch_abort(ac_job *j) { if (j) { cudaSetDevice(j->ctx->dev); job_free(j); } }

static unsigned long long space_max(const ac_job *j) { return j->use64 ? ~0ULL : 0xffffffffULL; }

static int read_jd(ac_ctx *c) {
    CK(cudaMemcpyAsync(c->h_jd, c->d_jd, sizeof(JobDev), cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaStreamSynchronize(c->s_comp));
    return AC_OK;
}

static int key_bits_for(unsigned long long tau) {
    int b = 64;
    while (b > 8 && !(tau >> (b - 1))) b--;
    return b;
}

// fold the candidate buffer into the table; one host sync to learn the sizes
static int consolidate(ac_job *j) {
    ac_ctx *c = j->ctx;
    int rc = read_jd(c);
    if (rc) return rc;
    const JobDev &h = *c->h_jd;
    if (h.overflow) return AC_OK;  // the pass is void; the caller re-runs it in safe mode
    const unsigned long long n_cand = h.cand_count;
    const uint32_t n_table = h.n_table;
    if (n_cand + n_table > c->sb.cap) return fail(AC_E_NOMEM, "candidate table needs %llu entries, capacity %zu (raise AC_CAND_CAP)", n_cand + n_table, c->sb.cap);
    LaunchCtx lc{c->s_comp, &c->launches};
    {
        ProfScope ps(c, 2);
        launch_consolidate(lc, c->sb, c->d_jd, n_cand, n_table, key_bits_for(h.tau), j->s, j->m);
    }
    CK(cudaGetLastError());
    j->cand_bound = 0;
    j->tau_host = h.tau;              // stale (pre-consolidation) value: an upper bound of the new tau
    j->n_table_host = (uint32_t)std::min<unsigned long long>(n_cand + n_table, 0xffffffffULL);  // upper bound
    return AC_OK;
}

static int start_pass(ac_job *j, unsigned long long tau0, bool safe) {
    ac_ctx *c = j->ctx;
    JobDev z;
    memset(&z, 0, sizeof(z));
    z.ps[0].at_line_start = 1;
    z.ps[0].n_zbuf = Z_CARRY;
    z.ps[0].z_begin = Z_CARRY;
    z.tau = tau0;
    j->seg_serial = 0; j->ticket_base = 0;
    *c->h_jd = z;
    CK(cudaMemcpyAsync(c->d_jd, c->h_jd, sizeof(JobDev), cudaMemcpyHostToDevice, c->s_comp));
    CK(cudaStreamSynchronize(c->s_comp));
    if (j->passes == 0) CK(cudaEventRecord(c->ev_job0, c->s_comp));   // device time of a job starts before its table is cleared
    j->table = !safe && tau0 != space_max(j);
    j->table_pending = false;
    if (j->table) {
        const size_t slots = (size_t)c->ct.mask + 1;
        CK(cudaMemsetAsync(c->ct.key, 0xff, slots * 8, c->s_comp));
        CK(cudaMemsetAsync(c->ct.cnt, 0, slots * 4, c->s_comp));
        CK(cudaMemsetAsync(c->ct.qhist, 0, ((size_t)1 << CT_BUCKET_BITS) * 4, c->s_comp));
        c->ct.min_copies = (uint32_t)j->m;
        c->ct.bshift = std::max(0, key_bits_for(tau0) - CT_BUCKET_BITS);
    }
    j->pass_started = true; j->tau0 = tau0; j->safe = safe; j->passes++;
    j->proc_piece = 0; j->proc_off = 0; j->first_of_stream = true;
    j->cand_bound = 0; j->n_table_host = 0; j->tau_host = tau0; j->bytes_done = 0; j->next_mark = c->seg_bytes;
    return AC_OK;
}

// descriptor of the segment [off, off+len) of piece pi, the next one of the pass (draws its epoch tag)
static int build_segdesc(ac_job *j, size_t pi, size_t off, size_t len, bool last_of_file, SegDesc *out) {
    ac_ctx *c = j->ctx;
    const Piece &p = j->pieces[pi];
    const bool append_nl = last_of_file && p.append_nl;
    const size_t seg_len = len + (append_nl ? 1 : 0);
    SegDesc sd;
    const uintptr_t addr = (uintptr_t)(p.ptr + off);
    sd.base = (const uint8_t *)(addr & ~(uintptr_t)15);
    sd.skip = (uint32_t)(addr & 15);
    sd.len = (uint32_t)seg_len;
    sd.n_tiles = seg_len ? div_up(sd.skip + seg_len, K1_TILE) : 1;  // an empty segment still threads the parser state
    sd.append_nl = append_nl ? 1 : 0;
    sd.fmt = p.fmt;
    sd.first_of_stream = j->first_of_stream ? 1 : 0;
    sd.last_of_file = last_of_file ? 1 : 0;
    sd.serial = j->seg_serial;
    if (++c->seg_epoch >= (1u << 30)) {  // status-word tags wrapped: start over with clean arrays
        CK(cudaMemsetAsync(c->d_st1, 0, c->max_tiles * 8, c->s_comp)); CK(cudaMemsetAsync(c->d_st2, 0, c->max_tiles * 8, c->s_comp));
        c->seg_epoch = 1;
    }
    sd.epoch = c->seg_epoch;
    sd.ticket_base = j->ticket_base;
    if (sd.n_tiles > c->max_tiles) return fail(AC_E_INTERNAL, "segment of %zu bytes exceeds the tile arrays", seg_len);
    *out = sd;
    return AC_OK;
}

// run K1 + K2 over bytes [off, off+len) of piece pi (len <= SEG_BYTES)
static int run_segment(ac_job *j, size_t pi, size_t off, size_t len, bool last_of_file) {
    ac_ctx *c = j->ctx;
    const Piece &p = j->pieces[pi];
    LaunchCtx lc{c->s_comp, &c->launches};
    const bool append_nl = last_of_file && p.append_nl;
    const size_t seg_len = len + (append_nl ? 1 : 0);
    // candidate-buffer headroom: consolidate first when this segment could overflow it
    const double frac = j->safe ? 1.0 : std::min(1.0, ((double)j->tau_host + 1.0) / ((double)space_max(j) + 1.0) * 2.0 + 1e-3);
    const double seg_bound = j->safe ? (double)seg_len : (double)seg_len * frac + 65536.0;
    if (!j->table && j->cand_bound > 0 && j->cand_bound + seg_bound + j->n_table_host > (double)c->sb.cap) {
        int rc = consolidate(j);
        if (rc) return rc;
    }
    SegDesc sd;
    if (p.ready) CK(cudaStreamWaitEvent(c->s_comp, p.ready, 0));
    {
        int rc = build_segdesc(j, pi, off, len, last_of_file, &sd);
        if (rc) return rc;
    }
    uint8_t *Z = c->d_Z[sd.serial & 1u];
    {
        ProfScope ps(c, 0);
        launch_parse_segment(lc, sd, c->d_st1, c->d_st2, c->d_jd, c->d_Z[(sd.serial + 1u) & 1u], Z);
    }
    {
        ProfScope ps(c, 1);
        CountTable off;
        memset(&off, 0, sizeof(off));
        const uint32_t zmax = (uint32_t)(Z_CARRY + seg_len + 1);
        const CountTable &ctu = j->table ? c->ct : off;
        // The first segment of a count-table pass starts at the speculative threshold (1/32 of the hash space: ~8x the
        // candidates of the later segments).  Its first third (text / 7 bytes of the stream: 4-line FASTQ is ~46 % bases)
        // is hashed on its own and the threshold pulled down from what that slice qualified -- about 3x coverage of a
        // 50x isolate, enough for s qualified hashes well below the speculative threshold -- before the rest runs.
        const uint32_t head = (j->table && j->seg_serial == 0 && seg_len >= ((size_t)32 << 20)) ? hash_ctas_for((uint32_t)(seg_len / 7)) : 0u;
        if (head) {
            if (!launch_hash_segment(lc, j->k, Z, zmax, (int)((sd.serial + 1u) & 1u), c->d_jd, c->sb.keys_a, c->sb.cap, j->seed, j->use64, ctu, 0u, head))
                return fail(AC_E_PARAM, "this build of libauriclass_cuda.so has no hash kernel for k = %d", j->k);
            launch_table_tau_update(lc, c->ct, c->d_jd, j->s);
        }
        if (!launch_hash_segment(lc, j->k, Z, zmax, (int)((sd.serial + 1u) & 1u), c->d_jd, c->sb.keys_a, c->sb.cap, j->seed, j->use64, ctu, head))
            return fail(AC_E_PARAM, "this build of libauriclass_cuda.so has no hash kernel for k = %d", j->k);
    }
    j->seg_serial++;
    if (sd.fmt == FMT_FASTA) j->ticket_base += sd.n_tiles;  // only k1_fasta draws tile tickets
    CK(cudaGetLastError());
    j->first_of_stream = false;
    j->cand_bound += seg_bound;
    j->bytes_done += seg_len;
    if (j->table) {
        // the threshold follows the histogram of qualified keys: a 5 us kernel, no host round trip.  Early
        // segments pull tau from the speculative 1/32 down to ~1.3x its final value; later ones barely move it.
        if (j->seg_serial <= 4 || (j->seg_serial & 3u) == 0) {
            ProfScope ps(c, 2);
            launch_table_tau_update(lc, c->ct, c->d_jd, j->s);
            CK(cudaGetLastError());
        }
        return AC_OK;
    }
    if (j->safe || j->bytes_done >= j->next_mark) {
        int rc = consolidate(j);
        if (rc) return rc;
        // one early consolidation pulls tau from the speculative 1/32 down to ~1.3x its final value; after that
        // the candidate stream is a trickle (~0.5 % of the k-mers) that is folded in when the buffer fills
        // (cand_bound guard above) or at the end of the job
        j->next_mark = (size_t)-1;
    }
    return AC_OK;
}

// Is there a segment to process now?  `all` = flush everything (job end), else only full segments once enough has
// been fed, keeping HOLD_BACK bytes of a still-open file (EOF may trim trailing blank lines).
static bool segment_available(const ac_job *j, bool all, size_t *usable_out, bool *open_out) {
    if (j->proc_piece >= j->pieces.size()) return false;
    const Piece &p = j->pieces[j->proc_piece];
    const bool is_last_piece = j->proc_piece + 1 == j->pieces.size();
    const bool open = is_last_piece && !p.last_of_file;
    const size_t avail = p.len - j->proc_off;
    const size_t usable = open ? (avail > HOLD_BACK ? avail - HOLD_BACK : 0) : avail;
    if (open && usable == 0) return false;
    if (!all) {
        if (j->pending < j->ctx->stream_seg_bytes + HOLD_BACK) return false;
        if (open && usable < j->ctx->stream_seg_bytes) return false;
    }
    *usable_out = usable; *open_out = open;
    return true;
}
// its extent (the pass must have started: the segment size depends on the pass kind)
static void segment_extent(const ac_job *j, size_t usable, size_t *len, bool *seg_last) {
    const Piece &p = j->pieces[j->proc_piece];
    *len = std::min(usable, j->safe ? std::min(SEG_SAFE_BYTES, j->ctx->seg_bytes) : j->ctx->seg_bytes);
    *seg_last = p.last_of_file && (j->proc_off + *len == p.len);
}

// process fed bytes, segment by segment
static int pump(ac_job *j, bool all) {
    size_t usable;
    bool open;
    while (segment_available(j, all, &usable, &open)) {
        if (!j->pass_started) {
            // the size fed so far decides the speculative threshold: big inputs start at 1/32 of
            // the hash space (64 expected candidates per 2,048-window sub-tile of K2, half its staging slots),
            // small ones keep everything (tau = max, nothing to verify)
            // (tau0 = 1/32 must leave >= s qualified hashes: ~128 bytes of text per sketch entry at the least)
            const size_t spec_min = std::max((size_t)4 << 20, (size_t)128 * (size_t)j->s);
            const bool big = j->pending >= std::min(spec_min, j->ctx->seg_bytes);
            int rc = start_pass(j, big ? (space_max(j) >> 5) : space_max(j), !big);
            if (rc) return rc;
        }
        size_t len;
        bool seg_last;
        segment_extent(j, usable, &len, &seg_last);
        const size_t piece_len = j->pieces[j->proc_piece].len;
        int rc = run_segment(j, j->proc_piece, j->proc_off, len, seg_last);
        if (rc) return rc;
        j->pending -= std::min(j->pending, len);
        j->proc_off += len;
        if (j->proc_off == piece_len && !open) { j->proc_piece++; j->proc_off = 0; }
    }
    return AC_OK;
}

// close the file being fed: trailing blank lines are dropped, a missing final newline is added virtually
static void close_file(ac_job *j) {
    if (!j->file_open) return;
    j->file_open = false;
    // last piece of this file
    Piece *last = nullptr;
    for (auto it = j->pieces.rbegin(); it != j->pieces.rend(); ++it)
        if (it->file_idx == j->cur_file) { last = &*it; break; }
    if (!last) return;
    last->last_of_file = true;
    size_t t = j->tail_len, blank = 0;
    while (blank < t && is_blank(j->tail[t - 1 - blank])) blank++;
    // inside the trailing blank run keep everything up to its first '\n'
    size_t first_nl = (size_t)-1;
    for (size_t i = t - blank; i < t; i++) if (j->tail[i] == '\n') { first_nl = i; break; }
    if (first_nl == (size_t)-1) last->append_nl = true;
    else {
        size_t drop = t - 1 - first_nl;
        drop = std::min(drop, last->len);
        last->len -= drop;
        j->pending -= std::min(j->pending, drop);
        last->append_nl = false;
    }
}

static void push_tail(ac_job *j, const uint8_t *p, size_t n) {
    if (n >= sizeof(j->tail)) { memcpy(j->tail, p + n - sizeof(j->tail), sizeof(j->tail)); j->tail_len = sizeof(j->tail); return; }
    const size_t keep = std::min(j->tail_len, sizeof(j->tail) - n);
    memmove(j->tail, j->tail + j->tail_len - keep, keep);
    memcpy(j->tail + keep, p, n);
    j->tail_len = keep + n;
}

static int feed_common(ac_job *j, const uint8_t *src, size_t n, int file_idx, int eof, bool on_device) {
    if (!j) return fail(AC_E_PARAM, "ac_sketch_feed: job is NULL");
    ac_ctx *c = j->ctx;
    if (j->ended_local) return fail(AC_E_STATE, "feed after end");
    if (n && !src) return fail(AC_E_PARAM, "ac_sketch_feed: NULL buffer with n = %zu", n);
    CK(cudaSetDevice(c->dev));
    if (!j->file_open || file_idx != j->cur_file) {
        if (j->file_open) close_file(j);
        if (j->mode == AC_MODE_FASTA && !j->pieces.empty())
            return fail(AC_E_STATE, "AC_MODE_FASTA sketches one file per job (mash sketch without -r)");
        j->cur_file = file_idx; j->file_open = true; j->cur_fmt = -1; j->tail_len = 0;
    }
    // host view of the head (format sniffing) and tail (EOF trimming) of the chunk
    uint8_t headbuf[64], tailbuf[64];
    const uint8_t *hsrc = src;
    size_t skip = 0;
    if (n) {
        if (j->cur_fmt < 0) {
            // skip leading blanks, sniff '>' / '@'
            if (on_device) {
                size_t pos = 0;
                while (pos < n && j->cur_fmt < 0) {
                    const size_t m = std::min(sizeof(headbuf), n - pos);
                    CK(cudaMemcpy(headbuf, src + pos, m, cudaMemcpyDeviceToHost));
                    size_t i = 0;
                    while (i < m && is_blank(headbuf[i])) i++;
                    pos += i;
                    if (i < m) { j->cur_fmt = headbuf[i] == '>' ? FMT_FASTA : (headbuf[i] == '@' ? FMT_FASTQ : -2); break; }
                }
                skip = pos;
            } else {
                while (skip < n && is_blank(hsrc[skip])) skip++;
                if (skip < n) j->cur_fmt = hsrc[skip] == '>' ? FMT_FASTA : (hsrc[skip] == '@' ? FMT_FASTQ : -2);
            }
        }
        if (j->cur_fmt == -2) {
            // neither FASTA nor FASTQ: kseq finds no record start in such a file;
            // we do not scan it for a later '>' / '@' -- report it
            return fail(AC_E_FORMAT, "file %d starts with neither '>' nor '@'", file_idx);
        }
        const size_t m = n - skip;
        if (m && j->cur_fmt >= 0) {
            if (on_device) {
                const size_t tl = std::min(sizeof(tailbuf), m);
                CK(cudaMemcpy(tailbuf, src + n - tl, tl, cudaMemcpyDeviceToHost));
                push_tail(j, tailbuf, tl);
                Piece p;
                p.ptr = src + skip; p.len = m; p.file_idx = file_idx; p.fmt = j->cur_fmt;
                p.first_of_file = true; p.last_of_file = false; p.append_nl = false; p.ready = nullptr;
                for (auto &q : j->pieces) if (q.file_idx == file_idx) p.first_of_file = false;
                j->pieces.push_back(p);
            } else {
                push_tail(j, hsrc + skip, m);
                // place the bytes in the arena
                Slab *sl = nullptr;
                if (c->cur_slab < c->slabs.size() && c->slabs[c->cur_slab].cap - c->slabs[c->cur_slab].used >= m + 64)
                    sl = &c->slabs[c->cur_slab];
                else {
                    // never reuse a slab that holds earlier bytes of this job: they are retained for re-runs
                    for (size_t i = c->cur_slab + (c->slabs.empty() ? 0 : 1); i < c->slabs.size(); i++)
                        if (c->slabs[i].used == 0 && c->slabs[i].cap >= m + 64) { sl = &c->slabs[i]; c->cur_slab = i; break; }
                    if (!sl && c->cur_slab < c->slabs.size() && c->slabs[c->cur_slab].used == 0 &&
                        c->slabs[c->cur_slab].cap >= m + 64)
                        sl = &c->slabs[c->cur_slab];
                }
                if (!sl) {
                    Slab ns;
                    ns.cap = std::max(SLAB_BYTES, m + 4096); ns.used = 0; ns.ptr = nullptr;
                    CK(cudaMalloc(&ns.ptr, ns.cap));
                    c->slabs.push_back(ns);
                    c->cur_slab = c->slabs.size() - 1;
                    sl = &c->slabs.back();
                }
                uint8_t *dst = sl->ptr + sl->used;
                CK(cudaMemcpyAsync(dst, hsrc + skip, m, cudaMemcpyHostToDevice, c->s_copy));
                sl->used += m;
                cudaEvent_t ev;
                if (!c->ev_pool.empty()) { ev = c->ev_pool.back(); c->ev_pool.pop_back(); }
                else CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                CK(cudaEventRecord(ev, c->s_copy));
                Piece *prev = j->pieces.empty() ? nullptr : &j->pieces.back();
                if (prev && prev->file_idx == file_idx && !prev->last_of_file && prev->ptr + prev->len == dst &&
                    j->proc_piece < j->pieces.size()) {
                    prev->len += m;
                    if (prev->ready) c->ev_pool.push_back(prev->ready);
                    prev->ready = ev;
                } else {
                    Piece p;
                    p.ptr = dst; p.len = m; p.file_idx = file_idx; p.fmt = j->cur_fmt;
                    p.first_of_file = true; p.last_of_file = false; p.append_nl = false; p.ready = ev;
                    for (auto &q : j->pieces) if (q.file_idx == file_idx) p.first_of_file = false;
                    j->pieces.push_back(p);
                }
            }
            j->pending += m;
        }
    }
    if (eof) close_file(j);
    return pump(j, false);
}

extern "C" int ac_sketch_feed(ac_job *j, const uint8_t *host, size_t n, int file_idx, int eof) {
    return feed_common(j, host, n, file_idx, eof, false);
}
extern "C" int ac_sketch_feed_device(ac_job *j, const uint8_t *dev, size_t n, int file_idx, int eof) {
    return feed_common(j, dev, n, file_idx, eof, true);
}
extern "C" int ac_sketch_feed_wait(ac_job *j) {
    if (!j) return fail(AC_E_PARAM, "job is NULL");
    CK(cudaSetDevice(j->ctx->dev));
    CK(cudaStreamSynchronize(j->ctx->s_copy));
    return AC_OK;
}

static void fill_info(const ac_job *j, const JobDev &h, ac_sketch_info *info) {
    if (!info) return;
    memset(info, 0, sizeof(*info));
    info->n_hashes = h.n_sketch;
    info->n_records = h.n_records;
    info->n_kept = h.n_kept;
    info->total_len = h.n_bases - h.n_short_len;
    info->n_kmers = h.n_kmers;
    info->n_candidates = h.cand_total;
    info->max_hash = h.max_hash;
    info->sum_counts = h.sum_counts;
    info->final_tau = h.tau;
    info->genome_size = h.genome_size;
    info->coverage = h.coverage;
    info->hash_bits = j->use64 ? 64 : 32;
    info->passes = j->passes;
    info->n_bases = h.n_bases;
}

// end of a count-table pass: final threshold, then the qualified keys <= tau (for_exchange: every key <= tau)
// are collected, sorted and laid out as table + sketch prefix exactly as a consolidation leaves them
static int finish_table(ac_job *j, bool for_exchange) {
    ac_ctx *c = j->ctx;
    LaunchCtx lc{c->s_comp, &c->launches};
    ProfScope ps(c, 2);
    launch_table_tau_update(lc, c->ct, c->d_jd, j->s);
    CK(cudaGetLastError());
    int rc = read_jd(c);
    if (rc) return rc;
    const JobDev &h = *c->h_jd;
    if (h.overflow) return AC_OK;                                   // table or list full: the caller repeats the pass by sorting
    if (h.n_qualified < (uint32_t)j->s) return AC_OK;               // speculation failed: the caller repeats it with tau = max
    const uint32_t n_q = h.n_qualified;
    if (for_exchange) {
        // nothing to materialise yet: every rank's tau_r goes into the all-reduce first, then ac_job_table_ranges
        // writes the entries <= tau_cut range by range straight from the count table (no sort on this side)
        j->table_pending = true;
        return AC_OK;
    }
    launch_table_collect(lc, c->ct, c->sb, c->d_jd, h.n_qlist);
    launch_table_finish(lc, c->sb, c->d_jd, n_q, n_q, key_bits_for(h.tau), j->s, 0);
    CK(cudaGetLastError());
    return AC_OK;
}

// finish the local pass (all ranks): returns with the final table on the device and h_jd current
static int finish_local(ac_job *j, bool for_exchange = false) {
    ac_ctx *c = j->ctx;
    if (j->file_open) close_file(j);
    Trace tr;
    for (int attempt = 0; attempt < 4; attempt++) {
        int rc = pump(j, true);
        if (rc) return rc;
        tr.mark("  pump(all)");
        if (!j->pass_started) {  // nothing was fed at all
            rc = start_pass(j, space_max(j), true);
            if (rc) return rc;
        }
        rc = j->table ? finish_table(j, for_exchange) : consolidate(j);
        if (rc) return rc;
        if (!j->table_pending) {       // (a pending exchange table: finish_table's read is still current, nothing ran since)
            rc = read_jd(c);
            if (rc) return rc;
        }
        tr.mark("  table/consolidate");
        const JobDev &h = *c->h_jd;
        uint32_t ferr = h.format_error;
        if (h.fq_seq_bytes != h.fq_qual_bytes) ferr |= FERR_FQ_TRUNC;  // sequence and quality byte totals of the job differ
        if (ferr) {
            const char *what = (ferr & FERR_FQ_HEADER) ? "a FASTQ record does not start with '@' (multi-line FASTQ is not supported)"
                             : (ferr & FERR_FQ_PLUS) ? "third line of a FASTQ record does not start with '+' (multi-line FASTQ is not supported)"
                             : (ferr & FERR_FQ_TRUNC) ? "FASTQ input is truncated or sequence/quality lengths differ"
                             : "'+' line inside FASTA input";
            return fail(AC_E_FORMAT, "input rejected by the GPU parser: %s", what);
        }
        const bool spec_failed = !h.overflow && j->tau0 != space_max(j) && h.n_qualified < (uint32_t)j->s;
        if (!h.overflow && !spec_failed) return AC_OK;
        if (h.overflow && j->safe) return fail(AC_E_NOMEM, "candidate buffer overflow in safe mode (raise AC_CAND_CAP, now %zu)", c->sb.cap);
        // repeat the pass over the retained input: same threshold in safe mode after an overflow,
        // full hash space after a failed speculation
        rc = start_pass(j, spec_failed ? space_max(j) : j->tau0, true);
        if (rc) return rc;
        j->pending = 0;
        for (auto &p : j->pieces) j->pending += p.len;
    }
    return fail(AC_E_INTERNAL, "threshold search did not converge");
}

static int make_sketch(ac_ctx *c, const unsigned long long *d_hash, const uint32_t *d_cnt, uint64_t n, int k, int s,
                       int hash_bits, uint64_t length, ac_sketch **out) {
    ac_sketch *sk = new ac_sketch();
    sk->ctx = c; sk->dev = c->dev; sk->n = n; sk->k = k; sk->s = s; sk->hash_bits = hash_bits; sk->length = length;
    SketchBuf b;
    cudaError_t e = sketch_buf_get(c, n, &b);
    sk->d_hash = b.h; sk->d_cnt = b.c; sk->cap = b.cap;
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(sk->d_hash, d_hash, n * 8, cudaMemcpyDeviceToDevice, c->s_comp);
    if (e == cudaSuccess && n) {
        if (d_cnt) e = cudaMemcpyAsync(sk->d_cnt, d_cnt, n * 4, cudaMemcpyDeviceToDevice, c->s_comp);
        else e = cudaMemsetAsync(sk->d_cnt, 0, n * 4, c->s_comp);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_comp);
    if (e != cudaSuccess) {
        sketch_buf_put(c, c->dev, b); delete sk;
        return fail(AC_E_CUDA, "sketch allocation/copy failed: %s", cudaGetErrorString(e));
    }
    *out = sk;
    return AC_OK;
}

extern "C" int ac_sketch_end(ac_job *j, ac_sketch **out, ac_sketch_info *info) {
    if (!j || !out) { if (j) job_free(j); return fail(AC_E_PARAM, "ac_sketch_end: NULL argument"); }
    *out = nullptr;
    ac_ctx *c = j->ctx;
    cudaSetDevice(c->dev);
    Trace tr;
    int rc = finish_local(j);
    if (rc) { job_free(j); return rc; }
    tr.mark("finish_local");
    LaunchCtx lc{c->s_comp, &c->launches};
    {
        ProfScope ps(c, 2);
        launch_finalize(lc, c->sb, c->d_jd, j->use64);
    }
    cudaEventRecord(c->ev_job1, c->s_comp);
    rc = read_jd(c);
    if (rc) { job_free(j); return rc; }
    tr.mark("finalize+read");
    prof_collect(c);
    const JobDev h = *c->h_jd;
    fill_info(j, h, info);
    if (h.n_sketch == 0) {
        job_free(j);
        // Mash: "ERROR: Did not find fasta records in ..." (no record >= k, or no hash reached -m)
        return fail(AC_E_NO_RECORDS, "did not find sequence records (records: %llu, of length >= k: %llu, k-mers: %llu)",
                    h.n_records, h.n_kept, h.n_kmers);
    }
    const uint64_t length = (j->mode == AC_MODE_FASTA) ? (h.n_bases - h.n_short_len) : (uint64_t)h.genome_size;
    rc = make_sketch(c, c->sb.sk_hash, c->sb.sk_cnt, h.n_sketch, j->k, j->s, j->use64 ? 64 : 32, length, out);
    tr.mark("make_sketch");
    job_free(j);
    tr.mark("job_free");
    return rc;
}

// ------------------------------------------------------------------ multi-GPU pieces
extern "C" int ac_sketch_end_local(ac_job *j, uint64_t *n_entries, uint64_t *tau, int *speculative, ac_sketch_info *info) {
    if (!j) return fail(AC_E_PARAM, "job is NULL");
    ac_ctx *c = j->ctx;
    CK(cudaSetDevice(c->dev));
    Trace tr;
    int rc = finish_local(j, true);
    if (rc) return rc;
    tr.mark("end_local: finish_local");
    CK(cudaEventRecord(c->ev_job1, c->s_comp));
    CK(cudaStreamSynchronize(c->s_comp));
    prof_collect(c);                      // ac_last_job_times covers the local pass of a sharded job
    const JobDev &h = *c->h_jd;
    j->ended_local = true;
    if (n_entries) *n_entries = h.n_table;
    if (tau) *tau = h.tau;
    if (speculative) *speculative = 0;  // finish_local already repeated a failed speculation locally
    fill_info(j, h, info);
    return AC_OK;
}

extern "C" int ac_job_table_ranges(ac_job *j, uint64_t tau_cut, int n_ranges, uint64_t *range_begin) {
    if (!j || !range_begin || n_ranges < 1 || n_ranges > MAX_RANGES) return fail(AC_E_PARAM, "ac_job_table_ranges: bad argument (1..%d ranges)", MAX_RANGES);
    if (!j->ended_local) return fail(AC_E_STATE, "call ac_sketch_end_local first");
    ac_ctx *c = j->ctx;
    CK(cudaSetDevice(c->dev));
    // owner of hash x = floor(x * P / (tau_cut + 1)); boundaries b_r = ceil(r * (tau_cut + 1) / P)
    RangeBounds rb;
    memset(&rb, 0, sizeof(rb));
    rb.n = n_ranges; rb.tau_cut = tau_cut;
    const unsigned __int128 span = (unsigned __int128)tau_cut + 1;
    for (int r = 0; r < n_ranges; r++)
        rb.b[r] = (unsigned long long)((span * (unsigned)r + (unsigned)n_ranges - 1) / (unsigned)n_ranges);
    LaunchCtx lc{c->s_comp, &c->launches};
    if (j->table_pending) {
        Trace tr;
        // count-table mode: count the entries <= tau_cut per range, lay the ranges out back to back, fill them
        if (tau_cut > c->h_jd->tau) return fail(AC_E_PARAM, "tau_cut above this rank's threshold: counts beyond it are not exact");
        launch_table_range_count(lc, c->ct, rb, c->d_jd);
        CK(cudaGetLastError());
        int rc = read_jd(c);
        if (rc) return rc;
        tr.mark("ranges: count+read");
        RangeOffsets off;
        memset(&off, 0, sizeof(off));
        uint64_t run = 0;
        for (int r = 0; r < n_ranges; r++) { off.o[r] = (uint32_t)run; range_begin[r] = run; run += c->h_jd->range_cnt[r]; }
        off.o[n_ranges] = (uint32_t)run; range_begin[n_ranges] = run;
        if (run > c->sb.tab_cap) return fail(AC_E_NOMEM, "exchange table needs %llu entries, capacity %zu (raise AC_CAND_CAP)", (unsigned long long)run, c->sb.tab_cap);
        launch_table_range_scatter(lc, c->ct, rb, off, c->sb, c->d_jd);
        CK(cudaGetLastError());
        if (!c->stream_injected) CK(cudaStreamSynchronize(c->s_comp));
        c->h_jd->n_table = (uint32_t)run;    // what ac_job_table_export checks its range against
        j->table_pending = false;
        tr.mark("ranges: scatter launch");
        return AC_OK;
    }
    // the sort scratch is idle once the local pass has ended: its head receives the n_ranges+1 entry indices
    unsigned long long *d_out = c->sb.keys_b;
    launch_table_ranges(lc, c->sb, c->h_jd->n_table, rb, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(range_begin, d_out, (size_t)(n_ranges + 1) * 8, cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaStreamSynchronize(c->s_comp));
    return AC_OK;
}

extern "C" int ac_job_table_export(ac_job *j, uint64_t begin, uint64_t end, uint64_t *dev_hashes, uint32_t *dev_counts) {
    if (!j) return fail(AC_E_PARAM, "job is NULL");
    if (!j->ended_local) return fail(AC_E_STATE, "call ac_sketch_end_local first");
    if (j->table_pending) return fail(AC_E_STATE, "call ac_job_table_ranges first: it lays the exchange table out");
    ac_ctx *c = j->ctx;
    CK(cudaSetDevice(c->dev));
    if (end < begin || end > c->h_jd->n_table) return fail(AC_E_PARAM, "table range [%llu,%llu) outside 0..%u",
                                                            (unsigned long long)begin, (unsigned long long)end, c->h_jd->n_table);
    const size_t n = end - begin;
    if (n) {
        CK(cudaMemcpyAsync(dev_hashes, c->sb.tab_hash + begin, n * 8, cudaMemcpyDeviceToDevice, c->s_comp));
        CK(cudaMemcpyAsync(dev_counts, c->sb.tab_cnt + begin, n * 4, cudaMemcpyDeviceToDevice, c->s_comp));
    }
    if (!c->stream_injected) CK(cudaStreamSynchronize(c->s_comp));
    return AC_OK;
}

extern "C" int ac_merge_counts(ac_ctx *c, uint64_t *dev_hashes, uint32_t *dev_counts, uint64_t n, int min_copies, uint64_t *n_out) {
    if (!c || !n_out) return fail(AC_E_PARAM, "ac_merge_counts: NULL argument");
    CK(cudaSetDevice(c->dev));
    *n_out = 0;
    if (n == 0) return AC_OK;
    int rc = alloc_select(c, 1024);
    if (rc) return rc;
    if (n > c->sb.cap) return fail(AC_E_NOMEM, "merge of %llu entries exceeds capacity %zu", (unsigned long long)n, c->sb.cap);
    if (c->active && !c->active->ended_local) return fail(AC_E_STATE, "ac_merge_counts while a sketch job is hashing on this context");
    LaunchCtx lc{c->s_comp, &c->launches};
    // through the count table when it can hold the entries at <= 25 % load: sums by atomicAdd, the keys whose sum
    // reaches min_copies come back through the qualified list and only those (about s / P) are sorted
    const size_t slots_full = (size_t)c->ct.mask + 1;
    if (n * 4 <= slots_full && n < 0xffffffffull) {
        size_t slots = 1u << 16;
        while (slots < n * 4) slots <<= 1;
        CountTable ct = c->ct;
        ct.mask = (uint32_t)(slots - 1);
        ct.min_copies = (uint32_t)min_copies;
        ct.bshift = 0;
        JobDev z;
        memset(&z, 0, sizeof(z));
        z.tau = ~0ULL;                       // k3t_collect keeps keys <= tau: all of them
        *c->h_jd = z;
        CK(cudaMemcpyAsync(c->d_jd, c->h_jd, sizeof(JobDev), cudaMemcpyHostToDevice, c->s_comp));
        CK(cudaMemsetAsync(ct.key, 0xff, slots * 8, c->s_comp));
        CK(cudaMemsetAsync(ct.cnt, 0, slots * 4, c->s_comp));
        launch_merge_insert(lc, ct, (const unsigned long long *)dev_hashes, dev_counts, (uint32_t)n, c->d_jd);
        CK(cudaGetLastError());
        rc = read_jd(c);
        if (rc) return rc;
        if (!c->h_jd->overflow) {
            const uint32_t nq = c->h_jd->n_qlist;
            if (nq) {
                launch_table_collect(lc, ct, c->sb, c->d_jd, nq);
                launch_table_finish(lc, c->sb, c->d_jd, nq, nq, key_bits_for(c->h_jd->max_hash), 0x7fffffff, 1);   // sorted pairs -> (tab_hash, tab_cnt)
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(dev_hashes, c->sb.tab_hash, (size_t)nq * 8, cudaMemcpyDeviceToDevice, c->s_comp));
                CK(cudaMemcpyAsync(dev_counts, c->sb.tab_cnt, (size_t)nq * 4, cudaMemcpyDeviceToDevice, c->s_comp));
            }
            if (!c->stream_injected) CK(cudaStreamSynchronize(c->s_comp));
            *n_out = nq;
            return AC_OK;
        }
        // (a key equal to the table's "empty" marker, or a pathological probe sequence: sort instead)
    }
    CK(cudaMemcpyAsync(c->sb.keys_a, dev_hashes, n * 8, cudaMemcpyDeviceToDevice, c->s_comp));
    CK(cudaMemcpyAsync(c->sb.vals_a, dev_counts, n * 4, cudaMemcpyDeviceToDevice, c->s_comp));
    launch_merge_counts(lc, c->sb, c->d_jd, n, min_copies);
    CK(cudaGetLastError());
    rc = read_jd(c);
    if (rc) return rc;
    const uint64_t nq = c->h_jd->n_qualified;
    if (nq) {
        CK(cudaMemcpyAsync(dev_hashes, c->sb.keys_a, nq * 8, cudaMemcpyDeviceToDevice, c->s_comp));
        CK(cudaMemcpyAsync(dev_counts, c->sb.vals_a, nq * 4, cudaMemcpyDeviceToDevice, c->s_comp));
        CK(cudaStreamSynchronize(c->s_comp));
    }
    *n_out = nq;
    return AC_OK;
}

__global__ void k_set_nsketch(JobDev *jd, uint32_t n) { jd->n_sketch = n; }

extern "C" int ac_sketch_from_table(ac_ctx *c, const uint64_t *dev_hashes, const uint32_t *dev_counts, uint64_t n, int k,
                                    int sketch_size, ac_sketch **out, ac_sketch_info *info) {
    if (!c || !out) return fail(AC_E_PARAM, "ac_sketch_from_table: NULL argument");
    *out = nullptr;
    CK(cudaSetDevice(c->dev));
    const uint64_t ns = std::min<uint64_t>(n, (uint64_t)sketch_size);
    const int use64 = k > 16;
    if (ns == 0) return fail(AC_E_NO_RECORDS, "did not find sequence records (no hash reached the minimum copy number)");
    ac_sketch *sk = nullptr;
    int rc = make_sketch(c, (const unsigned long long *)dev_hashes, dev_counts, ns, k, sketch_size, use64 ? 64 : 32, 0, &sk);
    if (rc) return rc;
    k_set_nsketch<<<1, 1, 0, c->s_comp>>>(c->d_jd, (uint32_t)ns);
    LaunchCtx lc{c->s_comp, &c->launches};
    SelectBufs tmp = c->sb;
    tmp.sk_hash = sk->d_hash; tmp.sk_cnt = sk->d_cnt;
    launch_finalize(lc, tmp, c->d_jd, use64);
    c->launches += 1;
    rc = read_jd(c);
    if (rc) { ac_sketch_free(sk); return rc; }
    const JobDev &h = *c->h_jd;
    sk->length = (uint64_t)h.genome_size;
    if (info) {
        memset(info, 0, sizeof(*info));
        info->n_hashes = ns; info->max_hash = h.max_hash; info->sum_counts = h.sum_counts;
        info->genome_size = h.genome_size; info->coverage = h.coverage; info->hash_bits = use64 ? 64 : 32; info->passes = 1;
    }
    *out = sk;
    return AC_OK;
}

// ------------------------------------------------------------------ peer-memory exchange (ac_xchg.cu)
struct ac_xchg {
    ac_ctx *ctx;
    XLayout L;
    XPeers peers;
    bool opened[X_MAX];            // peers.base[r] came from cudaIpcOpenMemHandle
    unsigned char *win;
    XLocal *d_xl, *h_xl;
    JobDev *h_jd2;                 // pinned copy of the job state read together with h_xl
    unsigned long long epoch;
    int phase;                     // 0 idle, 1 after _begin, 2 after _merge
    ac_job *job;
    int k, s, m, use64;
    int local_rc;
    char local_err[512];
    unsigned long long counters[5];
    SketchBuf skb;
};

extern "C" void ac_xchg_destroy(ac_xchg *x) {
    if (!x) return;
    cudaSetDevice(x->ctx->dev);
    cudaDeviceSynchronize();
    for (int r = 0; r < X_MAX; r++) if (x->opened[r]) cudaIpcCloseMemHandle(x->peers.base[r]);
    if (x->job) job_free(x->job);
    if (x->skb.h || x->skb.c) sketch_buf_put(x->ctx, x->ctx->dev, x->skb);
    cudaFree(x->win); cudaFree(x->d_xl); cudaFreeHost(x->h_xl); cudaFreeHost(x->h_jd2);
    cudaGetLastError();
    delete x;
}

extern "C" int ac_xchg_create(ac_ctx *c, int rank, int world, int sketch_cap, uint64_t inbox_entries, ac_xchg **out) {
    if (!c || !out) return fail(AC_E_PARAM, "ac_xchg_create: NULL argument");
    *out = nullptr;
    if (world < 1 || world > X_MAX || rank < 0 || rank >= world) return fail(AC_E_PARAM, "ac_xchg_create: rank %d of %d (at most %d ranks)", rank, world, X_MAX);
    if (sketch_cap < 1 || sketch_cap > (1 << 24)) return fail(AC_E_PARAM, "sketch capacity %d outside 1..2^24", sketch_cap);
    CK(cudaSetDevice(c->dev));
    if (!c->sb_ok) {            // (with a job in flight the scratch exists already; the window brings its own sketch buffers)
        int rc = alloc_select(c, (size_t)sketch_cap);
        if (rc) return rc;
    }
    size_t cap_in = inbox_entries ? (size_t)inbox_entries : ((size_t)2 << 20);
    cap_in = std::min(cap_in, c->sb.cap / (size_t)world) & ~(size_t)3;     // the staging area is the sort scratch
    if (cap_in < 1024) return fail(AC_E_PARAM, "inbox capacity too small");
    ac_xchg *x = new ac_xchg();
    x->ctx = c;
    x->L = xchg_layout((uint32_t)rank, (uint32_t)world, (uint32_t)cap_in, (uint32_t)sketch_cap);
    cudaError_t e = cudaMalloc(&x->win, x->L.bytes);
    if (e == cudaSuccess) e = cudaMemset(x->win, 0, (sizeof(XCtl) + 4095) & ~(size_t)4095);
    if (e == cudaSuccess) e = cudaMalloc(&x->d_xl, sizeof(XLocal));
    if (e == cudaSuccess) e = cudaMemset(x->d_xl, 0, sizeof(XLocal));
    if (e == cudaSuccess) e = cudaHostAlloc(&x->h_xl, sizeof(XLocal), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(&x->h_jd2, sizeof(JobDev), cudaHostAllocDefault);
    if (e != cudaSuccess) { ac_xchg_destroy(x); return fail(e == cudaErrorMemoryAllocation ? AC_E_NOMEM : AC_E_CUDA, "exchange window of %zu bytes: %s", x->L.bytes, cudaGetErrorString(e)); }
    x->peers.base[rank] = x->win;
    // the final sketch is assembled straight into the buffers the ac_sketch will own; they are taken here and re-armed
    // after every exchange, so no allocation (a potential device-wide synchronisation) sits between an exchange's launches
    e = sketch_buf_get(c, (uint64_t)x->L.s_cap, &x->skb);
    if (e != cudaSuccess) { ac_xchg_destroy(x); return fail(AC_E_NOMEM, "sketch buffers: %s", cudaGetErrorString(e)); }
    *out = x;
    return AC_OK;
}

extern "C" int ac_xchg_ipc_handle(ac_xchg *x, void *out64) {
    if (!x || !out64) return fail(AC_E_PARAM, "ac_xchg_ipc_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    CK(cudaSetDevice(x->ctx->dev));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, x->win));
    memcpy(out64, &h, 64);
    return AC_OK;
}

extern "C" int ac_xchg_open_peers(ac_xchg *x, const void *handles) {
    if (!x || !handles) return fail(AC_E_PARAM, "ac_xchg_open_peers: NULL argument");
    CK(cudaSetDevice(x->ctx->dev));
    for (uint32_t r = 0; r < x->L.P; r++) {
        if (r == x->L.rank || x->peers.base[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)r, 64);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->peers.base[r] = (unsigned char *)p;
        x->opened[r] = true;
    }
    return AC_OK;
}

extern "C" int ac_xchg_set_peer_local(ac_xchg *x, int r, ac_xchg *other) {
    if (!x || !other || r < 0 || r >= (int)x->L.P) return fail(AC_E_PARAM, "ac_xchg_set_peer_local: bad argument");
    if (other->L.P != x->L.P || other->L.cap_in != x->L.cap_in || other->L.s_cap != x->L.s_cap || (int)other->L.rank != r)
        return fail(AC_E_PARAM_MISMATCH, "windows of one exchange must share world size and capacities");
    if (other->ctx->dev != x->ctx->dev) {
        CK(cudaSetDevice(x->ctx->dev));
        cudaError_t e = cudaDeviceEnablePeerAccess(other->ctx->dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(AC_E_CUDA, "no peer access from device %d to %d: %s", x->ctx->dev, other->ctx->dev, cudaGetErrorString(e));
        cudaGetLastError();
    }
    x->peers.base[r] = other->win;
    return AC_OK;
}

// The owner merge runs in a 2 Mi-slot window of the count table (24 MB to clear instead of 96): an owner receives ~1/P of
// every rank's entries <= tau_cut (C2 on 8 GPUs: ~50 k entries).  An inbox that does not fit raises the table's overflow
// flag and ac_sketch_exchange_merge falls back to the sort-based merge.  AC_XCHG_MERGE_LOG2 resizes it (tests: 10).
static CountTable xchg_merge_table(const ac_ctx *c, int min_copies) {
    const char *e = getenv("AC_XCHG_MERGE_LOG2");
    int log2_slots = e ? atoi(e) : 21;
    if (log2_slots < 8) log2_slots = 8;
    CountTable ctm = c->ct;
    const int l2 = std::min(log2_slots, c->ct_log2);
    ctm.mask = (uint32_t)(((size_t)1 << l2) - 1);
    ctm.min_copies = (uint32_t)min_copies; ctm.bshift = 0;
    return ctm;
}

extern "C" int ac_sketch_exchange_begin(ac_ctx *c, ac_job *j, ac_xchg *x, int k, int sketch_size, int min_copies, int failed) {
    if (!c || !x || x->ctx != c) return fail(AC_E_PARAM, "ac_sketch_exchange_begin: bad argument");
    if (x->phase != 0) return fail(AC_E_STATE, "an exchange is already in flight on this window");
    if (j && j->ctx != c) return fail(AC_E_PARAM, "job belongs to another context");
    if (sketch_size < 1 || (uint32_t)sketch_size > x->L.s_cap) return fail(AC_E_PARAM, "sketch size %d exceeds the window's capacity %u", sketch_size, x->L.s_cap);
    for (uint32_t r = 0; r < x->L.P; r++) if (!x->peers.base[r]) return fail(AC_E_STATE, "window of rank %u has not been opened", r);
    CK(cudaSetDevice(c->dev));
    x->k = k; x->s = sketch_size; x->m = min_copies; x->use64 = k > 16;
    x->local_rc = AC_OK; x->local_err[0] = 0; x->job = j;
    memset(x->counters, 0, sizeof(x->counters));
    int have_table = 0, table_mode = 0;
    if (j && !failed) {
        int rc = finish_local(j, true);
        if (rc) { x->local_rc = rc; snprintf(x->local_err, sizeof(x->local_err), "%s", g_err); }
        else {
            const JobDev &h = *c->h_jd;
            x->counters[0] = h.n_records; x->counters[1] = h.n_kept; x->counters[2] = h.n_bases - h.n_short_len;
            x->counters[3] = h.n_kmers; x->counters[4] = h.cand_total;
            have_table = 1; table_mode = j->table_pending ? 1 : 0;
            j->ended_local = true;
        }
        cudaEventRecord(c->ev_job1, c->s_comp);      // ac_last_job_times: the local pass of a sharded job
    } else if (!failed && !j) {
        return fail(AC_E_PARAM, "ac_sketch_exchange_begin: no job (pass failed = 1 for a rank that could not feed its reads)");
    } else {
        x->local_rc = AC_E_STATE; snprintf(x->local_err, sizeof(x->local_err), "this rank failed before the exchange");
    }
    x->epoch++;
    LaunchCtx lc{c->s_comp, &c->launches};
    const CountTable ctm = xchg_merge_table(c, min_copies);
    launch_xchg_phase1(lc, x->peers, x->win, x->L, x->epoch, x->d_xl, c->d_jd, have_table, table_mode, x->local_rc ? 1ull : 0ull,
                       c->ct, c->sb, ctm);
    CK(cudaGetLastError());
    x->phase = 1;
    return AC_OK;
}

static int xchg_read(ac_xchg *x) {
    ac_ctx *c = x->ctx;
    CK(cudaMemcpyAsync(x->h_xl, x->d_xl, sizeof(XLocal), cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaMemcpyAsync(x->h_jd2, c->d_jd, sizeof(JobDev), cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaStreamSynchronize(c->s_comp));
    return AC_OK;
}

extern "C" int ac_sketch_exchange_merge(ac_xchg *x) {
    if (!x) return fail(AC_E_PARAM, "ac_sketch_exchange_merge: NULL argument");
    if (x->phase != 1) return fail(AC_E_STATE, "call ac_sketch_exchange_begin first");
    ac_ctx *c = x->ctx;
    CK(cudaSetDevice(c->dev));
    int rc = xchg_read(x);        // the one number the host needs: how many qualified keys this owner holds
    if (rc) return rc;
    LaunchCtx lc{c->s_comp, &c->launches};
    uint64_t nq = 0;
    const unsigned long long *keys = c->sb.tab_hash;
    const uint32_t *cnts = c->sb.tab_cnt;
    if (x->h_jd2->overflow) {
        // the table could not take the entries (or one of them is the table's "empty" marker): sort-based merge
        const uint64_t n = x->h_xl->n_in_total;
        if (n > c->sb.cap) return fail(AC_E_NOMEM, "merge of %llu entries exceeds capacity %zu", (unsigned long long)n, c->sb.cap);
        if (n) {
            launch_xchg_compact_inbox(lc, x->win, x->L, x->d_xl, c->sb);
            launch_merge_counts(lc, c->sb, c->d_jd, n, x->m);
            CK(cudaGetLastError());
            rc = read_jd(c);
            if (rc) return rc;
            nq = c->h_jd->n_qualified;
        }
        keys = c->sb.keys_a; cnts = c->sb.vals_a;
    } else {
        nq = x->h_jd2->n_qlist;
        if (nq) {
            const CountTable ctm = xchg_merge_table(c, x->m);
            launch_table_collect(lc, ctm, c->sb, c->d_jd, (uint32_t)nq);
            launch_table_finish(lc, c->sb, c->d_jd, (uint32_t)nq, (uint32_t)nq, key_bits_for(x->h_jd2->max_hash), 0x7fffffff, 1);
            CK(cudaGetLastError());
        }
    }
    const uint32_t n_keep = (uint32_t)std::min<uint64_t>(nq, (uint64_t)x->s);
    if (x->skb.h == nullptr || x->skb.cap < (uint64_t)x->s) return fail(AC_E_INTERNAL, "exchange window has no sketch buffers");
    XHeader hdr;
    hdr.v[0] = n_keep; hdr.v[1] = nq;
    for (int i = 0; i < 5; i++) hdr.v[2 + i] = x->counters[i];
    hdr.v[7] = x->local_rc ? (unsigned long long)XERR_PEER : 0ull;
    launch_xchg_phase2(lc, x->peers, x->win, x->L, x->epoch, x->d_xl, c->d_jd, keys, cnts, n_keep, hdr, (uint32_t)x->s, x->skb.h, x->skb.c);
    SelectBufs tmp = c->sb;
    tmp.sk_hash = x->skb.h; tmp.sk_cnt = x->skb.c;
    launch_finalize(lc, tmp, c->d_jd, x->use64);
    CK(cudaGetLastError());
    x->phase = 2;
    return AC_OK;
}

extern "C" int ac_sketch_exchange_end(ac_xchg *x, ac_sketch **out, ac_sketch_info *info, uint64_t *stats /*[4] or NULL*/) {
    if (!x || !out) return fail(AC_E_PARAM, "ac_sketch_exchange_end: NULL argument");
    *out = nullptr;
    if (x->phase != 2) return fail(AC_E_STATE, "call ac_sketch_exchange_merge first");
    ac_ctx *c = x->ctx;
    CK(cudaSetDevice(c->dev));
    int rc = xchg_read(x);
    x->phase = 0;
    if (x->job) { job_free(x->job); x->job = nullptr; }
    prof_collect(c);
    if (rc) return rc;
    const XLocal &xl = *x->h_xl;
    const JobDev &h = *x->h_jd2;
    if (stats) { stats[0] = xl.tau_cut; stats[1] = xl.n_in_total; stats[2] = xl.qualified_total; stats[3] = xl.err; }
    if (x->local_rc) return fail(x->local_rc, "%s", x->local_err);
    if (xl.err & XERR_TIMEOUT) {
        // which wait gave up: the flags of this rank's window next to the epoch they should have reached
        XCtl ctl;
        memset(&ctl, 0, sizeof(ctl));
        cudaMemcpy(&ctl, x->win, sizeof(ctl), cudaMemcpyDeviceToHost);
        char fl[256];
        int o = 0;
        for (uint32_t r = 0; r < x->L.P && o < 200; r++)
            o += snprintf(fl + o, sizeof(fl) - o, " r%u:%llu/%llu/%llu", r, ctl.tau_flag[r], ctl.in_flag[r], ctl.ga_flag[r]);
        return fail(AC_E_PEER, "exchange %llu timed out waiting for another rank (err bits %u; tau/inbox/gather flags seen per rank:%s)",
                    x->epoch, xl.err, fl);
    }
    if (xl.err & XERR_OVERFLOW) return fail(AC_E_NOMEM, "a rank's share of the exchange exceeds the inbox capacity of %u entries (raise AC_XCHG_INBOX)", x->L.cap_in);
    if (xl.err) return fail(AC_E_PEER, "sharded sketch aborted: another rank of the exchange failed (err bits %u)", xl.err);
    if (xl.n_final == 0) return fail(AC_E_NO_RECORDS, "did not find sequence records (no hash reached the minimum copy number on any rank)");
    ac_sketch *sk = new ac_sketch();
    sk->ctx = c; sk->dev = c->dev; sk->n = xl.n_final; sk->k = x->k; sk->s = x->s; sk->hash_bits = x->use64 ? 64 : 32;
    sk->length = (uint64_t)h.genome_size;
    sk->d_hash = x->skb.h; sk->d_cnt = x->skb.c; sk->cap = x->skb.cap;
    x->skb = SketchBuf{nullptr, nullptr, 0};       // the buffers now belong to the sketch; re-arm the window
    if (sketch_buf_get(c, (uint64_t)x->L.s_cap, &x->skb) != cudaSuccess) { cudaGetLastError(); x->skb = SketchBuf{nullptr, nullptr, 0}; }
    if (info) {
        memset(info, 0, sizeof(*info));
        info->n_hashes = xl.n_final; info->n_records = xl.counters[0]; info->n_kept = xl.counters[1]; info->total_len = xl.counters[2];
        info->n_kmers = xl.counters[3]; info->n_candidates = xl.counters[4];
        info->max_hash = h.max_hash; info->sum_counts = h.sum_counts; info->final_tau = xl.tau_cut;
        info->genome_size = h.genome_size; info->coverage = h.coverage; info->hash_bits = x->use64 ? 64 : 32; info->passes = 1;
    }
    *out = sk;
    return AC_OK;
}

// ------------------------------------------------------------------ sketches
extern "C" int ac_sketch_from_hashes(ac_ctx *c, const uint64_t *hashes, uint64_t n, int k, int sketch_size, uint64_t length, ac_sketch **out) {
    if (!c || !out || (n && !hashes)) return fail(AC_E_PARAM, "ac_sketch_from_hashes: NULL argument");
    *out = nullptr;
    if (k < 1 || k > 32) return fail(AC_E_PARAM, "k = %d outside 1..32", k);
    CK(cudaSetDevice(c->dev));
    ac_sketch *sk = new ac_sketch();
    sk->ctx = c; sk->dev = c->dev; sk->n = n; sk->k = k; sk->s = sketch_size; sk->hash_bits = k > 16 ? 64 : 32; sk->length = length;
    SketchBuf b;
    cudaError_t e = sketch_buf_get(c, n, &b);
    sk->d_hash = b.h; sk->d_cnt = b.c; sk->cap = b.cap;
    if (e == cudaSuccess && n) e = cudaMemcpy(sk->d_hash, hashes, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n) e = cudaMemset(sk->d_cnt, 0, n * 4);
    if (e != cudaSuccess) { sketch_buf_put(c, c->dev, b); delete sk; return fail(AC_E_CUDA, "sketch upload failed: %s", cudaGetErrorString(e)); }
    *out = sk;
    return AC_OK;
}
extern "C" uint64_t ac_sketch_size(const ac_sketch *sk) { return sk ? sk->n : 0; }
extern "C" uint64_t ac_sketch_length(const ac_sketch *sk) { return sk ? sk->length : 0; }
extern "C" int ac_sketch_hashes(const ac_sketch *sk, uint64_t *hashes, uint32_t *counts) {
    if (!sk) return fail(AC_E_PARAM, "sketch is NULL");
    CK(cudaSetDevice(sk->dev));
    if (sk->n && hashes) CK(cudaMemcpy(hashes, sk->d_hash, sk->n * 8, cudaMemcpyDeviceToHost));
    if (sk->n && counts) CK(cudaMemcpy(counts, sk->d_cnt, sk->n * 4, cudaMemcpyDeviceToHost));
    return AC_OK;
}
extern "C" void ac_sketch_free(ac_sketch *sk) {
    if (!sk) return;
    sketch_buf_put(sk->ctx, sk->dev, SketchBuf{sk->d_hash, sk->d_cnt, sk->cap});
    delete sk;
}

// ------------------------------------------------------------------ references + dist
extern "C" int ac_refs_load(ac_ctx *c, const uint64_t *hashes, const uint64_t *offsets, const uint64_t *lengths, int n_refs,
                            int k, int sketch_size, int hash_bits, ac_refs **out) {
    if (!c || !out || !offsets || !lengths || n_refs < 0) return fail(AC_E_PARAM, "ac_refs_load: bad argument");
    *out = nullptr;
    if (k < 1 || k > 32) return fail(AC_E_PARAM, "k = %d outside 1..32", k);
    if (hash_bits != 32 && hash_bits != 64) return fail(AC_E_PARAM, "hash_bits = %d", hash_bits);
    CK(cudaSetDevice(c->dev));
    const uint64_t total = offsets[n_refs];
    uint64_t max_n = 0;
    for (int i = 0; i < n_refs; i++) {
        if (offsets[i + 1] < offsets[i]) return fail(AC_E_PARAM, "offsets not ascending at %d", i);
        max_n = std::max<uint64_t>(max_n, offsets[i + 1] - offsets[i]);
    }
    ac_refs *r = new ac_refs();
    r->ctx = c; r->dev = c->dev; r->n_refs = n_refs; r->k = k; r->s = sketch_size; r->hash_bits = hash_bits; r->max_n = max_n;
    r->d_hash = r->d_off = r->d_len = nullptr;
    cudaError_t e = cudaMalloc(&r->d_hash, (total ? total : 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_off, (size_t)(n_refs + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_len, (size_t)(n_refs ? n_refs : 1) * 8);
    if (e == cudaSuccess && total) e = cudaMemcpy(r->d_hash, hashes, total * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_off, offsets, (size_t)(n_refs + 1) * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_refs) e = cudaMemcpy(r->d_len, lengths, (size_t)n_refs * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(r->d_hash); cudaFree(r->d_off); cudaFree(r->d_len); delete r; return fail(AC_E_CUDA, "reference upload failed: %s", cudaGetErrorString(e)); }
    *out = r;
    return AC_OK;
}
extern "C" int ac_refs_count(const ac_refs *r) { return r ? r->n_refs : 0; }
extern "C" void ac_refs_free(ac_refs *r) {
    if (!r) return;
    cudaSetDevice(r->dev);
    cudaFree(r->d_hash); cudaFree(r->d_off); cudaFree(r->d_len);
    delete r;
}

static int ensure_dist(ac_ctx *c, size_t n_out, size_t n_q) {
    if (n_out > c->dist_cap) {
        cudaFree(c->d_dist); cudaFree(c->d_p); cudaFree(c->d_common); cudaFree(c->d_denom);
        cudaFreeHost(c->h_dist); cudaFreeHost(c->h_p); cudaFreeHost(c->h_common); cudaFreeHost(c->h_denom);
        c->dist_cap = 0;
        const size_t cap = std::max<size_t>(n_out, 1024);
        if (!c->d_dist_parts) CK(cudaMalloc(&c->d_dist_parts, (size_t)DIST_SPLIT_MAX_PAIRS * 64));
        CK(cudaMalloc(&c->d_dist, cap * 8)); CK(cudaMalloc(&c->d_p, cap * 8));
        CK(cudaMalloc(&c->d_common, cap * 8)); CK(cudaMalloc(&c->d_denom, cap * 8));
        CK(cudaHostAlloc(&c->h_dist, cap * 8, cudaHostAllocDefault)); CK(cudaHostAlloc(&c->h_p, cap * 8, cudaHostAllocDefault));
        CK(cudaHostAlloc(&c->h_common, cap * 8, cudaHostAllocDefault)); CK(cudaHostAlloc(&c->h_denom, cap * 8, cudaHostAllocDefault));
        c->dist_cap = cap;
    }
    if (n_q > c->q_cap) {
        cudaFree(c->d_qmeta); cudaFreeHost(c->h_qmeta);
        c->q_cap = 0;
        const size_t cap = std::max<size_t>(n_q, 64);
        CK(cudaMalloc(&c->d_qmeta, cap * 3 * 8));
        CK(cudaHostAlloc(&c->h_qmeta, cap * 3 * 8, cudaHostAllocDefault));
        c->q_cap = cap;
    }
    return AC_OK;
}

extern "C" int ac_dist_batch(ac_ctx *c, const ac_refs *refs, const ac_sketch *const *queries, int n_q, double *distance,
                             double *pvalue, uint64_t *common, uint64_t *denom) {
    if (!c || !refs || !queries || n_q < 0) return fail(AC_E_PARAM, "ac_dist: bad argument");
    CK(cudaSetDevice(c->dev));
    const size_t n_out = (size_t)n_q * refs->n_refs;
    if (n_out == 0) return AC_OK;
    int rc = ensure_dist(c, n_out, n_q);
    if (rc) return rc;
    uint64_t max_q = 0;
    int s_min = refs->s;
    for (int i = 0; i < n_q; i++) {
        const ac_sketch *q = queries[i];
        if (!q) return fail(AC_E_PARAM, "query %d is NULL", i);
        if (q->k != refs->k || q->hash_bits != refs->hash_bits)
            return fail(AC_E_PARAM_MISMATCH, "query sketch k=%d/%d-bit does not match reference k=%d/%d-bit", q->k, q->hash_bits, refs->k, refs->hash_bits);
        if (q->s != queries[0]->s) return fail(AC_E_PARAM_MISMATCH, "queries of one batch must share the sketch size");
        c->h_qmeta[i] = (unsigned long long)(uintptr_t)q->d_hash;
        c->h_qmeta[c->q_cap + i] = q->n;
        c->h_qmeta[2 * c->q_cap + i] = q->length;
        max_q = std::max<uint64_t>(max_q, q->n);
    }
    s_min = std::min(s_min, queries[0]->s);
    if (max_q > (1u << 20)) return fail(AC_E_PARAM, "query sketch of %llu hashes exceeds 2^20", (unsigned long long)max_q);
    CK(cudaMemcpyAsync(c->d_qmeta, c->h_qmeta, c->q_cap * 3 * 8, cudaMemcpyHostToDevice, c->s_comp));
    LaunchCtx lc{c->s_comp, &c->launches};
    launch_dist(lc, refs->d_hash, refs->d_off, refs->d_len, refs->n_refs, (const unsigned long long *const *)c->d_qmeta,
                c->d_qmeta + c->q_cap, c->d_qmeta + 2 * c->q_cap, n_q, refs->k, (unsigned long long)s_min, max_q, c->d_dist,
                c->d_p, c->d_common, c->d_denom, c->d_dist_parts);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->h_dist, c->d_dist, n_out * 8, cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaMemcpyAsync(c->h_p, c->d_p, n_out * 8, cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaMemcpyAsync(c->h_common, c->d_common, n_out * 8, cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaMemcpyAsync(c->h_denom, c->d_denom, n_out * 8, cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaStreamSynchronize(c->s_comp));
    if (distance) memcpy(distance, c->h_dist, n_out * 8);
    if (pvalue) memcpy(pvalue, c->h_p, n_out * 8);
    if (common) memcpy(common, c->h_common, n_out * 8);
    if (denom) memcpy(denom, c->h_denom, n_out * 8);
    return AC_OK;
}

extern "C" int ac_dist_pairs(ac_ctx *c, const ac_refs *a, int i0, int ni, const ac_refs *b, int j0, int nj, double *distance,
                             double *pvalue, uint64_t *common, uint64_t *denom) {
    if (!c || !a || !b) return fail(AC_E_PARAM, "ac_dist_pairs: NULL argument");
    if (i0 < 0 || ni < 0 || j0 < 0 || nj < 0 || i0 + ni > a->n_refs || j0 + nj > b->n_refs)
        return fail(AC_E_PARAM, "ac_dist_pairs: tile [%d,%d) x [%d,%d) outside %d x %d sketches", i0, i0 + ni, j0, j0 + nj, a->n_refs, b->n_refs);
    if (a->k != b->k || a->hash_bits != b->hash_bits)
        return fail(AC_E_PARAM_MISMATCH, "sketch sets disagree: k=%d/%d-bit vs k=%d/%d-bit", a->k, a->hash_bits, b->k, b->hash_bits);
    CK(cudaSetDevice(c->dev));
    const size_t n_out = (size_t)ni * (size_t)nj;
    if (n_out == 0) return AC_OK;
    int rc = ensure_dist(c, n_out, 1);
    if (rc) return rc;
    LaunchCtx lc{c->s_comp, &c->launches};
    launch_dist_pairs(lc, a->d_hash, a->d_off, a->d_len, (uint32_t)i0, (uint32_t)ni, b->d_hash, b->d_off, b->d_len, (uint32_t)j0,
                      (uint32_t)nj, a->k, (unsigned long long)std::min(a->s, b->s), c->d_dist, c->d_p, c->d_common, c->d_denom);
    CK(cudaGetLastError());
    // results travel through the context's pinned buffers (a device-to-host copy into pageable memory runs at a few
    // GB/s: 0.9 ms for the 2 MB of a 256 x 256 tile, a tenth of the kernel), then one memcpy each into the caller's arrays
    if (distance) CK(cudaMemcpyAsync(c->h_dist, c->d_dist, n_out * 8, cudaMemcpyDeviceToHost, c->s_comp));
    if (pvalue) CK(cudaMemcpyAsync(c->h_p, c->d_p, n_out * 8, cudaMemcpyDeviceToHost, c->s_comp));
    if (common) CK(cudaMemcpyAsync(c->h_common, c->d_common, n_out * 8, cudaMemcpyDeviceToHost, c->s_comp));
    if (denom) CK(cudaMemcpyAsync(c->h_denom, c->d_denom, n_out * 8, cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaStreamSynchronize(c->s_comp));
    if (distance) memcpy(distance, c->h_dist, n_out * 8);
    if (pvalue) memcpy(pvalue, c->h_p, n_out * 8);
    if (common) memcpy(common, c->h_common, n_out * 8);
    if (denom) memcpy(denom, c->h_denom, n_out * 8);
    return AC_OK;
}

extern "C" int ac_dist(ac_ctx *c, const ac_refs *refs, const ac_sketch *query, double *distance, double *pvalue,
                       uint64_t *common, uint64_t *denom) {
    const ac_sketch *q[1] = {query};
    return ac_dist_batch(c, refs, q, 1, distance, pvalue, common, denom);
}
