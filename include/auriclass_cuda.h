/*
 * auriclass_cuda.h -- C ABI of libauriclass_cuda.so
 *
 * B200 (sm_100a) replacement for the four `mash` subprocess calls AuriClass makes
 * (paths relative to the upstream repository RIVM-bioinformatics/auriclass v0.5.4):
 *
 *   mash sketch -r -m M -k K -s S files...   auriclass/classes.py:576-596   -> ac_sketch_begin/feed/end (mode AC_MODE_READS)
 *   mash sketch -k K -s S files...           auriclass/classes.py:696-713   -> ac_sketch_begin/feed/end (mode AC_MODE_FASTA, one job per file)
 *   mash dist REF.msh QUERY.msh              auriclass/classes.py:92-104    -> ac_refs_load + ac_dist
 *   mash -h (presence probe)                 auriclass/general.py:198-205   -> ac_probe
 *   tmpfile.msh hand-off sketch -> dist      auriclass/classes.py:620-625   -> ac_sketch handle (device resident), ac_sketch_hashes
 *
 * Plain C: pointers and sizes only, no exceptions cross the boundary, every
 * call returns an AC_* status (0 = ok) and leaves a message for ac_last_error().
 * The library never computes on the CPU: without a CUDA device every compute
 * entry point fails with AC_E_CUDA.
 *
 * Ownership: the caller owns all host buffers (use ac_alloc_pinned for the
 * staging buffers handed to ac_sketch_feed); the library owns device memory,
 * jobs, sketches and reference sets, released by the matching *_free/_destroy.
 * Threading: one ac_ctx per calling thread; a context is not thread safe.
 * ctypes releases the GIL for the duration of each call, so a Python thread
 * can inflate the next gzip block while a feed is in flight.
 */
#ifndef AURICLASS_CUDA_H
#define AURICLASS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AC_ABI_VERSION 1

/* status codes */
enum {
    AC_OK = 0,
    AC_E_PARAM = 1,          /* bad argument (k outside 1..32, s < 1, ...) */
    AC_E_CUDA = 2,           /* CUDA runtime error / no device */
    AC_E_NOMEM = 3,          /* device or pinned allocation failed */
    AC_E_NO_RECORDS = 4,     /* no record with length >= k, or no hash reached min_copies:
                                Mash prints `ERROR: Did not find fasta records in ...`
                                (auriclass/classes.py:597-600 turns that into ValueError) */
    AC_E_FORMAT = 5,         /* input is not FASTA / 4-line FASTQ as the GPU parser requires */
    AC_E_PARAM_MISMATCH = 6, /* query and reference sketches disagree in k / hash width */
    AC_E_STATE = 7,          /* call sequence error (feed after end, ...) */
    AC_E_INTERNAL = 8,
    AC_E_PEER = 9            /* sharded sketch: another rank of the exchange failed or did not answer */
};

/* ac_sketch_begin mode */
enum {
    AC_MODE_READS = 0, /* `mash sketch -r -m M`: every fed file merges into ONE sketch      */
    AC_MODE_FASTA = 1  /* `mash sketch`: one job per file, min_copies forced to 1           */
};

typedef struct ac_ctx ac_ctx;       /* device context: streams, scratch arenas              */
typedef struct ac_job ac_job;       /* a sketch under construction                          */
typedef struct ac_sketch ac_sketch; /* finished bottom-s sketch, device resident            */
typedef struct ac_refs ac_refs;     /* reference sketches of one .msh, device resident      */
typedef struct ac_xchg ac_xchg;     /* this rank's window of the peer-memory exchange       */

/* What `mash sketch` reports on stderr / stores in the .msh, plus parser counters. */
typedef struct ac_sketch_info {
    uint64_t n_hashes;     /* sketch length = min(s, #distinct hashes with count >= m)       */
    uint64_t n_records;    /* records seen by the parser                                     */
    uint64_t n_kept;       /* records with sequence length >= k (Mash skips the others)      */
    uint64_t total_len;    /* sum of sequence lengths over kept records (FASTA `length`)     */
    uint64_t n_kmers;      /* valid windows hashed (no byte outside ACGTacgt)                */
    uint64_t n_candidates; /* hash occurrences that passed the running threshold             */
    uint64_t max_hash;     /* largest hash in the sketch                                     */
    uint64_t sum_counts;   /* sum of exact occurrence counts over the sketch                 */
    uint64_t final_tau;    /* threshold the last pass ended with                             */
    double genome_size;    /* 2^64 (2^32 for k<=16) * n / max_hash : `Estimated genome size` */
    double coverage;       /* sum_counts / n                      : `Estimated coverage`     */
    int32_t hash_bits;     /* 64 for k > 16 else 32                                          */
    int32_t passes;        /* hashing passes run (1 unless threshold speculation failed)     */
    uint64_t n_bases;      /* sequence characters of ALL records (what pyfastx `Fasta(...).size`
                              reports: auriclass/classes.py:746-751 sums it for the FASTA QC)   */
} ac_sketch_info;

/* ---- library / device ------------------------------------------------ */
int ac_abi_version(void);
/* 0 when a CUDA device of compute capability 10.x is visible (replaces the
 * `mash -h` probe of auriclass/general.py:198-205); AC_E_CUDA otherwise. */
int ac_probe(void);
int ac_device_count(void);
/* Message of the last failing call on this thread (never NULL). */
const char *ac_last_error(void);

int ac_create(int device, ac_ctx **out);
void ac_destroy(ac_ctx *ctx);
/* Block until everything queued on the context's streams has finished. */
int ac_sync(ac_ctx *ctx);

/* Run the context's kernels on the caller's CUDA stream (`stream` = a cudaStream_t, `injected` != 0), or go back to
 * the context's own stream (`injected` = 0).  With an injected stream, entry points that only produce DEVICE-side
 * results (ac_job_table_export, ac_merge_counts' arrays, ...) are ordered on that stream instead of blocking the host,
 * so collectives queued on the same stream (NCCL through torch.distributed) need no host synchronisation between them. */
int ac_set_stream(ac_ctx *ctx, void *stream, int injected);

void *ac_alloc_pinned(size_t n);
void ac_free_pinned(void *p);

/* ---- sketching  (mash sketch; classes.py:576-596, 696-713) ----------- */
int ac_sketch_begin(ac_ctx *ctx, int k, int sketch_size, uint32_t seed, int min_copies, int mode,
                    ac_job **out);
/* Append n raw (inflated) FASTA/FASTQ bytes of input file `file_idx` from host
 * memory.  Chunks may be cut anywhere; the library carries record state across
 * calls.  Files must be fed one after the other; `eof` != 0 marks the last
 * chunk of the file (n may be 0).  The copy is asynchronous when `host` is
 * pinned: call ac_sketch_feed_wait before overwriting the buffer. */
int ac_sketch_feed(ac_job *job, const uint8_t *host, size_t n, int file_idx, int eof);
/* Same, for bytes that already live in device memory of the context's GPU
 * (no copy is made; the buffer must stay valid until ac_sketch_end returns). */
int ac_sketch_feed_device(ac_job *job, const uint8_t *dev, size_t n, int file_idx, int eof);
int ac_sketch_feed_wait(ac_job *job);
/* Finish: exact min-copies filter, bottom-s selection, estimates.  Consumes the
 * job (also on error).  AC_E_NO_RECORDS mirrors Mash's empty-input error. */
int ac_sketch_end(ac_job *job, ac_sketch **out, ac_sketch_info *info);
void ac_sketch_abort(ac_job *job);

/* Multi-GPU (read-block sharded) variant of ac_sketch_end: stops after the
 * local pass and leaves the rank's exact (hash, count) table of everything
 * below its threshold on the device.  The caller exchanges table slices by
 * hash range (NCCL all-to-all), and finishes with ac_sketch_from_table. */
int ac_sketch_end_local(ac_job *job, uint64_t *n_entries, uint64_t *tau, int *speculative,
                        ac_sketch_info *info);
/* Entries of the local table with hash <= tau_cut, split at `n_ranges` equal
 * hash ranges of [0, tau_cut]: writes range_begin[0..n_ranges] (entry indices). */
int ac_job_table_ranges(ac_job *job, uint64_t tau_cut, int n_ranges, uint64_t *range_begin);
/* Copy table entries [begin, end) into caller-provided DEVICE buffers. */
int ac_job_table_export(ac_job *job, uint64_t begin, uint64_t end, uint64_t *dev_hashes,
                        uint32_t *dev_counts);
/* Sum counts of equal hashes over n (hash,count) DEVICE entries (any order),
 * keep count >= min_copies, sort: writes up to cap entries back to the DEVICE
 * buffers and the number kept to *n_out. */
int ac_merge_counts(ac_ctx *ctx, uint64_t *dev_hashes, uint32_t *dev_counts, uint64_t n,
                    int min_copies, uint64_t *n_out);
/* Build a sketch from n ascending qualified DEVICE (hash,count) entries (first
 * sketch_size are kept). */
int ac_sketch_from_table(ac_ctx *ctx, const uint64_t *dev_hashes, const uint32_t *dev_counts,
                         uint64_t n, int k, int sketch_size, ac_sketch **out, ac_sketch_info *info);

/* ---- the same exchange as kernels over NVLink peer memory (one process per GPU of ONE box) ----------------
 * Replaces, for the sharded FASTQ path, the single address space in which one `mash sketch -r -m` process counts
 * all reads (auriclass/classes.py:576-596).  Every rank creates a window, publishes its 64-byte CUDA IPC handle
 * (any transport: the host mirror uses torch.distributed's object all-gather, once), opens the others', and from
 * then on an exchange is three calls that only launch kernels and read two small structs back:
 *   ac_sketch_exchange_begin  ends the rank's local pass, posts its threshold to every rank, takes tau_cut = min,
 *                             writes its table entries <= tau_cut into their owners' inboxes (hash-range owner
 *                             r = floor(h P / (tau_cut + 1))) and adds up what arrives in its own inbox
 *   ac_sketch_exchange_merge  sorts the owner's qualified keys, pushes the first <= s to every rank, assembles the
 *                             ordered concatenation's first s entries and runs the estimates
 *   ac_sketch_exchange_end    returns the sketch (identical on every rank) and the job-wide counters
 * The calls are collective in the sense that every rank must make them, in the same order; a rank that failed
 * before (bad input in its blocks) calls _begin with job = NULL, failed = 1 and every rank gets AC_E_PEER (the
 * failing one its own error) -- nobody waits for more than 20 s for a rank that never shows up.
 * inbox_entries: capacity per (sender, owner) pair, 0 = 2 Mi entries; sketch_cap: largest sketch size the window
 * can carry.  ac_xchg_set_peer_local wires windows of the SAME process together (tests emulate ranks that way). */
int ac_xchg_create(ac_ctx *ctx, int rank, int world, int sketch_cap, uint64_t inbox_entries, ac_xchg **out);
int ac_xchg_ipc_handle(ac_xchg *x, void *out64);
int ac_xchg_open_peers(ac_xchg *x, const void *handles /* [world][64], own slot ignored */);
int ac_xchg_set_peer_local(ac_xchg *x, int rank, ac_xchg *other);
void ac_xchg_destroy(ac_xchg *x);
int ac_sketch_exchange_begin(ac_ctx *ctx, ac_job *job, ac_xchg *x, int k, int sketch_size, int min_copies, int failed);
int ac_sketch_exchange_merge(ac_xchg *x);
/* stats (may be NULL): tau_cut, entries received, qualified hashes over all owners, error bits */
int ac_sketch_exchange_end(ac_xchg *x, ac_sketch **out, ac_sketch_info *info, uint64_t stats[4]);

/* Build a device sketch from host hashes (a query .msh, or tests). */
int ac_sketch_from_hashes(ac_ctx *ctx, const uint64_t *hashes, uint64_t n, int k, int sketch_size,
                          uint64_t length, ac_sketch **out);
uint64_t ac_sketch_size(const ac_sketch *sk);
/* Copy the sorted hashes (and exact counts, may be NULL) to host arrays of ac_sketch_size entries. */
int ac_sketch_hashes(const ac_sketch *sk, uint64_t *hashes, uint32_t *counts);
/* Sequence length Mash stores with the sketch (p-value input): FASTA = total_len,
 * reads = (uint64) genome_size. */
uint64_t ac_sketch_length(const ac_sketch *sk);
void ac_sketch_free(ac_sketch *sk);

/* ---- distance  (mash dist; classes.py:92-104) ------------------------ */
/* hashes: concatenation of the ascending hash lists; offsets[n_refs+1]; lengths[n_refs]. */
int ac_refs_load(ac_ctx *ctx, const uint64_t *hashes, const uint64_t *offsets, const uint64_t *lengths,
                 int n_refs, int k, int sketch_size, int hash_bits, ac_refs **out);
int ac_refs_count(const ac_refs *refs);
void ac_refs_free(ac_refs *refs);
/* One query against every reference, in reference order; arrays of n_refs. */
int ac_dist(ac_ctx *ctx, const ac_refs *refs, const ac_sketch *query, double *distance, double *pvalue,
            uint64_t *common, uint64_t *denom);
/* n_q queries against every reference: outputs are [n_q][n_refs] row-major. */
int ac_dist_batch(ac_ctx *ctx, const ac_refs *refs, const ac_sketch *const *queries, int n_q,
                  double *distance, double *pvalue, uint64_t *common, uint64_t *denom);

/* Tile of an all-vs-all `mash dist` (BASELINE config 5; `mash dist A.msh B.msh` with many sketches per file):
 * sketches [i0, i0+ni) of set `a` against sketches [j0, j0+nj) of set `b` (the same handle for a self
 * comparison).  Outputs are [ni][nj] row-major host arrays (any may be NULL).  Mash's row for a pair
 * (reference = a[i], query = b[j]) carries exactly these four numbers. */
int ac_dist_pairs(ac_ctx *ctx, const ac_refs *a, int i0, int ni, const ac_refs *b, int j0, int nj, double *distance,
                  double *pvalue, uint64_t *common, uint64_t *denom);

/* ---- instrumentation ------------------------------------------------- */
/* Kernel launches issued by this context since creation (bench.py gpu_launches). */
uint64_t ac_launch_count(const ac_ctx *ctx);
/* CUDA-event stopwatch on the context's compute stream: ac_timer_start records an event, ac_timer_stop records a
 * second one, waits for it and returns the device time between the two in milliseconds -- everything enqueued on the
 * compute stream in between (a sketch job, the dist call after it; on an injected stream also the caller's
 * collectives).  bench.py times a whole per-isolate step with it (the reference times the `mash` child processes,
 * auriclass/classes.py:100-104, :592-596, from outside in the same way). */
int ac_timer_start(ac_ctx *ctx);
int ac_timer_stop(ac_ctx *ctx, float *ms);
/* Record CUDA events around every kernel group of subsequent jobs (off by default). */
int ac_set_profile(ac_ctx *ctx, int on);
/* Device time (ms, CUDA events on the compute stream) the last finished job
 * spent in: [0] parse (K1), [1] hash (K2), [2] select (K3/K4) -- these three only
 * with ac_set_profile(ctx, 1) -- and [3] first kernel to last kernel of the job. */
int ac_last_job_times(const ac_ctx *ctx, float ms[4]);
/* Measured integer-issue peak: runs an IMAD+LOP3 microbenchmark and returns
 * int32 thread-instructions per second (the sketch roofline denominator). */
int ac_int_peak(ac_ctx *ctx, double *instr_per_s);

/* Issue rates (thread-instructions per second) of the three instruction kinds kernel K2 is made of, each
 * issued alone in independent chains: [0] IMAD, [1] IMAD.WIDE.U32 (both on the fmaheavy pipe), [2] LOP3 (alu
 * pipe).  bench.py turns them into the pipe-cycle roofline of the hash kernel. */
int ac_pipe_rates(ac_ctx *ctx, double rates[3]);

/* Issue rates of the alu-pipe instruction kinds of kernel K2, each issued alone in independent chains:
 * [0] SHF (funnel shift), [1] PRMT, [2] IADD3 (carry pairs), [3] ISETP + SEL pairs, [4] LEA. */
int ac_alu_rates(ac_ctx *ctx, double rates[5]);

#ifdef __cplusplus
}
#endif
#endif /* AURICLASS_CUDA_H */
