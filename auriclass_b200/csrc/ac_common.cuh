// ac_common.cuh -- structures shared by the kernels and the C-ABI glue of
// libauriclass_cuda.so (see include/auriclass_cuda.h for the boundary).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ac {

// ---------------------------------------------------------------- geometry
constexpr int K1_THREADS = 256;                      // parser CTA
constexpr int K1_CHUNK = 32;                         // raw bytes per parser thread (one 32-bit mask)
constexpr int K1_TILE = K1_THREADS * K1_CHUNK;       // 8 KiB of raw text per parser CTA
constexpr int Z_CARRY = 32;                          // bytes of the base stream carried between segments (>= k-1)
constexpr size_t SEG_BYTES = 256u << 20;             // raw text per segment: few dependent launches (measured: 64 MiB 5.13, 128 MiB 4.85, 256 MiB 4.75, 512 MiB 4.78 ms per C2 isolate)
constexpr size_t SEG_SAFE_BYTES = 32u << 20;         // segment size when every window may be a candidate (tau = max / safe mode)
#ifndef AC_K2_THREADS
#define AC_K2_THREADS 128
#endif
#ifndef AC_K2_MINB
#define AC_K2_MINB 8
#endif
constexpr int K2_THREADS = AC_K2_THREADS;             // hash CTA: independent warps, no CTA barrier after the first
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr int K2_W = 32;                             // window starts per lane and round
#ifndef AC_K2_WT
#define AC_K2_WT 2048
#endif
constexpr int K2_WT = AC_K2_WT;                       // window starts per hash WARP (its own sub-tile: 64 rounds of 32)
constexpr int K2_TILE = K2_WARPS * K2_WT;            // window starts per hash CTA
constexpr int K2_STAGE = 128;                        // candidate staging slots per hash warp (shared memory)
constexpr int MAX_RANGES = 64;                       // hash ranges (= ranks) of the multi-GPU exchange

enum { FMT_FASTQ = 0, FMT_FASTA = 1 };
enum {  // JobDev::format_error bits
    FERR_FQ_HEADER = 1,   // a FASTQ record does not start with '@'
    FERR_FQ_PLUS = 2,     // third line of a FASTQ record does not start with '+'
    FERR_FQ_TRUNC = 4,    // file ends inside a record / sequence and quality lengths differ
    FERR_FA_PLUS = 8      // '+' line inside FASTA input (kseq would switch to quality mode)
};

// One contiguous piece of raw text handed to K1 (passed by value).
struct SegDesc {
    const uint8_t *base;  // 16-byte aligned (aligned-down) device pointer
    uint32_t skip;        // leading bytes of the aligned range that precede the segment
    uint32_t len;         // segment bytes incl. the virtual newline when append_nl
    uint32_t n_tiles;     // >= 1
    int32_t append_nl;    // the last byte of the segment is a virtual '\n' (file lacks a final newline)
    int32_t fmt;          // FMT_*
    int32_t first_of_stream;  // first segment of the job's base stream: carry is all separators
    int32_t last_of_file;
    uint32_t serial;      // segment number within the pass: parity selects the ParseState slot and the Z buffer
    uint32_t epoch;       // context-wide segment serial (< 2^30, never reused): tags the look-back status words
    uint32_t ticket_base; // tiles handed out by earlier segments of the pass (JobDev::tile_ticket keeps counting)
};

// Parser state between segments; double buffered (segment i reads slot i&1, writes slot (i+1)&1)
struct ParseState {
    uint32_t fq_phase;       // FASTQ: phase (0 header,1 sequence,2 plus,3 quality) of the line holding the next byte
    uint32_t fa_type;        // FASTA: type of that line (0 header, 1 sequence)
    uint32_t at_line_start;  // the next byte begins a line
    uint32_t n_zbuf;         // end of the segment's base stream in its Z buffer
    uint32_t z_begin;        // its first byte (the carry of the previous segment occupies the Z_CARRY bytes before it)
};

// Device-resident job state, updated by kernels, read back by the host at sync points.
struct JobDev {
    ParseState ps[2];
    uint32_t tile_ticket;   // K1 tile dispenser (monotonic over the pass)
    uint32_t format_error;
    uint32_t overflow;  // candidate buffer overflowed (pass must be repeated in safe mode)
    // counters over the whole job
    unsigned long long fq_seq_bytes, fq_qual_bytes;
    unsigned long long n_records, n_kept, n_short_len, n_bases, n_kmers;
    // selection state
    unsigned long long tau;         // keep hash <= tau
    unsigned long long cand_count;  // entries in the candidate buffer
    unsigned long long cand_total;  // candidates emitted over the pass
    uint32_t n_table;               // entries of the (hash,count) table
    uint32_t n_unique;              // scratch: distinct keys of the last consolidation
    uint32_t n_qualified;           // entries of the table with count >= m (capped view: see select)
    uint32_t n_sketch;              // min(s, n_qualified)
    unsigned long long total_vals;  // scratch
    uint32_t n_qlist;               // count-table mode: keys whose count reached min_copies so far (entries of the qualified list)
    uint32_t n_out;                 // count-table mode: entries written by the collect / export kernels
    uint32_t range_cnt[MAX_RANGES];         // multi-GPU exchange: table entries <= tau_cut per hash range (MAX_RANGES), then write cursors
    // result (K4)
    unsigned long long max_hash, sum_counts;
    double genome_size, coverage;
};

// Count table of the speculative pass (tau0 < max): an open-addressing hash table keyed by the hash itself
// (hashes <= tau only, so the all-ones key is free to mean "empty"), exact occurrence counts, the list of keys
// whose count reached min_copies, and a 65,536-bucket histogram of those keys by their top bits below tau0.
// K2 inserts its candidates straight into it; the threshold follows from the histogram (k3t_tau_update) and
// the sketch from the qualified list -- the candidate stream is never sorted.
constexpr unsigned long long CT_EMPTY = ~0ULL;
constexpr int CT_BUCKET_BITS = 16;
constexpr int CT_MAX_PROBE = 512;
struct CountTable {              // passed by value to the kernels
    unsigned long long *key;     // [mask + 1]
    uint32_t *cnt;               // [mask + 1]
    uint32_t *qhist;             // [1 << CT_BUCKET_BITS]
    unsigned long long *qlist;   // [qcap]
    uint32_t mask;               // slots - 1 (power of two); 0 = table mode off (candidates go to the sort buffer)
    uint32_t qcap;
    uint32_t min_copies;
    int bshift;                  // bucket of a key = key >> bshift
};

__host__ __device__ inline unsigned int div_up(unsigned long long a, unsigned long long b) {
    return (unsigned int)((a + b - 1) / b);
}

// ------------------------------------------------------------ launch glue
struct LaunchCtx {
    cudaStream_t stream;
    uint64_t *launches;  // counter owned by ac_ctx
};

// kernels' host-side launchers (defined in the .cu files)
// K1 over one segment: FASTQ -> k1q_fused (st1 = look-back status words, cleared here), FASTA -> k1_fasta (st1 / st2)
void launch_parse_segment(const LaunchCtx &lc, const SegDesc &sd, unsigned long long *st1, unsigned long long *st2,
                          JobDev *jd, const uint8_t *Zprev, uint8_t *Z);
// ps_slot: ParseState slot holding this segment's n_zbuf (= (serial + 1) & 1)
// ct.mask != 0: candidates are inserted into the count table instead of being appended to cand
// false: this build has no kernel for k (nothing launched)
bool launch_hash_segment(const LaunchCtx &lc, int k, const uint8_t *Z, uint32_t max_zbuf, int ps_slot, JobDev *jd,
                         unsigned long long *cand, unsigned long long cand_cap, uint32_t seed, int use64,
                         const CountTable &ct, uint32_t cta0 = 0, uint32_t cta1 = 0xffffffffu);   // CTAs [cta0, cta1) of the segment's grid
// CTAs of launch_hash_segment's grid that cover the first `zbytes` bytes of the base stream
uint32_t hash_ctas_for(uint32_t zbytes);

struct SelectBufs {
    unsigned long long *keys_a, *keys_b;  // candidate buffer (keys_a) and its ping-pong partner
    uint32_t *vals_a, *vals_b;
    unsigned long long *tab_hash;         // table (sorted unique hashes <= tau)
    uint32_t *tab_cnt;
    uint32_t *start_sum;                  // scratch: exclusive value prefix per unique key
    uint32_t *hist;                       // radix histograms [256][n_tiles]
    uint32_t *tile_tmp;                   // scan tile sums (2 x n_tiles)
    unsigned long long *sk_hash;          // sketch under construction (s entries)
    uint32_t *sk_cnt;
    size_t cap;                           // entries in keys/vals
    size_t tab_cap;
};
// Fold n_cand candidates (keys_a[0..n_cand), implicit count 1) and the current table into a new
// table; lower tau when >= s qualified entries exist; refresh the sketch prefix.
void launch_consolidate(const LaunchCtx &lc, const SelectBufs &sb, JobDev *jd, uint64_t n_cand,
                        uint32_t n_table, int key_bits, int sketch_size, int min_copies);
void launch_finalize(const LaunchCtx &lc, const SelectBufs &sb, JobDev *jd, int use64);
// sort (hash,count) pairs, sum equal hashes, keep count >= m; result in keys_a/vals_a, n in jd->n_unique
void launch_merge_counts(const LaunchCtx &lc, const SelectBufs &sb, JobDev *jd, uint64_t n, int min_copies);

// ---- count-table mode (ac_select.cu)
// tau <- upper edge of the histogram bucket holding the s-th smallest qualified key (if there are s);
// jd->n_qualified <- qualified keys <= tau (all of them when fewer than s exist)
void launch_table_tau_update(const LaunchCtx &lc, const CountTable &ct, JobDev *jd, int sketch_size);
// qualified keys <= tau with their counts -> (keys_a, vals_a), n in jd->n_out
void launch_table_collect(const LaunchCtx &lc, const CountTable &ct, const SelectBufs &sb, JobDev *jd, uint32_t n_qlist);
// every key <= tau with its count -> (keys_a, vals_a), n in jd->n_out  (multi-GPU exchange needs unqualified ones too)
void launch_table_export(const LaunchCtx &lc, const CountTable &ct, const SelectBufs &sb, JobDev *jd);
// sort n (hash,count) pairs of (keys_a, vals_a); all n -> (tab_hash, tab_cnt); first min(n_q, s) -> sketch unless
// all_entries (the pairs then include unqualified keys); sets n_sketch / n_qualified (= n_q) / n_table (= n)
void launch_table_finish(const LaunchCtx &lc, const SelectBufs &sb, JobDev *jd, uint32_t n, uint32_t n_q, int key_bits,
                         int sketch_size, int all_entries);

struct RangeBounds {             // passed by value
    unsigned long long b[MAX_RANGES];  // b[r] = first hash owned by range r
    unsigned long long tau_cut;
    int n;
};
void launch_table_ranges(const LaunchCtx &lc, const SelectBufs &sb, uint32_t n_table, const RangeBounds &bounds,
                         unsigned long long *d_out /*[n+1], device*/);
// count-table mode, no sort: entries <= bounds.tau_cut counted per hash range into jd->range_cnt ...
void launch_table_range_count(const LaunchCtx &lc, const CountTable &ct, const RangeBounds &bounds, JobDev *jd);
// ... and, with the ranges' exclusive offsets, written range by range (any order inside a range) to (tab_hash, tab_cnt)
struct RangeOffsets { uint32_t o[MAX_RANGES + 1]; };
void launch_table_range_scatter(const LaunchCtx &lc, const CountTable &ct, const RangeBounds &bounds, const RangeOffsets &off,
                                const SelectBufs &sb, JobDev *jd);
// owner merge through the count table: (hash, count) DEVICE pairs in any order are added up; every hash whose sum
// reaches min_copies lands in ct.qlist (jd->n_qlist); jd->overflow when the table cannot hold them
void launch_merge_insert(const LaunchCtx &lc, const CountTable &ct, const unsigned long long *hashes, const uint32_t *counts,
                         uint32_t n, JobDev *jd);

void launch_dist(const LaunchCtx &lc, const unsigned long long *ref_hashes, const unsigned long long *ref_off,
                 const unsigned long long *ref_len, int n_refs, const unsigned long long *const *q_hashes,
                 const unsigned long long *q_n, const unsigned long long *q_len, int n_q, int k,
                 unsigned long long sketch_size, unsigned long long max_q_n, double *d_dist, double *d_p,
                 unsigned long long *d_common, unsigned long long *d_denom, void *parts_scratch);
// batches of up to this many pairs are split over several CTAs per pair; parts_scratch: DIST_SPLIT_MAX_PAIRS * 64 bytes
constexpr int DIST_SPLIT_MAX_PAIRS = 256;

// tile of pairs: sketches [i0, i0+ni) of set A against [j0, j0+nj) of set B, outputs [ni][nj] row-major
void launch_dist_pairs(const LaunchCtx &lc, const unsigned long long *a_hashes, const unsigned long long *a_off,
                       const unsigned long long *a_len, uint32_t i0, uint32_t ni, const unsigned long long *b_hashes,
                       const unsigned long long *b_off, const unsigned long long *b_len, uint32_t j0, uint32_t nj, int k,
                       unsigned long long sketch_size, double *d_dist, double *d_p, unsigned long long *d_common,
                       unsigned long long *d_denom);

double run_int_peak(const LaunchCtx &lc);
void run_pipe_rates(const LaunchCtx &lc, double rates[3]);
void run_alu_rates(const LaunchCtx &lc, double rates[5]);


#if defined(__CUDACC__)
constexpr unsigned FULL = 0xffffffffu;

// One occurrence of hash h (<= tau < CT_EMPTY) into the count table.  Linear probing from the low bits of h
// (MurmurHash3 output: uniform).  A slot's key changes once, from empty to its owner, so an uncached read
// that shows a foreign key is final and one that shows "empty" is settled by the CAS.
__device__ __forceinline__ void ct_insert(const CountTable &ct, unsigned long long h, JobDev *jd) {
    uint32_t slot = (uint32_t)h & ct.mask;
    for (int probe = 0; probe < CT_MAX_PROBE; probe++, slot = (slot + 1u) & ct.mask) {
        unsigned long long cur = __ldcg(ct.key + slot);
        if (cur == CT_EMPTY) cur = atomicCAS(ct.key + slot, CT_EMPTY, h);
        if (cur == CT_EMPTY || cur == h) {
            const uint32_t c = atomicAdd(ct.cnt + slot, 1u) + 1u;
            if (c == ct.min_copies) {   // the key qualifies now (exactly once per key)
                atomicAdd(ct.qhist + (uint32_t)(h >> ct.bshift), 1u);
                const uint32_t i = atomicAdd(&jd->n_qlist, 1u);
                if (i < ct.qcap) ct.qlist[i] = h; else jd->overflow = 1u;
            }
            return;
        }
    }
    jd->overflow = 1u;   // table (nearly) full: the pass is repeated through the sort path
}
// count of a key that is known to be in the table
__device__ __forceinline__ uint32_t ct_lookup(const CountTable &ct, unsigned long long h) {
    uint32_t slot = (uint32_t)h & ct.mask;
    for (int probe = 0; probe < CT_MAX_PROBE; probe++, slot = (slot + 1u) & ct.mask)
        if (__ldcg(ct.key + slot) == h) return __ldcg(ct.cnt + slot);
    return 0u;
}

// exclusive sum scan over the CTA (NT threads, NT/32 <= 32 warps); total returned to all
template <int NT>
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t *s_w /*[33]*/, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t x = (lane < NT / 32) ? s_w[lane] : 0u, xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, xi, d);
            if (lane >= d) xi += t;
        }
        s_w[lane] = xi - x;
        if (lane == 31) s_w[32] = xi;
    }
    __syncthreads();
    const uint32_t r = inc - v + s_w[warp];
    total = s_w[32];
    __syncthreads();
    return r;
}
// inclusive max scan (used for "last defined" propagation)
template <int NT>
__device__ __forceinline__ uint32_t block_incl_max(uint32_t v, uint32_t *s_w /*[33]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc = max(inc, t);
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t x = (lane < NT / 32) ? s_w[lane] : 0u, xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, xi, d);
            if (lane >= d) xi = max(xi, t);
        }
        // exclusive across warps
        uint32_t ex = __shfl_up_sync(FULL, xi, 1);
        s_w[lane] = lane ? ex : 0u;
    }
    __syncthreads();
    const uint32_t r = max(inc, s_w[warp]);
    __syncthreads();
    return r;
}

#endif  // __CUDACC__

}  // namespace ac
