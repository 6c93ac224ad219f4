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
constexpr size_t SEG_BYTES = 128u << 20;             // raw text per segment: ~13 waves of K2 CTAs, few dependent launches
constexpr size_t SEG_SAFE_BYTES = 32u << 20;         // segment size when every window may be a candidate (tau = max / safe mode)
#ifndef AC_K2_THREADS
#define AC_K2_THREADS 256
#endif
#ifndef AC_K2_MINB
#define AC_K2_MINB 4
#endif
constexpr int K2_THREADS = AC_K2_THREADS;             // hash CTA
constexpr int K2_W = 32;                             // window starts per hash thread
constexpr int K2_TILE = K2_THREADS * K2_W;           // window starts per hash CTA
constexpr int K2_STAGE = K2_TILE / 8;                 // candidate staging slots per hash CTA (shared memory)

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
    // result (K4)
    unsigned long long max_hash, sum_counts;
    double genome_size, coverage;
};

struct JobCounters;  // (reserved)

__host__ __device__ inline unsigned int div_up(unsigned long long a, unsigned long long b) {
    return (unsigned int)((a + b - 1) / b);
}

// ------------------------------------------------------------ launch glue
struct LaunchCtx {
    cudaStream_t stream;
    uint64_t *launches;  // counter owned by ac_ctx
};

// kernels' host-side launchers (defined in the .cu files)
void launch_parse_segment(const LaunchCtx &lc, const SegDesc &sd, unsigned long long *st1, unsigned long long *st2,
                          uint32_t *tile_arr, size_t tile_cap, JobDev *jd, const uint8_t *Zprev, uint8_t *Z);
// words of the FASTQ per-tile arrays for segments of up to max_tiles tiles; *tile_cap = padded entries per array
size_t parse_tile_array_words(size_t max_tiles, size_t *tile_cap);
// ps_slot: ParseState slot holding this segment's n_zbuf (= (serial + 1) & 1)
void launch_hash_segment(const LaunchCtx &lc, int k, const uint8_t *Z, uint32_t max_zbuf, int ps_slot, JobDev *jd,
                         unsigned long long *cand, unsigned long long cand_cap, uint32_t seed, int use64);

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

constexpr int MAX_RANGES = 64;  // hash ranges (= ranks) of the multi-GPU exchange
struct RangeBounds {             // passed by value
    unsigned long long b[MAX_RANGES];  // b[r] = first hash owned by range r
    unsigned long long tau_cut;
    int n;
};
void launch_table_ranges(const LaunchCtx &lc, const SelectBufs &sb, uint32_t n_table, const RangeBounds &bounds,
                         unsigned long long *d_out /*[n+1], device*/);

void launch_dist(const LaunchCtx &lc, const unsigned long long *ref_hashes, const unsigned long long *ref_off,
                 const unsigned long long *ref_len, int n_refs, const unsigned long long *const *q_hashes,
                 const unsigned long long *q_n, const unsigned long long *q_len, int n_q, int k,
                 unsigned long long sketch_size, unsigned long long max_q_n, double *d_dist, double *d_p,
                 unsigned long long *d_common, unsigned long long *d_denom);

double run_int_peak(const LaunchCtx &lc);


#if defined(__CUDACC__)
constexpr unsigned FULL = 0xffffffffu;
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
