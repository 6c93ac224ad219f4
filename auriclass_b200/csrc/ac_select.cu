// ac_select.cu -- kernels K3 (exact min-copies count filter + bottom-s under a
// monotonically shrinking threshold) and K4 (genome-size / coverage estimate).
//
// Replaces Mash's MinHashHeap::tryInsert / hashesPending bookkeeping and
// estimateSetSize / estimateMultiplicity (SURVEY.md section 2.2 stages S4-S5,
// App. A.4/A.5) that run inside `mash sketch -r -m M`
// (auriclass/classes.py:576-596).
//
// State: a table of every distinct hash <= tau seen so far with its exact
// occurrence count, sorted ascending.  A consolidation folds the candidate buffer
// (hash occurrences <= tau emitted by K2 since the last one) into the table:
//   LSD radix sort of (hash,count) pairs -> run-length sums -> count >= m scan
//   -> if >= s qualified entries exist, tau drops to the s-th of them and the
//   table is truncated there.
// Entries above the s-th qualified hash can never enter the sketch (App. A.4), so
// tau only ever shrinks and counts of everything <= tau stay exact.
#include <stddef.h>

#include "ac_common.cuh"

namespace ac {

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per CTA

typedef unsigned long long u64;

// ------------------------------------------------------------- radix sort
__global__ void __launch_bounds__(RS_THREADS) rs_hist(const u64 *__restrict__ keys, uint32_t n, int shift,
                                                      uint32_t *__restrict__ hist, uint32_t n_tiles) {
    __shared__ uint32_t h[256];
    const uint32_t tid = threadIdx.x, tile = blockIdx.x;
    h[tid] = 0;
    __syncthreads();
    const uint32_t base = tile * RS_TILE;
#pragma unroll 4
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t idx = base + i * RS_THREADS + tid;
        if (idx < n) atomicAdd(&h[(uint32_t)(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[tid * n_tiles + tile] = h[tid];
}

// one CTA per digit: exclusive scan of the digit's row over tiles, row total -> totals[digit]
__global__ void __launch_bounds__(RS_THREADS) rs_scan_rows(uint32_t *__restrict__ hist, uint32_t n_tiles,
                                                           uint32_t *__restrict__ totals) {
    __shared__ uint32_t s_w[33];
    uint32_t *row = hist + (size_t)blockIdx.x * n_tiles;
    uint32_t carry = 0;
    for (uint32_t b = 0; b < n_tiles; b += RS_THREADS) {
        const uint32_t i = b + threadIdx.x;
        const uint32_t v = (i < n_tiles) ? row[i] : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_sum<RS_THREADS>(v, s_w, tot);
        if (i < n_tiles) row[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const u64 *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                         u64 *__restrict__ kout, uint32_t *__restrict__ vout,
                                                         uint32_t n, int shift, const uint32_t *__restrict__ hist,
                                                         const uint32_t *__restrict__ totals, uint32_t n_tiles) {
    __shared__ uint32_t s_cnt[RS_THREADS / 32][256];
    __shared__ uint32_t s_gofs[256];
    __shared__ uint32_t s_w[33];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, tile = blockIdx.x;
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; w++) s_cnt[w][tid] = 0;
    uint32_t tot;
    const uint32_t dbase = block_excl_sum<RS_THREADS>(totals[tid], s_w, tot);  // syncs
    s_gofs[tid] = dbase + hist[tid * n_tiles + tile];
    __syncthreads();

    const uint32_t base = tile * RS_TILE + warp * (32 * RS_ITEMS);
    u64 key[RS_ITEMS];
    uint32_t rk[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t idx = base + i * 32 + lane;
        const bool valid = idx < n;
        key[i] = valid ? kin[idx] : 0ULL;
        const uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & 255u) : (256u + lane);
        const uint32_t peers = __match_any_sync(FULL, d);
        const uint32_t r = __popc(peers & lt);
        const uint32_t prev = valid ? s_cnt[warp][d] : 0u;
        __syncwarp();
        if (valid && r == 0) s_cnt[warp][d] = prev + __popc(peers);
        __syncwarp();
        rk[i] = prev + r;
    }
    __syncthreads();
    {   // exclusive prefix of the per-warp digit counts across warps (thread = digit)
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; w++) {
            const uint32_t c = s_cnt[w][tid];
            s_cnt[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t idx = base + i * 32 + lane;
        if (idx < n) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
            const uint32_t pos = s_gofs[d] + s_cnt[warp][d] + rk[i];
            kout[pos] = key[i];
            vout[pos] = vin[idx];
        }
    }
}

// sorts n (key,val) pairs by the low key_bits of the key; returns 0 if the result is in (ka,va), 1 if in (kb,vb)
int radix_sort_pairs(const LaunchCtx &lc, u64 *ka, uint32_t *va, u64 *kb, uint32_t *vb, uint32_t n, int key_bits,
                     uint32_t *hist, uint32_t *totals) {
    if (n == 0) return 0;
    const uint32_t n_tiles = div_up(n, RS_TILE);
    const int passes = (key_bits + 7) / 8;
    int cur = 0;
    for (int p = 0; p < passes; p++) {
        const u64 *kin = cur ? kb : ka; const uint32_t *vin = cur ? vb : va;
        u64 *kout = cur ? ka : kb; uint32_t *vout = cur ? va : vb;
        rs_hist<<<n_tiles, RS_THREADS, 0, lc.stream>>>(kin, n, 8 * p, hist, n_tiles);
        rs_scan_rows<<<256, RS_THREADS, 0, lc.stream>>>(hist, n_tiles, totals);
        rs_scatter<<<n_tiles, RS_THREADS, 0, lc.stream>>>(kin, vin, kout, vout, n, 8 * p, hist, totals, n_tiles);
        *lc.launches += 3;
        cur ^= 1;
    }
    return cur;
}

// ------------------------------------------------------- small sort of hash values (bucket sort)
// The <= ~s pairs a count-table pass ends with are MurmurHash3 values spread evenly over [0, tau]: bucketing by
// the top 16 bits below tau's top bit leaves about one key per bucket, so hist -> scan -> scatter -> a per-bucket
// fix-up orders them in 4 launches where the 8-bit LSD radix sort needs 24.  Exact for any input (a crowded
// bucket is only slow: its fix-up is one thread's heap sort).
constexpr int BS_BITS = 16;
constexpr int BS_BUCKETS = 1 << BS_BITS;

__global__ void __launch_bounds__(256) bs_hist(const u64 *__restrict__ keys, uint32_t n, int shift, uint32_t *__restrict__ hist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(hist + (uint32_t)(keys[i] >> shift), 1u);
}

// one CTA: exclusive scan of the bucket counts -> start[], cursors zeroed.  Thread t owns buckets [64 t, 64 t + 64)
// (16-byte accesses; the counts are read twice rather than kept in 64 registers, which would spill at 1,024 threads)
__global__ void __launch_bounds__(1024) bs_scan(const uint32_t *__restrict__ hist, uint32_t *__restrict__ start,
                                                uint32_t *__restrict__ cursor) {
    __shared__ uint32_t s_w[33];
    constexpr int PER = BS_BUCKETS / 1024;
    const uint32_t t = threadIdx.x;
    const uint4 *h4 = reinterpret_cast<const uint4 *>(hist + (size_t)t * PER);
    uint32_t mine = 0;
#pragma unroll 4
    for (int i = 0; i < PER / 4; i++) { const uint4 v = h4[i]; mine += v.x + v.y + v.z + v.w; }
    uint32_t total;
    uint32_t run = block_excl_sum<1024>(mine, s_w, total);
    uint4 *s4 = reinterpret_cast<uint4 *>(start + (size_t)t * PER);
    uint4 *c4 = reinterpret_cast<uint4 *>(cursor + (size_t)t * PER);
#pragma unroll 4
    for (int i = 0; i < PER / 4; i++) {
        const uint4 v = h4[i];
        uint4 o;
        o.x = run; run += v.x; o.y = run; run += v.y; o.z = run; run += v.z; o.w = run; run += v.w;
        s4[i] = o;
        c4[i] = make_uint4(0, 0, 0, 0);
    }
    if (t == 1023) start[BS_BUCKETS] = run;
}

__global__ void __launch_bounds__(256) bs_scatter(const u64 *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n,
                                                  int shift, const uint32_t *__restrict__ start, uint32_t *__restrict__ cursor,
                                                  u64 *__restrict__ okeys, uint32_t *__restrict__ ovals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 k = keys[i];
    const uint32_t b = (uint32_t)(k >> shift);
    const uint32_t pos = start[b] + atomicAdd(cursor + b, 1u);
    okeys[pos] = k; ovals[pos] = vals[i];
}

__device__ void bs_sift_down(u64 *k, uint32_t *v, uint32_t root, uint32_t end) {
    for (;;) {
        uint32_t child = 2 * root + 1;
        if (child >= end) return;
        if (child + 1 < end && k[child] < k[child + 1]) child++;
        if (k[root] >= k[child]) return;
        const u64 tk = k[root]; k[root] = k[child]; k[child] = tk;
        const uint32_t tv = v[root]; v[root] = v[child]; v[child] = tv;
        root = child;
    }
}

// one thread per bucket: order its (few) entries
__global__ void __launch_bounds__(256) bs_fix(const uint32_t *__restrict__ start, u64 *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= BS_BUCKETS) return;
    const uint32_t lo = start[b], n = start[b + 1] - lo;
    if (n < 2) return;
    u64 *k = keys + lo;
    uint32_t *v = vals + lo;
    if (n <= 32) {
        for (uint32_t i = 1; i < n; i++) {
            const u64 kk = k[i];
            const uint32_t vv = v[i];
            uint32_t jx = i;
            while (jx > 0 && k[jx - 1] > kk) { k[jx] = k[jx - 1]; v[jx] = v[jx - 1]; jx--; }
            k[jx] = kk; v[jx] = vv;
        }
        return;
    }
    for (uint32_t i = n / 2; i-- > 0;) bs_sift_down(k, v, i, n);          // crowded bucket: in-place heap sort
    for (uint32_t end = n - 1; end > 0; end--) {
        const u64 tk = k[0]; k[0] = k[end]; k[end] = tk;
        const uint32_t tv = v[0]; v[0] = v[end]; v[end] = tv;
        bs_sift_down(k, v, 0, end);
    }
}

// sorts n pairs (keys <= 2^key_bits - 1) from (ka, va) into (kb, vb); hist scratch: 3 * BS_BUCKETS + 4 words
void bucket_sort_pairs(const LaunchCtx &lc, const u64 *ka, const uint32_t *va, u64 *kb, uint32_t *vb, uint32_t n, int key_bits,
                       uint32_t *scratch) {
    const int shift = key_bits > BS_BITS ? key_bits - BS_BITS : 0;
    uint32_t *hist = scratch, *start = scratch + BS_BUCKETS, *cursor = scratch + 2 * BS_BUCKETS + 4;   // 16-byte aligned rows
    cudaMemsetAsync(hist, 0, BS_BUCKETS * sizeof(uint32_t), lc.stream);
    bs_hist<<<div_up(n, 256), 256, 0, lc.stream>>>(ka, n, shift, hist);
    bs_scan<<<1, 1024, 0, lc.stream>>>(hist, start, cursor);
    bs_scatter<<<div_up(n, 256), 256, 0, lc.stream>>>(ka, va, n, shift, start, cursor, kb, vb);
    bs_fix<<<BS_BUCKETS / 256, 256, 0, lc.stream>>>(start, kb, vb);
    *lc.launches += 4;
}

// ------------------------------------------------------- consolidation
// vals of the fresh candidates are an implicit 1; table entries are appended behind them
__global__ void k3_prepare(u64 *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t n_cand,
                           const u64 *__restrict__ tab_hash, const uint32_t *__restrict__ tab_cnt, uint32_t n_table) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_cand) vals[i] = 1u;
    else if (i < n_cand + n_table) { keys[i] = tab_hash[i - n_cand]; vals[i] = tab_cnt[i - n_cand]; }
}

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;  // 2048 elements per CTA, thread owns 8 consecutive

// pass 1 over the sorted pairs: per tile (#run heads, sum of vals)
__global__ void __launch_bounds__(SC_THREADS) k3_runs_reduce(const u64 *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                             uint32_t n, uint32_t *__restrict__ tile_heads,
                                                             uint32_t *__restrict__ tile_vals) {
    __shared__ uint32_t s_w[33];
    const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    uint32_t heads = 0, vs = 0;
    u64 prev = (base > 0 && base - 1 < n) ? keys[base - 1] : 0ULL;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t idx = base + i;
        if (idx < n) {
            const u64 k = keys[idx];
            heads += (idx == 0 || k != prev) ? 1u : 0u;
            vs += vals[idx];
            prev = k;
        }
    }
    uint32_t th, tv;
    block_excl_sum<SC_THREADS>(heads, s_w, th);
    block_excl_sum<SC_THREADS>(vs, s_w, tv);
    if (threadIdx.x == 0) { tile_heads[blockIdx.x] = th; tile_vals[blockIdx.x] = tv; }
}

// single CTA: exclusive scan of the two tile arrays in place; totals -> jd (mode 0) or qualified count (mode 1)
__global__ void __launch_bounds__(1024) k3_scan_tiles(uint32_t *__restrict__ a, uint32_t *__restrict__ b, uint32_t n_tiles,
                                                      JobDev *__restrict__ jd, uint32_t *__restrict__ start_sum, int mode,
                                                      uint32_t sketch_size) {
    __shared__ uint32_t s_w[33];
    uint32_t ca = 0, cb = 0;
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += 1024) {
        const uint32_t i = t0 + threadIdx.x;
        const uint32_t va = (i < n_tiles) ? a[i] : 0u;
        const uint32_t vb = (b && i < n_tiles) ? b[i] : 0u;
        uint32_t ta, tb;
        const uint32_t ea = block_excl_sum<1024>(va, s_w, ta);
        const uint32_t eb = block_excl_sum<1024>(vb, s_w, tb);
        if (i < n_tiles) { a[i] = ca + ea; if (b) b[i] = cb + eb; }
        ca += ta; cb += tb;
    }
    if (threadIdx.x == 0) {
        if (mode == 0) {
            jd->n_unique = ca;
            jd->total_vals = cb;
            start_sum[ca] = cb;  // sentinel: exclusive prefix one past the last unique key
        } else {
            jd->n_qualified = ca;
            jd->n_sketch = ca < sketch_size ? ca : sketch_size;
            jd->n_table = jd->n_unique;  // default: keep everything (k3_qual_apply lowers it with tau)
            jd->cand_count = 0;
        }
    }
}

// pass 2: unique keys -> tab_hash, exclusive val prefix at each run head -> start_sum
__global__ void __launch_bounds__(SC_THREADS) k3_runs_apply(const u64 *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                            uint32_t n, const uint32_t *__restrict__ tile_heads,
                                                            const uint32_t *__restrict__ tile_vals, u64 *__restrict__ tab_hash,
                                                            uint32_t *__restrict__ start_sum) {
    __shared__ uint32_t s_w[33];
    const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    u64 k[SC_ITEMS];
    uint32_t v[SC_ITEMS];
    uint32_t heads = 0, vs = 0;
    const u64 before = (base > 0 && base - 1 < n) ? keys[base - 1] : 0ULL;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t idx = base + i;
        k[i] = (idx < n) ? keys[idx] : 0ULL;
        v[i] = (idx < n) ? vals[idx] : 0u;
    }
    u64 prev = before;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t idx = base + i;
        if (idx < n) { heads += (idx == 0 || k[i] != prev) ? 1u : 0u; vs += v[i]; prev = k[i]; }
    }
    uint32_t th, tv;
    uint32_t eh = block_excl_sum<SC_THREADS>(heads, s_w, th) + tile_heads[blockIdx.x];
    uint32_t ev = block_excl_sum<SC_THREADS>(vs, s_w, tv) + tile_vals[blockIdx.x];
    prev = before;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t idx = base + i;
        if (idx < n) {
            if (idx == 0 || k[i] != prev) { tab_hash[eh] = k[i]; start_sum[eh] = ev; eh++; }
            ev += v[i];
            prev = k[i];
        }
    }
}

// pass 3 over unique keys: qualified (count >= m) count per tile
__global__ void __launch_bounds__(SC_THREADS) k3_qual_reduce(const uint32_t *__restrict__ start_sum,
                                                             const JobDev *__restrict__ jd, uint32_t min_copies,
                                                             uint32_t *__restrict__ tile_q) {
    __shared__ uint32_t s_w[33];
    const uint32_t n = jd->n_unique;
    const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    uint32_t q = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t u = base + i;
        if (u < n) q += (start_sum[u + 1] - start_sum[u] >= min_copies) ? 1u : 0u;
    }
    uint32_t tq;
    block_excl_sum<SC_THREADS>(q, s_w, tq);
    if (threadIdx.x == 0) tile_q[blockIdx.x] = tq;
}

// pass 4: counts -> tab_cnt, qualified prefix -> sketch, s-th qualified -> tau + table truncation.
// compact != 0 (multi-GPU owner merge): write ALL qualified entries to (out_hash,out_cnt) instead.
__global__ void __launch_bounds__(SC_THREADS) k3_qual_apply(const u64 *__restrict__ tab_hash, const uint32_t *__restrict__ start_sum,
                                                            uint32_t *__restrict__ tab_cnt, JobDev *__restrict__ jd,
                                                            uint32_t min_copies, uint32_t sketch_size,
                                                            const uint32_t *__restrict__ tile_q, u64 *__restrict__ out_hash,
                                                            uint32_t *__restrict__ out_cnt, int compact) {
    __shared__ uint32_t s_w[33];
    const uint32_t n = jd->n_unique;
    const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    uint32_t cnt[SC_ITEMS];
    uint32_t q = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t u = base + i;
        cnt[i] = (u < n) ? (start_sum[u + 1] - start_sum[u]) : 0u;
        q += (u < n && cnt[i] >= min_copies) ? 1u : 0u;
    }
    uint32_t tq;
    uint32_t rank = block_excl_sum<SC_THREADS>(q, s_w, tq) + tile_q[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t u = base + i;
        if (u < n) {
            if (!compact) tab_cnt[u] = cnt[i];
            if (cnt[i] >= min_copies) {
                if (compact || rank < sketch_size) { out_hash[rank] = tab_hash[u]; out_cnt[rank] = cnt[i]; }
                if (!compact && rank + 1 == sketch_size) { jd->tau = tab_hash[u]; jd->n_table = u + 1; }
                rank++;
            }
        }
    }
}

// K4: estimates from the finished sketch (Mash estimateSetSize / multiplicity; App. A.5)
__global__ void __launch_bounds__(1024) k4_finalize(const u64 *__restrict__ sk_hash, const uint32_t *__restrict__ sk_cnt,
                                                    JobDev *__restrict__ jd, int use64) {
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const uint32_t n = jd->n_sketch;
    unsigned long long local = 0;
    for (uint32_t i = threadIdx.x; i < n; i += 1024) local += sk_cnt[i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) local += __shfl_xor_sync(FULL, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_sum, local);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (n) {
            const u64 mx = sk_hash[n - 1];
            const double space = use64 ? 18446744073709551616.0 : 4294967296.0;
            jd->max_hash = mx;
            jd->sum_counts = s_sum;
            jd->genome_size = space * (double)n / (double)mx;
            jd->coverage = (double)s_sum / (double)n;
        } else {
            jd->max_hash = 0; jd->sum_counts = 0; jd->genome_size = 0.0; jd->coverage = 0.0;
        }
    }
}

// multi-GPU hash-range split: out[r] = first table entry >= bounds.b[r] (r < n), out[n] = first entry > tau_cut
__global__ void k3_table_ranges(const u64 *__restrict__ tab_hash, uint32_t n_table, RangeBounds bounds,
                                unsigned long long *__restrict__ out) {
    const int r = threadIdx.x;
    if (r > bounds.n) return;
    const bool upper = r == bounds.n;          // last boundary: upper_bound(tau_cut)
    const u64 key = upper ? bounds.tau_cut : bounds.b[r];
    uint32_t lo = 0, hi = n_table;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const u64 v = tab_hash[mid];
        if (upper ? (v <= key) : (v < key)) lo = mid + 1; else hi = mid;
    }
    out[r] = lo;
}


// ------------------------------------------------------- count-table mode
// Threshold from the histogram of qualified keys: one CTA, thread t owns buckets [64t, 64t + 64).
__global__ void __launch_bounds__(1024) k3t_tau_update(const uint32_t *__restrict__ qhist, JobDev *__restrict__ jd,
                                                       uint32_t sketch_size, int bshift) {
    __shared__ uint32_t s_w[33];
    constexpr int PER = (1 << CT_BUCKET_BITS) / 1024;
    const uint32_t t = threadIdx.x;
    uint32_t mine = 0;
    const uint4 *row = reinterpret_cast<const uint4 *>(qhist + (size_t)t * PER);
#pragma unroll
    for (int i = 0; i < PER / 4; i++) { const uint4 v = __ldcg(row + i); mine += v.x + v.y + v.z + v.w; }
    uint32_t total;
    const uint32_t before = block_excl_sum<1024>(mine, s_w, total);
    if (t == 0) {
        jd->n_out = 0;
        if (total < sketch_size) { jd->n_qualified = total; jd->n_sketch = total; }
    }
    if (total >= sketch_size && before < sketch_size && before + mine >= sketch_size) {
        uint32_t cum = before;
        for (int i = 0; i < PER; i++) {
            cum += __ldcg(qhist + (size_t)t * PER + i);
            if (cum >= sketch_size) {
                const u64 edge = (((u64)t * PER + i + 1) << bshift) - 1ULL;   // last key of the bucket
                if (edge < jd->tau) jd->tau = edge;
                jd->n_qualified = cum;
                jd->n_sketch = sketch_size;
                break;
            }
        }
    }
}

// qualified keys <= tau, with their exact counts -> (keys, vals) in any order
__global__ void __launch_bounds__(256) k3t_collect(const CountTable ct, uint32_t n_qlist, JobDev *__restrict__ jd,
                                                   u64 *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_qlist) return;
    const u64 h = ct.qlist[i];
    if (h > jd->tau) return;
    const uint32_t c = ct_lookup(ct, h);
    const uint32_t pos = atomicAdd(&jd->n_out, 1u);
    if (pos < cap) { keys[pos] = h; vals[pos] = c; }
}

// every key <= tau with its count (multi-GPU exchange: counts below min_copies still add up across ranks)
__global__ void __launch_bounds__(256) k3t_export(const CountTable ct, JobDev *__restrict__ jd, u64 *__restrict__ keys,
                                                  uint32_t *__restrict__ vals, uint32_t cap) {
    const uint32_t tau_stride = gridDim.x * blockDim.x;
    const u64 tau = jd->tau;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot <= ct.mask; slot += tau_stride) {
        const u64 h = __ldcg(ct.key + slot);
        if (h != CT_EMPTY && h <= tau) {
            const uint32_t pos = atomicAdd(&jd->n_out, 1u);
            if (pos < cap) { keys[pos] = h; vals[pos] = __ldcg(ct.cnt + slot); }
        }
        if (slot + tau_stride < slot) break;   // 32-bit wrap (mask = 2^32 - 1 is not used, but stay safe)
    }
}

// sorted pairs -> table (all n) and sketch (first n_sk); job counters
__global__ void __launch_bounds__(256) k3t_finish(const u64 *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n,
                                                  uint32_t n_sk, uint32_t n_q, uint32_t n_sketch, u64 *__restrict__ tab_hash,
                                                  uint32_t *__restrict__ tab_cnt, u64 *__restrict__ sk_hash,
                                                  uint32_t *__restrict__ sk_cnt, JobDev *__restrict__ jd) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const u64 k = keys[i];
        const uint32_t v = vals[i];
        tab_hash[i] = k; tab_cnt[i] = v;
        if (i < n_sk) { sk_hash[i] = k; sk_cnt[i] = v; }
    }
    if (i == 0) { jd->n_table = n; jd->n_unique = n; jd->n_qualified = n_q; jd->n_sketch = n_sketch; jd->cand_count = 0; }
}

// ---- multi-GPU exchange straight from the count table (no sort): owner of hash x = #{r >= 1 : b[r] <= x}
__device__ __forceinline__ int range_of(const RangeBounds &rb, u64 x) {
    int r = 0;
    for (int j = 1; j < rb.n; j++) r += (x >= rb.b[j]) ? 1 : 0;
    return r;
}

constexpr int RG_THREADS = 256;
constexpr int RG_CTAS = 148 * 4;

__global__ void __launch_bounds__(RG_THREADS) k3t_range_count(const CountTable ct, const RangeBounds rb, JobDev *__restrict__ jd) {
    __shared__ uint32_t s_cnt[MAX_RANGES];
    if (threadIdx.x < MAX_RANGES) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u64 slots = (u64)ct.mask + 1, per = (slots + gridDim.x - 1) / gridDim.x;
    const u64 lo = (u64)blockIdx.x * per, hi = lo + per < slots ? lo + per : slots;
    for (u64 slot = lo + threadIdx.x; slot < hi; slot += RG_THREADS) {
        const u64 h = __ldcg(ct.key + slot);
        if (h != CT_EMPTY && h <= rb.tau_cut) atomicAdd(&s_cnt[range_of(rb, h)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < rb.n && s_cnt[threadIdx.x]) atomicAdd(&jd->range_cnt[threadIdx.x], s_cnt[threadIdx.x]);
}

// the same walk; a CTA reserves its share of every range with one atomic per range, then fills it
__global__ void __launch_bounds__(RG_THREADS) k3t_range_scatter(const CountTable ct, const RangeBounds rb, const RangeOffsets off,
                                                                JobDev *__restrict__ jd, u64 *__restrict__ out_hash,
                                                                uint32_t *__restrict__ out_cnt) {
    __shared__ uint32_t s_cnt[MAX_RANGES], s_base[MAX_RANGES];
    if (threadIdx.x < MAX_RANGES) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u64 slots = (u64)ct.mask + 1, per = (slots + gridDim.x - 1) / gridDim.x;
    const u64 lo = (u64)blockIdx.x * per, hi = lo + per < slots ? lo + per : slots;
    for (u64 slot = lo + threadIdx.x; slot < hi; slot += RG_THREADS) {
        const u64 h = __ldcg(ct.key + slot);
        if (h != CT_EMPTY && h <= rb.tau_cut) atomicAdd(&s_cnt[range_of(rb, h)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < rb.n) {
        const uint32_t c = s_cnt[threadIdx.x];
        s_base[threadIdx.x] = off.o[threadIdx.x] + (c ? atomicAdd(&jd->range_cnt[threadIdx.x], c) : 0u);
        s_cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    for (u64 slot = lo + threadIdx.x; slot < hi; slot += RG_THREADS) {
        const u64 h = __ldcg(ct.key + slot);
        if (h != CT_EMPTY && h <= rb.tau_cut) {
            const int r = range_of(rb, h);
            const uint32_t pos = s_base[r] + atomicAdd(&s_cnt[r], 1u);
            out_hash[pos] = h; out_cnt[pos] = __ldcg(ct.cnt + slot);
        }
    }
}

// owner merge: add (hash, count) pairs into the table; the addition that carries a sum across min_copies qualifies the key
__global__ void __launch_bounds__(256) k3t_merge_insert(const CountTable ct, const u64 *__restrict__ hashes,
                                                        const uint32_t *__restrict__ counts, uint32_t n, JobDev *__restrict__ jd) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 h = hashes[i];
    const uint32_t c = counts[i];
    if (h == CT_EMPTY) { jd->overflow = 1u; return; }     // the one value the table cannot key: the caller sorts instead
    uint32_t slot = (uint32_t)h & ct.mask;
    for (int probe = 0; probe < CT_MAX_PROBE; probe++, slot = (slot + 1u) & ct.mask) {
        u64 cur = __ldcg(ct.key + slot);
        if (cur == CT_EMPTY) cur = atomicCAS(ct.key + slot, CT_EMPTY, h);
        if (cur == CT_EMPTY || cur == h) {
            const uint32_t old = atomicAdd(ct.cnt + slot, c);
            if (old < ct.min_copies && old + c >= ct.min_copies) {
                atomicMax(&jd->max_hash, h);               // the sort of the qualified keys sizes its buckets with it
                const uint32_t q = atomicAdd(&jd->n_qlist, 1u);
                if (q < ct.qcap) ct.qlist[q] = h; else jd->overflow = 1u;
            }
            return;
        }
    }
    jd->overflow = 1u;
}

void runs_and_qualify(const LaunchCtx &lc, const SelectBufs &sb, JobDev *jd, const u64 *keys, const uint32_t *vals,
                      uint32_t n, int sketch_size, int min_copies, u64 *out_hash, uint32_t *out_cnt, int compact) {
    const uint32_t n_tiles = div_up(n ? n : 1, SC_TILE);
    uint32_t *t_heads = sb.tile_tmp, *t_vals = sb.tile_tmp + n_tiles;
    k3_runs_reduce<<<n_tiles, SC_THREADS, 0, lc.stream>>>(keys, vals, n, t_heads, t_vals);
    k3_scan_tiles<<<1, 1024, 0, lc.stream>>>(t_heads, t_vals, n_tiles, jd, sb.start_sum, 0, 0);
    k3_runs_apply<<<n_tiles, SC_THREADS, 0, lc.stream>>>(keys, vals, n, t_heads, t_vals, sb.tab_hash, sb.start_sum);
    k3_qual_reduce<<<n_tiles, SC_THREADS, 0, lc.stream>>>(sb.start_sum, jd, (uint32_t)min_copies, t_heads);
    k3_scan_tiles<<<1, 1024, 0, lc.stream>>>(t_heads, nullptr, n_tiles, jd, sb.start_sum, 1, (uint32_t)sketch_size);
    k3_qual_apply<<<n_tiles, SC_THREADS, 0, lc.stream>>>(sb.tab_hash, sb.start_sum, sb.tab_cnt, jd, (uint32_t)min_copies,
                                                         (uint32_t)sketch_size, t_heads, out_hash, out_cnt, compact);
    *lc.launches += 6;
}

}  // namespace

void launch_consolidate(const LaunchCtx &lc, const SelectBufs &sb, JobDev *jd, uint64_t n_cand, uint32_t n_table,
                        int key_bits, int sketch_size, int min_copies) {
    const uint32_t n = (uint32_t)n_cand + n_table;
    k3_prepare<<<div_up(n ? n : 1, 256), 256, 0, lc.stream>>>(sb.keys_a, sb.vals_a, (uint32_t)n_cand, sb.tab_hash,
                                                              sb.tab_cnt, n_table);
    ++*lc.launches;
    const int where = radix_sort_pairs(lc, sb.keys_a, sb.vals_a, sb.keys_b, sb.vals_b, n, key_bits, sb.hist,
                                       sb.hist + (size_t)256 * div_up(n ? n : 1, RS_TILE));
    const u64 *keys = where ? sb.keys_b : sb.keys_a;
    const uint32_t *vals = where ? sb.vals_b : sb.vals_a;
    runs_and_qualify(lc, sb, jd, keys, vals, n, sketch_size, min_copies, sb.sk_hash, sb.sk_cnt, 0);
}

void launch_table_tau_update(const LaunchCtx &lc, const CountTable &ct, JobDev *jd, int sketch_size) {
    k3t_tau_update<<<1, 1024, 0, lc.stream>>>(ct.qhist, jd, (uint32_t)sketch_size, ct.bshift);
    ++*lc.launches;
}

void launch_table_collect(const LaunchCtx &lc, const CountTable &ct, const SelectBufs &sb, JobDev *jd, uint32_t n_qlist) {
    if (!n_qlist) return;
    k3t_collect<<<div_up(n_qlist, 256), 256, 0, lc.stream>>>(ct, n_qlist, jd, sb.keys_a, sb.vals_a, (uint32_t)sb.cap);
    ++*lc.launches;
}

void launch_table_export(const LaunchCtx &lc, const CountTable &ct, const SelectBufs &sb, JobDev *jd) {
    k3t_export<<<148 * 8, 256, 0, lc.stream>>>(ct, jd, sb.keys_a, sb.vals_a, (uint32_t)sb.cap);
    ++*lc.launches;
}

void launch_table_finish(const LaunchCtx &lc, const SelectBufs &sb, JobDev *jd, uint32_t n, uint32_t n_q, int key_bits,
                         int sketch_size, int all_entries) {
    const u64 *keys;
    const uint32_t *vals;
    if (n >= 2 && n <= (1u << 20) && (size_t)256 * div_up(sb.cap, RS_TILE) >= (size_t)3 * BS_BUCKETS + 4) {
        // few, evenly spread hash values: bucket sort (4 launches) into the idle ping-pong pair
        bucket_sort_pairs(lc, sb.keys_a, sb.vals_a, sb.keys_b, sb.vals_b, n, key_bits, sb.hist);
        keys = sb.keys_b; vals = sb.vals_b;
    } else {
        const int where = radix_sort_pairs(lc, sb.keys_a, sb.vals_a, sb.keys_b, sb.vals_b, n, key_bits, sb.hist,
                                           sb.hist + (size_t)256 * div_up(n ? n : 1, RS_TILE));
        keys = where ? sb.keys_b : sb.keys_a;
        vals = where ? sb.vals_b : sb.vals_a;
    }
    const uint32_t n_sk = n_q < (uint32_t)sketch_size ? n_q : (uint32_t)sketch_size;
    // all_entries: the pairs include unqualified keys (exchange table), so there is no sketch prefix to copy
    k3t_finish<<<div_up(n ? n : 1, 256), 256, 0, lc.stream>>>(keys, vals, n, all_entries ? 0u : n_sk, n_q, n_sk, sb.tab_hash,
                                                              sb.tab_cnt, sb.sk_hash, sb.sk_cnt, jd);
    ++*lc.launches;
}

void launch_table_range_count(const LaunchCtx &lc, const CountTable &ct, const RangeBounds &bounds, JobDev *jd) {
    cudaMemsetAsync((char *)jd + offsetof(JobDev, range_cnt), 0, sizeof(uint32_t) * MAX_RANGES, lc.stream);
    k3t_range_count<<<RG_CTAS, RG_THREADS, 0, lc.stream>>>(ct, bounds, jd);
    ++*lc.launches;
}

void launch_table_range_scatter(const LaunchCtx &lc, const CountTable &ct, const RangeBounds &bounds, const RangeOffsets &off,
                                const SelectBufs &sb, JobDev *jd) {
    cudaMemsetAsync((char *)jd + offsetof(JobDev, range_cnt), 0, sizeof(uint32_t) * MAX_RANGES, lc.stream);   // now the per-range write cursors
    k3t_range_scatter<<<RG_CTAS, RG_THREADS, 0, lc.stream>>>(ct, bounds, off, jd, sb.tab_hash, sb.tab_cnt);
    ++*lc.launches;
}

void launch_merge_insert(const LaunchCtx &lc, const CountTable &ct, const unsigned long long *hashes, const uint32_t *counts,
                         uint32_t n, JobDev *jd) {
    if (!n) return;
    k3t_merge_insert<<<div_up(n, 256), 256, 0, lc.stream>>>(ct, hashes, counts, n, jd);
    ++*lc.launches;
}

void launch_finalize(const LaunchCtx &lc, const SelectBufs &sb, JobDev *jd, int use64) {
    k4_finalize<<<1, 1024, 0, lc.stream>>>(sb.sk_hash, sb.sk_cnt, jd, use64);
    ++*lc.launches;
}

void launch_table_ranges(const LaunchCtx &lc, const SelectBufs &sb, uint32_t n_table, const RangeBounds &bounds,
                         unsigned long long *d_out) {
    k3_table_ranges<<<1, MAX_RANGES + 1, 0, lc.stream>>>(sb.tab_hash, n_table, bounds, d_out);
    ++*lc.launches;
}

void launch_merge_counts(const LaunchCtx &lc, const SelectBufs &sb, JobDev *jd, uint64_t n, int min_copies) {
    const int where = radix_sort_pairs(lc, sb.keys_a, sb.vals_a, sb.keys_b, sb.vals_b, (uint32_t)n, 64, sb.hist,
                                       sb.hist + (size_t)256 * div_up(n ? n : 1, RS_TILE));
    const u64 *keys = where ? sb.keys_b : sb.keys_a;
    const uint32_t *vals = where ? sb.vals_b : sb.vals_a;
    // qualified entries are compacted into the buffer pair that does not hold the sorted input
    u64 *oh = where ? sb.keys_a : sb.keys_b;
    uint32_t *oc = where ? sb.vals_a : sb.vals_b;
    runs_and_qualify(lc, sb, jd, keys, vals, (uint32_t)n, 0x7fffffff, min_copies, oh, oc, 1);
    if (oh != sb.keys_a) {
        // results are expected in keys_a/vals_a: n_qualified <= n entries
        cudaMemcpyAsync(sb.keys_a, oh, n * sizeof(u64), cudaMemcpyDeviceToDevice, lc.stream);
        cudaMemcpyAsync(sb.vals_a, oc, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, lc.stream);
    }
}

}  // namespace ac
