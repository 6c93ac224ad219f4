// ac_hash.cu -- kernel K2: canonical k-mer MurmurHash3_x64_128(seed 42) over the
// base stream Z, threshold filter, candidate emission, record statistics.
//
// Replaces Mash's addMinHashes()/getHash() and the admission test of
// MinHashHeap::tryInsert (SURVEY.md section 2.2 stages S3-S4, App. A.3/A.4) that run
// inside `mash sketch` (auriclass/classes.py:576-596, :696-713).
//
// One CTA owns K2_TILE consecutive window starts; a thread owns K2_W of them and
// keeps its W+K bytes in registers (kmer_core.cuh: slices of the run and of its
// reverse complement feed MurmurHash3 directly as little-endian ASCII words).
// Integer-issue bound: ~46 IMAD + ~46 LOP3/SHF/IADD3 per k-mer for the hash, no
// shared-memory traffic on the hot path, 2 bytes of HBM/L2 read per k-mer.
// Hashes <= tau (device-resident, monotonically shrinking) are staged in shared
// memory and appended to the candidate buffer with one atomic per CTA.
#include "ac_common.cuh"
#include "kmer_core.cuh"

namespace ac {

namespace {

struct CandSink {
    unsigned long long tau;
    unsigned long long lowmask;  // 32-bit mode: keep the low 32 bits of h1
    unsigned long long *s_buf;
    uint32_t *s_cnt;
    unsigned long long *cand;
    unsigned long long *cand_count;
    unsigned long long cand_cap;
    uint32_t *overflow;
    __device__ __forceinline__ void operator()(int, uint64_t h1) {
        const unsigned long long h = h1 & lowmask;
        if (h <= tau) {
            const uint32_t idx = atomicAdd(s_cnt, 1u);
            if (idx < (uint32_t)K2_STAGE) {
                s_buf[idx] = h;
            } else {
                // staging full (tau close to max): opportunistic warp-aggregated global append
                const unsigned act = __activemask();
                const int lane = threadIdx.x & 31;
                const int leader = __ffs(act) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(cand_count, (unsigned long long)__popc(act));
                base = __shfl_sync(act, base, leader);
                const unsigned long long g = base + __popc(act & ((1u << lane) - 1u));
                if (g < cand_cap) cand[g] = h; else *overflow = 1u;
            }
        }
    }
};

template <int K>
__global__ void __launch_bounds__(K2_THREADS) k2_hash(const uint8_t *__restrict__ Z, JobDev *__restrict__ jd,
                                                      unsigned long long *__restrict__ cand,
                                                      unsigned long long cand_cap, uint32_t seed, int use64) {
    using C = ackmer::RunCfg<K, K2_W>;
    __shared__ unsigned long long s_buf[K2_STAGE];
    __shared__ uint32_t s_cnt;
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long s_stat[5];  // kmers, records, kept, short_len, bases

    const uint32_t n_zbuf = jd->n_zbuf;
    const long long n_win = (long long)n_zbuf - K + 1;             // window starts 0 .. n_win-1 exist
    const long long tile_start = (long long)blockIdx.x * K2_TILE;
    // a tile is needed while it holds window starts; record-end positions p = start + K < n_zbuf
    // always belong to such a tile (p - K <= n_win - 2)
    if (tile_start >= n_win) return;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) s_cnt = 0;
    if (tid < 5) s_stat[tid] = 0;
    __syncthreads();

    const long long run_start = tile_start + (long long)tid * K2_W;
    uint32_t r[C::RW];
#pragma unroll
    for (int v = 0; v < C::RW / 4; v++) {
        const uint4 x = __ldg(reinterpret_cast<const uint4 *>(Z + run_start) + v);
        r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
    }
    // windows of this thread allowed by the stream range: first unprocessed start is
    // Z_CARRY-(K-1) (earlier ones were hashed with the previous segment), last is n_win-1
    long long wlo = (long long)(Z_CARRY - (K - 1)) - run_start, whi = n_win - run_start;
    wlo = wlo < 0 ? 0 : wlo; whi = whi > K2_W ? K2_W : whi;
    uint32_t win_ok = 0;
    if (whi > wlo) win_ok = (whi >= 32 ? 0xffffffffu : ((1u << whi) - 1u)) & ~((1u << wlo) - 1u);

    CandSink sink;
    sink.tau = jd->tau;
    sink.lowmask = use64 ? ~0ULL : 0xffffffffULL;
    sink.s_buf = s_buf; sink.s_cnt = &s_cnt;
    sink.cand = cand; sink.cand_count = &jd->cand_count; sink.cand_cap = cand_cap;
    sink.overflow = &jd->overflow;
    const uint32_t vmask = ackmer::process_run<K, K2_W>(r, win_ok, seed, sink);

    // ---- record statistics (Mash skips records shorter than k; FASTA `length` = sum of kept lengths)
    // positions j in [K, K+W) of the run <-> stream positions p = run_start + j in [Z_CARRY, n_zbuf)
    uint32_t n_rec = 0, n_kept = 0, short_len = 0, n_bases = 0;
    {
        uint64_t sep = 0;  // bit j: byte j of the run is a record separator (0x00)
#pragma unroll
        for (int j = 0; j < C::RW; j++) {
            const uint32_t nz = ((r[j] >> 6) | r[j]) & 0x01010101u;
            sep |= (uint64_t)(((nz * 0x01020408u) >> 24) ^ 0xFu) << (4 * j);
        }
        long long plo = (long long)Z_CARRY - run_start, phi = (long long)n_zbuf - run_start;
        plo = plo < K ? K : plo; phi = phi > K + K2_W ? K + K2_W : phi;
        if (phi > plo) {
            const uint64_t range = ((phi >= 64) ? ~0ULL : ((1ULL << phi) - 1ULL)) & ~((1ULL << plo) - 1ULL);
            n_bases = __popcll(~sep & range);
            uint64_t todo = sep & range;
            while (todo) {
                const int j = __ffsll((long long)todo) - 1;
                todo &= todo - 1;
                const uint64_t back = sep & (((1ULL << j) - 1ULL) & ~((1ULL << (j - K)) - 1ULL));  // bytes j-K .. j-1
                if (back == 0) { n_rec++; n_kept++; }
                else {
                    const int pj = 63 - __clzll((long long)back);
                    const int L = j - pj - 1;
                    if (L > 0) { n_rec++; short_len += L; }
                }
            }
        }
    }
    uint32_t nk = __popc(vmask);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        nk += __shfl_xor_sync(FULL, nk, d);
        n_rec += __shfl_xor_sync(FULL, n_rec, d);
        n_kept += __shfl_xor_sync(FULL, n_kept, d);
        short_len += __shfl_xor_sync(FULL, short_len, d);
        n_bases += __shfl_xor_sync(FULL, n_bases, d);
    }
    if ((tid & 31) == 0) {
        if (nk) atomicAdd(&s_stat[0], (unsigned long long)nk);
        if (n_rec) atomicAdd(&s_stat[1], (unsigned long long)n_rec);
        if (n_kept) atomicAdd(&s_stat[2], (unsigned long long)n_kept);
        if (short_len) atomicAdd(&s_stat[3], (unsigned long long)short_len);
        if (n_bases) atomicAdd(&s_stat[4], (unsigned long long)n_bases);
    }
    __syncthreads();
    const uint32_t n_staged = min(s_cnt, (uint32_t)K2_STAGE);
    if (tid == 0) {
        if (s_stat[0]) atomicAdd(&jd->n_kmers, s_stat[0]);
        if (s_stat[1]) atomicAdd(&jd->n_records, s_stat[1]);
        if (s_stat[2]) atomicAdd(&jd->n_kept, s_stat[2]);
        if (s_stat[3]) atomicAdd(&jd->n_short_len, s_stat[3]);
        if (s_stat[4]) atomicAdd(&jd->n_bases, s_stat[4]);
        if (s_cnt) atomicAdd(&jd->cand_total, (unsigned long long)s_cnt);
        s_base = n_staged ? atomicAdd(&jd->cand_count, (unsigned long long)n_staged) : 0ULL;
    }
    __syncthreads();
    const unsigned long long base = s_base;
    for (uint32_t i = tid; i < n_staged; i += K2_THREADS) {
        const unsigned long long g = base + i;
        if (g < cand_cap) cand[g] = s_buf[i]; else jd->overflow = 1u;
    }
}

template <int K>
void launch_k(const LaunchCtx &lc, const uint8_t *Z, uint32_t max_zbuf, JobDev *jd, unsigned long long *cand,
              unsigned long long cand_cap, uint32_t seed, int use64) {
    const unsigned grid = div_up(max_zbuf, K2_TILE);
    k2_hash<K><<<grid, K2_THREADS, 0, lc.stream>>>(Z, jd, cand, cand_cap, seed, use64);
    ++*lc.launches;
}

}  // namespace

#define AC_K_CASE(KK) case KK: launch_k<KK>(lc, Z, max_zbuf, jd, cand, cand_cap, seed, use64); break;

void launch_hash_segment(const LaunchCtx &lc, int k, const uint8_t *Z, uint32_t max_zbuf, JobDev *jd,
                         unsigned long long *cand, unsigned long long cand_cap, uint32_t seed, int use64) {
    switch (k) {
        AC_K_CASE(1) AC_K_CASE(2) AC_K_CASE(3) AC_K_CASE(4) AC_K_CASE(5) AC_K_CASE(6) AC_K_CASE(7) AC_K_CASE(8)
        AC_K_CASE(9) AC_K_CASE(10) AC_K_CASE(11) AC_K_CASE(12) AC_K_CASE(13) AC_K_CASE(14) AC_K_CASE(15) AC_K_CASE(16)
        AC_K_CASE(17) AC_K_CASE(18) AC_K_CASE(19) AC_K_CASE(20) AC_K_CASE(21) AC_K_CASE(22) AC_K_CASE(23) AC_K_CASE(24)
        AC_K_CASE(25) AC_K_CASE(26) AC_K_CASE(27) AC_K_CASE(28) AC_K_CASE(29) AC_K_CASE(30) AC_K_CASE(31) AC_K_CASE(32)
        default: break;
    }
}

}  // namespace ac
