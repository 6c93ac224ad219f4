// ac_hash.cu -- kernel K2: canonical k-mer MurmurHash3_x64_128(seed 42) over the
// base stream Z, threshold filter, candidate emission, record statistics.
//
// Replaces Mash's addMinHashes()/getHash() and the admission test of
// MinHashHeap::tryInsert (SURVEY.md section 2.2 stages S3-S4, App. A.3/A.4) that run
// inside `mash sketch` (auriclass/classes.py:576-596, :696-713).
//
// One CTA owns K2_TILE consecutive window starts.  The tile (+64 bytes of halo) is
// brought into shared memory by one TMA bulk copy (cp.async.bulk + mbarrier); the
// CTA then builds, once, the reverse complement of the whole tile and per-32-byte
// validity / separator bitmaps.  A thread owns K2_W consecutive windows and walks
// them in groups of 8: the k-mer's ASCII words MurmurHash3 needs are byte-granular
// slices (one PRMT per word) of the forward / reverse-complement words it loads from
// shared memory -- no rolling state, no 2-bit <-> ASCII expansion.  The group loop
// keeps the unrolled body ~16 KB so it stays in the instruction cache.
// Integer-issue bound: ~46 IMAD + ~46 LOP3/SHF/IADD3 per k-mer for the hash alone,
// ~1 byte of HBM/L2 read per k-mer.
// Hashes <= tau (device-resident, monotonically shrinking) are staged in shared
// memory and appended to the candidate buffer with one atomic per CTA.
#include "ac_common.cuh"
#include "kmer_core.cuh"

namespace ac {

namespace {

// staging full (tau close to max): opportunistic warp-aggregated append straight to the candidate buffer
__device__ __noinline__ void cand_append_global(unsigned long long h, unsigned long long *cand,
                                                unsigned long long *cand_count, unsigned long long cand_cap,
                                                uint32_t *overflow) {
    const unsigned act = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(act) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(cand_count, (unsigned long long)__popc(act));
    base = __shfl_sync(act, base, leader);
    const unsigned long long g = base + __popc(act & ((1u << lane) - 1u));
    if (g < cand_cap) cand[g] = h; else *overflow = 1u;
}

template <bool USE64>
struct CandSink {
    unsigned long long tau;
    unsigned long long *s_buf;
    uint32_t *s_cnt;
    unsigned long long *cand;
    unsigned long long *cand_count;
    unsigned long long cand_cap;
    uint32_t *overflow;
    __device__ __forceinline__ void operator()(uint32_t vbits, uint32_t bit, uint64_t h1) {
        const unsigned long long h = USE64 ? h1 : (h1 & 0xffffffffULL);  // Mash: 32-bit hashes for k <= 16
        if (h <= tau) {                // rare: tau is a small fraction of the hash space
            if (vbits & bit) {         // windows touching an invalid byte are dropped here
                const uint32_t idx = atomicAdd(s_cnt, 1u);
                if (idx < (uint32_t)K2_STAGE) s_buf[idx] = h;
                else cand_append_global(h, cand, cand_count, cand_cap, overflow);
            }
        }
    }
};

// ---- TMA (1-D bulk copy) + mbarrier helpers: sm_90+ PTX, UBLKCP in SASS
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // make the init visible to the async proxy
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

constexpr int K2_G = 8;                         // windows per unrolled group (keeps the loop body ~16 KB: fits the I-cache)
constexpr int K2_TB = K2_TILE + 64;             // tile bytes staged in shared memory (window halo + statistics look-back)

template <int K>
__global__ void __launch_bounds__(K2_THREADS, 4) k2_hash(const uint8_t *__restrict__ Z, JobDev *__restrict__ jd,
                                                         unsigned long long *__restrict__ cand,
                                                         unsigned long long cand_cap, uint32_t seed, int ps_slot) {
    __shared__ __align__(128) uint8_t s_fwd[K2_TB + 64];  // base stream bytes of the tile (TMA destination)
    __shared__ __align__(16) uint8_t s_rc[K2_TB + 64];    // reverse complement of the tile: s_rc[p'] = comp(s_fwd[TB-1-p'])
    __shared__ uint32_t s_valid[K2_TB / 32 + 2];          // bit b of word c: byte 32c+b is a valid base
    __shared__ uint32_t s_sep[K2_TB / 32 + 2];            // ... is a record separator
    __shared__ unsigned long long s_buf[K2_STAGE];
    __shared__ uint32_t s_cnt;
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long s_stat[5];  // kmers, records, kept, short_len, bases
    __shared__ __align__(8) uint64_t s_bar;

    const uint32_t n_zbuf = jd->ps[ps_slot].n_zbuf;
    const long long n_win = (long long)n_zbuf - K + 1;             // window starts 0 .. n_win-1 exist
    const long long tile_start = (long long)blockIdx.x * K2_TILE;
    // a tile is needed while it holds window starts; record-end positions p = start + K < n_zbuf
    // always belong to such a tile (p - K <= n_win - 2)
    if (tile_start >= n_win) return;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        s_cnt = 0;
        mbar_init(&s_bar, 1);
    }
    if (tid < 5) s_stat[tid] = 0;
    __syncthreads();
    if (tid == 0) tma_load_1d(s_fwd, Z + tile_start, K2_TB, &s_bar);
    mbar_wait(&s_bar, 0);

    // ---- once per CTA: reverse-complement tile + validity / separator bitmaps (chunk = 32 bytes)
    for (uint32_t c = tid; c < K2_TB / 32; c += K2_THREADS) {
        uint32_t w[8];
        const uint4 a = reinterpret_cast<const uint4 *>(s_fwd)[2 * c], b4 = reinterpret_cast<const uint4 *>(s_fwd)[2 * c + 1];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b4.x; w[5] = b4.y; w[6] = b4.z; w[7] = b4.w;
        uint32_t vm = 0, sm = 0, rcw[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            vm |= ackmer::valid_nibble(w[j]) << (4 * j);
            const uint32_t nz = ((w[j] >> 6) | w[j]) & 0x01010101u;
            sm |= (((nz * 0x01020408u) >> 24) ^ 0xFu) << (4 * j);
            rcw[7 - j] = __byte_perm(ackmer::comp4(w[j]), 0, 0x0123);
        }
        s_valid[c] = vm; s_sep[c] = sm;
        uint4 *dst = reinterpret_cast<uint4 *>(s_rc + (K2_TB - 32 * (c + 1)));
        dst[0] = make_uint4(rcw[0], rcw[1], rcw[2], rcw[3]);
        dst[1] = make_uint4(rcw[4], rcw[5], rcw[6], rcw[7]);
    }
    __syncthreads();

    const long long run_start = tile_start + (long long)tid * K2_W;
    // windows of this thread allowed by the stream range: first unprocessed start is
    // Z_CARRY-(K-1) (earlier ones were hashed with the previous segment), last is n_win-1
    long long wlo = (long long)(Z_CARRY - (K - 1)) - run_start, whi = n_win - run_start;
    wlo = wlo < 0 ? 0 : wlo; whi = whi > K2_W ? K2_W : whi;
    uint32_t win_ok = 0;
    if (whi > wlo) win_ok = (whi >= 32 ? 0xffffffffu : ((1u << whi) - 1u)) & ~((1u << wlo) - 1u);
    const uint64_t vm64 = (uint64_t)s_valid[tid] | ((uint64_t)s_valid[tid + 1] << 32);
    const uint32_t vmask = (uint32_t)ackmer::windows_all_valid<K>(vm64) & win_ok;

    CandSink<(K > 16)> sink;
    sink.tau = jd->tau;
    sink.s_buf = s_buf; sink.s_cnt = &s_cnt;
    sink.cand = cand; sink.cand_count = &jd->cand_count; sink.cand_cap = cand_cap;
    sink.overflow = &jd->overflow;

    if (__any_sync(FULL, vmask != 0)) {
        // forward words of group g start at byte 32*tid + 8*g; reverse-complement window i of the
        // group starts at byte RCB - 32*tid - 8*g - i with RCB = TB - K; both word-aligned up to a
        // compile-time byte offset
        constexpr int NF = (K2_G - 1 + K + 3) / 4;                 // forward words per group
        constexpr int RC_LOW = K2_TB - K - (K2_G - 1);             // byte where window G-1 of (tid 0, g 0) starts
        constexpr int RC_WORD = RC_LOW >> 2;                       // its word
        constexpr int RC_PAR = RC_WORD & 1;                        // keep 8-byte aligned loads
        constexpr int RC_OFF = (RC_LOW & 3) + 4 * RC_PAR;          // byte offset of window G-1 inside the loaded words
        constexpr int NR = (RC_OFF + (K2_G - 1) + K + 3) / 4;      // reverse words per group
        const uint2 *pf = reinterpret_cast<const uint2 *>(s_fwd) + 4 * tid;
        const uint2 *pr = reinterpret_cast<const uint2 *>(s_rc) + ((RC_WORD - RC_PAR) >> 1) - 4 * tid;
#pragma unroll 1
        for (int g = 0; g < K2_W / K2_G; g++) {
            uint32_t fw[2 * ((NF + 1) / 2)], rw[2 * ((NR + 1) / 2)];
#pragma unroll
            for (int j = 0; j < (NF + 1) / 2; j++) { const uint2 v = pf[g + j]; fw[2 * j] = v.x; fw[2 * j + 1] = v.y; }
#pragma unroll
            for (int j = 0; j < (NR + 1) / 2; j++) { const uint2 v = pr[j - g]; rw[2 * j] = v.x; rw[2 * j + 1] = v.y; }
            ackmer::process_group<K, K2_G, RC_OFF + (K2_G - 1), 0>(fw, rw, vmask >> (K2_G * g), seed, sink);
        }
    }

    // ---- record statistics (Mash skips records shorter than k; FASTA `length` = sum of kept lengths)
    // positions j in [K, K+W) of the run <-> stream positions p = run_start + j in [Z_CARRY, n_zbuf)
    uint32_t n_rec = 0, n_kept = 0, short_len = 0, n_bases = 0;
    {
        const uint64_t sep = (uint64_t)s_sep[tid] | ((uint64_t)s_sep[tid + 1] << 32);
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
void launch_k(const LaunchCtx &lc, const uint8_t *Z, uint32_t max_zbuf, int ps_slot, JobDev *jd, unsigned long long *cand,
              unsigned long long cand_cap, uint32_t seed) {
    const unsigned grid = div_up(max_zbuf, K2_TILE);
    k2_hash<K><<<grid, K2_THREADS, 0, lc.stream>>>(Z, jd, cand, cand_cap, seed, ps_slot);
    ++*lc.launches;
}

}  // namespace

#define AC_K_CASE(KK) case KK: launch_k<KK>(lc, Z, max_zbuf, ps_slot, jd, cand, cand_cap, seed); break;

void launch_hash_segment(const LaunchCtx &lc, int k, const uint8_t *Z, uint32_t max_zbuf, int ps_slot, JobDev *jd,
                         unsigned long long *cand, unsigned long long cand_cap, uint32_t seed, int use64) {
    (void)use64;  // the hash width follows k (CandSink<(K > 16)>)
    switch (k) {
#ifdef AC_ONLY_K27  // development builds: one instantiation, seconds to compile
        AC_K_CASE(27)
#else
        AC_K_CASE(1) AC_K_CASE(2) AC_K_CASE(3) AC_K_CASE(4) AC_K_CASE(5) AC_K_CASE(6) AC_K_CASE(7) AC_K_CASE(8)
        AC_K_CASE(9) AC_K_CASE(10) AC_K_CASE(11) AC_K_CASE(12) AC_K_CASE(13) AC_K_CASE(14) AC_K_CASE(15) AC_K_CASE(16)
        AC_K_CASE(17) AC_K_CASE(18) AC_K_CASE(19) AC_K_CASE(20) AC_K_CASE(21) AC_K_CASE(22) AC_K_CASE(23) AC_K_CASE(24)
        AC_K_CASE(25) AC_K_CASE(26) AC_K_CASE(27) AC_K_CASE(28) AC_K_CASE(29) AC_K_CASE(30) AC_K_CASE(31) AC_K_CASE(32)
#endif
        default: break;
    }
}

}  // namespace ac
