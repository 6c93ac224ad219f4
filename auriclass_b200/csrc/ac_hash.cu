// ac_hash.cu -- kernel K2: canonical k-mer MurmurHash3_x64_128(seed 42) over the
// base stream Z, threshold filter, candidate emission, record statistics.
//
// Replaces Mash's addMinHashes()/getHash() and the admission test of
// MinHashHeap::tryInsert (SURVEY.md section 2.2 stages S3-S4, App. A.3/A.4) that run
// inside `mash sketch` (auriclass/classes.py:576-596, :696-713).
//
// Every warp owns K2_WT consecutive window starts (a CTA is K2_WARPS such warps and one barrier).  The warp's
// sub-tile (+64 bytes of halo) is brought into shared memory by one TMA bulk copy (cp.async.bulk + mbarrier); the
// warp then builds, once, the reverse complement of the sub-tile, per-32-byte validity / separator bitmaps, the
// bitmap of valid window starts, and from it the list of *groups*: every maximal run of valid windows is cut into
// pieces of 32, so lanes are only spent where there is a k-mer to hash (97 % of the slots for 150 bp reads instead
// of 82 % with a fixed assignment; gaps of any length cost nothing).
// A quad of lanes takes a group: lane c hashes windows g+c, g+c+4, ... (stride 4 bytes), so one byte alignment of
// the 15 words it loads per strand (funnel shifts) makes every key word of its 8 windows a whole register -- no
// per-window byte slicing, no rolling state, no 2-bit <-> ASCII expansion.
// Count-table passes (tau <= 1/32 of the hash space) run the FLAT loop: the 8 windows of a lane are one basic block
// (strand from the first four bases, threshold gate on the high words of the two last fmix products); the ~1 window
// in 120 that ties or passes the gate is parked in the warp's redo list and hashed again, exactly, by all lanes
// together when the sub-tile is done.  Passes that start at tau = max keep the branchy loop.
// Integer-issue bound: 14 IMAD.WIDE + 31 IMAD/IMAD.X and ~47 SHF/LOP3/IADD3/SEL/ISETP per k-mer, both integer pipes
// ~72 % busy; ~1 byte of HBM/L2 read per k-mer.
// Hashes <= tau (device-resident, monotonically shrinking) are staged in shared memory and inserted into the count
// table (or appended to the candidate buffer) by the warp when its sub-tile is done.
#include "ac_common.cuh"
#include "kmer_core.cuh"

namespace ac {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int K2_REDO = 128;                    // flagged windows a warp parks per sub-tile (flat hot loop)
constexpr int K2_TB = K2_WT + 64;               // sub-tile bytes staged per warp (window halo + statistics look-back)

// Where candidates go.  Lives in shared memory so that the hot loop carries no kernel parameters
// (ptxas otherwise hoists their constant-bank loads in front of every window's threshold branch).
struct SinkShared {
    unsigned long long tau;
    unsigned long long *cand;
    unsigned long long *cand_count;
    unsigned long long cand_cap;
    uint32_t *overflow;
    JobDev *jd;
    CountTable ct;                         // ct.mask != 0: candidates go into the count table
    uint32_t cnt;                          // staged candidates of this CTA (may exceed K2_STAGE: the excess went to global)
    uint32_t n_redo;                       // flagged windows of the sub-tile (flat hot loop); may exceed K2_REDO: the excess was redone on the spot
    unsigned long long buf[K2_STAGE];
    uint16_t redo[K2_REDO];                // their start bytes inside the sub-tile
};

// staging full (tau close to max): opportunistic warp-aggregated append straight to the candidate buffer
__device__ __noinline__ void cand_append_global(unsigned long long h, SinkShared *ss) {
    if (ss->ct.mask) { ct_insert(ss->ct, h, ss->jd); return; }
    const unsigned act = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(act) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(ss->cand_count, (unsigned long long)__popc(act));
    base = __shfl_sync(act, base, leader);
    const unsigned long long g = base + __popc(act & ((1u << lane) - 1u));
    if (g < ss->cand_cap) ss->cand[g] = h; else *ss->overflow = 1u;
}

// rare path (about 1 window in 100 gets here, 1 in 2 of those is kept): finish the hash, exact test, stage
template <bool USE64>
__device__ __noinline__ void cand_stage(uint64_t p1, uint64_t p2, uint32_t window_ok, SinkShared *ss) {
    if (!window_ok) return;            // windows touching an invalid byte are dropped here
    const uint64_t h1 = ackmer::murmur3_close(p1, p2);
    const unsigned long long h = USE64 ? h1 : (h1 & 0xffffffffULL);  // Mash: 32-bit hashes for k <= 16
    if (h > ss->tau) return;
    uint32_t idx;   // plain shared atomic: one lane at a time gets here, warp aggregation would cost more than it saves
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(idx) : "r"(smem_u32(&ss->cnt)) : "memory");
    if (idx < (uint32_t)K2_STAGE) ss->buf[idx] = h;
    else cand_append_global(h, ss);
}

// Flat hot loop: a flagged window (first four bases of the two strands tie, or the gate value passes) is hashed
// again from the sub-tile in shared memory, exactly, and staged if it is a candidate.
template <int K>
__device__ __noinline__ void redo_window(const uint32_t *fwd_w, const uint32_t *rc_w, uint32_t p, uint32_t rc_top,
                                         uint32_t seed, SinkShared *ss) {
    const unsigned long long h = ackmer::rehash_window<K>(fwd_w, rc_w, p, rc_top - p, seed);
    if (h > ss->tau) return;
    uint32_t idx;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(idx) : "r"(smem_u32(&ss->cnt)) : "memory");
    if (idx < (uint32_t)K2_STAGE) ss->buf[idx] = h;
    else cand_append_global(h, ss);
}

template <bool USE64>
struct CandSink {
    uint32_t gate;   // ackmer::gate_of_tau
    SinkShared *ss;
    __device__ __forceinline__ void operator()(uint32_t vbits, uint32_t bit, uint64_t p1, uint64_t p2) {
        // p1, p2: half-open fmix values (kmer_core.cuh).  64-bit hashes are rejected on the high word of the digest
        // alone (gate_value); cand_stage finishes the hash and tests exactly.
        const bool maybe = ackmer::gate_value<USE64>(p1, p2) <= gate;
        if (maybe) cand_stage<USE64>(p1, p2, vbits & bit, ss);
    }
};

// ---- TMA (1-D bulk copy) + mbarrier helpers: sm_90+ PTX, UBLKCP in SASS
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

constexpr int K2_NWV = K2_WT / 32;              // words of a warp's valid-window bitmap
constexpr int K2_ROUNDS = K2_NWV / 32;          // a lane owns word (32 r + lane) in round r
// group list capacity: two runs are at least K invalid windows apart, a run of L windows makes ceil(L/32) groups
template <int K> struct K2MaxGroups { static constexpr int value = K2_WT / (K + 1) + K2_WT / 32 + 8; };

__device__ __forceinline__ uint32_t warp_incl_max(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, v, d);
        if (lane >= (uint32_t)d) v = max(v, t);
    }
    return v;
}
__device__ __forceinline__ uint32_t warp_incl_sum(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, v, d);
        if (lane >= (uint32_t)d) v += t;
    }
    return v;
}

// Every WARP owns K2_WT consecutive window starts and works through them on its own: TMA load of its sub-tile,
// reverse complement + bitmaps, group list, hash loop, record statistics, candidate flush -- synchronised with
// __syncwarp only.  The prologue / epilogue of one warp (latency bound) thus overlaps the hash loops of the other
// warps on its sub-partition instead of idling whole CTAs at barriers: with CTA-wide phases only ~64 % of the time
// enough warps were inside the loop to keep the fmaheavy pipe busy.
// FLAT: the hot loop is one basic block per 8 windows (process_quad_windows_flat); flagged windows are parked in the
// warp's redo list and worked off, all lanes busy, when the sub-tile is done.  Used for count-table passes (tau <=
// 1/32 of the hash space: ~1 window in 120 is flagged); passes that start at tau = max keep the branchy loop, whose
// rare path is taken by every window there.
template <int K, bool FLAT>
__global__ void __launch_bounds__(K2_THREADS, AC_K2_MINB) k2_hash(const uint8_t *__restrict__ Z, JobDev *__restrict__ jd,
                                                         unsigned long long *__restrict__ cand,
                                                         unsigned long long cand_cap, uint32_t seed, int ps_slot,
                                                         const CountTable ct, uint32_t cta0) {
    using Q = ackmer::QuadCfg<K>;
    constexpr int K2_MAXG = K2MaxGroups<K>::value;
    __shared__ __align__(128) uint8_t s_fwd_all[K2_WARPS][K2_TB + 64];  // base stream bytes of the sub-tile (TMA destination)
    __shared__ __align__(16) uint8_t s_rc_all[K2_WARPS][K2_TB + 64];    // its reverse complement: rc[p'] = comp(fwd[TB-1-p'])
    __shared__ uint32_t s_valid_all[K2_WARPS][K2_TB / 32 + 2];          // bit b of word c: byte 32c+b is a valid base
    __shared__ uint32_t s_sep_all[K2_WARPS][K2_TB / 32 + 2];            // ... is a record separator
    __shared__ uint32_t s_wv_all[K2_WARPS][K2_NWV + 1];                 // bit b of word c: a k-mer to hash starts at byte 32c+b
    __shared__ uint16_t s_groups_all[K2_WARPS][K2_MAXG];                // start positions of the groups of <= 32 consecutive valid windows
    __shared__ SinkShared s_sink_all[K2_WARPS];
    __shared__ __align__(8) uint64_t s_bar_all[K2_WARPS];
#if defined(AC_K2_STAT64)
    __shared__ unsigned long long s_stat[6];  // kmers, records, kept, short_len, bases, candidates: summed over the CTA's warps
#else
    __shared__ uint32_t s_stat[6];            // kmers, records, kept, short_len, bases, candidates: summed over the CTA's warps
                                              // (32 bits hold a CTA's totals; 64-bit shared atomics are a CAS loop)
#endif
    __shared__ uint32_t s_done;               // warps that have finished; the last one adds s_stat to the job counters

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 6) s_stat[tid] = 0;
    if (tid == 0) s_done = 0;
    __syncthreads();   // the only CTA barrier: every warp is still at its first instruction

    const uint32_t n_zbuf = jd->ps[ps_slot].n_zbuf, z_begin = jd->ps[ps_slot].z_begin;
    const long long n_win = (long long)n_zbuf - K + 1;             // window starts 0 .. n_win-1 exist
    const long long tile_start = ((long long)(cta0 + blockIdx.x) * K2_WARPS + warp) * K2_WT;
    uint32_t nk = 0, n_rec = 0, n_kept = 0, short_len = 0, n_bases = 0, n_emitted = 0;

    // a sub-tile is needed while it holds window starts; record-end positions p = start + K < n_zbuf
    // always belong to such a sub-tile (p - K <= n_win - 2)
    if (tile_start < n_win) {
        uint8_t *s_fwd = s_fwd_all[warp], *s_rc = s_rc_all[warp];
        uint32_t *s_valid = s_valid_all[warp], *s_sep = s_sep_all[warp], *s_wv = s_wv_all[warp];
        uint16_t *s_groups = s_groups_all[warp];
        SinkShared &s_sink = s_sink_all[warp];
        uint64_t *s_bar = &s_bar_all[warp];
        if (lane == 0) {
            s_wv[K2_NWV] = 0;
            s_sink.cnt = 0; s_sink.n_redo = 0; s_sink.tau = jd->tau;
            s_sink.cand = cand; s_sink.cand_count = &jd->cand_count; s_sink.cand_cap = cand_cap;
            s_sink.overflow = &jd->overflow; s_sink.jd = jd; s_sink.ct = ct;
            mbar_init(s_bar, 1);
        }
        __syncwarp();
        if (lane == 0) tma_load_1d(s_fwd, Z + tile_start, K2_TB, s_bar);
        mbar_wait(s_bar, 0);

        // ---- once per sub-tile: reverse complement + validity / separator bitmaps (chunk = 32 bytes)
#pragma unroll
        for (uint32_t it = 0; it < (K2_TB / 32 + 31) / 32; it++) {   // (unrolled: two chunks in flight per lane)
            const uint32_t c = lane + 32u * it;
            if (c >= K2_TB / 32) break;
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
        __syncwarp();

        // ---- valid-window bitmap and group list.  Runs of valid windows are cut every 32 windows counted from the
        // run's start, so a run continuing from the previous word puts a cut at bit (its start position mod 32) of
        // this word.  Lane l owns word 32 r + l in round r; the run state is carried from round to round.
        uint32_t n_groups = 0, carry_last = 0, carry_enc = 0;
#pragma unroll 1
        for (int r = 0; r < K2_ROUNDS; r++) {
            const uint32_t wi = 32u * r + lane;
            const long long run_start = tile_start + (long long)wi * K2_W;
            // windows allowed by the stream range: first unprocessed start is z_begin-(K-1) (earlier ones were
            // hashed with the previous segment), last is n_win-1
            long long wlo = (long long)z_begin - (K - 1) - run_start, whi = n_win - run_start;
            wlo = wlo < 0 ? 0 : wlo; whi = whi > K2_W ? K2_W : whi;
            uint32_t win_ok = 0;
            if (whi > wlo) win_ok = (whi >= 32 ? 0xffffffffu : ((1u << whi) - 1u)) & ~((1u << wlo) - 1u);
            const uint64_t vm64 = (uint64_t)s_valid[wi] | ((uint64_t)s_valid[wi + 1] << 32);
            const uint32_t vmask = (uint32_t)ackmer::windows_all_valid<K>(vm64) & win_ok;
            s_wv[wi] = vmask;
            nk += __popc(vmask);

            uint32_t prev_last = __shfl_up_sync(FULL, vmask >> 31, 1);
            if (lane == 0) prev_last = carry_last;                       // sub-tile start = fresh run (carry_last = 0)
            const bool cont = (vmask & 1u) && prev_last;
            const uint32_t starts = vmask & ~((vmask << 1) | (cont ? 1u : 0u));
            // start bit of the run that reaches the end of this word (0 = this word passes an older run through / no run)
            uint32_t enc = 0;
            if ((vmask >> 31) && starts) enc = ((wi + 1u) << 5) | (31u - __clz(starts));
            const uint32_t inc = max(warp_incl_max(enc, lane), carry_enc);
            uint32_t before = __shfl_up_sync(FULL, inc, 1);
            if (lane == 0) before = carry_enc;
            uint32_t extra = 0xffffffffu;                  // cut of the continuing run inside this word, if any
            if (cont) {
                const uint32_t phi = before & 31u;         // before != 0: the run started at some word's start bit
                const uint32_t ones = (vmask == 0xffffffffu) ? 32u : (uint32_t)(__ffs(~vmask) - 1);
                if (phi < ones && phi != 0u) extra = phi;  // phi == 0: the run started at bit 0 of a word, its cut is bit 0 here
                else if (phi == 0u) extra = 0u;
            }
            const uint32_t n_mine = __popc(starts) + (extra != 0xffffffffu ? 1u : 0u);
            const uint32_t incl = warp_incl_sum(n_mine, lane);
            uint32_t slot = n_groups + incl - n_mine;
            if (extra != 0xffffffffu) { if (slot < K2_MAXG) s_groups[slot] = (uint16_t)(32u * wi + extra); slot++; }
            for (uint32_t st = starts; st; st &= st - 1, slot++)
                if (slot < K2_MAXG) s_groups[slot] = (uint16_t)(32u * wi + (uint32_t)(__ffs(st) - 1));
            n_groups += __shfl_sync(FULL, incl, 31);
            carry_last = __shfl_sync(FULL, vmask >> 31, 31);
            carry_enc = __shfl_sync(FULL, inc, 31);
        }
        __syncwarp();
        if (n_groups > (uint32_t)K2_MAXG) __trap();  // cannot happen (see K2MaxGroups)

        CandSink<(K > 16)> sink;
        sink.gate = ackmer::gate_of_tau(s_sink.tau, K > 16);
        sink.ss = &s_sink;
        const ackmer::MurmurConsts mc;   // (registers are too scarce to hide these from ptxas' constant folding: 64/thread)

        // ---- hash: a quad per group, lane c takes windows g+c+4j
        {
            const uint32_t quad = lane >> 2, c = lane & 3u;
            const uint32_t *fwd_w = reinterpret_cast<const uint32_t *>(s_fwd);
            const uint32_t *rc_w = reinterpret_cast<const uint32_t *>(s_rc);
#pragma unroll 1
            for (uint32_t gi = quad; gi < n_groups; gi += 8u) {
                const uint32_t g = s_groups[gi];
                // the 32 windows from g on, restricted to the run that starts at g (a later run has its own group)
                uint32_t vb = __funnelshift_r(s_wv[g >> 5], s_wv[(g >> 5) + 1], g & 31u);
                vb &= ~(vb + 1u);
                vb >>= c;
                const uint32_t A = g + c;                                  // byte of the lane's window 0
                const uint32_t QL = (uint32_t)(K2_TB - K) - A - 4u * (Q::G - 1);  // byte where its window G-1's reverse complement starts
                uint32_t in[Q::NL], fw[Q::NS], rw[Q::NS];
#pragma unroll
                for (int i = 0; i < Q::NL; i++) in[i] = fwd_w[(A >> 2) + i];
                ackmer::preshift<Q::NS>(in, A & 3u, fw);
#pragma unroll
                for (int i = 0; i < Q::NL; i++) in[i] = rc_w[(QL >> 2) + i];
                ackmer::preshift<Q::NS>(in, QL & 3u, rw);
                if constexpr (FLAT) {
                    uint32_t flags = 0;
                    ackmer::process_quad_windows_flat<K, 0, (K > 16)>(fw, rw, seed, mc, sink.gate, flags);
                    flags &= vb;                     // bit 4j: the lane's window j (byte A + 4j) is valid and flagged
                    while (flags) {                  // ~2 lanes of the warp per pass
                        const uint32_t pos = A + (uint32_t)(__ffs(flags) - 1);
                        flags &= flags - 1u;
                        uint32_t idx;
                        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(idx) : "r"(smem_u32(&s_sink.n_redo)) : "memory");
                        if (idx < (uint32_t)K2_REDO) s_sink.redo[idx] = (uint16_t)pos;
                        else redo_window<K>(fwd_w, rc_w, pos, (uint32_t)(K2_TB - K), seed, &s_sink);
                    }
                } else {
                    ackmer::process_quad_windows<K, 0>(fw, rw, vb, seed, mc, sink);
                }
            }
            if constexpr (FLAT) {
                __syncwarp();
                const uint32_t n_redo = min(s_sink.n_redo, (uint32_t)K2_REDO);
                for (uint32_t i = lane; i < n_redo; i += 32)
                    redo_window<K>(fwd_w, rc_w, s_sink.redo[i], (uint32_t)(K2_TB - K), seed, &s_sink);
            }
        }

        // ---- record statistics (Mash skips records shorter than k; FASTA `length` = sum of kept lengths)
        // positions j in [K, K+W) of the run <-> stream positions p = run_start + j in [z_begin, n_zbuf)
#pragma unroll 1
        for (int r = 0; r < K2_ROUNDS; r++) {
            const uint32_t wi = 32u * r + lane;
            const long long run_start = tile_start + (long long)wi * K2_W;
            const uint64_t sep = (uint64_t)s_sep[wi] | ((uint64_t)s_sep[wi + 1] << 32);
            long long plo = (long long)z_begin - run_start, phi = (long long)n_zbuf - run_start;
            plo = plo < K ? K : plo; phi = phi > K + K2_W ? K + K2_W : phi;
            if (phi > plo) {
                const uint64_t range = ((phi >= 64) ? ~0ULL : ((1ULL << phi) - 1ULL)) & ~((1ULL << plo) - 1ULL);
                n_bases += __popcll(~sep & range);
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

        // ---- candidates of this sub-tile
        __syncwarp();
        n_emitted = s_sink.cnt;
        const uint32_t n_staged = min(n_emitted, (uint32_t)K2_STAGE);
        if (ct.mask) {   // count-table mode: one insertion per staged candidate
            for (uint32_t i = lane; i < n_staged; i += 32) ct_insert(ct, s_sink.buf[i], jd);
        } else if (n_staged) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&jd->cand_count, (unsigned long long)n_staged);
            base = __shfl_sync(FULL, base, 0);
            for (uint32_t i = lane; i < n_staged; i += 32) {
                const unsigned long long g = base + i;
                if (g < cand_cap) cand[g] = s_sink.buf[i]; else jd->overflow = 1u;
            }
        }
    }

    // ---- job counters: warp -> CTA (shared atomics) -> one set of global atomics by the CTA's last warp
    nk = __reduce_add_sync(FULL, nk); n_rec = __reduce_add_sync(FULL, n_rec); n_kept = __reduce_add_sync(FULL, n_kept);
    short_len = __reduce_add_sync(FULL, short_len); n_bases = __reduce_add_sync(FULL, n_bases);
    if (lane == 0) {
#if defined(AC_K2_STAT64)
        typedef unsigned long long stat_t;
#else
        typedef uint32_t stat_t;
#endif
        if (nk) atomicAdd(&s_stat[0], (stat_t)nk);
        if (n_rec) atomicAdd(&s_stat[1], (stat_t)n_rec);
        if (n_kept) atomicAdd(&s_stat[2], (stat_t)n_kept);
        if (short_len) atomicAdd(&s_stat[3], (stat_t)short_len);
        if (n_bases) atomicAdd(&s_stat[4], (stat_t)n_bases);
        if (n_emitted) atomicAdd(&s_stat[5], (stat_t)n_emitted);
        __threadfence_block();
        if (atomicAdd(&s_done, 1u) == (uint32_t)K2_WARPS - 1u) {
            __threadfence_block();
            const volatile stat_t *st = s_stat;
            if (st[0]) atomicAdd(&jd->n_kmers, (unsigned long long)st[0]);
            if (st[1]) atomicAdd(&jd->n_records, (unsigned long long)st[1]);
            if (st[2]) atomicAdd(&jd->n_kept, (unsigned long long)st[2]);
            if (st[3]) atomicAdd(&jd->n_short_len, (unsigned long long)st[3]);
            if (st[4]) atomicAdd(&jd->n_bases, (unsigned long long)st[4]);
            if (st[5]) atomicAdd(&jd->cand_total, (unsigned long long)st[5]);
        }
    }
}

template <int K>
void launch_k(const LaunchCtx &lc, const uint8_t *Z, uint32_t max_zbuf, int ps_slot, JobDev *jd, unsigned long long *cand,
              unsigned long long cand_cap, uint32_t seed, const CountTable &ct, uint32_t cta0, uint32_t cta1) {
    unsigned grid = div_up(max_zbuf, K2_TILE);   // CTAs past the end of the stream leave at once (measured: no cost)
    if (cta1 < grid) grid = cta1;
    if (cta0 >= grid) return;
    grid -= cta0;
#if defined(AC_K2_BRANCHY)   // comparison builds: the branchy loop everywhere
    k2_hash<K, false><<<grid, K2_THREADS, 0, lc.stream>>>(Z, jd, cand, cand_cap, seed, ps_slot, ct, cta0);
#else
    if (ct.mask) k2_hash<K, true><<<grid, K2_THREADS, 0, lc.stream>>>(Z, jd, cand, cand_cap, seed, ps_slot, ct, cta0);
    else k2_hash<K, false><<<grid, K2_THREADS, 0, lc.stream>>>(Z, jd, cand, cand_cap, seed, ps_slot, ct, cta0);
#endif
    ++*lc.launches;
}

}  // namespace

uint32_t hash_ctas_for(uint32_t zbytes) { return div_up(zbytes, K2_TILE); }

#define AC_K_CASE(KK) case KK: launch_k<KK>(lc, Z, max_zbuf, ps_slot, jd, cand, cand_cap, seed, ct, cta0, cta1); break;

bool launch_hash_segment(const LaunchCtx &lc, int k, const uint8_t *Z, uint32_t max_zbuf, int ps_slot, JobDev *jd,
                         unsigned long long *cand, unsigned long long cand_cap, uint32_t seed, int use64,
                         const CountTable &ct, uint32_t cta0, uint32_t cta1) {
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
        default: return false;   // k outside the instantiated set (development builds with AC_ONLY_K27): nothing was launched
    }
    return true;
}

}  // namespace ac
