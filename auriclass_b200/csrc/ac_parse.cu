// ac_parse.cu -- kernel K1: FASTA / FASTQ record parser, single pass.
//
// Replaces Mash's kseq_read() loop (SURVEY.md section 2.2 stages S1-S2, App. A.2)
// that runs inside `mash sketch` (auriclass/classes.py:576-596, :696-713).
// Input: raw inflated text in HBM.  Output: the sanitised compacted base stream Z
// (kmer_core.cuh): sequence characters only, 0x00 after/between records.
//
// FASTA: one kernel (k1_fasta), HBM sees the text once.  FASTQ: see "count -> scan -> mask in place"
// below (no compaction needed).  k1_fasta: a CTA takes the next 8 KiB tile (ticket),
// every thread loads 32 bytes (two coalesced 16-byte loads) and turns them into
// newline / line-start bit masks.  What a tile emits depends on the parser state at
// its first byte (FASTQ: line phase 0..3, FASTA: header or sequence line) and on the
// output offset; both are chained through the tiles with decoupled look-back (one
// 64-bit status word per tile and quantity, no second pass, no scan kernel):
//   stage 1  state   FASTQ: sum of newline counts (mod 4)   FASTA: type of the last line start
//   stage 2  offset  sum of the bytes the earlier tiles emit
// The emitted byte runs are then compacted through shared memory (word-granular
// funnel copies), sanitised word-parallel on the way out and written to Z as aligned
// 16-byte vectors.  The last tile leaves the parser state for the next segment.
//
// FASTQ is parsed as strict 4-line records and *verified* (every record starts
// with '@', its third line with '+', sequence and quality byte totals match, the
// file ends on a record boundary); anything else raises AC_E_FORMAT -- the GPU
// path never guesses.  Multi-line FASTA is handled in full (k-mers span lines).
#include "ac_common.cuh"

namespace ac {

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t bit_range(int a, int b) {  // bits [a, b), 0 <= a <= b <= 32
    uint32_t hi = (b >= 32) ? 0xffffffffu : ((1u << b) - 1u);
    uint32_t lo = (a >= 32) ? 0xffffffffu : ((1u << a) - 1u);
    return hi & ~lo;
}

struct ChunkMasks {
    uint32_t valid, nl;
};

// 0x80 in every byte of w equal to '\n' (exact: no borrow artefacts)
__device__ __forceinline__ uint32_t newline_bits(uint32_t w) {
    const uint32_t x = w ^ 0x0A0A0A0Au;
    const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x | 0x7F7F7F7Fu);
}

// Loads the thread's 32 raw bytes (two 16-byte vector loads) and builds per-byte masks.
template <bool WITH_NL>
__device__ __forceinline__ void load_chunk_t(const SegDesc &sd, uint32_t tile, uint32_t tid, uint32_t w[8],
                                             ChunkMasks &m) {
    const long long pos = (long long)tile * K1_TILE + (long long)tid * K1_CHUNK;
    const long long lo = sd.skip, hi = (long long)sd.skip + sd.len;
    const long long real_hi = hi - (sd.append_nl ? 1 : 0);
    const long long tile_lo = (long long)tile * K1_TILE;
    if (tile_lo >= lo && tile_lo + K1_TILE <= real_hi) {
        // interior tile (all but the first and last of a segment): every byte is text, no range arithmetic per thread
        const uint4 *src = reinterpret_cast<const uint4 *>(sd.base + pos);
        const uint4 x = __ldg(src), y = __ldg(src + 1);
        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w; w[4] = y.x; w[5] = y.y; w[6] = y.z; w[7] = y.w;
        m.valid = 0xffffffffu;
    } else {
#pragma unroll
        for (int v = 0; v < 2; v++) {
            const long long p = pos + 16 * v;
            uint4 x = make_uint4(0, 0, 0, 0);
            if (p < real_hi && p + 16 > lo) x = __ldg(reinterpret_cast<const uint4 *>(sd.base + p));
            w[4 * v + 0] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
        }
        long long a = (lo > pos ? lo : pos) - pos, b = (hi < pos + 32 ? hi : pos + 32) - pos;
        m.valid = (b > a) ? bit_range((int)a, (int)b) : 0u;
        if (sd.append_nl && hi - 1 >= pos && hi - 1 < pos + 32) {
            const int j = (int)(hi - 1 - pos);
#pragma unroll
            for (int jj = 0; jj < 8; jj++)
                if (jj == (j >> 2)) w[jj] = (w[jj] & ~(0xFFu << (8 * (j & 3)))) | (0x0Au << (8 * (j & 3)));
        }
    }
    if (WITH_NL) {
        uint32_t nl = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) nl |= (((newline_bits(w[j]) >> 7) * 0x01020408u) >> 24) << (4 * j);
        m.nl = nl & m.valid;
    } else {
        m.nl = 0;   // the caller has the mask from an earlier pass
    }
}
__device__ __forceinline__ void load_chunk(const SegDesc &sd, uint32_t tile, uint32_t tid, uint32_t w[8], ChunkMasks &m) {
    load_chunk_t<true>(sd, tile, tid, w, m);
}

// Is the byte just before this thread's first valid byte a newline?  (seg_first_ls: value to use
// when that byte is the first byte of the segment.)
__device__ __forceinline__ bool prev_is_newline(const SegDesc &sd, uint32_t tile, const ChunkMasks &m,
                                                uint32_t seg_first_ls, uint32_t *s_last /*[8]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t last_bit = (m.nl >> 31) & 1u;
    const uint32_t up = __shfl_up_sync(FULL, last_bit, 1);
    if (lane == 31) s_last[warp] = last_bit;
    __syncthreads();
    bool r = false;
    if (m.valid) {
        const long long pos = (long long)tile * K1_TILE + (long long)threadIdx.x * K1_CHUNK;
        const int vlo = __ffs(m.valid) - 1;
        if (pos + vlo == (long long)sd.skip) r = seg_first_ls != 0;          // first byte of the segment
        else if (vlo > 0) r = false;                                          // cannot happen (valid starts at bit 0 except at the segment start)
        else if (lane > 0) r = up != 0;
        else if (warp > 0) r = s_last[warp - 1] != 0;
        else r = sd.base[pos - 1] == '\n';                                    // peek into the previous tile
    }
    __syncthreads();
    return r;
}

// four bytes -> upper-case A/C/G/T where valid, 0x01 elsewhere.  Bits 1-2 of a base letter index
// the table "ACTG"; a byte is valid iff the table letter equals its case-folded value.
__device__ __forceinline__ uint32_t sanitise4(uint32_t w) {
    const uint32_t u = w & 0xDFDFDFDFu;
    const uint32_t c = (w >> 1) & 0x03030303u;            // 2-bit index per byte
    const uint32_t t = c | (c >> 4);                      // nibble pairs in bytes 0 and 2
    const uint32_t sel = (t & 0xFFu) | ((t >> 8) & 0xFF00u);
    const uint32_t canon = __byte_perm(0x47544341u, 0u, sel);  // table A,C,T,G
    const uint32_t x = canon ^ u;
    const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);  // 0x80 where valid
    const uint32_t vm = (z >> 7) * 0xFFu;
    return (u & vm) | (~vm & 0x01010101u);
}
// same, and '\n' (the in-stream marker of a record boundary) -> 0x00
__device__ __forceinline__ uint32_t sanitise4_sep(uint32_t w) {
    return sanitise4(w) & ~(newline_bits(w) >> 7);  // newline bytes come out of sanitise4 as 0x01
}

// ---------------------------------------------------------------- decoupled look-back
// status word: [63:34] epoch (segment serial of the context, never reused) [33:32] flag [31:0] value
constexpr uint32_t ST_AGG = 1, ST_INC = 2;
__device__ __forceinline__ u64 st_pack(uint32_t epoch, uint32_t flag, uint32_t v) {
    return ((u64)epoch << 34) | ((u64)flag << 32) | v;
}
__device__ __forceinline__ u64 ld_status(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

enum { OP_SUM = 0, OP_LAST = 1 };  // OP_LAST: value bit0 = defined, bit1 = payload; the nearest defined value wins

// Called by one full warp.  Publishes the tile's aggregate, walks back over the earlier tiles until an
// inclusive prefix is met, publishes the tile's own inclusive prefix, returns the exclusive one.
// `init` is the exclusive prefix of tile 0 (for OP_LAST it must have bit0 set).
template <int OP>
__device__ uint32_t lookback(u64 *st, uint32_t tile, uint32_t epoch, uint32_t agg, uint32_t init) {
    const int lane = threadIdx.x & 31;
    uint32_t ex;
    if (tile == 0) {
        ex = init;
    } else {
        if (lane == 0) st_status(&st[tile], st_pack(epoch, ST_AGG, agg));
        uint32_t acc = 0;  // OP_SUM: sum of the aggregates walked over so far
        bool done = false;
        ex = 0;
        for (long long base = (long long)tile - 1; !done; base -= 32) {
            const long long idx = base - lane;
            uint32_t flag, v;
            if (idx >= 0) {
                u64 w;
                do {
                    w = ld_status(&st[idx]);
                    flag = (uint32_t)(w >> 32) & 3u;
                } while ((uint32_t)(w >> 34) != epoch || flag == 0);
                v = (uint32_t)w;
            } else {
                flag = ST_INC;  // virtual tile -1 carries the initial state
                v = init;
            }
            const unsigned inc_mask = __ballot_sync(FULL, flag == ST_INC);
            const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;  // nearest inclusive prefix in this window
            if (OP == OP_SUM) {
                uint32_t c = (lane <= first_inc) ? v : 0u;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(FULL, c, d);
                acc += c;
                if (first_inc < 32) { ex = acc; done = true; }
            } else {
                const unsigned def_mask = __ballot_sync(FULL, lane <= first_inc && (v & 1u));
                if (def_mask) { ex = __shfl_sync(FULL, v, __ffs(def_mask) - 1); done = true; }
                // an inclusive prefix is always defined, so first_inc < 32 implies def_mask != 0
            }
        }
    }
    uint32_t inc;
    if (OP == OP_SUM) inc = ex + agg; else inc = (agg & 1u) ? agg : ex;
    if (lane == 0) st_status(&st[tile], st_pack(epoch, ST_INC, inc));
    return ex;
}

// ---------------------------------------------------------------- compaction through shared memory
constexpr int SW_STRIDE = 9;  // words per thread in s_words (8 + 1 pad: conflict-free column access)

// copy the thread's emitted byte runs (mask em over its 32-byte chunk) to s_out starting at byte d
__device__ __forceinline__ void compact_runs(const uint32_t *s_words, uint8_t *s_out, uint32_t tid, uint32_t em, uint32_t d) {
    const uint32_t *src_w = s_words + tid * SW_STRIDE;
    const uint8_t *src_b = reinterpret_cast<const uint8_t *>(src_w);
    while (em) {
        const int a = __ffs(em) - 1;
        const uint32_t inv = ~em >> a;                 // bit 0 is clear; first set bit = run length
        const int len = inv ? (__ffs(inv) - 1) : (32 - a);
        int sa = a, n = len;
        // head: bytes until the destination is word aligned
        while (n > 0 && (d & 3u)) { s_out[d++] = src_b[sa++]; n--; }
        if (n >= 4) {
            const int sh = (sa & 3) * 8;
            int wi = sa >> 2;
            uint32_t lo = src_w[wi];
            uint32_t *dst = reinterpret_cast<uint32_t *>(s_out + d);
            const int nw = n >> 2;
            for (int q = 0; q < nw; q++) {
                const uint32_t hi = src_w[wi + 1];     // wi + 1 <= 8: the pad word, shifted out when sh == 0
                dst[q] = __funnelshift_r(lo, hi, sh);
                lo = hi; wi++;
            }
            d += 4 * nw; sa += 4 * nw; n -= 4 * nw;
        }
        while (n > 0) { s_out[d++] = src_b[sa++]; n--; }
        em = (a + len >= 32) ? 0u : (em & ~((1u << (a + len)) - 1u));
    }
}

// s_out[s_off .. s_off + cnt) -> Z[out_off ..), sanitised; aligned 16-byte vectors inside, bytes at the edges
__device__ __forceinline__ void write_out(const uint8_t *s_out, uint32_t s_off, uint32_t cnt, uint8_t *Z, uint32_t out_off) {
    uint8_t *zal = Z + (out_off - s_off);  // 16-byte aligned: s_off = out_off & 15
    const uint32_t end = s_off + cnt;
    const uint32_t nvec = (end + 15) >> 4;
    for (uint32_t i = threadIdx.x; i < nvec; i += K1_THREADS) {
        const uint32_t b0 = 16 * i;
        uint4 v = reinterpret_cast<const uint4 *>(s_out)[i];
        v.x = sanitise4_sep(v.x); v.y = sanitise4_sep(v.y); v.z = sanitise4_sep(v.z); v.w = sanitise4_sep(v.w);
        if (b0 >= s_off && b0 + 16 <= end) {
            reinterpret_cast<uint4 *>(zal)[i] = v;
        } else {
            const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t b = b0 + j;
                if (b >= s_off && b < end) zal[b] = (uint8_t)(ww[j >> 2] >> (8 * (j & 3)));
            }
        }
    }
}

// stream carry: the last Z_CARRY bytes of the previous segment's stream go in front of this one's
__device__ __forceinline__ void copy_carry(const SegDesc &sd, const ParseState &ps, const uint8_t *Zprev, uint8_t *Z) {
    if (threadIdx.x < Z_CARRY) {
        uint8_t c = 0;
        if (!sd.first_of_stream) c = Zprev[ps.n_zbuf - Z_CARRY + threadIdx.x];
        Z[threadIdx.x] = c;
    }
}

struct TileShared {
    uint32_t w[33];
    uint32_t last[8];
    uint32_t acc[6];
    uint32_t tile, state_in, out_off;
};

// ---------------------------------------------------------------- FASTQ line classification
// Line classification is closed form per 32-byte chunk: two prefix-xor scans over the newline mask give every byte's
// line index mod 4 (no per-line loop, no divergence).

// exclusive prefix parity: bit i of the result = parity of popc(x & ((1 << i) - 1))
__device__ __forceinline__ uint32_t prefix_xor_excl(uint32_t x) {
    x <<= 1;
    x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
    return x;
}

struct LineMasks {
    uint32_t p0, p1;  // bit i: bit 0 / bit 1 of (number of newlines in bytes [0, i) of the chunk)
};
__device__ __forceinline__ LineMasks line_index_masks(uint32_t nl) {
    LineMasks lm;
    lm.p0 = prefix_xor_excl(nl);
    lm.p1 = prefix_xor_excl(nl & lm.p0);  // bit 1 flips at a newline that follows an odd count
    return lm;
}
// bytes of the chunk that lie in lines whose index (relative to the chunk's first byte) is r (mod 4)
__device__ __forceinline__ uint32_t lines_of(const LineMasks &lm, uint32_t r, uint32_t valid) {
    return ((r & 1u) ? lm.p0 : ~lm.p0) & ((r & 2u) ? lm.p1 : ~lm.p1) & valid;
}

// ---------------------------------------------------------------- FASTQ, one pass: k1q_fused
// The text is read ONCE (TMA bulk copy of a 16 KiB tile into shared memory) and the base stream written once.
// What a tile emits depends on the line phase at its first byte and on its output offset; both are chained through
// the tiles by decoupled look-back over ONE 64-bit status word per tile:
//   aggregate  [63:62] = 1 | [61:60] newlines (mod 4) | 4 x 15 bits: bytes in lines of tile-local index 0..3 (mod 4)
//   inclusive  [63:62] = 2 | [61:60] phase after the tile | [31:0] output offset after the tile
// Aggregates compose ((nA, cA) then (nB, cB) = (nA + nB, cA[q] + cB[q - nA])), so a warp folds a window of 32
// predecessors with one scan, a 64-bit rotate per lane (the four 16-bit class counts rotate as one register) and four
// warp reductions.  The status words are cleared per segment by the launcher (no epoch tag fits beside 4 x 15 bits).
// The aggregate goes out as soon as the tile is counted; the walk itself comes LATE, after the emit side: the phase the
// emit side needs is first read off the text (a line that starts with '@' and whose next-but-one line starts with '+' is
// a header) and checked against the chained value afterwards -- a wrong or missing guess costs a second emit pass, never
// a wrong stream.
// Output side: the tile's sequence lines ("runs": maximal stretches of emitted bytes, newline included) are listed in
// shared memory as (output offset, source offset); every thread then PRODUCES 16-byte vectors of the stream --
// locates the run of its first byte (proportional guess + short walk), gathers 16 unaligned source bytes from the
// staged text (and 16 more from the next run when the vector straddles a record boundary), sanitises 4 bytes at a time
// -- into a staging area at tile-local alignment; once the output offset is known the staged stream goes to Z as
// aligned 16-byte vectors.  No lane idles on header / quality bytes and there is no per-thread byte loop.
// NT threads per CTA, tile = 64 NT bytes (thread t owns the 32-byte chunks t and NT + t).  Measured per C2 isolate with the
// look-back in FRONT of the emit side: 256 threads / 16 KiB at 6 CTAs per SM (40 registers) 1.29 ms, at 5 CTAs 1.40 ms,
// 128 threads / 8 KiB 1.55 ms; 64-tile windows 1.43 ms; all eight warps walking 1.52 ms.  With the late look-back (this
// kernel: 48 registers, 37.5 KB of shared memory, 5 CTAs per SM): 1.25 ms.
#ifndef AC_K1_NT
#define AC_K1_NT 256
#endif
#ifndef AC_K1_MINB
#define AC_K1_MINB 5
#endif
constexpr int FQ_NT = AC_K1_NT, FQ_MINB = AC_K1_MINB;
constexpr int FQ_MAXRUNS = 1024;                      // runs listed per tile; beyond that (reads of < ~12 bytes) a byte-wise path
constexpr int FQ_PAD = 32;                            // slack in front of / behind the staged text (unaligned gathers)
constexpr uint32_t FQS_AGG = 1, FQS_INC = 2;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NT>
__device__ __forceinline__ void block_excl_sum2(uint32_t &a, uint32_t &b, uint32_t *s_w /*[66]*/, uint32_t &tot_a, uint32_t &tot_b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ta = __shfl_up_sync(FULL, ia, d), tb = __shfl_up_sync(FULL, ib, d);
        if (lane >= d) { ia += ta; ib += tb; }
    }
    if (lane == 31) { s_w[warp] = ia; s_w[33 + warp] = ib; }
    __syncthreads();
    if (warp == 0) {
        uint32_t xa = (lane < NT / 32) ? s_w[lane] : 0u, xb = (lane < NT / 32) ? s_w[33 + lane] : 0u, ya = xa, yb = xb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ta = __shfl_up_sync(FULL, ya, d), tb = __shfl_up_sync(FULL, yb, d);
            if (lane >= d) { ya += ta; yb += tb; }
        }
        s_w[lane] = ya - xa; s_w[33 + lane] = yb - xb;
        if (lane == 31) { s_w[32] = ya; s_w[65] = yb; }
    }
    __syncthreads();
    a = ia - a + s_w[warp]; b = ib - b + s_w[33 + warp];
    tot_a = s_w[32]; tot_b = s_w[65];
    __syncthreads();
}

template <int NT>
struct FqShared {
    __align__(128) uint8_t text[FQ_PAD + 64 * NT + FQ_PAD];   // staged tile at text + FQ_PAD (TMA destination)
    __align__(16) uint8_t out[FQ_PAD + 64 * NT + FQ_PAD];     // the tile's base stream at tile-local alignment (out + FQ_PAD)
    uint16_t run_out[FQ_MAXRUNS + 2], run_src[FQ_MAXRUNS + 2];
    uint16_t nlpos[16];                                       // positions of the tile's first newlines (phase from the text)
    uint32_t w[66];
    uint32_t acc[4];
    uint32_t last_nl[16], last_em[16];
    uint32_t p_cur, p_true, o_true;
    __align__(8) uint64_t bar;
};

// 16 source bytes starting at byte offset s of the staged text (s in [-FQ_PAD, FQ_TILE + FQ_PAD - 20]), as 4 words
__device__ __forceinline__ void gather16(const uint8_t *text0, int s, uint32_t out[4]) {
    const uint32_t addr = smem_addr(text0) + (uint32_t)s;
    const uint32_t al = addr & ~3u, sh = (addr & 3u) * 8u;
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; i++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[i]) : "r"(al + 4u * i));
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = __funnelshift_r(v[i], v[i + 1], sh);
}

// The look-back of one tile (one full warp): walks back over the predecessors' status words, 32 per round trip, until an
// inclusive prefix is met; returns the phase at the tile's first byte and its output offset.  Virtual tile -1 carries the
// parser state the previous segment left.
__device__ __forceinline__ void fq_lookback(const u64 *__restrict__ status, uint32_t tile, uint32_t phase0, uint32_t lane,
                                            uint32_t &P, uint32_t &O) {
    P = 0; O = 0;
    uint32_t S_nl = 0, Sc[4] = {0, 0, 0, 0};     // composition of the tiles behind the current window, up to tile - 1
    for (long long top = (long long)tile - 1;; top -= 32) {
        const long long idx = top - 31 + lane;    // ascending with the lane
        u64 w = 0;
        if (idx >= 0) { do { w = ld_status(&status[idx]); } while ((w >> 62) == 0); }
        else if (idx == -1) w = ((u64)FQS_INC << 62) | ((u64)(phase0 & 3u) << 60) | (u64)Z_CARRY;   // virtual tile -1
        const uint32_t flag = (uint32_t)(w >> 62);
        const unsigned inc_mask = __ballot_sync(FULL, flag == FQS_INC);
        const int F = inc_mask ? 31 - __clz(inc_mask) : -1;   // nearest inclusive prefix of this window
        const bool use = (int)lane > F && flag == FQS_AGG;
        const uint32_t my_nl = use ? (uint32_t)(w >> 60) & 3u : 0u;
        uint32_t pre = my_nl;                                 // newlines of the used lanes before this one
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, pre, d); if ((int)lane >= d) pre += t; }
        const uint32_t W_nl = __shfl_sync(FULL, pre, 31);
        pre -= my_nl;
        // class counts as 4 x 16 bits; rotating left by 16 * pre re-indexes them to the window's first used tile
        u64 x = 0;
        if (use) x = (w & 0x7FFFu) | (((w >> 15) & 0x7FFFu) << 16) | (((w >> 30) & 0x7FFFu) << 32) | (((w >> 45) & 0x7FFFu) << 48);
        const uint32_t rot = 16u * (pre & 3u);
        x = rot ? ((x << rot) | (x >> (64u - rot))) : x;
        const uint32_t W0 = __reduce_add_sync(FULL, (uint32_t)x & 0xFFFFu), W1 = __reduce_add_sync(FULL, (uint32_t)(x >> 16) & 0xFFFFu);
        const uint32_t W2 = __reduce_add_sync(FULL, (uint32_t)(x >> 32) & 0xFFFFu), W3 = __reduce_add_sync(FULL, (uint32_t)(x >> 48) & 0xFFFFu);
        // window first, then what was folded before: c[q] = W[q] + S[q - W_nl]
        const uint32_t Wc[4] = {W0, W1, W2, W3};
        uint32_t N[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t from = (q - W_nl) & 3u;
            N[q] = Wc[q] + (from == 0 ? Sc[0] : from == 1 ? Sc[1] : from == 2 ? Sc[2] : Sc[3]);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) Sc[q] = N[q];
        S_nl += W_nl;
        if (inc_mask) {
            const u64 wi = __shfl_sync(FULL, w, F);
            const uint32_t Pi = (uint32_t)(wi >> 60) & 3u, Oi = (uint32_t)wi;
            const uint32_t seq = (1u - Pi) & 3u;      // tile-local line index of the sequence lines right behind that tile
            P = (Pi + S_nl) & 3u;
            O = Oi + (seq == 0 ? Sc[0] : seq == 1 ? Sc[1] : seq == 2 ? Sc[2] : Sc[3]);
            break;
        }
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k1q_fused(SegDesc sd, uint32_t n_tiles, u64 *__restrict__ status, JobDev *__restrict__ jd,
                                                      const uint8_t *__restrict__ Zprev, uint8_t *__restrict__ Z) {
    constexpr int FQ_THREADS = NT, FQ_TILE = 64 * NT, FQ_WARPS = NT / 32;
    __shared__ FqShared<NT> sh;
    const uint32_t tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *text0 = sh.text + FQ_PAD;
    const long long tile_lo = (long long)tile * FQ_TILE;
    const long long seg_lo = sd.skip, seg_hi = (long long)sd.skip + sd.len, real_hi = seg_hi - (sd.append_nl ? 1 : 0);
    const ParseState ps = jd->ps[sd.serial & 1u];

    // ---- stage the tile: one TMA bulk copy (whole 16-byte vectors that hold text; the allocation granule covers them)
    long long want = real_hi - tile_lo;
    want = want < 0 ? 0 : (want > FQ_TILE ? FQ_TILE : want);
    const uint32_t nbytes = ((uint32_t)want + 15u) & ~15u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sh.bar)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (nbytes) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&sh.bar)), "r"(nbytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr(text0)), "l"(sd.base + tile_lo), "r"(nbytes), "r"(smem_addr(&sh.bar)) : "memory");
        }
    }
    if (tid < 4) sh.acc[tid] = 0;
    if (tile == 0) copy_carry(sd, ps, Zprev, Z);
    __syncthreads();
    if (nbytes) {
        uint32_t ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(smem_addr(&sh.bar)), "r"(0u) : "memory");
        } while (!ok);
    }
    if (sd.append_nl && seg_hi - 1 >= tile_lo && seg_hi - 1 < tile_lo + FQ_TILE && tid == 0) text0[seg_hi - 1 - tile_lo] = '\n';
    const bool interior = tile_lo >= seg_lo && tile_lo + FQ_TILE <= real_hi;
    if (!interior) __syncthreads();   // the patched byte

    // ---- per chunk: newline mask, validity; rows r = 0, 1: chunk ch = 256 r + tid
    uint32_t nl[2], valid[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint32_t ch = (uint32_t)NT * r + tid;
        const uint4 x = reinterpret_cast<const uint4 *>(text0)[2 * ch], y = reinterpret_cast<const uint4 *>(text0)[2 * ch + 1];
        const uint32_t w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) m |= (((newline_bits(w[j]) >> 7) * 0x01020408u) >> 24) << (4 * j);
        uint32_t v = 0xffffffffu;
        if (!interior) {
            const long long pos = tile_lo + 32ll * ch;
            const long long a = (seg_lo > pos ? seg_lo : pos) - pos, b = (seg_hi < pos + 32 ? seg_hi : pos + 32) - pos;
            v = (b > a) ? bit_range((int)a, (int)b) : 0u;
        }
        valid[r] = v; nl[r] = m & v;
    }
    // newlines before each chunk inside the tile (both rows in one scan; row 1 follows row 0)
    uint32_t L0a = __popc(nl[0]), L0b = __popc(nl[1]), tot_nl0, tot_nl1;
    block_excl_sum2<FQ_THREADS>(L0a, L0b, sh.w, tot_nl0, tot_nl1);
    const uint32_t L0[2] = {L0a, L0b + tot_nl0};
    const uint32_t tile_nl = tot_nl0 + tot_nl1;
    // bytes per tile-local line index (mod 4)
    LineMasks lm[2];
    {
        uint32_t packed = 0;                   // byte q: bytes of the thread's chunks in lines of tile-local index q (<= 64)
#pragma unroll
        for (int r = 0; r < 2; r++) {
            lm[r] = line_index_masks(nl[r]);
            const uint32_t p0 = lm[r].p0, p1 = lm[r].p1, v = valid[r];
            // chunk-local classes 0..3, one count per byte; chunk-local index i is tile-local index (i + L0) & 3:
            // rotate the four bytes left by L0 bytes
            const uint32_t c = __popc(~p0 & ~p1 & v) | (__popc(p0 & ~p1 & v) << 8) | (__popc(~p0 & p1 & v) << 16) | (__popc(p0 & p1 & v) << 24);
            const uint32_t rot = 8u * (L0[r] & 3u);
            packed += __funnelshift_l(c, c, rot);
        }
        uint32_t lo = (packed & 0xFFu) | ((packed & 0xFF00u) << 8), hi = ((packed >> 16) & 0xFFu) | ((packed >> 24) << 16);   // <= 64 * 32 per 16-bit field and warp
        lo = __reduce_add_sync(FULL, lo); hi = __reduce_add_sync(FULL, hi);
        if (lane == 0) {
            atomicAdd(&sh.acc[0], lo & 0xFFFFu); atomicAdd(&sh.acc[1], lo >> 16);
            atomicAdd(&sh.acc[2], hi & 0xFFFFu); atomicAdd(&sh.acc[3], hi >> 16);
        }
    }
    // last-byte flags of every chunk (the next chunk's "previous byte" questions)
    {
        const uint32_t b0 = __ballot_sync(FULL, nl[0] >> 31), b1 = __ballot_sync(FULL, nl[1] >> 31);
        if (lane == 0) { sh.last_nl[warp] = b0; sh.last_nl[FQ_WARPS + warp] = b1; }
    }
    __syncthreads();

    // ---- aggregate out at once; the look-back itself comes LATE (after the emit side), when the predecessors' aggregates have
    // been out for a while.  The phase the emit side needs is first read off the text itself -- in 4-line FASTQ a line that
    // starts with '@' and whose next-but-one line starts with '+' is a header -- and checked against the chained value
    // afterwards (a wrong or missing guess costs a second emit pass, never a wrong stream).  With the look-back in front of
    // the emit side 37 % of all stall samples sat at the barrier behind it: the walk waits for the YOUNGEST predecessors'
    // aggregates, which are produced at the same time as this tile's (wider windows made it worse: 1.43 / 1.52 ms).
    const uint32_t c0 = sh.acc[0], c1 = sh.acc[1], c2 = sh.acc[2], c3 = sh.acc[3];
    if (warp == 0) {
        if (tile > 0 && lane == 0) {
            const u64 agg = ((u64)FQS_AGG << 62) | ((u64)(tile_nl & 3u) << 60) | ((u64)c3 << 45) | ((u64)c2 << 30) | ((u64)c1 << 15) | (u64)c0;
            st_status(&status[tile], agg);
        }
        uint32_t p_spec = 0xFFu;
        if (tile == 0) p_spec = ps.fq_phase & 3u;      // (exact)
        else {
            // positions of the first newlines of the tile (warp 0 owns chunks 0..31: its first KiB)
            const uint32_t cnt = __popc(nl[0]);
            uint32_t inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, d); if ((int)lane >= d) inc += t; }
            uint32_t kk = inc - cnt;
            for (uint32_t x = nl[0]; x && kk < 16u; x &= x - 1, kk++) sh.nlpos[kk] = (uint16_t)(32u * lane + (uint32_t)(__ffs(x) - 1));
            const uint32_t m = min(16u, __shfl_sync(FULL, inc, 31));
            __syncwarp();
            bool hit = false;
            if (lane + 2 < m) {
                const uint32_t h = sh.nlpos[lane] + 1u, q = sh.nlpos[lane + 2] + 1u;
                hit = (long long)q < want && text0[h] == '@' && text0[q] == '+';
            }
            const unsigned hm = __ballot_sync(FULL, hit);
            if (hm) p_spec = (0u - (uint32_t)__ffs(hm)) & 3u;     // the line behind newline k is a header: P + k + 1 = 0 (mod 4)
        }
        if (lane == 0) sh.p_cur = p_spec;
    }
    __syncthreads();
    uint32_t P = sh.p_cur, O = 0, n_out = 0;
    bool lb_done = false;
    uint32_t emit[2] = {0, 0}, err = 0, E[2] = {0, 0}, n_runs = 0;
    uint8_t *out0 = sh.out + FQ_PAD;
#pragma unroll 1
    for (int attempt = 0; attempt < 3; attempt++) {
        if (P == 0xFFu || attempt > 0) {
            // no phase from the text (lines longer than the first KiB), or the guess was wrong: the chained value, now
            if (!lb_done) {
                if (warp == 0) {
                    uint32_t Pt, Ot;
                    if (tile == 0) { Pt = ps.fq_phase & 3u; Ot = Z_CARRY; } else fq_lookback(status, tile, ps.fq_phase, lane, Pt, Ot);
                    if (lane == 0) { sh.p_true = Pt; sh.o_true = Ot; }
                }
                __syncthreads();
                lb_done = true;
            }
            P = sh.p_true;
        }
        n_out = ((1u - P) & 3u) == 0 ? c0 : ((1u - P) & 3u) == 1 ? c1 : ((1u - P) & 3u) == 2 ? c2 : c3;

        // ---- emitted bytes (sequence lines incl. their newline), format checks, run list
        err = 0;
    #pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint32_t ch = (uint32_t)NT * r + tid;
            const uint32_t ph = (P + L0[r]) & 3u;                     // phase of the line holding the chunk's first byte
            emit[r] = lines_of(lm[r], (1u - ph) & 3u, valid[r]);
            if (valid[r]) {
                // is the byte in front of the chunk's first valid byte a newline?
                bool prev_nl;
                const long long pos = tile_lo + 32ll * ch;
                const int vlo = __ffs(valid[r]) - 1;
                if (pos + vlo == seg_lo) prev_nl = ps.at_line_start != 0;                 // first byte of the segment
                else if (ch == 0) prev_nl = sd.base[tile_lo - 1] == '\n';                 // last byte of the previous tile
                else if (lane > 0) prev_nl = (sh.last_nl[FQ_WARPS * r + warp] >> (lane - 1)) & 1u;
                else if (warp > 0) prev_nl = sh.last_nl[FQ_WARPS * r + warp - 1] >> 31;
                else prev_nl = sh.last_nl[FQ_WARPS - 1] >> 31;                                       // row 1 starts behind row 0's last chunk
                const uint32_t ls = ((nl[r] << 1) | (prev_nl ? (1u << vlo) : 0u)) & valid[r];
                if (ls) {
                    const uint32_t hdr = lines_of(lm[r], (0u - ph) & 3u, valid[r]), plus = lines_of(lm[r], (2u - ph) & 3u, valid[r]);
                    for (uint32_t x = ls & (hdr | plus); x; x &= x - 1) {
                        const int b = __ffs(x) - 1;
                        const uint8_t c = text0[32u * ch + b];
                        if ((hdr >> b) & 1u) { if (c != '@') err |= FERR_FQ_HEADER; }
                        else if (c != '+') err |= FERR_FQ_PLUS;
                    }
                }
            }
        }
        {
            const uint32_t b0 = __ballot_sync(FULL, emit[0] >> 31), b1 = __ballot_sync(FULL, emit[1] >> 31);
            if (lane == 0) { sh.last_em[warp] = b0; sh.last_em[FQ_WARPS + warp] = b1; }
        }
        __syncthreads();
        uint32_t starts[2];
    #pragma unroll
        for (int r = 0; r < 2; r++) {
            uint32_t prev_em;      // is the byte in front of the chunk emitted?  (a tile always starts a new run)
            if (lane > 0) prev_em = (sh.last_em[FQ_WARPS * r + warp] >> (lane - 1)) & 1u;
            else if (warp > 0) prev_em = sh.last_em[FQ_WARPS * r + warp - 1] >> 31;
            else prev_em = r ? sh.last_em[FQ_WARPS - 1] >> 31 : 0u;
            starts[r] = emit[r] & ~((emit[r] << 1) | prev_em);
        }
        // exclusive prefixes over the chunks in stream order: emitted bytes, run starts (rows packed 16 | 16)
        uint32_t eb = __popc(emit[0]) | (__popc(emit[1]) << 16), rs = __popc(starts[0]) | (__popc(starts[1]) << 16), tot_e, tot_r;
        block_excl_sum2<FQ_THREADS>(eb, rs, sh.w, tot_e, tot_r);
        E[0] = eb & 0xFFFFu; E[1] = (eb >> 16) + (tot_e & 0xFFFFu);
        const uint32_t R[2] = {rs & 0xFFFFu, (rs >> 16) + (tot_r & 0xFFFFu)};
        n_runs = (tot_r & 0xFFFFu) + (tot_r >> 16);
    #pragma unroll
        for (int r = 0; r < 2; r++) {
            uint32_t k = R[r];
            for (uint32_t x = starts[r]; x; x &= x - 1, k++) {
                const int b = __ffs(x) - 1;
                if (k < (uint32_t)FQ_MAXRUNS) {
                    sh.run_out[k] = (uint16_t)(E[r] + __popc(emit[r] & ((1u << b) - 1u)));
                    sh.run_src[k] = (uint16_t)(32u * ((uint32_t)NT * r + tid) + b);
                }
            }
        }
        __syncthreads();

        // ---- produce the stream at tile-local alignment into the staging area (the output offset is not known yet)
        if (n_out) {
            if (n_runs <= (uint32_t)FQ_MAXRUNS) {
                const uint32_t v_end = (n_out + 15u) >> 4;
                const float runs_per_byte = (float)n_runs / (float)n_out;
                for (uint32_t v = tid; v < v_end; v += FQ_THREADS) {
                    const int o0 = (int)(v * 16u);                       // tile-local output offset of the vector's first byte
                    const uint32_t key = o0 < 0 ? 0u : (uint32_t)o0;
                    // last run with run_out <= key: reads of one length make runs of one length, so the proportional guess is
                    // right or one off; the walk behind it is exact for any input
                    uint32_t lo = min(n_runs - 1u, (uint32_t)((float)key * runs_per_byte));
                    while (sh.run_out[lo] > key) lo--;
                    while (lo + 1 < n_runs && sh.run_out[lo + 1] <= key) lo++;
                    const int avail = (lo + 1 < n_runs ? (int)sh.run_out[lo + 1] : 0x7fffffff) - o0;   // bytes of the vector served by run lo
                    uint32_t wv[4];
                    gather16(text0, (int)sh.run_src[lo] + (o0 - (int)sh.run_out[lo]), wv);
                    if (avail < 16) {
                        const int avail2 = (lo + 2 < n_runs ? (int)sh.run_out[lo + 2] : 0x7fffffff) - o0;
                        if (avail2 >= 16) {
                            uint32_t w2[4];
                            gather16(text0, (int)sh.run_src[lo + 1] - avail, w2);    // byte i >= avail is byte i - avail of run lo + 1
    #pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const int na = avail - 4 * i;                        // bytes of word i that still belong to run lo
                                const uint32_t m = na >= 4 ? 0xffffffffu : (na <= 0 ? 0u : ((1u << (8 * na)) - 1u));
                                wv[i] = (wv[i] & m) | (w2[i] & ~m);
                            }
                        } else {
                            // three or more runs inside 16 bytes (reads of a few bases): byte by byte
                            uint32_t rr = lo;
    #pragma unroll
                            for (int wi = 0; wi < 4; wi++) {
                                uint32_t acc = 0;
    #pragma unroll
                                for (int bi = 0; bi < 4; bi++) {
                                    const int o = o0 + 4 * wi + bi;
                                    while (rr + 1 < n_runs && (int)sh.run_out[rr + 1] <= o) rr++;
                                    const int sidx = (int)sh.run_src[rr] + (o - (int)sh.run_out[rr]);
                                    const uint32_t byte = (sidx >= -FQ_PAD && sidx < FQ_TILE + FQ_PAD) ? text0[sidx] : 0u;
                                    acc |= byte << (8 * bi);
                                }
                                wv[wi] = acc;
                            }
                        }
                    }
    #pragma unroll
                    for (int i = 0; i < 4; i++) wv[i] = sanitise4_sep(wv[i]);
                    reinterpret_cast<uint4 *>(out0)[v] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                }
            } else {
                // more runs than the list holds: every thread writes its own emitted bytes, one at a time
    #pragma unroll 1
                for (int r = 0; r < 2; r++) {
                    uint32_t d = E[r];
                    for (uint32_t x = emit[r]; x; x &= x - 1, d++) {
                        const int b = __ffs(x) - 1;
                        const uint32_t c = text0[32u * ((uint32_t)NT * r + tid) + b];
                        out0[d] = (uint8_t)(sanitise4_sep(c | 0x41414100u) & 0xFFu);
                    }
                }
            }
        }


        // ---- the look-back, late
        __syncthreads();
        if (!lb_done) {
            if (warp == 0) {
                uint32_t Pt, Ot;
                if (tile == 0) { Pt = ps.fq_phase & 3u; Ot = Z_CARRY; } else fq_lookback(status, tile, ps.fq_phase, lane, Pt, Ot);
                if (lane == 0) { sh.p_true = Pt; sh.o_true = Ot; }
            }
            __syncthreads();
            lb_done = true;
        }
        if (sh.p_true == P) break;
        __syncthreads();      // (wrong guess: once more with the chained phase; the run list and the staging area are rewritten)
    }
    O = sh.o_true;
    if (err) atomicOr(&jd->format_error, err);
    if (tid == 0) {
        const uint32_t qual = (3u - P) & 3u;
        const uint32_t q_bytes = qual == 0 ? c0 : qual == 1 ? c1 : qual == 2 ? c2 : c3;
        st_status(&status[tile], ((u64)FQS_INC << 62) | ((u64)((P + tile_nl) & 3u) << 60) | (u64)(O + n_out));
        if (n_out) atomicAdd(&jd->fq_seq_bytes, (u64)n_out);
        if (q_bytes) atomicAdd(&jd->fq_qual_bytes, (u64)q_bytes);
    }
    // ---- staged stream -> Z[O .. O + n_out): aligned 16-byte vectors of Z, each from 16 unaligned staged bytes
    if (n_out) {
        const uint32_t v_first = O >> 4, v_end = (O + n_out + 15u) >> 4;
        for (uint32_t v = v_first + tid; v < v_end; v += FQ_THREADS) {
            const int o0 = (int)(v * 16u) - (int)O;
            uint32_t wv[4];
            gather16(out0, o0, wv);
            uint8_t *dst = Z + (size_t)v * 16u;
            if (o0 >= 0 && (uint32_t)o0 + 16u <= n_out) {
                *reinterpret_cast<uint4 *>(dst) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int o = o0 + i;
                    if (o >= 0 && (uint32_t)o < n_out) dst[i] = (uint8_t)(wv[i >> 2] >> (8 * (i & 3)));
                }
            }
        }
    }

    // ---- last tile: parser state for the next segment
    if (tile + 1 == n_tiles && tid == 0) {
        const uint32_t end_state = (P + tile_nl) & 3u;
        uint32_t ends_nl = ps.at_line_start;
        if (seg_hi > seg_lo) ends_nl = text0[seg_hi - 1 - tile_lo] == '\n';
        ParseState nx;
        nx.z_begin = Z_CARRY;
        nx.n_zbuf = O + n_out;
        if (sd.last_of_file) {
            if (end_state != 0 || !ends_nl) atomicOr(&jd->format_error, (uint32_t)FERR_FQ_TRUNC);
            nx.fq_phase = 0; nx.fa_type = 0; nx.at_line_start = 1;
        } else {
            nx.fq_phase = end_state; nx.fa_type = ps.fa_type; nx.at_line_start = ends_nl;
        }
        jd->ps[(sd.serial + 1u) & 1u] = nx;
    }
}

// ---------------------------------------------------------------- FASTA
struct FaThread {
    uint32_t emit;       // bytes emitted whatever the incoming line type (lines that start inside the chunk, and
                         // the chunk's head when its type is known from an earlier line start of the tile)
    uint32_t head;       // bytes of the chunk's head line, emitted iff the incoming type is "sequence"
    uint32_t sepm;       // header line starts ('>' / '@'): emitted as a record separator
    uint32_t err;
};

__device__ __forceinline__ FaThread fa_classify(const ChunkMasks &m, uint32_t lsm, const uint8_t *bytes, bool known, uint32_t type) {
    FaThread o{0, 0, 0, 0};
    if (!m.valid) return o;
    const int vlo = __ffs(m.valid) - 1, vhi = 32 - __clz(m.valid);
    int pos = vlo;
    while (true) {
        if ((lsm >> pos) & 1u) {
            const uint8_t ch = bytes[pos];
            type = (ch == '>' || ch == '@') ? 0u : 1u;
            known = true;
            if (ch == '+') o.err |= FERR_FA_PLUS;
            if (type == 0u) { o.emit |= 1u << pos; o.sepm |= 1u << pos; }
        }
        const uint32_t next = (pos + 1 < 32) ? (lsm & (0xffffffffu << (pos + 1))) : 0u;
        const int end = next ? (__ffs(next) - 1) : vhi;  // exclusive
        const uint32_t body = bit_range(pos, end) & ~m.nl & m.valid;
        if (!known) o.head |= body;
        else if (type == 1u) o.emit |= body;
        if (!next) break;
        pos = end;
    }
    return o;
}

__device__ __forceinline__ uint32_t fa_line_starts(const ChunkMasks &m, bool prev_nl) {
    if (!m.valid) return 0u;
    const int vlo = __ffs(m.valid) - 1;
    return ((m.nl << 1) | (prev_nl ? (1u << vlo) : 0u)) & m.valid;
}

__global__ void __launch_bounds__(K1_THREADS) k1_fasta(SegDesc sd, JobDev *__restrict__ jd, u64 *__restrict__ st1,
                                                       u64 *__restrict__ st2, const uint8_t *__restrict__ Zprev,
                                                       uint8_t *__restrict__ Z) {
    __shared__ TileShared sh;
    __shared__ __align__(16) uint32_t s_words[K1_THREADS * SW_STRIDE + 4];
    __shared__ __align__(16) uint8_t s_out[K1_TILE + 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sh.tile = (uint32_t)(atomicAdd(&jd->tile_ticket, 1u) - sd.ticket_base);
    if (tid < 6) sh.acc[tid] = 0;
    __syncthreads();
    const uint32_t tile = sh.tile;
    const ParseState ps = jd->ps[sd.serial & 1u];
    if (tile == 0) copy_carry(sd, ps, Zprev, Z);

    uint32_t w[8];
    ChunkMasks m;
    load_chunk(sd, tile, tid, w, m);
#pragma unroll
    for (int j = 0; j < 8; j++) s_words[tid * SW_STRIDE + j] = w[j];
    const bool prev_nl = prev_is_newline(sd, tile, m, ps.at_line_start, sh.last);
    const uint32_t lsm = fa_line_starts(m, prev_nl);
    uint8_t *bytes = reinterpret_cast<uint8_t *>(s_words + tid * SW_STRIDE);
    uint32_t enc = 0;  // (thread + 1) << 1 | type of the thread's last line start
    if (lsm) {
        const uint8_t ch = bytes[31 - __clz(lsm)];
        enc = ((tid + 1u) << 1) | ((ch == '>' || ch == '@') ? 0u : 1u);
    }
    const uint32_t inc = block_incl_max<K1_THREADS>(enc, sh.w);
    uint32_t prev = __shfl_up_sync(FULL, inc, 1);  // exclusive value = inclusive value of the previous thread
    if (lane == 31) sh.last[warp] = inc;
    __syncthreads();
    if (lane == 0) prev = warp ? sh.last[warp - 1] : 0u;
    const FaThread o = fa_classify(m, lsm, bytes, prev != 0, prev & 1u);
    if (o.err) atomicOr(&jd->format_error, o.err);
    {
        uint32_t h = __popc(o.head), r = __popc(o.emit);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            h += __shfl_xor_sync(FULL, h, d);
            r += __shfl_xor_sync(FULL, r, d);
        }
        if (lane == 0) { atomicAdd(&sh.acc[0], h); atomicAdd(&sh.acc[1], r); }
    }
    if (m.valid) {
        const int vhi = 32 - __clz(m.valid);
        const long long pos0 = (long long)tile * K1_TILE + (long long)tid * K1_CHUNK;
        const long long hi = (long long)sd.skip + sd.len;
        const long long tile_hi = min(hi, (long long)(tile + 1) * K1_TILE);
        if (pos0 + vhi == tile_hi) sh.acc[4] = 0x80000000u | ((m.nl >> (vhi - 1)) & 1u);
    }
    if (tid == K1_THREADS - 1) sh.acc[5] = inc;  // tile's last line start (0: none)
    // header line starts become the record separator: mark them with '\n' in the staged text
    for (uint32_t sp = o.sepm; sp; sp &= sp - 1) bytes[__ffs(sp) - 1] = '\n';
    __syncthreads();
    if (warp == 0) {
        const uint32_t last = sh.acc[5];
        const uint32_t agg = last ? (1u | ((last & 1u) << 1)) : 0u;
        const uint32_t ty_ex = lookback<OP_LAST>(st1, tile, sd.epoch, agg, 1u | (ps.fa_type << 1));
        const uint32_t ty = (ty_ex >> 1) & 1u;
        const uint32_t out_bytes = sh.acc[1] + (ty ? sh.acc[0] : 0u);
        const uint32_t off = lookback<OP_SUM>(st2, tile, sd.epoch, out_bytes, Z_CARRY);
        if (lane == 0) { sh.state_in = ty; sh.out_off = off; }
    }
    __syncthreads();
    const uint32_t ty = sh.state_in, out_off = sh.out_off;
    const uint32_t emit = o.emit | (ty ? o.head : 0u);
    uint32_t total_emit;
    const uint32_t E = block_excl_sum<K1_THREADS>(__popc(emit), sh.w, total_emit);
    const uint32_t s_off = out_off & 15u;
    compact_runs(s_words, s_out, tid, emit, s_off + E);
    __syncthreads();
    write_out(s_out, s_off, total_emit, Z, out_off);

    if (tile + 1 == sd.n_tiles && tid == 0) {
        const uint32_t last = sh.acc[5];
        const uint32_t end_state = last ? (last & 1u) : ty;
        const uint32_t ends_nl = (sh.acc[4] & 0x80000000u) ? (sh.acc[4] & 1u) : ps.at_line_start;
        ParseState nx;
        uint32_t nz = out_off + total_emit;
        if (sd.last_of_file) {
            Z[nz++] = 0;  // terminate the last record (this CTA's own stores precede it in program order)
            nx.fq_phase = 0; nx.fa_type = 0; nx.at_line_start = 1;
        } else {
            nx.fq_phase = ps.fq_phase; nx.fa_type = end_state; nx.at_line_start = ends_nl;
        }
        nx.n_zbuf = nz;
        nx.z_begin = Z_CARRY;
        jd->ps[(sd.serial + 1u) & 1u] = nx;
    }
}

}  // namespace

void launch_parse_segment(const LaunchCtx &lc, const SegDesc &sd, unsigned long long *st1, unsigned long long *st2,
                          JobDev *jd, const uint8_t *Zprev, uint8_t *Z) {
    // sd.n_tiles >= 1 (an empty segment still threads the parser state through its single tile)
    if (sd.fmt == FMT_FASTQ) {
        static bool carveout_set = false;
        if (!carveout_set) {   // a preference, not a requirement: 6 CTAs x 21 KB of shared memory per SM
            cudaFuncSetAttribute(k1q_fused<FQ_NT, FQ_MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            carveout_set = true;
        }
        const uint32_t n_tiles = sd.len ? div_up((unsigned long long)sd.skip + sd.len, 64u * FQ_NT) : 1u;
        cudaMemsetAsync(st1, 0, (size_t)n_tiles * 8, lc.stream);      // status words of the look-back (no epoch tag in them)
        k1q_fused<FQ_NT, FQ_MINB><<<n_tiles, FQ_NT, 0, lc.stream>>>(sd, n_tiles, st1, jd, Zprev, Z);
    } else {
        k1_fasta<<<sd.n_tiles, K1_THREADS, 0, lc.stream>>>(sd, jd, st1, st2, Zprev, Z);
    }
    ++*lc.launches;
}

}  // namespace ac
