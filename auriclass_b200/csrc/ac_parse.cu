// ac_parse.cu -- kernel K1: FASTA / FASTQ record parser.
//
// Replaces Mash's kseq_read() loop (SURVEY.md section 2.2 stages S1-S2, App. A.2)
// that runs inside `mash sketch` (auriclass/classes.py:576-596, :696-713).
// Input: raw inflated text in HBM.  Output: the sanitised compacted base stream Z
// (kmer_core.cuh): sequence characters only, 0x00 after/between records.
//
// Shape: pass A (per 8 KiB tile: newline/line-start ballots as 32-bit masks per
// thread, what the tile would emit under every incoming parser state) -> a one-CTA
// scan that threads the parser state (FASTQ line phase / FASTA line type) and the
// output offsets through the tiles -> pass C (classify again with the known state,
// compact through shared memory, coalesced byte stores).  A segment (<= 32 MiB)
// stays L2 resident between the passes, so HBM sees the text once.
//
// FASTQ is parsed as strict 4-line records and *verified* (every record starts
// with '@', its third line with '+', sequence and quality byte totals match, the
// file ends on a record boundary); anything else raises AC_E_FORMAT -- the GPU
// path never guesses.  Multi-line FASTA is handled in full (k-mers span lines).
#include "ac_common.cuh"

namespace ac {

namespace {

__device__ __forceinline__ uint32_t gather_nib(uint32_t eq) {  // eq: 0xFF per selected byte -> 4-bit mask
    return ((eq & 0x01010101u) * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t bit_range(int a, int b) {  // bits [a, b), 0 <= a <= b <= 32
    uint32_t hi = (b >= 32) ? 0xffffffffu : ((1u << b) - 1u);
    uint32_t lo = (a >= 32) ? 0xffffffffu : ((1u << a) - 1u);
    return hi & ~lo;
}

struct ChunkMasks {
    uint32_t valid, nl;
    const uint8_t *bytes;  // address of the chunk's byte 0 (line-start characters are read from here, L1 hits)
};

// 0x80 in every byte of w equal to '\n' (exact: no borrow artefacts)
__device__ __forceinline__ uint32_t newline_bits(uint32_t w) {
    const uint32_t x = w ^ 0x0A0A0A0Au;
    const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x | 0x7F7F7F7Fu);
}

// Loads the thread's 32 raw bytes (two 16-byte vector loads) and builds per-byte masks.
__device__ __forceinline__ void load_chunk(const SegDesc &sd, uint32_t tile, uint32_t tid, uint32_t w[8],
                                           ChunkMasks &m) {
    const long long pos = (long long)tile * K1_TILE + (long long)tid * K1_CHUNK;
    const long long lo = sd.skip, hi = (long long)sd.skip + sd.len;
    const long long real_hi = hi - (sd.append_nl ? 1 : 0);
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
    uint32_t nl = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) nl |= (((newline_bits(w[j]) >> 7) * 0x01020408u) >> 24) << (4 * j);
    m.nl = nl & m.valid;
    m.bytes = sd.base + pos;
}

// Is the byte just before this thread's first valid byte a newline?  (seg_first: value to use
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

// ---------------------------------------------------------------- FASTQ
__global__ void __launch_bounds__(K1_THREADS) k1_fq_pass_a(SegDesc sd, TileSum *__restrict__ sums) {
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_cnt[4];
    __shared__ uint32_t s_ends;
    const uint32_t tile = blockIdx.x, tid = threadIdx.x;
    uint32_t w[8];
    ChunkMasks m;
    load_chunk(sd, tile, tid, w, m);
    if (tid < 4) s_cnt[tid] = 0;
    if (tid == 0) s_ends = 0;
    uint32_t total_nl;
    const uint32_t L0 = block_excl_sum<K1_THREADS>(__popc(m.nl), s_w, total_nl);
    uint32_t packed = 0;  // four 8-bit byte counts, index = local line index & 3
    if (m.valid) {
        const int vlo = __ffs(m.valid) - 1, vhi = 32 - __clz(m.valid);
        int pos = vlo;
        uint32_t L = L0;
        while (true) {
            const uint32_t r = m.nl & (0xffffffffu << pos);
            const int end = r ? (__ffs(r) - 1) : (vhi - 1);
            packed += (uint32_t)(end - pos + 1) << (8 * (L & 3));
            if (!r || end + 1 >= vhi) break;
            pos = end + 1; L++;
        }
        // does this thread own the last valid byte of the tile?
        const long long pos0 = (long long)tile * K1_TILE + (long long)tid * K1_CHUNK;
        const long long hi = (long long)sd.skip + sd.len;
        const long long tile_hi = min(hi, (long long)(tile + 1) * K1_TILE);
        if (pos0 + vhi == tile_hi) s_ends = (m.nl >> (vhi - 1)) & 1u;
    }
    // per-counter warp reduction: two 16-bit lanes per shuffle
    uint32_t lo = (packed & 0xFFu) | (((packed >> 8) & 0xFFu) << 16);
    uint32_t hi2 = ((packed >> 16) & 0xFFu) | (((packed >> 24) & 0xFFu) << 16);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        lo += __shfl_xor_sync(FULL, lo, d);
        hi2 += __shfl_xor_sync(FULL, hi2, d);
    }
    if ((tid & 31) == 0) {
        atomicAdd(&s_cnt[0], lo & 0xFFFFu); atomicAdd(&s_cnt[1], lo >> 16);
        atomicAdd(&s_cnt[2], hi2 & 0xFFFFu); atomicAdd(&s_cnt[3], hi2 >> 16);
    }
    __syncthreads();
    if (tid == 0) {
        TileSum s;
        s.a = total_nl;
        s.c[0] = s_cnt[0]; s.c[1] = s_cnt[1]; s.c[2] = s_cnt[2]; s.c[3] = s_cnt[3];
        s.ends_nl = s_ends;
        sums[tile] = s;
    }
}

__global__ void __launch_bounds__(K1_THREADS) k1_fq_pass_c(SegDesc sd, const TileIn *__restrict__ tin,
                                                           JobDev *__restrict__ jd, uint8_t *__restrict__ Z) {
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_last[8];
    __shared__ __align__(16) uint32_t s_words[K1_THREADS * 8];
    const uint32_t tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileIn ti = tin[tile];
    uint32_t w[8];
    ChunkMasks m;
    load_chunk(sd, tile, tid, w, m);
#pragma unroll
    for (int v = 0; v < 2; v++)
        reinterpret_cast<uint4 *>(s_words)[tid * 2 + v] = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
    uint32_t total_nl;
    const uint32_t L0 = block_excl_sum<K1_THREADS>(__popc(m.nl), s_w, total_nl);
    const bool prev_nl = prev_is_newline(sd, tile, m, ti.first_ls, s_last);
    uint32_t emit = 0, err = 0;
    if (m.valid) {
        const int vlo = __ffs(m.valid) - 1, vhi = 32 - __clz(m.valid);
        int pos = vlo;
        uint32_t ph = (ti.state + L0) & 3u;
        bool ls = prev_nl;
        while (true) {
            if (ls && !(ph & 1u)) {
                const uint8_t ch = __ldg(m.bytes + pos);
                if (ph == 0 && ch != '@') err |= FERR_FQ_HEADER;
                if (ph == 2 && ch != '+') err |= FERR_FQ_PLUS;
            }
            const uint32_t r = m.nl & (0xffffffffu << pos);
            const int end = r ? (__ffs(r) - 1) : (vhi - 1);
            if (ph == 1) emit |= bit_range(pos, end + 1);
            if (!r || end + 1 >= vhi) break;
            pos = end + 1; ph = (ph + 1) & 3u; ls = true;
        }
    }
    if (err) atomicOr(&jd->format_error, err);
    uint32_t total_emit;
    const uint32_t E = block_excl_sum<K1_THREADS>(__popc(emit), s_w, total_emit);  // also orders s_words
    // flat copy: a warp walks its own 1 KiB region 128 bytes at a time
#pragma unroll 1
    for (int step = 0; step < 8; step++) {
        const int srcl = 4 * step + (lane >> 3);
        const uint32_t em = __shfl_sync(FULL, emit, srcl);
        const uint32_t e0 = __shfl_sync(FULL, E, srcl);
        const int bitoff = (lane & 7) * 4;
        const uint32_t nib = (em >> bitoff) & 0xFu;
        if (nib) {
            const uint32_t word = s_words[warp * 256 + step * 32 + lane];
            uint32_t sw = sanitise4(word) & ~((newline_bits(word) >> 7) * 0xFFu);  // newline -> record separator
            uint32_t off = ti.out_off + e0 + __popc(em & ((1u << bitoff) - 1u));
#pragma unroll
            for (int k = 0; k < 4; k++)
                if ((nib >> k) & 1u) Z[off++] = (uint8_t)(sw >> (8 * k));
        }
    }
}

// ---------------------------------------------------------------- FASTA
// classification shared by pass A and pass C
struct FaThread {
    uint32_t emit, sepm;      // pass C masks
    uint32_t head_emit, rest_emit;
    uint32_t err;
};

template <bool PASS_C>
__device__ __forceinline__ FaThread fa_classify(const ChunkMasks &m, uint32_t lsm, bool known, uint32_t type) {
    FaThread o{0, 0, 0, 0, 0};
    if (!m.valid) return o;
    const int vlo = __ffs(m.valid) - 1, vhi = 32 - __clz(m.valid);
    int pos = vlo;
    while (true) {
        if ((lsm >> pos) & 1u) {
            const uint8_t ch = __ldg(m.bytes + pos);
            type = (ch == '>' || ch == '@') ? 0u : 1u;
            known = true;
            if (ch == '+') o.err |= FERR_FA_PLUS;
            if (type == 0u) { o.rest_emit += 1; if (PASS_C) { o.emit |= 1u << pos; o.sepm |= 1u << pos; } }
        }
        const uint32_t next = (pos + 1 < 32) ? (lsm & (0xffffffffu << (pos + 1))) : 0u;
        const int end = next ? (__ffs(next) - 1) : vhi;  // exclusive
        const uint32_t body = bit_range(pos, end) & ~m.nl & m.valid;
        if (!known) o.head_emit += __popc(body);
        else if (type == 1u) { o.rest_emit += __popc(body); if (PASS_C) o.emit |= body; }
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

__global__ void __launch_bounds__(K1_THREADS) k1_fa_pass_a(SegDesc sd, const JobDev *__restrict__ jd,
                                                           TileSum *__restrict__ sums) {
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_last[8];
    __shared__ uint32_t s_acc[3];  // head, rest, ends
    const uint32_t tile = blockIdx.x, tid = threadIdx.x;
    uint32_t w[8];
    ChunkMasks m;
    load_chunk(sd, tile, tid, w, m);
    if (tid < 3) s_acc[tid] = 0;
    const bool prev_nl = prev_is_newline(sd, tile, m, jd->at_line_start, s_last);
    const uint32_t lsm = fa_line_starts(m, prev_nl);
    uint32_t enc = 0;
    if (lsm) {
        const uint8_t ch = __ldg(m.bytes + (31 - __clz(lsm)));
        enc = ((tid + 1u) << 1) | ((ch == '>' || ch == '@') ? 0u : 1u);
    }
    const uint32_t inc = block_incl_max<K1_THREADS>(enc, s_w);
    // exclusive value = inclusive value of the previous thread
    uint32_t prev = __shfl_up_sync(FULL, inc, 1);
    if ((tid & 31) == 31) s_last[tid >> 5] = inc;
    __syncthreads();
    if ((tid & 31) == 0) prev = (tid >> 5) ? s_last[(tid >> 5) - 1] : 0u;
    FaThread o = fa_classify<false>(m, lsm, prev != 0, prev & 1u);
    uint32_t h = o.head_emit, r = o.rest_emit;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        h += __shfl_xor_sync(FULL, h, d);
        r += __shfl_xor_sync(FULL, r, d);
    }
    if ((tid & 31) == 0) { atomicAdd(&s_acc[0], h); atomicAdd(&s_acc[1], r); }
    if (m.valid) {
        const int vhi = 32 - __clz(m.valid);
        const long long pos0 = (long long)tile * K1_TILE + (long long)tid * K1_CHUNK;
        const long long hi = (long long)sd.skip + sd.len;
        const long long tile_hi = min(hi, (long long)(tile + 1) * K1_TILE);
        if (pos0 + vhi == tile_hi) s_acc[2] = (m.nl >> (vhi - 1)) & 1u;
    }
    __syncthreads();
    if (tid == K1_THREADS - 1) {
        TileSum s;
        s.a = (inc != 0 ? 1u : 0u) | ((inc & 1u) << 1);
        s.c[0] = s_acc[0]; s.c[1] = s_acc[1]; s.c[2] = 0; s.c[3] = 0;
        s.ends_nl = s_acc[2];
        sums[tile] = s;
    }
}

__global__ void __launch_bounds__(K1_THREADS) k1_fa_pass_c(SegDesc sd, const TileIn *__restrict__ tin,
                                                           JobDev *__restrict__ jd, uint8_t *__restrict__ Z) {
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_last[8];
    __shared__ __align__(16) uint32_t s_words[K1_THREADS * 8];
    const uint32_t tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileIn ti = tin[tile];
    uint32_t w[8];
    ChunkMasks m;
    load_chunk(sd, tile, tid, w, m);
#pragma unroll
    for (int v = 0; v < 2; v++)
        reinterpret_cast<uint4 *>(s_words)[tid * 2 + v] = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
    const bool prev_nl = prev_is_newline(sd, tile, m, ti.first_ls, s_last);
    const uint32_t lsm = fa_line_starts(m, prev_nl);
    uint32_t enc = 0;
    if (lsm) {
        const uint8_t ch = __ldg(m.bytes + (31 - __clz(lsm)));
        enc = ((tid + 1u) << 1) | ((ch == '>' || ch == '@') ? 0u : 1u);
    }
    const uint32_t inc = block_incl_max<K1_THREADS>(enc, s_w);
    uint32_t prev = __shfl_up_sync(FULL, inc, 1);
    if (lane == 31) s_last[warp] = inc;
    __syncthreads();
    if (lane == 0) prev = warp ? s_last[warp - 1] : 0u;
    const uint32_t type_in = prev ? (prev & 1u) : ti.state;
    FaThread o = fa_classify<true>(m, lsm, true, type_in);
    if (o.err) atomicOr(&jd->format_error, o.err);
    uint32_t total_emit;
    const uint32_t E = block_excl_sum<K1_THREADS>(__popc(o.emit), s_w, total_emit);
#pragma unroll 1
    for (int step = 0; step < 8; step++) {
        const int srcl = 4 * step + (lane >> 3);
        const uint32_t em = __shfl_sync(FULL, o.emit, srcl);
        const uint32_t sp = __shfl_sync(FULL, o.sepm, srcl);
        const uint32_t e0 = __shfl_sync(FULL, E, srcl);
        const int bitoff = (lane & 7) * 4;
        const uint32_t nib = (em >> bitoff) & 0xFu;
        if (nib) {
            const uint32_t word = s_words[warp * 256 + step * 32 + lane];
            const uint32_t spn = (sp >> bitoff) & 0xFu;
            // separator bytes -> 0x00: expand the 4-bit mask to a byte mask
            const uint32_t spm = ((spn & 1u) * 0xFFu) | (((spn >> 1) & 1u) * 0xFF00u) | (((spn >> 2) & 1u) * 0xFF0000u) |
                                 (((spn >> 3) & 1u) * 0xFF000000u);
            const uint32_t sw = sanitise4(word) & ~spm;
            uint32_t off = ti.out_off + e0 + __popc(em & ((1u << bitoff) - 1u));
#pragma unroll
            for (int k = 0; k < 4; k++)
                if ((nib >> k) & 1u) Z[off++] = (uint8_t)(sw >> (8 * k));
        }
    }
}

// ---------------------------------------------------------------- scan
constexpr int SCAN_T = 1024;

__global__ void __launch_bounds__(SCAN_T) k1_scan(SegDesc sd, const TileSum *__restrict__ sums,
                                                  TileIn *__restrict__ tin, JobDev *__restrict__ jd,
                                                  uint8_t *__restrict__ Z) {
    __shared__ uint32_t s_w[33];
    const uint32_t tid = threadIdx.x;
    const uint32_t n = sd.n_tiles;
    const uint32_t per = (n + SCAN_T - 1) / SCAN_T;
    const uint32_t t0 = min(n, tid * per), t1 = min(n, t0 + per);

    // carry: last Z_CARRY bytes of the previous segment's stream move to the front
    const uint32_t old_n = jd->n_zbuf;
    uint8_t carry = 0;
    if (tid < Z_CARRY && !sd.first_of_stream) carry = Z[old_n - Z_CARRY + tid];
    const uint32_t phase0 = jd->fq_phase, type0 = jd->fa_type, ls0 = jd->at_line_start;
    __syncthreads();
    if (tid < Z_CARRY) Z[tid] = carry;

    uint32_t total_out = 0, end_state = 0;
    if (sd.fmt == FMT_FASTQ) {
        uint32_t nl_local = 0;
        for (uint32_t t = t0; t < t1; t++) nl_local += sums[t].a;
        uint32_t total_nl;
        uint32_t ex_nl = block_excl_sum<SCAN_T>(nl_local, s_w, total_nl);
        uint32_t out_local = 0, qual_local = 0, run = ex_nl;
        for (uint32_t t = t0; t < t1; t++) {
            const uint32_t ph = (phase0 + run) & 3u;
            out_local += sums[t].c[(1u - ph) & 3u];
            qual_local += sums[t].c[(3u - ph) & 3u];
            run += sums[t].a;
        }
        uint32_t total_qual;
        const uint32_t ex_q = block_excl_sum<SCAN_T>(qual_local, s_w, total_qual);
        (void)ex_q;
        uint32_t off = block_excl_sum<SCAN_T>(out_local, s_w, total_out);
        run = ex_nl;
        for (uint32_t t = t0; t < t1; t++) {
            const uint32_t ph = (phase0 + run) & 3u;
            TileIn ti;
            ti.state = ph; ti.out_off = Z_CARRY + off; ti.first_ls = (t == 0) ? ls0 : 0u;
            tin[t] = ti;
            off += sums[t].c[(1u - ph) & 3u];
            run += sums[t].a;
        }
        end_state = (phase0 + total_nl) & 3u;
        if (tid == 0) {
            jd->fq_seq_bytes += total_out;
            jd->fq_qual_bytes += total_qual;
        }
    } else {
        // incoming line type per tile = type of the nearest earlier tile that saw a line start
        uint32_t enc = 0;
        for (uint32_t t = t0; t < t1; t++)
            if (sums[t].a & 1u) enc = ((t + 1u) << 1) | ((sums[t].a >> 1) & 1u);
        const uint32_t inc = block_incl_max<SCAN_T>(enc, s_w);
        // exclusive: value of the previous thread
        __shared__ uint32_t s_inc[SCAN_T];
        s_inc[tid] = inc;
        __syncthreads();
        uint32_t cur = tid ? s_inc[tid - 1] : 0u;
        uint32_t out_local = 0;
        uint32_t c2 = cur;
        for (uint32_t t = t0; t < t1; t++) {
            const uint32_t ty = c2 ? (c2 & 1u) : type0;
            out_local += sums[t].c[1] + (ty ? sums[t].c[0] : 0u);
            if (sums[t].a & 1u) c2 = ((t + 1u) << 1) | ((sums[t].a >> 1) & 1u);
        }
        uint32_t off = block_excl_sum<SCAN_T>(out_local, s_w, total_out);
        c2 = cur;
        for (uint32_t t = t0; t < t1; t++) {
            const uint32_t ty = c2 ? (c2 & 1u) : type0;
            TileIn ti;
            ti.state = ty; ti.out_off = Z_CARRY + off; ti.first_ls = (t == 0) ? ls0 : 0u;
            tin[t] = ti;
            off += sums[t].c[1] + (ty ? sums[t].c[0] : 0u);
            if (sums[t].a & 1u) c2 = ((t + 1u) << 1) | ((sums[t].a >> 1) & 1u);
        }
        const uint32_t last = s_inc[SCAN_T - 1];
        end_state = last ? (last & 1u) : type0;
    }
    if (tid == 0) {
        uint32_t nz = Z_CARRY + total_out;
        const uint32_t ends_nl = n ? sums[n - 1].ends_nl : ls0;
        if (sd.last_of_file) {
            if (sd.fmt == FMT_FASTA) Z[nz++] = 0;  // terminate the last record
            else if (end_state != 0 || !ends_nl || jd->fq_seq_bytes != jd->fq_qual_bytes)
                atomicOr(&jd->format_error, (uint32_t)FERR_FQ_TRUNC);
            jd->fq_phase = 0; jd->fa_type = 0; jd->at_line_start = 1;
        } else {
            if (sd.fmt == FMT_FASTQ) jd->fq_phase = end_state; else jd->fa_type = end_state;
            jd->at_line_start = ends_nl;
        }
        jd->n_zbuf = nz;
    }
}

}  // namespace

void launch_parse_segment(const LaunchCtx &lc, const SegDesc &sd, TileSum *sums, TileIn *tin, JobDev *jd,
                          uint8_t *Z) {
    if (sd.n_tiles) {
        if (sd.fmt == FMT_FASTQ) k1_fq_pass_a<<<sd.n_tiles, K1_THREADS, 0, lc.stream>>>(sd, sums);
        else k1_fa_pass_a<<<sd.n_tiles, K1_THREADS, 0, lc.stream>>>(sd, jd, sums);
        ++*lc.launches;
    }
    k1_scan<<<1, SCAN_T, 0, lc.stream>>>(sd, sums, tin, jd, Z);
    ++*lc.launches;
    if (sd.n_tiles) {
        if (sd.fmt == FMT_FASTQ) k1_fq_pass_c<<<sd.n_tiles, K1_THREADS, 0, lc.stream>>>(sd, tin, jd, Z);
        else k1_fa_pass_c<<<sd.n_tiles, K1_THREADS, 0, lc.stream>>>(sd, tin, jd, Z);
        ++*lc.launches;
    }
}

}  // namespace ac
