// kmer_core.cuh -- per-thread canonical k-mer + MurmurHash3_x64_128 core of kernel K2.
//
// Replaces Mash's Sketch.cpp addMinHashes()/getHash() (external binary called at
// auriclass/classes.py:576-596 and :696-713; algorithm: SURVEY.md App. A.3).
// Everything here is __host__ __device__ so tests/test_host.py can run the
// very same arithmetic on the CPU (through tests/hostcheck/hostcheck.cpp) against the oracle;
// the product path only ever calls it from device code.
//
// Data model.  K1 hands K2 a *sanitised compacted base stream* Z: one byte per
// sequence character, upper-case 'A','C','G','T' for valid bases, 0x00 for a
// record separator and 0x01 for any other character.  Valid bytes all have bit 6
// set, invalid ones do not.  A thread owns W consecutive window starts and holds
// its run of W+K-1 bytes as little-endian 32-bit words r[0..RW).  The k-mer ASCII
// words MurmurHash3 needs are then byte-granular *slices* of r (forward strand)
// and of q = reverse(complement(r)) (reverse strand): one PRMT per word, no
// rolling state and no 2-bit <-> ASCII expansion.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AC_HD __host__ __device__ __forceinline__
#else
#define AC_HD inline
#endif

namespace ackmer {

AC_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t s = (sel >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(v >> (8 * (s & 7))) & 0xFF;
        if (s & 8) byte = (byte & 0x80) ? 0xFF : 0;
        r |= byte << (8 * i);
    }
    return r;
#endif
}

// 64-bit rotate left by a compile-time amount: two funnel shifts on the device
template <int R>
AC_HD uint64_t rotl64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    uint32_t nlo, nhi;
    if constexpr (R == 32) { nlo = hi; nhi = lo; }
    else if constexpr (R < 32) { nhi = __funnelshift_l(lo, hi, R); nlo = __funnelshift_l(hi, lo, R); }
    else { nhi = __funnelshift_l(hi, lo, R - 32); nlo = __funnelshift_l(lo, hi, R - 32); }
    return ((uint64_t)nhi << 32) | nlo;
#else
    return (x << R) | (x >> (64 - R));
#endif
}

// x * c (mod 2^64): IMAD.WIDE + 2 IMAD on the device (nvcc's own expansion spends a 4th on an add)
AC_HD uint64_t mul64(uint64_t x, uint64_t c) {
#if defined(__CUDA_ARCH__)
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32), cl = (uint32_t)c, ch = (uint32_t)(c >> 32);
    uint64_t p;
    uint32_t hi;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(xl), "r"(cl));
    hi = (uint32_t)(p >> 32);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(xl), "r"(ch));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(xh), "r"(cl));
    return ((uint64_t)hi << 32) | (uint32_t)p;
#else
    return x * c;
#endif
}

// keeps a value in a register: ptxas otherwise rematerialises cheap SWAR results per use
AC_HD void pin_reg(uint32_t &v) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+r"(v));
#else
    (void)v;
#endif
}

AC_HD uint64_t fmix64(uint64_t x) {
    x ^= x >> 33; x = mul64(x, 0xff51afd7ed558ccdULL);
    x ^= x >> 33; x = mul64(x, 0xc4ceb9fe1a85ec53ULL);
    x ^= x >> 33;
    return x;
}

// word j of a K-byte key held as NW little-endian words, 0 beyond the key
template <int K, int J>
AC_HD uint32_t key_word(const uint32_t *w) {
    if constexpr (J < (K + 3) / 4) return w[J]; else return 0u;
}

// h * 5 + c (mod 2^64): IMAD.WIDE + IMAD on the device.  (The shift-and-add form on the alu pipe -- LEA, SHF, two
// carry adds -- takes 16 cycles off the fmaheavy pipe per window and made K2 3 % slower: the alu pipe is as full.)
AC_HD uint64_t mul5_add(uint64_t h, uint64_t c) {
#if defined(__CUDA_ARCH__)
    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    uint64_t p;
    uint32_t ph;
    asm("mad.wide.u32 %0, %1, 5, %2;" : "=l"(p) : "r"(lo), "l"(c));
    ph = (uint32_t)(p >> 32);
    asm("mad.lo.u32 %0, %1, 5, %0;" : "+r"(ph) : "r"(hi));
    return ((uint64_t)ph << 32) | (uint32_t)p;
#else
    return h * 5 + c;
#endif
}

// the two additive constants of the MurmurHash3 block step (a kernel keeps them in registers)
struct MurmurConsts {
    uint64_t a, b;
    AC_HD MurmurConsts() : a(0x52dce729ULL), b(0x38495ab5ULL) {}
};

// fmix64 up to, not including, its second multiply ("half-open"): q with fmix64(x) = f(q * FMIX_C2), f(p) = p ^ (p >> 33)
constexpr uint64_t FMIX_C1 = 0xff51afd7ed558ccdULL, FMIX_C2 = 0xc4ceb9fe1a85ec53ULL;
AC_HD uint64_t fmix64_half(uint64_t x) {
    x ^= x >> 33; x = mul64(x, FMIX_C1);
    x ^= x >> 33;
    return x;
}
// h1 of the digest from the two half-open fmix values: fmix(h1) + fmix(h2)
AC_HD uint64_t murmur3_close(uint64_t q1, uint64_t q2) {
    const uint64_t p1 = mul64(q1, FMIX_C2), p2 = mul64(q2, FMIX_C2);
    return (p1 ^ (p1 >> 33)) + (p2 ^ (p2 >> 33));
}
// MurmurHash3_x64_128(key, K, seed) up to the half-open fmix values (q1, q2); h1 of the digest =
// murmur3_close(q1, q2).  w[] = the K key bytes as (K+3)/4 little-endian words, bytes past K in the
// last word MUST be zero.  Mash: hash = first 8 bytes of the 128-bit digest = h1.
template <int K>
AC_HD void murmur3_open(const uint32_t *w, uint32_t seed, const MurmurConsts &mc, uint64_t &p1, uint64_t &p2) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    constexpr int NB = K / 16;
    if constexpr (NB >= 1) {
        uint64_t k1 = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
        uint64_t k2 = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
        k1 = mul64(k1, c1); k1 = rotl64<31>(k1); k1 = mul64(k1, c2); h1 ^= k1;
        h1 = rotl64<27>(h1); h1 += h2; h1 = mul5_add(h1, mc.a);
        k2 = mul64(k2, c2); k2 = rotl64<33>(k2); k2 = mul64(k2, c1); h2 ^= k2;
        h2 = rotl64<31>(h2); h2 += h1; h2 = mul5_add(h2, mc.b);
    }
    if constexpr (NB >= 2) {
        uint64_t k1 = (uint64_t)w[4] | ((uint64_t)w[5] << 32);
        uint64_t k2 = (uint64_t)w[6] | ((uint64_t)w[7] << 32);
        k1 = mul64(k1, c1); k1 = rotl64<31>(k1); k1 = mul64(k1, c2); h1 ^= k1;
        h1 = rotl64<27>(h1); h1 += h2; h1 = mul5_add(h1, mc.a);
        k2 = mul64(k2, c2); k2 = rotl64<33>(k2); k2 = mul64(k2, c1); h2 ^= k2;
        h2 = rotl64<31>(h2); h2 += h1; h2 = mul5_add(h2, mc.b);
    }
    constexpr int T = K & 15;
    if constexpr (T > 8) {
        uint64_t k2 = (uint64_t)key_word<K, 4 * NB + 2>(w) | ((uint64_t)key_word<K, 4 * NB + 3>(w) << 32);
        k2 = mul64(k2, c2); k2 = rotl64<33>(k2); k2 = mul64(k2, c1); h2 ^= k2;
    }
    if constexpr (T > 0) {
        uint64_t k1 = (uint64_t)key_word<K, 4 * NB>(w) | ((uint64_t)key_word<K, 4 * NB + 1>(w) << 32);
        k1 = mul64(k1, c1); k1 = rotl64<31>(k1); k1 = mul64(k1, c2); h1 ^= k1;
    }
    h1 ^= (uint64_t)K; h2 ^= (uint64_t)K;
    h1 += h2; h2 += h1;
    p1 = fmix64_half(h1); p2 = fmix64_half(h2);
}

// h1 of MurmurHash3_x64_128(key, K, seed)
template <int K>
AC_HD uint64_t murmur3_h1(const uint32_t *w, uint32_t seed) {
    const MurmurConsts mc;
    uint64_t p1, p2;
    murmur3_open<K>(w, seed, mc, p1, p2);
    return murmur3_close(p1, p2);
}

// SWAR complement of four sanitised bases: A(41)<->T(54) differ by 0x15,
// C(43)<->G(47) by 0x04; bit 1 of the byte tells the pair.
AC_HD uint32_t comp4(uint32_t w) {
    uint32_t b = (w >> 1) & 0x01010101u;
    return w ^ (0x15151515u ^ (b * 0x11u));
}

// bit j of the result = byte j of w is a valid base (bit 6 set)
AC_HD uint32_t valid_nibble(uint32_t w) {
    uint32_t t = (w >> 6) & 0x01010101u;
    return (t * 0x01020408u) >> 24;
}

template <int K, int W>
struct RunCfg {
    static constexpr int RUN_BYTES = W + K - 1;             // bytes the W windows touch
    static constexpr int RW = 4 * ((W + K + 15) / 16);      // words held per thread (16-byte loads); >= W+K bytes so
                                                            // K2's record statistics can look K bytes back from byte K+W-1
    static constexpr int NW = (K + 3) / 4;                  // words per k-mer
    static_assert(4 * RW <= 64, "validity mask is 64 bits");
    static_assert(W <= 32, "window mask is 32 bits");
};

// bit i of the result: bytes i .. i+K-1 of the run are all valid
template <int K>
AC_HD uint64_t windows_all_valid(uint64_t m) {
    uint64_t p[6];
    p[0] = m;
    p[1] = p[0] & (p[0] >> 1);
    p[2] = p[1] & (p[1] >> 2);
    p[3] = p[2] & (p[2] >> 4);
    p[4] = p[3] & (p[3] >> 8);
    p[5] = p[4] & (p[4] >> 16);
    uint64_t res = ~0ULL;
    int off = 0;
#pragma unroll
    for (int b = 5; b >= 0; b--) {
        if ((K >> b) & 1) { res &= p[b] >> off; off += 1 << b; }
    }
    return res;
}

// word j (bytes off+4j .. off+4j+3) of the byte string held in words a[]
template <int OFF>
AC_HD uint32_t slice_word(const uint32_t *a) {
    constexpr int wi = OFF >> 2, sh = OFF & 3;
    if constexpr (sh == 0) return a[wi];
    else if constexpr (sh == 1) return byte_perm(a[wi], a[wi + 1], 0x4321);
    else if constexpr (sh == 2) return byte_perm(a[wi], a[wi + 1], 0x5432);
    else return byte_perm(a[wi], a[wi + 1], 0x6543);
}

template <int K, int OFF, int J, int NW>
AC_HD void slice_key(const uint32_t *a, uint32_t *out) {
    if constexpr (J < NW) {
        uint32_t v;
        if constexpr (J == NW - 1 && (K & 3) != 0) {
            // last word: only K&3 bytes belong to the key; avoid reading a[wi+1] past the run
            constexpr int o = OFF + 4 * J, wi = o >> 2, sh = o & 3, nb = K & 3;
            if constexpr (sh + nb <= 4) v = (a[wi] >> (8 * sh)) & ((1u << (8 * nb)) - 1u);
            else v = slice_word<o>(a) & ((1u << (8 * nb)) - 1u);
        } else {
            v = slice_word<OFF + 4 * J>(a);
        }
        out[J] = v;
        slice_key<K, OFF, J + 1, NW>(a, out);
    }
}

// true when the reverse-complement key c sorts strictly before the forward key f
// (memcmp order over the K ASCII bytes).
template <int NW>
AC_HD bool rc_is_smaller(const uint32_t *f, const uint32_t *c) {
    // big-endian view of the first word decides unless the first 4 bases tie
    uint32_t fb = byte_perm(f[0], 0, 0x0123), cb = byte_perm(c[0], 0, 0x0123);
    if (fb != cb) return cb < fb;
#pragma unroll
    for (int j = 1; j < NW; j++) {
        fb = byte_perm(f[j], 0, 0x0123); cb = byte_perm(c[j], 0, 0x0123);
        if (fb != cb) return cb < fb;
    }
    return false;
}

template <int K, int W, int I, class Emit>
AC_HD void process_window(const uint32_t *r, const uint32_t *q, uint32_t vmask, uint32_t seed, Emit &emit) {
    if constexpr (I < W) {
        using C = RunCfg<K, W>;
        {
            // every lane hashes every window (a warp runs the hash as soon as one lane needs it);
            // windows touching an invalid byte are dropped at emission
            uint32_t f[C::NW], c[C::NW];
            slice_key<K, I, 0, C::NW>(r, f);
            slice_key<K, 4 * C::RW - K - I, 0, C::NW>(q, c);
            const bool use_rc = rc_is_smaller<C::NW>(f, c);
            uint32_t key[C::NW];
#pragma unroll
            for (int j = 0; j < C::NW; j++) key[j] = use_rc ? c[j] : f[j];
            emit(vmask, 1u << I, murmur3_h1<K>(key, seed));
        }
        process_window<K, W, I + 1, Emit>(r, q, vmask, seed, emit);
    }
}

// Hash G consecutive windows.  fw: forward bytes as LE words, window i starts at byte i.
// rw: reverse-complement bytes as LE words, window i starts at byte RC0 - i (RC0 compile time).
// vbits: bit i = window i valid.  emit(vbits, 1<<i, h1) is called for every window.
template <int K, int G, int RC0, int I, class Emit>
AC_HD void process_group(const uint32_t *fw, const uint32_t *rw, uint32_t vbits, uint32_t seed, Emit &emit) {
    if constexpr (I < G) {
        constexpr int NW = (K + 3) / 4;
        uint32_t f[NW], c[NW];
        slice_key<K, I, 0, NW>(fw, f);
        slice_key<K, RC0 - I, 0, NW>(rw, c);
        const bool use_rc = rc_is_smaller<NW>(f, c);
        uint32_t key[NW];
#pragma unroll
        for (int j = 0; j < NW; j++) key[j] = use_rc ? c[j] : f[j];
        emit(vbits, 1u << I, murmur3_h1<K>(key, seed));
        process_group<K, G, RC0, I + 1, Emit>(fw, rw, vbits, seed, emit);
    }
}

// Hash every valid window of one thread's run.  r[RW]: run bytes as LE words
// (bytes past the run may hold anything; they only enter windows >= W).
// win_ok: bit i set when window i may be emitted (range clipping by the caller).
// emit(vmask, 1<<i, h1) is called for every window (valid iff vmask & bit); returns the window mask.
template <int K, int W, class Emit>
AC_HD uint32_t process_run(const uint32_t *r, uint32_t win_ok, uint32_t seed, Emit &emit) {
    using C = RunCfg<K, W>;
    uint64_t m = 0;
#pragma unroll
    for (int j = 0; j < C::RW; j++) m |= (uint64_t)valid_nibble(r[j]) << (4 * j);
    uint32_t vmask = (uint32_t)windows_all_valid<K>(m) & win_ok;
    if (W < 32) vmask &= (1u << (W & 31)) - 1u;
    if (vmask == 0) return 0;
    uint32_t q[C::RW];
#pragma unroll
    for (int j = 0; j < C::RW; j++) { q[j] = byte_perm(comp4(r[C::RW - 1 - j]), 0, 0x0123); pin_reg(q[j]); }
    process_window<K, W, 0, Emit>(r, q, vmask, seed, emit);
    return vmask;
}

// ---- quad layout (kernel K2): a lane owns G windows at a stride of 4 bytes, so that after ONE byte
// alignment of its words (preshift) every window's key words are whole registers: window j of the lane
// reads forward words j .. j+NW-1 and reverse-complement words (G-1-j) .. (G-1-j)+NW-1.
AC_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t bits /*0..31*/) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, bits);
#else
    return bits ? ((lo >> bits) | (hi << (32 - bits))) : lo;
#endif
}

template <int K>
struct QuadCfg {
    static constexpr int G = 8;                  // windows per lane and group (4 lanes x 8 = 32 consecutive windows)
    static constexpr int NW = (K + 3) / 4;       // words per k-mer
    static constexpr int NS = G - 1 + NW;        // aligned words a lane needs per strand
    static constexpr int NL = NS + 1;            // words loaded per strand (one more feeds the funnel shift)
};

// in[N+1] aligned words, off = byte offset (0..3) of the first wanted byte inside in[0] -> out[N] words starting there
template <int N>
AC_HD void preshift(const uint32_t *in, uint32_t off, uint32_t *out) {
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = funnel_r(in[i], in[i + 1], 8u * off);
}

// fw / rw: the lane's pre-shifted forward / reverse-complement words (QuadCfg::NS each).
// vbits: bit 4j set = the lane's window j may be emitted.  emit(vbits, 1 << 4j, p1, p2) is called for every
// window with the half-open fmix values: the hash is murmur3_close(p1, p2), which the caller only needs to
// finish for the few windows whose gate value (gate_value) can still pass its threshold.
template <int K, int J, class Emit>
AC_HD void process_quad_windows(const uint32_t *fw, const uint32_t *rw, uint32_t vbits, uint32_t seed,
                                const MurmurConsts &mc, Emit &emit) {
    using C = QuadCfg<K>;
    if constexpr (J < C::G) {
        uint32_t f[C::NW], c[C::NW];
#pragma unroll
        for (int m = 0; m < C::NW; m++) { f[m] = fw[J + m]; c[m] = rw[C::G - 1 - J + m]; }
        constexpr uint32_t last_mask = (K & 3) ? ((1u << (8 * (K & 3))) - 1u) : 0xffffffffu;
        // strand choice: the first four bases decide unless they tie (1 in 256); only the tie path needs
        // the last words trimmed to the k-mer
        bool use_rc;
        if constexpr (C::NW == 1) {
            use_rc = byte_perm(c[0] & last_mask, 0, 0x0123) < byte_perm(f[0] & last_mask, 0, 0x0123);
        } else {
            const uint32_t fb = byte_perm(f[0], 0, 0x0123), cb = byte_perm(c[0], 0, 0x0123);
            if (fb != cb) use_rc = cb < fb;
            else {
                uint32_t ft[C::NW], ct[C::NW];
#pragma unroll
                for (int m = 0; m < C::NW; m++) { ft[m] = f[m]; ct[m] = c[m]; }
                ft[C::NW - 1] &= last_mask; ct[C::NW - 1] &= last_mask;
                use_rc = rc_is_smaller<C::NW>(ft, ct);
            }
        }
        uint32_t key[C::NW];
#pragma unroll
        for (int m = 0; m < C::NW; m++) key[m] = use_rc ? c[m] : f[m];
        key[C::NW - 1] &= last_mask;  // bytes past the k-mer in its last word are not part of the key
        uint64_t p1, p2;
        murmur3_open<K>(key, seed, mc, p1, p2);
        emit(vbits, 1u << (4 * J), p1, p2);
        process_quad_windows<K, J + 1, Emit>(fw, rw, vbits, seed, mc, emit);
    }
}

// ---- branch-free form of the quad pass (K2's hot loop) + the exact per-window re-hash behind it.
// The value a threshold test compares with gate_of_tau(tau).  64-bit hashes are tested on the high word of the digest
// alone: h = f(P1) + f(P2) with P = q * FMIX_C2 and f(p) = p ^ (p >> 33) touching only the low word, so
// hi(h) = hi(P1) + hi(P2) + carry, carry in {0, 1}, and hi(h) <= hi(tau) implies t <= hi(tau) or t == 2^32 - 1 for
// t = hi(P1) + hi(P2), i.e. t + 1 <= hi(tau) + 1 (mod 2^32).  32-bit hashes (k <= 16): the hash itself against tau.
// (One product instead of two -- hi((q1 + q2) * FMIX_C2) is t or t + 1 by distributivity -- saves 8 fmaheavy
// cycles per window and did not change K2's time; the two-product form is kept.)
template <bool USE64>
AC_HD uint32_t gate_value(uint64_t q1, uint64_t q2) {
    const uint64_t p1 = mul64(q1, FMIX_C2), p2 = mul64(q2, FMIX_C2);
    if constexpr (USE64) return (uint32_t)(p1 >> 32) + (uint32_t)(p2 >> 32) + 1u;
    else return ((uint32_t)p1 ^ (uint32_t)(p1 >> 33)) + ((uint32_t)p2 ^ (uint32_t)(p2 >> 33));
}
AC_HD uint32_t gate_of_tau(uint64_t tau, bool use64) {
    const uint32_t top = (uint32_t)(tau >> 32);
    if (!use64) return (uint32_t)tau;
    return top == 0xffffffffu ? top : top + 1u;
}

// Same windows as process_quad_windows, as ONE basic block: no branch, no call.  The strand is chosen from the
// first four bases (cb < fb); a window is *flagged* (bit 4j of flags) when those four bases tie (1 in 256: the
// choice may be wrong) or when its gate value passes `gate` (it may be a candidate).  Flagged windows are hashed
// again, exactly, by rehash_window; for all others the only fact needed -- hash > tau -- is already established.
// The caller masks flags with the window validity bits.
template <int K, int J, bool USE64>
AC_HD void process_quad_windows_flat(const uint32_t *fw, const uint32_t *rw, uint32_t seed, const MurmurConsts &mc,
                                     uint32_t gate, uint32_t &flags) {
    using C = QuadCfg<K>;
    if constexpr (J < C::G) {
        uint32_t f[C::NW], c[C::NW];
#pragma unroll
        for (int m = 0; m < C::NW; m++) { f[m] = fw[J + m]; c[m] = rw[C::G - 1 - J + m]; }
        constexpr uint32_t last_mask = (K & 3) ? ((1u << (8 * (K & 3))) - 1u) : 0xffffffffu;
        bool use_rc;
        uint32_t fb = 0, cb = 1;      // compared big-endian first words; equal = the first four bases tie
        if constexpr (C::NW == 1) {   // the whole k-mer is in the compared word: equal words = equal keys, nothing to redo
            use_rc = byte_perm(c[0] & last_mask, 0, 0x0123) < byte_perm(f[0] & last_mask, 0, 0x0123);
        } else {
            fb = byte_perm(f[0], 0, 0x0123); cb = byte_perm(c[0], 0, 0x0123);
            use_rc = cb < fb;
        }
        uint32_t key[C::NW];
#pragma unroll
        for (int m = 0; m < C::NW; m++) key[m] = use_rc ? c[m] : f[m];
        key[C::NW - 1] &= last_mask;
        uint64_t q1, q2;
        murmur3_open<K>(key, seed, mc, q1, q2);
        const uint32_t gv = gate_value<USE64>(q1, q2);
#if defined(__CUDA_ARCH__)
        // opaque to the compiler on purpose: as C++ the tie test turns into a branch around the gate arithmetic
        asm("{\n.reg .pred t, g;\nsetp.eq.u32 t, %1, %2;\nsetp.le.u32 g, %3, %4;\nor.pred t, t, g;\n@t or.b32 %0, %0, %5;\n}"
            : "+r"(flags) : "r"(fb), "r"(cb), "r"(gv), "r"(gate), "n"(1u << (4 * J)));
#else
        if (fb == cb || gv <= gate) flags |= 1u << (4 * J);
#endif
        process_quad_windows_flat<K, J + 1, USE64>(fw, rw, seed, mc, gate, flags);
    }
}

// Exact canonical hash of the window that starts at byte p of a byte string held as little-endian words fwd_w[],
// its reverse complement starting at byte q of rc_w[] (any alignment; reads words (p >> 2) .. (p >> 2) + NW).
template <int K>
AC_HD uint64_t rehash_window(const uint32_t *fwd_w, const uint32_t *rc_w, uint32_t p, uint32_t q, uint32_t seed) {
    constexpr int NW = (K + 3) / 4;
    constexpr uint32_t last_mask = (K & 3) ? ((1u << (8 * (K & 3))) - 1u) : 0xffffffffu;
    uint32_t f[NW], c[NW];
#pragma unroll
    for (int m = 0; m < NW; m++) {
        f[m] = funnel_r(fwd_w[(p >> 2) + m], fwd_w[(p >> 2) + m + 1], 8u * (p & 3u));
        c[m] = funnel_r(rc_w[(q >> 2) + m], rc_w[(q >> 2) + m + 1], 8u * (q & 3u));
    }
    f[NW - 1] &= last_mask; c[NW - 1] &= last_mask;
    const bool use_rc = rc_is_smaller<NW>(f, c);
    uint32_t key[NW];
#pragma unroll
    for (int m = 0; m < NW; m++) key[m] = use_rc ? c[m] : f[m];
    const uint64_t h = murmur3_h1<K>(key, seed);
    return K > 16 ? h : (h & 0xffffffffULL);   // Mash: 32-bit hashes for k <= 16
}

// K1's per-byte sanitiser (scalar form; the kernel uses it per byte of a word).
// valid ACGTacgt -> upper case, everything else -> 0x01.
AC_HD uint32_t sanitise_byte(uint32_t c) {
    const uint32_t u = c & 0xDFu;
    const bool ok = ((c & 0xC0u) == 0x40u) && ((0x0010008Au >> (c & 0x1Fu)) & 1u);
    return ok ? u : 0x01u;
}

}  // namespace ackmer
