// hostcheck.cpp -- CPU harness around kmer_core.cuh for tests/test_host.py (test infrastructure, not product).
//
// TEST INFRASTRUCTURE.  Runs the *same* __host__ __device__ arithmetic that kernel
// K2 executes (slices of the sanitised base stream, canonical strand choice,
// MurmurHash3_x64_128 h1) on the CPU so it can be compared with the oracle without
// a GPU.  It is not part of libauriclass_cuda.so and nothing in the product path
// loads it.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "../../auriclass_b200/csrc/kmer_core.cuh"

namespace {
struct Collect {
    std::vector<uint64_t> *out;
    uint64_t base;
    void operator()(uint32_t vb, uint32_t bit, uint64_t h) { if (vb & bit) out->push_back(h); }
};

template <int K>
long run_k(const uint8_t *z, size_t n, uint64_t *out, size_t cap) {
    constexpr int W = 32;
    using C = ackmer::RunCfg<K, W>;
    std::vector<uint64_t> hs;
    const long long n_win = (long long)n - K + 1;
    for (long long rs = 0; rs < n_win; rs += W) {
        uint8_t buf[4 * C::RW];
        memset(buf, 0, sizeof(buf));
        const size_t take = (size_t)rs + sizeof(buf) <= n ? sizeof(buf) : n - (size_t)rs;
        memcpy(buf, z + rs, take);
        uint32_t r[C::RW];
        memcpy(r, buf, sizeof(buf));
        long long whi = n_win - rs;
        if (whi > W) whi = W;
        const uint32_t win_ok = whi >= 32 ? 0xffffffffu : ((1u << whi) - 1u);
        Collect c{&hs, 0};
        ackmer::process_run<K, W>(r, win_ok, 42u, c);
    }
    const size_t m = hs.size() < cap ? hs.size() : cap;
    if (out && m) memcpy(out, hs.data(), m * 8);
    return (long)hs.size();
}
// The quad layout of kernel K2 (QuadCfg / preshift / process_quad_windows), emulated lane by lane over
// one big "tile": forward bytes, reverse-complement array rc[q] = comp(fwd[TB-1-q]), groups of 32
// consecutive windows, lane c of the quad hashing windows g + c + 4j.
struct CollectPos {
    std::vector<std::pair<uint64_t, uint64_t>> *out;
    uint64_t base;  // stream position of the lane's window 0
    void operator()(uint32_t vb, uint32_t bit, uint64_t p1, uint64_t p2) {
        const uint64_t h = ackmer::murmur3_close(p1, p2);
        if (vb & bit) { int j = 0; while (!((bit >> (4 * j)) & 1u)) j++; out->push_back({base + 4 * (uint64_t)j, h}); }
    }
};

template <int K>
long run_quad_k(const uint8_t *z, size_t n, uint64_t *out, size_t cap) {
    using Q = ackmer::QuadCfg<K>;
    const long long n_win = (long long)n - K + 1;
    if (n_win <= 0) return 0;
    const size_t TB = ((n + 95) / 4) * 4;                 // multiple of 4, room for the window halo
    std::vector<uint8_t> fwd(TB + 128, 0), rc(TB + 128, 0);
    memcpy(fwd.data(), z, n);
    for (size_t w = 0; w < TB / 4; w++) {                 // word-wise, as the kernel prologue does
        uint32_t x; memcpy(&x, &fwd[4 * w], 4);
        const uint32_t y = ackmer::byte_perm(ackmer::comp4(x), 0, 0x0123);
        memcpy(&rc[TB - 4 * (w + 1)], &y, 4);
    }
    // window validity
    std::vector<uint8_t> okb(n + K + 64, 0);
    for (size_t i = 0; i < n; i++) okb[i] = (z[i] >> 6) & 1;
    std::vector<std::pair<uint64_t, uint64_t>> hs;
    for (long long g = 0; g < n_win; g += 32) {
        for (int c = 0; c < 4; c++) {
            uint32_t vb = 0;
            for (int j = 0; j < Q::G; j++) {
                const long long p = g + c + 4 * j;
                bool ok = p < n_win;
                for (int t = 0; ok && t < K; t++) ok = okb[p + t];
                if (ok) vb |= 1u << (4 * j);
            }
            const size_t A = (size_t)g + c;
            uint32_t in[Q::NL], fw[Q::NS], rw[Q::NS];
            memcpy(in, &fwd[A & ~(size_t)3], sizeof(in));
            ackmer::preshift<Q::NS>(in, (uint32_t)(A & 3), fw);
            const size_t QL = TB - K - A - 4 * (Q::G - 1);
            memcpy(in, &rc[QL & ~(size_t)3], sizeof(in));
            ackmer::preshift<Q::NS>(in, (uint32_t)(QL & 3), rw);
            CollectPos col{&hs, (uint64_t)A};
            const ackmer::MurmurConsts mc;
            ackmer::process_quad_windows<K, 0>(fw, rw, vb, 42u, mc, col);
        }
    }
    std::sort(hs.begin(), hs.end());
    const size_t m = hs.size() < cap ? hs.size() : cap;
    for (size_t i = 0; i < m && out; i++) out[i] = hs[i].second;
    return (long)hs.size();
}

// The flat hot loop of kernel K2 (process_quad_windows_flat + rehash_window), emulated lane by lane: only flagged
// windows are hashed exactly, everything else is dropped as "above tau".  Writes the hashes <= tau in stream order;
// *n_flagged = windows that were flagged.
template <int K>
long run_quad_flat_k(const uint8_t *z, size_t n, uint64_t tau, uint64_t *out, size_t cap, long *n_flagged) {
    using Q = ackmer::QuadCfg<K>;
    const long long n_win = (long long)n - K + 1;
    if (n_flagged) *n_flagged = 0;
    if (n_win <= 0) return 0;
    const size_t TB = ((n + 95) / 4) * 4;
    std::vector<uint32_t> fwd_w(TB / 4 + 32, 0), rc_w(TB / 4 + 32, 0);
    uint8_t *fwd = reinterpret_cast<uint8_t *>(fwd_w.data()), *rc = reinterpret_cast<uint8_t *>(rc_w.data());
    memcpy(fwd, z, n);
    for (size_t w = 0; w < TB / 4; w++) rc_w[TB / 4 - 1 - w] = ackmer::byte_perm(ackmer::comp4(fwd_w[w]), 0, 0x0123);
    std::vector<uint8_t> okb(n + K + 64, 0);
    for (size_t i = 0; i < n; i++) okb[i] = (z[i] >> 6) & 1;
    const uint32_t gate = ackmer::gate_of_tau(tau, K > 16);
    std::vector<std::pair<uint64_t, uint64_t>> hs;
    long flagged = 0;
    for (long long g = 0; g < n_win; g += 32) {
        for (int c = 0; c < 4; c++) {
            uint32_t vb = 0;
            for (int j = 0; j < Q::G; j++) {
                const long long p = g + c + 4 * j;
                bool ok = p < n_win;
                for (int t = 0; ok && t < K; t++) ok = okb[p + t];
                if (ok) vb |= 1u << (4 * j);
            }
            const size_t A = (size_t)g + c;
            uint32_t in[Q::NL], fw[Q::NS], rw[Q::NS];
            memcpy(in, &fwd[A & ~(size_t)3], sizeof(in));
            ackmer::preshift<Q::NS>(in, (uint32_t)(A & 3), fw);
            const size_t QL = TB - K - A - 4 * (Q::G - 1);
            memcpy(in, &rc[QL & ~(size_t)3], sizeof(in));
            ackmer::preshift<Q::NS>(in, (uint32_t)(QL & 3), rw);
            const ackmer::MurmurConsts mc;
            uint32_t flags = 0;
            ackmer::process_quad_windows_flat<K, 0, (K > 16)>(fw, rw, 42u, mc, gate, flags);
            flags &= vb;
            for (int j = 0; j < Q::G; j++) {
                if (!((flags >> (4 * j)) & 1u)) continue;
                flagged++;
                const uint32_t p = (uint32_t)(A + 4 * j);
                const uint64_t h = ackmer::rehash_window<K>(fwd_w.data(), rc_w.data(), p, (uint32_t)(TB - K) - p, 42u);
                if (h <= tau) hs.push_back({p, h});
            }
        }
    }
    std::sort(hs.begin(), hs.end());
    const size_t m = hs.size() < cap ? hs.size() : cap;
    for (size_t i = 0; i < m && out; i++) out[i] = hs[i].second;
    if (n_flagged) *n_flagged = flagged;
    return (long)hs.size();
}
}  // namespace

#define HC_CASE(KK) case KK: return run_k<KK>(z, n, out, cap);

extern "C" {
// z: sanitised stream (A,C,G,T / 0x00 separator / 0x01 invalid).  Writes h1 of every valid
// window in stream order; returns their number (or -1 for unsupported k).
long hc_hash_stream(const uint8_t *z, size_t n, int k, uint64_t *out, size_t cap) {
    switch (k) {
        HC_CASE(1) HC_CASE(2) HC_CASE(3) HC_CASE(4) HC_CASE(5) HC_CASE(6) HC_CASE(7) HC_CASE(8)
        HC_CASE(9) HC_CASE(10) HC_CASE(11) HC_CASE(12) HC_CASE(13) HC_CASE(14) HC_CASE(15) HC_CASE(16)
        HC_CASE(17) HC_CASE(18) HC_CASE(19) HC_CASE(20) HC_CASE(21) HC_CASE(22) HC_CASE(23) HC_CASE(24)
        HC_CASE(25) HC_CASE(26) HC_CASE(27) HC_CASE(28) HC_CASE(29) HC_CASE(30) HC_CASE(31) HC_CASE(32)
    }
    return -1;
}
#undef HC_CASE
#define HC_CASE(KK) case KK: return run_quad_k<KK>(z, n, out, cap);
// same contract as hc_hash_stream, through the quad layout of kernel K2
long hc_hash_stream_quad(const uint8_t *z, size_t n, int k, uint64_t *out, size_t cap) {
    switch (k) {
        HC_CASE(1) HC_CASE(2) HC_CASE(3) HC_CASE(4) HC_CASE(5) HC_CASE(6) HC_CASE(7) HC_CASE(8)
        HC_CASE(9) HC_CASE(10) HC_CASE(11) HC_CASE(12) HC_CASE(13) HC_CASE(14) HC_CASE(15) HC_CASE(16)
        HC_CASE(17) HC_CASE(18) HC_CASE(19) HC_CASE(20) HC_CASE(21) HC_CASE(22) HC_CASE(23) HC_CASE(24)
        HC_CASE(25) HC_CASE(26) HC_CASE(27) HC_CASE(28) HC_CASE(29) HC_CASE(30) HC_CASE(31) HC_CASE(32)
    }
    return -1;
}
#undef HC_CASE
#define HC_CASE(KK) case KK: return run_quad_flat_k<KK>(z, n, tau, out, cap, n_flagged);
// hashes <= tau of the stream, in stream order, through the flat hot loop of kernel K2 (flag, then re-hash exactly)
long hc_hash_stream_flat(const uint8_t *z, size_t n, int k, uint64_t tau, uint64_t *out, size_t cap, long *n_flagged) {
    switch (k) {
        HC_CASE(1) HC_CASE(2) HC_CASE(3) HC_CASE(4) HC_CASE(5) HC_CASE(6) HC_CASE(7) HC_CASE(8)
        HC_CASE(9) HC_CASE(10) HC_CASE(11) HC_CASE(12) HC_CASE(13) HC_CASE(14) HC_CASE(15) HC_CASE(16)
        HC_CASE(17) HC_CASE(18) HC_CASE(19) HC_CASE(20) HC_CASE(21) HC_CASE(22) HC_CASE(23) HC_CASE(24)
        HC_CASE(25) HC_CASE(26) HC_CASE(27) HC_CASE(28) HC_CASE(29) HC_CASE(30) HC_CASE(31) HC_CASE(32)
    }
    return -1;
}
void hc_sanitise(const uint8_t *in, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)ackmer::sanitise_byte(in[i]);
}
}
