// hostcheck.cpp -- CPU harness around kmer_core.cuh for tests/test_core_host.py.
//
// TEST INFRASTRUCTURE.  Runs the *same* __host__ __device__ arithmetic that kernel
// K2 executes (slices of the sanitised base stream, canonical strand choice,
// MurmurHash3_x64_128 h1) on the CPU so it can be compared with the oracle without
// a GPU.  It is not part of libauriclass_cuda.so and nothing in the product path
// loads it.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "kmer_core.cuh"

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
void hc_sanitise(const uint8_t *in, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)ackmer::sanitise_byte(in[i]);
}
}
