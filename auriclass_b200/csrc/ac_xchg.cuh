// ac_xchg.cuh -- structures of the peer-memory hash-range exchange (ac_xchg.cu; host side in ac_api.cu).
#pragma once
#include "ac_common.cuh"

namespace ac {

constexpr int X_MAX = 16;                                  // ranks of one exchange (GPUs of one box)
constexpr int X_PUSH_CTAS = 8;                             // CTAs that copy one rank's share into its owner's inbox
constexpr int X_GATHER_CTAS = 2;                           // ... one owner's sketch prefix into a rank's gather area
constexpr unsigned long long X_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;   // a spin gives up after 20 s
constexpr unsigned long long X_OVERFLOW = ~0ULL;           // inbox count of a sender whose share did not fit
enum { XERR_TIMEOUT = 1, XERR_OVERFLOW = 2, XERR_PEER = 4 };

// Head of every rank's window.  Written by the peers with plain stores over NVLink (slot [src] by rank src only),
// read by the owner.  Flags hold the epoch of the exchange that wrote the slot and only ever grow.
struct XCtl {
    unsigned long long tau_flag[X_MAX], tau_val[X_MAX], tau_err[X_MAX];
    unsigned long long in_flag[X_MAX], in_n[X_MAX];
    unsigned long long ga_flag[X_MAX], ga_hdr[X_MAX][8];   // n_keep, n_qualified, 5 job counters, error bits
};

struct XHeader { unsigned long long v[8]; };               // passed by value

// Window = [XCtl | inbox keys P x cap_in x 8 | inbox counts P x cap_in x 4 | gather keys P x s_cap x 8 | gather counts
// P x s_cap x 4]; identical on every rank, so a peer's areas are its base pointer plus these offsets.
struct XLayout {
    uint32_t rank, P, cap_in, s_cap;
    size_t off_in_keys, off_in_cnt, off_ga_keys, off_ga_cnt, bytes;
};
struct XPeers { unsigned char *base[X_MAX]; };             // window of every rank as seen from this process

// per-rank scratch of one exchange (local device memory; read back by the host at the two sync points)
struct XLocal {
    unsigned long long tau_cut;
    unsigned long long bound[X_MAX];                       // first hash owned by rank r
    unsigned long long counters[5];                        // job-wide: records, kept records, total length, k-mers, candidates
    unsigned long long qualified_total;
    uint32_t cursor[X_MAX], push_done[X_MAX], in_n[X_MAX], take[X_MAX], take_off[X_MAX];
    uint32_t err, n_in_total, n_final, pad;
};

XLayout xchg_layout(uint32_t rank, uint32_t P, uint32_t cap_in, uint32_t s_cap);

// post tau, wait, scatter the local table by owner, push, wait for the inboxes, merge into ct_merge (all asynchronous)
void launch_xchg_phase1(const LaunchCtx &lc, const XPeers &peers, unsigned char *win, const XLayout &L,
                        unsigned long long epoch, XLocal *xl, JobDev *jd, int have_table, int table_mode,
                        unsigned long long err, const CountTable &ct_local, const SelectBufs &sb, const CountTable &ct_merge);
void launch_xchg_compact_inbox(const LaunchCtx &lc, unsigned char *win, const XLayout &L, XLocal *xl, const SelectBufs &sb);
// push this owner's prefix to every rank, wait for all of them, write the final sketch
void launch_xchg_phase2(const LaunchCtx &lc, const XPeers &peers, unsigned char *win, const XLayout &L,
                        unsigned long long epoch, XLocal *xl, JobDev *jd, const unsigned long long *keys, const uint32_t *cnts,
                        uint32_t n_keep, const XHeader &hdr, uint32_t sketch_size, unsigned long long *sk_hash, uint32_t *sk_cnt);

}  // namespace ac
