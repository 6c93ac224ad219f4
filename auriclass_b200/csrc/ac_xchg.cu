// ac_xchg.cu -- the hash-range exchange of the read-block sharded sketch as kernels over NVLink peer memory.
//
// Replaces, for one box of P GPUs (one process per GPU), what a single `mash sketch -r -m` process does with all
// the reads in one address space (auriclass/classes.py:576-596): the per-rank count tables are merged so that the
// min-copies counts are exact (BASELINE.json configs[2]).
//
// Every rank owns a *window* (one cudaMalloc, opened by the peers through CUDA IPC) that the other ranks write into
// with plain stores over NVLink; a rank only ever spins on flags in its OWN memory.  One exchange =
//   post tau_r to everybody -> wait for P taus, tau_cut = min, range bounds
//   -> scatter the local table entries <= tau_cut by owner into a staging area -> push every owner's share into its
//      inbox[rank] (coalesced 16-byte stores), then the count, then the flag
//   -> wait for P inboxes -> add the counts up in the count table (the qualified keys fall out of it)
//   -> (host learns one number: how many qualified keys this owner has) sort them
//   -> push the first <= s (hash, count) pairs + a header to everybody's gather[rank] -> wait for P of them
//   -> the ordered concatenation's first s entries are the sketch, written straight into the ac_sketch.
// No collective library call, no host round trip between the kernels of a phase; flags carry the exchange's epoch so
// nothing is ever cleared.  A spin gives up after X_TIMEOUT_NS and reports it: a lost peer is an error, not a hang.
#include "ac_common.cuh"
#include "ac_xchg.cuh"

namespace ac {

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 global_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin on a flag in local memory until it reaches `epoch` (flags only grow); false = timed out
__device__ bool wait_flag(const u64 *p, u64 epoch) {
    const u64 t0 = global_ns();
    while (ld_acquire_sys(p) < epoch) {
        __nanosleep(200);
        if (global_ns() - t0 > X_TIMEOUT_NS) return false;
    }
    return true;
}

__device__ __forceinline__ XCtl *ctl_of(unsigned char *base) { return reinterpret_cast<XCtl *>(base); }

// ---- phase 1 ---------------------------------------------------------------------------------------------------
// tau_r (from the job state, or "everything" when this rank has no table) and the error word go to every rank; then
// tau_cut = min over ranks, range bounds b[r] = ceil(r * (tau_cut + 1) / P), cursors cleared.  (One launch: the post only
// depends on this rank's own state.)
__global__ void kx_wait_tau(XPeers peers, unsigned char *win, XLayout L, u64 epoch, XLocal *__restrict__ xl,
                            const JobDev *__restrict__ jd, int have_table, u64 my_err) {
    const int r = threadIdx.x;
    if (r < (int)L.P) {
        XCtl *pc = ctl_of(peers.base[r]);
        pc->tau_val[L.rank] = have_table ? jd->tau : ~0ULL;
        pc->tau_err[L.rank] = my_err;
        __threadfence_system();
        st_release_sys(&pc->tau_flag[L.rank], epoch);
    }
    XCtl *c = ctl_of(win);
    u64 tau = ~0ULL;
    uint32_t err = 0;
    if (r < (int)L.P) {
        if (!wait_flag(&c->tau_flag[r], epoch)) err |= XERR_TIMEOUT;
        else { tau = c->tau_val[r]; if (c->tau_err[r]) err |= XERR_PEER; }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const u64 o = __shfl_xor_sync(FULL, tau, d);
        tau = o < tau ? o : tau;
        err |= __shfl_xor_sync(FULL, err, d);
    }
    if (r < X_MAX) {
        const unsigned __int128 span = (unsigned __int128)tau + 1;
        xl->bound[r] = r < (int)L.P ? (u64)((span * (unsigned)r + L.P - 1) / L.P) : ~0ULL;
        xl->cursor[r] = 0; xl->push_done[r] = 0; xl->in_n[r] = 0; xl->take[r] = 0;
    }
    if (r == 0) { xl->tau_cut = tau; xl->err = err; xl->n_in_total = 0; xl->n_final = 0; }
}

__device__ __forceinline__ int owner_of(const u64 *s_bound, int P, u64 x) {
    int r = 0;
    for (int j = 1; j < P; j++) r += (x >= s_bound[j]) ? 1 : 0;
    return r;
}

constexpr int XS_THREADS = 256;
constexpr int XS_CTAS = 148 * 4;

// local table entries <= tau_cut -> staging[owner][...]; a CTA reserves its share of every owner's area with one
// atomic per owner.  TABLE: the open-addressing count table; else the sorted (tab_hash, tab_cnt) arrays.
template <bool TABLE>
__global__ void __launch_bounds__(XS_THREADS) kx_scatter(const CountTable ct, const u64 *__restrict__ tab_hash,
                                                         const uint32_t *__restrict__ tab_cnt, const JobDev *__restrict__ jd,
                                                         XLayout L, XLocal *__restrict__ xl, u64 *__restrict__ st_keys,
                                                         uint32_t *__restrict__ st_cnt) {
    __shared__ uint32_t s_cnt[X_MAX], s_base[X_MAX];
    __shared__ u64 s_bound[X_MAX];
    if (threadIdx.x < X_MAX) { s_cnt[threadIdx.x] = 0; s_bound[threadIdx.x] = xl->bound[threadIdx.x]; }
    __syncthreads();
    const u64 tau_cut = xl->tau_cut;
    const u64 n = TABLE ? (u64)ct.mask + 1 : (u64)jd->n_table;
    const u64 per = (n + gridDim.x - 1) / gridDim.x;
    const u64 lo = (u64)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    for (u64 i = lo + threadIdx.x; i < hi; i += XS_THREADS) {
        const u64 h = TABLE ? __ldcg(ct.key + i) : tab_hash[i];
        if ((!TABLE || h != CT_EMPTY) && h <= tau_cut) atomicAdd(&s_cnt[owner_of(s_bound, L.P, h)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < L.P) {
        const uint32_t c = s_cnt[threadIdx.x];
        s_base[threadIdx.x] = c ? atomicAdd(&xl->cursor[threadIdx.x], c) : 0u;
        s_cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    for (u64 i = lo + threadIdx.x; i < hi; i += XS_THREADS) {
        const u64 h = TABLE ? __ldcg(ct.key + i) : tab_hash[i];
        if ((!TABLE || h != CT_EMPTY) && h <= tau_cut) {
            const int r = owner_of(s_bound, L.P, h);
            const uint32_t pos = s_base[r] + atomicAdd(&s_cnt[r], 1u);
            if (pos < L.cap_in) {
                st_keys[(size_t)r * L.cap_in + pos] = h;
                st_cnt[(size_t)r * L.cap_in + pos] = TABLE ? __ldcg(ct.cnt + i) : tab_cnt[i];
            }   // (beyond the area: kx_push reports the overflow to the owner)
        }
    }
}

// staging[dst] -> inbox[rank] of rank dst, X_PUSH_CTAS CTAs per destination; the last one to finish posts count + flag
__global__ void __launch_bounds__(256) kx_push(XPeers peers, XLayout L, u64 epoch, XLocal *__restrict__ xl,
                                               const u64 *__restrict__ st_keys, const uint32_t *__restrict__ st_cnt) {
    const uint32_t dst = blockIdx.x / X_PUSH_CTAS, part = blockIdx.x % X_PUSH_CTAS;
    const uint32_t have = xl->cursor[dst];
    const uint32_t n = have < L.cap_in ? have : L.cap_in;
    unsigned char *pb = peers.base[dst];
    u64 *ik = reinterpret_cast<u64 *>(pb + L.off_in_keys) + (size_t)L.rank * L.cap_in;
    uint32_t *ic = reinterpret_cast<uint32_t *>(pb + L.off_in_cnt) + (size_t)L.rank * L.cap_in;
    const u64 *sk = st_keys + (size_t)dst * L.cap_in;
    const uint32_t *sc = st_cnt + (size_t)dst * L.cap_in;
    // 16-byte vectors: 2 keys / 4 counts each (all areas are 16-byte aligned: cap_in is a multiple of 4)
    const uint32_t nk2 = (n + 1) / 2, nc4 = (n + 3) / 4;
    for (uint32_t i = part * 256 + threadIdx.x; i < nk2; i += X_PUSH_CTAS * 256)
        reinterpret_cast<uint4 *>(ik)[i] = reinterpret_cast<const uint4 *>(sk)[i];
    for (uint32_t i = part * 256 + threadIdx.x; i < nc4; i += X_PUSH_CTAS * 256)
        reinterpret_cast<uint4 *>(ic)[i] = reinterpret_cast<const uint4 *>(sc)[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&xl->push_done[dst], 1u) == X_PUSH_CTAS - 1) {
            __threadfence_system();
            XCtl *c = ctl_of(pb);
            c->in_n[L.rank] = have > L.cap_in ? X_OVERFLOW : (u64)n;
            if (have > L.cap_in) atomicOr(&xl->err, (uint32_t)XERR_OVERFLOW);
            __threadfence_system();
            st_release_sys(&c->in_flag[L.rank], epoch);
        }
    }
}

// (also leaves the job state the owner merge expects: empty qualified list, threshold "everything")
__global__ void kx_wait_inbox(unsigned char *win, XLayout L, u64 epoch, XLocal *__restrict__ xl, JobDev *__restrict__ jd) {
    XCtl *c = ctl_of(win);
    const int r = threadIdx.x;
    if (r == 0) { jd->tau = ~0ULL; jd->n_out = 0; jd->n_qlist = 0; jd->max_hash = 0; jd->overflow = 0; jd->n_qualified = 0; jd->n_sketch = 0; }
    uint32_t n = 0, err = 0;
    if (r < (int)L.P) {
        if (!wait_flag(&c->in_flag[r], epoch)) err |= XERR_TIMEOUT;
        else {
            const u64 v = c->in_n[r];
            if (v == X_OVERFLOW) err |= XERR_OVERFLOW; else n = (uint32_t)v;
        }
        xl->in_n[r] = n;
    }
    uint32_t tot = n;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { tot += __shfl_xor_sync(FULL, tot, d); err |= __shfl_xor_sync(FULL, err, d); }
    if (r == 0) { xl->n_in_total = tot; if (err) atomicOr(&xl->err, err); }
}

// owner merge: every (hash, count) of the P inboxes into the count table; a sum that crosses min_copies qualifies the key
__global__ void __launch_bounds__(256) kx_merge(const CountTable ct, const unsigned char *__restrict__ win, XLayout L,
                                                const XLocal *__restrict__ xl, JobDev *__restrict__ jd) {
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    for (uint32_t src = 0; src < L.P; src++) {
        const uint32_t n = xl->in_n[src];
        const u64 *ik = reinterpret_cast<const u64 *>(win + L.off_in_keys) + (size_t)src * L.cap_in;
        const uint32_t *ic = reinterpret_cast<const uint32_t *>(win + L.off_in_cnt) + (size_t)src * L.cap_in;
        for (uint32_t i = gtid; i < n; i += gstride) {
            const u64 h = __ldcg(ik + i);
            const uint32_t cn = __ldcg(ic + i);
            if (h == CT_EMPTY) { jd->overflow = 1u; continue; }   // the one value the table cannot key: the caller sorts instead
            uint32_t slot = (uint32_t)h & ct.mask;
            bool placed = false;
            for (int probe = 0; probe < CT_MAX_PROBE; probe++, slot = (slot + 1u) & ct.mask) {
                u64 cur = __ldcg(ct.key + slot);
                if (cur == CT_EMPTY) cur = atomicCAS(ct.key + slot, CT_EMPTY, h);
                if (cur == CT_EMPTY || cur == h) {
                    const uint32_t old = atomicAdd(ct.cnt + slot, cn);
                    if (old < ct.min_copies && old + cn >= ct.min_copies) {
                        atomicMax(&jd->max_hash, h);            // the sort of the qualified keys sizes its buckets with it
                        const uint32_t q = atomicAdd(&jd->n_qlist, 1u);
                        if (q < ct.qcap) ct.qlist[q] = h; else jd->overflow = 1u;
                    }
                    placed = true;
                    break;
                }
            }
            if (!placed) jd->overflow = 1u;
        }
    }
}

// fallback of the owner merge (table overflow): the inboxes back to back in (keys, vals) for the sort-based merge
__global__ void __launch_bounds__(256) kx_compact_inbox(const unsigned char *__restrict__ win, XLayout L,
                                                        const XLocal *__restrict__ xl, u64 *__restrict__ keys,
                                                        uint32_t *__restrict__ vals) {
    uint32_t base = 0;
    for (uint32_t src = 0; src < L.P; src++) {
        const uint32_t n = xl->in_n[src];
        const u64 *ik = reinterpret_cast<const u64 *>(win + L.off_in_keys) + (size_t)src * L.cap_in;
        const uint32_t *ic = reinterpret_cast<const uint32_t *>(win + L.off_in_cnt) + (size_t)src * L.cap_in;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            keys[base + i] = ik[i]; vals[base + i] = ic[i];
        }
        base += n;
    }
}

// ---- phase 2 ---------------------------------------------------------------------------------------------------
// this owner's first n_keep qualified (hash, count) pairs + header -> gather[rank] of every rank
__global__ void __launch_bounds__(256) kx_push_gather(XPeers peers, XLayout L, u64 epoch, XLocal *__restrict__ xl,
                                                      const u64 *__restrict__ keys, const uint32_t *__restrict__ cnts,
                                                      uint32_t n_keep, XHeader hdr) {
    const uint32_t dst = blockIdx.x / X_GATHER_CTAS, part = blockIdx.x % X_GATHER_CTAS;
    unsigned char *pb = peers.base[dst];
    u64 *gk = reinterpret_cast<u64 *>(pb + L.off_ga_keys) + (size_t)L.rank * L.s_cap;
    uint32_t *gc = reinterpret_cast<uint32_t *>(pb + L.off_ga_cnt) + (size_t)L.rank * L.s_cap;
    for (uint32_t i = part * 256 + threadIdx.x; i < n_keep; i += X_GATHER_CTAS * 256) { gk[i] = keys[i]; gc[i] = cnts[i]; }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&xl->push_done[dst], 1u) == X_PUSH_CTAS + X_GATHER_CTAS - 1) {   // (push_done already counts kx_push's CTAs)
            __threadfence_system();
            XCtl *c = ctl_of(pb);
            hdr.v[7] |= (u64)xl->err;                       // whatever went wrong on this rank so far travels with the header
#pragma unroll
            for (int i = 0; i < 8; i++) c->ga_hdr[L.rank][i] = hdr.v[i];
            __threadfence_system();
            st_release_sys(&c->ga_flag[L.rank], epoch);
        }
    }
}

// wait for P headers; the ordered concatenation's first s entries: take[r] of owner r; job-wide counters
__global__ void kx_wait_gather(unsigned char *win, XLayout L, u64 epoch, XLocal *__restrict__ xl, uint32_t sketch_size) {
    XCtl *c = ctl_of(win);
    const int r = threadIdx.x;
    uint32_t keep = 0, err = 0;
    u64 h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < (int)L.P) {
        if (!wait_flag(&c->ga_flag[r], epoch)) err |= XERR_TIMEOUT;
        else {
#pragma unroll
            for (int i = 0; i < 8; i++) h[i] = c->ga_hdr[r][i];
            keep = (uint32_t)h[0];
            if (h[7]) err |= (uint32_t)h[7] | XERR_PEER;
        }
    }
    // exclusive prefix of keep over the ranks (owner r holds a lower hash range than owner r + 1)
    uint32_t inc = keep;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, inc, d);
        if (r >= d) inc += t;
    }
    const uint32_t before = inc - keep;
    const uint32_t take = before >= sketch_size ? 0u : (keep < sketch_size - before ? keep : sketch_size - before);
    if (r < (int)L.P) { xl->take[r] = take; xl->take_off[r] = before < sketch_size ? before : sketch_size; }
    uint32_t n_final = take;
    u64 nq = h[1];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        n_final += __shfl_xor_sync(FULL, n_final, d);
        nq += __shfl_xor_sync(FULL, nq, d);
        err |= __shfl_xor_sync(FULL, err, d);
#pragma unroll
        for (int i = 2; i < 7; i++) h[i] += __shfl_xor_sync(FULL, h[i], d);
    }
    if (r == 0) {
        xl->n_final = n_final; xl->qualified_total = nq;
        for (int i = 0; i < 5; i++) xl->counters[i] = h[2 + i];
        if (err) atomicOr(&xl->err, err);
    }
}

// gather[r][0 .. take[r]) -> sketch[take_off[r] ..); JobDev::n_sketch for K4
__global__ void __launch_bounds__(256) kx_concat(const unsigned char *__restrict__ win, XLayout L, const XLocal *__restrict__ xl,
                                                 u64 *__restrict__ sk_hash, uint32_t *__restrict__ sk_cnt, JobDev *__restrict__ jd) {
    const uint32_t r = blockIdx.x / X_GATHER_CTAS, part = blockIdx.x % X_GATHER_CTAS;
    const uint32_t n = xl->take[r], off = xl->take_off[r];
    const u64 *gk = reinterpret_cast<const u64 *>(win + L.off_ga_keys) + (size_t)r * L.s_cap;
    const uint32_t *gc = reinterpret_cast<const uint32_t *>(win + L.off_ga_cnt) + (size_t)r * L.s_cap;
    for (uint32_t i = part * 256 + threadIdx.x; i < n; i += X_GATHER_CTAS * 256) {
        sk_hash[off + i] = __ldcg(gk + i); sk_cnt[off + i] = __ldcg(gc + i);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) jd->n_sketch = xl->n_final;
}

}  // namespace

XLayout xchg_layout(uint32_t rank, uint32_t P, uint32_t cap_in, uint32_t s_cap) {
    XLayout L;
    L.rank = rank; L.P = P; L.cap_in = cap_in & ~3u; L.s_cap = (s_cap + 3u) & ~3u;
    size_t off = (sizeof(XCtl) + 4095) & ~(size_t)4095;
    L.off_in_keys = off; off += (size_t)P * L.cap_in * 8;
    L.off_in_cnt = off;  off += (size_t)P * L.cap_in * 4;
    L.off_ga_keys = off; off += (size_t)P * L.s_cap * 8;
    L.off_ga_cnt = off;  off += (size_t)P * L.s_cap * 4;
    L.bytes = (off + 4095) & ~(size_t)4095;
    return L;
}

void launch_xchg_phase1(const LaunchCtx &lc, const XPeers &peers, unsigned char *win, const XLayout &L, u64 epoch, XLocal *xl,
                        JobDev *jd, int have_table, int table_mode, u64 err, const CountTable &ct_local, const SelectBufs &sb,
                        const CountTable &ct_merge) {
    kx_wait_tau<<<1, 32, 0, lc.stream>>>(peers, win, L, epoch, xl, jd, have_table, err);
    ++*lc.launches;
    if (have_table) {
        if (table_mode) kx_scatter<true><<<XS_CTAS, XS_THREADS, 0, lc.stream>>>(ct_local, nullptr, nullptr, jd, L, xl, sb.keys_a, sb.vals_a);
        else kx_scatter<false><<<XS_CTAS, XS_THREADS, 0, lc.stream>>>(ct_local, sb.tab_hash, sb.tab_cnt, jd, L, xl, sb.keys_a, sb.vals_a);
        ++*lc.launches;
    }
    kx_push<<<L.P * X_PUSH_CTAS, 256, 0, lc.stream>>>(peers, L, epoch, xl, sb.keys_a, sb.vals_a);
    // the owner's table is cleared while the other ranks' shares are still on their way
    const size_t slots = (size_t)ct_merge.mask + 1;
    cudaMemsetAsync(ct_merge.key, 0xff, slots * 8, lc.stream);
    cudaMemsetAsync(ct_merge.cnt, 0, slots * 4, lc.stream);
    kx_wait_inbox<<<1, 32, 0, lc.stream>>>(win, L, epoch, xl, jd);
    kx_merge<<<148 * 4, 256, 0, lc.stream>>>(ct_merge, win, L, xl, jd);
    *lc.launches += 3;
}

void launch_xchg_compact_inbox(const LaunchCtx &lc, unsigned char *win, const XLayout &L, XLocal *xl, const SelectBufs &sb) {
    kx_compact_inbox<<<148, 256, 0, lc.stream>>>(win, L, xl, sb.keys_a, sb.vals_a);
    ++*lc.launches;
}

void launch_xchg_phase2(const LaunchCtx &lc, const XPeers &peers, unsigned char *win, const XLayout &L, u64 epoch, XLocal *xl,
                        JobDev *jd, const u64 *keys, const uint32_t *cnts, uint32_t n_keep, const XHeader &hdr, uint32_t sketch_size,
                        u64 *sk_hash, uint32_t *sk_cnt) {
    kx_push_gather<<<L.P * X_GATHER_CTAS, 256, 0, lc.stream>>>(peers, L, epoch, xl, keys, cnts, n_keep, hdr);
    kx_wait_gather<<<1, 32, 0, lc.stream>>>(win, L, epoch, xl, sketch_size);
    kx_concat<<<L.P * X_GATHER_CTAS, 256, 0, lc.stream>>>(win, L, xl, sk_hash, sk_cnt, jd);
    *lc.launches += 3;
}

}  // namespace ac
