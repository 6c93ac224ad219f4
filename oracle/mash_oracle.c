/*
 * mash_oracle.c -- CPU restatement of the Mash sketch / dist path that AuriClass
 * delegates to the external `mash` binary.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it,
 * and only as the checker / the timed CPU baseline.  The product path
 * (auriclass_b200 + libauriclass_cuda.so) never links, imports or calls it.
 *
 * Provenance.  The arithmetic lives in marbl/Mash 2.x (bioconda `mash=2`,
 * /root/reference/env.yaml:6), which is NOT vendored in /root/reference.  This
 * file restates its published algorithm and is pinned
 * against the reference's own fixtures (tests/test_oracle.py):
 *   - tests/data/ref_sketch.msh       : bit-exact sorted hash lists (k=27, s=50000)
 *   - tests/test_correct_workflow.py:18-38,99,105 : 48451/48476, 9.55405e-06,
 *     0/50000, 48454.7, bounds text (:26)
 * Call sites being replaced: auriclass/classes.py:92-104 (dist), :305-318
 * (bounds), :576-596 (sketch -r -m), :696-713 (sketch).
 *
 * Parity status: 64-bit ACGT FASTA / 4-line FASTQ path is PINNED by the
 * fixtures above.  Lower-case folding, N skipping, 32-bit mode (k<=16),
 * truncation at s, p-values other than 0/1 and "Estimated coverage" follow the
 * recalled upstream semantics and are "parity unpinned" (SURVEY.md §8c).
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* MurmurHash3_x64_128 (Appleby, public domain algorithm; restated)    */
/* Mash: Sketch.cpp getHash() -> MurmurHash3_x64_128(seq,k,seed,out)   */
/* ------------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33; return x;
}
static inline uint64_t load_le64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

void orc_murmur3_x64_128(const void *key, int len, uint32_t seed, uint64_t out[2]) {
    const uint8_t *data = (const uint8_t *)key;
    const int nblocks = len / 16;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    for (int i = 0; i < nblocks; i++) {
        uint64_t k1 = load_le64(data + 16 * i), k2 = load_le64(data + 16 * i + 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    int t = len & 15;
    for (int i = t - 1; i >= 8; i--) k2 |= (uint64_t)tail[i] << (8 * (i - 8));
    if (t > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (int i = (t > 8 ? 8 : t) - 1; i >= 0; i--) k1 |= (uint64_t)tail[i] << (8 * i);
    if (t > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

/* ------------------------------------------------------------------ */
/* k-mer hashing of one record (Mash Sketch.cpp addMinHashes)          */
/* SURVEY App. A.3                                                     */
/* ------------------------------------------------------------------ */
static inline uint8_t comp_base(uint8_t c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; }
    return 'N';
}

typedef void (*hash_sink_fn)(uint64_t h, void *ctx);

/* Emits the hash of every valid canonical k-mer of seq[0..len).  seq is
 * upper-cased in place first (Mash: a-z only, preserveCase=false). Returns the
 * number of hashed windows. */
static uint64_t hash_record(uint8_t *seq, int64_t len, int k, uint32_t seed, int use64,
                            hash_sink_fn sink, void *ctx) {
    if (len < k) return 0;
    for (int64_t i = 0; i < len; i++) if (seq[i] > 96 && seq[i] < 123) seq[i] -= 32;
    uint64_t n = 0;
    int64_t bad = -1; /* index of the last non-ACGT byte seen in [0, i+k) */
    for (int64_t j = 0; j < k - 1 && j < len; j++) {
        uint8_t c = seq[j];
        if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) bad = j;
    }
    uint8_t rc[32];
    uint64_t out[2];
    for (int64_t i = 0; i + k <= len; i++) {
        uint8_t c = seq[i + k - 1];
        if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) bad = i + k - 1;
        if (bad >= i) continue;
        const uint8_t *fwd = seq + i;
        for (int j = 0; j < k; j++) rc[j] = comp_base(fwd[k - 1 - j]);
        const uint8_t *kmer = (memcmp(fwd, rc, (size_t)k) <= 0) ? fwd : rc;
        orc_murmur3_x64_128(kmer, k, seed, out);
        uint64_t h = use64 ? out[0] : (out[0] & 0xffffffffULL);
        sink(h, ctx);
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------ */
/* kseq-style record reader over an in-memory buffer */
/* ------------------------------------------------------------------ */
typedef struct {
    const uint8_t *buf; size_t n, pos;
    int last_char;
    uint8_t *seq; size_t seq_len, seq_cap;
    size_t qual_len;
} kreader;

static void kr_init(kreader *r, const uint8_t *buf, size_t n) {
    memset(r, 0, sizeof(*r)); r->buf = buf; r->n = n;
}
static void kr_free(kreader *r) { free(r->seq); r->seq = NULL; }
static inline int kr_getc(kreader *r) { return r->pos < r->n ? r->buf[r->pos++] : -1; }
static void kr_push(kreader *r, const uint8_t *p, size_t m) {
    if (r->seq_len + m + 1 > r->seq_cap) {
        size_t nc = r->seq_cap ? r->seq_cap * 2 : 256;
        while (nc < r->seq_len + m + 1) nc *= 2;
        r->seq = (uint8_t *)realloc(r->seq, nc); r->seq_cap = nc;
    }
    memcpy(r->seq + r->seq_len, p, m); r->seq_len += m;
}
/* read to end of line (consumes '\n'); returns span (without '\n', trailing '\r' kept) */
static void kr_line(kreader *r, const uint8_t **p, size_t *m) {
    size_t s = r->pos;
    const uint8_t *nl = (const uint8_t *)memchr(r->buf + s, '\n', r->n - s);
    size_t e = nl ? (size_t)(nl - r->buf) : r->n;
    *p = r->buf + s; *m = e - s;
    r->pos = nl ? e + 1 : r->n;
}
/* returns seq length >=0, -1 EOF, -2 truncated quality */
static int64_t kr_read(kreader *r) {
    int c;
    if (r->last_char == 0) {
        while ((c = kr_getc(r)) != -1 && c != '>' && c != '@') {}
        if (c == -1) return -1;
        r->last_char = c;
    }
    r->seq_len = 0; r->qual_len = 0;
    const uint8_t *p; size_t m;
    kr_line(r, &p, &m); /* name + comment */
    while ((c = kr_getc(r)) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        r->pos--; /* put the char back, take the whole line */
        kr_line(r, &p, &m);
        kr_push(r, p, m);
    }
    if (c == '>' || c == '@') r->last_char = c;
    if (c != '+') { if (c == -1) r->last_char = 0; return (int64_t)r->seq_len; }
    kr_line(r, &p, &m); /* rest of '+' line */
    if (r->pos >= r->n && m == 0 && r->seq_len > 0) { r->last_char = 0; return -2; }
    int got = 0;
    do {
        if (r->pos >= r->n) { if (!got) { r->last_char = 0; return -2; } break; }
        kr_line(r, &p, &m); r->qual_len += m; got = 1;
    } while (r->qual_len < r->seq_len);
    r->last_char = 0;
    if (r->qual_len != r->seq_len) return -2;
    return (int64_t)r->seq_len;
}

/* ------------------------------------------------------------------ */
/* Exact selection: bottom-s of {h : count(h) >= m}   */
/* ------------------------------------------------------------------ */
typedef struct { uint64_t *v; size_t n, cap; } u64vec;
static void vec_push(uint64_t h, void *ctx) {
    u64vec *v = (u64vec *)ctx;
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : (1u << 16); v->v = (uint64_t *)realloc(v->v, v->cap * 8); }
    v->v[v->n++] = h;
}
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return (x > y) - (x < y);
}
/* LSD radix sort, 8 passes of 8 bits (constant-digit passes skipped) */
static void radix_sort_u64_safe(uint64_t *a, size_t n) {
    if (n < 4096) { qsort(a, n, 8, cmp_u64); return; }
    uint64_t *src = a, *dst = (uint64_t *)malloc(n * 8), *tmp0 = dst;
    for (int pass = 0; pass < 8; pass++) {
        size_t cnt[257]; memset(cnt, 0, sizeof(cnt));
        int sh = pass * 8;
        for (size_t i = 0; i < n; i++) cnt[((src[i] >> sh) & 255) + 1]++;
        int skip = 0;
        for (int i = 1; i <= 256; i++) if (cnt[i] == n) { skip = 1; break; }
        if (skip) continue;
        for (int i = 0; i < 256; i++) cnt[i + 1] += cnt[i];
        for (size_t i = 0; i < n; i++) dst[cnt[(src[i] >> sh) & 255]++] = src[i];
        uint64_t *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * 8);
    free(tmp0);
}

typedef struct {
    uint64_t n_records;      /* records returned by the reader */
    uint64_t n_kept;         /* records with L >= k */
    uint64_t total_len;      /* sum of L over kept records */
    uint64_t n_kmers;        /* hashed (valid) windows */
    uint64_t n_distinct;     /* distinct hashes */
    uint64_t n_qualified;    /* distinct with count >= m */
    uint64_t n_sketch;       /* min(s, n_qualified) */
    double genome_size;      /* 2^bits * n / max_hash (App. A.5) */
    double coverage;         /* exact-count multiplicity: sum(counts)/n */
    int error;               /* 0 ok, -2 truncated record */
} orc_stats;

/* sorted all-hash vector -> sketch (hashes+counts); returns n */
static size_t select_from_sorted(const uint64_t *all, size_t n_all, int s, int m, int use64,
                                 uint64_t *out_h, uint32_t *out_c, orc_stats *st) {
    size_t n = 0; uint64_t nd = 0, nq = 0, csum = 0;
    for (size_t i = 0; i < n_all;) {
        size_t j = i + 1;
        while (j < n_all && all[j] == all[i]) j++;
        nd++;
        if ((int64_t)(j - i) >= m) {
            nq++;
            if (n < (size_t)s) { out_h[n] = all[i]; out_c[n] = (uint32_t)(j - i); csum += (j - i); n++; }
        }
        i = j;
    }
    st->n_distinct = nd; st->n_qualified = nq; st->n_sketch = n;
    if (n) {
        double space = use64 ? 18446744073709551616.0 : 4294967296.0;
        st->genome_size = space * (double)n / (double)out_h[n - 1];
        st->coverage = (double)csum / (double)n;
    } else { st->genome_size = 0; st->coverage = 0; }
    return n;
}

/* Sketch the concatenation of nbufs in-memory (inflated) files into ONE sketch
 * (mash sketch -r: all files merged; FASTA mode: call once per file with m=1).
 * out_h/out_c must hold s entries.  Returns n_sketch, or <0 on error. */
int64_t orc_sketch(const uint8_t *const *bufs, const size_t *lens, int nbufs, int k, int s, int m,
                   uint32_t seed, uint64_t *out_h, uint32_t *out_c, orc_stats *st) {
    memset(st, 0, sizeof(*st));
    int use64 = k > 16; /* 4^k > 2^32 */
    u64vec all = {0};
    for (int f = 0; f < nbufs; f++) {
        kreader r; kr_init(&r, bufs[f], lens[f]);
        int64_t l;
        while ((l = kr_read(&r)) >= 0) {
            st->n_records++;
            if (l < k) continue;
            st->n_kept++; st->total_len += (uint64_t)l;
            st->n_kmers += hash_record(r.seq, l, k, seed, use64, vec_push, &all);
        }
        kr_free(&r);
        if (l == -2) { st->error = -2; free(all.v); return -2; }
    }
    radix_sort_u64_safe(all.v, all.n);
    size_t n = select_from_sorted(all.v, all.n, s, m, use64, out_h, out_c, st);
    free(all.v);
    return (int64_t)n;
}

/* All valid-window hashes in input order (for kernel-level multiset parity). */
int64_t orc_all_hashes(const uint8_t *buf, size_t len, int k, uint32_t seed, uint64_t *out, size_t cap) {
    int use64 = k > 16;
    u64vec all = {0};
    kreader r; kr_init(&r, buf, len);
    int64_t l;
    while ((l = kr_read(&r)) >= 0) { if (l >= k) hash_record(r.seq, l, k, seed, use64, vec_push, &all); }
    kr_free(&r);
    if (l == -2) { free(all.v); return -2; }
    size_t n = all.n < cap ? all.n : cap;
    if (out && n) memcpy(out, all.v, n * 8);
    int64_t total = (int64_t)all.n;
    free(all.v);
    return total;
}

/* ------------------------------------------------------------------ */
/* Faithful streaming MinHashHeap (what `mash sketch -r -m` runs;      */
/* single thread) -- the "what AuriClass runs today" CPU baseline and  */
/* the source of Mash's order-dependent coverage figure (App. A.4).    */
/* ------------------------------------------------------------------ */
typedef struct { uint64_t *key; uint32_t *val; uint8_t *state; size_t cap, n, used; } hmap; /* state 0 empty 1 full 2 tomb */
static inline uint64_t mix_slot(uint64_t h) { return h * 0x9E3779B97F4A7C15ULL; }
static void hm_init(hmap *m, size_t cap) {
    m->cap = cap; m->n = m->used = 0;
    m->key = (uint64_t *)malloc(cap * 8); m->val = (uint32_t *)malloc(cap * 4);
    m->state = (uint8_t *)calloc(cap, 1);
}
static void hm_free(hmap *m) { free(m->key); free(m->val); free(m->state); }
static void hm_put(hmap *m, uint64_t k, uint32_t add);
static void hm_grow(hmap *m) {
    hmap o = *m;
    hm_init(m, o.n * 4 > o.cap ? o.cap * 2 : o.cap);
    for (size_t i = 0; i < o.cap; i++) if (o.state[i] == 1) hm_put(m, o.key[i], o.val[i]);
    hm_free(&o);
}
static uint32_t hm_get(const hmap *m, uint64_t k) {
    size_t i = (size_t)(mix_slot(k) >> 20) & (m->cap - 1);
    while (m->state[i]) { if (m->state[i] == 1 && m->key[i] == k) return m->val[i]; i = (i + 1) & (m->cap - 1); }
    return 0;
}
static void hm_put(hmap *m, uint64_t k, uint32_t add) {
    if ((m->used + 1) * 2 > m->cap) hm_grow(m);
    size_t i = (size_t)(mix_slot(k) >> 20) & (m->cap - 1), tomb = (size_t)-1;
    while (m->state[i]) {
        if (m->state[i] == 1 && m->key[i] == k) { m->val[i] += add; return; }
        if (m->state[i] == 2 && tomb == (size_t)-1) tomb = i;
        i = (i + 1) & (m->cap - 1);
    }
    if (tomb != (size_t)-1) i = tomb; else m->used++;
    m->state[i] = 1; m->key[i] = k; m->val[i] = add; m->n++;
}
static void hm_del(hmap *m, uint64_t k) {
    size_t i = (size_t)(mix_slot(k) >> 20) & (m->cap - 1);
    while (m->state[i]) {
        if (m->state[i] == 1 && m->key[i] == k) { m->state[i] = 2; m->n--; return; }
        i = (i + 1) & (m->cap - 1);
    }
}
typedef struct { uint64_t *a; size_t n, cap; } maxheap;
static void mh_push(maxheap *h, uint64_t v) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 1024; h->a = (uint64_t *)realloc(h->a, h->cap * 8); }
    size_t i = h->n++; h->a[i] = v;
    while (i && h->a[(i - 1) / 2] < h->a[i]) { uint64_t t = h->a[i]; h->a[i] = h->a[(i - 1) / 2]; h->a[(i - 1) / 2] = t; i = (i - 1) / 2; }
}
static void mh_pop(maxheap *h) {
    h->a[0] = h->a[--h->n]; size_t i = 0;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, b = i;
        if (l < h->n && h->a[l] > h->a[b]) b = l;
        if (r < h->n && h->a[r] > h->a[b]) b = r;
        if (b == i) break;
        uint64_t t = h->a[i]; h->a[i] = h->a[b]; h->a[b] = t; i = b;
    }
}
typedef struct {
    hmap hashes, pending; maxheap q, qpending; size_t smax; uint32_t m; uint64_t mult_sum;
} mheap;
static void mheap_try(uint64_t h, void *ctx) {
    mheap *H = (mheap *)ctx;
    if (H->hashes.n < H->smax || h < H->q.a[0]) {
        uint32_t c = hm_get(&H->hashes, h);
        if (c == 0) {
            if (H->m == 1 || hm_get(&H->pending, h) == H->m - 1) {
                hm_put(&H->hashes, h, H->m); mh_push(&H->q, h); H->mult_sum += H->m;
                if (H->m > 1) hm_del(&H->pending, h);
            } else {
                if (hm_get(&H->pending, h) == 0) mh_push(&H->qpending, h);
                hm_put(&H->pending, h, 1);
            }
            if (H->hashes.n > H->smax) {
                uint64_t top = H->q.a[0];
                H->mult_sum -= hm_get(&H->hashes, top);
                hm_del(&H->hashes, top);
                mh_pop(&H->q);
                /* purge pending entries above the new top (Mash compares with the
                 * queue top before popping it; entries > old top are unreachable) */
                while (H->qpending.n && top < H->qpending.a[0]) {
                    if (hm_get(&H->pending, H->qpending.a[0])) hm_del(&H->pending, H->qpending.a[0]);
                    mh_pop(&H->qpending);
                }
            }
        } else { hm_put(&H->hashes, h, 1); H->mult_sum++; }
    }
}

/* Streaming sketch; files are consumed round-robin one record each (mash -r). */
int64_t orc_sketch_stream(const uint8_t *const *bufs, const size_t *lens, int nbufs, int k, int s, int m,
                          uint32_t seed, uint64_t *out_h, uint32_t *out_c, orc_stats *st) {
    memset(st, 0, sizeof(*st));
    int use64 = k > 16;
    mheap H; memset(&H, 0, sizeof(H));
    hm_init(&H.hashes, 1u << 17); hm_init(&H.pending, 1u << 20);
    H.smax = (size_t)s; H.m = (uint32_t)m;
    kreader *rs = (kreader *)calloc((size_t)nbufs, sizeof(kreader));
    int *done = (int *)calloc((size_t)nbufs, sizeof(int));
    for (int f = 0; f < nbufs; f++) kr_init(&rs[f], bufs[f], lens[f]);
    int live = nbufs, err = 0;
    while (live > 0 && !err) {
        for (int f = 0; f < nbufs; f++) {
            if (done[f]) continue;
            int64_t l = kr_read(&rs[f]);
            if (l < 0) { done[f] = 1; live--; if (l == -2) err = 1; continue; }
            st->n_records++;
            if (l < k) continue;
            st->n_kept++; st->total_len += (uint64_t)l;
            st->n_kmers += hash_record(rs[f].seq, l, k, seed, use64, mheap_try, &H);
        }
    }
    for (int f = 0; f < nbufs; f++) kr_free(&rs[f]);
    free(rs); free(done);
    size_t n = 0;
    if (!err) {
        for (size_t i = 0; i < H.hashes.cap; i++) if (H.hashes.state[i] == 1) out_h[n++] = H.hashes.key[i];
        qsort(out_h, n, 8, cmp_u64);
        for (size_t i = 0; i < n; i++) out_c[i] = hm_get(&H.hashes, out_h[i]);
        st->n_sketch = n;
        if (n) {
            double space = use64 ? 18446744073709551616.0 : 4294967296.0;
            st->genome_size = space * (double)n / (double)out_h[n - 1];
            st->coverage = (double)H.mult_sum / (double)n;
        }
    } else st->error = -2;
    hm_free(&H.hashes); hm_free(&H.pending); free(H.q.a); free(H.qpending.a);
    return err ? -2 : (int64_t)n;
}

/* ------------------------------------------------------------------ */
/* All-core variant (OpenMP): records are indexed once, threads hash   */
/* disjoint record ranges keeping hashes <= tau, then sort+count.      */
/* tau starts speculative and is widened until >= s qualified hashes   */
/* are proven below it (same exactness argument as the GPU path).      */
/* ------------------------------------------------------------------ */
typedef struct { uint64_t tau; u64vec *v; } tau_sink;
static void tau_push(uint64_t h, void *ctx) { tau_sink *t = (tau_sink *)ctx; if (h <= t->tau) vec_push(h, t->v); }

int64_t orc_sketch_parallel(const uint8_t *const *bufs, const size_t *lens, int nbufs, int k, int s, int m,
                            uint32_t seed, int nthreads, uint64_t *out_h, uint32_t *out_c, orc_stats *st) {
    memset(st, 0, sizeof(*st));
    int use64 = k > 16;
    /* pass 0: index records (sequential, kseq semantics), copying sequences out */
    typedef struct { uint8_t *p; int64_t l; } rec;
    rec *recs = NULL; size_t nrec = 0, caprec = 0;
    for (int f = 0; f < nbufs; f++) {
        kreader r; kr_init(&r, bufs[f], lens[f]);
        int64_t l;
        while ((l = kr_read(&r)) >= 0) {
            st->n_records++;
            if (l < k) continue;
            st->n_kept++; st->total_len += (uint64_t)l;
            if (nrec == caprec) { caprec = caprec ? caprec * 2 : (1u << 16); recs = (rec *)realloc(recs, caprec * sizeof(rec)); }
            recs[nrec].p = (uint8_t *)malloc((size_t)l); memcpy(recs[nrec].p, r.seq, (size_t)l); recs[nrec].l = l; nrec++;
        }
        kr_free(&r);
        if (l == -2) { st->error = -2; for (size_t i = 0; i < nrec; i++) free(recs[i].p); free(recs); return -2; }
    }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    uint64_t space_max = use64 ? UINT64_MAX : 0xffffffffULL;
    uint64_t tau = st->total_len > (uint64_t)64 * (uint64_t)s ? space_max / 64 : space_max;
    size_t n = 0;
    for (;;) {
        u64vec *vs = (u64vec *)calloc((size_t)nthreads, sizeof(u64vec));
        uint64_t nk = 0;
#pragma omp parallel num_threads(nthreads) reduction(+ : nk)
        {
#ifdef _OPENMP
            int t = omp_get_thread_num();
#else
            int t = 0;
#endif
            tau_sink ts = {tau, &vs[t]};
#pragma omp for schedule(dynamic, 256)
            for (size_t i = 0; i < nrec; i++) nk += hash_record(recs[i].p, recs[i].l, k, seed, use64, tau_push, &ts);
        }
        st->n_kmers = nk;
        size_t tot = 0; for (int t = 0; t < nthreads; t++) tot += vs[t].n;
        uint64_t *all = (uint64_t *)malloc((tot ? tot : 1) * 8); size_t o = 0;
        for (int t = 0; t < nthreads; t++) { if (vs[t].n) memcpy(all + o, vs[t].v, vs[t].n * 8); o += vs[t].n; free(vs[t].v); }
        free(vs);
        radix_sort_u64_safe(all, tot);
        n = select_from_sorted(all, tot, s, m, use64, out_h, out_c, st);
        free(all);
        if (st->n_qualified >= (uint64_t)s || tau == space_max) break;
        tau = tau > space_max / 16 ? space_max : tau * 16; /* speculation failed: widen, redo */
    }
    for (size_t i = 0; i < nrec; i++) free(recs[i].p);
    free(recs);
    return (int64_t)n;
}

/* ------------------------------------------------------------------ */
/* Distance + p-value  (Mash CommandDistance.cpp compareSketches,      */
/* pValue)                                            */
/* ------------------------------------------------------------------ */
static double stirlerr(double n) {
    /* lgamma(n+1) - ((n+0.5)ln n - n + 0.5 ln 2pi), Loader's saddle-point term */
    if (n <= 15.0) {
        double f = 1.0; for (int i = 2; i <= (int)n; i++) f *= i; /* exact */
        if (n == 0) return 0.0;
        return log(f) - ((n + 0.5) * log(n) - n + 0.918938533204672741780329736406);
    }
    double nn = n * n;
    if (n > 500) return (1.0 / 12 - (1.0 / 360) / nn) / n;
    if (n > 80) return (1.0 / 12 - (1.0 / 360 - (1.0 / 1260) / nn) / nn) / n;
    if (n > 35) return (1.0 / 12 - (1.0 / 360 - (1.0 / 1260 - (1.0 / 1680) / nn) / nn) / nn) / n;
    return (1.0 / 12 - (1.0 / 360 - (1.0 / 1260 - (1.0 / 1680 - (1.0 / 1188) / nn) / nn) / nn) / nn) / n;
}
static double bd0(double x, double np) {
    /* x ln(x/np) + np - x, evaluated without cancellation */
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np), s = (x - np) * v, ej = 2 * x * v; v = v * v;
        for (int j = 1; j < 1000; j++) { ej *= v; double s1 = s + ej / (2 * j + 1); if (s1 == s) return s1; s = s1; }
    }
    return x * log(x / np) + np - x;
}
/* log of binomial pmf, relative accuracy ~1e-15 in the pmf */
static double log_dbinom(double x, double n, double p) {
    double q = 1 - p;
    if (x == 0) return n * log1p(-p);
    if (x == n) return n * log(p);
    double lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q);
    double lf = 1.837877066409345483560659472811 + log(x) + log1p(-x / n); /* ln(2pi x (n-x)/n) */
    return lc - 0.5 * lf;
}
/* P(X > kk), X ~ Binomial(n, p)  ==  gsl_cdf_binomial_Q(kk, p, n) */
double orc_binomial_Q(uint64_t kk, double p, uint64_t n) {
    if (kk >= n) return 0.0;
    if (p <= 0) return 0.0;
    if (p >= 1) return 1.0;
    double a = (double)kk + 1, N = (double)n, odds = p / (1 - p);
    if (a > N * p) { /* sum the upper tail upward from its largest term */
        double lt = log_dbinom(a, N, p);
        if (lt < -745.2) return 0.0; /* below the smallest subnormal */
        double t = exp(lt), sum = t;
        for (double i = a; i < N; i += 1) { t *= (N - i) / (i + 1) * odds; sum += t; if (t < sum * 1e-18) break; }
        return sum;
    }
    /* kk below the mean: 1 - lower tail, summed downward from kk */
    double t = exp(log_dbinom((double)kk, N, p)), sum = t;
    for (double i = (double)kk; i > 0; i -= 1) { t *= i / (N - i + 1) / odds; sum += t; if (t < sum * 1e-18) break; }
    return 1.0 - sum;
}

typedef struct { uint64_t common, denom; double distance, pvalue; } orc_dist_t;

void orc_dist(const uint64_t *ref, uint64_t n_ref, const uint64_t *qry, uint64_t n_qry, uint64_t sketch_size,
              int k, uint64_t len_ref, uint64_t len_qry, orc_dist_t *out) {
    uint64_t i = 0, j = 0, common = 0, denom = 0;
    while (denom < sketch_size && i < n_ref && j < n_qry) {
        if (ref[i] < qry[j]) i++;
        else if (qry[j] < ref[i]) j++;
        else { i++; j++; common++; }
        denom++;
    }
    if (denom < sketch_size) {
        if (i < n_ref) denom += n_ref - i;
        if (j < n_qry) denom += n_qry - j;
        if (denom > sketch_size) denom = sketch_size;
    }
    double jac = denom ? (double)common / (double)denom : 0.0, d;
    if (common == denom) d = 0; else if (common == 0) d = 1.0; else d = -log(2 * jac / (1.0 + jac)) / k;
    double pv;
    if (common == 0) pv = 1.0;
    else {
        double space = pow(4.0, k);
        double pX = 1.0 / (1.0 + space / (double)len_ref), pY = 1.0 / (1.0 + space / (double)len_qry);
        double r = pX * pY / (pX + pY - pX * pY);
        pv = orc_binomial_Q(common - 1, r, denom);
    }
    out->common = common; out->denom = denom; out->distance = d; out->pvalue = pv;
}

/* ------------------------------------------------------------------ */
/* mash bounds table (CommandBounds.cpp)              */
/* ------------------------------------------------------------------ */
/* smallest x >= 0 with BinomCDF(x; n, p) > q2, capped at n */
static uint64_t binom_quantile_gt(uint64_t n, double p, double q2) {
    double N = (double)n;
    if (p >= 1.0) return n;
    double t = exp(N * log1p(-p)), cdf = t, odds = p / (1 - p);
    if (t > 0) {
        uint64_t x = 0;
        while (!(cdf > q2) && x < n) { t *= (N - (double)x) / ((double)x + 1) * odds; x++; cdf += t; }
        return x;
    }
    /* pmf(0) underflows: start near the lower tail where pmf ~ 1e-300 */
    /* (everything below that point contributes < 1e-290 to the CDF) */
    double lo = 0, hi = floor(N * p);
    while (hi - lo > 1) { double mid = floor((lo + hi) / 2); if (log_dbinom(mid, N, p) < -690) lo = mid; else hi = mid; }
    double x0 = lo;
    t = exp(log_dbinom(x0, N, p)); cdf = t;
    uint64_t x = (uint64_t)x0;
    while (!(cdf > q2) && x < n) { t *= (N - (double)x) / ((double)x + 1) * odds; x++; cdf += t; }
    return x;
}
static int fmt_g(char *b, double v) { return sprintf(b, "%g", v); }

/* Writes Mash's `mash bounds -k K -p P` stdout into buf; returns length. */
int orc_bounds_text(int k, double prob, char *buf, size_t cap) {
    static const int sizes[9] = {100, 500, 1000, 5000, 10000, 50000, 100000, 500000, 1000000};
    static const double dists[8] = {0.05, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4};
    char *o = buf; (void)cap;
    double q2 = (1.0 - prob) / 2.0;
    o += sprintf(o, "\nParameters (run with -h for details):\n   k:   %d\n   p:   ", k); o += fmt_g(o, prob); o += sprintf(o, "\n\n");
    for (int tbl = 0; tbl < 2; tbl++) {
        o += sprintf(o, "\t%s\n", tbl == 0 ? "Mash distance" : "Screen distance");
        o += sprintf(o, "Sketch");
        for (int c = 0; c < 8; c++) { *o++ = '\t'; o += fmt_g(o, dists[c]); }
        *o++ = '\n';
        for (int r = 0; r < 9; r++) {
            o += sprintf(o, "%d", sizes[r]);
            for (int c = 0; c < 8; c++) {
                double d = dists[c], m2j = tbl == 0 ? 1.0 / (2.0 * exp(k * d) - 1.0) : pow(1.0 - d, k);
                uint64_t x = binom_quantile_gt((uint64_t)sizes[r], m2j, q2);
                double je = (double)x / sizes[r], j2m;
                if (tbl == 0) j2m = -1.0 / k * log(2.0 * je / (1.0 + je)); else j2m = 1.0 - pow(je, 1.0 / k);
                *o++ = '\t'; o += fmt_g(o, j2m - d);
            }
            *o++ = '\n';
        }
        *o++ = '\n';
    }
    *o = 0;
    return (int)(o - buf);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
