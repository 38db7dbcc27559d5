/*
 * pw_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the *semantics* of Polarway's (Polars 0.52 fork)
 * filter -> group_by -> agg and group_by_dynamic hot path.  It is NOT part of
 * the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The CUDA product path
 * never links or calls anything in this directory.
 *
 * Parity status: PINNED — checked against the reference's own known-answer
 * tests transcribed under tests/golden/ (see tests/golden/README.md for the
 * file:line of every vector).  The reference engine itself (Rust) cannot be
 * built in this image (no cargo/rustc), so there is no oracle/_ref binary.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/crates).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- shared small enums (documented in include/polarway_b200.h) ---------- */
enum { ORC_EQ = 0, ORC_NE = 1, ORC_LT = 2, ORC_LE = 3, ORC_GT = 4, ORC_GE = 5 };
enum { ORC_SUM = 0, ORC_MEAN = 1, ORC_MIN = 2, ORC_MAX = 3, ORC_COUNT = 4,
       ORC_LEN = 5, ORC_FIRST = 6, ORC_LAST = 7,
       /* the other reductions the streaming engine pre-aggregates (reduce/convert.rs:46-150) */
       ORC_VAR = 8, ORC_STD = 9, ORC_FIRST_NN = 10, ORC_LAST_NN = 11, ORC_NULL_COUNT = 12,
       ORC_BIT_AND = 13, ORC_BIT_OR = 14, ORC_BIT_XOR = 15, ORC_ANY = 16, ORC_ALL = 17 };
/* value classes the Python wrapper widens to (exact widenings) */
enum { ORC_I64 = 0, ORC_U64 = 1, ORC_F64 = 2, ORC_F32 = 3 };
enum { ORC_CLOSED_LEFT = 0, ORC_CLOSED_RIGHT = 1, ORC_CLOSED_BOTH = 2, ORC_CLOSED_NONE = 3 };

/* =========================================================================
 * a1/a2 — predicate evaluation, null => false
 *   polars-compute/src/comparisons/mod.rs:56-76 (tot_*_kernel_broadcast)
 *   polars-compute/src/filter/mod.rs:18-28      (null mask bit => false)
 * Values are passed widened (i64 / u64 / f64); float compares follow the
 * TotalOrd rules the reference uses (NaN == NaN, NaN greater than everything:
 * polars-utils/src/total_ord.rs).
 * ========================================================================= */
static inline int tot_cmp_f64(double a, double b) {
    int an = isnan(a), bn = isnan(b);
    if (an || bn) return an - bn; /* NaN is the largest value */
    return (a > b) - (a < b);
}

int orc_predicate_mask(int vclass, const void *values, const uint8_t *valid,
                       int64_t n, int op, const void *scalar, uint8_t *mask) {
    for (int64_t i = 0; i < n; i++) {
        int c;
        if (vclass == ORC_I64) {
            int64_t a = ((const int64_t *)values)[i], b = *(const int64_t *)scalar;
            c = (a > b) - (a < b);
        } else if (vclass == ORC_U64) {
            uint64_t a = ((const uint64_t *)values)[i], b = *(const uint64_t *)scalar;
            c = (a > b) - (a < b);
        } else {
            c = tot_cmp_f64(((const double *)values)[i], *(const double *)scalar);
        }
        int r;
        switch (op) {
        case ORC_EQ: r = c == 0; break;
        case ORC_NE: r = c != 0; break;
        case ORC_LT: r = c < 0; break;
        case ORC_LE: r = c <= 0; break;
        case ORC_GT: r = c > 0; break;
        default: r = c >= 0; break;
        }
        if (valid && !valid[i]) r = 0;
        mask[i] = (uint8_t)r;
    }
    return 0;
}

/* =========================================================================
 * a3/a4/a6/a10 — hash grouping + per-group reductions
 *
 * Group identity: rows are equal iff every key word is equal and the null
 * flags are equal; a null key is its own group
 *   polars-core/src/frame/group_by/hashing.rs:75-111, into_groups.rs:30-58
 *   polars-expr/src/groups/single_key.rs:41-47, 93-147
 * Group order (maintain_order): first occurrence
 *   polars-core/src/frame/group_by/hashing.rs:26-63 (finish_group_order)
 *
 * Reductions (polars-expr/src/reduce/*):
 *   sum   sum.rs:25-47,93-110   null adds 0, all-null group = 0; integer
 *         sums wrap (the wrapper truncates i32/u32 to their width, which is
 *         exact because wrapping add is a ring homomorphism); f64 sum uses
 *         Kahan like polars-core/.../aggregations/mod.rs:594-606; f32 sums
 *         run in f32 arrival order (NumSumReducer<Float32Type>::Value = f32)
 *   mean  mean.rs:29-52,82-131  state (f64 sum, count non-null); count==0 => null
 *   min/max min_max.rs:110-162 + polars-utils/src/min_max.rs:12-123: floats
 *         ignore NaN unless every non-null value is NaN; all-null => null
 *   count count.rs:6-124 (non-null)   len len.rs (all rows)
 *   first/last first_last.rs:50-168: value of the first/last ROW of the
 *         group, nulls included
 *
 * Structure mirrors the streaming node (polars-stream/src/nodes/group_by.rs
 * :116-441): thread-local tables over row shards, then a merge; with
 * n_threads == 1 it is the plain sequential oracle.
 * ========================================================================= */
typedef struct {
    int32_t kind;        /* ORC_SUM.. */
    int32_t vclass;      /* ORC_I64.. */
    const void *values;  /* widened 8-byte values (or float for F32) */
    const uint8_t *valid;/* byte per row or NULL */
    int32_t ddof;        /* VAR / STD */
    int32_t is_bool;     /* the source column is Boolean (values widened to 0 / 1) */
} OrcAgg;

typedef struct { uint64_t w0, w1; int64_t cnt; int64_t aux; } OrcState;
/* layout by kind:
 *  SUM f64/f32: w0=sum bits w1=kahan comp   | SUM int: w0=sum
 *  MEAN: w0=sum w1=comp cnt=non-null
 *  MIN/MAX: w0=current cnt=non-null aux=non-NaN count
 *  COUNT/LEN: cnt
 *  FIRST/LAST: w0=value bits w1=row index cnt=has aux=valid
 *  VAR/STD: w0=mean w1=dp cnt=weight (polars-compute/src/moment.rs:40-44 VarState);
 *           Boolean input: w0=number of true values cnt=non-null (var_std.rs:144-160)
 *  FIRST_NN/LAST_NN: as FIRST/LAST over the non-null rows only (first_last_nonnull.rs)
 *  NULL_COUNT: cnt=nulls (count.rs NullCountReduce)
 *  BIT_AND/OR/XOR: w0=accumulator cnt=non-null (bitwise.rs; masked: no non-null value => null)
 *  ANY: w0=1 once a true was seen; ALL: w0=number of non-null false values (any_all.rs, ignore_nulls) */

static inline double u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

static inline void kahan_add(OrcState *s, double x) {
    /* polars-utils/src/kahan_sum.rs:35-46 */
    double sum = u2d(s->w0), err = u2d(s->w1);
    if (isfinite(x)) {
        double y = x - err;
        double t = sum + y;
        err = (t - sum) - y;
        sum = t;
    } else {
        sum += x;
    }
    s->w0 = d2u(sum); s->w1 = d2u(err);
}

static inline double get_f64(const OrcAgg *a, int64_t i) {
    if (a->vclass == ORC_F64) return ((const double *)a->values)[i];
    if (a->vclass == ORC_F32) return (double)((const float *)a->values)[i];
    if (a->vclass == ORC_U64) return (double)((const uint64_t *)a->values)[i];
    return (double)((const int64_t *)a->values)[i];
}
static inline uint64_t get_bits(const OrcAgg *a, int64_t i) {
    if (a->vclass == ORC_F32) return d2u((double)((const float *)a->values)[i]);
    return ((const uint64_t *)a->values)[i];
}

static inline void state_init(OrcState *s) { memset(s, 0, sizeof *s); }

static inline void state_update(OrcState *s, const OrcAgg *a, int64_t row) {
    int ok = !a->valid || a->valid[row];
    switch (a->kind) {
    case ORC_SUM:
        if (!ok) break;
        if (a->vclass == ORC_F64) kahan_add(s, ((const double *)a->values)[row]);
        else if (a->vclass == ORC_F32) {
            float f; uint32_t b = (uint32_t)s->w0; memcpy(&f, &b, 4);
            f += ((const float *)a->values)[row];
            memcpy(&b, &f, 4); s->w0 = b;
        } else s->w0 += ((const uint64_t *)a->values)[row];
        break;
    case ORC_MEAN:
        if (!ok) break;
        kahan_add(s, get_f64(a, row)); s->cnt++;
        break;
    case ORC_MIN: case ORC_MAX: {
        if (!ok) break;
        int is_min = a->kind == ORC_MIN;
        if (a->vclass == ORC_F64 || a->vclass == ORC_F32) {
            double v = get_f64(a, row);
            s->cnt++;
            if (isnan(v)) break;
            if (s->aux == 0) s->w0 = d2u(v);
            else {
                double c = u2d(s->w0);
                /* total order on zeros (-0 < +0) so the result does not depend
                 * on visit order; the reference's f64::min leaves it open */
                int lt = (v < c) || (v == c && signbit(v) && !signbit(c));
                int gt = (v > c) || (v == c && !signbit(v) && signbit(c));
                if (is_min ? lt : gt) s->w0 = d2u(v);
            }
            s->aux++;
        } else if (a->vclass == ORC_U64) {
            uint64_t v = ((const uint64_t *)a->values)[row];
            if (s->cnt == 0 || (is_min ? v < s->w0 : v > s->w0)) s->w0 = v;
            s->cnt++;
        } else {
            int64_t v = ((const int64_t *)a->values)[row];
            if (s->cnt == 0 || (is_min ? v < (int64_t)s->w0 : v > (int64_t)s->w0)) s->w0 = (uint64_t)v;
            s->cnt++;
        }
        break; }
    case ORC_COUNT: s->cnt += ok; break;
    case ORC_LEN: s->cnt++; break;
    case ORC_FIRST:
        if (!s->cnt) { s->cnt = 1; s->w1 = (uint64_t)row; s->aux = ok; s->w0 = ok ? get_bits(a, row) : 0; }
        break;
    case ORC_LAST:
        s->cnt = 1; s->w1 = (uint64_t)row; s->aux = ok; s->w0 = ok ? get_bits(a, row) : 0;
        break;
    case ORC_VAR: case ORC_STD:
        if (!ok) break;
        if (a->is_bool) { s->w0 += ((const uint64_t *)a->values)[row] != 0; s->cnt++; break; }
        {   /* VarState::insert_one, moment.rs:87-97 */
            const double x = get_f64(a, row);
            const double new_weight = (double)s->cnt + 1.0;
            const double mean = u2d(s->w0), delta_mean = x - mean;
            const double new_mean = mean + delta_mean / new_weight;
            s->w1 = d2u(u2d(s->w1) + (x - new_mean) * delta_mean);
            s->w0 = d2u(new_mean);
            s->cnt++;
        }
        break;
    case ORC_FIRST_NN:
        if (ok && !s->cnt) { s->cnt = 1; s->w1 = (uint64_t)row; s->aux = 1; s->w0 = get_bits(a, row); }
        break;
    case ORC_LAST_NN:
        if (ok) { s->cnt = 1; s->w1 = (uint64_t)row; s->aux = 1; s->w0 = get_bits(a, row); }
        break;
    case ORC_NULL_COUNT: s->cnt += !ok; break;
    case ORC_BIT_AND: case ORC_BIT_OR: case ORC_BIT_XOR:
        if (!ok) break;
        {
            const uint64_t v = ((const uint64_t *)a->values)[row];
            if (!s->cnt) s->w0 = a->kind == ORC_BIT_AND ? ~0ull : 0ull;   /* the identity element (bitwise.rs:68-120) */
            s->w0 = a->kind == ORC_BIT_AND ? (s->w0 & v) : a->kind == ORC_BIT_OR ? (s->w0 | v) : (s->w0 ^ v);
            s->cnt++;
        }
        break;
    case ORC_ANY: if (ok && ((const uint64_t *)a->values)[row]) s->w0 = 1; break;
    case ORC_ALL: if (ok && !((const uint64_t *)a->values)[row]) s->w0++; break;
    }
}

/* merge b (later rows) into a (earlier rows): GroupedReduction::combine_subset
 * polars-expr/src/reduce/mod.rs:94-105 and each reducer's `combine`. */
static inline void state_combine(OrcState *a, const OrcState *b, const OrcAgg *g) {
    switch (g->kind) {
    case ORC_SUM:
        if (g->vclass == ORC_F64) { kahan_add(a, u2d(b->w0)); kahan_add(a, -u2d(b->w1)); }
        else if (g->vclass == ORC_F32) {
            float x, y; uint32_t p = (uint32_t)a->w0, q = (uint32_t)b->w0;
            memcpy(&x, &p, 4); memcpy(&y, &q, 4); x += y; memcpy(&p, &x, 4); a->w0 = p;
        } else a->w0 += b->w0;
        break;
    case ORC_MEAN:
        kahan_add(a, u2d(b->w0)); kahan_add(a, -u2d(b->w1)); a->cnt += b->cnt; break;
    case ORC_MIN: case ORC_MAX: {
        int is_min = g->kind == ORC_MIN;
        if (g->vclass == ORC_F64 || g->vclass == ORC_F32) {
            if (b->aux) {
                if (!a->aux) a->w0 = b->w0;
                else {
                    double v = u2d(b->w0), c = u2d(a->w0);
                    int lt = (v < c) || (v == c && signbit(v) && !signbit(c));
                    int gt = (v > c) || (v == c && !signbit(v) && signbit(c));
                    if (is_min ? lt : gt) a->w0 = b->w0;
                }
            }
            a->aux += b->aux; a->cnt += b->cnt;
        } else if (g->vclass == ORC_U64) {
            if (b->cnt && (!a->cnt || (is_min ? b->w0 < a->w0 : b->w0 > a->w0))) a->w0 = b->w0;
            a->cnt += b->cnt;
        } else {
            if (b->cnt && (!a->cnt || (is_min ? (int64_t)b->w0 < (int64_t)a->w0 : (int64_t)b->w0 > (int64_t)a->w0))) a->w0 = b->w0;
            a->cnt += b->cnt;
        }
        break; }
    case ORC_COUNT: case ORC_LEN: a->cnt += b->cnt; break;
    case ORC_FIRST:
        if (b->cnt && (!a->cnt || b->w1 < a->w1)) *a = *b;
        break;
    case ORC_LAST:
        if (b->cnt && (!a->cnt || b->w1 >= a->w1)) *a = *b;
        break;
    case ORC_VAR: case ORC_STD:
        if (g->is_bool) { a->w0 += b->w0; a->cnt += b->cnt; break; }
        if (b->cnt == 0) break;
        {   /* VarState::combine, moment.rs:111-124 */
            const double bw = (double)b->cnt, new_weight = (double)a->cnt + bw;
            const double frac = bw / new_weight;
            const double delta_mean = u2d(b->w0) - u2d(a->w0);
            const double new_mean = u2d(a->w0) + delta_mean * frac;
            a->w1 = d2u(u2d(a->w1) + u2d(b->w1) + bw * (u2d(b->w0) - new_mean) * delta_mean);
            a->w0 = d2u(new_mean);
            a->cnt += b->cnt;
        }
        break;
    case ORC_FIRST_NN:
        if (b->cnt && (!a->cnt || b->w1 < a->w1)) *a = *b;
        break;
    case ORC_LAST_NN:
        if (b->cnt && (!a->cnt || b->w1 >= a->w1)) *a = *b;
        break;
    case ORC_NULL_COUNT: a->cnt += b->cnt; break;
    case ORC_BIT_AND: case ORC_BIT_OR: case ORC_BIT_XOR:
        if (!b->cnt) break;
        if (!a->cnt) a->w0 = b->w0;
        else a->w0 = g->kind == ORC_BIT_AND ? (a->w0 & b->w0) : g->kind == ORC_BIT_OR ? (a->w0 | b->w0) : (a->w0 ^ b->w0);
        a->cnt += b->cnt;
        break;
    case ORC_ANY: a->w0 |= b->w0; break;
    case ORC_ALL: a->w0 += b->w0; break;
    }
}

typedef struct {
    int64_t cap, n;          /* slots (pow2), groups */
    int32_t *slot_gid;       /* -1 empty */
    uint64_t *keys;          /* n_words per group */
    uint32_t *nullmask;      /* per group */
    int64_t *first_row;      /* per group */
    OrcState *st;            /* n_aggs per group */
    int64_t gcap;
} OrcTable;

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

static void table_init(OrcTable *t, int n_words, int n_aggs, int64_t cap) {
    t->cap = cap; t->n = 0; t->gcap = cap / 2 + 1;
    t->slot_gid = (int32_t *)malloc(sizeof(int32_t) * cap);
    memset(t->slot_gid, 0xff, sizeof(int32_t) * cap);
    t->keys = (uint64_t *)malloc(sizeof(uint64_t) * t->gcap * (n_words ? n_words : 1));
    t->nullmask = (uint32_t *)malloc(sizeof(uint32_t) * t->gcap);
    t->first_row = (int64_t *)malloc(sizeof(int64_t) * t->gcap);
    t->st = (OrcState *)malloc(sizeof(OrcState) * t->gcap * (n_aggs ? n_aggs : 1));
}
static void table_free(OrcTable *t) {
    free(t->slot_gid); free(t->keys); free(t->nullmask); free(t->first_row); free(t->st);
}
static void table_grow(OrcTable *t, int n_words, int n_aggs) {
    int64_t ncap = t->cap * 2;
    free(t->slot_gid);
    t->slot_gid = (int32_t *)malloc(sizeof(int32_t) * ncap);
    memset(t->slot_gid, 0xff, sizeof(int32_t) * ncap);
    t->gcap = ncap / 2 + 1;
    t->keys = (uint64_t *)realloc(t->keys, sizeof(uint64_t) * t->gcap * (n_words ? n_words : 1));
    t->nullmask = (uint32_t *)realloc(t->nullmask, sizeof(uint32_t) * t->gcap);
    t->first_row = (int64_t *)realloc(t->first_row, sizeof(int64_t) * t->gcap);
    t->st = (OrcState *)realloc(t->st, sizeof(OrcState) * t->gcap * (n_aggs ? n_aggs : 1));
    t->cap = ncap;
    for (int64_t g = 0; g < t->n; g++) {
        uint64_t h = t->nullmask[g] * 0x9E3779B97F4A7C15ULL;
        for (int w = 0; w < n_words; w++) h = mix64(h ^ t->keys[g * n_words + w]);
        int64_t s = (int64_t)(h & (uint64_t)(ncap - 1));
        while (t->slot_gid[s] >= 0) s = (s + 1) & (ncap - 1);
        t->slot_gid[s] = (int32_t)g;
    }
}

static inline int64_t table_find_or_insert(OrcTable *t, const uint64_t *kw, uint32_t nm,
                                           int n_words, int n_aggs, int64_t row) {
    uint64_t h = nm * 0x9E3779B97F4A7C15ULL;
    for (int w = 0; w < n_words; w++) h = mix64(h ^ kw[w]);
    int64_t s = (int64_t)(h & (uint64_t)(t->cap - 1));
    for (;;) {
        int32_t g = t->slot_gid[s];
        if (g < 0) break;
        if (t->nullmask[g] == nm && memcmp(&t->keys[(int64_t)g * n_words], kw, 8 * (size_t)n_words) == 0) return g;
        s = (s + 1) & (t->cap - 1);
    }
    if ((t->n + 1) * 2 > t->cap) {
        table_grow(t, n_words, n_aggs);
        return table_find_or_insert(t, kw, nm, n_words, n_aggs, row);
    }
    int64_t g = t->n++;
    t->slot_gid[s] = (int32_t)g;
    memcpy(&t->keys[g * n_words], kw, 8 * (size_t)n_words);
    t->nullmask[g] = nm;
    t->first_row[g] = row;
    for (int a = 0; a < n_aggs; a++) state_init(&t->st[g * n_aggs + a]);
    return g;
}

/* Result handle so the Python side can size its output arrays. */
typedef struct {
    int64_t n_groups;
    int32_t n_words, n_aggs;
    uint64_t *keys; uint32_t *nullmask; int64_t *first_row; OrcState *st;
} OrcResult;

static int cmp_first_row(const void *a, const void *b, void *ctx) {
    const int64_t *fr = (const int64_t *)ctx;
    int64_t x = fr[*(const int64_t *)a], y = fr[*(const int64_t *)b];
    return (x > y) - (x < y);
}

typedef struct {
    const uint64_t *const *keys; const uint8_t *const *key_valid; int n_words;
    const uint8_t *sel; const OrcAgg *aggs; int n_aggs; int64_t lo, hi; OrcTable *tb;
} ShardJob;

/* Morsel-wise update of one aggregate: the shape of GroupedReduction::update_groups_subset
 * (polars-expr/src/reduce/mod.rs:292-312) — group indices first, then ONE typed loop per
 * aggregate with the dispatch outside of it.  Same arithmetic as state_update (which stays the
 * reference for every other kind / class and for the slice path). */
static void update_morsel(OrcState *st, int n_aggs, int a, const OrcAgg *g,
                          const int32_t *gid, const int64_t *rows, int m) {
    const uint8_t *valid = g->valid;
    if (g->kind == ORC_LEN) {
        for (int k = 0; k < m; k++) st[(int64_t)gid[k] * n_aggs + a].cnt++;
    } else if (g->kind == ORC_COUNT) {
        for (int k = 0; k < m; k++) st[(int64_t)gid[k] * n_aggs + a].cnt += !valid || valid[rows[k]];
    } else if (g->vclass == ORC_F64 && (g->kind == ORC_SUM || g->kind == ORC_MEAN)) {
        const double *v = (const double *)g->values;
        const int mean = g->kind == ORC_MEAN;
        for (int k = 0; k < m; k++) {
            if (valid && !valid[rows[k]]) continue;
            OrcState *s = &st[(int64_t)gid[k] * n_aggs + a];
            kahan_add(s, v[rows[k]]);
            s->cnt += mean;
        }
    } else if (g->vclass == ORC_F64 && (g->kind == ORC_MIN || g->kind == ORC_MAX)) {
        const double *v = (const double *)g->values;
        const int is_min = g->kind == ORC_MIN;
        for (int k = 0; k < m; k++) {
            if (valid && !valid[rows[k]]) continue;
            OrcState *s = &st[(int64_t)gid[k] * n_aggs + a];
            const double x = v[rows[k]];
            s->cnt++;
            if (isnan(x)) continue;
            if (s->aux == 0) s->w0 = d2u(x);
            else {
                const double c = u2d(s->w0);
                const int lt = (x < c) || (x == c && signbit(x) && !signbit(c));
                const int gt = (x > c) || (x == c && !signbit(x) && signbit(c));
                if (is_min ? lt : gt) s->w0 = d2u(x);
            }
            s->aux++;
        }
    } else {
        for (int k = 0; k < m; k++) state_update(&st[(int64_t)gid[k] * n_aggs + a], g, rows[k]);
    }
}

/* one pipeline of the sink phase (nodes/group_by.rs:116-214): a thread-local table over a
 * contiguous row shard, fed morsel by morsel: probe -> group index per row, then one pass per
 * aggregate (update_morsel) */
#define ORC_MORSEL 2048
static void *shard_main(void *arg) {
    ShardJob *j = (ShardJob *)arg;
    OrcTable *tb = j->tb;
    int n_words = j->n_words, n_aggs = j->n_aggs;
    table_init(tb, n_words, n_aggs, 1024);
    uint64_t kw[16];
    int32_t gid[ORC_MORSEL];
    int64_t rows[ORC_MORSEL];
    for (int64_t lo = j->lo; lo < j->hi; lo += ORC_MORSEL) {
        const int64_t hi = lo + ORC_MORSEL < j->hi ? lo + ORC_MORSEL : j->hi;
        int m = 0;
        for (int64_t i = lo; i < hi; i++) {
            if (j->sel && !j->sel[i]) continue;
            uint32_t nm = 0;
            for (int w = 0; w < n_words; w++) {
                int ok = !j->key_valid || !j->key_valid[w] || j->key_valid[w][i];
                kw[w] = ok ? j->keys[w][i] : 0;
                nm |= (uint32_t)(!ok) << w;
            }
            gid[m] = (int32_t)table_find_or_insert(tb, kw, nm, n_words, n_aggs, i);
            rows[m++] = i;
        }
        /* the table may have grown (realloc) while probing: take the state pointer afterwards */
        for (int a = 0; a < n_aggs; a++) update_morsel(tb->st, n_aggs, a, &j->aggs[a], gid, rows, m);
    }
    return NULL;
}

/* keys[w][row]: widened key words; key_valid[w][row] byte or NULL pointer;
 * sel: byte mask of surviving rows or NULL.  Output groups are ordered by
 * first occurrence. */
OrcResult *orc_groupby(const uint64_t *const *keys, const uint8_t *const *key_valid,
                       int n_words, int64_t n, const uint8_t *sel,
                       const OrcAgg *aggs, int n_aggs, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    OrcTable *tabs = (OrcTable *)calloc((size_t)n_threads, sizeof(OrcTable));
    ShardJob *jobs = (ShardJob *)calloc((size_t)n_threads, sizeof(ShardJob));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        ShardJob j = { keys, key_valid, n_words, sel, aggs, n_aggs,
                       n * t / n_threads, n * (t + 1) / n_threads, &tabs[t] };
        jobs[t] = j;
        if (t > 0) pthread_create(&th[t], NULL, shard_main, &jobs[t]);
    }
    shard_main(&jobs[0]);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    free(jobs); free(th);
    /* merge thread tables in shard order (earlier rows first) */
    OrcTable *m = &tabs[0];
    for (int t = 1; t < n_threads; t++) {
        OrcTable *tb = &tabs[t];
        for (int64_t g = 0; g < tb->n; g++) {
            int64_t before = m->n;
            int64_t mg = table_find_or_insert(m, &tb->keys[g * n_words], tb->nullmask[g], n_words, n_aggs, tb->first_row[g]);
            if (mg == before) memcpy(&m->st[mg * n_aggs], &tb->st[g * n_aggs], sizeof(OrcState) * (size_t)n_aggs);
            else for (int a = 0; a < n_aggs; a++) state_combine(&m->st[mg * n_aggs + a], &tb->st[g * n_aggs + a], &aggs[a]);
        }
        table_free(tb);
    }
    OrcResult *r = (OrcResult *)calloc(1, sizeof *r);
    r->n_groups = m->n; r->n_words = n_words; r->n_aggs = n_aggs;
    /* order by first occurrence */
    int64_t *perm = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m->n + 1));
    for (int64_t g = 0; g < m->n; g++) perm[g] = g;
    qsort_r(perm, (size_t)m->n, sizeof(int64_t), cmp_first_row, m->first_row);
    r->keys = (uint64_t *)malloc(8 * (size_t)(m->n * (n_words ? n_words : 1) + 1));
    r->nullmask = (uint32_t *)malloc(4 * (size_t)(m->n + 1));
    r->first_row = (int64_t *)malloc(8 * (size_t)(m->n + 1));
    r->st = (OrcState *)malloc(sizeof(OrcState) * (size_t)(m->n * (n_aggs ? n_aggs : 1) + 1));
    for (int64_t i = 0; i < m->n; i++) {
        int64_t g = perm[i];
        memcpy(&r->keys[i * n_words], &m->keys[g * n_words], 8 * (size_t)n_words);
        r->nullmask[i] = m->nullmask[g];
        r->first_row[i] = m->first_row[g];
        memcpy(&r->st[i * n_aggs], &m->st[g * n_aggs], sizeof(OrcState) * (size_t)n_aggs);
    }
    free(perm);
    table_free(m);
    free(tabs);
    return r;
}

int64_t orc_result_ngroups(const OrcResult *r) { return r->n_groups; }

/* copy out group keys (word w) + null flag */
void orc_result_keys(const OrcResult *r, int w, uint64_t *out, uint8_t *out_valid) {
    for (int64_t g = 0; g < r->n_groups; g++) {
        out[g] = r->keys[g * r->n_words + w];
        out_valid[g] = !((r->nullmask[g] >> w) & 1);
    }
}
void orc_result_first_rows(const OrcResult *r, int64_t *out) {
    memcpy(out, r->first_row, 8 * (size_t)r->n_groups);
}

/* Finalise aggregation a into 8-byte outputs (+valid byte).  Output classes:
 * SUM int -> i64 bits, SUM f64 -> f64, SUM f32 -> f32 widened to f64;
 * MEAN -> f64 (mean.rs:29-52: count==0 => null);  MIN/MAX -> input class;
 * COUNT/LEN -> u64 (the wrapper narrows to IdxSize u32,
 * polars-utils/src/index.rs:9-11);  FIRST/LAST -> input bits. */
void orc_result_agg(const OrcResult *r, const OrcAgg *aggs, int a, uint64_t *out, uint8_t *out_valid) {
    const OrcAgg *g = &aggs[a];
    for (int64_t i = 0; i < r->n_groups; i++) {
        const OrcState *s = &r->st[i * r->n_aggs + a];
        uint64_t v = 0; uint8_t ok = 1;
        switch (g->kind) {
        case ORC_SUM:
            if (g->vclass == ORC_F64) v = s->w0; /* kahan: sum already compensated per step */
            else if (g->vclass == ORC_F32) { float f; uint32_t b = (uint32_t)s->w0; memcpy(&f, &b, 4); v = d2u((double)f); }
            else v = s->w0;
            break;
        case ORC_MEAN:
            if (s->cnt == 0) ok = 0; else v = d2u(u2d(s->w0) / (double)s->cnt);
            break;
        case ORC_MIN: case ORC_MAX:
            if (s->cnt == 0) ok = 0;
            else if ((g->vclass == ORC_F64 || g->vclass == ORC_F32) && s->aux == 0) v = d2u(NAN);
            else v = s->w0;
            break;
        case ORC_COUNT: case ORC_LEN: v = (uint64_t)s->cnt; break;
        case ORC_FIRST: case ORC_LAST: case ORC_FIRST_NN: case ORC_LAST_NN:
            if (!s->cnt || !s->aux) ok = 0; else v = s->w0;
            break;
        case ORC_VAR: case ORC_STD: {
            /* VarState::finalize, moment.rs:126-139; Boolean: var_std.rs:144-160 */
            if (s->cnt <= (int64_t)g->ddof) { ok = 0; break; }
            double var;
            if (g->is_bool) {
                const double sum = (double)s->w0;
                var = sum * (1.0 - sum / (double)s->cnt) / (double)(s->cnt - g->ddof);
            } else {
                var = u2d(s->w1) / ((double)s->cnt - (double)g->ddof);
                if (var < 0.0) var = 0.0;
            }
            v = d2u(g->kind == ORC_STD ? sqrt(var) : var);
            break; }
        case ORC_NULL_COUNT: v = (uint64_t)s->cnt; break;
        case ORC_BIT_AND: case ORC_BIT_OR: case ORC_BIT_XOR:
            if (!s->cnt) ok = 0; else v = s->w0;
            break;
        case ORC_ANY: v = s->w0 != 0; break;
        case ORC_ALL: v = s->w0 == 0; break;
        }
        out[i] = v; out_valid[i] = ok;
    }
}

void orc_result_free(OrcResult *r) {
    if (!r) return;
    free(r->keys); free(r->nullmask); free(r->first_row); free(r->st); free(r);
}

/* =========================================================================
 * a4 — GroupsIdx construction: (first, all) per group in first-occurrence
 * order, row ids ascending inside a group.
 *   polars-core/src/frame/group_by/hashing.rs:75-111, position.rs:16-20
 * gid_out[row] = group number (or -1 when !sel[row]).
 * ========================================================================= */
int64_t orc_group_ids(const uint64_t *const *keys, const uint8_t *const *key_valid,
                      int n_words, int64_t n, const uint8_t *sel, int32_t *gid_out,
                      int64_t *first_rows) {
    OrcTable tb; table_init(&tb, n_words, 0, 1024);
    uint64_t kw[16];
    for (int64_t i = 0; i < n; i++) {
        if (sel && !sel[i]) { gid_out[i] = -1; continue; }
        uint32_t nm = 0;
        for (int w = 0; w < n_words; w++) {
            int ok = !key_valid || !key_valid[w] || key_valid[w][i];
            kw[w] = ok ? keys[w][i] : 0;
            nm |= (uint32_t)(!ok) << w;
        }
        gid_out[i] = (int32_t)table_find_or_insert(&tb, kw, nm, n_words, 0, i);
    }
    int64_t ng = tb.n;
    if (first_rows) memcpy(first_rows, tb.first_row, 8 * (size_t)ng);
    table_free(&tb);
    return ng;
}

/* =========================================================================
 * a5 — sorted-key fast path: run boundaries -> [start,len]
 *   polars-core/src/frame/group_by/into_groups.rs:65-129
 *   polars-arrow/src/legacy/kernels/sort_partition.rs:168 (partition_to_groups)
 * ========================================================================= */
int64_t orc_partition_to_groups(const uint64_t *keys, const uint8_t *valid, int64_t n,
                                int64_t *starts, int64_t *lens) {
    int64_t ng = 0;
    for (int64_t i = 0; i < n; i++) {
        int ok = !valid || valid[i];
        int same = 0;
        if (i > 0) {
            int pok = !valid || valid[i - 1];
            same = (ok == pok) && (!ok || keys[i] == keys[i - 1]);
        }
        if (!same) { starts[ng] = i; lens[ng] = 1; ng++; }
        else lens[ng - 1]++;
    }
    return ng;
}

/* =========================================================================
 * a12 — dynamic windows, fixed-duration `every`/`period`/`offset`
 * (integers in the index column's own unit; no time zone, no calendar months)
 *   polars-time/src/windows/bounds.rs:33-76          membership predicates
 *   polars-time/src/windows/duration.rs:681-685      truncate (floor-mod)
 *   polars-time/src/windows/window.rs:25-53          ensure_t_in_or_in_front_of_window
 *   polars-time/src/windows/window.rs:115-170        get_earliest_bounds_*
 *   polars-time/src/windows/window.rs:342-438        BoundsIter::{next,nth,get_stride}
 *   polars-time/src/windows/group_by.rs:79-151       update_groups_and_bounds
 *   polars-time/src/windows/group_by.rs:165-246      group_by_windows
 * This is a deliberately literal restatement (two-pointer sweep, strides and
 * the "last value" special case included) so that it can be pinned by
 * polars-time/src/windows/test.rs; the CUDA kernels use a closed form.
 * ========================================================================= */
typedef struct { int64_t start, stop; } Bounds;

static inline int b_is_member(Bounds b, int64_t t, int c) {
    switch (c) {
    case ORC_CLOSED_RIGHT: return t > b.start && t <= b.stop;
    case ORC_CLOSED_LEFT: return t >= b.start && t < b.stop;
    case ORC_CLOSED_NONE: return t > b.start && t < b.stop;
    default: return t >= b.start && t <= b.stop;
    }
}
static inline int b_is_member_entry(Bounds b, int64_t t, int c) {
    return (c == ORC_CLOSED_RIGHT || c == ORC_CLOSED_NONE) ? t > b.start : t >= b.start;
}
static inline int b_is_member_exit(Bounds b, int64_t t, int c) {
    return (c == ORC_CLOSED_RIGHT || c == ORC_CLOSED_BOTH) ? t <= b.stop : t < b.stop;
}
static inline int b_is_future(Bounds b, int64_t t, int c) {
    return (c == ORC_CLOSED_LEFT || c == ORC_CLOSED_NONE) ? b.stop <= t : b.stop < t;
}
static inline int b_is_past(Bounds b, int64_t t, int c) {
    return (c == ORC_CLOSED_LEFT || c == ORC_CLOSED_BOTH) ? b.start > t : b.start >= t;
}

static inline int64_t truncate_fixed(int64_t t, int64_t every) {
    int64_t r = t % every;
    if (r < 0) r += every;
    return t - r;
}

typedef struct { int64_t every, period; Bounds boundary, bi; } BoundsIter;

static int bi_next(BoundsIter *it, Bounds *out) {
    if (it->bi.start < it->boundary.stop) {
        *out = it->bi;
        it->bi.start += it->every;
        it->bi.stop = it->bi.start + it->period;
        return 1;
    }
    return 0;
}
static int bi_nth(BoundsIter *it, int64_t n, Bounds *out) {
    if (it->bi.start < it->boundary.stop) {
        it->bi.start += it->every * n;
        it->bi.stop = it->bi.start + it->period;
        return bi_next(it, out);
    }
    return 0;
}
static int64_t bi_get_stride(const BoundsIter *it, int64_t target) {
    int64_t stride = 0;
    if (it->bi.start < it->boundary.stop && target > it->bi.start) {
        int64_t gap = target - it->bi.start;
        if (gap > it->every + it->period) stride = (gap - it->period) / it->every;
    }
    return stride;
}

/* returns number of (non-empty) windows; starts/lens/lower/upper sized by the
 * caller with orc_group_by_windows(…, NULL…) first (two-call protocol). */
int64_t orc_group_by_windows(const int64_t *time, int64_t n, int64_t every, int64_t period,
                             int64_t offset, int closed, int64_t *starts, int64_t *lens,
                             int64_t *lower, int64_t *upper) {
    if (n == 0) return 0;
    Bounds boundary;
    boundary.start = time[0];
    boundary.stop = (n > 1) ? time[n - 1] + 1 : time[0] + 1;
    /* get_earliest_bounds + ensure_t_in_or_in_front_of_window */
    int64_t t0 = boundary.start;
    int64_t start = truncate_fixed(t0, every) + offset;
    int64_t stop = start + period;
    for (;;) {
        Bounds b = { start, stop };
        if (!b_is_past(b, t0, closed)) break;
        int64_t gap = start - t0;
        if (closed == ORC_CLOSED_RIGHT || closed == ORC_CLOSED_NONE) gap += 1;
        int64_t stride = (gap + every - 1) / every;
        if (stride < 1) stride = 1;
        start -= every * stride;
        stop = start + period;
    }
    BoundsIter it = { every, period, boundary, { start, stop } };

    int64_t ng = 0, s = 0, stride = 0;
    Bounds bi;
    while (bi_nth(&it, stride, &bi)) {
        int has_member = 0, skipped = 0;
        int64_t lim = n > 0 ? n - 1 : 0;
        for (int64_t k = s; k < lim; k++) {
            int64_t t = time[k];
            if (b_is_future(bi, t, closed)) { stride = bi_get_stride(&it, t); skipped = 1; break; }
            if (b_is_member_entry(bi, t, closed)) { has_member = 1; break; }
            s++;
        }
        if (skipped) continue;
        stride = has_member ? 0 : bi_get_stride(&it, time[s]);
        int64_t e = s;
        if (e == n - 1) {
            int64_t t = time[e];
            if (b_is_member(bi, t, closed)) {
                if (starts) { starts[ng] = e; lens[ng] = 1; lower[ng] = bi.start; upper[ng] = bi.stop; }
                ng++;
            }
            continue;
        }
        for (int64_t k = e; k < n; k++) {
            if (!b_is_member_exit(bi, time[k], closed)) break;
            e++;
        }
        if (starts) { starts[ng] = s; lens[ng] = e - s; lower[ng] = bi.start; upper[ng] = bi.stop; }
        ng++;
    }
    return ng;
}

/* =========================================================================
 * a7 — contiguous-slice reductions (GroupsType::Slice), used after
 * group_by_windows:  polars-core/.../aggregations/mod.rs:184-191 ->
 * ChunkAgg (chunked_array/ops/aggregate/mod.rs:44-130).  Same null/NaN rules
 * as above; implemented by re-using the state machine over [start,len].
 * ========================================================================= */
void orc_agg_slices(const OrcAgg *agg, const int64_t *starts, const int64_t *lens,
                    int64_t n_slices, uint64_t *out, uint8_t *out_valid) {
    OrcResult r; memset(&r, 0, sizeof r);
    r.n_groups = n_slices; r.n_aggs = 1; r.n_words = 0;
    r.st = (OrcState *)malloc(sizeof(OrcState) * (size_t)(n_slices + 1));
    for (int64_t g = 0; g < n_slices; g++) {
        state_init(&r.st[g]);
        for (int64_t i = starts[g]; i < starts[g] + lens[g]; i++) state_update(&r.st[g], agg, i);
    }
    orc_result_agg(&r, agg, 0, out, out_valid);
    free(r.st);
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
