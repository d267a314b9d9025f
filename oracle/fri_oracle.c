/*
 * fri_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; PARITY UNPINNED — see fri_oracle.h).
 *
 * Restates, in plain C, the semantics of the reference's hot path.  Citations are
 * file:line relative to /root/reference/.  The code is written from the semantics of the
 * cited lines (Option-valued lifting, truncating division, bounds-checked gather/scatter);
 * no reference source is copied.
 *
 * Reference quirk worth knowing (and the one deliberate deviation of this oracle):
 *   wavelet_transform.rs:180-184 always allocates THREE coefficient vectors filled with
 *   None, :191 fills only `num_channels` of them, and the retain at :415-416 keeps a tile
 *   only if `channel[0].is_some()` for ALL THREE vectors.  For a 1-channel (Luma) image
 *   that predicate is false for every tile, the lattice becomes empty and
 *   sort_lattice indexes `keys[0]` of an empty Vec (:663-664) -> panic.  The reference is
 *   therefore only runnable on 3-channel images.  This oracle applies the retain predicate
 *   to the ACTIVE channels, which is identical for 3 channels and gives 1-channel images
 *   the obviously intended behaviour.  Validity is geometric (identical for every channel).
 */
#include "fri_oracle.h"

#include <stdlib.h>
#include <string.h>

/* crates/libfri/src/fractal.rs:51-86 */
const int32_t FRI_ORACLE_LITERALS[30][2] = {
    {0, 1},         {-1, 1},       {2, 0},        {-3, -1},       {5, -1},
    {1, 3},         {-11, -1},     {9, -5},       {13, 7},        {-31, 3},
    {5, -17},       {57, 11},      {-67, 23},     {-47, -45},     {181, -1},
    {-87, 91},      {-275, -89},   {449, -93},    {101, 271},     {-999, -85},
    {797, -457},    {1201, 627},   {-2795, 287},  {393, -1541},   {5197, 967},
    {-5983, 2115},  {-4411, -4049}, {16377, -181}, {-7555, 8279}, {-25199, -7917},
};

/* utils.rs:5-14 — smear the top bit down, keep only the top bit. */
size_t fri_oracle_prev_power_two(size_t x)
{
    size_t n = x;
    n |= n >> 1;
    n |= n >> 2;
    n |= n >> 4;
    n |= n >> 8;
    n |= n >> 16;
    return n ^ (n >> 1);
}

/* wavelet_transform.rs:71-90 */
void fri_oracle_nearby_vectors(int depth, int32_t out[6][2])
{
    int32_t zl[2], zmd[2];
    if (depth == 1) {
        zl[0] = -1; zl[1] = 1; zmd[0] = 0; zmd[1] = 2;
    } else if (depth == 2) {
        zl[0] = -2; zl[1] = 0; zmd[0] = 0; zmd[1] = -2;
    } else if (depth == 3) {
        zl[0] = -3; zl[1] = -1; zmd[0] = -1; zmd[1] = -3;
    } else {
        zl[0] = FRI_ORACLE_LITERALS[depth][0];
        zl[1] = FRI_ORACLE_LITERALS[depth][1];
        zmd[0] = FRI_ORACLE_LITERALS[depth + 1][0] + zl[0];
        zmd[1] = FRI_ORACLE_LITERALS[depth + 1][1] + zl[1];
    }
    for (int k = 0; k < 2; ++k) {
        out[0][k] = zl[k];
        out[1][k] = zl[k] - zmd[k];
        out[2][k] = -zmd[k];
        out[3][k] = -zl[k];
        out[4][k] = zmd[k] - zl[k];
        out[5][k] = zmd[k];
    }
}

/* wavelet_transform.rs:42-54: [0] = [1] = centre; [2p] = [p]; [2p+1] = [p] + LITERALS[depth-level-1] */
void fri_oracle_image_positions(int depth, int32_t cx, int32_t cy, int32_t *pos)
{
    pos[0] = cx; pos[1] = cy;
    pos[2] = cx; pos[3] = cy;
    for (int level = 0; level < depth; ++level) {
        const int32_t *lit = FRI_ORACLE_LITERALS[depth - level - 1];
        for (size_t p = (size_t)1 << level; p < (size_t)1 << (level + 1); ++p) {
            pos[2 * (2 * p) + 0] = pos[2 * p + 0];
            pos[2 * (2 * p) + 1] = pos[2 * p + 1];
            pos[2 * (2 * p + 1) + 0] = pos[2 * p + 0] + lit[0];
            pos[2 * (2 * p + 1) + 1] = pos[2 * p + 1] + lit[1];
        }
    }
}

/* ---- small open-addressing set of (re, im) keys, used by the BFS ------------------- */
typedef struct {
    int64_t *keys;
    size_t cap, n;
} keyset;
#define KS_EMPTY INT64_MIN

static int64_t ks_key(int32_t re, int32_t im) { return ((int64_t)re << 32) | (uint32_t)im; }
static size_t ks_hash(int64_t k, size_t cap)
{
    uint64_t h = (uint64_t)k * 0x9E3779B97F4A7C15ull;
    return (size_t)(h >> 17) & (cap - 1);
}
static void ks_init(keyset *s, size_t cap)
{
    s->cap = cap; s->n = 0;
    s->keys = (int64_t *)malloc(cap * sizeof(int64_t));
    for (size_t i = 0; i < cap; ++i) s->keys[i] = KS_EMPTY;
}
static int ks_has(const keyset *s, int64_t k)
{
    for (size_t i = ks_hash(k, s->cap);; i = (i + 1) & (s->cap - 1)) {
        if (s->keys[i] == k) return 1;
        if (s->keys[i] == KS_EMPTY) return 0;
    }
}
static void ks_put(keyset *s, int64_t k);
static void ks_grow(keyset *s)
{
    keyset t;
    ks_init(&t, s->cap * 2);
    for (size_t i = 0; i < s->cap; ++i)
        if (s->keys[i] != KS_EMPTY) ks_put(&t, s->keys[i]);
    free(s->keys);
    *s = t;
}
static void ks_put(keyset *s, int64_t k)
{
    if ((s->n + 1) * 2 > s->cap) ks_grow(s);
    for (size_t i = ks_hash(k, s->cap);; i = (i + 1) & (s->cap - 1)) {
        if (s->keys[i] == k) return;
        if (s->keys[i] == KS_EMPTY) { s->keys[i] = k; s->n++; return; }
    }
}

/*
 * wavelet_transform.rs:450-484.  The reference keeps a VecDeque `to_add`, a map
 * `fractal_lattice` and a `boundary` queue; a neighbour is queued unless it is already a
 * lattice key or already waiting in to_add (:470).  A position outside [0,w]x[0,h]
 * (inclusive, :459-463) is moved to `boundary` without being expanded and is inserted
 * afterwards (:478-481).  Note a boundary position can be queued more than once (it is in
 * neither container while it sits in `boundary`); the final HashMap insert dedups it.
 * Here `to_add.contains` is answered by a set mirroring the queue's content.
 */
int fri_oracle_fractal_divide(uint32_t width, uint32_t height, int depth, int32_t **centers_out,
                              size_t *n_out)
{
    int32_t vec[6][2];
    fri_oracle_nearby_vectors(depth, vec);

    size_t qcap = 1024, qhead = 0, qtail = 0;
    int32_t *queue = (int32_t *)malloc(qcap * 2 * sizeof(int32_t));
    size_t ocap = 1024, on = 0;
    int32_t *out = (int32_t *)malloc(ocap * 2 * sizeof(int32_t));
    size_t bcap = 1024, bn = 0;
    int32_t *boundary = (int32_t *)malloc(bcap * 2 * sizeof(int32_t));
    keyset lattice, queued;
    ks_init(&lattice, 1024);
    ks_init(&queued, 1024);
    if (!queue || !out || !boundary) return -1;

    queue[0] = (int32_t)width / 2;
    queue[1] = (int32_t)height / 2;
    qtail = 1;
    ks_put(&queued, ks_key(queue[0], queue[1]));

    while (qhead < qtail) {
        int32_t re = queue[2 * qhead], im = queue[2 * qhead + 1];
        qhead++;
        /* pop_front: the element leaves to_add.  A set cannot count duplicates, but an
         * in-bounds position is never queued twice (it is in `queued` until popped and in
         * `lattice` right after), and for out-of-bounds ones duplicates are harmless. */
        if (re < 0 || im < 0 || re > (int32_t)width || im > (int32_t)height) {
            if (bn == bcap) { bcap *= 2; boundary = (int32_t *)realloc(boundary, bcap * 2 * sizeof(int32_t)); }
            boundary[2 * bn] = re; boundary[2 * bn + 1] = im; bn++;
            continue;
        }
        for (int k = 0; k < 6; ++k) {
            int32_t nre = re + vec[k][0], nim = im + vec[k][1];
            int64_t key = ks_key(nre, nim);
            if (!ks_has(&lattice, key) && !ks_has(&queued, key)) {
                if (qtail == qcap) { qcap *= 2; queue = (int32_t *)realloc(queue, qcap * 2 * sizeof(int32_t)); }
                queue[2 * qtail] = nre; queue[2 * qtail + 1] = nim; qtail++;
                ks_put(&queued, key);
            }
        }
        ks_put(&lattice, ks_key(re, im));
        if (on == ocap) { ocap *= 2; out = (int32_t *)realloc(out, ocap * 2 * sizeof(int32_t)); }
        out[2 * on] = re; out[2 * on + 1] = im; on++;
    }
    for (size_t i = 0; i < bn; ++i) {
        int64_t key = ks_key(boundary[2 * i], boundary[2 * i + 1]);
        if (ks_has(&lattice, key)) continue;
        ks_put(&lattice, key);
        if (on == ocap) { ocap *= 2; out = (int32_t *)realloc(out, ocap * 2 * sizeof(int32_t)); }
        out[2 * on] = boundary[2 * i]; out[2 * on + 1] = boundary[2 * i + 1]; on++;
    }
    free(queue); free(boundary); free(lattice.keys); free(queued.keys);
    *centers_out = out;
    *n_out = on;
    return 0;
}

/* images.rs:89-100 */
fri_opt_i32 fri_oracle_get_pixel(const fri_oracle_raster *img, int32_t x, int32_t y, uint32_t channel)
{
    fri_opt_i32 r = {0, 0};
    if (x >= 0 && y >= 0 && x < (int32_t)img->width && y < (int32_t)img->height) {
        size_t position = ((size_t)y * img->width + (size_t)x) * img->channels + channel;
        r.v = img->sample_bytes == 2 ? (int32_t)((const uint16_t *)img->data)[position]
                                     : (int32_t)((const uint8_t *)img->data)[position];
        r.some = 1;
    }
    return r;
}

/* wavelet_transform.rs:14-26 with the two closures used at :212 and :216.
 * Integer arithmetic is done in uint32_t so that overflow wraps like release-mode Rust. */
static int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

static fri_opt_i32 apply_diff(fri_opt_i32 l, fri_opt_i32 r)
{
    fri_opt_i32 o = {0, 0};
    if (l.some && r.some) { o.v = wsub(l.v, r.v); o.some = 1; }
    else if (l.some)      { o.v = wsub(l.v, 0);   o.some = 1; }
    else if (r.some)      { o.v = wsub(0, r.v);   o.some = 1; }
    return o;
}
static fri_opt_i32 apply_lowpass(fri_opt_i32 r, fri_opt_i32 d)
{
    /* |l, r| l + r / 2  — Rust `/` on i32 truncates toward zero, as C99 `/` does. */
    fri_opt_i32 o = {0, 0};
    if (r.some && d.some) { o.v = wadd(r.v, d.v / 2); o.some = 1; }
    else if (r.some)      { o.v = wadd(r.v, 0 / 2);   o.some = 1; }
    else if (d.some)      { o.v = wadd(0, d.v / 2);   o.some = 1; }
    return o;
}

/* wavelet_transform.rs:179-225 */
void fri_oracle_extract_coefficients(const fri_oracle_raster *img, int depth, int32_t cx, int32_t cy,
                                     fri_opt_i32 *coef)
{
    const size_t n = (size_t)1 << depth;
    int32_t *pos = (int32_t *)malloc(2 * n * 2 * sizeof(int32_t));
    fri_opt_i32 *low = (fri_opt_i32 *)malloc(n * sizeof(fri_opt_i32));
    fri_oracle_image_positions(depth, cx, cy, pos);
    for (uint32_t ch = 0; ch < img->channels; ++ch) {
        fri_opt_i32 *c = coef + (size_t)ch * n;
        memset(c, 0, n * sizeof(fri_opt_i32));
        memset(low, 0, n * sizeof(fri_opt_i32));
        for (int level = depth - 1; level >= 0; --level) {
            for (size_t p = (size_t)1 << level; p < (size_t)1 << (level + 1); ++p) {
                fri_opt_i32 l, r;
                if (level == depth - 1) {
                    l = fri_oracle_get_pixel(img, pos[2 * (2 * p)], pos[2 * (2 * p) + 1], ch);
                    r = fri_oracle_get_pixel(img, pos[2 * (2 * p + 1)], pos[2 * (2 * p + 1) + 1], ch);
                } else {
                    l = low[2 * p];
                    r = low[2 * p + 1];
                }
                c[p] = apply_diff(l, r);
                low[p] = apply_lowpass(r, c[p]);
            }
        }
        c[0] = low[1];
    }
    free(pos);
    free(low);
}


/* ---- tiny pthread parallel-for (the reference is single-threaded; nthreads > 1 is only
 * used by bench.py's "all host cores" CPU arm) ---------------------------------------- */
#include <pthread.h>
typedef void (*pf_body)(size_t t, void *ctx);
typedef struct { pf_body body; void *ctx; size_t lo, hi; } pf_job;
static void *pf_run(void *p)
{
    pf_job *j = (pf_job *)p;
    for (size_t t = j->lo; t < j->hi; ++t) j->body(t, j->ctx);
    return NULL;
}
static void parallel_for(size_t n, int nthreads, pf_body body, void *ctx)
{
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    if (nthreads == 1) { for (size_t t = 0; t < n; ++t) body(t, ctx); return; }
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    pf_job *jobs = (pf_job *)malloc((size_t)nthreads * sizeof(pf_job));
    for (int i = 0; i < nthreads; ++i) {
        jobs[i].body = body; jobs[i].ctx = ctx;
        jobs[i].lo = n * (size_t)i / (size_t)nthreads;
        jobs[i].hi = n * (size_t)(i + 1) / (size_t)nthreads;
        pthread_create(&th[i], NULL, pf_run, &jobs[i]);
    }
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th); free(jobs);
}

typedef struct {
    const fri_oracle_raster *img; int depth; const int32_t *centers; int32_t *coef; uint8_t *some;
} xt_ctx;
static void xt_body(size_t t, void *p)
{
    xt_ctx *c = (xt_ctx *)p;
    const size_t per = (size_t)c->img->channels << c->depth;
    fri_opt_i32 *tmp = (fri_opt_i32 *)malloc(per * sizeof(fri_opt_i32));
    fri_oracle_extract_coefficients(c->img, c->depth, c->centers[2 * t], c->centers[2 * t + 1], tmp);
    for (size_t i = 0; i < per; ++i) {
        c->coef[t * per + i] = tmp[i].some ? tmp[i].v : 0;
        if (c->some) c->some[t * per + i] = tmp[i].some;
    }
    free(tmp);
}

void fri_oracle_extract_tiles(const fri_oracle_raster *img, int depth, const int32_t *centers, size_t n,
                              int32_t *coef, uint8_t *some, int nthreads)
{
    xt_ctx c = {img, depth, centers, coef, some};
    parallel_for(n, nthreads, xt_body, &c);
}

static int cmp_center(const void *a, const void *b)
{
    const int32_t *x = (const int32_t *)a, *y = (const int32_t *)b;
    if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return 0;
}

/* wavelet_transform.rs:405-416 (retain predicate over the active channels — see header). */
int fri_oracle_from_raster(const fri_oracle_raster *img, int depth, int32_t **centers_out,
                           int32_t **coef_out, uint8_t **some_out, size_t *n_out)
{
    int32_t *built = NULL;
    size_t nb = 0;
    if (fri_oracle_fractal_divide(img->width, img->height, depth, &built, &nb)) return -1;
    qsort(built, nb, 2 * sizeof(int32_t), cmp_center);

    const size_t per = (size_t)img->channels << depth;
    int32_t *centers = (int32_t *)malloc((nb ? nb : 1) * 2 * sizeof(int32_t));
    int32_t *coef = (int32_t *)malloc((nb ? nb : 1) * per * sizeof(int32_t));
    uint8_t *some = (uint8_t *)malloc((nb ? nb : 1) * per);
    fri_opt_i32 *tmp = (fri_opt_i32 *)malloc(per * sizeof(fri_opt_i32));
    size_t n = 0;
    for (size_t t = 0; t < nb; ++t) {
        fri_oracle_extract_coefficients(img, depth, built[2 * t], built[2 * t + 1], tmp);
        int keep = 1;
        for (uint32_t ch = 0; ch < img->channels; ++ch)
            keep &= tmp[(size_t)ch << depth].some;
        if (!keep) continue;
        centers[2 * n] = built[2 * t];
        centers[2 * n + 1] = built[2 * t + 1];
        for (size_t i = 0; i < per; ++i) {
            coef[n * per + i] = tmp[i].some ? tmp[i].v : 0;
            some[n * per + i] = tmp[i].some;
        }
        n++;
    }
    free(tmp);
    free(built);
    *centers_out = centers; *coef_out = coef; *some_out = some; *n_out = n;
    return 0;
}

/* quantization.rs:7-25 / :27-45.  layer = trailing_zeros(prev_power_two(i + 1)). */
void fri_oracle_quantize(int32_t *coef, const uint8_t *some, size_t n_tiles, uint32_t channels, int depth,
                         const int32_t q[32], int multiply)
{
    const size_t n = (size_t)1 << depth;
    for (size_t t = 0; t < n_tiles * channels; ++t) {
        for (size_t i = 0; i < n; ++i) {
            if (some && !some[t * n + i]) continue;
            unsigned layer = (unsigned)__builtin_ctzll((unsigned long long)fri_oracle_prev_power_two(i + 1));
            int32_t *c = &coef[t * n + i];
            if (multiply) *c = (int32_t)((uint32_t)*c * (uint32_t)q[layer]);
            else if (q[layer] == -1) *c = wsub(0, *c); /* avoid INT_MIN / -1 trap */
            else *c = *c / q[layer];
        }
    }
}

static void set_pixel(void *out, uint32_t width, uint32_t height, uint32_t channels, uint32_t sample_bytes,
                      int32_t x, int32_t y, int32_t value, uint32_t ch)
{
    /* images.rs:103-111; the u16 branch is the 16-bit extension (clamp to the sample range) */
    if (x >= 0 && y >= 0 && x < (int32_t)width && y < (int32_t)height) {
        size_t position = ((size_t)y * width + (size_t)x) * channels + ch;
        if (sample_bytes == 2) {
            int32_t v = value < 0 ? 0 : (value > 65535 ? 65535 : value);
            ((uint16_t *)out)[position] = (uint16_t)v;
        } else {
            int32_t v = value < 0 ? 0 : (value > 255 ? 255 : value);
            ((uint8_t *)out)[position] = (uint8_t)v;
        }
    }
}

/* wavelet_transform.rs:358-381 for each tile of the lattice (:319-321). */
typedef struct {
    const int32_t *centers, *coef; const uint8_t *some; int depth;
    uint32_t width, height, channels, sample_bytes; void *out;
} xv_ctx;
static void xv_body(size_t t, void *p)
{
    xv_ctx *x = (xv_ctx *)p;
    const int depth = x->depth;
    const size_t n = (size_t)1 << depth;
    int32_t *pos = (int32_t *)malloc(2 * n * 2 * sizeof(int32_t));
    int32_t *low = (int32_t *)malloc(n * sizeof(int32_t));
    fri_oracle_image_positions(depth, x->centers[2 * t], x->centers[2 * t + 1], pos);
    for (uint32_t ch = 0; ch < x->channels; ++ch) {
        const int32_t *c = x->coef + (t * x->channels + ch) * n;
        const uint8_t *s = x->some ? x->some + (t * x->channels + ch) * n : NULL;
        memset(low, 0, n * sizeof(int32_t));
        low[1] = c[0]; /* .unwrap(): retained tiles always have Some(DC) */
        for (int level = 0; level < depth; ++level) {
            for (size_t q = (size_t)1 << level; q < (size_t)1 << (level + 1); ++q) {
                if (s && !s[q]) continue; /* :365 `if let Some(dif)` */
                int32_t dif = c[q];
                int32_t right = wsub(low[q], dif / 2);
                int32_t left = wadd(dif, right);
                if (level == depth - 1) {
                    set_pixel(x->out, x->width, x->height, x->channels, x->sample_bytes, pos[2 * (2 * q)],
                              pos[2 * (2 * q) + 1], left, ch);
                    set_pixel(x->out, x->width, x->height, x->channels, x->sample_bytes, pos[2 * (2 * q + 1)],
                              pos[2 * (2 * q + 1) + 1], right, ch);
                } else {
                    low[2 * q] = left;
                    low[2 * q + 1] = right;
                }
            }
        }
    }
    free(pos);
    free(low);
}

void fri_oracle_extract_values(const int32_t *centers, const int32_t *coef, const uint8_t *some, size_t n_tiles,
                               int depth, uint32_t width, uint32_t height, uint32_t channels,
                               uint32_t sample_bytes, void *out, int nthreads)
{
    xv_ctx x = {centers, coef, some, depth, width, height, channels, sample_bytes, out};
    parallel_for(n_tiles, nthreads, xv_body, &x);
}

/* ---- whole-path drivers used as the timed CPU baseline (bench.py): the reference's
 * wavelet_transform::encode + quantization::encode (encoder.rs:26-33) and quantization::decode +
 * wavelet_transform::decode (decoder.rs:27-34) over a given tile list, tiles split over threads. */
typedef struct {
    const fri_oracle_raster *img; int depth; const int32_t *centers; int32_t *coef; const int32_t *q;
} et_ctx;
static void et_body(size_t t, void *p)
{
    et_ctx *c = (et_ctx *)p;
    const size_t n = (size_t)1 << c->depth, per = (size_t)c->img->channels * n;
    fri_opt_i32 *tmp = (fri_opt_i32 *)malloc(per * sizeof(fri_opt_i32));
    fri_oracle_extract_coefficients(c->img, c->depth, c->centers[2 * t], c->centers[2 * t + 1], tmp);
    for (size_t i = 0; i < per; ++i) {
        unsigned layer = (unsigned)__builtin_ctzll((unsigned long long)fri_oracle_prev_power_two((i & (n - 1)) + 1));
        c->coef[t * per + i] = tmp[i].some ? tmp[i].v / c->q[layer] : 0;
    }
    free(tmp);
}
void fri_oracle_encode_tiles(const fri_oracle_raster *img, int depth, const int32_t *centers, size_t n_tiles,
                             const int32_t q[32], int32_t *coef, int nthreads)
{
    et_ctx c = {img, depth, centers, coef, q};
    parallel_for(n_tiles, nthreads, et_body, &c);
}

typedef struct { xv_ctx x; const int32_t *q; int32_t *tmp_all; } dt_ctx;
static void dt_body(size_t t, void *p)
{
    dt_ctx *d = (dt_ctx *)p;
    const size_t n = (size_t)1 << d->x.depth, per = (size_t)d->x.channels * n;
    int32_t *tmp = (int32_t *)malloc(per * sizeof(int32_t));
    for (size_t i = 0; i < per; ++i) {
        unsigned layer = (unsigned)__builtin_ctzll((unsigned long long)fri_oracle_prev_power_two((i & (n - 1)) + 1));
        tmp[i] = d->x.coef[t * per + i] / d->q[layer]; /* quantization.rs:37 divides */
    }
    xv_ctx one = d->x;
    one.centers = d->x.centers + 2 * t;
    one.coef = tmp;
    one.some = d->x.some ? d->x.some + t * per : NULL;
    xv_body(0, &one);
    free(tmp);
}
void fri_oracle_decode_tiles(const int32_t *centers, const int32_t *coef, const uint8_t *some, size_t n_tiles,
                             int depth, uint32_t width, uint32_t height, uint32_t channels, uint32_t sample_bytes,
                             const int32_t q[32], void *out, int nthreads)
{
    dt_ctx d = {{centers, coef, some, depth, width, height, channels, sample_bytes, out}, q, NULL};
    parallel_for(n_tiles, nthreads, dt_body, &d);
}

/* ------------------------------------------------------------------------------------------
 * "Reference-shaped" cost estimate — NOT a restatement, a cost model.
 * The reference spends most of its transform time not in the lifting arithmetic but in the
 * containers Fractal::new builds for every tile (wavelet_transform.rs:42-69): image_positions
 * (Vec of 2^(depth+1) Complex<i32>), one HashMap<Complex<i32>, usize> per tree level with 2^level
 * inserts (std HashMap = hashbrown with SipHash-1-3 and doubling growth from empty), the
 * coefficient and value Vecs.  This function performs the same allocations, hashes and inserts for
 * every tile of a list so that bench.py can report "arithmetic + containers" beside the
 * arithmetic-only figure.  It returns a checksum so the work cannot be optimised away.
 * ------------------------------------------------------------------------------------------ */
#define ROTL64(x, b) (uint64_t)(((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND(v0, v1, v2, v3) \
    do { v0 += v1; v1 = ROTL64(v1, 13); v1 ^= v0; v0 = ROTL64(v0, 32); v2 += v3; v3 = ROTL64(v3, 16); v3 ^= v2; \
         v0 += v3; v3 = ROTL64(v3, 21); v3 ^= v0; v2 += v1; v1 = ROTL64(v1, 17); v1 ^= v2; v2 = ROTL64(v2, 32); } while (0)

static uint64_t siphash13_u64(uint64_t m, uint64_t k0, uint64_t k1)
{
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL, v1 = k1 ^ 0x646f72616e646f6dULL;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ULL, v3 = k1 ^ 0x7465646279746573ULL;
    v3 ^= m; SIPROUND(v0, v1, v2, v3); v0 ^= m;            /* one 8-byte block, 1 compression round */
    const uint64_t b = (uint64_t)8 << 56;                   /* length byte */
    v3 ^= b; SIPROUND(v0, v1, v2, v3); v0 ^= b;
    v2 ^= 0xff; SIPROUND(v0, v1, v2, v3); SIPROUND(v0, v1, v2, v3); SIPROUND(v0, v1, v2, v3);  /* 3 finalisation rounds */
    return v0 ^ v1 ^ v2 ^ v3;
}

typedef struct { uint64_t *keys; size_t *vals; uint8_t *used; size_t cap, len; } pos_map;

static void pm_insert_raw(pos_map *m, uint64_t key, size_t val)
{
    size_t i = (size_t)siphash13_u64(key, 0x0706050403020100ULL, 0x0f0e0d0c0b0a0908ULL) & (m->cap - 1);
    while (m->used[i]) {
        if (m->keys[i] == key) { m->vals[i] = val; return; }
        i = (i + 1) & (m->cap - 1);
    }
    m->used[i] = 1; m->keys[i] = key; m->vals[i] = val; ++m->len;
}

static void pm_insert(pos_map *m, uint64_t key, size_t val)
{
    if (m->cap == 0 || (m->len + 1) * 8 > m->cap * 7) {  /* hashbrown: grow at 7/8 load, double, re-insert */
        pos_map n;
        n.cap = m->cap ? m->cap * 2 : 4;
        n.len = 0;
        n.keys = (uint64_t *)malloc(n.cap * sizeof(uint64_t));
        n.vals = (size_t *)malloc(n.cap * sizeof(size_t));
        n.used = (uint8_t *)calloc(n.cap, 1);
        for (size_t i = 0; i < m->cap; ++i)
            if (m->used[i]) pm_insert_raw(&n, m->keys[i], m->vals[i]);
        free(m->keys); free(m->vals); free(m->used);
        *m = n;
    }
    pm_insert_raw(m, key, val);
}

uint64_t fri_oracle_fractal_new_cost(int depth, const int32_t *centers, size_t n_tiles, uint32_t channels)
{
    uint64_t sum = 0;
    const size_t nodes = (size_t)1 << (depth + 1), leaves = (size_t)1 << depth;
    for (size_t t = 0; t < n_tiles; ++t) {
        int32_t *pos = (int32_t *)malloc(nodes * 2 * sizeof(int32_t));  /* image_positions */
        fri_oracle_image_positions(depth, centers[2 * t], centers[2 * t + 1], pos);
        pos_map *maps = (pos_map *)calloc((size_t)depth, sizeof(pos_map));  /* position_map[level] */
        for (int level = 0; level < depth; ++level)
            for (size_t p = (size_t)1 << level; p < ((size_t)2 << level); ++p)
                pm_insert(&maps[level], (uint64_t)ks_key(pos[2 * p], pos[2 * p + 1]), p);
        /* coefficients: [channels] Vec<Option<i32>> (8 B each), values: Vec<Option<i32>> per channel */
        fri_opt_i32 *coef = (fri_opt_i32 *)calloc((size_t)channels * leaves, sizeof(fri_opt_i32));
        fri_opt_i32 *vals = (fri_opt_i32 *)calloc((size_t)channels * nodes, sizeof(fri_opt_i32));
        for (int level = 0; level < depth; ++level) {
            sum += maps[level].len + (maps[level].used[maps[level].cap / 2] ? maps[level].vals[maps[level].cap / 2] : 0);
            free(maps[level].keys); free(maps[level].vals); free(maps[level].used);
        }
        sum += (uint64_t)coef[leaves - 1].v + (uint64_t)vals[nodes - 1].v + (uint64_t)(uint32_t)pos[2 * nodes - 1];
        free(maps); free(coef); free(vals); free(pos);
    }
    return sum;
}

void fri_oracle_free(void *p) { free(p); }
