/* FLOAT64 TRUTH MODEL -- TEST INFRASTRUCTURE ONLY (same rules as sspsd_oracle.c: never linked or loaded
 * by the product package).
 *
 * A streaming float64 model of the reference's cascaded PSD (src/psd.rs:196-269, 445-468) used for the
 * accuracy statements of SURVEY.md A.7 at the FULL size of the BASELINE configs (200e6 ... 4.8e9 samples),
 * where oracle/model_f64.py (whole-stream numpy) does not fit in memory.  It is written from the closed-form
 * semantics of SURVEY.md App. A.2, deliberately NOT from sspsd_oracle.c:
 *   - stage i sees the stream s_i; segment k covers s_i[k hop, k hop + N); after L samples
 *     craw = L < N ? 0 : 1 + (L - N) / hop segments are complete and D = N + (craw - 1) hop samples
 *     have been decimated (psd.rs:235-253); s_{i+1} = decimate8(s_i[0, D)) without its first R outputs
 *     (psd.rs:254-260);
 *   - decimate8 = three direct-form half-band FIRs y[j] = sum_k h[k] x[2j + 1 - k] (zero history) with the
 *     full 4M-1 tap impulse response (the f32 oracle uses a folded polyphase form);
 *   - the FFT is a plain iterative radix-2 complex transform (the f32 oracle uses a four-step Stockham);
 *   - averaging psd.rs:215-233 with the f32 factor g = avg / count promoted to f64.
 * The window table and the half-band taps are DATA of the algorithm: they are taken as the f32 values the
 * reference would compute (orc_window, hbf_taps.h) and promoted, so that differences to this model measure
 * arithmetic rounding only.  Segments of a block are transformed in parallel (pthreads; the image's gcc has
 * no libgomp); all sums are f64.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <unistd.h>

#include "hbf_taps.h"
#include "sspsd_oracle.h"

/* ---- minimal parallel-for: body(ctx, lo, hi, thread index) over [0, n) split into contiguous ranges ---- */
#define F64_MAX_THREADS 32
typedef void (*par_body)(void *ctx, uint64_t lo, uint64_t hi, int t);
typedef struct {
    par_body body;
    void *ctx;
    uint64_t lo, hi;
    int t;
} par_job;
static void *par_tramp(void *a)
{
    par_job *j = (par_job *)a;
    j->body(j->ctx, j->lo, j->hi, j->t);
    return NULL;
}
static int par_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    const char *e = getenv("SSPSD_F64_THREADS");
    if (e)
        n = atol(e);
    return n < 1 ? 1 : n > F64_MAX_THREADS ? F64_MAX_THREADS : (int)n;
}
static void par_for(uint64_t n, par_body body, void *ctx)
{
    int nt = par_threads();
    if ((uint64_t)nt > n)
        nt = (int)(n ? n : 1);
    par_job jobs[F64_MAX_THREADS];
    pthread_t th[F64_MAX_THREADS];
    for (int t = 0; t < nt; t++) {
        jobs[t] = (par_job){body, ctx, n * (uint64_t)t / (uint64_t)nt, n * (uint64_t)(t + 1) / (uint64_t)nt, t};
        if (t > 0 && pthread_create(&th[t], NULL, par_tramp, &jobs[t]) != 0)
            abort();
    }
    par_tramp(&jobs[0]);
    for (int t = 1; t < nt; t++)
        pthread_join(th[t], NULL);
}

#define F64_MAX_STAGES 24
#define F64_DEPTH 3

typedef struct {
    int m;        /* unique taps */
    int len;      /* 4m - 1 */
    double *h;    /* full impulse response */
    double *hist; /* last len - 1 inputs */
} fir2;

typedef struct {
    uint64_t L, craw, count, emitted; /* samples received, segments done, effective count, handed on */
    uint32_t avg;
    double *spec;  /* N/2 + 1 */
    double *buf;   /* samples [base, L) */
    uint64_t base;
    size_t cap;
    fir2 fir[3];
    uint64_t drain_left;
} f64_stage;

typedef struct f64_cascade {
    int n, hop, overlap, window, preset, detrend, max_stages;
    uint32_t avg_limit, avg_count;
    double *win;
    double *twr, *twi; /* radix-2 twiddles W_N^k, k < N/2 */
    int *rev;
    f64_stage st[F64_MAX_STAGES];
    int n_stages;
} f64_cascade;

static void fir2_init(fir2 *f, const float *t, int m)
{
    f->m = m;
    f->len = 4 * m - 1;
    f->h = (double *)calloc((size_t)f->len, sizeof(double));
    f->hist = (double *)calloc((size_t)f->len, sizeof(double));
    const int c = 2 * m - 1;
    f->h[c] = 1.0; /* DC gain 2 per half-band stage: centre tap 1, odd taps sum to 1 */
    for (int i = 0; i < m; i++) {
        const int k = m - 1 - i;
        f->h[c - (2 * k + 1)] = (double)t[i];
        f->h[c + (2 * k + 1)] = (double)t[i];
    }
}

typedef struct {
    const fir2 *f;
    const double *e;
    double *y;
} fir_ctx;
static void fir_body(void *vc, uint64_t lo, uint64_t hi, int t)
{
    (void)t;
    const fir_ctx *c = (const fir_ctx *)vc;
    const double *h = c->f->h;
    const int len = c->f->len, hl = len - 1;
    for (uint64_t j = lo; j < hi; j++) {
        /* y[j] = sum_k h[k] x[2j + 1 - k]; x[p] sits at e[p + hl] */
        const double *top = c->e + 2 * j + 1 + hl;
        double acc = 0.0;
        for (int k = 0; k < len; k++)
            acc += h[k] * top[-k];
        c->y[j] = acc;
    }
}

/* x: n (even) new inputs -> y: n / 2 outputs */
static void fir2_run(fir2 *f, const double *x, size_t n, double *y)
{
    const int hl = f->len - 1;
    double *e = (double *)malloc((n + (size_t)hl) * sizeof(double));
    memcpy(e, f->hist, (size_t)hl * sizeof(double));
    memcpy(e + hl, x, n * sizeof(double));
    fir_ctx c = {f, e, y};
    par_for(n / 2, fir_body, &c);
    memcpy(f->hist, e + n, (size_t)hl * sizeof(double));
    free(e);
}

f64_cascade *f64_cascade_new(int n, int window, int preset, int detrend, uint32_t avg_limit, uint32_t avg_count,
                             int max_stages)
{
    if (n < 16 || (n & (n - 1)) || preset < 0 || preset >= ORC_HBF_NPRESET || detrend < 0 || detrend > 3)
        return NULL;
    f64_cascade *c = (f64_cascade *)calloc(1, sizeof(*c));
    c->n = n;
    c->window = window;
    c->preset = preset;
    c->detrend = detrend;
    c->avg_limit = avg_limit;
    c->avg_count = avg_count;
    c->max_stages = max_stages > 0 && max_stages < F64_MAX_STAGES ? max_stages : F64_MAX_STAGES;
    float *w32 = (float *)malloc(sizeof(float) * (size_t)n);
    float power, nenbw;
    size_t ov;
    orc_window(n, window, w32, &power, &nenbw, &ov); /* the reference's f32 table (psd.rs:24-55), promoted */
    c->overlap = (int)ov;
    c->hop = n - c->overlap;
    c->win = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++)
        c->win[i] = (double)w32[i];
    free(w32);
    c->twr = (double *)malloc(sizeof(double) * (size_t)(n / 2));
    c->twi = (double *)malloc(sizeof(double) * (size_t)(n / 2));
    for (int k = 0; k < n / 2; k++) {
        c->twr[k] = cos(-2.0 * M_PI * (double)k / (double)n);
        c->twi[k] = sin(-2.0 * M_PI * (double)k / (double)n);
    }
    c->rev = (int *)malloc(sizeof(int) * (size_t)n);
    int lg = 0;
    while ((1 << lg) < n)
        lg++;
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < lg; b++)
            if (i & (1 << b))
                r |= 1 << (lg - 1 - b);
        c->rev[i] = r;
    }
    return c;
}

void f64_cascade_free(f64_cascade *c)
{
    if (!c)
        return;
    for (int i = 0; i < c->n_stages; i++) {
        free(c->st[i].spec);
        free(c->st[i].buf);
        for (int s = 0; s < 3; s++) {
            free(c->st[i].fir[s].h);
            free(c->st[i].fir[s].hist);
        }
    }
    free(c->win);
    free(c->twr);
    free(c->twi);
    free(c->rev);
    free(c);
}

static f64_stage *add_stage(f64_cascade *c)
{
    f64_stage *s = &c->st[c->n_stages];
    memset(s, 0, sizeof(*s));
    const int i = c->n_stages;
    const unsigned sh = (unsigned)(F64_DEPTH * i);
    uint32_t a = sh >= 32 ? 0u : (c->avg_count >> sh); /* psd.rs:434,449 */
    s->avg = a < c->avg_limit ? a : c->avg_limit;
    s->spec = (double *)calloc((size_t)c->n / 2 + 1, sizeof(double));
    for (int q = 0; q < 3; q++) {
        const int set = 2 - q; /* highest rate first */
        fir2_init(&s->fir[q], orc_hbf_taps[c->preset][set], orc_hbf_ntaps[c->preset][set]);
    }
    s->drain_left = (uint64_t)orc_hbf_drain[c->preset];
    c->n_stages++;
    return s;
}

/* |FFT(detrend(seg) * win)|^2 for bins 0..N/2 into p; work: 2N doubles */
static void segment_power(const f64_cascade *c, const double *seg, double *work, double *p)
{
    const int n = c->n;
    double *re = work, *im = work + n;
    double off = 0.0, slope = 0.0;
    switch (c->detrend) {
    case 1: off = seg[n / 2]; break;                                         /* psd.rs:87-93 */
    case 2: off = seg[0]; slope = (seg[n - 1] - seg[0]) / (double)(n - 1); break; /* psd.rs:94-102 */
    case 3: {                                                                /* psd.rs:103-109 */
        double sum = 0.0;
        for (int i = 0; i < n; i++)
            sum += seg[i];
        off = sum / (double)n;
        break;
    }
    default: break;
    }
    for (int i = 0; i < n; i++) {
        const int r = c->rev[i];
        re[r] = (seg[i] - (off + slope * (double)i)) * c->win[i];
        im[r] = 0.0;
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len / 2, step = n / len;
        for (int b = 0; b < n; b += len)
            for (int k = 0; k < half; k++) {
                const double wr = c->twr[k * step], wi = c->twi[k * step];
                const int u = b + k, v = u + half;
                const double tr = re[v] * wr - im[v] * wi, ti = re[v] * wi + im[v] * wr;
                re[v] = re[u] - tr;
                im[v] = im[u] - ti;
                re[u] += tr;
                im[u] += ti;
            }
    }
    for (int k = 0; k <= n / 2; k++)
        p[k] = re[k] * re[k] + im[k] * im[k];
}

typedef struct {
    const f64_cascade *c;
    const f64_stage *s;
    uint64_t k0;                   /* first segment of the range */
    double *P;                     /* if set: per-segment powers [k][nb]; else per-thread sums in acc */
    double *acc[F64_MAX_THREADS];
} seg_ctx;
static void seg_body(void *vc, uint64_t lo, uint64_t hi, int t)
{
    seg_ctx *sc = (seg_ctx *)vc;
    const f64_cascade *c = sc->c;
    const int n = c->n, nb = n / 2 + 1;
    double *work = (double *)malloc(sizeof(double) * (size_t)(2 * n + nb));
    double *p = work + 2 * n;
    double *acc = NULL;
    if (!sc->P)
        acc = sc->acc[t] = (double *)calloc((size_t)nb, sizeof(double));
    for (uint64_t k = lo; k < hi; k++) {
        const double *seg = sc->s->buf + ((sc->k0 + k) * (uint64_t)c->hop - sc->s->base);
        if (sc->P) {
            segment_power(c, seg, work, sc->P + k * (uint64_t)nb);
        } else {
            segment_power(c, seg, work, p);
            for (int q = 0; q < nb; q++)
                acc[q] += p[q];
        }
    }
    free(work);
}

static void stage_feed(f64_cascade *c, int i, const double *x, size_t nx)
{
    if (nx == 0)
        return;
    if (i >= c->n_stages)
        add_stage(c);
    f64_stage *s = &c->st[i];
    const int n = c->n, hop = c->hop;
    const size_t have = (size_t)(s->L - s->base);
    if (have + nx > s->cap) {
        s->cap = (have + nx) * 2 + 64;
        s->buf = (double *)realloc(s->buf, s->cap * sizeof(double));
    }
    memcpy(s->buf + have, x, nx * sizeof(double));
    s->L += nx;
    const uint64_t craw1 = s->L < (uint64_t)n ? 0 : 1 + (s->L - (uint64_t)n) / (uint64_t)hop;
    const uint64_t k0 = s->craw, ks = craw1 - k0;
    const int nb = n / 2 + 1;
    if (ks > 0) {
        seg_ctx sc = {c, s, k0, NULL, {NULL}};
        const int boxcar = (s->count + ks) <= (uint64_t)s->avg;
        if (boxcar) {
            par_for(ks, seg_body, &sc);
            for (int t = 0; t < F64_MAX_THREADS; t++)
                if (sc.acc[t]) {
                    for (int q = 0; q < nb; q++)
                        s->spec[q] += sc.acc[t][q];
                    free(sc.acc[t]);
                }
            s->count += ks;
        } else {
            /* EWMA is a recurrence over segments: powers in parallel per block, recurrence in order */
            const uint64_t blk = 256;
            sc.P = (double *)malloc(sizeof(double) * (size_t)blk * (size_t)nb);
            for (uint64_t b0 = 0; b0 < ks; b0 += blk) {
                const uint64_t bn = ks - b0 < blk ? ks - b0 : blk;
                sc.k0 = k0 + b0;
                par_for(bn, seg_body, &sc);
                for (uint64_t k = 0; k < bn; k++) {
                    double g = 1.0;
                    if (s->count > (uint64_t)s->avg) { /* psd.rs:218-225 */
                        g = (double)((float)s->avg / (float)(uint32_t)s->count);
                        s->count = s->avg;
                    }
                    s->count += 1;
                    const double *p = sc.P + k * (uint64_t)nb;
                    for (int q = 0; q < nb; q++)
                        s->spec[q] = g * s->spec[q] + p[q];
                }
            }
            free(sc.P);
        }
    }
    const uint64_t D0 = s->craw ? (uint64_t)n + (s->craw - 1) * (uint64_t)hop : 0;
    s->craw = craw1;
    const uint64_t D1 = s->craw ? (uint64_t)n + (s->craw - 1) * (uint64_t)hop : 0;
    if (D1 > D0 && i + 1 < c->max_stages) {
        size_t nin = (size_t)(D1 - D0);
        double *a = (double *)malloc(sizeof(double) * (nin / 2 + 1));
        double *b = (double *)malloc(sizeof(double) * (nin / 4 + 1));
        double *y = (double *)malloc(sizeof(double) * (nin / 8 + 1));
        fir2_run(&s->fir[0], s->buf + (D0 - s->base), nin, a);
        fir2_run(&s->fir[1], a, nin / 2, b);
        fir2_run(&s->fir[2], b, nin / 4, y);
        size_t ny = nin / 8, skip = 0;
        if (s->drain_left) { /* psd.rs:254-260 */
            skip = s->drain_left < ny ? (size_t)s->drain_left : ny;
            s->drain_left -= skip;
        }
        free(a);
        free(b);
        s->emitted += ny - skip;
        stage_feed(c, i + 1, y + skip, ny - skip); /* (may realloc c->st[..].buf of deeper stages only) */
        free(y);
    }
    /* drop what no later segment or decimation needs: everything before craw * hop (<= D1) */
    s = &c->st[i];
    const uint64_t keep_from = s->craw * (uint64_t)hop;
    if (keep_from > s->base) {
        memmove(s->buf, s->buf + (keep_from - s->base), (size_t)(s->L - keep_from) * sizeof(double));
        s->base = keep_from;
    }
}

void f64_cascade_process(f64_cascade *c, const float *x, size_t n)
{
    const size_t blk = (size_t)1 << 24;
    double *d = (double *)malloc(sizeof(double) * (n < blk ? n : blk));
    for (size_t pos = 0; pos < n; pos += blk) {
        const size_t m = n - pos < blk ? n - pos : blk;
        for (size_t i = 0; i < m; i++)
            d[i] = (double)x[pos + i];
        stage_feed(c, 0, d, m);
    }
    free(d);
}

int f64_cascade_num_stages(const f64_cascade *c) { return c->n_stages; }

/* spectrum: N/2+1 accumulated powers; count = effective averaging count; craw = segments; L = samples */
void f64_cascade_stage(const f64_cascade *c, int i, double *spectrum, uint64_t *count, uint64_t *craw, uint64_t *L)
{
    const f64_stage *s = &c->st[i];
    if (spectrum)
        memcpy(spectrum, s->spec, sizeof(double) * ((size_t)c->n / 2 + 1));
    if (count)
        *count = s->count;
    if (craw)
        *craw = s->craw;
    if (L)
        *L = s->L;
}
