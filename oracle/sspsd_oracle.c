/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see sspsd_oracle.h for scope and parity status).
 *
 * Restates, in plain C and in the reference's own f32 evaluation order (no FMA contraction:
 * build with -ffp-contract=off), the hot path of quartiq/stabilizer-stream:
 *   src/psd.rs      Window, Detrend, Psd<N>, PsdCascade<N>, Break, MergeOpts, AvgOpts
 *   src/de/{frame,data}.rs  Header::parse, Frame::from_bytes, AdcDac/Fls/ThermostatEem/Mpll::traces
 *   src/loss.rs     Loss::update / analyze
 *   src/var.rs      Var::eval
 * plus the two crates.io dependencies on that path that are not on disk:
 *   rustfft 6.4.1   -> orc_fft_forward (own Stockham radix-4/2 FFT, same DFT definition)
 *   idsp 0.20.0 hbf -> orc_hbf8_* (taps from tools/gen_hbf_taps.py; PARITY UNPINNED)
 */
#include "sspsd_oracle.h"
#include "hbf_taps.h"

#include <assert.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * idsp::hbf restatement
 * One half-band decimate-by-2 stage with M unique taps is the causal linear-phase FIR
 *   y[j] = x[c] + sum_{i=0}^{M-1} t[i] * (x[c-(2k+1)] + x[c+(2k+1)]),
 *   k = M-1-i,  c = 2*j - 2*M + 2,   newest input used: x[2*j+1]
 * with zero initial state.  The taps are 2*remez(...) (centre tap 1.0, sum of t = 0.5), so every
 * stage has DC gain 2 and the divide-by-8 cascade has amplitude gain 8.  That gain is what makes
 * the reference's stage normalisation 1/(gain * decimation) (psd.rs:516) continuous across stage
 * breaks for noise: power gain 64, bandwidth 1/8 => net 8 = decimation per stage (this is what the
 * reference test psd.rs:634-643 asserts: 0.5*p ~ 1 in every break).  HbfDec8 chains three of them, highest rate first, using tap sets
 * 2, 1, 0 (the lowest-rate stage has the longest filter).  psd.rs:246-253 feeds it chunks of
 * 8 items and gets one item per chunk.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int m;            /* unique taps */
    const float *t;   /* outermost -> innermost */
    float *hist;      /* last 4m-2 inputs */
} hbf2;

struct orc_hbf8 {
    int preset;
    hbf2 st[3]; /* st[0] = highest rate (tap set 2) ... st[2] = lowest rate (tap set 0) */
    float *work[3];
    size_t work_cap;
};

orc_hbf8 *orc_hbf8_new(int preset)
{
    if (preset < 0 || preset >= ORC_HBF_NPRESET)
        return NULL;
    orc_hbf8 *h = (orc_hbf8 *)calloc(1, sizeof(*h));
    h->preset = preset;
    for (int s = 0; s < 3; s++) {
        int set = 2 - s;
        h->st[s].m = orc_hbf_ntaps[preset][set];
        h->st[s].t = orc_hbf_taps[preset][set];
        h->st[s].hist = (float *)calloc((size_t)(4 * h->st[s].m - 2), sizeof(float));
    }
    return h;
}

static orc_hbf8 *hbf8_clone(const orc_hbf8 *src)
{
    orc_hbf8 *h = orc_hbf8_new(src->preset);
    for (int s = 0; s < 3; s++)
        memcpy(h->st[s].hist, src->st[s].hist, (size_t)(4 * h->st[s].m - 2) * sizeof(float));
    return h;
}

void orc_hbf8_free(orc_hbf8 *h)
{
    if (!h)
        return;
    for (int s = 0; s < 3; s++) {
        free(h->st[s].hist);
        free(h->work[s]);
    }
    free(h);
}

/* e = hist ++ x (2k new inputs) -> y (k outputs); hist <- last 4m-2 of e.
 * Per output the taps are accumulated in the same order as a scalar loop would (outermost first);
 * the loops are merely arranged tap-outer / output-inner over de-interleaved even/odd planes so
 * that the compiler vectorises across outputs. */
#define HBF_BLK 512
static void hbf2_block(hbf2 *f, float *e, size_t k, float *y)
{
    const int m = f->m;
    const int hl = 4 * m - 2;
    const float *t = f->t;
    float od[HBF_BLK + 2 * ORC_HBF_MAXTAPS + 8], ev[HBF_BLK], acc[HBF_BLK];
    for (size_t j0 = 0; j0 < k; j0 += HBF_BLK) {
        const size_t nb = k - j0 < HBF_BLK ? k - j0 : HBF_BLK;
        /* output j: centre e[2m + 2j]; odd taps e[2m + 2j +- (2q+1)], q < m
         * od[i] = e[2*j0 + 1 + 2i] covers i in [0, nb + 2m - 1) */
        const float *base = e + 2 * j0;
        for (size_t i = 0; i < nb + 2 * (size_t)m - 1; i++)
            od[i] = base[1 + 2 * i];
        for (size_t j = 0; j < nb; j++) {
            ev[j] = base[2 * m + 2 * j];
            acc[j] = 0.0f;
        }
        for (int i = 0; i < m; i++) {
            /* tap i pairs od[j + i] (below the centre) with od[j + 2m - 1 - i] (above) */
            const float ti = t[i];
            const float *lo = od + i, *hi = od + (2 * m - 1 - i);
            for (size_t j = 0; j < nb; j++)
                acc[j] += (lo[j] + hi[j]) * ti;
        }
        for (size_t j = 0; j < nb; j++)
            y[j0 + j] = ev[j] + acc[j];
    }
    memmove(f->hist, e + 2 * k, (size_t)hl * sizeof(float));
}

void orc_hbf8_block(orc_hbf8 *h, const float *x, size_t n_chunks, float *y)
{
    if (n_chunks == 0)
        return;
    if (n_chunks > h->work_cap) {
        for (int s = 0; s < 3; s++) {
            free(h->work[s]);
            size_t n_in = (n_chunks * 8) >> s;
            h->work[s] = (float *)malloc((n_in + (size_t)(4 * h->st[s].m - 2)) * sizeof(float));
        }
        h->work_cap = n_chunks;
    }
    /* work[s] = hist_s ++ input_s; each stage writes its output directly behind the history slot
     * of the next stage's work buffer */
    memcpy(h->work[0] + (4 * h->st[0].m - 2), x, n_chunks * 8 * sizeof(float));
    for (int s = 0; s < 3; s++) {
        size_t n_in = (n_chunks * 8) >> s;
        int hl = 4 * h->st[s].m - 2;
        memcpy(h->work[s], h->st[s].hist, (size_t)hl * sizeof(float));
        float *out = (s == 2) ? y : (h->work[s + 1] + (4 * h->st[s + 1].m - 2));
        hbf2_block(&h->st[s], h->work[s], n_in / 2, out);
    }
}

int orc_hbf_response_length(int preset)
{
    return orc_hbf_drain[preset];
}

float orc_hbf_passband(void)
{
    return 0.4f;
}

/* ------------------------------------------------------------------------------------------
 * rustfft restatement: forward, unnormalised, X[k] = sum_n x[n] e^{-2 pi i nk/N}.
 * Four-step FFT over split re/im planes; the sub-transforms are Stockham autosort radix-4 (+ one
 * radix-2) passes over interleaved batches, so every inner loop is long and contiguous and gcc
 * vectorises it (the reference's rustfft picks an AVX kernel at run time; this keeps the CPU
 * baseline within a small factor of it).  Twiddles from f64.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int n, n1, n2;
    float *w1r, *w1i; /* W_n1^k */
    float *w2r, *w2i; /* W_n2^k */
    float *tr, *ti;   /* W_n^(k1*c), [n1][n2] */
} fft_plan;

/* Plans hold read-only tables only (built once under a mutex); every thread owns its work planes, so
 * cascades may run on different threads concurrently (bench.py --impl reference, one channel per thread). */
static fft_plan g_plans[32];
static pthread_mutex_t g_plan_lock = PTHREAD_MUTEX_INITIALIZER;

typedef struct {
    int cap;
    float *ar, *ai, *br, *bi; /* work planes of fft_soa, cap floats each */
    int ccap;
    float *cre, *cim;         /* split planes of orc_fft_forward */
} fft_scratch;
static __thread fft_scratch t_scratch;

/* Planes are placed at distinct offsets modulo 4 KiB: the radix-4 passes stream 16 arrays whose
 * mutual distances are powers of two, which would otherwise all map to the same L1 sets. */
static float *alloc_plane(size_t n, size_t skew_bytes)
{
    void *p = NULL;
    if (posix_memalign(&p, 4096, n * sizeof(float) + 4096) != 0)
        abort();
    return (float *)((char *)p + skew_bytes);
}

static void make_roots(int n, float **wr, float **wi)
{
    *wr = (float *)malloc(sizeof(float) * (size_t)n);
    *wi = (float *)malloc(sizeof(float) * (size_t)n);
    for (int k = 0; k < n; k++) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        (*wr)[k] = (float)cos(a);
        (*wi)[k] = (float)sin(a);
    }
}

static fft_plan *fft_get_plan(int n)
{
    int l = 0;
    while ((1 << l) < n)
        l++;
    assert((1 << l) == n && l < 32 && l >= 2);
    fft_plan *p = &g_plans[l];
    if (__atomic_load_n(&p->n, __ATOMIC_ACQUIRE) == n)
        return p;
    pthread_mutex_lock(&g_plan_lock);
    if (p->n != n) {
        p->n1 = 1 << ((l + 1) / 2);
        p->n2 = n / p->n1;
        make_roots(p->n1, &p->w1r, &p->w1i);
        make_roots(p->n2, &p->w2r, &p->w2i);
        p->tr = (float *)malloc(sizeof(float) * (size_t)n);
        p->ti = (float *)malloc(sizeof(float) * (size_t)n);
        for (int k1 = 0; k1 < p->n1; k1++)
            for (int c = 0; c < p->n2; c++) {
                double a = -2.0 * M_PI * (double)k1 * (double)c / (double)n;
                p->tr[k1 * p->n2 + c] = (float)cos(a);
                p->ti[k1 * p->n2 + c] = (float)sin(a);
            }
        __atomic_store_n(&p->n, n, __ATOMIC_RELEASE); /* published last: tables are complete */
    }
    pthread_mutex_unlock(&g_plan_lock);
    return p;
}

static fft_scratch *fft_get_scratch(int n)
{
    fft_scratch *s = &t_scratch;
    if (n > s->cap) { /* planes are never freed (they are skewed pointers); grow-only, per thread */
        s->ar = alloc_plane((size_t)n, 2048);
        s->ai = alloc_plane((size_t)n, 3072);
        s->br = alloc_plane((size_t)n, 512);
        s->bi = alloc_plane((size_t)n, 1536);
        s->cap = n;
    }
    return s;
}

static inline void r4_inner(const float *restrict ar, const float *restrict ai, const float *restrict br,
                            const float *restrict bi, const float *restrict cr, const float *restrict ci,
                            const float *restrict dr, const float *restrict di, float *restrict y0r,
                            float *restrict y0i, float *restrict y1r, float *restrict y1i, float *restrict y2r,
                            float *restrict y2i, float *restrict y3r, float *restrict y3i, size_t n, float w1r,
                            float w1i, float w2r, float w2i, float w3r, float w3i)
{
    for (size_t q = 0; q < n; q++) {
        float apcr = ar[q] + cr[q], apci = ai[q] + ci[q];
        float amcr = ar[q] - cr[q], amci = ai[q] - ci[q];
        float bpdr = br[q] + dr[q], bpdi = bi[q] + di[q];
        float jr = bi[q] - di[q], ji = -(br[q] - dr[q]); /* -i (b - d) */
        y0r[q] = apcr + bpdr;
        y0i[q] = apci + bpdi;
        float t1r = amcr + jr, t1i = amci + ji;
        y1r[q] = t1r * w1r - t1i * w1i;
        y1i[q] = t1r * w1i + t1i * w1r;
        float t2r = apcr - bpdr, t2i = apci - bpdi;
        y2r[q] = t2r * w2r - t2i * w2i;
        y2i[q] = t2r * w2i + t2i * w2r;
        float t3r = amcr - jr, t3i = amci - ji;
        y3r[q] = t3r * w3r - t3i * w3i;
        y3i[q] = t3r * w3i + t3i * w3r;
    }
}

static inline void r2_inner(const float *restrict ar, const float *restrict ai, const float *restrict br,
                            const float *restrict bi, float *restrict y0r, float *restrict y0i,
                            float *restrict y1r, float *restrict y1i, size_t n, float w1r, float w1i)
{
    for (size_t q = 0; q < n; q++) {
        y0r[q] = ar[q] + br[q];
        y0i[q] = ai[q] + bi[q];
        float tr = ar[q] - br[q], ti = ai[q] - bi[q];
        y1r[q] = tr * w1r - ti * w1i;
        y1i[q] = tr * w1i + ti * w1r;
    }
}

/* B interleaved length-n transforms (element idx of transform b at [idx*B + b]), Stockham autosort,
 * radix 4 (+ one radix 2).  Result ends up in (xr, xi) or (yr, yi); returns 1 if in y. */
static int fft_batch(float *xr, float *xi, float *yr, float *yi, int n, int B, const float *wr, const float *wi)
{
    int len = n, s = 1, flip = 0;
    while (len > 1) {
        const size_t sb = (size_t)s * B;
        if (len % 4 == 0) {
            const int m = len / 4, tws = n / len;
            for (int p = 0; p < m; p++) {
                float *y0r = yr + sb * (4 * p), *y0i = yi + sb * (4 * p);
                r4_inner(xr + sb * p, xi + sb * p, xr + sb * (p + m), xi + sb * (p + m), xr + sb * (p + 2 * m),
                         xi + sb * (p + 2 * m), xr + sb * (p + 3 * m), xi + sb * (p + 3 * m), y0r, y0i, y0r + sb,
                         y0i + sb, y0r + 2 * sb, y0i + 2 * sb, y0r + 3 * sb, y0i + 3 * sb, sb, wr[p * tws],
                         wi[p * tws], wr[2 * p * tws], wi[2 * p * tws], wr[3 * p * tws], wi[3 * p * tws]);
            }
            len = m;
            s *= 4;
        } else {
            const int m = len / 2, tws = n / len;
            for (int p = 0; p < m; p++) {
                float *y0r = yr + sb * (2 * p), *y0i = yi + sb * (2 * p);
                r2_inner(xr + sb * p, xi + sb * p, xr + sb * (p + m), xi + sb * (p + m), y0r, y0i, y0r + sb,
                         y0i + sb, sb, wr[p * tws], wi[p * tws]);
            }
            len = m;
            s *= 2;
        }
        float *t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
        flip ^= 1;
    }
    return flip;
}

/* Four-step FFT on split planes: (re, im) of length n, in place.  n = n1*n2, input index r*n2 + c:
 * n2 interleaved column transforms of length n1, twiddle W_n^(k1 c), transpose, n1 interleaved
 * transforms of length n2; X[k1 + n1 k2] lands at k2*n1 + k1, i.e. in natural order. */
static void fft_soa(float *re, float *im, int n)
{
    fft_plan *pl = fft_get_plan(n);
    fft_scratch *sc = fft_get_scratch(n);
    const int n1 = pl->n1, n2 = pl->n2;
    float *xr = re, *xi = im, *yr = sc->ar, *yi = sc->ai;
    if (fft_batch(xr, xi, yr, yi, n1, n2, pl->w1r, pl->w1i)) {
        xr = sc->ar;
        xi = sc->ai;
    }
    /* twiddle (contiguous), then blocked transpose [n1][n2] -> [n2][n1] */
    float *tr = sc->br, *ti = sc->bi;
    {
        const float *restrict wr = pl->tr, *restrict wi = pl->ti;
        float *restrict ar = xr, *restrict ai = xi;
        for (int i = 0; i < n; i++) {
            float vr = ar[i] * wr[i] - ai[i] * wi[i];
            float vi = ar[i] * wi[i] + ai[i] * wr[i];
            ar[i] = vr;
            ai[i] = vi;
        }
    }
    enum { TB = 16 };
    for (int r0 = 0; r0 < n1; r0 += TB)
        for (int c0 = 0; c0 < n2; c0 += TB) {
            const int rb = n1 - r0 < TB ? n1 - r0 : TB, cb = n2 - c0 < TB ? n2 - c0 : TB;
            for (int c = 0; c < cb; c++) {
                float *restrict dr = tr + (size_t)(c0 + c) * n1 + r0, *restrict di = ti + (size_t)(c0 + c) * n1 + r0;
                const float *restrict sr = xr + (size_t)r0 * n2 + c0 + c, *restrict si = xi + (size_t)r0 * n2 + c0 + c;
                for (int r = 0; r < rb; r++) {
                    dr[r] = sr[(size_t)r * n2];
                    di[r] = si[(size_t)r * n2];
                }
            }
        }
    float *zr = (xr == re) ? sc->ar : re, *zi = (xr == re) ? sc->ai : im;
    if (fft_batch(tr, ti, zr, zi, n2, n1, pl->w2r, pl->w2i)) {
        if (zr != re) {
            memcpy(re, zr, sizeof(float) * (size_t)n);
            memcpy(im, zi, sizeof(float) * (size_t)n);
        }
    } else {
        memcpy(re, tr, sizeof(float) * (size_t)n);
        memcpy(im, ti, sizeof(float) * (size_t)n);
    }
}

void orc_fft_forward(float *c, int n)
{
    if (n <= 1)
        return;
    if (n == 2) {
        float ar = c[0], ai = c[1], br = c[2], bi = c[3];
        c[0] = ar + br; c[1] = ai + bi; c[2] = ar - br; c[3] = ai - bi;
        return;
    }
    fft_scratch *sc = &t_scratch;
    if (n > sc->ccap) {
        sc->cre = alloc_plane((size_t)n, 0);
        sc->cim = alloc_plane((size_t)n, 1024);
        sc->ccap = n;
    }
    float *re = sc->cre, *im = sc->cim;
    for (int i = 0; i < n; i++) {
        re[i] = c[2 * i];
        im[i] = c[2 * i + 1];
    }
    fft_soa(re, im, n);
    for (int i = 0; i < n; i++) {
        c[2 * i] = re[i];
        c[2 * i + 1] = im[i];
    }
}

/* ------------------------------------------------------------------------------------------
 * Window (psd.rs:12-56)
 * ------------------------------------------------------------------------------------------ */
void orc_window(int n, int window, float *win, float *power, float *nenbw, size_t *overlap)
{
    if (window == ORC_WINDOW_RECT) { /* psd.rs:24-32 */
        for (int i = 0; i < n; i++)
            win[i] = 1.0f;
        *power = 1.0f;
        *nenbw = 1.0f;
        *overlap = 0;
    } else { /* psd.rs:42-55: (PI / N as f32 * i as f32).sin().powi(2) */
        const float pi = 3.14159265358979323846f;
        float df = pi / (float)n;
        for (int i = 0; i < n; i++) {
            float s = sinf(df * (float)i);
            win[i] = s * s;
        }
        *power = 0.25f;
        *nenbw = 1.5f;
        *overlap = (size_t)n / 2;
    }
}

/* Detrend::apply (psd.rs:75-113), real part only (the imaginary part is 0) */
static int detrend_real(int detrend, const float *x, const float *win, int n, float *re)
{
    switch (detrend) {
    case ORC_DETREND_NONE: /* psd.rs:81-86 */
        for (int i = 0; i < n; i++)
            re[i] = x[i] * win[i];
        return 0;
    case ORC_DETREND_MIDPOINT: { /* psd.rs:87-93 */
        float offset = x[n / 2];
        for (int i = 0; i < n; i++)
            re[i] = (x[i] - offset) * win[i];
        return 0;
    }
    case ORC_DETREND_SPAN: { /* psd.rs:94-102: offset accumulated sequentially */
        float offset = x[0];
        float slope = (x[n - 1] - x[0]) / (float)(n - 1);
        for (int i = 0; i < n; i++) {
            re[i] = (x[i] - offset) * win[i];
            offset += slope;
        }
        return 0;
    }
    case ORC_DETREND_MEAN: { /* psd.rs:103-109: sequential f32 sum */
        float sum = 0.0f;
        for (int i = 0; i < n; i++)
            sum += x[i];
        float offset = sum / (float)n;
        for (int i = 0; i < n; i++)
            re[i] = (x[i] - offset) * win[i];
        return 0;
    }
    default: /* psd.rs:110 unimplemented!() */
        return -1;
    }
}

int orc_detrend_apply(int detrend, const float *x, const float *win, int n, float *c)
{
    float *re = (float *)malloc(sizeof(float) * (size_t)n);
    int r = detrend_real(detrend, x, win, n, re);
    if (r == 0)
        for (int i = 0; i < n; i++) {
            c[2 * i] = re[i];
            c[2 * i + 1] = 0.0f;
        }
    free(re);
    return r;
}

/* ------------------------------------------------------------------------------------------
 * Psd<N> (psd.rs:123-288)
 * ------------------------------------------------------------------------------------------ */
struct orc_stage {
    int n;
    int window;
    int hbf_preset;
    orc_hbf8 *hbf;
    float *buf;
    size_t idx;
    float *spectrum;
    uint32_t count;
    size_t drain;
    float *win;
    float power, nenbw;
    size_t overlap;
    int detrend;
    uint32_t avg;
    float *c_re, *c_im; /* scratch planes, N floats each (skewed, see alloc_plane) */
    void *c_base;
};

orc_stage *orc_stage_new(int n, int window, int hbf_preset)
{
    /* psd.rs:137-152; as_chunks::<8> remainder assert psd.rs:246-247 */
    if (n < 16 || (n & (n - 1)) != 0 || hbf_preset < 0 || hbf_preset >= ORC_HBF_NPRESET)
        return NULL;
    orc_stage *s = (orc_stage *)calloc(1, sizeof(*s));
    s->n = n;
    s->window = window;
    s->hbf_preset = hbf_preset;
    s->hbf = orc_hbf8_new(hbf_preset);
    s->buf = (float *)calloc((size_t)n, sizeof(float));
    s->spectrum = (float *)calloc((size_t)n, sizeof(float));
    s->win = (float *)calloc((size_t)n, sizeof(float));
    if (posix_memalign(&s->c_base, 4096, 2 * (size_t)n * sizeof(float) + 8192) != 0)
        abort();
    s->c_re = (float *)s->c_base;
    s->c_im = (float *)((char *)s->c_base + (((size_t)n * sizeof(float) + 4095) & ~(size_t)4095) + 1024);
    orc_window(n, window, s->win, &s->power, &s->nenbw, &s->overlap);
    s->count = 0;
    s->idx = 0;
    s->drain = (size_t)orc_hbf_response_length(hbf_preset); /* psd.rs:149 */
    s->detrend = ORC_DETREND_NONE;
    s->avg = UINT32_MAX; /* psd.rs:150 */
    return s;
}

orc_stage *orc_stage_clone(const orc_stage *o)
{
    orc_stage *s = orc_stage_new(o->n, o->window, o->hbf_preset);
    orc_hbf8_free(s->hbf);
    s->hbf = hbf8_clone(o->hbf);
    memcpy(s->buf, o->buf, sizeof(float) * (size_t)o->n);
    memcpy(s->spectrum, o->spectrum, sizeof(float) * (size_t)o->n);
    s->idx = o->idx;
    s->count = o->count;
    s->drain = o->drain;
    s->detrend = o->detrend;
    s->avg = o->avg;
    return s;
}

void orc_stage_free(orc_stage *s)
{
    if (!s)
        return;
    orc_hbf8_free(s->hbf);
    free(s->buf);
    free(s->spectrum);
    free(s->win);
    free(s->c_base);
    free(s);
}

void orc_stage_set_avg(orc_stage *s, uint32_t avg) { s->avg = avg; }       /* psd.rs:154-156 */
void orc_stage_set_detrend(orc_stage *s, int d) { s->detrend = d; }        /* psd.rs:158-160 */

size_t orc_stage_process(orc_stage *s, const float *x, size_t nx, float *y)
{
    const size_t N = (size_t)s->n;
    size_t n = 0;
    while (nx > 0) { /* psd.rs:199 */
        /* load, psd.rs:201-208 */
        size_t take = nx < N - s->idx ? nx : N - s->idx;
        memcpy(s->buf + s->idx, x, take * sizeof(float));
        x += take;
        nx -= take;
        s->idx += take;
        if (s->idx < N)
            break;

        /* detrend and window, psd.rs:211 (re plane; im = 0); fft in place, psd.rs:213 */
        if (detrend_real(s->detrend, s->buf, s->win, s->n, s->c_re) != 0)
            abort(); /* Detrend::Linear panics in the reference */
        memset(s->c_im, 0, N * sizeof(float));
        fft_soa(s->c_re, s->c_im, s->n);

        int is_first = s->count == 0; /* psd.rs:215 */

        /* normalize and keep for EWMA, psd.rs:218-225 */
        float g;
        if (s->count > s->avg) {
            g = (float)s->avg / (float)s->count;
            s->count = s->avg;
        } else {
            g = 1.0f;
        }
        s->count += 1;

        /* power accumulate, psd.rs:228-233 (norm_sqr = re*re + im*im, unfused) */
        for (size_t k = 0; k <= N / 2; k++) {
            float re = s->c_re[k], im = s->c_im[k];
            s->spectrum[k] = g * s->spectrum[k] + (re * re + im * im);
        }

        size_t start; /* psd.rs:235-243 */
        if (is_first) {
            start = 0;
        } else {
            memmove(s->buf, s->buf + (N - s->overlap), s->overlap * sizeof(float));
            start = s->overlap;
        }

        /* decimate, psd.rs:246-253 */
        assert((N - start) % 8 == 0);
        size_t chunks = (N - start) / 8;
        orc_hbf8_block(s->hbf, s->buf + start, chunks, y + n);
        /* drain, psd.rs:255-260 */
        size_t skip = s->drain < chunks ? s->drain : chunks;
        if (skip > 0) {
            s->drain -= skip;
            memmove(y + n, y + n + skip, (chunks - skip) * sizeof(float));
        }
        n += chunks - skip;

        if (is_first) /* psd.rs:262-265 */
            memmove(s->buf, s->buf + (N - s->overlap), s->overlap * sizeof(float));
        s->idx = s->overlap; /* psd.rs:266 */
    }
    return n;
}

const float *orc_stage_spectrum(const orc_stage *s) { return s->spectrum; } /* psd.rs:271 */
uint32_t orc_stage_count(const orc_stage *s) { return s->count; }           /* psd.rs:275 */

float orc_stage_gain(const orc_stage *s)
{
    /* psd.rs:279-283: (N as u32 / 2 * count) as f32 * nenbw * power.
     * The reference multiplies in u32 (overflows for count > 2^32/(N/2)); the oracle
     * multiplies in u64 -- identical whenever the reference does not overflow. */
    uint64_t nc = (uint64_t)((uint32_t)s->n / 2) * (uint64_t)s->count;
    return (float)nc * s->nenbw * s->power;
}

size_t orc_stage_buf(const orc_stage *s, float *out)
{
    if (out)
        memcpy(out, s->buf, s->idx * sizeof(float));
    return s->idx;
}

/* ------------------------------------------------------------------------------------------
 * PsdCascade<N> (psd.rs:399-544)
 * ------------------------------------------------------------------------------------------ */
#define ORC_MAX_STAGES 24
struct orc_cascade {
    int n;
    int hbf_preset;
    orc_stage *stages[ORC_MAX_STAGES];
    size_t n_stages;
    int detrend;
    uint32_t avg_limit, avg_count;
    float *a0, *a1;
};

orc_cascade *orc_cascade_new(int n, int hbf_preset)
{
    orc_stage *probe = orc_stage_new(n, ORC_WINDOW_HANN, hbf_preset);
    if (!probe)
        return NULL;
    orc_stage_free(probe);
    orc_cascade *c = (orc_cascade *)calloc(1, sizeof(*c));
    c->n = n;
    c->hbf_preset = hbf_preset;
    c->detrend = ORC_DETREND_NONE;              /* psd.rs:418 */
    c->avg_limit = UINT32_MAX;                  /* psd.rs:369-376 */
    c->avg_count = UINT32_MAX;
    /* The reference ping-pongs two [f32; N] arrays (psd.rs:457).  That is one decimated item short
     * in a corner it never meets with its own sources: when a stage's first-ever segment completes
     * inside an 8N chunk that also completes 15 more (possible only if an earlier short call left
     * more than N/2 items pending), the chunk yields N/8 - drain + 15 N/16 > N items and the slice
     * `y[n..][..xb.len()]` (psd.rs:253) panics.  The restatement continues the arithmetic instead
     * (buffers of 2N), which is also what the device library does; tests/test_gpu_psd.py covers it. */
    c->a0 = (float *)calloc(2 * (size_t)n, sizeof(float));
    c->a1 = (float *)calloc(2 * (size_t)n, sizeof(float));
    return c;
}

orc_cascade *orc_cascade_clone(const orc_cascade *o)
{
    orc_cascade *c = orc_cascade_new(o->n, o->hbf_preset);
    c->detrend = o->detrend;
    c->avg_limit = o->avg_limit;
    c->avg_count = o->avg_count;
    c->n_stages = o->n_stages;
    for (size_t i = 0; i < o->n_stages; i++)
        c->stages[i] = orc_stage_clone(o->stages[i]);
    return c;
}

void orc_cascade_free(orc_cascade *c)
{
    if (!c)
        return;
    for (size_t i = 0; i < c->n_stages; i++)
        orc_stage_free(c->stages[i]);
    free(c->a0);
    free(c->a1);
    free(c);
}

float orc_cascade_rbw(const orc_cascade *c)
{
    /* psd.rs:427-429 */
    return (float)(1 << ORC_DEPTH) / ((float)c->n * orc_hbf_passband());
}

static uint32_t stage_avg(const orc_cascade *c, size_t i)
{
    /* (avg.count >> (DEPTH * i)).min(avg.limit), psd.rs:434,449.  A shift >= 32 panics in a debug
     * build of the reference and is masked in release; stages that deep are unreachable. */
    unsigned sh = (unsigned)(ORC_DEPTH * i);
    uint32_t v = sh >= 32 ? 0u : (c->avg_count >> sh);
    return v < c->avg_limit ? v : c->avg_limit;
}

void orc_cascade_set_avg(orc_cascade *c, uint32_t limit, uint32_t count)
{
    c->avg_limit = limit; /* psd.rs:431-436 */
    c->avg_count = count;
    for (size_t i = 0; i < c->n_stages; i++)
        orc_stage_set_avg(c->stages[i], stage_avg(c, i));
}

void orc_cascade_set_detrend(orc_cascade *c, int d)
{
    c->detrend = d; /* psd.rs:438-443 */
    for (size_t i = 0; i < c->n_stages; i++)
        orc_stage_set_detrend(c->stages[i], d);
}

static orc_stage *get_or_add(orc_cascade *c, size_t i)
{
    while (i >= c->n_stages) { /* psd.rs:445-453 */
        assert(c->n_stages < ORC_MAX_STAGES);
        orc_stage *s = orc_stage_new(c->n, ORC_WINDOW_HANN, c->hbf_preset);
        orc_stage_set_detrend(s, c->detrend);
        orc_stage_set_avg(s, stage_avg(c, c->n_stages));
        c->stages[c->n_stages++] = s;
    }
    return c->stages[i];
}

void orc_cascade_process(orc_cascade *c, const float *x, size_t n)
{
    /* psd.rs:456-468 */
    const size_t chunk = (size_t)c->n << ORC_DEPTH;
    float *y = c->a0, *z = c->a1;
    for (size_t off = 0; off < n; off += chunk) {
        const float *xp = x + off;
        size_t len = n - off < chunk ? n - off : chunk;
        size_t i = 0;
        while (len > 0) {
            size_t m = orc_stage_process(get_or_add(c, i), xp, len, y);
            float *t = z;
            z = y;
            y = t;
            xp = z;
            len = m;
            i++;
        }
    }
}

size_t orc_cascade_num_stages(const orc_cascade *c) { return c->n_stages; }
const orc_stage *orc_cascade_stage(const orc_cascade *c, size_t i) { return c->stages[i]; }

size_t orc_cascade_psd(const orc_cascade *c, int keep_overlap, uint32_t min_count,
                       int keep_transition_band, float *p, orc_break *b, size_t *nb)
{
    /* psd.rs:479-543 */
    const size_t N = (size_t)c->n;
    size_t plen = 0, blen = 0;
    uint64_t decimation = (uint64_t)1 << (ORC_DEPTH * c->n_stages);
    size_t end = 0;
    for (size_t r = c->n_stages; r-- > 0;) {
        const orc_stage *st = c->stages[r];
        decimation >>= ORC_DEPTH;
        size_t start = !keep_overlap ? (end + (1u << ORC_DEPTH) - 1) >> ORC_DEPTH : 0;
        end = (decimation > 1 && !keep_transition_band) ? 2 * N / 5 : N / 2 + 1;
        int include = st->count >= min_count;
        orc_break *bk = &b[blen++];
        memset(bk, 0, sizeof(*bk));
        bk->start = plen;
        bk->count = st->count;
        bk->include = (uint32_t)include;
        bk->avg = st->avg;
        bk->bins_start = start;
        bk->bins_end = end;
        bk->fft_size = N;
        bk->decimation = decimation;
        uint32_t cm1 = st->count > 0 ? st->count - 1 : 0; /* saturating_sub(1) */
        bk->processed = (uint64_t)N * st->count - (uint64_t)st->overlap * cm1;
        bk->pending = st->idx;
        if (include) {
            float g = 1.0f / (orc_stage_gain(st) * (float)decimation);
            for (size_t k = start; k < end; k++)
                p[plen++] = st->spectrum[k] * g;
        } else {
            end = start;
        }
    }
    *nb = blen;
    return plen;
}

size_t orc_break_frequencies(const orc_break *b, size_t nb, float *f)
{
    /* psd.rs:315-327, rbw psd.rs:334-336 */
    size_t n = 0;
    for (size_t i = 0; i < nb; i++) {
        if (!b[i].include)
            continue;
        float rbw = 1.0f / (float)(b[i].fft_size * b[i].decimation);
        for (uint64_t k = b[i].bins_start; k < b[i].bins_end; k++)
            f[n++] = (float)k * rbw;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Frame decode (de/frame.rs:25-60, de/data.rs) and loss (loss.rs)
 * ------------------------------------------------------------------------------------------ */
static inline int16_t rd_i16(const uint8_t *p) { return (int16_t)((uint16_t)p[0] | ((uint16_t)p[1] << 8)); }
static inline int32_t rd_i32(const uint8_t *p)
{
    return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}
static inline int64_t rd_i64(const uint8_t *p)
{
    uint64_t lo = (uint32_t)rd_i32(p), hi = (uint32_t)rd_i32(p + 4);
    return (int64_t)(lo | (hi << 32));
}
static inline float rd_f32(const uint8_t *p)
{
    int32_t v = rd_i32(p);
    float f;
    memcpy(&f, &v, 4);
    return f;
}

int orc_frame_decode(const uint8_t *buf, size_t len, orc_header *hdr, float *const *traces,
                     size_t *samples_per_trace, int *n_traces)
{
    if (len < 8)
        return ORC_ESHORT; /* frame.rs:50 slice index panics */
    /* Header::parse, frame.rs:25-38 */
    if (buf[0] != 0x7b || buf[1] != 0x05)
        return ORC_EHEADER;
    uint8_t format = buf[2];
    if (format < 1 || format > 4)
        return ORC_EFORMAT; /* mod.rs:12-17 */
    hdr->format = format;
    hdr->batches = buf[3];
    hdr->seq = (uint32_t)rd_i32(buf + 4);
    const uint8_t *d = buf + 8;
    size_t dl = len - 8;
    const size_t batches = hdr->batches;
    static const size_t bsize[5] = {0, 64, 56, 80, 24};
    if (dl % bsize[format] != 0)
        return ORC_ESIZE; /* bytemuck::try_cast_slice -> PodCastError, data.rs:23,91,149,173 */
    if (dl / bsize[format] != batches)
        return ORC_EBATCHES; /* assert_eq!, data.rs:24,93,150,174 */

    switch (format) {
    case 1: { /* AdcDac::traces, data.rs:28-82 */
        const float volt_per_lsb = 4.096f * 2.5f / 32768.0f; /* data.rs:31 */
        for (size_t b = 0; b < batches; b++) {
            const uint8_t *bp = d + 64 * b;
            for (int i = 0; i < 8; i++) {
                traces[0][8 * b + i] = (float)rd_i16(bp + 2 * i) * volt_per_lsb;
                traces[1][8 * b + i] = (float)rd_i16(bp + 16 + 2 * i) * volt_per_lsb;
                /* wrapping_add(i16::MIN): offset binary -> two's complement, data.rs:64,75 */
                traces[2][8 * b + i] =
                    (float)(int16_t)((uint16_t)rd_i16(bp + 32 + 2 * i) ^ 0x8000u) * volt_per_lsb;
                traces[3][8 * b + i] =
                    (float)(int16_t)((uint16_t)rd_i16(bp + 48 + 2 * i) ^ 0x8000u) * volt_per_lsb;
            }
        }
        *n_traces = 4;
        *samples_per_trace = 8 * batches;
        return ORC_OK;
    }
    case 2: { /* Fls::traces, data.rs:97-139; batch = [[i32;7];2] */
        const float i32max = (float)INT32_MAX; /* 2^31 after rounding */
        const float tau = 6.28318530717958647692f;
        for (size_t b = 0; b < batches; b++) {
            const uint8_t *bp = d + 56 * b;
            float re = (float)rd_i32(bp), im = (float)rd_i32(bp + 4);
            traces[0][b] = sqrtf(re * re + im * im) * (1.0f / i32max);
            traces[1][b] = (float)rd_i64(bp + 8) * (tau / 65536.0f);
            traces[2][b] = (float)rd_i32(bp + 28) / i32max;
            traces[3][b] = (float)rd_i32(bp + 32) / i32max;
        }
        *n_traces = 4;
        *samples_per_trace = batches;
        return ORC_OK;
    }
    case 3: { /* ThermostatEem::traces, data.rs:154-163; f32 words 0, 8, 13, 16 of 20 */
        static const int idx[4] = {0, 8, 13, 16};
        for (size_t b = 0; b < batches; b++)
            for (int t = 0; t < 4; t++)
                traces[t][b] = rd_f32(d + 80 * b + 4 * idx[t]);
        *n_traces = 4;
        *samples_per_trace = batches;
        return ORC_OK;
    }
    default: { /* Mpll::traces, data.rs:178-211 */
        const float tau = 6.28318530717958647692f;
        const float two32 = 4294967296.0f;
        const float k_phase = tau / two32;
        const float k_freq = 1.0f / 1.28e-3f / two32;
        const float k_amp = 10.24f / 10.0f * 2.0f * 2.0f / two32;
        for (size_t b = 0; b < batches; b++) {
            const uint8_t *bp = d + 24 * b;
            traces[0][b] = (float)rd_i32(bp + 16) * k_phase;
            traces[1][b] = (float)rd_i32(bp + 20) * k_freq;
            float re = (float)rd_i32(bp), im = (float)rd_i32(bp + 4);
            traces[2][b] = sqrtf(re * re + im * im) * k_amp;
        }
        *n_traces = 3;
        *samples_per_trace = batches;
        return ORC_OK;
    }
    }
}

void orc_loss_update(orc_loss *l, uint32_t seq, uint8_t batches)
{
    /* loss.rs:11-26 */
    l->received += batches;
    if (l->has_seq) {
        uint64_t missing = (uint32_t)(seq - l->seq);
        l->dropped += missing;
    }
    l->seq = seq + batches;
    l->has_seq = 1;
}

float orc_loss_ratio(const orc_loss *l)
{
    /* loss.rs:29-30 */
    return (float)l->dropped / (float)(l->received + l->dropped);
}

/* ------------------------------------------------------------------------------------------
 * Var::eval (var.rs:26-45)
 * ------------------------------------------------------------------------------------------ */
static float powi_f32(float x, int e)
{
    /* f32::powi: repeated multiplication (llvm.powi), negative exponent -> reciprocal */
    int n = e < 0 ? -e : e;
    float r = 1.0f, b = x;
    while (n) {
        if (n & 1)
            r *= b;
        b *= b;
        n >>= 1;
    }
    return e < 0 ? 1.0f / r : r;
}

float orc_var_eval(int x_exp, int sinx_exp, float clip, size_t dc_cut, const float *phase_psd,
                   const float *frequencies, size_t n, float tau)
{
    const float pi = 3.14159265358979323846f;
    float accu = 0.0f, a0 = 0.0f, f0 = 0.0f;
    for (size_t i = dc_cut; i < n; i++) {
        float sp = phase_psd[i], f = frequencies[i];
        if (!(f <= clip / tau))
            break; /* take_while */
        float sy = sp * f * f;
        float pft = pi * (f * tau);
        float hahd = powi_f32(sinf(pft), sinx_exp) * powi_f32(pft, x_exp);
        float a = sy * hahd;
        accu = accu + (a + a0) * (f - f0);
        a0 = a;
        f0 = f;
    }
    return accu;
}

/* Trace::plot + Trapezoidal (bin/psd.rs:96-157).  Returns the number of plot points written to xy. */
size_t orc_trace_plot(float fs, float integral_start, float integral_end, int integrate, const float *psd,
                      const float *frequencies, size_t n, float *integral, double *xy)
{
    const float logfs = log10f(fs);                /* bin/psd.rs:126 */
    float x0 = 0.0f, y0 = 0.0f, acc = 0.0f;        /* Trapezoidal::default(), bin/psd.rs:96-101 */
    float pi = 0.0f;
    size_t np = 0;
    for (size_t i = 0; i < n; i++) {
        float p = psd[i], f = frequencies[i];
        float di = (p + y0) * 0.5f * (f - x0);     /* push, bin/psd.rs:104-110 */
        x0 = f;
        y0 = p;
        acc += di;
        float ff = fs * f;
        if (integral_start <= ff && ff <= integral_end) /* RangeInclusive::contains, bin/psd.rs:136 */
            pi += di;
        if (fpclassify(f) == FP_NORMAL) {          /* f.is_normal(), bin/psd.rs:140 */
            if (xy) {
                xy[2 * np] = (double)(log10f(f) + logfs);
                xy[2 * np + 1] = (double)(integrate ? sqrtf(acc) : 10.0f * (log10f(p) - logfs));
            }
            np++;
        }
    }
    if (integral)
        *integral = sqrtf(pi);                     /* bin/psd.rs:155 */
    return np;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic sources (source.rs:66-73, 104-134)
 * ------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    /* Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11), Philox-4x32, 10 rounds */
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

struct orc_source {
    int kind;
    uint64_t pos;
    /* noise */
    uint32_t key[2];
    int diff, order;
    float state[ORC_SOURCE_MAX_ORDER];
    /* dsm */
    uint32_t ftw, a[3];
    int c2z, c3z, c3zz;
};

orc_source *orc_source_new(int kind, int64_t param, uint64_t seed)
{
    orc_source *s = calloc(1, sizeof(*s));
    if (!s)
        return NULL;
    s->kind = kind;
    if (kind == ORC_SOURCE_NOISE) {
        int64_t o = param < 0 ? -param : param;
        if (o > ORC_SOURCE_MAX_ORDER) {
            free(s);
            return NULL;
        }
        s->diff = param > 0; /* source.rs:70 */
        s->order = (int)o;   /* source.rs:71 */
        s->key[0] = (uint32_t)seed;
        s->key[1] = (uint32_t)(seed >> 32);
    } else if (kind == ORC_SOURCE_DSM) {
        s->ftw = (uint32_t)param;
    } else {
        free(s);
        return NULL;
    }
    return s;
}

void orc_source_free(orc_source *s)
{
    free(s);
}

static float uniform_open01(uint32_t r)
{
    /* 23 random mantissa bits in [1, 2), shifted to (0, 1): never 0, never 1 (rand's Open01) */
    union {
        uint32_t u;
        float f;
    } v;
    v.u = 0x3f800000u | (r >> 9);
    return v.f - (1.0f - 5.9604645e-8f);
}

uint32_t orc_dsm_input(uint64_t i, uint32_t ftw)
{
    /* source.rs:119-126: x starts at 1 and advances by ftw (wrapping) after each sample.  The sine
     * is evaluated in double and rounded to f32 so that host and device agree bit for bit. */
    const float M = 4294967296.0f;
    uint32_t x = 1u + (uint32_t)i * ftw;
    float arg = (float)x * (6.28318530717958647692f / M);
    float sn = (float)sin((double)arg);
    float v = (sn * 0.4999f + 0.5f) * M;
    return (uint32_t)v;
}

void orc_source_get(orc_source *s, float *out, size_t n)
{
    if (s->kind == ORC_SOURCE_NOISE) {
        const float scale = sqrtf(12.0f);
        for (size_t j = 0; j < n; j++, s->pos++) {
            uint32_t ctr[4] = {(uint32_t)(s->pos >> 2), (uint32_t)(s->pos >> 34), 0, 0}, r[4];
            orc_philox4x32_10(ctr, s->key, r);
            float x = (uniform_open01(r[s->pos & 3]) - 0.5f) * scale; /* source.rs:109 */
            for (int k = 0; k < s->order; k++) {                      /* source.rs:110-114 */
                float st = s->state[k];
                if (s->diff) {
                    s->state[k] = x;
                    x = x - st;
                } else {
                    s->state[k] = x + st;
                    x = st;
                }
            }
            out[j] = x;
        }
    } else {
        for (size_t j = 0; j < n; j++, s->pos++) {
            uint32_t x = orc_dsm_input(s->pos, s->ftw), t;
            int c1, c2, c3;
            t = s->a[0] + x;
            c1 = t < x;
            s->a[0] = t;
            t = s->a[1] + s->a[0];
            c2 = t < s->a[0];
            s->a[1] = t;
            t = s->a[2] + s->a[1];
            c3 = t < s->a[1];
            s->a[2] = t;
            int y = c1 + (c2 - s->c2z) + (c3 - 2 * s->c3z + s->c3zz);
            s->c2z = c2;
            s->c3zz = s->c3z;
            s->c3z = c3;
            out[j] = (float)y - 0.5f; /* source.rs:127 */
        }
    }
}
