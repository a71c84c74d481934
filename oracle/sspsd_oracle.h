/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's cascaded-PSD hot path (quartiq/stabilizer-stream,
 * src/psd.rs, src/de/{frame,data}.rs, src/loss.rs, src/var.rs).  It exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg can check and time
 * the CUDA library against the reference semantics.  Nothing under stabilizer_stream_b200/ may
 * include, link or call it: the product has no CPU path.
 *
 * PARITY STATUS
 *   - src/psd.rs, src/de, src/loss.rs, src/var.rs semantics: restated line by line (citations at
 *     every function) and pinned by the reference's own tests (tests/test_oracle_*.py):
 *     psd.rs:599-644 (white-noise flatness + output-length identity), psd.rs:562-597 (Hann
 *     known-answer vector, generalised), psd.rs:601 (HBF_PASSBAND), var.rs:52-60 (Var KAT).
 *   - rustfft 6.4.1 (Cargo.lock:2362-2363, not on disk): forward unnormalised complex DFT,
 *     e^{-2 pi i nk/N}; restated with an own Stockham FFT (pinned by the DFT definition, checked
 *     against numpy in the tests).
 *   - idsp 0.20.0 `hbf` (Cargo.lock:1365-1366, not on disk): PARITY UNPINNED.  Taps are restated
 *     from idsp's published remez recipe (tools/gen_hbf_taps.py reproduces the recalled 98 dB
 *     literals to 1e-8); which tap family HBF_DEC_CASCADE uses and the value of
 *     hbf_dec_response_length(3) cannot be verified offline.  Both families are selectable.
 *   - The reference cannot be compiled here (no cargo/rustc): there is no oracle/_ref.
 */
#ifndef SSPSD_ORACLE_H
#define SSPSD_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_WINDOW_RECT = 0, ORC_WINDOW_HANN = 1 };
enum { ORC_DETREND_NONE = 0, ORC_DETREND_MIDPOINT = 1, ORC_DETREND_SPAN = 2, ORC_DETREND_MEAN = 3,
       ORC_DETREND_LINEAR = 4 };
enum { ORC_HBF_98 = 0, ORC_HBF_140 = 1 };
#define ORC_DEPTH 3 /* psd.rs:117 */

/* ---- half-band decimate-by-8 (idsp::hbf::HbfDec8 restatement) ---- */
typedef struct orc_hbf8 orc_hbf8;
orc_hbf8 *orc_hbf8_new(int preset);
void orc_hbf8_free(orc_hbf8 *h);
/* x: n_chunks*8 inputs -> y: n_chunks outputs; state persists across calls (psd.rs:246-253) */
void orc_hbf8_block(orc_hbf8 *h, const float *x, size_t n_chunks, float *y);
int orc_hbf_response_length(int preset); /* hbf_dec_response_length(3), psd.rs:149 */
float orc_hbf_passband(void);            /* HBF_PASSBAND, psd.rs:601 */

/* ---- forward complex FFT (rustfft restatement), interleaved re/im, in place ---- */
void orc_fft_forward(float *c, int n);

/* ---- Psd<N>: one stage (psd.rs:123-288) ---- */
typedef struct orc_stage orc_stage;
orc_stage *orc_stage_new(int n, int window, int hbf_preset);
orc_stage *orc_stage_clone(const orc_stage *s);
void orc_stage_free(orc_stage *s);
void orc_stage_set_avg(orc_stage *s, uint32_t avg);
void orc_stage_set_detrend(orc_stage *s, int detrend);
/* PsdStage::process(x, y) -> number of items written to y (psd.rs:196-269) */
size_t orc_stage_process(orc_stage *s, const float *x, size_t nx, float *y);
const float *orc_stage_spectrum(const orc_stage *s); /* N/2+1 values (psd.rs:271-273) */
uint32_t orc_stage_count(const orc_stage *s);
float orc_stage_gain(const orc_stage *s);
size_t orc_stage_buf(const orc_stage *s, float *out); /* pending items incl. overlap (psd.rs:285) */
void orc_window(int n, int window, float *win, float *power, float *nenbw, size_t *overlap);
/* Detrend::apply (psd.rs:75-113): writes n interleaved complex values */
int orc_detrend_apply(int detrend, const float *x, const float *win, int n, float *c);

/* ---- PsdCascade<N> (psd.rs:399-544) ---- */
typedef struct {
    uint64_t start;      /* start index in PSD and frequencies */
    uint32_t include;    /* was included in output */
    uint32_t count;      /* number of averages */
    uint32_t avg;        /* averaging limit */
    uint32_t _pad;
    uint64_t bins_start; /* bins: Range<usize> */
    uint64_t bins_end;
    uint64_t fft_size;
    uint64_t decimation;
    uint64_t pending;    /* unprocessed input items (includes overlap) */
    uint64_t processed;  /* items processed (excluding overlap, ignoring averaging) */
} orc_break;

typedef struct orc_cascade orc_cascade;
orc_cascade *orc_cascade_new(int n, int hbf_preset);
orc_cascade *orc_cascade_clone(const orc_cascade *c);
void orc_cascade_free(orc_cascade *c);
float orc_cascade_rbw(const orc_cascade *c);
void orc_cascade_set_avg(orc_cascade *c, uint32_t limit, uint32_t count);
void orc_cascade_set_detrend(orc_cascade *c, int detrend);
void orc_cascade_process(orc_cascade *c, const float *x, size_t n);
size_t orc_cascade_num_stages(const orc_cascade *c);
const orc_stage *orc_cascade_stage(const orc_cascade *c, size_t i);
/* psd(): p must hold stages*(N/2+1) floats, b must hold `stages` breaks.  Returns len(p); *nb = len(b) */
size_t orc_cascade_psd(const orc_cascade *c, int keep_overlap, uint32_t min_count,
                       int keep_transition_band, float *p, orc_break *b, size_t *nb);
/* Break::frequencies (psd.rs:315-327); returns number written */
size_t orc_break_frequencies(const orc_break *b, size_t nb, float *f);

/* ---- frame decode (de/frame.rs, de/data.rs) and loss (loss.rs) ---- */
enum { ORC_OK = 0, ORC_EHEADER = 5, ORC_EFORMAT = 6, ORC_ESIZE = 7, ORC_EBATCHES = 8, ORC_ESHORT = 9 };
typedef struct {
    uint8_t format;
    uint8_t batches;
    uint32_t seq;
} orc_header;
/* Frame::from_bytes + payload.traces().  traces[t] must hold ORC_MAX_TRACE_SAMPLES floats. */
#define ORC_MAX_TRACES 4
int orc_frame_decode(const uint8_t *buf, size_t len, orc_header *hdr, float *const *traces,
                     size_t *samples_per_trace, int *n_traces);
typedef struct {
    uint64_t received;
    uint64_t dropped;
    uint32_t seq;
    uint8_t has_seq;
} orc_loss;
void orc_loss_update(orc_loss *l, uint32_t seq, uint8_t batches); /* loss.rs:11-26 */
float orc_loss_ratio(const orc_loss *l);                          /* loss.rs:29-30 */

/* ---- Var::eval (var.rs:26-45) ---- */
float orc_var_eval(int x_exp, int sinx_exp, float clip, size_t dc_cut, const float *phase_psd,
                   const float *frequencies, size_t n, float tau);

/* ---- Trace::plot + Trapezoidal (bin/psd.rs:96-157): trapezoidal integral over the merged spectrum ---- */
size_t orc_trace_plot(float fs, float integral_start, float integral_end, int integrate, const float *psd,
                      const float *frequencies, size_t n, float *integral, double *xy);

/* ---- synthetic sources (source.rs:66-73, 104-134) ----
 * Noise: uniform (0,1) -> (x - 0.5) * sqrt(12), folded through |noise| first-order integrators
 * (noise < 0) or differentiators (noise > 0) with f32 state, exactly the fold at source.rs:110-114.
 * The reference draws from rand 0.10 SmallRng (not under /root/reference; SURVEY.md 8c says "we use
 * our own seeded generators"): here the uniform stream is Philox4x32-10, key = seed, counter = i/4,
 * word i%4, pinned against the Random123 known-answer vectors (tests/test_oracle_source.py).
 * Dsm: phase accumulator + MASH-1-1-1 (idsp::Dsm<3>, not under /root/reference: textbook form
 * y = c1 + D c2 + D^2 c3 over three wrapping u32 accumulators -- parity unpinned). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
enum { ORC_SOURCE_NOISE = 0, ORC_SOURCE_DSM = 1 };
#define ORC_SOURCE_MAX_ORDER 8
typedef struct orc_source orc_source;
orc_source *orc_source_new(int kind, int64_t param, uint64_t seed);
void orc_source_free(orc_source *s);
void orc_source_get(orc_source *s, float *out, size_t n);
/* the DSM input word of sample index i (source.rs:122-125) */
uint32_t orc_dsm_input(uint64_t i, uint32_t ftw);

#ifdef __cplusplus
}
#endif
#endif
