/* TEST INFRASTRUCTURE ONLY.  CPU restatement ("port") of the DCTZ hot path, used as the parity
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing in the
 * product path (dctz_b200/, include/) may call, link or import this file.
 *
 * Parity status: the reference (swson/DCTZ v0.2.2) ships NO golden vectors or known-answer tests
 * (SURVEY.md §8c), so this oracle is pinned two other ways, both exercised by tests/test_oracle.py:
 *   (1) against the UNMODIFIED reference sources compiled into oracle/_ref/ (Makefile in this
 *       directory; FFTW replaced by oracle/fftw_standin) -- bin indices, DC, outliers, qtable,
 *       statistics and reconstructions must be bit-identical;
 *   (2) against genuine FFTW outputs (REDFT10/REDFT01 vectors shipped with SciPy's test-suite,
 *       committed under tests/golden/ with the script that extracted them).
 *
 * Every function cites the reference lines it restates.  Arithmetic types, operation order and
 * float/double promotions follow the reference exactly; the code is written from scratch.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "fftw3.h" /* oracle/fftw_standin: the DFT the reference obtains from FFTW */

#define ORACLE_BLK 64   /* dctz.h:28 BLK_SZ */
#define ORACLE_NBINS 255 /* dctz.h:63-66: t_bin_id = unsigned char -> NBINS = 255 */

typedef struct {
  double max, min, sum, mean, sf; /* float path: the float results, widened exactly */
} oracle_stat;

/* ------------------------------------------------------------------------------------------
 * util.c:12-44  calc_data_stat
 * ---------------------------------------------------------------------------------------- */
void oracle_calc_stat_d(const double *x, long n, oracle_stat *st) {
  double mx = fabs(x[0]), mn = fabs(x[0]), sum = 0.0; /* util.c:17-19: sum starts at 0, loop from 1 */
  long i;
  for (i = 1; i < n; i++) {
    double a = fabs(x[i]);
    if (a > mx) mx = a;
    if (a < mn) mn = a;
    sum += x[i]; /* util.c:24 -- element 0 is never added */
  }
  st->max = mx; st->min = mn; st->sum = sum;
  st->mean = sum / n;                        /* util.c:27 */
  st->sf = pow(10, ceil(log10(mx)) - 1);     /* util.c:28, SF_ADJ_AMT = 1 */
}

void oracle_calc_stat_f(const float *x, long n, oracle_stat *st) {
  float mx = fabsf(x[0]), mn = fabsf(x[0]), sum = 0.0f; /* util.c:31-33 */
  long i;
  for (i = 1; i < n; i++) {
    float a = fabsf(x[i]);
    if (a > mx) mx = a;
    if (a < mn) mn = a;
    sum += x[i];
  }
  st->max = mx; st->min = mn; st->sum = sum;
  st->mean = (float)(sum / n);                        /* util.c:41 (int n -> float division) */
  st->sf = (float)powf(10, ceil(log10f(mx)) - 1);     /* util.c:42: ceil() is the double one */
}

/* ------------------------------------------------------------------------------------------
 * dct.c:24-53 dct_init + dct.c:55-103 dct_fftw  (orthonormal DCT-II through one complex DFT)
 * dct-float.c:24-54, 56-104 for the float twin.
 * ---------------------------------------------------------------------------------------- */
/* Plan + weight cache: the reference initialises once per length (dct_init) and re-initialises
 * only for the partial tail block; the oracle keeps one context per (direction, type) and rebuilds
 * it when the length changes.  Not thread safe -- neither is the reference (dct.c:18-22). */
typedef struct { int dn; fftw_complex *in, *out; double *wa, *wb; fftw_plan p; } ctx_d;
typedef struct { int dn; fftwf_complex *in, *out; float *wa, *wb; fftwf_plan p; } ctx_f;
static ctx_d g_fwd_d, g_inv_d;
static ctx_f g_fwd_f, g_inv_f;

static void ctx_d_reset(ctx_d *c, int dn) {
  if (c->dn) { fftw_destroy_plan(c->p); fftw_free(c->in); fftw_free(c->out); free(c->wa); free(c->wb); }
  c->dn = dn;
  c->in = (fftw_complex *)fftw_malloc(sizeof(fftw_complex) * 2 * dn);
  c->out = (fftw_complex *)fftw_malloc(sizeof(fftw_complex) * 2 * dn);
  c->wa = (double *)malloc(sizeof(double) * dn);
  c->wb = (double *)malloc(sizeof(double) * dn);
}
static void ctx_f_reset(ctx_f *c, int dn) {
  if (c->dn) { fftwf_destroy_plan(c->p); fftwf_free(c->in); fftwf_free(c->out); free(c->wa); free(c->wb); }
  c->dn = dn;
  c->in = (fftwf_complex *)fftwf_malloc(sizeof(fftwf_complex) * 2 * dn);
  c->out = (fftwf_complex *)fftwf_malloc(sizeof(fftwf_complex) * 2 * dn);
  c->wa = (float *)malloc(sizeof(float) * dn);
  c->wb = (float *)malloc(sizeof(float) * dn);
}

void oracle_dct_d(const double *a, double *b, int dn) {
  ctx_d *c = &g_fwd_d;
  int i, j, k;
  if (c->dn != dn) { /* dct.c:24-53 dct_init */
    ctx_d_reset(c, dn);
    for (i = 0; i < dn; i++) { /* dct.c:37-41 */
      double y = -i * M_PI / (2 * dn);
      c->wa[i] = exp(0.0) * cos(y) / sqrt(2.0 * dn);
      c->wb[i] = exp(0.0) * sin(y) / sqrt(2.0 * dn);
    }
    c->wa[0] = c->wa[0] / sqrt(2.0); /* dct.c:42 */
    if (dn % 2 == 0) {               /* dct.c:43-49 */
      for (i = 0; i < dn; i++) { c->wa[i] = c->wa[i] * 2; c->wb[i] = c->wb[i] * 2; }
      c->p = fftw_plan_dft_1d(dn, c->in, c->out, FFTW_FORWARD, FFTW_ESTIMATE);
    } else {                         /* dct.c:50-52 */
      c->p = fftw_plan_dft_1d(2 * dn, c->in, c->out, FFTW_FORWARD, FFTW_ESTIMATE);
    }
  }
  memset(c->in, 0, sizeof(fftw_complex) * 2 * dn);
  if (dn % 2 == 0) {          /* dct.c:73-92: even/odd reorder, length-dn DFT */
    for (i = 0, j = 0, k = dn - 1; i < dn; i++) {
      if (i % 2) c->in[k--][0] = a[i]; else c->in[j++][0] = a[i];
    }
  } else {                    /* dct.c:59-72: mirrored length-2dn DFT */
    for (i = 0; i < dn; i++) { c->in[i][0] = a[i]; c->in[dn + i][0] = a[dn - 1 - i]; }
  }
  fftw_execute(c->p);
  for (i = 0; i < dn; i++) b[i] = c->wa[i] * c->out[i][0] - c->wb[i] * c->out[i][1]; /* dct.c:100-102 */
}

void oracle_dct_f(const float *a, float *b, int dn) {
  ctx_f *c = &g_fwd_f;
  int i, j, k;
  if (c->dn != dn) { /* dct-float.c:24-54 dct_init_f */
    ctx_f_reset(c, dn);
    for (i = 0; i < dn; i++) { /* dct-float.c:38-43 */
      float y = -i * (float)M_PI / (2 * dn);
      c->wa[i] = expf(0.0f) * cosf(y) / sqrtf(2.0 * dn);
      c->wb[i] = expf(0.0f) * sinf(y) / sqrtf(2.0 * dn);
    }
    c->wa[0] = c->wa[0] / sqrtf(2.0);
    if (dn % 2 == 0) {
      for (i = 0; i < dn; i++) { c->wa[i] = c->wa[i] * 2; c->wb[i] = c->wb[i] * 2; }
      c->p = fftwf_plan_dft_1d(dn, c->in, c->out, FFTW_FORWARD, FFTW_ESTIMATE);
    } else {
      c->p = fftwf_plan_dft_1d(2 * dn, c->in, c->out, FFTW_FORWARD, FFTW_ESTIMATE);
    }
  }
  memset(c->in, 0, sizeof(fftwf_complex) * 2 * dn);
  if (dn % 2 == 0) {
    for (i = 0, j = 0, k = dn - 1; i < dn; i++) {
      if (i % 2) c->in[k--][0] = a[i]; else c->in[j++][0] = a[i];
    }
  } else {
    for (i = 0; i < dn; i++) { c->in[i][0] = a[i]; c->in[dn + i][0] = a[dn - 1 - i]; }
  }
  fftwf_execute(c->p);
  for (i = 0; i < dn; i++) b[i] = c->wa[i] * c->out[i][0] - c->wb[i] * c->out[i][1];
}

/* ------------------------------------------------------------------------------------------
 * dct.c:115-205 ifft_idct (orthonormal DCT-III = inverse of the above); dct-float.c:116-208.
 * ---------------------------------------------------------------------------------------- */
void oracle_idct_d(const double *a, double *data, int dn) {
  ctx_d *c = &g_inv_d;
  double w0;
  int i, j, k;
  if (c->dn != dn) { /* dct.c:120-140: lazy initialisation on first use */
    ctx_d_reset(c, dn);
    for (i = 0; i < dn; i++) {
      double y = i * M_PI / (2 * dn);
      c->wa[i] = exp(0.0) * cos(y) * sqrt(2.0 * dn);
      c->wb[i] = exp(0.0) * sin(y) * sqrt(2.0 * dn);
    }
    c->p = fftw_plan_dft_1d((dn % 2) ? 2 * dn : dn, c->in, c->out, FFTW_BACKWARD, FFTW_ESTIMATE);
  }
  memset(c->in, 0, sizeof(fftw_complex) * 2 * dn);
  memset(c->out, 0, sizeof(fftw_complex) * 2 * dn);
  if (dn % 2 == 1) { /* dct.c:144-203 odd branch */
    w0 = c->wa[0] * sqrt(2.0);
    c->in[0][0] = w0 * a[0]; c->in[0][1] = c->wb[0] * a[0];
    for (i = 1; i < dn; i++) {
      c->in[i][0] = c->wa[i] * a[i];            c->in[i][1] = c->wb[i] * a[i];
      c->in[dn + i][0] = c->wb[i] * a[dn - i];  c->in[dn + i][1] = -c->wa[i] * a[dn - i];
    }
    fftw_execute(c->p);
    for (i = 0; i < dn; i++) data[i] = c->out[i][0] / dn / 2;
  } else { /* dct.c:144-203 even branch */
    w0 = c->wa[0] / sqrt(2.0);
    c->in[0][0] = w0 * a[0]; c->in[0][1] = c->wb[0] * a[0];
    for (i = 1; i < dn; i++) { c->in[i][0] = c->wa[i] * a[i]; c->in[i][1] = c->wb[i] * a[i]; }
    fftw_execute(c->p);
    for (i = 0; i < dn; i++) { c->out[i][0] = c->out[i][0] / dn; c->out[i][1] = c->out[i][1] / dn; }
    for (i = 0, j = 0, k = dn - 1; i < dn; i++) {
      if (i % 2) data[i] = c->out[k--][0]; else data[i] = c->out[j++][0];
    }
  }
}

void oracle_idct_f(const float *a, float *data, int dn) {
  ctx_f *c = &g_inv_f;
  float w0;
  int i, j, k;
  if (c->dn != dn) { /* dct-float.c:124-141: lazy initialisation on first use */
    ctx_f_reset(c, dn);
    for (i = 0; i < dn; i++) {
      float y = i * (float)M_PI / (2 * dn);
      c->wa[i] = expf(0.0f) * cosf(y) * sqrtf(2.0 * dn);
      c->wb[i] = expf(0.0f) * sinf(y) * sqrtf(2.0 * dn);
    }
    c->p = fftwf_plan_dft_1d((dn % 2) ? 2 * dn : dn, c->in, c->out, FFTW_BACKWARD, FFTW_ESTIMATE);
  }
  memset(c->in, 0, sizeof(fftwf_complex) * 2 * dn);
  memset(c->out, 0, sizeof(fftwf_complex) * 2 * dn);
  if (dn % 2 == 1) { /* dct-float.c:147-206 odd branch */
    w0 = c->wa[0] * sqrtf(2.0);
    c->in[0][0] = w0 * a[0]; c->in[0][1] = c->wb[0] * a[0];
    for (i = 1; i < dn; i++) {
      c->in[i][0] = c->wa[i] * a[i];            c->in[i][1] = c->wb[i] * a[i];
      c->in[dn + i][0] = c->wb[i] * a[dn - i];  c->in[dn + i][1] = -c->wa[i] * a[dn - i];
    }
    fftwf_execute(c->p);
    for (i = 0; i < dn; i++) data[i] = c->out[i][0] / dn / 2;
  } else { /* dct-float.c:147-206 even branch */
    w0 = c->wa[0] / sqrtf(2.0);
    c->in[0][0] = w0 * a[0]; c->in[0][1] = c->wb[0] * a[0];
    for (i = 1; i < dn; i++) { c->in[i][0] = c->wa[i] * a[i]; c->in[i][1] = c->wb[i] * a[i]; }
    fftwf_execute(c->p);
    for (i = 0; i < dn; i++) { c->out[i][0] = c->out[i][0] / dn; c->out[i][1] = c->out[i][1] / dn; }
    for (i = 0, j = 0, k = dn - 1; i < dn; i++) {
      if (i % 2) data[i] = c->out[k--][0]; else data[i] = c->out[j++][0];
    }
  }
}

/* Definition-based "truth": b[k] = sqrt(2/dn) c_k sum_n a[n] cos(pi (2n+1) k / (2 dn)), c_0 = 1/sqrt2,
 * accumulated in long double.  Used to classify quantisation-boundary ties and to measure the
 * coefficient error of every implementation (reference, oracle, GPU) against the same yardstick. */
void oracle_dct_exact(const double *a, double *b, int dn) {
  const long double pi = 3.14159265358979323846264338327950288L;
  int k, n;
  for (k = 0; k < dn; k++) {
    long double s = 0.0L;
    for (n = 0; n < dn; n++) {
      /* reduce the angle index mod 4*dn so cosl sees a small argument */
      long idx = ((long)(2 * n + 1) * k) % (4L * dn);
      s += (long double)a[n] * cosl(pi * (long double)idx / (2.0L * dn));
    }
    s *= sqrtl(2.0L / dn);
    if (k == 0) s /= sqrtl(2.0L);
    b[k] = (double)s;
  }
}

void oracle_idct_exact(const double *a, double *x, int dn) {
  const long double pi = 3.14159265358979323846264338327950288L;
  int k, n;
  for (n = 0; n < dn; n++) {
    long double s = (long double)a[0] / sqrtl(2.0L);
    for (k = 1; k < dn; k++) {
      long idx = ((long)(2 * n + 1) * k) % (4L * dn);
      s += (long double)a[k] * cosl(pi * (long double)idx / (2.0L * dn));
    }
    x[n] = (double)(s * sqrtl(2.0L / dn));
  }
}

/* ------------------------------------------------------------------------------------------
 * binning.c:12-30 gen_bins, 32-50 gen_bins_f  (BRSF = 1.0, dctz.h:29)
 * ---------------------------------------------------------------------------------------- */
void oracle_gen_bins_d(double *center, int nbins, double eb) {
  double bw = eb * 2 * 1.0;
  int i;
  center[0] = 0.0;
  for (i = 1; i < nbins; i++) {
    int t = (i % 2) ? ((i / 2) + 1) : -(i / 2);
    center[i] = t * bw;
  }
}

void oracle_gen_bins_f(float *center, int nbins, float eb) {
  float bw = eb * 2 * 1.0; /* float*int -> float, *double -> double, stored as float */
  int i;
  center[0] = 0.0;
  for (i = 1; i < nbins; i++) {
    int t = (i % 2) ? ((i / 2) + 1) : -(i / 2);
    center[i] = t * bw;
  }
}

/* dctz-comp-lib.c:27-43 conv_tbl: ordinal bin t (0..254, t=127 is the zero-centred bin) to the
 * centre-out id used in the stream: 127->0, 128->1, 126->2, 129->3, ... 0->254, 254->253. */
static unsigned char conv_ordinal(int t) { return (unsigned char)((t <= 127) ? (254 - 2 * t) : (2 * (t - 127) - 1)); }
unsigned char oracle_conv_tbl(int t) { return conv_ordinal(t); }

/* ------------------------------------------------------------------------------------------
 * Compress core = dctz-comp-lib.c:186-217 (stats + in-place scale), 271-281 (quantiser params),
 * 318-420 (DCT + bin indices + qtable max), 443-476 (qtable clamp), 478-544 (ordered outliers).
 *
 * `buf` is scaled in place exactly like the caller's buffer in the reference.
 * `coef` (optional, N values) receives the coefficients (the -DDCT_FILE_DEBUG dump).
 * `qtable_raw` / `qtable` (64 values, element type; may be NULL when !qt) receive the table
 * before (qtable.bin) and after the >=1.0 clamp (the stream trailer).
 * `n_edge` counts coefficients that hit ordinal 255 (item == range_max after rounding): the
 * reference reads conv_tbl[255], one past the table (undefined behaviour).  The oracle -- and every
 * GPU path -- stores such a coefficient as an OUTLIER (bin id 255, value kept in AC_exact: no error at
 * all) and reports the count; in QT mode it is rescaled with the range_max branch.
 * ---------------------------------------------------------------------------------------- */
int oracle_compress_core_d(double *buf, long n, double eb, int qt, unsigned char *bin_index, float *dc,
                           float *ac_exact, unsigned *n_out, double *qtable_raw, double *qtable,
                           oracle_stat *st, double *coef, unsigned long *n_edge) {
  const long nblk = (n + ORACLE_BLK - 1) / ORACLE_BLK; /* CEIL, dctz.h:42 */
  const int rem = (int)(n % ORACLE_BLK);
  const int half = ORACLE_NBINS / 2;
  const double bin_width = eb * 2.0 * 1.0;            /* dctz-comp-lib.c:273 */
  const double range_min = -(half * 2 + 1) * (eb * 1.0); /* :274 */
  const double range_max = (half * 2 + 1) * (eb * 1.0);  /* :275 */
  const double qt_factor = 10.0;                         /* :473 (NBINS == 255) */
  double *ax = coef ? coef : (double *)malloc(sizeof(double) * (size_t)n);
  double qtab[ORACLE_BLK];
  unsigned cnt = 0;
  unsigned long edge = 0;
  long i;
  int j;

  oracle_calc_stat_d(buf, n, st);                                   /* :186 */
  if (st->sf != 1.0) for (i = 0; i < n; i++) buf[i] /= st->sf;      /* :193-201 */
  if (n > 0) memset(bin_index, 0, (size_t)n);                       /* :151 */
  for (j = 0; j < ORACLE_BLK; j++) qtab[j] = 0.0;                   /* :160-162 */

  for (i = 0; i < nblk; i++) { /* :325-416 */
    const int l = (i == nblk - 1 && rem != 0) ? rem : ORACLE_BLK;
    oracle_dct_d(buf + i * ORACLE_BLK, ax + i * ORACLE_BLK, l);
    dc[i] = (float)ax[i * ORACLE_BLK];       /* :351 USE_TRUNCATE */
    qtab[0] = ax[i * ORACLE_BLK];            /* :357 -> ends as the last block's DC */
    bin_index[i * ORACLE_BLK] = ORACLE_NBINS; /* :361 */
    for (j = 1; j < l; j++) {
      double item = ax[i * ORACLE_BLK + j];
      unsigned char id;
      if (item < range_min || item > range_max) { /* :367 */
        id = ORACLE_NBINS;
        if (fabs(item) >= qtab[j]) qtab[j] = fabs(item); /* :371-372 (QT only; harmless otherwise) */
      } else {
        int t = (unsigned char)((item - range_min) / bin_width); /* :377 */
        if (t > 254) { id = ORACLE_NBINS; edge++; if (fabs(item) >= qtab[j]) qtab[j] = fabs(item); }
        else id = conv_ordinal(t); /* :378 */
      }
      bin_index[i * ORACLE_BLK + j] = id;
    }
  }
  if (qtable_raw) memcpy(qtable_raw, qtab, sizeof qtab); /* qtable.bin, :443-448 */
  for (j = 1; j < ORACLE_BLK; j++) if (qtab[j] < 1.0) qtab[j] = 1.0; /* :450-461 */
  if (qtable) memcpy(qtable, qtab, sizeof qtab);

  for (i = 0; i < nblk; i++) { /* :478-544 */
    const int l = (i == nblk - 1 && rem != 0) ? rem : ORACLE_BLK;
    for (j = 1; j < l; j++) {
      if (bin_index[i * ORACLE_BLK + j] != ORACLE_NBINS) continue;
      if (qt) {
        double item = ax[i * ORACLE_BLK + j];
        if (item < range_min) item = (item / qtab[j]) * eb * qt_factor + range_min;      /* :489 */
        else if (item > range_max || item > 0) item = (item / qtab[j]) * eb * qt_factor + range_max; /* :491; `item > 0` only
                                                         adds the n_edge elements (ordinal 255 although item <= range_max) */
        /* the reference also stores the rescaled value back into a_x (:493); nothing reads it afterwards, and `coef`
         * is meant to equal the -DDCT_FILE_DEBUG dump taken before this loop (:422-433), so it is not mirrored */
        if (item < range_min || item > range_max) ac_exact[cnt++] = (float)item; /* :494-497 */
        /* else: the reference computes a bin id and drops it (:502-506): nothing is stored */
      } else {
        ac_exact[cnt++] = (float)ax[i * ORACLE_BLK + j]; /* :537 */
      }
    }
  }
  *n_out = cnt;
  if (n_edge) *n_edge = edge;
  if (!coef) free(ax);
  return 1;
}

int oracle_compress_core_f(float *buf, long n, double eb, int qt, unsigned char *bin_index, float *dc,
                           float *ac_exact, unsigned *n_out, float *qtable_raw, float *qtable,
                           oracle_stat *st, float *coef, unsigned long *n_edge) {
  const long nblk = (n + ORACLE_BLK - 1) / ORACLE_BLK;
  const int rem = (int)(n % ORACLE_BLK);
  const int half = ORACLE_NBINS / 2;
  const float bin_width = eb * 2.0 * 1.0;               /* dctz-comp-lib.c:278: double, stored float */
  const float range_min = -(half * 2 + 1) * (eb * 1.0); /* :279 */
  const float range_max = (half * 2 + 1) * (eb * 1.0);  /* :280 */
  const float qt_factor = 10.0;                         /* :475 */
  float *ax = coef ? coef : (float *)malloc(sizeof(float) * (size_t)n);
  float qtab[ORACLE_BLK];
  unsigned cnt = 0;
  unsigned long edge = 0;
  long i;
  int j;

  oracle_calc_stat_f(buf, n, st);
  {
    const float sf = (float)st->sf;
    if (sf != 1.0) for (i = 0; i < n; i++) buf[i] /= sf; /* :208-216 */
  }
  if (n > 0) memset(bin_index, 0, (size_t)n);
  for (j = 0; j < ORACLE_BLK; j++) qtab[j] = 0.0;

  for (i = 0; i < nblk; i++) {
    const int l = (i == nblk - 1 && rem != 0) ? rem : ORACLE_BLK;
    oracle_dct_f(buf + i * ORACLE_BLK, ax + i * ORACLE_BLK, l);
    dc[i] = ax[i * ORACLE_BLK];
    qtab[0] = ax[i * ORACLE_BLK]; /* :359 */
    bin_index[i * ORACLE_BLK] = ORACLE_NBINS;
    for (j = 1; j < l; j++) {
      float item = ax[i * ORACLE_BLK + j]; /* :390 */
      unsigned char id;
      if (item < range_min || item > range_max) { /* :392 */
        id = ORACLE_NBINS;
        if (fabsf(item) >= qtab[j]) qtab[j] = fabsf(item); /* :396-397 */
      } else {
        int t = (unsigned char)((item - range_min) / bin_width); /* :402, float arithmetic */
        if (t > 254) { id = ORACLE_NBINS; edge++; if (fabsf(item) >= qtab[j]) qtab[j] = fabsf(item); }
        else id = conv_ordinal(t);
      }
      bin_index[i * ORACLE_BLK + j] = id;
    }
  }
  if (qtable_raw) memcpy(qtable_raw, qtab, sizeof qtab);
  for (j = 1; j < ORACLE_BLK; j++) if (qtab[j] < 1.0) qtab[j] = 1.0; /* :456-459 */
  if (qtable) memcpy(qtable, qtab, sizeof qtab);

  for (i = 0; i < nblk; i++) {
    const int l = (i == nblk - 1 && rem != 0) ? rem : ORACLE_BLK;
    for (j = 1; j < l; j++) {
      if (bin_index[i * ORACLE_BLK + j] != ORACLE_NBINS) continue;
      if (qt) {
        float item = ax[i * ORACLE_BLK + j];
        /* :515/:517 -- (float/float) is float, then promoted to double by error_bound */
        if (item < range_min) item = (item / qtab[j]) * eb * qt_factor + range_min;
        else if (item > range_max || item > 0) item = (item / qtab[j]) * eb * qt_factor + range_max; /* `item > 0`: n_edge elements only */
        if (item < range_min || item > range_max) ac_exact[cnt++] = item; /* :520-523 (write-back to a_x not mirrored, see above) */
      } else {
        ac_exact[cnt++] = ax[i * ORACLE_BLK + j]; /* :537 */
      }
    }
  }
  *n_out = cnt;
  if (n_edge) *n_edge = edge;
  if (!coef) free(ax);
  return 1;
}

/* ------------------------------------------------------------------------------------------
 * Decompress core = dctz-decomp-lib.c:358-361 (bin centres), 370-381 (ranges), 389-483
 * (dequantise + IDCT), 494-511 (de-scale).  conv_tbl_i (:23-39) is the identity.
 * `coef` (optional) receives the rebuilt coefficients.
 * ---------------------------------------------------------------------------------------- */
int oracle_decompress_core_d(const unsigned char *bin_index, const float *dc, const float *ac_exact,
                             const double *qtable, long n, double eb, double sf, int qt, double *out,
                             double *coef) {
  const long nblk = (n + ORACLE_BLK - 1) / ORACLE_BLK;
  const int rem = (int)(n % ORACLE_BLK);
  const double range_max = eb * ORACLE_NBINS, range_min = -eb * ORACLE_NBINS; /* :373-374 */
  const double qt_factor = 10.0;
  double center[ORACLE_NBINS];
  double *xr = coef ? coef : (double *)malloc(sizeof(double) * (size_t)n);
  unsigned pos = 0;
  long i;
  int j;
  oracle_gen_bins_d(center, ORACLE_NBINS, eb);
  for (i = 0; i < nblk; i++) {
    const int l = (i == nblk - 1 && rem != 0) ? rem : ORACLE_BLK;
    xr[i * ORACLE_BLK] = dc[i]; /* :392 */
    for (j = 1; j < l; j++) {
      unsigned char id = bin_index[i * ORACLE_BLK + j];
      if (id == ORACLE_NBINS) {
        double v = ac_exact[pos++]; /* :402-403 */
        if (qt) {
          if (v > 0) v = ((v - range_max) / (eb * qt_factor)) * qtable[j]; /* :405 */
          else v = ((v - range_min) / (eb * qt_factor)) * qtable[j];       /* :408 */
        }
        xr[i * ORACLE_BLK + j] = v;
      } else {
        xr[i * ORACLE_BLK + j] = center[id]; /* :416 */
      }
    }
    oracle_idct_d(xr + i * ORACLE_BLK, out + i * ORACLE_BLK, l); /* :428 */
  }
  if (sf != 1.0) for (i = 0; i < n; i++) out[i] *= sf; /* :496-502 */
  if (!coef) free(xr);
  return 1;
}

int oracle_decompress_core_f(const unsigned char *bin_index, const float *dc, const float *ac_exact,
                             const float *qtable, long n, double eb, float sf, int qt, float *out,
                             float *coef) {
  const long nblk = (n + ORACLE_BLK - 1) / ORACLE_BLK;
  const int rem = (int)(n % ORACLE_BLK);
  const float range_max = eb * ORACLE_NBINS, range_min = -eb * ORACLE_NBINS; /* :378-379 */
  const float qt_factor = 10.0;
  float center[ORACLE_NBINS];
  float *xr = coef ? coef : (float *)malloc(sizeof(float) * (size_t)n);
  unsigned pos = 0;
  long i;
  int j;
  oracle_gen_bins_f(center, ORACLE_NBINS, eb); /* :361: double eb converted to the float parameter */
  for (i = 0; i < nblk; i++) {
    const int l = (i == nblk - 1 && rem != 0) ? rem : ORACLE_BLK;
    xr[i * ORACLE_BLK] = dc[i];
    for (j = 1; j < l; j++) {
      unsigned char id = bin_index[i * ORACLE_BLK + j];
      if (id == ORACLE_NBINS) {
        float v = ac_exact[pos++];
        if (qt) { /* :450-454: float subtraction, then double division and product, stored float */
          if (v > 0) v = ((v - range_max) / (eb * qt_factor)) * qtable[j];
          else v = ((v - range_min) / (eb * qt_factor)) * qtable[j];
        }
        xr[i * ORACLE_BLK + j] = v;
      } else {
        xr[i * ORACLE_BLK + j] = center[id];
      }
    }
    oracle_idct_f(xr + i * ORACLE_BLK, out + i * ORACLE_BLK, l);
  }
  if (sf != 1.0) for (i = 0; i < n; i++) out[i] *= sf; /* :505-510 */
  if (!coef) free(xr);
  return 1;
}
