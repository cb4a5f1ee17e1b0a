/* TEST INFRASTRUCTURE ONLY (oracle/): FFTW3-API stand-in, see fftw3.h in this directory.
 * Power-of-two lengths: iterative radix-2 decimation-in-time with a precomputed twiddle table.
 * Any other length (the partial tail block of the codec: rem even -> n=rem, rem odd -> n=2*rem):
 * table-driven O(n^2) DFT. */
#include "fftw3.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct standin_plan_s {
  int n, sign, is_float, pow2, log2n;
  void *in, *out;
  double *wd;   /* n twiddles (re,im) in double */
  float *wf;    /* same rounded to float */
  int *rev;     /* bit reversal permutation (pow2 only) */
};

static struct standin_plan_s *make_plan(int n, void *in, void *out, int sign, int is_float) {
  struct standin_plan_s *p = (struct standin_plan_s *)calloc(1, sizeof *p);
  int k;
  p->n = n; p->sign = sign; p->is_float = is_float; p->in = in; p->out = out;
  p->pow2 = (n > 0) && ((n & (n - 1)) == 0);
  p->wd = (double *)malloc(sizeof(double) * 2 * (size_t)n);
  p->wf = (float *)malloc(sizeof(float) * 2 * (size_t)n);
  for (k = 0; k < n; k++) {
    double ang = (double)sign * 2.0 * M_PI * (double)k / (double)n;
    p->wd[2 * k] = cos(ang); p->wd[2 * k + 1] = sin(ang);
    p->wf[2 * k] = (float)p->wd[2 * k]; p->wf[2 * k + 1] = (float)p->wd[2 * k + 1];
  }
  if (p->pow2) {
    int l = 0; while ((1 << l) < n) l++;
    p->log2n = l;
    p->rev = (int *)malloc(sizeof(int) * (size_t)n);
    for (k = 0; k < n; k++) {
      int r = 0, b;
      for (b = 0; b < l; b++) if (k & (1 << b)) r |= 1 << (l - 1 - b);
      p->rev[k] = r;
    }
  }
  return p;
}

#define DEFINE_EXEC(NAME, T, W)                                                         \
  static void NAME(const struct standin_plan_s *p) {                                    \
    const int n = p->n;                                                                 \
    T(*in)[2] = (T(*)[2])p->in;                                                         \
    T(*out)[2] = (T(*)[2])p->out;                                                       \
    const T *w = p->W;                                                                  \
    int i, j, k;                                                                        \
    if (p->pow2) {                                                                      \
      int half, step;                                                                   \
      for (i = 0; i < n; i++) { out[p->rev[i]][0] = in[i][0]; out[p->rev[i]][1] = in[i][1]; } \
      for (half = 1, step = n >> 1; half < n; half <<= 1, step >>= 1) {                 \
        for (i = 0; i < n; i += 2 * half) {                                             \
          for (j = 0; j < half; j++) {                                                  \
            const T wr = w[2 * j * step], wi = w[2 * j * step + 1];                     \
            const T xr = out[i + j + half][0], xi = out[i + j + half][1];               \
            const T tr = wr * xr - wi * xi, ti = wr * xi + wi * xr;                     \
            const T ur = out[i + j][0], ui = out[i + j][1];                             \
            out[i + j][0] = ur + tr; out[i + j][1] = ui + ti;                           \
            out[i + j + half][0] = ur - tr; out[i + j + half][1] = ui - ti;             \
          }                                                                             \
        }                                                                               \
      }                                                                                 \
    } else {                                                                            \
      for (k = 0; k < n; k++) {                                                         \
        T sr = 0, si = 0; int idx = 0;                                                  \
        for (j = 0; j < n; j++) {                                                       \
          const T wr = w[2 * idx], wi = w[2 * idx + 1];                                 \
          sr += in[j][0] * wr - in[j][1] * wi;                                          \
          si += in[j][0] * wi + in[j][1] * wr;                                          \
          idx += k; if (idx >= n) idx -= n;                                             \
        }                                                                               \
        out[k][0] = sr; out[k][1] = si;                                                 \
      }                                                                                 \
    }                                                                                   \
  }

DEFINE_EXEC(exec_d, double, wd)
DEFINE_EXEC(exec_f, float, wf)

static void destroy(struct standin_plan_s *p) {
  if (!p) return;
  free(p->wd); free(p->wf); free(p->rev); free(p);
}

void *fftw_malloc(size_t n) { void *q = NULL; if (posix_memalign(&q, 64, n ? n : 64)) return NULL; return q; }
void fftw_free(void *p) { free(p); }
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags) {
  (void)flags; return make_plan(n, in, out, sign, 0);
}
void fftw_execute(const fftw_plan p) { exec_d(p); }
void fftw_destroy_plan(fftw_plan p) { destroy(p); }
void fftw_cleanup(void) {}

void *fftwf_malloc(size_t n) { return fftw_malloc(n); }
void fftwf_free(void *p) { free(p); }
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags) {
  (void)flags; return make_plan(n, in, out, sign, 1);
}
void fftwf_execute(const fftwf_plan p) { exec_f(p); }
void fftwf_destroy_plan(fftwf_plan p) { destroy(p); }
void fftwf_cleanup(void) {}
