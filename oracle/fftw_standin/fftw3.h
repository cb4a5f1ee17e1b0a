/* TEST INFRASTRUCTURE ONLY (oracle/): minimal FFTW3-API stand-in.
 *
 * FFTW3 (the reference's only numerical dependency: Makefile:4 `-lfftw3 -lfftw3f`, no version
 * pin; README.md:25-46 suggests 3.3.10) is not installed in this image and there is no network.
 * This header + fftw_standin.c provide exactly the handful of entry points the reference calls
 * (dct.c:28-29,48,51,72,91,107-112,157,160,179,182; dct-float.c twins) so that the UNMODIFIED
 * reference sources under /root/reference compile into oracle/_ref/.  The transform computed is
 * FFTW's documented one:  out[k] = sum_j in[j] * exp(sign * 2*pi*i * j*k / n)  (unnormalised).
 * Rounding differs from real FFTW at the 1e-16 (double) / 1e-7 (float) relative level.
 */
#ifndef DCTZ_ORACLE_FFTW3_STANDIN_H
#define DCTZ_ORACLE_FFTW3_STANDIN_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef float fftwf_complex[2];
typedef struct standin_plan_s *fftw_plan;
typedef struct standin_plan_s *fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void fftw_cleanup(void);

void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
void fftwf_cleanup(void);

#ifdef __cplusplus
}
#endif
#endif
