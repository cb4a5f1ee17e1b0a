/* dctz_compat.h -- declarations of DCTZ's public C API, as libdctz_ec.so / libdctz_qt.so export it.
 *
 * These are the symbols and data layouts of the reference's dctz.h / dct.h (swson/DCTZ v0.2.2) that a
 * caller such as dctz-test.c or dct-test.c links against; a program built against the reference's own
 * headers is binary compatible with these libraries.  Every declaration cites the line it mirrors.
 * The hot path behind them runs on the GPU through include/dctz_gpu.h; zlib stays on the host.
 */
#ifndef DCTZ_COMPAT_H
#define DCTZ_COMPAT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCTZ_BLK_SZ 64  /* BLK_SZ, dctz.h:28 */
#define DCTZ_NBINS 255  /* NBINS for the 8-bit t_bin_id, dctz.h:63-66 */

typedef enum { FLOAT = 0, DOUBLE = 1 } t_datatype; /* dctz.h:44-47 */

typedef struct { /* dctz.h:49-59; only datatype and buf are read by the library */
  t_datatype datatype;
  double err_bound;
  char *var_name;
  union { float *f; double *d; } buf;
} t_var;

typedef unsigned char t_bin_id; /* dctz.h:63 */

typedef union { double d; float f; } dctz_num; /* the double/float unions of dctz.h:68-94 */
typedef struct { dctz_num mean, min, max, range, sf; } t_bstat; /* dctz.h:68-94, same member order */

/* Stream header, dctz.h:96-119: 56 bytes in both build modes (bindex_count exists only in the
 * -DUSE_QTABLE build; the EC build leaves those 4 bytes as padding). */
struct header {
  t_datatype datatype;
  unsigned int num_elements;
  double error_bound;
  unsigned int tot_AC_exact_count;
  dctz_num scaling_factor;
  dctz_num mean;
  unsigned int bindex_sz_compressed;
  unsigned int DC_sz_compressed;
  unsigned int AC_exact_sz_compressed;
#ifdef USE_QTABLE
  unsigned int bindex_count;
#endif
};

/* dctz.h:121-128 */
void calc_data_stat(t_var *in, t_bstat *bs, int N);
void gen_bins(double min, double max, double *bin_center, int nbins, double error_bound);
void gen_bins_f(float min, float max, float *bin_center, int nbins, float error_bound);
int dctz_compress(t_var *var, int N, size_t *outSize, t_var *var_z, double error_bound);
int dctz_decompress(t_var *var_z, t_var *var_r);
double calc_psnr(t_var *var, t_var *var_r, int N, double error_bound);

/* dct.h:17-27 */
void dct_init(int dn);
void dct_init_f(int dn);
void dct_fftw(double *a, double *b, int dn, int nblk);
void dct_fftw_f(float *a, float *b, int dn, int nblk);
void ifft_idct(int dn, double *a, double *data);
void ifft_idct_f(int dn, float *a, float *data);
void dct_finish(void);
void dct_finish_f(void);
void idct_finish(void);
void idct_finish_f(void);

/* Extensions (not in the reference): which build this is, and the device to use (default 0, or the
 * DCTZ_GPU_DEVICE environment variable).  */
int dctz_build_is_qt(void);
void dctz_set_device(int device);
/* stage timers (ms: upload, statistics, transform, wall to kernels, wall of downloads + scaling) and PCIe bytes of the
 * GPU call inside the last dctz_compress / dctz_decompress; 0 on success */
int dctz_host_last_call_stats(double times_ms[8], unsigned long long *h2d_bytes, unsigned long long *d2h_bytes);
/* Large fields (beyond the `int N` / 32-bit header of one stream): a container of block-aligned pieces, each a
 * complete standard DCTZ stream compressed with the GLOBAL scaling factor; DCTZ_GPUS=<n> spreads the pieces over
 * n devices.  dctz_compress_large returns the container size (out must hold dctz_large_bound bytes);
 * dctz_decompress_large returns the element count.  dctz_large_set_piece changes the piece length (tests). */
size_t dctz_large_bound(size_t N, t_datatype dt);
size_t dctz_compress_large(const void *data, size_t N, t_datatype dt, double error_bound, void *out, size_t out_cap);
size_t dctz_decompress_large(const void *in, size_t in_size, void *out, size_t out_elements);
void dctz_large_set_piece(size_t elements);
/* deflate one stream section the way dctz_compress does (chunk-parallel above 2 MiB); returns the size or 0 */
size_t dctz_host_deflate(const void *src, size_t n, void *dst, size_t cap);
/* the same while the section arrives in pieces (what dctz_compress does with the GPU's downloads); byte-identical */
size_t dctz_host_deflate_streamed(const void *src, size_t n, void *dst, size_t cap, size_t piece);
/* the same for stream section `section` (0 bin_index, 1 DC, 2 AC_exact; the float sections are cut into smaller chunks) */
size_t dctz_host_deflate_section(const void *src, size_t n, void *dst, size_t cap, int section, size_t piece);

#ifdef __cplusplus
}
#endif
#endif /* DCTZ_COMPAT_H */
