/* dctz_cli.c -- round-trip test driver with the command line of the reference's dctz-test.c
 * (dctz-test.c:18-92): built as dctz-ec-test and dctz-qt-test like the reference's Makefile:12-17.
 *
 *   dctz-ec-test -d|-f <err bound> <var name> <srcFilePath> <dim1> [dim2 [dim3 [dim4]]]
 *
 * Reads the raw array, compresses it through dctz_compress() (GPU hot path + host zlib), writes
 * <src>.<ec|qt>.<bound>.z, decompresses, writes <src>.<ec|qt>.<bound>.z.r and prints the same
 * "CR = ..., PSNR = ..." summary the reference's scripts collect (tests/test-dctz.sh).
 * Extra: with TIME=1 in the environment it also prints wall-clock times of the two calls.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../../include/dctz_compat.h"

static double now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

static void usage(const char *me) {
  printf("Test case: %s -d|-f [err bound] [var name] [srcFilePath] [dimension sizes...] \n", me);
  printf("Example: %s -d 1E-3 sedov testdata/x86/testfloat_8_8_128.dat 8 8 128 \n", me);
}

/* Fields beyond one stream's `int N` (dctz.h:126): the multi-stream container of dctz_compress_large, written
 * to <src>.<ec|qt>.<bound>.zms; DCTZ_GPUS=<n> spreads its pieces over n devices. */
static int run_large_field(char *argv[], t_datatype dt, size_t n, double eb, const char *mode) {
  const size_t esize = dt == DOUBLE ? sizeof(double) : sizeof(float);
  const size_t cap = dctz_large_bound(n, dt);
  void *data = malloc(n * esize), *rec = malloc(n * esize), *z = malloc(cap);
  char path[1024];
  size_t zsize;
  double t0, t1, t2;
  t_var va, vr;
  FILE *f;
  if (!data || !rec || !z) { fprintf(stderr, "Out of memory\n"); return 1; }
  f = fopen(argv[4], "rb");
  if (!f) { printf("File Not Found\n"); return 1; }
  if (fread(data, esize, n, f) != n) { fprintf(stderr, "short read from %s\n", argv[4]); return 1; }
  fclose(f);
  t0 = now();
  zsize = dctz_compress_large(data, n, dt, eb, z, cap);
  t1 = now();
  snprintf(path, sizeof path, "%s.%s.%s.zms", argv[4], mode, argv[2]);
  f = fopen(path, "wb");
  if (!f || fwrite(z, zsize, 1, f) != 1) { printf("Write zms file failed\n"); return 1; }
  fclose(f);
  printf("outsize = %zu\n", zsize);
  dctz_decompress_large(z, zsize, rec, n);
  t2 = now();
  snprintf(path, sizeof path, "%s.%s.%s.zms.r", argv[4], mode, argv[2]);
  f = fopen(path, "wb");
  if (!f || fwrite(rec, n * esize, 1, f) != 1) { printf("Write zms.r file failed\n"); return 1; }
  fclose(f);
  if (getenv("TIME")) printf("comp_time = %f (s), decomp_time = %f (s) [wall clock]\n", t1 - t0, t2 - t1);
  memset(&va, 0, sizeof va);
  va.datatype = dt; vr = va;
  va.buf.d = (double *)data; vr.buf.d = (double *)rec;
  if (n <= 0x7FFFFFFFu) printf("CR = %.2f, PSNR = %.2f\n", (double)(n * esize) / (double)zsize, calc_psnr(&va, &vr, (int)n, eb));
  else printf("CR = %.2f\n", (double)(n * esize) / (double)zsize);
  free(data); free(rec); free(z);
  printf("done\n");
  return 0;
}

int main(int argc, char *argv[]) {
  const char *mode = dctz_build_is_qt() ? "qt" : "ec";
  size_t dims[4] = {0, 0, 0, 0}, n = 1, esize, out_size = 0;
  t_datatype dt;
  t_var var, var_z, var_r;
  char path[1024];
  FILE *f;
  double t0, t1, t2, eb;
  int i, ndims;

  if (argc < 6 || argc > 9 || (strcmp(argv[1], "-d") && strcmp(argv[1], "-f"))) { usage(argv[0]); return 1; }
  dt = strcmp(argv[1], "-d") ? FLOAT : DOUBLE;
  esize = dt == DOUBLE ? sizeof(double) : sizeof(float);
  eb = atof(argv[2]);
  ndims = argc - 5;
  for (i = 0; i < ndims; i++) { dims[i] = (size_t)atoll(argv[5 + i]); n *= dims[i]; }
  if (n == 0) { usage(argv[0]); return 1; }
  printf("total number of elements = %zu\n", n);
  if (n > ((size_t)1 << 30) || getenv("DCTZ_FORCE_LARGE")) return run_large_field(argv, dt, n, eb, mode);

  memset(&var, 0, sizeof var);
  var.datatype = dt; var.err_bound = eb; var.var_name = argv[3];
  var_z = var; var_r = var;
  var.buf.d = (double *)malloc(n * esize);
  var_r.buf.d = (double *)malloc(n * esize);
  var_z.buf.d = (double *)malloc(2 * n * esize + 4096); /* room even when every coefficient is an outlier */
  if (!var.buf.d || !var_r.buf.d || !var_z.buf.d) { fprintf(stderr, "Out of memory\n"); return 1; }

  f = fopen(argv[4], "rb");
  if (!f) { printf("File Not Found\n"); return 1; }
  if (fread(var.buf.d, esize, n, f) != n) { fprintf(stderr, "short read from %s\n", argv[4]); return 1; }
  fclose(f);

  t0 = now();
  dctz_compress(&var, (int)n, &out_size, &var_z, eb);
  t1 = now();
  snprintf(path, sizeof path, "%s.%s.%s.z", argv[4], mode, argv[2]);
  printf("oriFilePath = %s, outputFilePath = %s, datatype = %s, error = %s, dim1 = %zu, dim2 = %zu, dim3 = %zu, dim4 = %zu\n", argv[4],
         path, dt == FLOAT ? "float" : "double", argv[2], dims[0], dims[1], dims[2], dims[3]);
  printf("outsize = %zu\n", out_size);

  { /* dctz_compress leaves the input divided by the scaling factor (dctz-comp-lib.c:198,213);
       undo that before the quality metrics, as dctz-test.c:186-210 does */
    struct header h;
    memcpy(&h, var_z.buf.d, sizeof h);
    if (dt == DOUBLE) { if (h.scaling_factor.d != 1.0) for (size_t k = 0; k < n; k++) var.buf.d[k] *= h.scaling_factor.d; }
    else { if (h.scaling_factor.f != 1.0) for (size_t k = 0; k < n; k++) var.buf.f[k] *= h.scaling_factor.f; }
  }

  f = fopen(path, "wb");
  if (!f || fwrite(var_z.buf.d, out_size, 1, f) != 1) { printf("Write qtz file failed\n"); return 1; }
  fclose(f);

  t2 = now();
  dctz_decompress(&var_z, &var_r);
  t2 = now() - t2;
  snprintf(path, sizeof path, "%s.%s.%s.z.r", argv[4], mode, argv[2]);
  f = fopen(path, "wb");
  if (!f || fwrite(var_r.buf.d, n * esize, 1, f) != 1) { printf("Write qtz.r file failed\n"); return 1; }
  fclose(f);

  if (getenv("TIME"))
    printf("comp_time = %f (s), decomp_time = %f (s) [wall clock, GPU hot path + host zlib + PCIe]\n", t1 - t0, t2);
  {
    const double psnr = calc_psnr(&var, &var_r, (int)n, eb);
    printf("CR = %.2f, PSNR = %.2f\n", (double)(n * esize) / (double)out_size, psnr);
  }
  free(var.buf.d); free(var_r.buf.d); free(var_z.buf.d);
  printf("done\n");
  return 0;
}
