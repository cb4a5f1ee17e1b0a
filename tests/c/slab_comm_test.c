/* slab_comm_test.c -- two ranks (one host thread + one GPU each, as two MPI ranks would be) compress the two slabs of
 * one field through dctz_gpu_compress_slab_comm: the statistics exchange (and the QT table reduction) happen inside the
 * library over NCCL.  The concatenation of the two slabs' outputs must equal the single-GPU result bit for bit.
 * Plain C against include/dctz_gpu.h; built by dctz_b200/csrc/host/Makefile as bin/slab-comm-test. */
#include <cuda_runtime.h>
#include <nccl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dctz_gpu.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)
#define NK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { fprintf(stderr, "%s: %s\n", #x, ncclGetErrorString(r_)); exit(2); } } while (0)
#define DK(ctx, x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "%s: %s\n", #x, dctz_gpu_last_error(ctx)); exit(2); } } while (0)

enum { NRANKS = 2 };
static const size_t N_TOTAL = 64u * 32u * 3000u + 64u * 7u + 29u; /* a partial last tile and a partial last block */
static const double EB = 1e-3;

typedef struct {
  int rank, mode_qt;
  ncclUniqueId id;
  size_t start, n;
  uint8_t *bins;  /* host copies of the slab's outputs */
  float *dc, *ac;
  double qtable[64];
  dctz_gpu_info info;
} rank_job;

static void fill(dctz_gpu_ctx *ctx, double *d, size_t start, size_t n, int noisy) {
  DK(ctx, dctz_gpu_fill_hash_field(ctx, d, start, n, 2048, 20261018u + (noisy ? 1u : 0u), NULL));
}

static void *rank_main(void *arg) {
  rank_job *j = (rank_job *)arg;
  ncclComm_t comm;
  dctz_gpu_ctx *ctx = NULL;
  cudaStream_t st;
  double *d_in, *d_q, *d_qraw;
  uint8_t *d_bins;
  float *d_dc, *d_ac;
  dctz_gpu_info *d_info;
  const size_t nblk = (j->n + 63) / 64;
  CK(cudaSetDevice(j->rank));
  NK(ncclCommInitRank(&comm, NRANKS, j->id, j->rank));
  DK(NULL, dctz_gpu_create(&ctx, j->rank));
  CK(cudaStreamCreate(&st));
  CK(cudaMalloc((void **)&d_in, j->n * 8)); CK(cudaMalloc((void **)&d_bins, j->n)); CK(cudaMalloc((void **)&d_dc, nblk * 4));
  CK(cudaMalloc((void **)&d_ac, j->n * 4)); CK(cudaMalloc((void **)&d_q, 512)); CK(cudaMalloc((void **)&d_qraw, 512));
  CK(cudaMalloc((void **)&d_info, sizeof(dctz_gpu_info)));
  CK(cudaMemset(d_qraw, 0, 512));
  fill(ctx, d_in, j->start, j->n, 0);
  CK(cudaDeviceSynchronize());
  DK(ctx, dctz_gpu_compress_slab_comm(ctx, comm, j->rank, NRANKS, NRANKS - 1, d_in, j->n, N_TOTAL, DCTZ_GPU_DOUBLE, EB, j->mode_qt, d_bins, d_dc, d_ac,
                                      d_q, d_qraw, d_info, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaMemcpy(&j->info, d_info, sizeof j->info, cudaMemcpyDeviceToHost));
  j->bins = (uint8_t *)malloc(j->n); j->dc = (float *)malloc(nblk * 4); j->ac = (float *)malloc(j->n * 4 + 4);
  CK(cudaMemcpy(j->bins, d_bins, j->n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(j->dc, d_dc, nblk * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(j->ac, d_ac, (size_t)j->info.n_outliers * 4, cudaMemcpyDeviceToHost));
  if (j->mode_qt) CK(cudaMemcpy(j->qtable, d_q, 512, cudaMemcpyDeviceToHost));
  cudaFree(d_in); cudaFree(d_bins); cudaFree(d_dc); cudaFree(d_ac); cudaFree(d_q); cudaFree(d_qraw); cudaFree(d_info);
  cudaStreamDestroy(st);
  dctz_gpu_destroy(ctx);
  ncclCommDestroy(comm);
  return NULL;
}

int main(void) {
  int ndev = 0, mode_qt;
  cudaGetDeviceCount(&ndev);
  if (ndev < NRANKS) { printf("slab_comm_test SKIPPED: %d device(s), %d needed\n", ndev, NRANKS); return 77; }
  for (mode_qt = 0; mode_qt < 2; mode_qt++) {
    rank_job jobs[NRANKS];
    pthread_t th[NRANKS];
    ncclUniqueId id;
    const size_t nblk = (N_TOTAL + 63) / 64, per = (nblk + NRANKS - 1) / NRANKS;
    int r;
    NK(ncclGetUniqueId(&id));
    for (r = 0; r < NRANKS; r++) {
      const size_t b0 = (size_t)r * per < nblk ? (size_t)r * per : nblk, b1 = (size_t)(r + 1) * per < nblk ? (size_t)(r + 1) * per : nblk;
      memset(&jobs[r], 0, sizeof jobs[r]);
      jobs[r].rank = r; jobs[r].mode_qt = mode_qt; jobs[r].id = id;
      jobs[r].start = b0 * 64;
      jobs[r].n = (b1 * 64 < N_TOTAL ? b1 * 64 : N_TOTAL) - b0 * 64;
      pthread_create(&th[r], NULL, rank_main, &jobs[r]);
    }
    for (r = 0; r < NRANKS; r++) pthread_join(th[r], NULL);
    { /* the same field on ONE GPU */
      dctz_gpu_ctx *ctx = NULL;
      double *d_in, *d_q, *d_qraw, q[64];
      uint8_t *d_bins, *bins = (uint8_t *)malloc(N_TOTAL);
      float *d_dc, *d_ac, *dc = (float *)malloc(nblk * 4), *ac;
      dctz_gpu_info *d_info, info;
      size_t off_e = 0, off_b = 0, off_a = 0;
      CK(cudaSetDevice(0));
      DK(NULL, dctz_gpu_create(&ctx, 0));
      CK(cudaMalloc((void **)&d_in, N_TOTAL * 8)); CK(cudaMalloc((void **)&d_bins, N_TOTAL)); CK(cudaMalloc((void **)&d_dc, nblk * 4));
      CK(cudaMalloc((void **)&d_ac, N_TOTAL * 4)); CK(cudaMalloc((void **)&d_q, 512)); CK(cudaMalloc((void **)&d_qraw, 512));
      CK(cudaMalloc((void **)&d_info, sizeof info));
      fill(ctx, d_in, 0, N_TOTAL, 0);
      DK(ctx, dctz_gpu_compress_field_dev(ctx, d_in, N_TOTAL, DCTZ_GPU_DOUBLE, EB, mode_qt, d_bins, d_dc, d_ac, d_q, d_qraw, d_info, NULL));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&info, d_info, sizeof info, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(bins, d_bins, N_TOTAL, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(dc, d_dc, nblk * 4, cudaMemcpyDeviceToHost));
      ac = (float *)malloc((size_t)info.n_outliers * 4 + 4);
      CK(cudaMemcpy(ac, d_ac, (size_t)info.n_outliers * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(q, d_q, 512, cudaMemcpyDeviceToHost));
      for (r = 0; r < NRANKS; r++) {
        const size_t nb = (jobs[r].n + 63) / 64;
        if (jobs[r].info.status != 0 || jobs[r].info.sf != info.sf) { printf("rank %d: status %d sf %g vs %g\n", r, jobs[r].info.status, jobs[r].info.sf, info.sf); return 1; }
        if (memcmp(jobs[r].bins, bins + off_e, jobs[r].n)) { printf("rank %d: bin indices differ from the single-GPU result\n", r); return 1; }
        if (memcmp(jobs[r].dc, dc + off_b, nb * 4)) { printf("rank %d: DC differs\n", r); return 1; }
        if (off_a + jobs[r].info.n_outliers > info.n_outliers || memcmp(jobs[r].ac, ac + off_a, (size_t)jobs[r].info.n_outliers * 4)) { printf("rank %d: outliers differ\n", r); return 1; }
        if (mode_qt && memcmp(jobs[r].qtable, q, 512)) { printf("rank %d: qtable differs\n", r); return 1; }
        off_e += jobs[r].n; off_b += nb; off_a += (size_t)jobs[r].info.n_outliers;
      }
      if (off_a != info.n_outliers) { printf("outlier counts do not add up: %zu vs %llu\n", off_a, (unsigned long long)info.n_outliers); return 1; }
      printf("mode %s: %d ranks over NCCL == one GPU (N = %zu, sf = %g, %llu outliers)\n", mode_qt ? "QT" : "EC", NRANKS, N_TOTAL, info.sf,
             (unsigned long long)info.n_outliers);
      cudaFree(d_in); cudaFree(d_bins); cudaFree(d_dc); cudaFree(d_ac); cudaFree(d_q); cudaFree(d_qraw); cudaFree(d_info);
      dctz_gpu_destroy(ctx);
      free(bins); free(dc); free(ac);
    }
    for (r = 0; r < NRANKS; r++) { free(jobs[r].bins); free(jobs[r].dc); free(jobs[r].ac); }
  }
  printf("slab_comm_test OK\n");
  return 0;
}
