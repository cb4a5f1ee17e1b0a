/* dctz_host.c -- DCTZ's public C API (include/dctz_compat.h) on top of the GPU hot path.
 *
 * This is the host half of the drop-in: it keeps what the reference keeps on the CPU (allocation,
 * the three zlib streams on their own threads, header + stream assembly, the debug dump files, the
 * stdout lines its test scripts grep) and hands everything SURVEY.md §8 calls the hot path to
 * libdctz_gpu.so through the C-ABI of include/dctz_gpu.h:
 *
 *   dctz_compress    replaces dctz-comp-lib.c:186-217, 271-281, 318-544 by ONE dctz_gpu_compress_core call
 *   dctz_decompress  replaces dctz-decomp-lib.c:358-511 by ONE dctz_gpu_decompress_core call
 *
 * The compressed stream is the reference's: `struct header` (56 bytes) | zlib(bin_index[N]) |
 * zlib(DC[nblk] float) | zlib(AC_exact[n] float) | QT build only: raw qtable[64]
 * (dctz-comp-lib.c:775-820), so either side can decode the other's output.
 * Build with -DUSE_QTABLE for the "qt" flavour, exactly like the reference's Makefile:12-17.
 * Error convention of the reference: message on stderr, exit(1).  No CPU fallback exists.
 */
#include <fcntl.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

#include "../../../include/dctz_compat.h"
#include "../../../include/dctz_gpu.h"

#ifdef USE_QTABLE
#define MODE_QT 1
#else
#define MODE_QT 0
#endif

static dctz_gpu_ctx *g_ctx = NULL;
static int g_device = -1;

static void die(const char *what, const char *detail) {
  fprintf(stderr, "dctz: %s%s%s\n", what, detail ? ": " : "", detail ? detail : "");
  exit(1);
}

void dctz_set_device(int device) {
  if (g_ctx && device != g_device) { dctz_gpu_destroy(g_ctx); g_ctx = NULL; }
  g_device = device;
}

int dctz_build_is_qt(void) { return MODE_QT; }

/* extension: stage timers / PCIe bytes of the GPU call inside the last dctz_compress or dctz_decompress
 * (dctz_gpu_last_call_stats of the library's own context) */
int dctz_host_last_call_stats(double times_ms[8], unsigned long long *h2d_bytes, unsigned long long *d2h_bytes) {
  uint64_t h = 0, d = 0;
  if (!g_ctx) return -1;
  if (dctz_gpu_last_call_stats(g_ctx, times_ms, &h, &d) != DCTZ_GPU_OK) return -1;
  if (h2d_bytes) *h2d_bytes = h;
  if (d2h_bytes) *d2h_bytes = d;
  return 0;
}

static dctz_gpu_ctx *gpu(void) {
  if (!g_ctx) {
    if (g_device < 0) {
      const char *e = getenv("DCTZ_GPU_DEVICE");
      g_device = e ? atoi(e) : 0;
    }
    if (dctz_gpu_create(&g_ctx, g_device) != DCTZ_GPU_OK) die("cannot open the GPU", dctz_gpu_last_error(NULL));
  }
  return g_ctx;
}

static void *xmalloc(size_t n, const char *name) {
  void *p = malloc(n ? n : 1);
  if (!p) die("Out of memory", name);
  return p;
}

static int dumps_enabled(void) {
  const char *e = getenv("DCTZ_NO_DUMPS");
  return !(e && *e && *e != '0');
}

static void dump(const char *name, const void *p, size_t bytes) {
  FILE *f = fopen(name, "wb");
  if (!f) return; /* the reference does not check either (dctz-comp-lib.c:586-588) */
  if (bytes) fwrite(p, bytes, 1, f);
  fclose(f);
}

/* ---- zlib sections --------------------------------------------------------------------------------
 * The reference deflates the three sections on three threads (dctz-comp-lib.c:620-706) and inflates them
 * one after the other (dctz-decomp-lib.c:244-322).  Once the GPU has taken the hot path, that host zlib
 * work is essentially the whole wall-clock of dctz_compress, so the drop-in does
 * better without touching the stream format:
 *   - deflate: every section is cut into chunks that are compressed concurrently on all host cores as raw
 *     deflate blocks (each primed with the last 32 KiB of its predecessor as dictionary, ended with a sync
 *     flush so it finishes on a byte boundary) and concatenated behind one zlib header, the Adler-32 of the
 *     whole section appended -- ONE valid zlib stream per section, inflated by any zlib (the reference's
 *     inflate() included), at most a few bytes per chunk larger than the single-threaded result.
 *     Sections up to DCTZ_Z_SERIAL bytes use the reference's exact single-stream call (byte-identical).
 *   - inflate: the three sections are inflated concurrently (a zlib stream cannot be split).
 * DCTZ_ZLIB_THREADS=<n> overrides the worker count; 1 reproduces the reference byte for byte.
 */
#define DCTZ_Z_CHUNK ((size_t)1 << 20)
#define DCTZ_Z_CHUNK_FLOAT ((size_t)1 << 17)
#define DCTZ_Z_SERIAL ((size_t)1 << 21)
/* Section 0 (bin_index: long runs, deflates at several hundred MB/s per core) is cut into 1 MiB chunks; sections 1 and 2
 * (DC, AC_exact: float data, ~30 MB/s per core at the default level) into 128 KiB chunks, or the eight 1 MiB chunks of an
 * 8 MiB DC section would each keep one core busy for 40 ms while the others idle.  A chunk costs ~100 bytes of stream. */
static size_t zchunk_size(int section) { return section == 0 ? DCTZ_Z_CHUNK : DCTZ_Z_CHUNK_FLOAT; }

typedef struct {
  const void *src;
  size_t n_src;
  unsigned char *dst;
  size_t cap, n_dst;
  int inflate_mode, rc;
} zjob;

static void *zjob_run(void *arg) {
  zjob *j = (zjob *)arg;
  const size_t piece = (size_t)1 << 30; /* zlib counts avail_in/avail_out in 32 bits */
  const unsigned char *src = (const unsigned char *)j->src;
  size_t in_left = j->n_src, out_left = j->cap;
  z_stream s;
  memset(&s, 0, sizeof s);
  if (j->inflate_mode) j->rc = inflateInit(&s);
  else j->rc = deflateInit2(&s, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15, 8, Z_DEFAULT_STRATEGY); /* dctz-comp-lib.c:642-643 */
  if (j->rc != Z_OK) return NULL;
  s.data_type = Z_UNKNOWN;
  s.next_in = (Bytef *)src;
  s.next_out = j->dst;
  for (;;) {
    if (s.avail_in == 0 && in_left) { s.avail_in = (uInt)(in_left < piece ? in_left : piece); in_left -= s.avail_in; }
    if (s.avail_out == 0 && out_left) { s.avail_out = (uInt)(out_left < piece ? out_left : piece); out_left -= s.avail_out; }
    j->rc = j->inflate_mode ? inflate(&s, in_left ? Z_NO_FLUSH : Z_FINISH) : deflate(&s, in_left ? Z_NO_FLUSH : Z_FINISH);
    if (j->rc == Z_OK) continue;
    if (j->rc == Z_BUF_ERROR && ((s.avail_in == 0 && in_left) || (s.avail_out == 0 && out_left))) continue;
    break; /* Z_STREAM_END, or an error reported by run_zjobs */
  }
  j->n_dst = j->cap - out_left - s.avail_out;
  if (j->inflate_mode) inflateEnd(&s); else deflateEnd(&s);
  return NULL;
}

static void run_zjobs(zjob *jobs, int n) {
  pthread_t th[3];
  int i;
  for (i = 0; i < n; i++)
    if (pthread_create(&th[i], NULL, zjob_run, &jobs[i])) die("Error creating thread", NULL);
  for (i = 0; i < n; i++) pthread_join(th[i], NULL);
  for (i = 0; i < n; i++)
    if (jobs[i].rc != Z_STREAM_END) die("zlib stream error", jobs[i].inflate_mode ? "inflate" : "deflate");
}

/* chunk-parallel deflate */
typedef struct {
  const unsigned char *src; /* start of the section */
  size_t begin, len;        /* this chunk */
  int last;
  unsigned char *dst;
  size_t cap, n_dst;
  uLong adler;
  int rc;
} zchunk;

typedef struct {
  zchunk *chunks;
  size_t n;
  size_t next; /* work queue cursor */
  pthread_mutex_t mu;
} zqueue;

static void zchunk_run(zchunk *c) {
  z_stream s;
  memset(&s, 0, sizeof s);
  c->rc = deflateInit2(&s, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY); /* raw deflate */
  if (c->rc != Z_OK) return;
  if (c->begin) { /* prime the window with the 32 KiB that precede the chunk */
    const size_t d = c->begin < 32768 ? c->begin : 32768;
    deflateSetDictionary(&s, c->src + c->begin - d, (uInt)d);
  }
  s.next_in = (Bytef *)(c->src + c->begin);
  s.avail_in = (uInt)c->len;
  s.next_out = c->dst;
  s.avail_out = (uInt)c->cap;
  c->rc = deflate(&s, c->last ? Z_FINISH : Z_SYNC_FLUSH);
  if (c->rc == Z_STREAM_END || (c->rc == Z_OK && !c->last && s.avail_in == 0 && s.avail_out != 0)) c->rc = Z_OK;
  else c->rc = Z_BUF_ERROR;
  c->n_dst = c->cap - s.avail_out;
  deflateEnd(&s);
  c->adler = adler32(adler32(0L, Z_NULL, 0), c->src + c->begin, (uInt)c->len);
}

static void *zworker(void *arg) {
  zqueue *q = (zqueue *)arg;
  for (;;) {
    size_t i;
    pthread_mutex_lock(&q->mu);
    i = q->next++;
    pthread_mutex_unlock(&q->mu);
    if (i >= q->n) return NULL;
    zchunk_run(&q->chunks[i]);
  }
}

static int zlib_threads(void) {
  const char *e = getenv("DCTZ_ZLIB_THREADS");
  long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
  if (n < 1) n = 1;
  if (n > 256) n = 256;
  return (int)n;
}

/* Deflate the three sections into freshly allocated buffers (jobs[i].dst / n_dst). */
static void deflate_sections(zjob *jobs, int nsec) {
  const int nthreads = zlib_threads();
  size_t nchunks = 0, k = 0, c0[3];
  zqueue q;
  pthread_t *th;
  int i, t, started = 0;
  int parallel[3];
  for (i = 0; i < nsec; i++) {
    parallel[i] = nthreads > 1 && jobs[i].n_src > DCTZ_Z_SERIAL;
    c0[i] = nchunks;
    if (parallel[i]) nchunks += (jobs[i].n_src + zchunk_size(i) - 1) / zchunk_size(i);
  }
  if (nchunks == 0) { /* small sections: the reference's own three single-stream calls */
    for (i = 0; i < nsec; i++) {
      jobs[i].cap = compressBound((uLong)jobs[i].n_src);
      jobs[i].dst = (unsigned char *)xmalloc(jobs[i].cap, "zlib output");
    }
    run_zjobs(jobs, nsec);
    return;
  }
  q.chunks = (zchunk *)xmalloc(nchunks * sizeof(zchunk), "zlib chunks");
  q.n = nchunks;
  q.next = 0;
  pthread_mutex_init(&q.mu, NULL);
  for (i = 0; i < nsec; i++) {
    size_t off;
    if (!parallel[i]) continue;
    for (off = 0; off < jobs[i].n_src; off += zchunk_size(i), k++) {
      zchunk *c = &q.chunks[k];
      c->src = (const unsigned char *)jobs[i].src;
      c->begin = off;
      c->len = jobs[i].n_src - off < zchunk_size(i) ? jobs[i].n_src - off : zchunk_size(i);
      c->last = (off + c->len == jobs[i].n_src);
      c->cap = compressBound((uLong)c->len) + 16;
      c->dst = (unsigned char *)xmalloc(c->cap, "zlib chunk");
    }
  }
  th = (pthread_t *)xmalloc((size_t)nthreads * sizeof(pthread_t), "threads");
  for (t = 0; t < nthreads && (size_t)t < nchunks; t++, started++)
    if (pthread_create(&th[t], NULL, zworker, &q)) die("Error creating thread", NULL);
  /* the small sections meanwhile, on this thread */
  for (i = 0; i < nsec; i++) {
    if (parallel[i]) continue;
    jobs[i].cap = compressBound((uLong)jobs[i].n_src);
    jobs[i].dst = (unsigned char *)xmalloc(jobs[i].cap, "zlib output");
    zjob_run(&jobs[i]);
    if (jobs[i].rc != Z_STREAM_END) die("zlib stream error", "deflate");
  }
  for (t = 0; t < started; t++) pthread_join(th[t], NULL);
  free(th);
  pthread_mutex_destroy(&q.mu);
  /* stitch: zlib header | raw blocks ... | Adler-32 (big endian) */
  for (i = 0; i < nsec; i++) {
    size_t n = (jobs[i].n_src + zchunk_size(i) - 1) / zchunk_size(i), total = 2 + 4, j;
    uLong ad = adler32(0L, Z_NULL, 0);
    unsigned char *o;
    if (!parallel[i]) continue;
    for (j = 0; j < n; j++) {
      if (q.chunks[c0[i] + j].rc != Z_OK) die("zlib stream error", "parallel deflate");
      total += q.chunks[c0[i] + j].n_dst;
    }
    o = jobs[i].dst = (unsigned char *)xmalloc(total, "zlib output");
    *o++ = 0x78; *o++ = 0x9C; /* deflate, 32 KiB window, default level, no preset dictionary */
    for (j = 0; j < n; j++) {
      zchunk *c = &q.chunks[c0[i] + j];
      memcpy(o, c->dst, c->n_dst);
      o += c->n_dst;
      ad = adler32_combine(ad, c->adler, (z_off_t)c->len);
      free(c->dst);
    }
    *o++ = (unsigned char)(ad >> 24); *o++ = (unsigned char)(ad >> 16); *o++ = (unsigned char)(ad >> 8); *o++ = (unsigned char)ad;
    jobs[i].n_dst = total;
    jobs[i].rc = Z_STREAM_END;
  }
  free(q.chunks);
}

/* ---- deflate that starts while the sections are still arriving from the GPU ---------------------------------------
 * dctz_gpu_compress_core_cb reports every piece of bin_index / DC / AC_exact as it lands in host memory; zpipe turns
 * the pieces of a large section into the same 1 MiB chunk jobs deflate_sections uses and feeds them to worker threads
 * at once, so the deflate of the first pieces overlaps the PCIe transfer of the later ones and the host-side scaling
 * of the caller's buffer.  The result is byte-identical to deflate_sections on the complete arrays. */
typedef struct {
  pthread_mutex_t mu;
  pthread_cond_t cv;
  const unsigned char *src[3];
  size_t total[3], ready[3], queued[3], nchunks[3];
  zchunk *chunks[3];
  int known[3], parallel[3];
  size_t *fifo; /* (section << 56) | chunk index */
  size_t head, tail, cap;
  int closing, nthreads, started;
  pthread_t *th;
} zpipe;

static void *zpipe_worker(void *arg) {
  zpipe *z = (zpipe *)arg;
  for (;;) {
    size_t item;
    pthread_mutex_lock(&z->mu);
    while (z->head == z->tail && !z->closing) pthread_cond_wait(&z->cv, &z->mu);
    if (z->head == z->tail) { pthread_mutex_unlock(&z->mu); return NULL; }
    item = z->fifo[z->head++];
    pthread_mutex_unlock(&z->mu);
    zchunk_run(&z->chunks[item >> 56][item & (((size_t)1 << 56) - 1)]);
  }
}

static void zpipe_init(zpipe *z) {
  int t;
  memset(z, 0, sizeof *z);
  pthread_mutex_init(&z->mu, NULL);
  pthread_cond_init(&z->cv, NULL);
  z->nthreads = zlib_threads();
  if (z->nthreads < 2) return; /* DCTZ_ZLIB_THREADS=1: the reference's single-stream calls, after the fact */
  z->th = (pthread_t *)xmalloc((size_t)z->nthreads * sizeof(pthread_t), "threads");
  for (t = 0; t < z->nthreads; t++, z->started++)
    if (pthread_create(&z->th[t], NULL, zpipe_worker, z)) die("Error creating thread", NULL);
}

/* a section's base pointer and final size become known with its first piece */
static void zpipe_feed(zpipe *z, int sec, const void *base, size_t total, size_t off, size_t len) {
  size_t k;
  pthread_mutex_lock(&z->mu);
  if (!z->known[sec]) {
    z->known[sec] = 1;
    z->src[sec] = (const unsigned char *)base;
    z->total[sec] = total;
    z->parallel[sec] = z->nthreads > 1 && total > DCTZ_Z_SERIAL;
    if (z->parallel[sec]) {
      z->nchunks[sec] = (total + zchunk_size(sec) - 1) / zchunk_size(sec);
      z->chunks[sec] = (zchunk *)xmalloc(z->nchunks[sec] * sizeof(zchunk), "zlib chunks");
      z->fifo = (size_t *)realloc(z->fifo, (z->cap + z->nchunks[sec]) * sizeof(size_t));
      if (!z->fifo) die("Out of memory", "zlib queue");
      z->cap += z->nchunks[sec];
    }
  }
  if (off + len > z->ready[sec]) z->ready[sec] = off + len;
  if (z->parallel[sec]) {
    for (k = z->queued[sec]; k < z->nchunks[sec]; k++) { /* every chunk whose bytes are all there */
      const size_t cs = zchunk_size(sec), begin = k * cs, end = begin + cs < z->total[sec] ? begin + cs : z->total[sec];
      zchunk *c = &z->chunks[sec][k];
      if (end > z->ready[sec]) break;
      c->src = z->src[sec];
      c->begin = begin;
      c->len = end - begin;
      c->last = (end == z->total[sec]);
      c->cap = compressBound((uLong)c->len) + 16;
      c->dst = (unsigned char *)xmalloc(c->cap, "zlib chunk");
      z->fifo[z->tail++] = ((size_t)sec << 56) | k;
    }
    z->queued[sec] = k;
    pthread_cond_broadcast(&z->cv);
  }
  pthread_mutex_unlock(&z->mu);
}

/* all pieces have been fed: finish the three sections into jobs[i].dst / n_dst (same contract as deflate_sections) */
static void zpipe_finish(zpipe *z, zjob *jobs, int nsec) {
  int i, t;
  pthread_mutex_lock(&z->mu);
  z->closing = 1;
  pthread_cond_broadcast(&z->cv);
  pthread_mutex_unlock(&z->mu);
  for (i = 0; i < nsec; i++) { /* the small sections meanwhile, on this thread: the reference's own single-stream call */
    if (z->known[i] && z->parallel[i]) continue;
    jobs[i].cap = compressBound((uLong)jobs[i].n_src);
    jobs[i].dst = (unsigned char *)xmalloc(jobs[i].cap, "zlib output");
    zjob_run(&jobs[i]);
    if (jobs[i].rc != Z_STREAM_END) die("zlib stream error", "deflate");
  }
  for (t = 0; t < z->started; t++) pthread_join(z->th[t], NULL);
  for (i = 0; i < nsec; i++) {
    size_t total = 2 + 4, j;
    uLong ad = adler32(0L, Z_NULL, 0);
    unsigned char *o;
    if (!(z->known[i] && z->parallel[i])) continue;
    if (z->queued[i] != z->nchunks[i] || z->total[i] != jobs[i].n_src) die("internal error", "a stream section was not delivered completely");
    for (j = 0; j < z->nchunks[i]; j++) {
      if (z->chunks[i][j].rc != Z_OK) die("zlib stream error", "parallel deflate");
      total += z->chunks[i][j].n_dst;
    }
    o = jobs[i].dst = (unsigned char *)xmalloc(total, "zlib output");
    *o++ = 0x78; *o++ = 0x9C;
    for (j = 0; j < z->nchunks[i]; j++) {
      zchunk *c = &z->chunks[i][j];
      memcpy(o, c->dst, c->n_dst);
      o += c->n_dst;
      ad = adler32_combine(ad, c->adler, (z_off_t)c->len);
      free(c->dst);
    }
    *o++ = (unsigned char)(ad >> 24); *o++ = (unsigned char)(ad >> 16); *o++ = (unsigned char)(ad >> 8); *o++ = (unsigned char)ad;
    jobs[i].n_dst = total;
    jobs[i].rc = Z_STREAM_END;
    free(z->chunks[i]);
  }
  free(z->fifo);
  free(z->th);
  pthread_mutex_destroy(&z->mu);
  pthread_cond_destroy(&z->cv);
}

/* extension, used by the CPU tests: deflate one section exactly the way dctz_compress does.
 * Returns the compressed size, or 0 if `cap` is too small. */
size_t dctz_host_deflate(const void *src, size_t n, void *dst, size_t cap) {
  zjob j;
  size_t out;
  memset(&j, 0, sizeof j);
  j.src = src;
  j.n_src = n;
  deflate_sections(&j, 1);
  out = j.n_dst <= cap ? j.n_dst : 0;
  if (out) memcpy(dst, j.dst, out);
  free(j.dst);
  return out;
}

/* extension, used by the CPU tests: section number `section` (0 bin_index, 1 DC, 2 AC_exact: they differ in chunk size)
 * deflated from the complete array (piece == 0) or while it arrives in pieces of `piece` bytes. */
size_t dctz_host_deflate_section(const void *src, size_t n, void *dst, size_t cap, int section, size_t piece) {
  zjob j[3];
  size_t out;
  if (section < 0 || section > 2) return 0;
  memset(j, 0, sizeof j);
  j[section].src = src;
  j[section].n_src = n;
  if (piece == 0) {
    deflate_sections(j, section + 1);
  } else {
    zpipe z;
    size_t off;
    zpipe_init(&z);
    for (off = 0; off < n; off += piece) zpipe_feed(&z, section, src, n, off, n - off < piece ? n - off : piece);
    if (n == 0) zpipe_feed(&z, section, src, 0, 0, 0);
    zpipe_finish(&z, j, section + 1);
  }
  out = j[section].n_dst <= cap ? j[section].n_dst : 0;
  if (out) memcpy(dst, j[section].dst, out);
  for (int i = 0; i <= section; i++) free(j[i].dst);
  return out;
}

/* extension, used by the CPU tests: the same section deflated while it "arrives" in pieces of `piece` bytes (zpipe) */
size_t dctz_host_deflate_streamed(const void *src, size_t n, void *dst, size_t cap, size_t piece) {
  zpipe z;
  zjob j;
  size_t off, out;
  memset(&j, 0, sizeof j);
  j.src = src;
  j.n_src = n;
  zpipe_init(&z);
  if (piece == 0) piece = n ? n : 1;
  for (off = 0; off < n; off += piece) zpipe_feed(&z, 0, src, n, off, n - off < piece ? n - off : piece);
  if (n == 0) zpipe_feed(&z, 0, src, 0, 0, 0);
  zpipe_finish(&z, &j, 1);
  out = j.n_dst <= cap ? j.n_dst : 0;
  if (out) memcpy(dst, j.dst, out);
  free(j.dst);
  return out;
}

static double wall(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static int env_on(const char *name) {
  const char *e = getenv(name);
  return e && *e && *e != '0';
}

typedef struct {
  const void *qtable_raw, *bin_index, *ac;
  size_t qbytes, n, ac_bytes;
} dump_job;

static void *dump_worker(void *arg) {
  dump_job *d = (dump_job *)arg;
  if (d->qtable_raw) dump("qtable.bin", d->qtable_raw, d->qbytes);
  dump("bin_index.bin", d->bin_index, d->n);
  dump("AC_exact.bin", d->ac, d->ac_bytes);
  return NULL;
}

/* ---- stream assembly (dctz-comp-lib.c:583-846): side files, three deflates, header, concatenation ---- */
static size_t assemble_stream(t_datatype dt, size_t n, double error_bound, const dctz_gpu_info *info, const t_bin_id *bin_index,
                              const float *DC, const float *AC_exact, const void *qtable, const void *qtable_raw, unsigned char *out,
                              int write_dumps, zpipe *zp) {
  const int is_double = (dt == DOUBLE);
  const size_t es = is_double ? sizeof(double) : sizeof(float);
  const size_t nblk = (n + DCTZ_BLK_SZ - 1) / DCTZ_BLK_SZ, qbytes = MODE_QT ? DCTZ_BLK_SZ * es : 0;
  zjob jobs[3];
  struct header h;
  size_t total;
  int i;
  dump_job dj;
  pthread_t dump_th;
  int dumping = 0;
  double t_a = wall(), t_b;
  if (info->n_outliers > 0xFFFFFFFFull) die("too many outliers for the stream header", NULL);
  if (write_dumps && dumps_enabled()) { /* dctz-comp-lib.c:443-448, 583-595: side files the reference's scripts rename;
                                           written by their own thread while the sections are deflated */
    dj.qtable_raw = MODE_QT ? qtable_raw : NULL; dj.qbytes = qbytes;
    dj.bin_index = bin_index; dj.n = n;
    dj.ac = AC_exact; dj.ac_bytes = (size_t)info->n_outliers * sizeof(float);
    dumping = !pthread_create(&dump_th, NULL, dump_worker, &dj);
    if (!dumping) dump_worker(&dj);
  }
  memset(jobs, 0, sizeof jobs);
  jobs[0].src = bin_index; jobs[0].n_src = n;
  jobs[1].src = DC;        jobs[1].n_src = nblk * sizeof(float);
  jobs[2].src = AC_exact;  jobs[2].n_src = (size_t)info->n_outliers * sizeof(float);
  if (zp) zpipe_finish(zp, jobs, 3); /* most of the work is done already */
  else deflate_sections(jobs, 3);

  memset(&h, 0, sizeof h); /* the reference leaves padding uninitialised; zero is as valid and reproducible */
  h.datatype = dt;
  h.num_elements = (unsigned int)n;
  h.error_bound = error_bound;
  h.tot_AC_exact_count = (unsigned int)info->n_outliers;
  if (is_double) { h.scaling_factor.d = info->sf; h.mean.d = info->mean; }
  else { h.scaling_factor.f = (float)info->sf; h.mean.f = (float)info->mean; }
  h.bindex_sz_compressed = (unsigned int)jobs[0].n_dst;
  h.DC_sz_compressed = (unsigned int)jobs[1].n_dst;
  h.AC_exact_sz_compressed = (unsigned int)jobs[2].n_dst;
#ifdef USE_QTABLE
  h.bindex_count = (unsigned int)n;
#endif
  total = sizeof h + jobs[0].n_dst + jobs[1].n_dst + jobs[2].n_dst + qbytes;
  memcpy(out, &h, sizeof h);
  out += sizeof h;
  for (i = 0; i < 3; i++) {
    memcpy(out, jobs[i].dst, jobs[i].n_dst);
    out += jobs[i].n_dst;
    free(jobs[i].dst);
  }
  if (MODE_QT) memcpy(out, qtable, qbytes);
  t_b = wall();
  if (dumping) pthread_join(dump_th, NULL);
  if (env_on("DCTZ_PROFILE")) fprintf(stderr, "dctz profile: deflate+assemble %.1f ms, waiting for the side files %.1f ms\n", 1e3 * (t_b - t_a), 1e3 * (wall() - t_b));
  return total;
}

/* ---- side files written while the sections arrive -----------------------------------------------------------------
 * bin_index.bin / AC_exact.bin (dctz-comp-lib.c:583-595) are as large as the sections themselves; one thread writing them
 * after the GPU call was the longest stage of dctz_compress.  dpipe writes every piece of a section to its place in the
 * file (pwrite) as soon as it has landed, on a few threads of its own. */
#define DPIPE_THREADS 4
#define DPIPE_PIECE ((size_t)4 << 20)
typedef struct { int fd; const unsigned char *p; size_t off, len; } ditem;
typedef struct {
  pthread_mutex_t mu;
  pthread_cond_t cv;
  ditem *items;
  size_t head, tail, cap;
  int closing, started, opened, fd[3];
  pthread_t th[DPIPE_THREADS];
} dpipe;

static void dpipe_open(dpipe *d) { /* (truncating last call's files costs milliseconds: done by a worker, not by the caller) */
  const int f0 = open("bin_index.bin", O_WRONLY | O_CREAT | O_TRUNC, 0666);
  const int f2 = open("AC_exact.bin", O_WRONLY | O_CREAT | O_TRUNC, 0666);
  pthread_mutex_lock(&d->mu);
  d->fd[0] = f0; d->fd[1] = -1; d->fd[2] = f2; /* DC has no side file */
  d->opened = 1;
  pthread_cond_broadcast(&d->cv);
  pthread_mutex_unlock(&d->mu);
}

static void *dpipe_worker(void *arg) {
  dpipe *d = (dpipe *)arg;
  int first;
  pthread_mutex_lock(&d->mu);
  first = (d->opened == 0);
  if (first) d->opened = -1; /* being opened */
  pthread_mutex_unlock(&d->mu);
  if (first) dpipe_open(d);
  for (;;) {
    ditem it;
    pthread_mutex_lock(&d->mu);
    while ((d->opened != 1 || d->head == d->tail) && !(d->closing && d->opened == 1 && d->head == d->tail)) pthread_cond_wait(&d->cv, &d->mu);
    if (d->head == d->tail) { pthread_mutex_unlock(&d->mu); return NULL; }
    it = d->items[d->head++];
    it.fd = d->fd[it.fd]; /* the item carries the section number */
    pthread_mutex_unlock(&d->mu);
    while (it.fd >= 0 && it.len) { /* (errors are ignored like the reference's unchecked fwrite) */
      const ssize_t w = pwrite(it.fd, it.p, it.len, (off_t)it.off);
      if (w <= 0) break;
      it.p += w; it.off += (size_t)w; it.len -= (size_t)w;
    }
  }
}

static void dpipe_init(dpipe *d, size_t total_bytes) {
  int t;
  memset(d, 0, sizeof *d);
  d->fd[0] = d->fd[1] = d->fd[2] = -1;
  pthread_mutex_init(&d->mu, NULL);
  pthread_cond_init(&d->cv, NULL);
  d->cap = total_bytes / DPIPE_PIECE + 4096;
  d->items = (ditem *)xmalloc(d->cap * sizeof(ditem), "side-file queue");
  for (t = 0; t < DPIPE_THREADS; t++, d->started++)
    if (pthread_create(&d->th[t], NULL, dpipe_worker, d)) break;
}

static void dpipe_feed(dpipe *d, int sec, const void *base, size_t off, size_t len) {
  if (sec == 1 || len == 0) return;
  pthread_mutex_lock(&d->mu);
  while (len && d->tail < d->cap) { /* (the queue holds every piece of both files: see dpipe_init) */
    const size_t l = len < DPIPE_PIECE ? len : DPIPE_PIECE;
    ditem *it = &d->items[d->tail++];
    it->fd = sec; it->p = (const unsigned char *)base + off; it->off = off; it->len = l;
    off += l; len -= l;
  }
  pthread_cond_broadcast(&d->cv);
  pthread_mutex_unlock(&d->mu);
}

static void dpipe_finish(dpipe *d) {
  int t;
  pthread_mutex_lock(&d->mu);
  d->closing = 1;
  pthread_cond_broadcast(&d->cv);
  pthread_mutex_unlock(&d->mu);
  if (d->started == 0) dpipe_worker(d); /* no thread could be created: open, drain the queue and return here */
  for (t = 0; t < d->started; t++) pthread_join(d->th[t], NULL);
  for (t = 0; t < 3; t++) if (d->fd[t] >= 0) close(d->fd[t]);
  free(d->items);
  pthread_mutex_destroy(&d->mu);
  pthread_cond_destroy(&d->cv);
}

/* ---- dctz_compress (dctz.h:126) ------------------------------------------------------------------ */
typedef struct {
  zpipe *zp;
  dpipe *dp; /* NULL: no side files */
  const void *base[3];
  size_t n, nblk;
  const dctz_gpu_info *info;
} feed_ctx;

static void on_section(void *user, int section, size_t off, size_t bytes) {
  feed_ctx *f = (feed_ctx *)user;
  const size_t total = section == 0 ? f->n : section == 1 ? f->nblk * sizeof(float) : (size_t)f->info->n_outliers * sizeof(float);
  zpipe_feed(f->zp, section, f->base[section], total, off, bytes);
  if (f->dp) dpipe_feed(f->dp, section, f->base[section], off, bytes);
}

/* -DDCT_FILE_DEBUG of the reference (dctz-comp-lib.c:422-433): the coefficients of the scaled data and the DC array.
 * The fused kernels never materialise the coefficients, so this debug dump transforms the (by now scaled) input once
 * more through the DCT-only entry point.  DCTZ_DCT_FILE_DEBUG=1 enables it. */
static void dump_coefficients(const t_var *var, size_t n, const float *DC, size_t nblk) {
  const int is_double = (var->datatype == DOUBLE);
  const size_t es = is_double ? sizeof(double) : sizeof(float), nfull = n / DCTZ_BLK_SZ, rem = n % DCTZ_BLK_SZ;
  const unsigned char *src = is_double ? (const unsigned char *)var->buf.d : (const unsigned char *)var->buf.f;
  unsigned char *coef = (unsigned char *)xmalloc(n * es, "dct_result");
  const int code = is_double ? DCTZ_GPU_DOUBLE : DCTZ_GPU_FLOAT;
  if (nfull && dctz_gpu_dct_blocks(gpu(), src, coef, nfull, DCTZ_BLK_SZ, code, 0) != DCTZ_GPU_OK) die("GPU DCT failed", dctz_gpu_last_error(g_ctx));
  if (rem && dctz_gpu_dct_blocks(gpu(), src + nfull * DCTZ_BLK_SZ * es, coef + nfull * DCTZ_BLK_SZ * es, 1, (int)rem, code, 0) != DCTZ_GPU_OK)
    die("GPU DCT failed", dctz_gpu_last_error(g_ctx));
  dump("dct_result.bin", coef, n * es);
  dump("DC.bin", DC, nblk * sizeof(float));
  free(coef);
}

int dctz_compress(t_var *var, int N, size_t *outSize, t_var *var_z, double error_bound) {
  const int is_double = (var->datatype == DOUBLE);
  const int timing = env_on("DCTZ_TIME_DEBUG");
  size_t n, nblk;
  t_bin_id *bin_index;
  float *DC, *AC_exact;
  unsigned char qtable[DCTZ_BLK_SZ * sizeof(double)], qtable_raw[DCTZ_BLK_SZ * sizeof(double)];
  dctz_gpu_info info;
  zpipe zp;
  dpipe dp;
  int piped_dumps = 0;
  feed_ctx fc;
  double t0 = wall(), t1, t2;
  void *data = is_double ? (void *)var->buf.d : (void *)var->buf.f;
  int rc;

  if (error_bound < 1E-6) { /* dctz-comp-lib.c:135-138 */
    fprintf(stderr, "ERROR: error bound should be no less than 1E-6.\n");
    exit(1);
  }
  if (N <= 0) die("nothing to compress", "N <= 0");
  n = (size_t)N;
  nblk = (n + DCTZ_BLK_SZ - 1) / DCTZ_BLK_SZ;
  bin_index = (t_bin_id *)xmalloc(n, "bin_index");
  DC = (float *)xmalloc(nblk * sizeof(float), "DC");
  AC_exact = (float *)xmalloc(n * sizeof(float), "AC_exact");

  /* the whole hot path: statistics, scaling (left in the caller's buffer like dctz-comp-lib.c:198,213),
   * block DCT, binning quantiser, ordered outliers, QT table + rescale.  The sections are deflated as they arrive. */
  dctz_gpu_set_timing(gpu(), timing);
  if (!timing) {
    zpipe_init(&zp);
    piped_dumps = dumps_enabled();
    if (piped_dumps) dpipe_init(&dp, n * (1 + sizeof(float)));
    fc.dp = piped_dumps ? &dp : NULL;
    fc.zp = &zp; fc.base[0] = bin_index; fc.base[1] = DC; fc.base[2] = AC_exact; fc.n = n; fc.nblk = nblk; fc.info = &info;
    rc = dctz_gpu_compress_core_cb(g_ctx, data, n, is_double ? DCTZ_GPU_DOUBLE : DCTZ_GPU_FLOAT, error_bound, MODE_QT, data, bin_index, DC,
                                   AC_exact, MODE_QT ? qtable : NULL, MODE_QT ? qtable_raw : NULL, &info, on_section, &fc);
  } else { /* the reference's -DTIME_DEBUG lines: every stage on its own, nothing overlapped */
    rc = dctz_gpu_compress_core(g_ctx, data, n, is_double ? DCTZ_GPU_DOUBLE : DCTZ_GPU_FLOAT, error_bound, MODE_QT, data, bin_index, DC,
                                AC_exact, MODE_QT ? qtable : NULL, MODE_QT ? qtable_raw : NULL, &info);
  }
  if (rc != DCTZ_GPU_OK) die("GPU compress failed", dctz_gpu_last_error(g_ctx));
  t1 = wall();
  if (env_on("DCTZ_PROFILE")) {
    double ms[8];
    dctz_gpu_last_call_stats(g_ctx, ms, NULL, NULL);
    fprintf(stderr, "dctz profile: GPU call %.1f ms (upload + kernels %.1f ms, downloads + host scaling %.1f ms)\n", 1e3 * (t1 - t0), ms[3], ms[4]);
  }
  if (env_on("DCTZ_DCT_FILE_DEBUG") && dumps_enabled()) dump_coefficients(var, n, DC, nblk);
  *outSize = assemble_stream(var->datatype, n, error_bound, &info, bin_index, DC, AC_exact, qtable, qtable_raw,
                             is_double ? (unsigned char *)var_z->buf.d : (unsigned char *)var_z->buf.f, !piped_dumps, timing ? NULL : &zp);
  if (piped_dumps) {
    const double t_d = wall();
    if (MODE_QT) dump("qtable.bin", qtable_raw, DCTZ_BLK_SZ * (is_double ? sizeof(double) : sizeof(float))); /* dctz-comp-lib.c:443-448 */
    dpipe_finish(&dp);
    if (env_on("DCTZ_PROFILE")) fprintf(stderr, "dctz profile: waiting for the side files %.1f ms\n", 1e3 * (wall() - t_d));
  }
  t2 = wall();
  free(bin_index);
  free(DC);
  free(AC_exact);
  if (timing) { /* same lines as dctz-comp-lib.c:762-773; sf_t / dct_t are CUDA-event times of the two kernel stages */
    double ms[8];
    dctz_gpu_last_call_stats(g_ctx, ms, NULL, NULL);
    printf("sf_t=%f(s), dct_t=%f(s), zlib_t(compress)=%f(s)\n", ms[1] / 1e3, ms[2] / 1e3, t2 - t1);
    printf("h2d_t=%f(s), d2h_scale_t=%f(s)\n", ms[0] / 1e3, ms[4] / 1e3);
    printf("comp_time = %f (s), compression rate = %f (MB/s)\n", t2 - t0, ((double)n * sizeof(double) / (1024.0 * 1024.0)) / (t2 - t0));
  }
  printf("outSize = %zu\n", *outSize);
  return 1;
}

/* ---- dctz_decompress (dctz.h:127) ---------------------------------------------------------------- */
static double g_inflate_s = 0.0; /* wall clock of the three inflates of the last dctz_decompress call */

/* decode one standard stream at `p` into `out` (room for `out_cap` elements; 0 = the caller vouches for the room, as
 * dctz_decompress's interface does); returns the stream's element count */
static size_t decode_stream(dctz_gpu_ctx *ctx, t_datatype dt, const unsigned char *p, void *out, size_t out_cap, int chatty) {
  const int is_double = (dt == DOUBLE);
  const size_t es = is_double ? sizeof(double) : sizeof(float);
  struct header h;
  size_t n, nblk, n_out;
  t_bin_id *bin_index;
  float *DC, *AC_exact;
  unsigned char qtable[DCTZ_BLK_SZ * sizeof(double)];
  zjob jobs[3];
  double sf;

  memcpy(&h, p, sizeof h); /* dctz-decomp-lib.c:84-100 */
  p += sizeof h;
  n = h.num_elements;
  nblk = (n + DCTZ_BLK_SZ - 1) / DCTZ_BLK_SZ;
  n_out = h.tot_AC_exact_count;
  if (n == 0) die("corrupt stream", "num_elements == 0");
  if (out_cap && n > out_cap) die("corrupt stream", "num_elements exceeds the room left in the output");
  if (h.datatype != dt) die("corrupt stream", "datatype in the header differs from the caller's");
  bin_index = (t_bin_id *)xmalloc(n, "bin_index");
  DC = (float *)xmalloc(nblk * sizeof(float), "DC");
  AC_exact = (float *)xmalloc((n_out ? n_out : 1) * sizeof(float), "AC_exact");

  memset(jobs, 0, sizeof jobs);
  jobs[0].src = p;                              jobs[0].n_src = h.bindex_sz_compressed;   jobs[0].dst = bin_index;                 jobs[0].cap = n;
  jobs[1].src = p + h.bindex_sz_compressed;     jobs[1].n_src = h.DC_sz_compressed;       jobs[1].dst = (unsigned char *)DC;       jobs[1].cap = nblk * sizeof(float);
  jobs[2].src = (const unsigned char *)jobs[1].src + h.DC_sz_compressed;
  jobs[2].n_src = h.AC_exact_sz_compressed;     jobs[2].dst = (unsigned char *)AC_exact;  jobs[2].cap = (n_out ? n_out : 1) * sizeof(float);
  jobs[0].inflate_mode = jobs[1].inflate_mode = jobs[2].inflate_mode = 1;
  {
    const double tz = wall();
    run_zjobs(jobs, 3);
    if (chatty) g_inflate_s = wall() - tz;
  }
  if (jobs[0].n_dst != n || jobs[1].n_dst != nblk * sizeof(float) || jobs[2].n_dst != n_out * sizeof(float))
    die("corrupt stream", "section sizes do not match the header");
  if (chatty) printf("uncompressed bin_index size is: %lu\n", (unsigned long)jobs[0].n_dst);
  if (MODE_QT) memcpy(qtable, (const unsigned char *)jobs[2].src + h.AC_exact_sz_compressed, DCTZ_BLK_SZ * es);

  sf = is_double ? h.scaling_factor.d : (double)h.scaling_factor.f;
  /* dequantise + inverse DCT + de-scale: dctz-decomp-lib.c:358-511 */
  if (dctz_gpu_decompress_core(ctx, bin_index, DC, AC_exact, n_out, MODE_QT ? qtable : NULL, n,
                               is_double ? DCTZ_GPU_DOUBLE : DCTZ_GPU_FLOAT, h.error_bound, sf, MODE_QT, out) != DCTZ_GPU_OK)
    die("GPU decompress failed", dctz_gpu_last_error(ctx));
  free(bin_index);
  free(DC);
  free(AC_exact);
  return n;
}

int dctz_decompress(t_var *var_z, t_var *var_r) {
  const int is_double = (var_z->datatype == DOUBLE);
  const int timing = env_on("DCTZ_TIME_DEBUG");
  const double t0 = wall();
  size_t n;
  dctz_gpu_set_timing(gpu(), timing);
  n = decode_stream(g_ctx, var_z->datatype, is_double ? (const unsigned char *)var_z->buf.d : (const unsigned char *)var_z->buf.f,
                    is_double ? (void *)var_r->buf.d : (void *)var_r->buf.f, 0, 1);
  if (timing) { /* dctz-decomp-lib.c:513-528 */
    double ms[8];
    const double t1 = wall();
    dctz_gpu_last_call_stats(g_ctx, ms, NULL, NULL);
    printf("sf_t=%f(s), idct_t=%f(s), zlib_t(uncompress)=%f(s)\n", 0.0, ms[2] / 1e3, g_inflate_s);
    printf("h2d_t=%f(s), d2h_t=%f(s)\n", ms[0] / 1e3, ms[4] / 1e3);
    printf("decomp_time = %f (s), decompression rate = %f (MB/s)\n", t1 - t0, ((double)n * sizeof(double) / (1024.0 * 1024.0)) / (t1 - t0));
  }
  return 1;
}

/* ---- large fields: several standard streams sharing ONE global scaling factor -----------
 * `dctz_compress` takes `int N` and the header stores 32-bit counts (dctz.h:99-114), so a 2048^3 field (2^33
 * elements) cannot be one stream.  The container below frames it as block-aligned pieces of at most
 * `g_piece` elements; every piece is a complete, standard DCTZ stream whose header carries the GLOBAL scaling
 * factor (the reference's dctz_decompress decodes each of them), compressed by one of DCTZ_GPUS devices:
 *     "DCTZMS01" | u64 N_total | u32 datatype | u32 n_streams | u64 size[n_streams] | stream 0 | stream 1 | ...
 */
static size_t g_piece = (size_t)1 << 30;
void dctz_large_set_piece(size_t elements) { g_piece = elements < 64 ? 64 : (elements / 64) * 64; }

typedef struct {
  int device, ndev, failed;
  t_datatype dt;
  const unsigned char *data;
  size_t N, npieces;
  double eb;
  /* pass 1 */
  double *pmax, *pmin, *psum;
  /* pass 2 */
  const double *stats3;
  unsigned char **streams;
  size_t *sizes;
  /* decode */
  const unsigned char *in;
  const size_t *offs;      /* byte offset of every stream in the container */
  const size_t *elem_offs; /* element offset of every stream in the output (prefix sum of the streams' num_elements) */
  const size_t *elem_cnt;
  unsigned char *out;
} large_job;

static size_t piece_len(const large_job *j, size_t i) {
  const size_t start = i * g_piece;
  return j->N - start < g_piece ? j->N - start : g_piece;
}

static void *large_stats_worker(void *arg) {
  large_job *j = (large_job *)arg;
  const size_t es = j->dt == DOUBLE ? 8 : 4;
  dctz_gpu_ctx *ctx = NULL;
  size_t i;
  if (dctz_gpu_create(&ctx, j->device) != DCTZ_GPU_OK) { j->failed = 1; return NULL; }
  for (i = (size_t)j->device; i < j->npieces; i += (size_t)j->ndev) {
    const unsigned char *p = j->data + i * g_piece * es;
    dctz_gpu_info info;
    if (dctz_gpu_stats(ctx, p, piece_len(j, i), j->dt == DOUBLE ? DCTZ_GPU_DOUBLE : DCTZ_GPU_FLOAT, &info) != DCTZ_GPU_OK) { j->failed = 1; break; }
    j->pmax[i] = info.max_abs;
    j->pmin[i] = info.min_abs;
    j->psum[i] = info.sum + (j->dt == DOUBLE ? *(const double *)p : (double)*(const float *)p); /* the kernel skipped the piece's element 0 */
  }
  dctz_gpu_destroy(ctx);
  return NULL;
}

static void *large_compress_worker(void *arg) {
  large_job *j = (large_job *)arg;
  const size_t es = j->dt == DOUBLE ? 8 : 4;
  dctz_gpu_ctx *ctx = NULL;
  size_t i;
  if (dctz_gpu_create(&ctx, j->device) != DCTZ_GPU_OK) { j->failed = 1; return NULL; }
  for (i = (size_t)j->device; i < j->npieces; i += (size_t)j->ndev) {
    const size_t n = piece_len(j, i), nblk = (n + DCTZ_BLK_SZ - 1) / DCTZ_BLK_SZ;
    t_bin_id *bin_index = (t_bin_id *)xmalloc(n, "bin_index");
    float *DC = (float *)xmalloc(nblk * sizeof(float), "DC"), *AC_exact = (float *)xmalloc(n * sizeof(float), "AC_exact");
    unsigned char qtable[DCTZ_BLK_SZ * sizeof(double)], qtable_raw[DCTZ_BLK_SZ * sizeof(double)];
    dctz_gpu_info info;
    if (dctz_gpu_compress_core_with_stats(ctx, j->data + i * g_piece * es, n, j->N, j->stats3, i == 0, j->dt == DOUBLE ? DCTZ_GPU_DOUBLE : DCTZ_GPU_FLOAT,
                                          j->eb, MODE_QT, NULL, bin_index, DC, AC_exact, MODE_QT ? qtable : NULL, MODE_QT ? qtable_raw : NULL,
                                          &info) != DCTZ_GPU_OK) {
      fprintf(stderr, "dctz: %s\n", dctz_gpu_last_error(ctx));
      j->failed = 1;
    } else {
      j->streams[i] = (unsigned char *)xmalloc(64 + n + n / 8 + nblk * 4 + (size_t)info.n_outliers * 4 + 4096, "stream");
      j->sizes[i] = assemble_stream(j->dt, n, j->eb, &info, bin_index, DC, AC_exact, qtable, qtable_raw, j->streams[i], 0, NULL);
    }
    free(bin_index); free(DC); free(AC_exact);
    if (j->failed) break;
  }
  dctz_gpu_destroy(ctx);
  return NULL;
}

static void *large_decode_worker(void *arg) {
  large_job *j = (large_job *)arg;
  const size_t es = j->dt == DOUBLE ? 8 : 4;
  dctz_gpu_ctx *ctx = NULL;
  size_t i;
  if (dctz_gpu_create(&ctx, j->device) != DCTZ_GPU_OK) { j->failed = 1; return NULL; }
  for (i = (size_t)j->device; i < j->npieces; i += (size_t)j->ndev)
    decode_stream(ctx, j->dt, j->in + j->offs[i], j->out + j->elem_offs[i] * es, j->elem_cnt[i], 0);
  dctz_gpu_destroy(ctx);
  return NULL;
}

static int large_devices(void) {
  const char *e = getenv("DCTZ_GPUS");
  int n = e ? atoi(e) : 1, have = dctz_gpu_device_count();
  if (have < 1) die("cannot open the GPU", "no CUDA device");
  if (n < 1) n = 1;
  return n > have ? have : n;
}

static void run_large(large_job *proto, void *(*fn)(void *)) {
  const int ndev = large_devices();
  large_job jobs[16];
  pthread_t th[16];
  int t, n = ndev > 16 ? 16 : ndev;
  for (t = 0; t < n; t++) {
    jobs[t] = *proto;
    jobs[t].device = t;
    jobs[t].ndev = n;
    jobs[t].failed = 0;
    if (pthread_create(&th[t], NULL, fn, &jobs[t])) die("Error creating thread", NULL);
  }
  for (t = 0; t < n; t++) pthread_join(th[t], NULL);
  for (t = 0; t < n; t++) if (jobs[t].failed) die("GPU worker failed", dctz_gpu_last_error(NULL));
}

size_t dctz_large_bound(size_t N, t_datatype dt) {
  const size_t es = dt == DOUBLE ? 8 : 4, np = (N + g_piece - 1) / g_piece;
  return 24 + 8 * np + N * (1 + es) + N / 8 + np * 8192; /* every coefficient an outlier + zlib slack */
}

size_t dctz_compress_large(const void *data, size_t N, t_datatype dt, double error_bound, void *out, size_t out_cap) {
  large_job j;
  double stats3[3];
  size_t i, total, np;
  unsigned char *o = (unsigned char *)out;
  if (error_bound < 1E-6) { fprintf(stderr, "ERROR: error bound should be no less than 1E-6.\n"); exit(1); }
  if (!data || N == 0) die("nothing to compress", "N == 0");
  memset(&j, 0, sizeof j);
  np = (N + g_piece - 1) / g_piece;
  j.dt = dt; j.data = (const unsigned char *)data; j.N = N; j.npieces = np; j.eb = error_bound;
  j.pmax = (double *)xmalloc(3 * np * sizeof(double), "stats");
  j.pmin = j.pmax + np; j.psum = j.pmin + np;
  run_large(&j, large_stats_worker); /* pass 1: statistics of every piece */
  stats3[0] = j.pmax[0]; stats3[1] = j.pmin[0]; stats3[2] = 0.0;
  for (i = 0; i < np; i++) { /* piece order: deterministic */
    if (j.pmax[i] > stats3[0]) stats3[0] = j.pmax[i];
    if (j.pmin[i] < stats3[1]) stats3[1] = j.pmin[i];
    stats3[2] += j.psum[i];
  }
  j.stats3 = stats3;
  j.streams = (unsigned char **)xmalloc(np * sizeof(unsigned char *), "streams");
  j.sizes = (size_t *)xmalloc(np * sizeof(size_t), "sizes");
  memset(j.streams, 0, np * sizeof(unsigned char *));
  run_large(&j, large_compress_worker); /* pass 2: every piece with the global scaling factor */
  total = 24 + 8 * np;
  for (i = 0; i < np; i++) total += j.sizes[i];
  if (total > out_cap) die("output buffer too small", "dctz_compress_large");
  memcpy(o, "DCTZMS01", 8);
  { const unsigned long long nt = N; const unsigned int d = (unsigned int)dt, ns = (unsigned int)np; memcpy(o + 8, &nt, 8); memcpy(o + 16, &d, 4); memcpy(o + 20, &ns, 4); }
  o += 24;
  for (i = 0; i < np; i++) { const unsigned long long sz = j.sizes[i]; memcpy(o, &sz, 8); o += 8; }
  for (i = 0; i < np; i++) { memcpy(o, j.streams[i], j.sizes[i]); o += j.sizes[i]; free(j.streams[i]); }
  free(j.streams); free(j.sizes); free(j.pmax);
  return total;
}

size_t dctz_decompress_large(const void *in, size_t in_size, void *out, size_t out_elements) {
  const unsigned char *p = (const unsigned char *)in;
  unsigned long long nt;
  unsigned int d, ns;
  size_t *offs, *eoffs, *ecnt, off, eoff, i;
  large_job j;
  if (in_size < 24 || memcmp(p, "DCTZMS01", 8)) die("corrupt stream", "not a DCTZ multi-stream container");
  memcpy(&nt, p + 8, 8); memcpy(&d, p + 16, 4); memcpy(&ns, p + 20, 4);
  if (nt > out_elements) die("output buffer too small", "dctz_decompress_large");
  if (d != (unsigned int)FLOAT && d != (unsigned int)DOUBLE) die("corrupt stream", "container datatype");
  if (ns == 0 || (size_t)ns > (in_size - 24) / 8) die("corrupt stream", "container stream count");
  offs = (size_t *)xmalloc(3 * (size_t)ns * sizeof(size_t), "offsets");
  eoffs = offs + ns; ecnt = eoffs + ns;
  /* The container does not store a piece length: every stream's place in the output follows from the element counts
   * in the streams' own headers (prefix sum), checked against the container's total before anything is decoded. */
  off = 24 + 8 * (size_t)ns;
  eoff = 0;
  for (i = 0; i < ns; i++) {
    unsigned long long sz;
    struct header h;
    memcpy(&sz, p + 24 + 8 * i, 8);
    if (sz < sizeof h || sz > in_size - off) die("corrupt stream", "container sizes exceed the buffer");
    memcpy(&h, p + off, sizeof h);
    if ((unsigned int)h.datatype != d) die("corrupt stream", "a piece's datatype differs from the container's");
    if (h.num_elements == 0 || (unsigned long long)h.num_elements > nt - eoff) die("corrupt stream", "the pieces hold more elements than the container states");
    if (i + 1 < ns && h.num_elements % DCTZ_BLK_SZ) die("corrupt stream", "only the last piece may end with a partial block");
    offs[i] = off; eoffs[i] = eoff; ecnt[i] = h.num_elements;
    off += (size_t)sz;
    eoff += h.num_elements;
  }
  if (eoff != nt) die("corrupt stream", "the pieces' element counts do not add up to the container's total");
  memset(&j, 0, sizeof j);
  j.dt = (t_datatype)d; j.N = (size_t)nt; j.npieces = ns; j.in = p; j.offs = offs; j.elem_offs = eoffs; j.elem_cnt = ecnt; j.out = (unsigned char *)out;
  run_large(&j, large_decode_worker);
  free(offs);
  return (size_t)nt;
}

/* ---- the fine-grained legacy symbols (dctz.h:121-124, dct.h:17-27) ----------------------------------- */
void calc_data_stat(t_var *in, t_bstat *bs, int N) { /* util.c:12-44, reduced on the GPU */
  const int is_double = (in->datatype == DOUBLE);
  dctz_gpu_info info;
  if (dctz_gpu_stats(gpu(), is_double ? (void *)in->buf.d : (void *)in->buf.f, (size_t)N, is_double ? DCTZ_GPU_DOUBLE : DCTZ_GPU_FLOAT,
                     &info) != DCTZ_GPU_OK)
    die("GPU statistics failed", dctz_gpu_last_error(g_ctx));
  if (is_double) { bs->max.d = info.max_abs; bs->min.d = info.min_abs; bs->mean.d = info.mean; bs->sf.d = info.sf; }
  else { bs->max.f = (float)info.max_abs; bs->min.f = (float)info.min_abs; bs->mean.f = (float)info.mean; bs->sf.f = (float)info.sf; }
}

/* binning.c:12-50: centre-out ids; the table is 255 entries, nothing to offload */
void gen_bins(double min, double max, double *bin_center, int nbins, double error_bound) {
  const double width = error_bound * 2 * 1.0;
  int id;
  (void)min; (void)max;
  for (id = 0; id < nbins; id++) {
    const int steps = (id & 1) ? (id / 2 + 1) : -(id / 2);
    bin_center[id] = id ? steps * width : 0.0;
  }
}
void gen_bins_f(float min, float max, float *bin_center, int nbins, float error_bound) {
  const float width = error_bound * 2 * 1.0;
  int id;
  (void)min; (void)max;
  for (id = 0; id < nbins; id++) {
    const int steps = (id & 1) ? (id / 2 + 1) : -(id / 2);
    bin_center[id] = id ? steps * width : 0.0f;
  }
}

/* dct.h:17-27: one block per call, transformed on the GPU.  init/finish have nothing to plan. */
void dct_init(int dn) { (void)dn; (void)gpu(); }
void dct_init_f(int dn) { (void)dn; (void)gpu(); }
void dct_finish(void) {}
void dct_finish_f(void) {}
void idct_finish(void) {}
void idct_finish_f(void) {}
static void dct_one(const void *a, void *b, int dn, int datatype, int inverse) {
  if (dctz_gpu_dct_blocks(gpu(), a, b, 1, dn, datatype, inverse) != DCTZ_GPU_OK) die("GPU DCT failed", dctz_gpu_last_error(g_ctx));
}
void dct_fftw(double *a, double *b, int dn, int nblk) { (void)nblk; dct_one(a, b, dn, DCTZ_GPU_DOUBLE, 0); }
void dct_fftw_f(float *a, float *b, int dn, int nblk) { (void)nblk; dct_one(a, b, dn, DCTZ_GPU_FLOAT, 0); }
void ifft_idct(int dn, double *a, double *data) { dct_one(a, data, dn, DCTZ_GPU_DOUBLE, 1); }
void ifft_idct_f(int dn, float *a, float *data) { dct_one(a, data, dn, DCTZ_GPU_FLOAT, 1); }

/* util.c:54-104: the driver's quality metric; the reductions run on the GPU */
double calc_psnr(t_var *var, t_var *var_r, int N, double error_bound) {
  const int is_double = (var->datatype == DOUBLE);
  double q[4], range;
  (void)error_bound;
  if (dctz_gpu_quality(gpu(), is_double ? (void *)var->buf.d : (void *)var->buf.f, is_double ? (void *)var_r->buf.d : (void *)var_r->buf.f,
                       (size_t)N, is_double ? DCTZ_GPU_DOUBLE : DCTZ_GPU_FLOAT, q) != DCTZ_GPU_OK)
    die("GPU quality metrics failed", dctz_gpu_last_error(g_ctx));
  range = q[1] - q[0];
  printf("Max relative error = %.6f\n", q[2] / range);
  return 20 * log10(range / sqrt(q[3] / N));
}
