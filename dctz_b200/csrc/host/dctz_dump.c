/* dctz_dump.c -- print the header of a DCTZ stream (what the reference's tools/dctz-dump.c prints), plus the
 * section sizes, and walk a multi-stream container (dctz_compress_large).   usage: dctz-dump <file> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/dctz_compat.h"

static void print_header(const struct header *h, const char *indent) {
  printf("%sdata type=%s\n", indent, h->datatype == DOUBLE ? "double" : "float");
  printf("%sN=%u\n", indent, h->num_elements);
  printf("%serror_bound=%f\n", indent, h->error_bound);
  printf("%stotal # of AC_exact=%u\n", indent, h->tot_AC_exact_count);
  printf("%sSF=%f\n", indent, h->datatype == DOUBLE ? h->scaling_factor.d : (double)h->scaling_factor.f);
  printf("%smean=%g\n", indent, h->datatype == DOUBLE ? h->mean.d : (double)h->mean.f);
  printf("%ssections (compressed bytes): bin_index=%u DC=%u AC_exact=%u\n", indent, h->bindex_sz_compressed, h->DC_sz_compressed,
         h->AC_exact_sz_compressed);
}

int main(int argc, char *argv[]) {
  unsigned char head[24];
  struct header h;
  FILE *f;
  if (argc != 2) { printf("Usage: %s filename\n", argv[0]); return 0; }
  f = fopen(argv[1], "rb");
  if (!f) { perror("Failed: "); printf("File Not Found\n"); return 0; }
  printf("File Name=%s\n", argv[1]);
  if (fread(head, 1, sizeof head, f) != sizeof head) { printf("file too short\n"); return 1; }
  if (!memcmp(head, "DCTZMS01", 8)) {
    unsigned long long nt, *sizes, off;
    unsigned int dt, ns, i;
    memcpy(&nt, head + 8, 8); memcpy(&dt, head + 16, 4); memcpy(&ns, head + 20, 4);
    printf("multi-stream container: N=%llu, data type=%s, streams=%u\n", nt, dt == DOUBLE ? "double" : "float", ns);
    sizes = (unsigned long long *)malloc(8ull * ns);
    if (!sizes || fread(sizes, 8, ns, f) != ns) { printf("truncated container\n"); return 1; }
    off = 24 + 8ull * ns;
    for (i = 0; i < ns; i++) {
      if (fseek(f, (long)off, SEEK_SET) || fread(&h, sizeof h, 1, f) != 1) { printf("truncated container\n"); return 1; }
      printf("stream %u at offset %llu, %llu bytes\n", i, off, sizes[i]);
      print_header(&h, "  ");
      off += sizes[i];
    }
    free(sizes);
  } else {
    rewind(f);
    if (fread(&h, sizeof h, 1, f) != 1) { printf("file too short\n"); return 1; }
    print_header(&h, "");
  }
  fclose(f);
  return 0;
}
