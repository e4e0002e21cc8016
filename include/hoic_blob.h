/* hoic_blob.h -- reader for the flat model blob written by hoic_b200/blob.py (format documented there).
 * Header-only, plain C99 / C++.  Used by the CUDA library (hoic_b200/csrc) and by the CPU oracle; it only
 * parses the container format, no model semantics live here.
 */
#ifndef HOIC_BLOB_H
#define HOIC_BLOB_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOIC_BLOB_F64 0
#define HOIC_BLOB_I32 1
#define HOIC_BLOB_BYTES 2

typedef struct hoic_blob_view {
  const unsigned char* base;
  size_t nbytes;
  uint32_t n_entries;
} hoic_blob_view;

/* entry: char name[40]; u32 dtype; u32 ndim; u32 shape[4]; u64 offset; u64 nbytes  => 80 bytes */
#define HOIC_BLOB_ENTRY_SIZE 80

static inline int hoic_blob_open(hoic_blob_view* v, const void* blob, size_t nbytes) {
  const unsigned char* b = (const unsigned char*)blob;
  uint32_t version, n;
  if (!blob || nbytes < 16 || memcmp(b, "HOICMDL1", 8) != 0) return -1;
  memcpy(&version, b + 8, 4);
  memcpy(&n, b + 12, 4);
  if (version != 1) return -2;
  if (16 + (size_t)n * HOIC_BLOB_ENTRY_SIZE > nbytes) return -3;
  v->base = b;
  v->nbytes = nbytes;
  v->n_entries = n;
  return 0;
}

/* returns pointer to the array data (8-byte aligned inside the blob) or NULL; count = number of elements */
static inline const void* hoic_blob_find(const hoic_blob_view* v, const char* name, int* dtype, size_t* count) {
  uint32_t i;
  for (i = 0; i < v->n_entries; i++) {
    const unsigned char* e = v->base + 16 + (size_t)i * HOIC_BLOB_ENTRY_SIZE;
    if (strncmp((const char*)e, name, 40) == 0) {
      uint32_t dt;
      uint64_t off, nb;
      memcpy(&dt, e + 40, 4);
      memcpy(&off, e + 64, 8);
      memcpy(&nb, e + 72, 8);
      if (off + nb > v->nbytes) return NULL;
      if (dtype) *dtype = (int)dt;
      if (count) *count = (size_t)(dt == HOIC_BLOB_F64 ? nb / 8 : dt == HOIC_BLOB_I32 ? nb / 4 : nb);
      return v->base + off;
    }
  }
  return NULL;
}

#ifdef __cplusplus
}
#endif
#endif /* HOIC_BLOB_H */
