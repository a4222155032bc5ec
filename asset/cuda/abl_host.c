/* abl_host.c — see abl_host.h */
#include "abl_host.h"

#include <limits.h>
#include <stdio.h>

/* xorshift128+ (Vigna 2014) with the fixed seed the reference uses, so that the initial
 * population of every model is the same as with the reference `c` backend. */
static uint64_t rng_s0 = 0xdeadbeefULL, rng_s1 = 0xbeefdeadULL;

void abl_host_rng_reset(void) { rng_s0 = 0xdeadbeefULL; rng_s1 = 0xbeefdeadULL; }

static uint64_t rng_next(void) {
  uint64_t a = rng_s0;
  const uint64_t b = rng_s1;
  rng_s0 = b;
  a ^= a << 23;
  rng_s1 = a ^ b ^ (a >> 17) ^ (b >> 26);
  return rng_s1 + b;
}

abl_real random_float(abl_real lo, abl_real hi) {
  uint64_t x = rng_next();
  return lo + (abl_real)x / (abl_real)(UINT64_MAX / (hi - lo));
}

int random_int(int lo, int hi) {
  unsigned n = hi - lo + 1;
  if ((n & (n - 1)) == 0) return rng_next() & (n - 1); /* power of two: `lo` is not added */
  unsigned r = UINT_MAX % n;
  unsigned x;
  do {
    x = rng_next();
  } while (x >= UINT_MAX - r);
  return lo + x % n;
}

void abl_host_check(int rc, const char *what) {
  if (rc == 0) return;
  fprintf(stderr, "abl_cuda: %s failed (%d): %s\n", what, rc, abl_cuda_last_error());
  exit(1);
}

static void json_member(FILE *f, const char *rec, const abl_member_desc *m) {
  const char *p = rec + m->offset;
  fprintf(f, "\"%s\":", m->name);
  switch (m->type) {
    case ABL_TYPE_BOOL: fputs(*(const bool *)p ? "true" : "false", f); break;
    case ABL_TYPE_INT: fprintf(f, "%d", *(const int *)p); break;
    case ABL_TYPE_FLOAT: fprintf(f, "%f", (double)*(const abl_real *)p); break;
    case ABL_TYPE_FLOAT2: {
      const abl_real *v = (const abl_real *)p;
      fprintf(f, "[%f,%f]", (double)v[0], (double)v[1]);
      break;
    }
    case ABL_TYPE_FLOAT3: {
      const abl_real *v = (const abl_real *)p;
      fprintf(f, "[%f,%f,%f]", (double)v[0], (double)v[1], (double)v[2]);
      break;
    }
    default: break;
  }
}

int abl_host_save_json(const abl_host_type *types, int n_types, const char *path) {
  FILE *f = fopen(path, "w");
  if (!f) { fprintf(stderr, "save: cannot open %s\n", path); return 1; }
  fputc('{', f);
  for (int t = 0; t < n_types; t++) {
    const abl_host_type *ty = &types[t];
    if (t) fputc(',', f);
    fprintf(f, "\"%s\":[", ty->desc.name);
    const char *rec = (const char *)ty->agents->data;
    for (size_t i = 0; i < ty->agents->len; i++, rec += ty->desc.stride) {
      if (i) fputs(",\n", f);
      fputc('{', f);
      for (int m = 0; m < ty->desc.n_members; m++) {
        if (m) fputc(',', f);
        json_member(f, rec, &ty->desc.members[m]);
      }
      fputc('}', f);
    }
    fputc(']', f);
  }
  fputc('}', f);
  fclose(f);
  return 0;
}

static void xml_member(FILE *f, const char *rec, const abl_member_desc *m, int for_gpu) {
  const char *p = rec + m->offset;
  const char *name = m->name;
  if (for_gpu && m->is_pos) {
    const abl_real *v = (const abl_real *)p;
    if (m->type == ABL_TYPE_FLOAT2) fprintf(f, "<x>%f</x>\n<y>%f</y>\n<z>0.0</z>\n", (double)v[0], (double)v[1]);
    else if (m->type == ABL_TYPE_FLOAT3) fprintf(f, "<x>%f</x>\n<y>%f</y>\n<z>%f</z>\n", (double)v[0], (double)v[1], (double)v[2]);
    return;
  }
  switch (m->type) {
    case ABL_TYPE_BOOL: fprintf(f, "<%s>%d</%s>\n", name, *(const bool *)p ? 1 : 0, name); break;
    case ABL_TYPE_INT: fprintf(f, "<%s>%d</%s>\n", name, *(const int *)p, name); break;
    case ABL_TYPE_FLOAT: fprintf(f, "<%s>%f</%s>\n", name, (double)*(const abl_real *)p, name); break;
    case ABL_TYPE_FLOAT2: {
      const abl_real *v = (const abl_real *)p;
      fprintf(f, "<%s_x>%f</%s_x>\n<%s_y>%f</%s_y>\n", name, (double)v[0], name, name, (double)v[1], name);
      break;
    }
    case ABL_TYPE_FLOAT3: {
      const abl_real *v = (const abl_real *)p;
      fprintf(f, "<%s_x>%f</%s_x>\n<%s_y>%f</%s_y>\n<%s_z>%f</%s_z>\n", name, (double)v[0], name, name,
              (double)v[1], name, name, (double)v[2], name);
      break;
    }
    default: break;
  }
}

int abl_host_save_flame_xml(const abl_host_type *types, int n_types, const char *path, int for_gpu) {
  FILE *f = fopen(path, "w");
  if (!f) { fprintf(stderr, "save: cannot open %s\n", path); return 1; }
  fputs("<states>\n<itno>0</itno>\n", f);
  for (int t = 0; t < n_types; t++) {
    const abl_host_type *ty = &types[t];
    const char *rec = (const char *)ty->agents->data;
    for (size_t i = 0; i < ty->agents->len; i++, rec += ty->desc.stride) {
      fprintf(f, "<xagent>\n<name>%s</name>\n", ty->desc.name);
      for (int m = 0; m < ty->desc.n_members; m++) xml_member(f, rec, &ty->desc.members[m], for_gpu);
      fputs("</xagent>\n", f);
    }
  }
  fputs("</states>\n", f);
  fclose(f);
  return 0;
}

int abl_host_save_raw(const abl_host_type *types, int n_types, const char *path) {
  FILE *f = fopen(path, "wb");
  if (!f) return 1;
  for (int t = 0; t < n_types; t++) {
    uint64_t n = types[t].agents->len;
    uint32_t stride = types[t].desc.stride;
    fwrite(&n, sizeof n, 1, f);
    fwrite(&stride, sizeof stride, 1, f);
    fwrite(types[t].agents->data, stride, n, f);
  }
  fclose(f);
  return 0;
}

static FILE *log_file = NULL;
void abl_host_log_open(const char *path) { if (!log_file) log_file = fopen(path, "w"); }
void abl_host_log_int(int first, int v) { if (log_file) fprintf(log_file, first ? "%d" : ",%d", v); }
void abl_host_log_float(int first, double v) { if (log_file) fprintf(log_file, first ? "%f" : ",%f", v); }
void abl_host_log_end(void) { if (log_file) { fputc('\n', log_file); fflush(log_file); } }

/* ---- frames of a visualised run (abl_host.h) ------------------------------------------------- */
extern int mkdir(const char *path, unsigned int mode);   /* <sys/stat.h>, POSIX: not declared under -std=c99 */
void abl_host_make_dir(const char *path) { (void)mkdir(path, 0777); }

int abl_host_frame_begin(abl_frame *f, int size, double min_x, double min_y, double max_x, double max_y) {
  double w = max_x - min_x, h = max_y - min_y, side = w > h ? w : h;
  f->size = size;
  f->min_x = min_x;
  f->min_y = min_y;
  f->scale = side > 0 ? size / side : 1.0;
  f->rgb = (unsigned char *)malloc((size_t)size * size * 3);
  if (!f->rgb) return 1;
  memset(f->rgb, 0xff, (size_t)size * size * 3);   /* display.setBackdrop(Color.white) */
  return 0;
}

void abl_host_frame_dot(abl_frame *f, double x, double y, int rgb, double size) {
  if (!(x == x) || !(y == y) || !(size > 0)) return;
  const double cx = (x - f->min_x) * f->scale, cy = (y - f->min_y) * f->scale;
  const double r = 2.0 * size;                       /* an oval of 4 * getSize pixels across */
  int x0 = (int)floor(cx - r), x1 = (int)ceil(cx + r), y0 = (int)floor(cy - r), y1 = (int)ceil(cy + r);
  if (x1 < 0 || y1 < 0 || x0 >= f->size || y0 >= f->size) return;
  if (x0 < 0) x0 = 0;
  if (y0 < 0) y0 = 0;
  if (x1 >= f->size) x1 = f->size - 1;
  if (y1 >= f->size) y1 = f->size - 1;
  const unsigned char red = (unsigned char)(rgb >> 16), green = (unsigned char)(rgb >> 8), blue = (unsigned char)rgb;
  for (int py = y0; py <= y1; py++)
    for (int px = x0; px <= x1; px++) {
      const double dx = px + 0.5 - cx, dy = py + 0.5 - cy;   /* pixel centres */
      if (dx * dx + dy * dy > r * r) continue;
      unsigned char *q = f->rgb + ((size_t)py * f->size + px) * 3;
      q[0] = red; q[1] = green; q[2] = blue;
    }
}

int abl_host_frame_end(abl_frame *f, const char *path) {
  int rc = 1;
  FILE *out = fopen(path, "wb");
  if (out) {
    fprintf(out, "P6\n%d %d\n255\n", f->size, f->size);
    rc = fwrite(f->rgb, 3, (size_t)f->size * f->size, out) == (size_t)f->size * f->size ? 0 : 1;
    if (fclose(out) != 0) rc = 1;
  }
  free(f->rgb);
  f->rgb = NULL;
  return rc;
}
