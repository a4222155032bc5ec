/* Link-time wrapper around the reference runtime's save() (reference asset/c/libabl.c:215-231),
 * used only when building reference-generated programs for parity tests:
 *     gcc ... main.c libabl.c save_wrap.c -Wl,--wrap=save
 * Besides calling the real save() it dumps the raw agent records next to the JSON file
 * (`<path>.bin`: per agent type `u64 n, u32 stride`, then n records), because the JSON keeps
 * only 6 decimals and parity is checked to 1e-9.  With ABL_REF_XML set it also has the
 * reference write its two other formats (`<path>.flame.xml`, `<path>.flamegpu.xml`; reference
 * asset/c/libabl.c:126-213), the goldens of tests/test_save_formats.py.  The reference sources
 * are not modified. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "libabl.h"

void __real_save(void *agents, const agent_info *info, const char *path, save_type type);

void __wrap_save(void *agents, const agent_info *info, const char *path, save_type type) {
  char raw[4096];
  snprintf(raw, sizeof raw, "%s.bin", path);
  FILE *f = fopen(raw, "wb");
  if (f) {
    for (const agent_info *ai = info; ai->name; ai++) {
      const dyn_array *arr = (const dyn_array *)((const char *)agents + ai->offset);
      const type_info *ti = ai->info;
      while (ti->type != TYPE_END) ti++;
      uint64_t n = arr->len;
      uint32_t stride = ti->offset;
      fwrite(&n, sizeof n, 1, f);
      fwrite(&stride, sizeof stride, 1, f);
      fwrite(arr->values, stride, n, f);
    }
    fclose(f);
  }
  __real_save(agents, info, path, type);
  if (getenv("ABL_REF_XML")) {
    snprintf(raw, sizeof raw, "%s.flame.xml", path);
    __real_save(agents, info, raw, SAVE_FLAME_XML);
    snprintf(raw, sizeof raw, "%s.flamegpu.xml", path);
    __real_save(agents, info, raw, SAVE_FLAMEGPU_XML);
  }
}
