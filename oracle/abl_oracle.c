/* abl_oracle.c — CPU restatement of the reference `c` backend's hot path, in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  May be imported/linked/executed only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, and there
 * only as the checker or the reported CPU baseline — never by the product path.
 *
 * What is restated, and from where (all paths relative to /root/reference):
 *   - vector helpers, sqrtf-in-double length/dist ............ asset/c/libabl.h:69-175
 *   - xorshift128+ / random_float / random_int ................ asset/c/libabl.c:8-39
 *   - the simulate loop: for every agent call step(in,out) .... src/backend/CPrinter.cpp:189-233
 *   - for-near: scan ALL agents in index order, skip when
 *     dist(nx.pos, in.pos) > radius (inclusive, self included)   src/backend/CPrinter.cpp:148-173
 *   - operator lowering (vec +,-,*s,/s; a += b -> a = add(a,b)) src/backend/GenericCPrinter.cpp:20-75
 *   - constants re-printed with 6 significant digits ........... src/backend/GenericPrinter.cpp:39-58,
 *                                                                 src/AnalysisVisitor.cpp:403-431
 *   - per-model step functions: the C the reference printer emits for examples/{circle,
 *     circle3d,boids2d,game_of_life}.abl (obtained by running the real reference compiler,
 *     oracle/_ref/OpenABL_ref; see oracle/refgen.py).
 *
 * Pinning: tests/test_oracle.py checks this file against golden vectors produced by the
 * real reference (tests/golden/, generator oracle/refgen.py): initial states bit-equal,
 * brute-force mode bit-equal after 10-100 steps.
 *
 * Two neighbour modes:
 *   MODE_BRUTE  reference order (all agents, ascending index) — bit-equal to the reference.
 *   MODE_GRID   uniform grid, cells visited in ascending cell key and agents in ascending id
 *               inside a cell, exactly the traversal order of the CUDA kernels, so GPU results
 *               can be compared bit for bit at sizes the O(N^2) reference cannot run.
 *
 * Build twice: default (abl_float = double) and -DLIBABL_USE_FLOAT=1 (float), like the
 * reference (src/backend/CBackend.cpp:21-27), with `gcc -O2 -std=c99 -fopenmp`.
 */
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* OpenMP threads the parallel loops below will use, and a way to set them: torchrun exports
 * OMP_NUM_THREADS=1 to its children, which would silently time the CPU baseline on one core. */
int oracle_omp_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

#ifdef LIBABL_USE_FLOAT
typedef float abl_float;
#else
typedef double abl_float;
#endif

enum { MODE_BRUTE = 0, MODE_GRID = 1 };

/* ---- libabl.h:69-175 ------------------------------------------------------------------ */
typedef struct { abl_float x, y; } float2;
typedef struct { abl_float x, y, z; } float3;

static inline float2 float2_create(abl_float x, abl_float y) { return (float2){x, y}; }
static inline float2 float2_fill(abl_float x) { return (float2){x, x}; }
static inline float2 float2_add(float2 a, float2 b) { return (float2){a.x + b.x, a.y + b.y}; }
static inline float2 float2_sub(float2 a, float2 b) { return (float2){a.x - b.x, a.y - b.y}; }
static inline float2 float2_mul_scalar(float2 a, abl_float s) { return (float2){a.x * s, a.y * s}; }
static inline float2 float2_div_scalar(float2 a, abl_float s) { return (float2){a.x / s, a.y / s}; }
static inline float3 float3_create(abl_float x, abl_float y, abl_float z) { return (float3){x, y, z}; }
static inline float3 float3_fill(abl_float x) { return (float3){x, x, x}; }
static inline float3 float3_add(float3 a, float3 b) { return (float3){a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline float3 float3_sub(float3 a, float3 b) { return (float3){a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float3 float3_mul_scalar(float3 a, abl_float s) { return (float3){a.x * s, a.y * s, a.z * s}; }
static inline float3 float3_div_scalar(float3 a, abl_float s) { return (float3){a.x / s, a.y / s, a.z / s}; }
/* libabl.h:156-168: float square root even when abl_float is double */
static inline abl_float length_float2(float2 v) { return sqrtf(v.x * v.x + v.y * v.y); }
static inline abl_float length_float3(float3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
static inline abl_float dist_float2(float2 a, float2 b) { return length_float2(float2_sub(a, b)); }
static inline abl_float dist_float3(float3 a, float3 b) { return length_float3(float3_sub(a, b)); }

/* ---- libabl.c:8-39 -------------------------------------------------------------------- */
static uint64_t xorshift_state[2] = {0xdeadbeef, 0xbeefdead};

void oracle_rng_reset(void) { xorshift_state[0] = 0xdeadbeef; xorshift_state[1] = 0xbeefdead; }

static uint64_t xorshift128plus(void) {
  uint64_t x = xorshift_state[0];
  uint64_t const y = xorshift_state[1];
  xorshift_state[0] = y;
  x ^= x << 23;
  xorshift_state[1] = x ^ y ^ (x >> 17) ^ (y >> 26);
  return xorshift_state[1] + y;
}
static abl_float random_float(abl_float min, abl_float max) {
  uint64_t x = xorshift128plus();
  return min + (abl_float)x / (abl_float)(UINT64_MAX / (max - min));
}

/* ---- constant folding + 6-digit re-print (GenericPrinter.cpp:42-50) ------------------- */
/* A global `float W = <expr>` is folded in double and printed by `ostream << double`
 * (precision 6), so the value the generated program computes with is strtod("%.6g"). */
double oracle_fold6(double v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.6g", v);
  return strtod(buf, NULL);
}

/* ---- uniform grid (CUDA traversal order) ------------------------------------------------ */
typedef struct {
  int dim;
  int n_cell[3];
  abl_float origin[3];
  abl_float cell;
  abl_float inv_cell; /* 1/cell in abl_float precision, as in the CUDA runtime */
  int n_cells;
  int *cell_start; /* [n_cells + 1] */
  int *order;      /* agent indices sorted by (cell, id) */
} grid_t;

/* same formula and precision as abl_cell_coord (asset/cuda/abl_device.cuh) */
static inline int cell_coord(abl_float p, abl_float origin, abl_float inv_cell, int n) {
  int c = (int)floor((p - origin) * inv_cell);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

static double oracle_padded_cell(double granularity) {
  return granularity * (1.0 + (sizeof(abl_float) == 8 ? 0x1p-20 : 0x1p-10));
}

static void grid_setup(grid_t *g, int dim, const double *env_min, const double *env_max, double granularity) {
  g->dim = dim;
  /* cells a shade larger than the granularity, as in abl_cuda_set_environment (ABL_CELL_PAD_*) */
  granularity = oracle_padded_cell(granularity);
  g->cell = (abl_float)granularity;
  g->inv_cell = (abl_float)1 / (abl_float)granularity;
  g->n_cells = 1;
  for (int a = 0; a < 3; a++) {
    if (a < dim) {
      long nc = (long)ceil((env_max[a] - env_min[a]) / granularity);
      g->n_cell[a] = nc < 1 ? 1 : (int)nc;
      g->origin[a] = (abl_float)env_min[a];
    } else {
      g->n_cell[a] = 1;
      g->origin[a] = 0;
    }
    g->n_cells *= g->n_cell[a];
  }
  g->cell_start = NULL;
  g->order = NULL;
}

static inline int cell_of(const grid_t *g, const abl_float *p) {
  int cx = cell_coord(p[0], g->origin[0], g->inv_cell, g->n_cell[0]);
  int cy = cell_coord(p[1], g->origin[1], g->inv_cell, g->n_cell[1]);
  int c = cy * g->n_cell[0] + cx;
  if (g->dim == 3) c += cell_coord(p[2], g->origin[2], g->inv_cell, g->n_cell[2]) * g->n_cell[0] * g->n_cell[1];
  return c;
}

/* pos: pointer to the position of agent 0, stride in bytes between agents */
static void grid_bin(grid_t *g, const char *pos, size_t stride, int n) {
  free(g->cell_start);
  free(g->order);
  g->cell_start = calloc((size_t)g->n_cells + 1, sizeof(int));
  g->order = malloc(sizeof(int) * (size_t)(n ? n : 1));
  int *key = malloc(sizeof(int) * (size_t)(n ? n : 1));
  for (int i = 0; i < n; i++) {
    key[i] = cell_of(g, (const abl_float *)(pos + stride * i));
    g->cell_start[key[i] + 1]++;
  }
  for (int c = 0; c < g->n_cells; c++) g->cell_start[c + 1] += g->cell_start[c];
  int *fill = malloc(sizeof(int) * (size_t)g->n_cells);
  memcpy(fill, g->cell_start, sizeof(int) * (size_t)g->n_cells);
  for (int i = 0; i < n; i++) g->order[fill[key[i]]++] = i; /* ascending id inside a cell */
  free(fill);
  free(key);
}

static void grid_free(grid_t *g) { free(g->cell_start); free(g->order); g->cell_start = g->order = NULL; }

/* Neighbour candidate enumeration shared by all models.  BODY sees `j` (agent index). */
#define FOR_CANDIDATES(g, mode, n, self_pos, reach, BODY)                                        \
  if ((mode) == MODE_BRUTE) {                                                                     \
    for (int j = 0; j < (n); j++) { BODY }                                                        \
  } else {                                                                                        \
    int cx_ = cell_coord((self_pos)[0], (g)->origin[0], (g)->inv_cell, (g)->n_cell[0]);               \
    int cy_ = cell_coord((self_pos)[1], (g)->origin[1], (g)->inv_cell, (g)->n_cell[1]);               \
    int cz_ = (g)->dim == 3 ? cell_coord((self_pos)[2], (g)->origin[2], (g)->inv_cell, (g)->n_cell[2]) : 0; \
    int x0_ = cx_ - (reach) < 0 ? 0 : cx_ - (reach);                                              \
    int x1_ = cx_ + (reach) >= (g)->n_cell[0] ? (g)->n_cell[0] - 1 : cx_ + (reach);               \
    int y0_ = cy_ - (reach) < 0 ? 0 : cy_ - (reach);                                              \
    int y1_ = cy_ + (reach) >= (g)->n_cell[1] ? (g)->n_cell[1] - 1 : cy_ + (reach);               \
    int z0_ = 0, z1_ = 0;                                                                         \
    if ((g)->dim == 3) {                                                                          \
      z0_ = cz_ - (reach) < 0 ? 0 : cz_ - (reach);                                                \
      z1_ = cz_ + (reach) >= (g)->n_cell[2] ? (g)->n_cell[2] - 1 : cz_ + (reach);                 \
    }                                                                                             \
    for (int z_ = z0_; z_ <= z1_; z_++)                                                           \
      for (int y_ = y0_; y_ <= y1_; y_++) {                                                       \
        int base_ = (z_ * (g)->n_cell[1] + y_) * (g)->n_cell[0];                                  \
        int b_ = (g)->cell_start[base_ + x0_], e_ = (g)->cell_start[base_ + x1_ + 1];             \
        for (int q_ = b_; q_ < e_; q_++) { int j = (g)->order[q_]; BODY }                         \
      }                                                                                           \
  }

static int reach_for(double radius, double cell) {
  int r = (int)ceil(radius / oracle_padded_cell(cell) - 1e-12);
  return r < 1 ? 1 : r;
}

/* ===================================================================================== */
/* circle.abl / circle3d.abl                                                             */
/* ===================================================================================== */
typedef struct { float2 pos; } Point2;
typedef struct { float3 pos; } Point3;

typedef struct {
  abl_float rho, k_rep, k_att, r, W; /* as the generated program declares them (6 digits) */
  double W_exact;                   /* folded, unrounded: environment bound */
  double radius_exact;              /* folded 2*r: granularity */
} circle_consts;

static circle_consts circle_fold(int dim, int num_agents, double rho, double k_rep, double k_att, double r) {
  circle_consts c;
  double w = dim == 2 ? sqrt(num_agents / rho) : cbrt(num_agents / rho); /* AnalysisVisitor.cpp:422-431 */
  c.W_exact = w;
  c.W = (abl_float)oracle_fold6(w);
  c.rho = (abl_float)oracle_fold6(rho);
  c.k_rep = (abl_float)oracle_fold6(k_rep);
  c.k_att = (abl_float)oracle_fold6(k_att);
  c.r = (abl_float)oracle_fold6(r);
  c.radius_exact = 2 * r;
  return c;
}

/* main(): for i in 0..num_agents: add(Point{pos: random(floatN(W))}); gcc evaluates the
 * arguments of floatN_create(random_float(..x), random_float(..y)[, ..z]) right to left. */
void oracle_circle_init(int dim, int num_agents, double rho, abl_float *pos) {
  circle_consts c = circle_fold(dim, num_agents, rho, 0, 0, 0);
  for (int i = 0; i < num_agents; i++) {
    if (dim == 2) {
      abl_float y = random_float(0, c.W);
      abl_float x = random_float(0, c.W);
      pos[2 * i] = x; pos[2 * i + 1] = y;
    } else {
      abl_float z = random_float(0, c.W);
      abl_float y = random_float(0, c.W);
      abl_float x = random_float(0, c.W);
      pos[3 * i] = x; pos[3 * i + 1] = y; pos[3 * i + 2] = z;
    }
  }
}

static inline float2 clamp_1(float2 pos, float2 min, float2 max) { /* lib.abl clamp(float2,float2,float2) */
  return float2_create(((pos.x < min.x) ? min.x : ((pos.x > max.x) ? max.x : pos.x)),
                       ((pos.y < min.y) ? min.y : ((pos.y > max.y) ? max.y : pos.y)));
}
static inline float3 clamp_2(float3 pos, float3 min, float3 max) {
  return float3_create(((pos.x < min.x) ? min.x : ((pos.x > max.x) ? max.x : pos.x)),
                       ((pos.y < min.y) ? min.y : ((pos.y > max.y) ? max.y : pos.y)),
                       ((pos.z < min.z) ? min.z : ((pos.z > max.z) ? max.z : pos.z)));
}

/* One timestep of move_point for agents [i0, i1) against the whole population. */
/* granularity <= 0: automatic (the for-near radius, AnalysisVisitor.cpp:1378-1399) */
void oracle_circle_step_g(int dim, int num_agents, double rho_, double k_rep_, double k_att_, double r_,
                          const abl_float *in_pos, abl_float *out_pos, int n, int mode, int i0, int i1,
                          double granularity);

void oracle_circle_step(int dim, int num_agents, double rho_, double k_rep_, double k_att_, double r_,
                        const abl_float *in_pos, abl_float *out_pos, int n, int mode, int i0, int i1) {
  oracle_circle_step_g(dim, num_agents, rho_, k_rep_, k_att_, r_, in_pos, out_pos, n, mode, i0, i1, 0.0);
}

void oracle_circle_step_g(int dim, int num_agents, double rho_, double k_rep_, double k_att_, double r_,
                          const abl_float *in_pos, abl_float *out_pos, int n, int mode, int i0, int i1,
                          double granularity) {
  const circle_consts c = circle_fold(dim, num_agents, rho_, k_rep_, k_att_, r_);
  const double cell_size = granularity > 0 ? granularity : c.radius_exact;
  const abl_float r = c.r, k_att = c.k_att, k_rep = c.k_rep, W = c.W;
  grid_t g;
  double lo[3] = {0, 0, 0}, hi[3] = {c.W_exact, c.W_exact, c.W_exact};
  grid_setup(&g, dim, lo, hi, cell_size);
  if (mode == MODE_GRID) grid_bin(&g, (const char *)in_pos, sizeof(abl_float) * dim, n);
  const int reach = reach_for(c.radius_exact, cell_size);
  if (dim == 2) {
    const Point2 *buf = (const Point2 *)in_pos;
    Point2 *dbuf = (Point2 *)out_pos;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = i0; i < i1; i++) {
      const Point2 *in = &buf[i];
      float2 new_pos = in->pos;
      FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
        const Point2 *nx = &buf[j];
        if (dist_float2(nx->pos, in->pos) > (2.0 * r)) continue;
        abl_float pos_dist = dist_float2(in->pos, nx->pos);
        if ((pos_dist == 0)) continue;
        abl_float sep_dist = (pos_dist - r);
        float2 force = float2_fill(0);
        if ((sep_dist > 0.0)) {
          force = float2_div_scalar(float2_mul_scalar(float2_sub(nx->pos, in->pos), (k_att * ((2.0 * r) - pos_dist))), pos_dist);
        } else {
          force = float2_mul_scalar(float2_sub(in->pos, nx->pos), k_rep);
        }
        new_pos = float2_add(new_pos, force);
      })
      dbuf[i].pos = clamp_1(new_pos, float2_fill(1), float2_fill((W - 1.0)));
    }
  } else {
    const Point3 *buf = (const Point3 *)in_pos;
    Point3 *dbuf = (Point3 *)out_pos;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = i0; i < i1; i++) {
      const Point3 *in = &buf[i];
      float3 new_pos = in->pos;
      FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
        const Point3 *nx = &buf[j];
        if (dist_float3(nx->pos, in->pos) > (2.0 * r)) continue;
        abl_float pos_dist = dist_float3(in->pos, nx->pos);
        if ((pos_dist == 0)) continue;
        abl_float sep_dist = (pos_dist - r);
        float3 force = float3_fill(0);
        if ((sep_dist > 0.0)) {
          force = float3_div_scalar(float3_mul_scalar(float3_sub(nx->pos, in->pos), (k_att * ((2.0 * r) - pos_dist))), pos_dist);
        } else {
          force = float3_mul_scalar(float3_sub(in->pos, nx->pos), k_rep);
        }
        new_pos = float3_add(new_pos, force);
      })
      dbuf[i].pos = clamp_2(new_pos, float3_fill(1), float3_fill((W - 1.0)));
    }
  }
  grid_free(&g);
}

/* ===================================================================================== */
/* boids2d.abl                                                                           */
/* ===================================================================================== */
typedef struct { float2 pos; float2 velocity; } Boid;

typedef struct {
  abl_float interaction_radius, separation_radius, time_scale, global_scale, steer_scale,
      collision_scale, match_scale, min_pos, max_pos;
  double max_pos_exact, radius_exact;
} boids_consts;

/* agent_density is an *integer* literal in the model (`param float agent_density = 500`),
 * and the folded symbol value stays an integer, so num_agents/agent_density folds as
 * integer division (AnalysisVisitor.cpp:422-471). */
static boids_consts boids_fold(int num_agents, int agent_density, double interaction_radius,
                               double separation_radius) {
  boids_consts c;
  double mp = sqrt((double)(num_agents / agent_density));
  c.max_pos_exact = mp;
  c.max_pos = (abl_float)oracle_fold6(mp);
  c.min_pos = (abl_float)0.0;
  c.interaction_radius = (abl_float)oracle_fold6(interaction_radius);
  c.separation_radius = (abl_float)oracle_fold6(separation_radius);
  c.time_scale = (abl_float)0.0005;
  c.global_scale = (abl_float)0.15;
  c.steer_scale = (abl_float)0.65;
  c.collision_scale = (abl_float)0.75;
  c.match_scale = (abl_float)1.25;
  c.radius_exact = interaction_radius;
  return c;
}

void oracle_boids2d_init(int num_agents, int agent_density, Boid *boids) {
  boids_consts c = boids_fold(num_agents, agent_density, 0.05, 0.005);
  for (int i = 0; i < num_agents; i++) {
    /* struct initialiser members in textual order; call arguments right to left */
    abl_float py = random_float(c.min_pos, c.max_pos);
    abl_float px = random_float(c.min_pos, c.max_pos);
    abl_float vy = random_float((-1), 1);
    abl_float vx = random_float((-1), 1);
    boids[i].pos = float2_create(px, py);
    boids[i].velocity = float2_create(vx, vy);
  }
}

void oracle_boids2d_step(int num_agents, int agent_density, double interaction_radius_,
                         double separation_radius_, const Boid *buf, Boid *dbuf, int n, int mode,
                         int i0, int i1) {
  const boids_consts c = boids_fold(num_agents, agent_density, interaction_radius_, separation_radius_);
  const abl_float interaction_radius = c.interaction_radius, separation_radius = c.separation_radius;
  const abl_float time_scale = c.time_scale, global_scale = c.global_scale, steer_scale = c.steer_scale;
  const abl_float collision_scale = c.collision_scale, match_scale = c.match_scale;
  const abl_float min_pos = c.min_pos, max_pos = c.max_pos;
  grid_t g;
  double lo[3] = {0, 0, 0}, hi[3] = {c.max_pos_exact, c.max_pos_exact, 0};
  grid_setup(&g, 2, lo, hi, c.radius_exact);
  if (mode == MODE_GRID) grid_bin(&g, (const char *)buf, sizeof(Boid), n);
  const int reach = 1;
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = i0; i < i1; i++) {
    const Boid *in = &buf[i];
    Boid *out = &dbuf[i];
    float2 global_velocity = float2_create(0, 0);
    float2 global_center = float2_create(0, 0);
    float2 collision_center = float2_create(0, 0);
    int interaction_count = 0;
    int collision_count = 0;
    FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
      const Boid *nx = &buf[j];
      if (dist_float2(nx->pos, in->pos) > interaction_radius) continue;
      global_center = float2_add(global_center, nx->pos);
      global_velocity = float2_add(global_velocity, nx->velocity);
      interaction_count += 1;
      abl_float separation = dist_float2(in->pos, nx->pos);
      if ((separation < separation_radius)) {
        collision_center = float2_add(collision_center, nx->pos);
        collision_count += 1;
      }
    })
    float2 velocity_change = float2_create(0, 0);
    float2 steer_velocity = float2_create(0, 0);
    if ((interaction_count > 0)) {
      global_center = float2_div_scalar(global_center, interaction_count);
      steer_velocity = float2_mul_scalar(float2_sub(global_center, in->pos), steer_scale);
    }
    velocity_change = float2_add(velocity_change, steer_velocity);
    float2 match_velocity = float2_create(0, 0);
    if ((interaction_count > 0)) {
      global_velocity = float2_div_scalar(global_velocity, interaction_count);
      match_velocity = float2_mul_scalar(global_velocity, match_scale);
    }
    velocity_change = float2_add(velocity_change, match_velocity);
    float2 avoid_velocity = float2_create(0, 0);
    if ((collision_count > 0)) {
      collision_center = float2_div_scalar(collision_center, collision_count);
      avoid_velocity = float2_mul_scalar(float2_sub(in->pos, collision_center), collision_scale);
    }
    velocity_change = float2_add(velocity_change, avoid_velocity);
    float2 new_velocity = float2_add(in->velocity, float2_mul_scalar(velocity_change, global_scale));
    abl_float new_velocity_scale = length_float2(new_velocity);
    if ((new_velocity_scale > 1)) {
      new_velocity = float2_div_scalar(new_velocity, new_velocity_scale);
    }
    float2 new_pos = float2_add(in->pos, float2_mul_scalar(new_velocity, time_scale));
    /* boundPosition */
    new_pos = float2_create(((new_pos.x < min_pos) ? max_pos : ((new_pos.x > max_pos) ? min_pos : new_pos.x)),
                            ((new_pos.y < min_pos) ? max_pos : ((new_pos.y > max_pos) ? min_pos : new_pos.y)));
    out->pos = new_pos;
    out->velocity = new_velocity;
  }
  grid_free(&g);
}

/* ===================================================================================== */
/* game_of_life.abl                                                                      */
/* ===================================================================================== */
typedef struct { float2 pos; bool alive; } Cell;

int oracle_gol_size(int num_agents) { return (int)(long)sqrt((double)num_agents); }
int oracle_gol_record_size(void) { return (int)sizeof(Cell); }

void oracle_gol_init(int num_agents, double alive_fraction_, Cell *cells) {
  const int size = oracle_gol_size(num_agents);
  const abl_float alive_fraction = (abl_float)oracle_fold6(alive_fraction_);
  int k = 0;
  for (int x = 0; x < size; x++) {
    for (int y = 0; y < size; y++) {
      memset(&cells[k], 0, sizeof(Cell));
      cells[k].pos = float2_create(((abl_float)x + 0.5), ((abl_float)y + 0.5));
      cells[k].alive = (random_float(0, 1) < alive_fraction);
      k++;
    }
  }
}

void oracle_gol_step(int num_agents, const Cell *buf, Cell *dbuf, int n, int mode, int i0, int i1) {
  const int size = oracle_gol_size(num_agents);
  const abl_float radius = (abl_float)1.5;
  grid_t g;
  double lo[3] = {0, 0, 0}, hi[3] = {(double)size, (double)size, 0};
  grid_setup(&g, 2, lo, hi, 1.5);
  if (mode == MODE_GRID) grid_bin(&g, (const char *)buf, sizeof(Cell), n);
  const int reach = 1;
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = i0; i < i1; i++) {
    const Cell *in = &buf[i];
    Cell *out = &dbuf[i];
    int living_neighbors = 0;
    FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
      const Cell *nx = &buf[j];
      if (dist_float2(nx->pos, in->pos) > radius) continue;
      living_neighbors += (nx->alive ? 1 : 0);
    })
    if (in->alive) {
      living_neighbors -= 1;
      out->alive = ((living_neighbors >= 2) && (living_neighbors <= 3));
    } else {
      out->alive = (living_neighbors == 3);
    }
    out->pos = in->pos;
  }
  grid_free(&g);
}

int oracle_real_size(void) { return (int)sizeof(abl_float); }

/* ===================================================================================== */
/* predator_prey.abl — PARITY UNPINNED BY THE REFERENCE                                  */
/* ===================================================================================== */
/* The reference `c` backend rejects this model (run-time add/remove: BackendError,
 * src/backend/CBackend.cpp:30-32) and no other reference backend can run in this
 * environment, so nothing pins these semantics numerically.  What is frozen here (and
 * implemented by the CUDA runtime) follows the Mason/FlameGPU printers where they agree:
 *   - `out` starts as a copy of `in` for every step function (MasonPrinter.cpp:480-486);
 *   - step functions run in the listed order, each on the committed state of the previous;
 *   - removeCurrent(): the agent disappears when its step function commits; survivors keep
 *     their order (MasonPrinter.cpp:358-367, FlameGPUPrinter.cpp:501-519);
 *   - add(): at most one new agent per parent and step function; new agents are appended
 *     after the survivors in parent order, get the next ids, and take part from the next
 *     step function on;
 *   - in-step random(): counter-based stream keyed by (seed, timestep, step index, agent id)
 *     (asset/cuda/abl_device.cuh; the reference has no reproducible definition);
 *   - the sequential step runs after all step functions (MasonPrinter.cpp:535-541):
 *     count(T) = live agents, sum(bool member) = number of true values (:623-682).
 * Per-step arithmetic is the reference lowering (GenericCPrinter.cpp:20-75) of the model. */
typedef struct { float2 pos; float2 dir; float2 steer; int life; } PPAnimal;  /* Predator and Prey */
typedef struct { float2 pos; int dead_cycles; bool avail; } PPGrass;

typedef struct {
  void *data; unsigned *ids; int n, cap; unsigned next_id; size_t stride;
} pp_pool;

typedef struct {
  pp_pool pred, prey, grass;
  /* constants as the generated program declares them */
  abl_float PI, SPACE_MULT, REPRODUCE_PREY_PROB, REPRODUCE_PREDATOR_PROB;
  int GAIN_FROM_FOOD_PREDATOR, GAIN_FROM_FOOD_PREY, GRASS_REGROW_CYCLES;
  abl_float PRED_PREY_INTERACTION_RADIUS, PREY_GROUP_COHESION_RADIUS, SAME_SPECIES_AVOIDANCE_RADIUS,
      GRASS_EAT_DISTANCE, PRED_KILL_DISTANCE, DELTA_TIME, PRED_SPEED_ADVANTAGE, env_size;
  double env_exact, granularity_exact;
  int num_agents;
  unsigned timestep;
  uint64_t seed;
} pp_world;

static void pp_pool_init(pp_pool *p, size_t stride) { memset(p, 0, sizeof *p); p->stride = stride; }
static void pp_pool_reserve(pp_pool *p, int want) {
  if (want <= p->cap) return;
  int cap = p->cap ? p->cap : 1024;
  while (cap < want) cap *= 2;
  p->data = realloc(p->data, (size_t)cap * p->stride);
  p->ids = realloc(p->ids, (size_t)cap * sizeof(unsigned));
  p->cap = cap;
}
static void *pp_pool_push(pp_pool *p) {
  pp_pool_reserve(p, p->n + 1);
  p->ids[p->n] = p->next_id++;
  void *slot = (char *)p->data + (size_t)p->n * p->stride;
  memset(slot, 0, p->stride);
  p->n++;
  return slot;
}

static int random_int(int min, int max) { /* libabl.c:26-39 */
  unsigned n = max - min + 1;
  if ((n & (n - 1)) == 0) return xorshift128plus() & (n - 1);
  unsigned r = UINT_MAX % n;
  unsigned x;
  do { x = xorshift128plus(); } while (x >= UINT_MAX - r);
  return min + x % n;
}

/* counter-based in-step RNG, identical to abl_ctx_init / abl_rng_next / random_float of
 * asset/cuda/abl_device.cuh */
static inline uint64_t pp_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static inline uint64_t pp_rng_init(uint64_t seed, unsigned timestep, unsigned step, unsigned id) {
  uint64_t k = pp_mix64(seed ^ 0x9e3779b97f4a7c15ull);
  k = pp_mix64(k + ((uint64_t)timestep << 32 | step));
  return pp_mix64(k + id);
}
static inline abl_float pp_random_float(uint64_t *state, abl_float lo, abl_float hi) {
  *state += 0x9e3779b97f4a7c15ull;
  uint64_t x = pp_mix64(*state);
  return lo + (abl_float)x / ((abl_float)UINT64_MAX / (hi - lo));
}

pp_world *oracle_pp_create(int num_agents) {
  pp_world *w = calloc(1, sizeof *w);
  pp_pool_init(&w->pred, sizeof(PPAnimal));
  pp_pool_init(&w->prey, sizeof(PPAnimal));
  pp_pool_init(&w->grass, sizeof(PPGrass));
  w->num_agents = num_agents;
  w->seed = 0x0123456789abcdefull;
  const double SPACE_MULT = 1.0;
  w->PI = (abl_float)oracle_fold6(3.14159265358979323846);
  w->SPACE_MULT = (abl_float)oracle_fold6(SPACE_MULT);
  w->REPRODUCE_PREY_PROB = (abl_float)0.005;
  w->REPRODUCE_PREDATOR_PROB = (abl_float)0.001;
  w->GAIN_FROM_FOOD_PREDATOR = 35;
  w->GAIN_FROM_FOOD_PREY = 60;
  w->GRASS_REGROW_CYCLES = 60;
  w->PRED_PREY_INTERACTION_RADIUS = (abl_float)oracle_fold6(0.100 * SPACE_MULT);
  w->PREY_GROUP_COHESION_RADIUS = (abl_float)oracle_fold6(0.120 * SPACE_MULT);
  w->SAME_SPECIES_AVOIDANCE_RADIUS = (abl_float)oracle_fold6(0.035 * SPACE_MULT);
  w->GRASS_EAT_DISTANCE = (abl_float)oracle_fold6(0.020 * SPACE_MULT);
  w->PRED_KILL_DISTANCE = (abl_float)oracle_fold6(0.020 * SPACE_MULT);
  w->DELTA_TIME = (abl_float)0.0015;
  w->PRED_SPEED_ADVANTAGE = (abl_float)2.0;
  /* agent_density = 1600 is an integer literal: num_agents / agent_density folds as integer
   * division (see boids_fold) */
  w->env_exact = sqrt((double)(num_agents / 1600)) * SPACE_MULT;
  w->env_size = (abl_float)oracle_fold6(w->env_exact);
  w->granularity_exact = 0.120 * SPACE_MULT; /* largest for-near radius */

  /* main(): sequential population set-up, reference RNG */
  oracle_rng_reset();
  int num_predators = (int)((abl_float)0.025 * num_agents);
  int num_prey = (int)((abl_float)0.35 * num_agents);
  int num_grass = num_agents - num_predators - num_prey;
  for (int i = 0; i < num_grass; i++) {
    abl_float y = random_float(0, w->env_size);
    abl_float x = random_float(0, w->env_size);
    PPGrass *g = pp_pool_push(&w->grass);
    g->pos = float2_create(x, y);
    g->dead_cycles = 0;
    g->avail = true;
  }
  for (int k = 0; k < 2; k++) {
    pp_pool *pool = k == 0 ? &w->prey : &w->pred;
    int count = k == 0 ? num_prey : num_predators;
    int gain = k == 0 ? w->GAIN_FROM_FOOD_PREY : w->GAIN_FROM_FOOD_PREDATOR;
    for (int i = 0; i < count; i++) {
      abl_float y = random_float(0, w->env_size);
      abl_float x = random_float(0, w->env_size);
      abl_float phi = random_float(0, (2.0 * w->PI));
      int life = gain + random_int(0, 10);
      PPAnimal *a = pp_pool_push(pool);
      a->pos = float2_create(x, y);
      a->dir = float2_create(sin(phi), cos(phi));
      a->steer = float2_fill(0);
      a->life = life;
    }
  }
  return w;
}

void oracle_pp_destroy(pp_world *w) {
  free(w->pred.data); free(w->pred.ids);
  free(w->prey.data); free(w->prey.ids);
  free(w->grass.data); free(w->grass.ids);
  free(w);
}

int oracle_pp_count(pp_world *w, int type) { return type == 0 ? w->pred.n : type == 1 ? w->prey.n : w->grass.n; }
int oracle_pp_record_size(int type) { return type == 2 ? (int)sizeof(PPGrass) : (int)sizeof(PPAnimal); }
void oracle_pp_read(pp_world *w, int type, void *records, unsigned *ids) {
  pp_pool *p = type == 0 ? &w->pred : type == 1 ? &w->prey : &w->grass;
  memcpy(records, p->data, (size_t)p->n * p->stride);
  if (ids) memcpy(ids, p->ids, (size_t)p->n * sizeof(unsigned));
}
int oracle_pp_sum_avail(pp_world *w) {
  int s = 0;
  const PPGrass *g = w->grass.data;
  for (int i = 0; i < w->grass.n; i++) s += g[i].avail ? 1 : 0;
  return s;
}

/* commit of one step function: new state becomes current, dead agents are compacted away
 * (stable), new agents are appended in parent order */
static void pp_commit(pp_pool *self, void *next, const bool *dead, pp_pool *target, const bool *added,
                      const void *staged, size_t staged_stride) {
  memcpy(self->data, next, (size_t)self->n * self->stride);
  const int n_before = self->n;
  if (target && added) {
    for (int i = 0; i < n_before; i++) {
      if (!added[i]) continue;
      void *slot = pp_pool_push(target);
      memcpy(slot, (const char *)staged + (size_t)i * staged_stride, staged_stride);
    }
  }
  if (dead) {
    int k = 0;
    const int n_now = self->n; /* includes appended agents when target == self */
    for (int i = 0; i < n_now; i++) {
      if (i < n_before && dead[i]) continue;
      if (k != i) {
        memcpy((char *)self->data + (size_t)k * self->stride, (char *)self->data + (size_t)i * self->stride, self->stride);
        self->ids[k] = self->ids[i];
      }
      k++;
    }
    self->n = k;
  }
}

static inline float2 clamp_4(float2 pos, float2 max) { return clamp_1(pos, float2_fill(0), max); }
static inline bool float2_not_equals(float2 a, float2 b) { return a.x != b.x || a.y != b.y; }
static inline float2 normalize_float2(float2 v) { return float2_div_scalar(v, length_float2(v)); }

#define PP_GRID(g, pool, type)                                                       \
  grid_t g;                                                                          \
  {                                                                                  \
    double lo_[3] = {0, 0, 0}, hi_[3] = {w->env_exact, w->env_exact, 0};             \
    grid_setup(&g, 2, lo_, hi_, w->granularity_exact);                               \
    if (mode == MODE_GRID) grid_bin(&g, (const char *)(pool).data, sizeof(type), (pool).n); \
  }

/* One full timestep: the 13 parallel step functions in schedule order. */
void oracle_pp_timestep(pp_world *w, int mode) {
  const abl_float R_INT = w->PRED_PREY_INTERACTION_RADIUS, R_COH = w->PREY_GROUP_COHESION_RADIUS,
                  R_AVOID = w->SAME_SPECIES_AVOIDANCE_RADIUS, R_EAT = w->GRASS_EAT_DISTANCE,
                  R_KILL = w->PRED_KILL_DISTANCE;
  const abl_float DELTA_TIME = w->DELTA_TIME, SPACE_MULT = w->SPACE_MULT,
                  PRED_SPEED_ADVANTAGE = w->PRED_SPEED_ADVANTAGE, env_size = w->env_size;
  const int reach = 1;

#define ANIMALS(p) ((PPAnimal *)(p).data)
#define GRASSES(p) ((PPGrass *)(p).data)
#define NEXT_OF(pool) void *next = malloc((size_t)((pool).n ? (pool).n : 1) * (pool).stride); \
                      memcpy(next, (pool).data, (size_t)(pool).n * (pool).stride)

  { /* 0: pred_follow_prey (Predator, near Prey) */
    NEXT_OF(w->pred);
    PP_GRID(g, w->prey, PPAnimal);
    const PPAnimal *nbrs = ANIMALS(w->prey);
    const int n = w->prey.n;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < w->pred.n; i++) {
      const PPAnimal *in = &ANIMALS(w->pred)[i];
      PPAnimal *out = &((PPAnimal *)next)[i];
      float2 agent_position = in->pos;
      float2 agent_steer = float2_create(0, 0);
      float2 closest_prey_position = float2_create(0, 0);
      abl_float closest_prey_distance = R_INT;
      bool can_see_prey = false;
      FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
        const PPAnimal *prey = &nbrs[j];
        if (dist_float2(prey->pos, in->pos) > R_INT) continue;
        abl_float separation = length_float2(float2_sub(in->pos, prey->pos));
        if ((separation < closest_prey_distance)) {
          closest_prey_position = prey->pos;
          closest_prey_distance = separation;
          can_see_prey = true;
        }
      })
      if (can_see_prey) agent_steer = float2_sub(closest_prey_position, agent_position);
      out->steer = agent_steer;
    }
    grid_free(&g);
    pp_commit(&w->pred, next, NULL, NULL, NULL, NULL, 0);
    free(next);
  }
  { /* 1: prey_avoid_pred (Prey, near Predator) */
    NEXT_OF(w->prey);
    PP_GRID(g, w->pred, PPAnimal);
    const PPAnimal *nbrs = ANIMALS(w->pred);
    const int n = w->pred.n;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < w->prey.n; i++) {
      const PPAnimal *in = &ANIMALS(w->prey)[i];
      PPAnimal *out = &((PPAnimal *)next)[i];
      float2 avoid_velocity = float2_fill(0.0);
      FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
        const PPAnimal *pred = &nbrs[j];
        if (dist_float2(pred->pos, in->pos) > R_INT) continue;
        abl_float separation = length_float2(float2_sub(in->pos, pred->pos));
        if ((separation < R_INT)) {
          if ((separation > 0.0)) {
            avoid_velocity = float2_add(avoid_velocity, float2_mul_scalar(float2_sub(in->pos, pred->pos), (R_INT / separation)));
          }
        }
      })
      out->steer = avoid_velocity;
    }
    grid_free(&g);
    pp_commit(&w->prey, next, NULL, NULL, NULL, NULL, 0);
    free(next);
  }
  { /* 2: prey_flock (Prey, near Prey) */
    NEXT_OF(w->prey);
    PP_GRID(g, w->prey, PPAnimal);
    const PPAnimal *nbrs = ANIMALS(w->prey);
    const int n = w->prey.n;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < w->prey.n; i++) {
      const PPAnimal *in = &nbrs[i];
      PPAnimal *out = &((PPAnimal *)next)[i];
      float2 group_center = float2_fill(0.0);
      float2 group_velocity = float2_fill(0.0);
      float2 avoid_velocity = float2_fill(0.0);
      int group_centre_count = 0;
      FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
        const PPAnimal *prey = &nbrs[j];
        if (dist_float2(prey->pos, in->pos) > R_COH) continue;
        abl_float separation = length_float2(float2_sub(in->pos, prey->pos));
        group_center = float2_add(group_center, prey->pos);
        group_centre_count += 1;
        if (((separation < R_AVOID) && float2_not_equals(prey->pos, in->pos))) {
          if ((separation > 0.0)) {
            avoid_velocity = float2_add(avoid_velocity, float2_mul_scalar(float2_sub(in->pos, prey->pos), (R_AVOID / separation)));
          }
        }
      })
      if ((group_centre_count > 0)) {
        group_center = float2_div_scalar(group_center, group_centre_count);
        group_velocity = float2_sub(group_center, in->pos);
      }
      out->steer = float2_add(float2_add(in->steer, group_velocity), avoid_velocity);
    }
    grid_free(&g);
    pp_commit(&w->prey, next, NULL, NULL, NULL, NULL, 0);
    free(next);
  }
  { /* 3: pred_avoid (Predator, near Predator) */
    NEXT_OF(w->pred);
    PP_GRID(g, w->pred, PPAnimal);
    const PPAnimal *nbrs = ANIMALS(w->pred);
    const int n = w->pred.n;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < w->pred.n; i++) {
      const PPAnimal *in = &nbrs[i];
      PPAnimal *out = &((PPAnimal *)next)[i];
      float2 avoid_velocity = float2_create(0, 0);
      FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
        const PPAnimal *pred = &nbrs[j];
        if (dist_float2(pred->pos, in->pos) > R_AVOID) continue;
        abl_float separation = length_float2(float2_sub(in->pos, pred->pos));
        if (((separation < R_AVOID) && float2_not_equals(pred->pos, in->pos))) {
          if ((separation > 0.0))
            avoid_velocity = float2_add(avoid_velocity, float2_mul_scalar(float2_sub(in->pos, pred->pos), (R_AVOID / separation)));
        }
      })
      out->steer = float2_add(in->steer, avoid_velocity);
    }
    grid_free(&g);
    pp_commit(&w->pred, next, NULL, NULL, NULL, NULL, 0);
    free(next);
  }
  for (int k = 0; k < 2; k++) { /* 4: prey_move, 5: pred_move */
    pp_pool *pool = k == 0 ? &w->prey : &w->pred;
    NEXT_OF(*pool);
    for (int i = 0; i < pool->n; i++) {
      const PPAnimal *in = &ANIMALS(*pool)[i];
      PPAnimal *out = &((PPAnimal *)next)[i];
      float2 agent_position = in->pos;
      float2 agent_velocity = in->dir;
      float2 agent_steer = in->steer;
      agent_velocity = float2_add(agent_velocity, agent_steer);
      abl_float current_speed = length_float2(agent_velocity);
      if ((current_speed > 1.0)) agent_velocity = normalize_float2(agent_velocity);
      if (k == 0)
        agent_position = float2_add(agent_position, float2_mul_scalar(float2_mul_scalar(agent_velocity, DELTA_TIME), SPACE_MULT));
      else
        agent_position = float2_add(agent_position, float2_mul_scalar(float2_mul_scalar(float2_mul_scalar(agent_velocity, DELTA_TIME), PRED_SPEED_ADVANTAGE), SPACE_MULT));
      agent_position = clamp_4(agent_position, float2_create(env_size, env_size)); /* boundPosition */
      out->pos = agent_position;
      out->dir = agent_velocity;
      out->life = (in->life - 1);
    }
    pp_commit(pool, next, NULL, NULL, NULL, NULL, 0);
    free(next);
  }
  { /* 6: prey_eat_or_starve (Prey, near Grass) */
    NEXT_OF(w->prey);
    bool *dead = calloc((size_t)(w->prey.n ? w->prey.n : 1), 1);
    PP_GRID(g, w->grass, PPGrass);
    const PPGrass *nbrs = GRASSES(w->grass);
    const int n = w->grass.n;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < w->prey.n; i++) {
      const PPAnimal *in = &ANIMALS(w->prey)[i];
      PPAnimal *out = &((PPAnimal *)next)[i];
      int life = in->life;
      FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
        const PPGrass *grass = &nbrs[j];
        if (dist_float2(grass->pos, in->pos) > R_EAT) continue;
        if (grass->avail) life += w->GAIN_FROM_FOOD_PREY;
      })
      out->life = life;
      if ((life < 1)) dead[i] = true;
    }
    grid_free(&g);
    pp_commit(&w->prey, next, dead, NULL, NULL, NULL, 0);
    free(next); free(dead);
  }
  { /* 7: pred_eat_or_starve (Predator, near Prey) */
    NEXT_OF(w->pred);
    bool *dead = calloc((size_t)(w->pred.n ? w->pred.n : 1), 1);
    PP_GRID(g, w->prey, PPAnimal);
    const PPAnimal *nbrs = ANIMALS(w->prey);
    const int n = w->prey.n;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < w->pred.n; i++) {
      const PPAnimal *in = &ANIMALS(w->pred)[i];
      PPAnimal *out = &((PPAnimal *)next)[i];
      int pred_life = in->life;
      FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
        const PPAnimal *prey = &nbrs[j];
        if (dist_float2(prey->pos, in->pos) > R_KILL) continue;
        pred_life += w->GAIN_FROM_FOOD_PREDATOR;
      })
      out->life = pred_life;
      if ((pred_life < 1)) dead[i] = true;
    }
    grid_free(&g);
    pp_commit(&w->pred, next, dead, NULL, NULL, NULL, 0);
    free(next); free(dead);
  }
  { /* 8: grass_eaten (Grass, near Prey) */
    NEXT_OF(w->grass);
    PP_GRID(g, w->prey, PPAnimal);
    const PPAnimal *nbrs = ANIMALS(w->prey);
    const int n = w->prey.n;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < w->grass.n; i++) {
      const PPGrass *in = &GRASSES(w->grass)[i];
      PPGrass *out = &((PPGrass *)next)[i];
      abl_float closest_prey = R_EAT;
      bool eaten = false;
      FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
        const PPAnimal *prey = &nbrs[j];
        if (dist_float2(prey->pos, in->pos) > R_EAT) continue;
        abl_float distance = length_float2(float2_sub(in->pos, prey->pos));
        if (((distance < closest_prey) && in->avail)) {
          closest_prey = distance;
          eaten = true;
        }
      })
      if (eaten) {
        out->dead_cycles = 0;
        out->avail = false;
      }
    }
    grid_free(&g);
    pp_commit(&w->grass, next, NULL, NULL, NULL, NULL, 0);
    free(next);
  }
  { /* 9: prey_eaten (Prey, near Predator) */
    NEXT_OF(w->prey);
    bool *dead = calloc((size_t)(w->prey.n ? w->prey.n : 1), 1);
    PP_GRID(g, w->pred, PPAnimal);
    const PPAnimal *nbrs = ANIMALS(w->pred);
    const int n = w->pred.n;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < w->prey.n; i++) {
      const PPAnimal *in = &ANIMALS(w->prey)[i];
      bool eaten = false;
      abl_float closest_pred = R_KILL;
      FOR_CANDIDATES(&g, mode, n, (const abl_float *)&in->pos, reach, {
        const PPAnimal *pred = &nbrs[j];
        if (dist_float2(pred->pos, in->pos) > R_KILL) continue;
        abl_float distance = length_float2(float2_sub(pred->pos, in->pos));
        if ((distance < closest_pred)) {
          closest_pred = distance;
          eaten = true;
        }
      })
      if (eaten) dead[i] = true;
    }
    grid_free(&g);
    pp_commit(&w->prey, next, dead, NULL, NULL, NULL, 0);
    free(next); free(dead);
  }
  for (int k = 0; k < 2; k++) { /* 10: pred_reproduction, 11: prey_reproduction */
    pp_pool *pool = k == 0 ? &w->pred : &w->prey;
    const unsigned step_index = k == 0 ? 10 : 11;
    const abl_float prob = k == 0 ? w->REPRODUCE_PREDATOR_PROB : w->REPRODUCE_PREY_PROB;
    NEXT_OF(*pool);
    bool *added = calloc((size_t)(pool->n ? pool->n : 1), 1);
    PPAnimal *staged = calloc((size_t)(pool->n ? pool->n : 1), sizeof(PPAnimal));
    for (int i = 0; i < pool->n; i++) {
      const PPAnimal *in = &ANIMALS(*pool)[i];
      PPAnimal *out = &((PPAnimal *)next)[i];
      uint64_t rng = pp_rng_init(w->seed, w->timestep, step_index, pool->ids[i]);
      if ((pp_random_float(&rng, 0, 1.0) < prob)) {
        added[i] = true;
        staged[i].pos = in->pos;
        staged[i].dir = float2_mul_scalar(in->dir, -1.0);
        staged[i].steer = float2_mul_scalar(in->steer, -1.0);
        staged[i].life = (in->life / 2);
        out->life = (in->life / 2);
      }
    }
    pp_commit(pool, next, NULL, pool, added, staged, sizeof(PPAnimal));
    free(next); free(added); free(staged);
  }
  { /* 12: grass_growth */
    NEXT_OF(w->grass);
    for (int i = 0; i < w->grass.n; i++) {
      const PPGrass *in = &GRASSES(w->grass)[i];
      PPGrass *out = &((PPGrass *)next)[i];
      if ((in->dead_cycles == w->GRASS_REGROW_CYCLES)) {
        int cycle_start = 0;
        out->dead_cycles = cycle_start;
        out->avail = true;
      }
      if ((!in->avail)) out->dead_cycles = (in->dead_cycles + 1);
    }
    pp_commit(&w->grass, next, NULL, NULL, NULL, NULL, 0);
    free(next);
  }
  w->timestep++;
}
