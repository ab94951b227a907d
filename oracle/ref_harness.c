/*
 * ref_harness.c -- drives the REFERENCE's own C sources (compiled in place
 * from /root/reference/clib/src by oracle/Makefile into oracle/_ref/) so the
 * oracle restatement can be pinned against them and the reference sampler can
 * be timed / compared statistically.  TEST INFRASTRUCTURE ONLY.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <libxml/parser.h>
#include "kdsource/kdsource.h"

/* ---- in-memory mcpl shim ------------------------------------------- */
typedef struct {
  const mcpl_particle_t *parts;
  uint64_t n, pos;
} memfile;
static const mcpl_particle_t *g_parts = NULL;
static uint64_t g_nparts = 0;

mcpl_file_t mcpl_open_file(const char *filename) {
  (void)filename;
  memfile *m = (memfile *)malloc(sizeof(memfile));
  m->parts = g_parts;
  m->n = g_nparts;
  m->pos = 0;
  mcpl_file_t f;
  f.internal = m;
  return f;
}
uint64_t mcpl_hdr_nparticles(mcpl_file_t f) { return ((memfile *)f.internal)->n; }
const mcpl_particle_t *mcpl_read(mcpl_file_t f) {
  memfile *m = (memfile *)f.internal;
  if (m->pos >= m->n)
    return NULL;
  return &m->parts[m->pos++];
}
void mcpl_rewind(mcpl_file_t f) { ((memfile *)f.internal)->pos = 0; }
void mcpl_close_file(mcpl_file_t f) { free(f.internal); }

/* ---- libxml2 stubs (never reached: KDS_open is not called) ---------- */
int xmlKeepBlanksDefault(int v) { return v; }
xmlDocPtr xmlReadFile(const char *a, const char *b, int c) {
  (void)a; (void)b; (void)c;
  return NULL;
}
xmlNodePtr xmlDocGetRootElement(xmlDocPtr d) { (void)d; return NULL; }
xmlChar *xmlNodeGetContent(xmlNodePtr n) { (void)n; return NULL; }
xmlChar *xmlGetProp(xmlNodePtr n, const xmlChar *p) { (void)n; (void)p; return NULL; }
void xmlFreeDoc(xmlDocPtr d) { (void)d; }
void xmlCleanupParser(void) {}

/* ---- scripted RNG ---------------------------------------------------- */
typedef struct {
  const double *u;
  int n, next;
} script_rng;
static double script_next(void *st) {
  script_rng *s = (script_rng *)st;
  double v = s->u[s->next % s->n];
  s->next++;
  return v;
}

/* xorshift64* for bulk sampling */
typedef struct { uint64_t s; } xs_rng;
static double xs_next(void *st) {
  xs_rng *r = (xs_rng *)st;
  uint64_t x = r->s;
  x ^= x >> 12;
  x ^= x << 25;
  x ^= x >> 27;
  r->s = x;
  return (double)((x * 0x2545F4914F6CDD1DULL) >> 11) * (1.0 / 9007199254740992.0);
}

static void load_part(mcpl_particle_t *p, const double *v8) {
  memset(p, 0, sizeof(*p));
  p->ekin = v8[0];
  p->position[0] = v8[1];
  p->position[1] = v8[2];
  p->position[2] = v8[3];
  p->direction[0] = v8[4];
  p->direction[1] = v8[5];
  p->direction[2] = v8[6];
  p->time = v8[7];
  p->weight = 1.0;
  p->pdgcode = 2112;
}
static void store_part(const mcpl_particle_t *p, double *v8) {
  v8[0] = p->ekin;
  v8[1] = p->position[0];
  v8[2] = p->position[1];
  v8[3] = p->position[2];
  v8[4] = p->direction[0];
  v8[5] = p->direction[1];
  v8[6] = p->direction[2];
  v8[7] = p->time;
}

/* the reference's own X_perturb, selected by index into _metric_perturbs */
int ref_metric_perturb(int type, int dim, const double *scaling, int nps,
                       const double *params, char kernel, double bw,
                       double *part8, const double *uniforms, int nuni) {
  Metric *m = Metric_create(dim, scaling, _metric_perturbs[type], nps, params);
  mcpl_particle_t p;
  load_part(&p, part8);
  script_rng s = {uniforms, nuni, 0};
  m->perturb(script_next, &s, m, &p, bw, kernel);
  store_part(&p, part8);
  Metric_destroy(m);
  return s.next;
}

double ref_rand_type(char kernel, const double *uniforms, int nuni,
                     int *consumed) {
  script_rng s = {uniforms, nuni, 0};
  double v = kds_rand_type(script_next, &s, kernel);
  *consumed = s.next;
  return v;
}

void ref_rotv(double *v, const double *rv, int inverse) { rotv(v, rv, inverse); }
int ref_pt2pdg(char pt) { return pt2pdg(pt); }
int ref_n_metrics(void) { return _n_metrics; }
const char *ref_metric_name(int i) { return _metric_names[i]; }

/*
 * Full reference sampler: KDS_create over an in-memory list, then the exact
 * call sequence of resamplemcpl.c:35-43:
 *   w_crit = KDS_w_mean(kds, 1000, NULL); nout x KDS_rand_sample2(..., 1, w_crit, NULL, 1)
 * parts8/weights/pdg: the raw list (N entries); metric_* describe `order`
 * metrics; bwfile may be NULL/"" for the scalar bw.  rng: xorshift64*(seed),
 * or the scripted uniforms when uniforms != NULL.
 * out8: nout x 8 doubles; out_w: nout weights.
 */
static int run_sampler(kds_rng_fct_t rng, void *st, const double *parts8,
                       const double *weights, const int32_t *pdg, int64_t N, char pt,
                       int order, const int *types, const int *dims,
                       const double *scalings, const int *npss, const double *paramss,
                       char kernel, double bw, const char *bwfile,
                       const double *plist_trasl, const double *plist_rot, int x2z,
                       const double *geom_trasl, const double *geom_rot, int64_t nout,
                       int nskip_wmean, double *out8, double *out_w) {
  mcpl_particle_t *arr =
      (mcpl_particle_t *)calloc((size_t)N, sizeof(mcpl_particle_t));
  for (int64_t i = 0; i < N; i++) {
    load_part(&arr[i], parts8 + i * 8);
    arr[i].weight = weights[i];
    arr[i].pdgcode = pdg[i];
  }
  g_parts = arr;
  g_nparts = (uint64_t)N;
  Metric *ms[16];
  int so = 0, po = 0;
  for (int i = 0; i < order; i++) {
    ms[i] = Metric_create(dims[i], scalings + so, _metric_perturbs[types[i]],
                          npss[i], paramss + po);
    so += dims[i];
    po += npss[i];
  }
  PList *pl = PList_create(pt, "mem", plist_trasl, plist_rot, x2z);
  Geometry *g = Geom_create(order, ms, bw, bwfile, kernel, geom_trasl, geom_rot);
  KDSource *k = KDS_create(1.0, kernel, pl, g);
  double w_crit = KDS_rand_w_mean(rng, st, k, nskip_wmean, NULL);
  mcpl_particle_t part;
  int wraps = 0;
  for (int64_t n = 0; n < nout; n++) {
    wraps += KDS_rand_sample2(rng, st, k, &part, 1, w_crit, NULL, 1);
    if (out8)
      store_part(&part, out8 + n * 8);
    if (out_w)
      out_w[n] = part.weight;
  }
  KDS_destroy(k);
  free(arr);
  g_parts = NULL;
  g_nparts = 0;
  return wraps;
}

int ref_sampler_run(const double *parts8, const double *weights,
                    const int32_t *pdg, int64_t N, char pt, int order,
                    const int *types, const int *dims, const double *scalings,
                    const int *npss, const double *paramss, char kernel,
                    double bw, const char *bwfile, const double *plist_trasl,
                    const double *plist_rot, int x2z, const double *geom_trasl,
                    const double *geom_rot, uint64_t seed,
                    const double *uniforms, int nuni, int64_t nout,
                    int nskip_wmean, double *out8, double *out_w) {
  xs_rng xr = {seed ? seed : 0x9E3779B97F4A7C15ULL};
  script_rng sr = {uniforms, nuni, 0};
  kds_rng_fct_t rng = uniforms ? script_next : xs_next;
  void *st = uniforms ? (void *)&sr : (void *)&xr;
  return run_sampler(rng, st, parts8, weights, pdg, N, pt, order, types, dims, scalings,
                     npss, paramss, kernel, bw, bwfile, plist_trasl, plist_rot, x2z,
                     geom_trasl, geom_rot, nout, nskip_wmean, out8, out_w);
}

/* The same loop driven the way the command-line tool drives it: a STATELESS callback
 * double(*)(void) -- python/kdsource/_chooks.py:29-41 hands random.Random(seed).random
 * through ctypes, one C -> Python call per uniform -- wrapped as in
 * clib/src/resamplemcpl.c:4-26. */
typedef double (*stateless_rng_t)(void);
static double stateless_next(void *st) { return (*(stateless_rng_t *)st)(); }
int ref_sampler_run_stateless(stateless_rng_t rng0, const double *parts8,
                              const double *weights, const int32_t *pdg, int64_t N, char pt,
                              int order, const int *types, const int *dims,
                              const double *scalings, const int *npss,
                              const double *paramss, char kernel, double bw,
                              const char *bwfile, int64_t nout, int nskip_wmean,
                              double *out8, double *out_w) {
  return run_sampler(stateless_next, (void *)&rng0, parts8, weights, pdg, N, pt, order,
                     types, dims, scalings, npss, paramss, kernel, bw, bwfile, NULL, NULL, 0,
                     NULL, NULL, nout, nskip_wmean, out8, out_w);
}
