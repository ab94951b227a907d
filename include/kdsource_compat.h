/*
 * kdsource_compat.h -- the C sampling interface of KDSource served by the
 * B200 sampler.  Same function names, argument meaning and struct fields as
 * clib/include/kdsource/{kdsource.h.in,geom.h,plist.h,utils.h} of the reference
 * (v0.2.3), so a host that does
 *     KDS_open -> KDS_w_mean(kds, 1000, NULL) -> N x KDS_sample2(...) -> KDS_destroy
 * (python/kdsource/test.py:97-124, legacy McStas/TRIPOLI components) links
 * against libkdsb200.so unchanged.  Particles are generated on the GPU in
 * batches (kdsb200_resample) and handed out one at a time.
 *
 * Differences, by design:
 *  - particles are drawn i.i.d. from the weight distribution (prefix sum +
 *    binary search) instead of the reference's sequential accept/advance walk
 *    (kdsource.c:282-330): same marginal distribution, different sequence;
 *  - the non-`rand` functions use counter-based Philox (seed: KDSB200_SEED,
 *    default 0) instead of rand(); the `rand` variants draw ONE number from
 *    the caller's generator per batch of 65536 particles and mix it into the
 *    Philox key (the counter is the source's running output index);
 *  - a WeightFun bias is evaluated once per list entry (on the host); particles
 *    are then picked with probability ~ w * bias and carry weight 1 / bias, the
 *    distribution the reference's accept/advance walk samples (kdsource.c:294-325).
 * The single-particle functions (X_perturb, Geom_perturb, kds_rand_*, PList_get/next,
 * traslv, rotv) run on the host, as in the reference.
 */
#ifndef KDSOURCE_COMPAT_H
#define KDSOURCE_COMPAT_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MCPL_H
/* layout of libmcpl's mcpl_particle_t (96 bytes) */
typedef struct {
  double ekin;
  double polarisation[3];
  double position[3];
  double direction[3];
  double time;
  double weight;
  int32_t pdgcode;
  uint32_t userflags;
} mcpl_particle_t;
typedef struct { void* internal; } mcpl_file_t;
#endif

#define MAX_RESAMPLES 1000000
#define NAME_MAX_LEN 256

#ifndef KDS_RNG_FCT
#define KDS_RNG_FCT
typedef double (*kds_rng_fct_t)(void*);
#endif
typedef double (*WeightFun)(const mcpl_particle_t* part);

/* ---- plist.h:18-41 ---- */
typedef struct PList {
  char pt;
  int pdgcode;
  char* filename;
  mcpl_file_t file;   /* internal: the decoded list */
  double* trasl;
  double* rot;
  int x2z;
  const mcpl_particle_t* part;
  long long npts;
} PList;
PList* PList_create(char pt, const char* filename, const double* trasl, const double* rot,
                    int switch_x2z);
PList* PList_copy(const PList* from);
int PList_get(const PList* plist, mcpl_particle_t* part);
int PList_next(PList* plist, int loop);
void PList_destroy(PList* plist);

/* ---- geom.h:20-97 ---- */
typedef struct Metric Metric;
typedef int (*PerturbFun)(kds_rng_fct_t, void* rngstate, const Metric* metric,
                          mcpl_particle_t* part, double bw, char kernel);
struct Metric {
  int dim;
  float* scaling;
  PerturbFun perturb;  /* identifies the metric; see the X_perturb tokens below */
  int nps;
  double* params;
};
Metric* Metric_create(int dim, const double* scaling, PerturbFun perturb, int nps,
                      const double* params);
Metric* Metric_copy(const Metric* from);
void Metric_destroy(Metric* metric);
typedef struct Geometry {
  int ord;
  Metric** ms;
  char* bwfilename;
  FILE* bwfile;       /* always NULL here: the bandwidth file is loaded at creation */
  double bw;
  char kernel;
  double* trasl;
  double* rot;
  float* bws;         /* extension: per-particle bandwidths (NULL if constant) */
  long long nbws;
  long long bwpos;    /* extension: position of Geom_next in bws */
} Geometry;
Geometry* Geom_create(int ord, Metric** metrics, double bw, const char* bwfilename, char kernel,
                      const double* trasl, const double* rot);
Geometry* Geom_copy(const Geometry* from);
int Geom_perturb(kds_rng_fct_t, void* rngstate, const Geometry* geom, mcpl_particle_t* part);
int Geom_next(Geometry* geom, int loop);
void Geom_destroy(Geometry* geom);
/* The 14 metric perturbations (geom.c:192-650).  The batched sampler identifies a metric
 * by its function and perturbs on the device; called directly they perturb ONE particle on
 * the host with the caller's rng, the same templated maps compiled for the host. */
#define KDS_DECL_PERTURB(name)                                                            \
  int name(kds_rng_fct_t, void* rngstate, const Metric* metric, mcpl_particle_t* part, \
           double bw, char kernel)
KDS_DECL_PERTURB(E_perturb);
KDS_DECL_PERTURB(Let_perturb);
KDS_DECL_PERTURB(wl_perturb);
KDS_DECL_PERTURB(t_perturb);
KDS_DECL_PERTURB(Dec_perturb);
KDS_DECL_PERTURB(Vol_perturb);
KDS_DECL_PERTURB(SurfXY_perturb);
KDS_DECL_PERTURB(SurfR_perturb);
KDS_DECL_PERTURB(SurfR2_perturb);
KDS_DECL_PERTURB(SurfCircle_perturb);
KDS_DECL_PERTURB(Guide_perturb);
KDS_DECL_PERTURB(Isotrop_perturb);
KDS_DECL_PERTURB(Polar_perturb);
KDS_DECL_PERTURB(PolarMu_perturb);
extern const int _n_metrics;
extern const char* _metric_names[];
extern const PerturbFun _metric_perturbs[];

/* ---- utils.h:22-31 ---- */
double kds_rand_norm(kds_rng_fct_t, void* rngstate);
double kds_rand_epan(kds_rng_fct_t, void* rngstate);
double kds_rand_box(kds_rng_fct_t, void* rngstate);
double kds_rand_type(kds_rng_fct_t, void* rngstate, char kernel);
double* traslv(double* vect, const double* trasl, int inverse);
double* rotv(double* vect, const double* rotvec, int inverse);
int pt2pdg(char pt);
char pdg2pt(int pdgcode);

/* ---- kdsource.h.in:41-92 ---- */
typedef struct KDSource {
  double J;
  char kernel;
  PList* plist;
  Geometry* geom;
  void* impl;         /* extension: device state */
} KDSource;
KDSource* KDS_create(double J, char kernel, PList* plist, Geometry* geom);
KDSource* KDS_open(const char* xmlfilename);
int KDS_rand_sample2(kds_rng_fct_t, void* rngstate, KDSource* kds, mcpl_particle_t* part,
                     int perturb, double w_crit, WeightFun bias, int loop);
int KDS_rand_sample(kds_rng_fct_t, void* rngstate, KDSource* kds, mcpl_particle_t* part);
double KDS_rand_w_mean(kds_rng_fct_t, void* rngstate, KDSource* kds, int N, WeightFun bias);
int KDS_sample2(KDSource* kds, mcpl_particle_t* part, int perturb, double w_crit, WeightFun bias,
                int loop);
int KDS_sample(KDSource* kds, mcpl_particle_t* part);
double KDS_w_mean(KDSource* kds, int N, WeightFun bias);
void KDS_destroy(KDSource* kds);

typedef struct MultiSource {
  int len;
  KDSource** s;
  double J;
  double* ws;
  double* cdf;
} MultiSource;
MultiSource* MS_create(int len, KDSource** s, const double* ws);
MultiSource* MS_open(int len, const char** xmlfilenames, const double* ws);
int MS_rand_sample2(kds_rng_fct_t, void* rngstate, MultiSource* ms, mcpl_particle_t* part,
                    int perturb, double w_crit, WeightFun bias, int loop);
int MS_rand_sample(kds_rng_fct_t, void* rngstate, MultiSource* ms, mcpl_particle_t* part);
double MS_rand_w_mean(kds_rng_fct_t, void* rngstate, MultiSource* ms, int N, WeightFun bias);
int MS_sample2(MultiSource* ms, mcpl_particle_t* part, int perturb, double w_crit, WeightFun bias,
               int loop);
int MS_sample(MultiSource* ms, mcpl_particle_t* part);
double MS_w_mean(MultiSource* ms, int N, WeightFun bias);
void MS_destroy(MultiSource* ms);

/* ---- ctypes entry points: resamplemcpl.c:17, writemcpl.c:15 ---- */
void kdsource_resample_to_mcpl(double (*rng)(void), const char* kds_sourcefile,
                               const char* destination_mcpl, uint64_t nout);
void kdsource_write_mcpl(const char* filename, uint64_t nparticles, unsigned flags,
                         const double* ekin, const double* x, const double* y, const double* z,
                         const double* ux, const double* uy, const double* uz,
                         const double* time, const double* weight, const int32_t* pdgcode,
                         const double* polx, const double* poly, const double* polz,
                         const uint32_t* userflags);

#ifdef __cplusplus
}
#endif
#endif /* KDSOURCE_COMPAT_H */
