// C sampling interface of KDSource (include/kdsource_compat.h) on top of the
// batched device sampler.  Host-side glue only: XML source files (read the way
// KDS_open reads them, clib/src/kdsource.c:79-280), MCPL-3 files (own reader and
// writer over zlib; libmcpl is not a dependency), the KDS_*/MS_* handles, and the
// two ctypes entry points of clib/src/resamplemcpl.c:17 and writemcpl.c:15.
// All particle generation happens in kdsb200_resample (resample.cu).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kdsb200.h"
#include "kdsource_compat.h"
#include "perturb.cuh"

namespace {

// ---------------------------------------------------------------------------
// errors (kdsource.c:10-13, 43-46): no error codes, print and exit
// ---------------------------------------------------------------------------
[[noreturn]] void kds_error(const char* msg) {
  printf("KDSource error: %s\n", msg);
  exit(EXIT_FAILURE);
}
[[noreturn]] void kds_end(const char* msg) {
  printf("KDSource terminate: %s\n", msg);
  exit(EXIT_SUCCESS);
}
void cuda_or_die(int rc) {
  if (rc) kds_error(kdsb200_last_error());
}
void cu(cudaError_t e) {
  if (e != cudaSuccess) kds_error(cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------
// minimal XML reader: elements, attributes, text; prolog and comments skipped
// ---------------------------------------------------------------------------
struct XNode {
  std::string name, text;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::vector<XNode*> kids;
  ~XNode() {
    for (XNode* k : kids) delete k;
  }
  const char* attr(const char* key) const {
    for (auto& a : attrs)
      if (a.first == key) return a.second.c_str();
    return nullptr;
  }
  // like xmlNodeGetContent: own text plus descendants'
  std::string content() const {
    std::string s = text;
    for (XNode* k : kids) s += k->content();
    return s;
  }
};

struct XParser {
  const std::string& s;
  size_t p = 0;
  explicit XParser(const std::string& src) : s(src) {}
  void skip_ws() {
    while (p < s.size() && isspace((unsigned char)s[p])) p++;
  }
  bool starts(const char* t) const { return s.compare(p, strlen(t), t) == 0; }
  void skip_misc() {
    for (;;) {
      skip_ws();
      if (starts("<?")) {
        size_t e = s.find("?>", p);
        p = e == std::string::npos ? s.size() : e + 2;
      } else if (starts("<!--")) {
        size_t e = s.find("-->", p);
        p = e == std::string::npos ? s.size() : e + 3;
      } else if (starts("<!")) {
        size_t e = s.find('>', p);
        p = e == std::string::npos ? s.size() : e + 1;
      } else {
        return;
      }
    }
  }
  static std::string unescape(const std::string& t) {
    std::string o;
    for (size_t i = 0; i < t.size(); i++) {
      if (t[i] == '&') {
        if (t.compare(i, 4, "&lt;") == 0) { o += '<'; i += 3; continue; }
        if (t.compare(i, 4, "&gt;") == 0) { o += '>'; i += 3; continue; }
        if (t.compare(i, 5, "&amp;") == 0) { o += '&'; i += 4; continue; }
        if (t.compare(i, 6, "&quot;") == 0) { o += '"'; i += 5; continue; }
        if (t.compare(i, 6, "&apos;") == 0) { o += '\''; i += 5; continue; }
      }
      o += t[i];
    }
    return o;
  }
  XNode* element() {
    skip_misc();
    if (p >= s.size() || s[p] != '<') return nullptr;
    p++;
    XNode* n = new XNode;
    while (p < s.size() && !isspace((unsigned char)s[p]) && s[p] != '>' && s[p] != '/')
      n->name += s[p++];
    for (;;) {
      skip_ws();
      if (p >= s.size()) return n;
      if (s[p] == '/') {  // <name ... />
        p += 2;
        return n;
      }
      if (s[p] == '>') {
        p++;
        break;
      }
      std::string key, val;
      while (p < s.size() && s[p] != '=' && !isspace((unsigned char)s[p])) key += s[p++];
      skip_ws();
      if (p < s.size() && s[p] == '=') p++;
      skip_ws();
      if (p < s.size() && (s[p] == '"' || s[p] == '\'')) {
        const char qc = s[p++];
        while (p < s.size() && s[p] != qc) val += s[p++];
        p++;
      }
      n->attrs.emplace_back(key, unescape(val));
    }
    for (;;) {  // content
      size_t lt = s.find('<', p);
      if (lt == std::string::npos) {
        p = s.size();
        return n;
      }
      n->text += unescape(s.substr(p, lt - p));
      p = lt;
      if (starts("</")) {
        size_t e = s.find('>', p);
        p = e == std::string::npos ? s.size() : e + 1;
        break;
      }
      if (starts("<!--") || starts("<?")) {
        skip_misc();
        continue;
      }
      XNode* k = element();
      if (k) n->kids.push_back(k);
    }
    // xmlKeepBlanksDefault(0): whitespace-only text between children is dropped
    bool blank = true;
    for (char c : n->text) blank = blank && isspace((unsigned char)c);
    if (blank) n->text.clear();
    return n;
  }
};

XNode* parse_xml_file(const char* fn) {
  FILE* f = fopen(fn, "rb");
  if (!f) return nullptr;
  std::string src;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) src.append(buf, n);
  fclose(f);
  XParser xp(src);
  return xp.element();
}

// ---------------------------------------------------------------------------
// MCPL-3 reader / writer (layout: kdsource_b200/mcpl.py docstring)
// ---------------------------------------------------------------------------
struct McplList {
  std::vector<mcpl_particle_t> parts;
  int refs = 1;  // PList_copy shares the decoded list; every PList has its own cursor
};
// what PList.file.internal points to: the list and the read position of this PList
// (index of the next record mcpl_read would return, plist.c:91-109)
struct McplCursor {
  McplList* list;
  uint64_t pos = 0;
};
McplList* list_of(const PList* pl) { return ((McplCursor*)pl->file.internal)->list; }

void unpack_dir(double p0, double p1, double sek, double* d) {
  const double sg = signbit(sek) ? -1.0 : 1.0;
  if (fabs(p0) > 1.0) {
    d[1] = p1;
    d[2] = 1.0 / p0;
    d[0] = sg * sqrt(fmax(0.0, 1.0 - (p1 * p1 + d[2] * d[2])));
  } else if (fabs(p1) > 1.0) {
    d[0] = p0;
    d[2] = 1.0 / p1;
    d[1] = sg * sqrt(fmax(0.0, 1.0 - (p0 * p0 + d[2] * d[2])));
  } else {
    d[0] = p0;
    d[1] = p1;
    d[2] = sg * sqrt(fmax(0.0, 1.0 - (p0 * p0 + p1 * p1)));
  }
}

void pack_dir(const double* d, double ekin, double* p0, double* p1, double* sek) {
  const double ax = fabs(d[0]), ay = fabs(d[1]), az = fabs(d[2]);
  double sg;
  if (az < fmax(ax, ay)) {
    const double invz = d[2] != 0.0 ? 1.0 / d[2] : INFINITY;
    if (ax >= ay) { *p0 = invz; *p1 = d[1]; sg = d[0]; }
    else { *p0 = d[0]; *p1 = invz; sg = d[1]; }
  } else {
    *p0 = d[0]; *p1 = d[1]; sg = d[2];
  }
  *sek = copysign(fabs(ekin), sg);
}

bool gz_read_exact(gzFile f, void* dst, size_t n) {
  char* c = (char*)dst;
  while (n) {
    const int chunk = n > (1u << 30) ? (1 << 30) : (int)n;
    const int r = gzread(f, c, chunk);
    if (r <= 0) return false;
    c += r;
    n -= r;
  }
  return true;
}

McplList* mcpl_load(const char* filename) {
  gzFile f = gzopen(filename, "rb");  // also reads uncompressed files
  if (!f) {
    printf("Could not open file %s\n", filename);
    kds_error("Error in PList_create");
  }
  gzbuffer(f, 1 << 20);
  unsigned char magic[8];
  uint64_t np;
  uint32_t a[8];
  if (!gz_read_exact(f, magic, 8) || memcmp(magic, "MCPL", 4) != 0 || magic[7] != 'L' ||
      !gz_read_exact(f, &np, 8) || !gz_read_exact(f, a, 32))
    kds_error("not a little-endian MCPL file");
  const uint32_t ncomments = a[0], nblobs = a[1];
  const bool uflags = a[2], pol = a[3], single = a[4];
  const int32_t updg = (int32_t)a[5];
  const uint32_t psize = a[6];
  double uweight = 0.0;
  if (a[7] && !gz_read_exact(f, &uweight, 8)) kds_error("truncated MCPL header");
  auto skip_string = [&]() {
    uint32_t len;
    if (!gz_read_exact(f, &len, 4)) kds_error("truncated MCPL header");
    std::vector<char> tmp(len);
    if (len && !gz_read_exact(f, tmp.data(), len)) kds_error("truncated MCPL header");
  };
  skip_string();  // source name
  for (uint32_t i = 0; i < ncomments; i++) skip_string();
  for (uint32_t i = 0; i < 2 * nblobs; i++) skip_string();  // keys then blobs
  const size_t fp = single ? 4 : 8;
  const size_t expect = (pol ? 3 : 0) * fp + 7 * fp + (a[7] ? 0 : fp) + (updg ? 0 : 4) + (uflags ? 4 : 0);
  if (expect != psize) kds_error("MCPL particle size does not match header options");
  McplList* L = new McplList;
  L->parts.resize(np);
  const size_t chunk_n = 1 << 16;
  std::vector<unsigned char> buf(chunk_n * psize);
  for (uint64_t done = 0; done < np;) {
    const size_t n = (size_t)std::min<uint64_t>(chunk_n, np - done);
    if (!gz_read_exact(f, buf.data(), n * psize)) kds_error("truncated MCPL particle data");
    for (size_t i = 0; i < n; i++) {
      const unsigned char* r = buf.data() + i * psize;
      auto rd = [&](size_t k) -> double {
        if (single) {
          float v;
          memcpy(&v, r + k * 4, 4);
          return (double)v;
        }
        double v;
        memcpy(&v, r + k * 8, 8);
        return v;
      };
      mcpl_particle_t& p = L->parts[done + i];
      memset(&p, 0, sizeof(p));
      size_t k = 0;
      if (pol)
        for (int j = 0; j < 3; j++) p.polarisation[j] = rd(k++);
      for (int j = 0; j < 3; j++) p.position[j] = rd(k++);
      const double p0 = rd(k), p1 = rd(k + 1), sek = rd(k + 2);
      k += 3;
      unpack_dir(p0, p1, sek, p.direction);
      p.ekin = fabs(sek);
      p.time = rd(k++);
      p.weight = a[7] ? uweight : rd(k++);
      size_t off = k * fp;
      if (updg) {
        p.pdgcode = updg;
      } else {
        memcpy(&p.pdgcode, r + off, 4);
        off += 4;
      }
      if (uflags) memcpy(&p.userflags, r + off, 4);
    }
    done += n;
  }
  gzclose(f);
  return L;
}

std::string mcpl_gz_name(const char* filename) {
  // mcpl_name_helper(fn,'B') + mcpl_create_outfile + mcpl_closeandgzip_outfile:
  // strip .gz / .mcpl, then <bare>.mcpl.gz
  std::string s(filename);
  if (s.size() > 3 && s.compare(s.size() - 3, 3, ".gz") == 0) s.resize(s.size() - 3);
  if (s.size() > 5 && s.compare(s.size() - 5, 5, ".mcpl") == 0) s.resize(s.size() - 5);
  return s + ".mcpl.gz";
}

void gz_write_all(gzFile f, const void* src, size_t n) {
  const char* c = (const char*)src;
  while (n) {
    const unsigned chunk = n > (1u << 30) ? (1u << 30) : (unsigned)n;
    if (gzwrite(f, c, chunk) <= 0) kds_error("could not write MCPL file");
    c += chunk;
    n -= chunk;
  }
}

std::vector<unsigned char> mcpl_header_bytes(uint64_t np, const char* srcname, bool single,
                                            bool pol, int32_t updg, bool has_uw, double uweight,
                                            bool uflags) {
  const uint32_t fp = single ? 4 : 8;
  const uint32_t psize = (pol ? 3 : 0) * fp + 7 * fp + (has_uw ? 0 : fp) + (updg ? 0 : 4) +
                         (uflags ? 4 : 0);
  std::vector<unsigned char> h;
  auto put = [&](const void* p, size_t n) {
    h.insert(h.end(), (const unsigned char*)p, (const unsigned char*)p + n);
  };
  put("MCPL003L", 8);
  put(&np, 8);
  const uint32_t a[8] = {0, 0, uflags, pol, single, (uint32_t)updg, psize, has_uw};
  put(a, 32);
  if (has_uw) put(&uweight, 8);
  const uint32_t len = (uint32_t)strlen(srcname);
  put(&len, 4);
  put(srcname, len);
  return h;
}

void mcpl_write_header(gzFile f, uint64_t np, const char* srcname, bool single, bool pol,
                       int32_t updg, bool has_uw, double uweight, bool uflags) {
  const std::vector<unsigned char> h =
      mcpl_header_bytes(np, srcname, single, pol, updg, has_uw, uweight, uflags);
  gz_write_all(f, h.data(), h.size());
}

// rotv (Rodrigues, axis-angle; utils.c:82-103), host, used once when a list is loaded
void rotv_host(double* v, const double* rv, int inverse) {
  double theta = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
  if (theta == 0) return;
  if (inverse) theta = -theta;
  const double kd = rv[0] * v[0] + rv[1] * v[1] + rv[2] * v[2];
  const double kx[3] = {rv[1] * v[2] - rv[2] * v[1], rv[2] * v[0] - rv[0] * v[2],
                        rv[0] * v[1] - rv[1] * v[0]};
  double o[3];
  for (int i = 0; i < 3; i++)
    o[i] = v[i] * cos(theta) + (kx[i] / fabs(theta)) * sin(theta) +
           rv[i] * kd / (theta * theta) * (1 - cos(theta));
  for (int i = 0; i < 3; i++) v[i] = o[i];
}

double* dup3(const double* v) {
  if (!v) return nullptr;
  double* o = (double*)malloc(3 * sizeof(double));
  memcpy(o, v, 3 * sizeof(double));
  return o;
}

// ---------------------------------------------------------------------------
// device state of one KDSource
// ---------------------------------------------------------------------------
constexpr uint64_t BATCH = 1 << 16;

struct Impl {
  int64_t N = 0;                         // valid particles
  std::vector<uint32_t> src_index;       // valid -> index in the raw list
  std::vector<double> w;                 // weights of the valid particles
  void* d_src = nullptr;
  int src_f32 = 1;
  uint64_t* d_cdf = nullptr;
  uint32_t* d_guide = nullptr;
  int guide_bits = 0;
  float* d_bw = nullptr;
  kdsb200_geom geom;
  double* d_parts = nullptr;
  int64_t* d_idx = nullptr;
  std::vector<double> h_parts;
  std::vector<int64_t> h_idx;
  uint64_t have = 0, used = 0;           // particles in the host batch / handed out
  int batch_perturb = -1, batch_seq = -1;
  uint64_t seed = 0, counter = 0;        // Philox key and global output index
  int contract = 2;                      // uniform-consumption contract (KDSB200_RESAMPLE_CONTRACT)
  uint64_t cursor = 0;                   // list cursor (sequential mode, w_mean)
  cudaStream_t stream = nullptr;
  // WeightFun bias (kdsource.c:294-325): the pick follows w_i * bias(part_i); the CDF of
  // the most recently used bias function is kept
  WeightFun bias_fn = nullptr;
  std::vector<double> bias_vals;         // bias(part_i) of the valid particles
  uint64_t* d_cdf_bias = nullptr;
  uint32_t* d_guide_bias = nullptr;
  WeightFun batch_bias = nullptr;        // bias the pending host batch was drawn with
  std::vector<double> parts8;            // valid particles after PList_get (host copy)
};

int metric_type_of(PerturbFun f) {
  for (int i = 0; i < _n_metrics; i++)
    if (_metric_perturbs[i] == f) return i;
  kds_error("unknown metric perturbation function");
}

void build_cdf(Impl* I, const std::vector<double>& w, uint64_t** d_cdf, uint32_t** d_guide);

Impl* ensure_impl(KDSource* kds) {
  if (kds->impl) return (Impl*)kds->impl;
  Impl* I = new Impl;
  PList* pl = kds->plist;
  Geometry* g = kds->geom;
  McplList* L = list_of(pl);
  std::vector<double>& parts8 = I->parts8;
  for (size_t i = 0; i < L->parts.size(); i++) {
    const mcpl_particle_t& p = L->parts[i];
    if (!(p.weight > 0 && p.pdgcode == pl->pdgcode)) continue;  // PList_next plist.c:105
    double pos[3] = {p.position[0], p.position[1], p.position[2]};
    double dir[3] = {p.direction[0], p.direction[1], p.direction[2]};
    if (pl->trasl)  // PList_get plist.c:69-89
      for (int k = 0; k < 3; k++) pos[k] += pl->trasl[k];
    if (pl->rot) {
      rotv_host(pos, pl->rot, 0);
      rotv_host(dir, pl->rot, 0);
    }
    if (pl->x2z) {
      const double x = pos[0], y = pos[1], z = pos[2], dx = dir[0], dy = dir[1], dz = dir[2];
      pos[0] = y; pos[1] = z; pos[2] = x;
      dir[0] = dy; dir[1] = dz; dir[2] = dx;
    }
    const double row[8] = {p.ekin, pos[0], pos[1], pos[2], dir[0], dir[1], dir[2], p.time};
    parts8.insert(parts8.end(), row, row + 8);
    I->w.push_back(p.weight);
    I->src_index.push_back((uint32_t)i);
  }
  I->N = (int64_t)I->w.size();
  if (I->N == 0) kds_error("Error in PList_next: Could not get particle");
  memset(&I->geom, 0, sizeof(I->geom));
  if (g->ord > KDSB200_MAX_METRICS) kds_error("too many metrics in geometry");
  I->geom.order = g->ord;
  for (int i = 0; i < g->ord; i++) {
    const Metric* m = g->ms[i];
    kdsb200_metric& d = I->geom.ms[i];
    d.type = metric_type_of(m->perturb);
    d.dim = m->dim;
    for (int k = 0; k < m->dim && k < 6; k++) d.scaling[k] = m->scaling[k];
    d.nps = m->nps;
    for (int k = 0; k < m->nps && k < 8; k++) d.params[k] = m->params[k];
  }
  I->geom.kernel = g->kernel;
  I->geom.has_trasl = g->trasl != nullptr;
  I->geom.has_rot = g->rot != nullptr;
  for (int k = 0; k < 3; k++) {
    I->geom.trasl[k] = g->trasl ? g->trasl[k] : 0.0;
    I->geom.rot[k] = g->rot ? g->rot[k] : 0.0;
  }
  cu(cudaStreamCreate(&I->stream));
  void* st = (void*)I->stream;
  double* d_tmp;
  cu(cudaMalloc(&d_tmp, parts8.size() * sizeof(double)));
  cu(cudaMemcpy(d_tmp, parts8.data(), parts8.size() * sizeof(double), cudaMemcpyHostToDevice));
  I->src_f32 = kdsb200_resample_fp32_ok(&I->geom);  // else the fp64 store (see kdsb200.h)
  cu(cudaMalloc(&I->d_src, (size_t)I->N * 8 * (I->src_f32 ? sizeof(float) : sizeof(double))));
  cuda_or_die(kdsb200_convert_particles(d_tmp, I->N, I->src_f32, I->d_src, st));
  I->guide_bits = 4;
  while (I->guide_bits < 24 && (1ll << I->guide_bits) * 8 < I->N) I->guide_bits++;
  build_cdf(I, I->w, &I->d_cdf, &I->d_guide);
  if (g->bws) {
    if (g->nbws != I->N)
      printf("Warning: Particle list and bandwidths file have different size.\n");
    std::vector<float> bw((size_t)I->N);
    for (int64_t i = 0; i < I->N; i++) bw[i] = g->bws[i % g->nbws];  // Geom_next rewinds
    cu(cudaMalloc(&I->d_bw, (size_t)I->N * sizeof(float)));
    cu(cudaMemcpy(I->d_bw, bw.data(), (size_t)I->N * sizeof(float), cudaMemcpyHostToDevice));
  }
  cu(cudaMalloc(&I->d_parts, BATCH * 8 * sizeof(double)));
  cu(cudaMalloc(&I->d_idx, BATCH * sizeof(int64_t)));
  I->h_parts.resize(BATCH * 8);
  I->h_idx.resize(BATCH);
  const char* e = getenv("KDSB200_SEED");
  I->seed = e ? strtoull(e, nullptr, 10) : 0;
  const char* ec = getenv("KDSB200_RESAMPLE_CONTRACT");
  I->contract = (ec && atoi(ec) == 1) ? 1 : 2;
  cu(cudaStreamSynchronize(I->stream));
  cudaFree(d_tmp);
  kds->impl = I;
  return I;
}

// weights -> device CDF + guide table
void build_cdf(Impl* I, const std::vector<double>& w, uint64_t** d_cdf, uint32_t** d_guide) {
  void* st = (void*)I->stream;
  double sum_w = 0.0;
  for (double v : w) sum_w += v;
  if (!(sum_w > 0)) kds_error("all particles have zero weight");
  double* d_w;
  uint64_t* d_scratch;
  cu(cudaMalloc(&d_w, (size_t)I->N * sizeof(double)));
  cu(cudaMemcpy(d_w, w.data(), (size_t)I->N * sizeof(double), cudaMemcpyHostToDevice));
  cu(cudaMalloc(&d_scratch, kdsb200_cdf_scratch_words(I->N) * sizeof(uint64_t)));
  if (!*d_cdf) cu(cudaMalloc(d_cdf, (size_t)I->N * sizeof(uint64_t)));
  cuda_or_die(kdsb200_weights_cdf(d_w, I->N, sum_w, d_scratch, *d_cdf, st));
  if (!*d_guide) cu(cudaMalloc(d_guide, ((1ull << I->guide_bits) + 1) * sizeof(uint32_t)));
  cuda_or_die(kdsb200_cdf_guide(*d_cdf, I->N, I->guide_bits, *d_guide, st));
  cu(cudaStreamSynchronize(I->stream));
  cudaFree(d_w);
  cudaFree(d_scratch);
}

// the particle PList_get would hand out for valid entry i (transformed, with its weight)
void valid_particle(const KDSource* kds, const Impl* I, int64_t i, mcpl_particle_t* part) {
  *part = list_of(kds->plist)->parts[I->src_index[i]];
  const double* o = &I->parts8[(size_t)i * 8];
  part->ekin = o[0];
  for (int k = 0; k < 3; k++) {
    part->position[k] = o[1 + k];
    part->direction[k] = o[4 + k];
  }
  part->time = o[7];
}

// CDF of w_i * bias(part_i) (the distribution the reference's accept/advance walk with a
// bias function converges to, kdsource.c:294-321)
void ensure_bias(KDSource* kds, Impl* I, WeightFun bias) {
  if (I->bias_fn == bias) return;
  I->bias_vals.resize((size_t)I->N);
  std::vector<double> wb((size_t)I->N);
  mcpl_particle_t part;
  for (int64_t i = 0; i < I->N; i++) {
    valid_particle(kds, I, i, &part);
    const double b = bias(&part);
    I->bias_vals[i] = b;
    wb[i] = b > 0 ? I->w[i] * b : 0.0;
  }
  build_cdf(I, wb, &I->d_cdf_bias, &I->d_guide_bias);
  I->bias_fn = bias;
}

void free_impl(Impl* I) {
  if (!I) return;
  cudaFree(I->d_cdf_bias);
  cudaFree(I->d_guide_bias);
  cudaFree(I->d_src);
  cudaFree(I->d_cdf);
  cudaFree(I->d_guide);
  cudaFree(I->d_bw);
  cudaFree(I->d_parts);
  cudaFree(I->d_idx);
  if (I->stream) cudaStreamDestroy(I->stream);
  delete I;
}

double default_rng(void*) { return rand() / (RAND_MAX + 1.0); }  // kdsource.c:37-41

uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

void fill_batch(KDSource* kds, Impl* I, kds_rng_fct_t rng, void* st, int perturb, bool seq,
                WeightFun bias) {
  // Philox counter = global output index of this source (never reused); a caller-supplied
  // generator contributes one draw per batch to the key
  uint64_t seed = I->seed;
  const uint64_t first = I->counter;
  if (rng != default_rng) {
    const double u = rng(st);
    uint64_t bits;
    memcpy(&bits, &u, sizeof(bits));
    seed ^= splitmix64(bits);
  }
  int64_t seq_start = -1;
  if (seq) {
    seq_start = (int64_t)(I->cursor % (uint64_t)I->N) - (int64_t)(first % (uint64_t)I->N);
    if (seq_start < 0) seq_start += I->N;
  }
  const bool biased = bias && !seq;
  cuda_or_die(kdsb200_resample(
      I->d_src, I->src_f32, biased ? I->d_cdf_bias : I->d_cdf, biased ? I->d_guide_bias : I->d_guide,
      I->guide_bits, 0, I->d_bw, kds->geom->bw, I->N, &I->geom, kds->plist->pdgcode, seed, first,
      BATCH, seq_start, nullptr, 0, perturb, I->contract, nullptr, I->d_idx, I->d_parts,
      (void*)I->stream));
  cu(cudaMemcpyAsync(I->h_parts.data(), I->d_parts, BATCH * 8 * sizeof(double),
                     cudaMemcpyDeviceToHost, I->stream));
  cu(cudaMemcpyAsync(I->h_idx.data(), I->d_idx, BATCH * sizeof(int64_t), cudaMemcpyDeviceToHost,
                     I->stream));
  cu(cudaStreamSynchronize(I->stream));
  I->counter += BATCH;
  I->have = BATCH;
  I->used = 0;
  I->batch_perturb = perturb;
  I->batch_seq = seq;
  I->batch_bias = biased ? bias : nullptr;
}

// uniforms from the caller's generator, for the single-particle host entry points
struct HostStream {
  static constexpr bool EXT = true;
  static constexpr bool V2 = false;
  kds_rng_fct_t rng;
  void* st;
  // __host__ __device__ only so that the shared templates compile without warnings;
  // this stream exists on the host alone
  __host__ __device__ double ext_next() { return rng(st); }
  __host__ __device__ int kint() { return 0; }  // Philox-only member, never reached with EXT
};

// X_perturb on the host (geom.c:192-650): the same templated maps the device sampler
// runs (perturb.cuh), in fp64, fed by rng(rngstate) one uniform at a time
int host_perturb(int type, kds_rng_fct_t rng, void* st, const Metric* m, mcpl_particle_t* part,
                 double bw, char kernel) {
  kdsb200_metric km;
  memset(&km, 0, sizeof(km));
  km.type = type;
  km.dim = m->dim;
  for (int k = 0; k < m->dim && k < 6; k++) km.scaling[k] = m->scaling[k];
  km.nps = m->nps;
  for (int k = 0; k < m->nps && k < 8; k++) km.params[k] = m->params[k];
  kdsb::Part<double> p;
  p.ekin = part->ekin;
  p.time = part->time;
  for (int k = 0; k < 3; k++) {
    p.pos[k] = part->position[k];
    p.dir[k] = part->direction[k];
  }
  HostStream s{rng ? rng : default_rng, st};
  kdsb::metric_perturb<double, HostStream, -1, false>(s, km, p, bw, kernel);
  part->ekin = p.ekin;
  part->time = p.time;
  for (int k = 0; k < 3; k++) {
    part->position[k] = p.pos[k];
    part->direction[k] = p.dir[k];
  }
  return 0;
}

}  // namespace

// ===========================================================================
// exported C interface
// ===========================================================================
extern "C" {

const int _n_metrics = 14;
const char* _metric_names[] = {"Energy", "Lethargy", "Wavelength", "Vol", "SurfXY", "SurfR",
                               "SurfR2", "SurfCircle", "Guide", "Isotrop", "Polar", "PolarMu",
                               "Time", "Decade"};

#define KDS_PERTURB(name, type)                                                           \
  int name(kds_rng_fct_t rng, void* rngstate, const Metric* metric, mcpl_particle_t* part, \
           double bw, char kernel) {                                                       \
    return host_perturb(type, rng, rngstate, metric, part, bw, kernel);                    \
  }
KDS_PERTURB(E_perturb, kdsb::M_ENERGY)
KDS_PERTURB(Let_perturb, kdsb::M_LETHARGY)
KDS_PERTURB(wl_perturb, kdsb::M_WAVELENGTH)
KDS_PERTURB(t_perturb, kdsb::M_TIME)
KDS_PERTURB(Dec_perturb, kdsb::M_DECADE)
KDS_PERTURB(Vol_perturb, kdsb::M_VOL)
KDS_PERTURB(SurfXY_perturb, kdsb::M_SURFXY)
KDS_PERTURB(SurfR_perturb, kdsb::M_SURFR)
KDS_PERTURB(SurfR2_perturb, kdsb::M_SURFR2)
KDS_PERTURB(SurfCircle_perturb, kdsb::M_SURFCIRCLE)
KDS_PERTURB(Guide_perturb, kdsb::M_GUIDE)
KDS_PERTURB(Isotrop_perturb, kdsb::M_ISOTROP)
KDS_PERTURB(Polar_perturb, kdsb::M_POLAR)
KDS_PERTURB(PolarMu_perturb, kdsb::M_POLARMU)
const PerturbFun _metric_perturbs[] = {
    E_perturb, Let_perturb, wl_perturb, Vol_perturb, SurfXY_perturb, SurfR_perturb,
    SurfR2_perturb, SurfCircle_perturb, Guide_perturb, Isotrop_perturb, Polar_perturb,
    PolarMu_perturb, t_perturb, Dec_perturb};

// ---- utils.c:25-103 ------------------------------------------------------------
double kds_rand_norm(kds_rng_fct_t rng, void* rngstate) {
  HostStream s{rng ? rng : default_rng, rngstate};
  return kdsb::rand_norm<double, HostStream, false>(s);
}
double kds_rand_epan(kds_rng_fct_t rng, void* rngstate) {
  HostStream s{rng ? rng : default_rng, rngstate};
  return kdsb::rand_epan<double, HostStream>(s);
}
double kds_rand_box(kds_rng_fct_t rng, void* rngstate) {
  return (rng ? rng : default_rng)(rngstate) * 2.0 - 1.0;
}
double kds_rand_type(kds_rng_fct_t rng, void* rngstate, char kernel) {
  if (kernel != 'g' && kernel != 'e' && kernel != 'b') {
    printf("Cannot perturbate with current kernel. \n");
    return 0;
  }
  HostStream s{rng ? rng : default_rng, rngstate};
  return kdsb::rand_type<double, HostStream, false>(s, kernel);
}
double* traslv(double* vect, const double* trasl, int inverse) {
  for (int i = 0; i < 3; i++) vect[i] += inverse ? -trasl[i] : trasl[i];
  return vect;
}
double* rotv(double* vect, const double* rotvec, int inverse) {
  kdsb::rotv<double>(vect, rotvec, inverse);
  return vect;
}

int pt2pdg(char pt) { return pt == 'n' ? 2112 : pt == 'p' ? 22 : pt == 'e' ? 11 : 0; }
char pdg2pt(int pdg) { return pdg == 2112 ? 'n' : pdg == 22 ? 'p' : pdg == 11 ? 'e' : '0'; }

// ---- PList (plist.c:15-117) ------------------------------------------------
PList* PList_create(char pt, const char* filename, const double* trasl, const double* rot,
                    int switch_x2z) {
  PList* pl = (PList*)calloc(1, sizeof(PList));
  pl->pt = pt;
  pl->pdgcode = pt2pdg(pt);
  McplCursor* c = new McplCursor;
  c->list = mcpl_load(filename);
  pl->npts = (long long)c->list->parts.size();
  pl->filename = (char*)malloc(NAME_MAX_LEN);
  strncpy(pl->filename, filename, NAME_MAX_LEN - 1);
  pl->filename[NAME_MAX_LEN - 1] = 0;
  pl->file.internal = c;
  pl->trasl = dup3(trasl);
  pl->rot = dup3(rot);
  pl->x2z = switch_x2z;
  pl->part = nullptr;
  PList_next(pl, 1);
  return pl;
}

PList* PList_copy(const PList* from) {  // plist.c:47-67: same file, its own cursor
  PList* pl = (PList*)calloc(1, sizeof(PList));
  *pl = *from;
  pl->filename = (char*)malloc(NAME_MAX_LEN);
  strncpy(pl->filename, from->filename, NAME_MAX_LEN - 1);
  pl->filename[NAME_MAX_LEN - 1] = 0;
  McplCursor* c = new McplCursor;
  c->list = list_of(from);
  c->list->refs++;
  pl->file.internal = c;
  pl->trasl = dup3(from->trasl);
  pl->rot = dup3(from->rot);
  pl->part = nullptr;
  PList_next(pl, 1);
  return pl;
}

int PList_get(const PList* plist, mcpl_particle_t* part) {  // plist.c:69-89
  *part = *plist->part;
  if (plist->trasl) traslv(part->position, plist->trasl, 0);
  if (plist->rot) {
    rotv(part->position, plist->rot, 0);
    rotv(part->direction, plist->rot, 0);
  }
  if (plist->x2z) {
    const double x = part->position[0], y = part->position[1], z = part->position[2];
    part->position[0] = y;
    part->position[1] = z;
    part->position[2] = x;
    const double dx = part->direction[0], dy = part->direction[1], dz = part->direction[2];
    part->direction[0] = dy;
    part->direction[1] = dz;
    part->direction[2] = dx;
  }
  return 0;
}

int PList_next(PList* plist, int loop) {  // plist.c:91-109
  McplCursor* c = (McplCursor*)plist->file.internal;
  const std::vector<mcpl_particle_t>& parts = c->list->parts;
  int ret = 0, resamples = 0;
  while (resamples++ < MAX_RESAMPLES) {
    if (c->pos >= parts.size()) {  // end of file: rewind
      if (!loop) kds_end("PList_next: End of particle list reached.");
      ret = 1;
      c->pos = 0;
      if (parts.empty()) kds_error("Error in PList_next: Could not get particle");
    }
    plist->part = &parts[c->pos++];
    if (plist->part->weight > 0 && plist->part->pdgcode == plist->pdgcode) return ret;
  }
  kds_error("Error in PList_next: MAX_RESAMPLES reached.");
}

void PList_destroy(PList* pl) {
  if (!pl) return;
  McplCursor* c = (McplCursor*)pl->file.internal;
  if (c) {
    if (--c->list->refs == 0) delete c->list;
    delete c;
  }
  free(pl->filename);
  free(pl->trasl);
  free(pl->rot);
  free(pl);
}

// ---- Metric / Geometry (geom.c:30-190) --------------------------------------
Metric* Metric_create(int dim, const double* scaling, PerturbFun perturb, int nps,
                      const double* params) {
  if (dim < 0 || nps < 0) kds_error("trying to allocate space for negative number of elements");
  Metric* m = (Metric*)calloc(1, sizeof(Metric));
  m->dim = dim;
  m->scaling = (float*)calloc(dim ? dim : 1, sizeof(float));
  for (int i = 0; i < dim; i++) m->scaling[i] = scaling ? (float)scaling[i] : 0.f;
  m->perturb = perturb;
  m->nps = nps;
  m->params = (double*)calloc(nps ? nps : 1, sizeof(double));
  for (int i = 0; i < nps; i++) m->params[i] = params[i];
  return m;
}

Metric* Metric_copy(const Metric* from) {  // geom.c:50-61
  Metric* m = (Metric*)calloc(1, sizeof(Metric));
  *m = *from;
  m->scaling = (float*)calloc(from->dim ? from->dim : 1, sizeof(float));
  for (int i = 0; i < from->dim; i++) m->scaling[i] = from->scaling[i];
  m->params = (double*)calloc(from->nps ? from->nps : 1, sizeof(double));
  for (int i = 0; i < from->nps; i++) m->params[i] = from->params[i];
  return m;
}

void Metric_destroy(Metric* m) {
  if (!m) return;
  free(m->scaling);
  free(m->params);
  free(m);
}

Geometry* Geom_create(int ord, Metric** metrics, double bw, const char* bwfilename, char kernel,
                      const double* trasl, const double* rot) {
  Geometry* g = (Geometry*)calloc(1, sizeof(Geometry));
  g->ord = ord;
  g->ms = (Metric**)calloc(ord > 0 ? ord : 1, sizeof(Metric*));
  for (int i = 0; i < ord; i++) g->ms[i] = metrics[i];
  g->bw = bw;
  g->kernel = kernel;
  if (bwfilename && strlen(bwfilename)) {
    FILE* f = fopen(bwfilename, "rb");
    if (!f) {
      printf("Could not open file %s\n", bwfilename);
      kds_error("Error in Geom_create");
    }
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    g->nbws = bytes / (long)sizeof(float);
    if (g->nbws < 1) kds_error("Error in Geom_next: Could not read BW.");
    g->bws = (float*)malloc((size_t)g->nbws * sizeof(float));
    if (fread(g->bws, sizeof(float), (size_t)g->nbws, f) != (size_t)g->nbws)
      kds_error("Error in Geom_next: Could not read BW.");
    fclose(f);
    g->bwfilename = (char*)malloc(NAME_MAX_LEN);
    strncpy(g->bwfilename, bwfilename, NAME_MAX_LEN - 1);
    g->bwfilename[NAME_MAX_LEN - 1] = 0;
    Geom_next(g, 1);  // geom.c:96: the first bandwidth is read at creation
  }
  g->trasl = dup3(trasl);
  g->rot = dup3(rot);
  return g;
}

// geom.c:108-137.  The reference's copy loses a variable bandwidth (it tests the file
// handle it has just cleared); here the copy keeps the bandwidth list and restarts it.
Geometry* Geom_copy(const Geometry* from) {
  Geometry* g = (Geometry*)calloc(1, sizeof(Geometry));
  *g = *from;
  g->ms = (Metric**)calloc(from->ord > 0 ? from->ord : 1, sizeof(Metric*));
  for (int i = 0; i < from->ord; i++) g->ms[i] = Metric_copy(from->ms[i]);
  g->bwfilename = nullptr;
  g->bwfile = nullptr;
  g->bws = nullptr;
  g->bwpos = 0;
  if (from->bws) {
    g->bws = (float*)malloc((size_t)from->nbws * sizeof(float));
    memcpy(g->bws, from->bws, (size_t)from->nbws * sizeof(float));
    g->bwfilename = (char*)malloc(NAME_MAX_LEN);
    memcpy(g->bwfilename, from->bwfilename, NAME_MAX_LEN);
    Geom_next(g, 1);
  }
  g->trasl = dup3(from->trasl);
  g->rot = dup3(from->rot);
  return g;
}

int Geom_perturb(kds_rng_fct_t rng, void* rngstate, const Geometry* geom,
                 mcpl_particle_t* part) {  // geom.c:139-158
  int ret = 0;
  if (geom->trasl) traslv(part->position, geom->trasl, 1);
  if (geom->rot) {
    rotv(part->position, geom->rot, 1);
    rotv(part->direction, geom->rot, 1);
  }
  for (int i = 0; i < geom->ord; i++)
    ret += geom->ms[i]->perturb(rng, rngstate, geom->ms[i], part, geom->bw, geom->kernel);
  if (geom->rot) {
    rotv(part->position, geom->rot, 0);
    rotv(part->direction, geom->rot, 0);
  }
  if (geom->trasl) traslv(part->position, geom->trasl, 0);
  return ret;
}

// geom.c:160-177: next bandwidth of the variable-bandwidth list (loaded at creation)
int Geom_next(Geometry* geom, int loop) {
  if (!geom->bws) return 0;
  if (geom->bwpos < geom->nbws) {
    geom->bw = geom->bws[geom->bwpos++];
    return 0;
  }
  if (!loop) kds_end("Geom_next: End of BW file reached.");
  geom->bwpos = 0;
  geom->bw = geom->bws[geom->bwpos++];
  return 1;
}

void Geom_destroy(Geometry* g) {
  if (!g) return;
  for (int i = 0; i < g->ord; i++) Metric_destroy(g->ms[i]);
  free(g->ms);
  free(g->trasl);
  free(g->rot);
  free(g->bwfilename);
  free(g->bws);
  free(g);
}

// ---- KDSource (kdsource.c:70-373) -------------------------------------------
KDSource* KDS_create(double J, char kernel, PList* plist, Geometry* geom) {
  KDSource* k = (KDSource*)calloc(1, sizeof(KDSource));
  k->J = J;
  k->kernel = kernel;
  k->plist = plist;
  k->geom = geom;
  return k;
}

KDSource* KDS_open(const char* xmlfilename) {
  printf("Reading xmlfile %s...\n", xmlfilename);
  XNode* root = parse_xml_file(xmlfilename);
  if (!root) {
    printf("Could not open file %s\n", xmlfilename);
    kds_error("Error in KDS_open");
  }
  if (root->name != "KDSource") {
    printf("Invalid format in source XML file %s\n", xmlfilename);
    kds_error("Error in KDS_open");
  }
  // positional walk, as the reference (children order: J, [kernel], PList, Geom, scaling, BW)
  size_t c = 0;
  auto need = [&](size_t i) -> XNode* {
    if (i >= root->kids.size()) kds_error("Error in KDS_open: truncated source file");
    return root->kids[i];
  };
  double J = 0;
  sscanf(need(c++)->content().c_str(), "%lf", &J);
  char kernel = 'g';
  if (need(c)->name == "kernel") {
    const std::string k = need(c++)->content();
    for (char ch : k)
      if (!isspace((unsigned char)ch)) {
        kernel = ch;
        break;
      }
  } else {
    printf("No kernel specified. Using gaussian as default.\n");
  }
  XNode* pltree = need(c++);
  XNode* gtree = need(c++);
  XNode* sctree = need(c++);
  XNode* bwtree = need(c++);
  if (pltree->kids.size() < 5) kds_error("Error in KDS_open: invalid PList");
  char pt = 'n';
  sscanf(pltree->kids[0]->content().c_str(), "%c", &pt);
  std::string mcplfile = pltree->kids[1]->content();
  if (mcplfile.size() > NAME_MAX_LEN) kds_error("Error in KDS_open");
  auto vec3 = [](XNode* n, double* out) -> bool {
    const std::string s = n->content();
    if (s.size() <= 1) return false;
    sscanf(s.c_str(), "%lf %lf %lf", &out[0], &out[1], &out[2]);
    return true;
  };
  double tp[3], rp[3], tg[3], rg[3];
  const bool has_tp = vec3(pltree->kids[2], tp), has_rp = vec3(pltree->kids[3], rp);
  int x2z = 0;
  sscanf(pltree->kids[4]->content().c_str(), "%d", &x2z);
  int order = 0;
  if (!gtree->attr("order") || sscanf(gtree->attr("order"), "%d", &order) != 1 || order < 1)
    kds_error("Error in KDS_open: invalid value of \"order\"");
  if ((int)gtree->kids.size() < order + 2) kds_error("Error in KDS_open: invalid Geom");
  std::vector<Metric*> metrics(order);
  std::vector<int> dims(order);
  std::vector<std::vector<double>> params(order);
  std::vector<PerturbFun> perturbs(order);
  for (int i = 0; i < order; i++) {
    XNode* m = gtree->kids[i];
    int j;
    for (j = 0; j < _n_metrics; j++)
      if (m->name == _metric_names[j]) break;
    if (j == _n_metrics) {
      printf("Invalid %s metric.\n", m->name.c_str());
      kds_error("Error in KDS_open");
    }
    perturbs[i] = _metric_perturbs[j];
    if (m->kids.size() < 2) kds_error("Error in KDS_open: invalid metric");
    sscanf(m->kids[0]->content().c_str(), "%d", &dims[i]);
    if (dims[i] < 1) kds_error("Error in KDS_open: invalid value of \"dims[i]\"");
    int nps = 0;
    if (m->kids[1]->attr("nps")) sscanf(m->kids[1]->attr("nps"), "%d", &nps);
    params[i].assign(nps > 0 ? nps : 0, 0.0);
    const std::string ptxt = m->kids[1]->content();
    const char* buf = ptxt.c_str();
    for (int k = 0; k < nps; k++) {
      int n = 0;
      if (sscanf(buf, "%lf %n", &params[i][k], &n) < 1) break;  // "None" stays 0
      buf += n;
    }
  }
  const bool has_tg = vec3(gtree->kids[order], tg), has_rg = vec3(gtree->kids[order + 1], rg);
  {
    const std::string stxt = sctree->content();
    const char* buf = stxt.c_str();
    for (int i = 0; i < order; i++) {
      std::vector<double> sc(dims[i], 0.0);
      for (int k = 0; k < dims[i]; k++) {
        int n = 0;
        if (sscanf(buf, "%lf %n", &sc[k], &n) < 1) break;
        buf += n;
      }
      metrics[i] = Metric_create(dims[i], sc.data(), perturbs[i], (int)params[i].size(),
                                 params[i].data());
    }
  }
  int variable_bw = 0;
  if (bwtree->attr("variable")) sscanf(bwtree->attr("variable"), "%d", &variable_bw);
  double bw = 0;
  std::string bwfile;
  if (variable_bw) {
    bwfile = bwtree->content();
    if (bwfile.size() > NAME_MAX_LEN) kds_error("Error in KDS_open");
  } else {
    sscanf(bwtree->content().c_str(), "%lf", &bw);
  }
  PList* plist = PList_create(pt, mcplfile.c_str(), has_tp ? tp : nullptr, has_rp ? rp : nullptr,
                              x2z);
  Geometry* geom = Geom_create(order, metrics.data(), bw, variable_bw ? bwfile.c_str() : nullptr,
                               kernel, has_tg ? tg : nullptr, has_rg ? rg : nullptr);
  KDSource* s = KDS_create(J, kernel, plist, geom);
  delete root;
  printf("Done.\n");
  return s;
}

int KDS_rand_sample2(kds_rng_fct_t rng, void* rngstate, KDSource* kds, mcpl_particle_t* part,
                     int perturb, double w_crit, WeightFun bias, int loop) {
  Impl* I = ensure_impl(kds);
  const bool seq = !(w_crit > 0);
  if (seq) bias = nullptr;  // the list walk ignores the bias (kdsource.c:286-292)
  perturb = perturb ? 1 : 0;
  if (seq && !loop && I->cursor >= (uint64_t)I->N)
    kds_end("PList_next: End of particle list reached.");
  if (bias) ensure_bias(kds, I, bias);
  if (I->used >= I->have || I->batch_perturb != perturb || I->batch_seq != (int)seq ||
      I->batch_bias != bias)
    fill_batch(kds, I, rng ? rng : default_rng, rngstate, perturb, seq, bias);
  const uint64_t b = I->used++;
  const int64_t idx = I->h_idx[b];
  const double* o = &I->h_parts[b * 8];
  const McplList* L = list_of(kds->plist);
  *part = L->parts[I->src_index[idx]];  // polarisation, userflags, pdgcode of the source entry
  part->ekin = o[0];
  for (int k = 0; k < 3; k++) {
    part->position[k] = o[1 + k];
    part->direction[k] = o[4 + k];
  }
  part->time = o[7];
  int ret = 0;
  if (seq) {
    part->weight = I->w[idx];
    I->cursor++;
    if (I->cursor % (uint64_t)I->N == 0 && loop) ret = 1;
  } else {
    // particles are picked with probability ~ w * bias and carry 1 / bias (kdsource.c:325)
    part->weight = bias ? 1.0 / I->bias_vals[idx] : 1.0;
  }
  return ret;
}

int KDS_rand_sample(kds_rng_fct_t rng, void* rngstate, KDSource* kds, mcpl_particle_t* part) {
  return KDS_rand_sample2(rng, rngstate, kds, part, 1, 1, nullptr, 1);
}

int KDS_sample2(KDSource* kds, mcpl_particle_t* part, int perturb, double w_crit, WeightFun bias,
                int loop) {
  return KDS_rand_sample2(default_rng, nullptr, kds, part, perturb, w_crit, bias, loop);
}

int KDS_sample(KDSource* kds, mcpl_particle_t* part) {
  return KDS_rand_sample(default_rng, nullptr, kds, part);
}

// mean of w (times bias) over the next N list entries; advances the list cursor
// (kdsource.c:353-367)
double KDS_rand_w_mean(kds_rng_fct_t, void*, KDSource* kds, int N, WeightFun bias) {
  Impl* I = ensure_impl(kds);
  double s = 0;
  mcpl_particle_t part;
  for (int i = 0; i < N; i++) {
    const int64_t j = (int64_t)((I->cursor + (uint64_t)i) % (uint64_t)I->N);
    if (bias) {
      valid_particle(kds, I, j, &part);
      part.weight = I->w[j];
      s += I->w[j] * bias(&part);
    } else {
      s += I->w[j];
    }
  }
  I->cursor += (uint64_t)(N > 0 ? N : 0);
  I->have = I->used = 0;  // a pending sequential batch no longer matches the cursor
  return N > 0 ? s / N : 0.0;
}

double KDS_w_mean(KDSource* kds, int N, WeightFun bias) {
  return KDS_rand_w_mean(default_rng, nullptr, kds, N, bias);
}

void KDS_destroy(KDSource* kds) {
  if (!kds) return;
  free_impl((Impl*)kds->impl);
  PList_destroy(kds->plist);
  Geom_destroy(kds->geom);
  free(kds);
}

// ---- MultiSource (kdsource.c:375-473) ----------------------------------------
MultiSource* MS_create(int len, KDSource** s, const double* ws) {
  if (len < 1) kds_error("Error in MS_create: invalid value of \"len\"");
  MultiSource* ms = (MultiSource*)calloc(1, sizeof(MultiSource));
  ms->len = len;
  ms->s = (KDSource**)calloc(len, sizeof(KDSource*));
  ms->ws = (double*)calloc(len, sizeof(double));
  ms->cdf = (double*)calloc(len, sizeof(double));
  double run = 0;
  for (int i = 0; i < len; i++) {
    ms->s[i] = s[i];
    ms->ws[i] = ws[i];
    ms->J += s[i]->J;
    run += ws[i];
    ms->cdf[i] = run;
  }
  return ms;
}

MultiSource* MS_open(int len, const char** xmlfilenames, const double* ws) {
  if (len < 1) kds_error("Error in MS_open: invalid value of \"len\"");
  std::vector<KDSource*> s(len);
  for (int i = 0; i < len; i++) s[i] = KDS_open(xmlfilenames[i]);
  return MS_create(len, s.data(), ws);
}

int MS_rand_sample2(kds_rng_fct_t rng, void* rngstate, MultiSource* ms, mcpl_particle_t* part,
                    int perturb, double w_crit, WeightFun bias, int loop) {
  if (!rng) rng = default_rng;
  const double y = rng(rngstate);
  const double tot = ms->cdf[ms->len - 1];
  int i = 0;
  if (tot <= 0) {
    i = (int)(y * ms->len);
  } else {
    while (i < ms->len - 1 && y * tot > ms->cdf[i]) i++;
  }
  const int ret = KDS_rand_sample2(rng, rngstate, ms->s[i], part, perturb, w_crit, bias, loop);
  if (tot > 0)
    part->weight *= (ms->s[i]->J / ms->J) / (ms->ws[i] / tot);
  else
    part->weight *= (ms->s[i]->J / ms->J) * ms->len;
  return ret;
}

int MS_rand_sample(kds_rng_fct_t rng, void* st, MultiSource* ms, mcpl_particle_t* part) {
  return MS_rand_sample2(rng, st, ms, part, 1, 1, nullptr, 1);
}
int MS_sample2(MultiSource* ms, mcpl_particle_t* part, int perturb, double w_crit, WeightFun bias,
               int loop) {
  return MS_rand_sample2(default_rng, nullptr, ms, part, perturb, w_crit, bias, loop);
}
int MS_sample(MultiSource* ms, mcpl_particle_t* part) {
  return MS_rand_sample(default_rng, nullptr, ms, part);
}
double MS_rand_w_mean(kds_rng_fct_t rng, void* st, MultiSource* ms, int N, WeightFun bias) {
  double wm = 0;
  for (int i = 0; i < ms->len; i++) wm += ms->ws[i] * KDS_rand_w_mean(rng, st, ms->s[i], N, bias);
  return wm / ms->cdf[ms->len - 1];
}
double MS_w_mean(MultiSource* ms, int N, WeightFun bias) {
  return MS_rand_w_mean(default_rng, nullptr, ms, N, bias);
}
void MS_destroy(MultiSource* ms) {
  if (!ms) return;
  for (int i = 0; i < ms->len; i++) KDS_destroy(ms->s[i]);
  free(ms->s);
  free(ms->ws);
  free(ms->cdf);
  free(ms);
}

// ---- ctypes entry points -------------------------------------------------------
void kdsource_resample_to_mcpl(double (*rng_stateless)(void), const char* kds_sourcefile,
                               const char* destination_mcpl, uint64_t nout) {
  KDSource* kds = KDS_open(kds_sourcefile);
  Impl* I = ensure_impl(kds);
  const std::string out = mcpl_gz_name(destination_mcpl);
  (void)KDS_w_mean(kds, 1000, nullptr);  // resamplemcpl.c:35 (its value, w_crit, is not needed)
  // the caller's generator keys the counter-based stream with one draw
  const uint64_t seed = rng_stateless ? (uint64_t)(rng_stateless() * 9007199254740992.0) : I->seed;
  printf("Resampling...\n");
  // records are produced, gzip-compressed (gzip.cu) and staged on the device; a worker
  // thread writes the previous batch while the next one is generated
  const uint64_t Bmax = nout > (1ull << 27) ? (1ull << 22) : (1ull << 20);  // see resample.py
  const uint64_t B = nout < Bmax ? nout : Bmax;
  kdsb200_gz_writer* w = kdsb200_gz_writer_open(out.c_str(), B * 36, 36, 12, (void*)I->stream);
  if (!w) kds_error(kdsb200_last_error());
  const std::vector<unsigned char> hdr =
      mcpl_header_bytes(nout, "KDSource resample", true, false, 0, false, 0.0, false);
  cuda_or_die(kdsb200_gz_writer_put_host(w, hdr.data(), hdr.size()));
  unsigned char* d_rec;
  cu(cudaMalloc(&d_rec, B * 36));
  // the bulk path picks through the pick table (one HBM line per particle, kdsb200.h) where
  // the fp32 Philox kernel serves the geometry and the table fits; else the window table
  const void* d_table = I->d_guide;
  int table_bits = I->guide_bits, table_kind = 0;
  void* d_pick = nullptr;
  if (I->src_f32 && I->contract == 2 && nout >= (uint64_t)I->N) {
    int bits = 4;
    while (bits < 30 && (1ll << bits) < 2 * I->N) bits++;
    size_t free_b = 0, total_b = 0;
    cu(cudaMemGetInfo(&free_b, &total_b));
    while (bits > 4 && kdsb200_pick_table_bytes(bits) > free_b / 3) bits--;
    if (cudaMalloc(&d_pick, kdsb200_pick_table_bytes(bits)) == cudaSuccess) {
      cuda_or_die(kdsb200_pick_table(I->d_cdf, I->N, bits, (const float*)I->d_src, I->d_bw,
                                     kds->geom->bw, d_pick, (void*)I->stream));
      d_table = d_pick;
      table_bits = bits;
      table_kind = 1;
    } else {
      (void)cudaGetLastError();
      d_pick = nullptr;
    }
  }
  for (uint64_t first = 0; first < nout; first += B) {
    const uint64_t n = nout - first < B ? nout - first : B;
    cuda_or_die(kdsb200_resample(I->d_src, I->src_f32, I->d_cdf, d_table, table_bits, table_kind,
                                 I->d_bw, kds->geom->bw, I->N, &I->geom, kds->plist->pdgcode, seed,
                                 first, n, -1, nullptr, 0, 1, I->contract, d_rec, nullptr, nullptr,
                                 (void*)I->stream));
    cuda_or_die(kdsb200_gz_writer_put_device(w, d_rec, n * 36));
  }
  cuda_or_die(kdsb200_gz_writer_close(w, nullptr));
  cudaFree(d_rec);
  cudaFree(d_pick);
  KDS_destroy(kds);
  printf("Successfully sampled %llu particles.\n", (unsigned long long)nout);
}

void kdsource_write_mcpl(const char* filename, uint64_t n, unsigned flags, const double* ekin,
                         const double* x, const double* y, const double* z, const double* ux,
                         const double* uy, const double* uz, const double* time,
                         const double* weight, const int32_t* pdgcode, const double* polx,
                         const double* poly, const double* polz, const uint32_t* userflags) {
  const bool dbl = flags & 1, upd = flags & 2, uw = flags & 4;  // writemcpl.c:4-13
  const bool pol = polx || poly || polz;
  const std::string out = mcpl_gz_name(filename);
  gzFile f = gzopen(out.c_str(), "wb6");
  if (!f) kds_error("could not create output MCPL file");
  gzbuffer(f, 1 << 20);
  mcpl_write_header(f, n, "KDSource MCPL writer", !dbl, pol, upd ? pdgcode[0] : 0, uw,
                    uw ? weight[0] : 0.0, userflags != nullptr);
  std::vector<unsigned char> rec;
  rec.reserve(128);
  auto put = [&](double v) {
    if (dbl) {
      const unsigned char* b = (const unsigned char*)&v;
      rec.insert(rec.end(), b, b + 8);
    } else {
      const float fv = (float)v;
      const unsigned char* b = (const unsigned char*)&fv;
      rec.insert(rec.end(), b, b + 4);
    }
  };
  for (uint64_t i = 0; i < n; i++) {
    rec.clear();
    if (pol) {
      put(polx ? polx[i] : 0.0);
      put(poly ? poly[i] : 0.0);
      put(polz ? polz[i] : 0.0);
    }
    put(x[i]);
    put(y[i]);
    put(z[i]);
    const double d[3] = {ux[i], uy[i], uz[i]};
    double p0, p1, sek;
    pack_dir(d, ekin[i], &p0, &p1, &sek);
    put(p0);
    put(p1);
    put(sek);
    put(time[i]);
    if (!uw) put(weight[i]);
    if (!upd) {
      const unsigned char* b = (const unsigned char*)&pdgcode[i];
      rec.insert(rec.end(), b, b + 4);
    }
    if (userflags) {
      const unsigned char* b = (const unsigned char*)&userflags[i];
      rec.insert(rec.end(), b, b + 4);
    }
    gz_write_all(f, rec.data(), rec.size());
  }
  gzclose(f);
}

}  // extern "C"
