// oracle/dana_oracle.cpp — CPU oracle for the din-mol-Li hot path.
//
// TEST INFRASTRUCTURE ONLY (see dana_oracle.h).  A from-scratch C++ restatement of what the
// reference program `dana` computes per time step, written so that iteration orders, rounding
// and RNG consumption are those of the default (-O0, no FMA) gfortran build that produced
// tests/*/ref.xyz.  Every routine cites the reference lines it follows (paths relative to the
// reference tree).  Build: g++ -O2 -ffp-contract=off (no -ffast-math), glibc logf/exp/sqrt.
//
// Parity status: PINNED by tests/golden/{ermak,brown,gcmc}/ref.xyz (bit-for-bit).
//
// Data structures deliberately mirror the reference (heap atom objects with separately
// allocated pos/vel/force/acel, circular doubly linked membership lists, pointer index arrays,
// head/next linked cells, dense column-major neighbour table list(n,mnb)) because this code is
// also the timed CPU baseline.

#include "dana_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Halt : std::runtime_error { using std::runtime_error::runtime_error; };

// ---------------------------------------------------------------------------------------------
// RNG — dana.F90:1407-1428 (ran) and dana.F90:1379-1404 (gasdev)
// ---------------------------------------------------------------------------------------------
struct Rng {
  int32_t idum = 0, ix = -1, iy = -1;
  bool stored = false;
  double g = 0.0;
  uint64_t calls = 0;
  // am = nearest(1.0,-1.0)/real(im,dp): the float just below 1, divided in double
  static double am() { return (double)std::nextafterf(1.0f, -1.0f) / 2147483647.0; }

  double ran() {
    ++calls;
    if (idum <= 0 || iy < 0) {
      int32_t a = idum < 0 ? -idum : idum;
      iy = (888889999 ^ a) | 1;
      ix = 777755555 ^ a;
      idum = a + 1;
    }
    uint32_t u = (uint32_t)ix;
    u ^= u << 13; u ^= u >> 17; u ^= u << 5;       // ishft is a logical shift
    ix = (int32_t)u;
    int32_t k = iy / 127773;
    iy = 16807 * (iy - k * 127773) - 2836 * k;
    if (iy < 0) iy += 2147483647;
    int32_t m = (2147483647 & (ix ^ iy)) | 1;
    return am() * (double)m;
  }
  double gasdev() {
    if (stored) { stored = false; return g; }
    double v1, v2; float rsq;
    for (;;) {
      v1 = 2.0 * ran() - 1.0;
      v2 = 2.0 * ran() - 1.0;
      rsq = (float)(v1 * v1 + v2 * v2);            // real(sp) :: rsq
      if ((double)rsq > 0.0 && (double)rsq < 1.0) break;
    }
    // rsq=sqrt(-2.0_dp*log(rsq)/rsq): log of a real(sp) is logf; the result is stored in sp
    float fac = (float)std::sqrt(-2.0 * (double)logf(rsq) / (double)rsq);
    g = v2 * (double)fac; stored = true;
    return v1 * (double)fac;
  }
};

// libgcc __powidf2 order: what gfortran -O0 evaluates for x**6 / x**7
inline double pow6(double x) { double x2 = x * x; double x4 = x2 * x2; return x2 * x4; }
inline double pow7(double x) { double x2 = x * x; double y = x * x2; double x4 = x2 * x2; return y * x4; }
inline double idnint(double x) { return std::round(x); }   // half away from zero

// ---------------------------------------------------------------------------------------------
// Object model — Groups.F90
// ---------------------------------------------------------------------------------------------
struct Group;
struct Dana;

struct Atom {                       // Groups.F90:251-328, allocation 370-387
  double *pos, *vel, *force, *acel; // four separate heap blocks, like the reference
  double pos_old[3] = {1e8, 1e8, 1e8}, old_cg[3] = {1e8, 1e8, 1e8};
  double epot = 0.0, mass = 1.0;
  int z = 119;
  bool pbc[3] = {false, false, false};
  bool skip = false;
  int ngr = 0;
  std::vector<int> gr, id;          // sorted group ids + per-group index
  int64_t uid = -1;                 // creation counter (oracle bookkeeping, not in the reference)
  Atom() : gr(10, 0), id(10, 0) {
    pos = new double[3](); vel = new double[3](); force = new double[3](); acel = new double[3]();
  }
  ~Atom() { delete[] pos; delete[] vel; delete[] force; delete[] acel; }
  void setz(int zz) { z = zz; mass = (zz >= 1 && zz <= 3) ? 6.94 : 1.0; }  // Elements: dana.F90:82-84
  // binleft — Groups.F90:1555-1598 (1-based m; found flag)
  bool binleft(int val, int &m) const {
    int l = 0, r = ngr + 1; m = 0;
    while (r > l + 1) {
      m = (r + l) / 2; int j = gr[m - 1];
      if (val == j) return true;
      if (val > j) l = m; else r = m;
    }
    if (r <= ngr && gr[r - 1] == val) { m = r; return true; }
    m = l; return false;
  }
  int gri(int gid_) const { int m; return binleft(gid_, m) ? m : 0; }          // Groups.F90:565-577
  int gid(int gid_) const { int i = gri(gid_); return i == 0 ? -1 : id[i - 1]; } // Groups.F90:595-610
};

struct Node { Node *next, *prev; Atom *o; };

struct Group {                      // Groups.F90:34-175
  int id = 0, nat = 0;
  Node *alist = nullptr;
  Dana *D = nullptr;
  virtual ~Group() {
    if (alist) { Node *n = alist->next; while (n != alist) { Node *nx = n->next; delete n; n = nx; } delete alist; }
  }
  void group_construct(Dana *d);
  // group_attach_atom — Groups.F90:730-766; returns l (1-based position in a->gr) or 0 if present
  int group_attach(Atom *a) {
    int l;
    if (a->binleft(id, l)) return 0;
    if (a->ngr == (int)a->gr.size()) { a->gr.resize(a->ngr + 5, 0); a->id.resize(a->ngr + 5, 0); }
    l += 1;                                           // sorted insertion, Groups.F90:536-547
    for (int i = a->ngr; i >= l; --i) { a->gr[i] = a->gr[i - 1]; a->id[i] = a->id[i - 1]; }
    a->gr[l - 1] = id; a->id[l - 1] = 0; a->ngr++;
    Node *nw = new Node;                              // add_before(head) == append at tail, cdlist_body.inc:43-58
    nw->next = alist; nw->prev = alist->prev; alist->prev->next = nw; alist->prev = nw; nw->o = a;
    nat++;
    return l;
  }
  // group_detach_atom + group_detach_link — Groups.F90:783-844
  void group_detach(Atom *a, Node **la_) {
    Node *la = alist;
    for (int i = 1; i <= nat; ++i) {
      la = la->next;
      if (la->o != a) continue;
      int j;
      if (!a->binleft(id, j)) return;                 // delgr, Groups.F90:549-563
      a->ngr--;
      for (int k = j; k <= a->ngr; ++k) { a->gr[k - 1] = a->gr[k]; a->id[k - 1] = a->id[k]; }
      Node *prev = la->prev;
      la->prev->next = la->next; la->next->prev = la->prev; delete la;
      nat--;
      if (la_) *la_ = prev;
      return;
    }
  }
  virtual int attach(Atom *a) { return group_attach(a); }
  virtual void detach(Atom *a, Node **la_ = nullptr) { group_detach(a, la_); }
};

struct IGroup : Group {             // Groups.F90:179-219
  std::vector<Atom *> a;            // 1-based: a[0] unused
  int amax = 0, nlimbo = 0, pad = 100;
  bool b_limbo = false;
  Atom *limbo;
  IGroup() { limbo = new Atom; }
  ~IGroup() override { delete limbo; }
  int asize() const { return (int)a.size() - 1; }
  void igroup_construct(Dana *d) { group_construct(d); a.assign(pad + 1, nullptr); }
  void clean() {                    // igroup_clean — Groups.F90:1036-1053
    if (!b_limbo) return;
    for (int i = 1; i <= amax; ++i) if (a[i] == limbo) a[i] = nullptr;
    nlimbo = 0; b_limbo = false;
  }
  int igroup_attach(Atom *at) {     // Groups.F90:1058-1100
    int l = group_attach(at);
    if (l == 0) return 0;
    int n = asize(), m = nat + nlimbo;
    if (n < m) a.resize(m + pad + 1, nullptr);
    if (amax >= nat + nlimbo) {
      for (n = 1; n <= amax; ++n) if (a[n] == nullptr) break;
      if (n > amax) throw Halt("Index inconsistency");
    } else { amax++; n = amax; }
    a[n] = at; at->id[l - 1] = n;
    return l;
  }
  void igroup_detach(Atom *at, Node **la_) {   // Groups.F90:1105-1135
    int i = at->gid(id);
    if (i == -1) return;
    a[i] = nullptr;
    group_detach(at, la_);
  }
  int attach(Atom *at) override { return igroup_attach(at); }
  void detach(Atom *at, Node **la_ = nullptr) override { igroup_detach(at, la_); }
};

// 27-cell stencil, Cells.F90:28-36 (order matters: it fixes the order of every neighbour row)
static const int MAP[27][3] = {
  {0,0,0},{1,0,0},{1,1,0},{0,1,0},{-1,1,0},{1,0,-1},{1,1,-1},{0,1,-1},{-1,1,-1},
  {1,0,1},{1,1,1},{0,1,1},{-1,1,1},{0,0,1},{-1,0,0},{-1,-1,0},{0,-1,0},{1,-1,0},
  {-1,0,1},{-1,-1,1},{0,-1,1},{1,-1,1},{-1,0,-1},{-1,-1,-1},{0,-1,-1},{1,-1,-1},{0,0,-1}};

struct CGroup : IGroup {            // Cells.F90:39-84
  std::vector<int> head, next;      // head(0:nx+1,0:ny+1,0:nz+1), next(:) 1-based
  int ncells[3] = {1, 1, 1};
  int hd[3] = {0, 0, 0};            // allocated extents of head (n+2)
  double cell[3] = {0, 0, 0};
  double rcut = 1e10;
  bool tessellated = false;
  void cgroup_construct(Dana *d) { igroup_construct(d); next.assign(pad + 1, 0); }
  int &H(int i, int j, int k) { return head[(size_t)i + (size_t)hd[0] * ((size_t)j + (size_t)hd[1] * (size_t)k)]; }
  void tessellate();
  void sort() {                     // cgroup_sort — Cells.F90:267-279
    std::fill(head.begin(), head.end(), 0);
    for (int i = 1; i <= amax; ++i) { if (a[i] == nullptr) continue; sort_atom(i, false); }
  }
  // cgroup_sort_atom — Cells.F90:281-302; returns errf (only meaningful when silent)
  bool sort_atom(int i, bool silent) {
    Atom *at = a[i]; int ci[3];
    for (int k = 0; k < 3; ++k) ci[k] = (int)(at->pos[k] / cell[k]) + 1;   // division, truncation
    bool neg = ci[0] < 0 || ci[1] < 0 || ci[2] < 0;
    bool big = ci[0] > ncells[0] + 1 || ci[1] > ncells[1] + 1 || ci[2] > ncells[2] + 1;
    if (!silent && (neg || big)) throw Halt("Particle out of tessellation");
    if (big) return true;           // second werr leaves errf set
    if (neg) throw Halt("oracle: reference would index head() out of bounds (silent ci<0)");
    next[i] = H(ci[0], ci[1], ci[2]); H(ci[0], ci[1], ci[2]) = i;
    return false;
  }
  void unsort_atom(int i) {         // cgroup_unsort_atom — Cells.F90:304-352
    Atom *at = a[i]; int ci[3];
    for (int k = 0; k < 3; ++k) ci[k] = (int)(at->pos[k] / cell[k]) + 1;
    for (int k = 0; k < 3; ++k) if (ci[k] < 0) { tessellated = false; return; }
    for (int k = 0; k < 3; ++k) if (ci[k] > ncells[k]) { tessellated = false; return; }
    bool removed = false;
    int &h = H(ci[0], ci[1], ci[2]);
    if (h == i) { h = next[i]; removed = true; }
    else {
      int j = h;
      while (j > 0) { int k = next[j]; if (k == i) { next[j] = next[i]; removed = true; break; } j = k; }
    }
    if (!removed) throw Halt("Unsort cgroup particle fail");
  }
  int attach(Atom *at) override {   // cgroup_attach_atom — Cells.F90:105-144
    int l = igroup_attach(at);
    if (l == 0) return 0;
    if ((int)next.size() < (int)a.size()) next.resize(a.size(), 0);
    if (tessellated) { int i = at->gid(id); if (sort_atom(i, true)) tessellated = false; }
    return l;
  }
  void detach(Atom *at, Node **la_ = nullptr) override {   // cgroup_detach_atom — Cells.F90:146-175
    if (tessellated) { int i = at->gid(id); if (i > 0) unsort_atom(i); }
    igroup_detach(at, la_);
  }
};

struct NGroup : IGroup {            // Neighbor.F90:32-79
  Group ref; CGroup b;
  double rcut = 1e10, rcut2 = 1e10;
  int mnb = 10000;
  std::vector<int> nn;              // 1-based
  int32_t *list = nullptr; size_t ld = 0;   // list(ld,mnb) column-major, 1-based (i,m) -> list[(m-1)*ld+(i-1)]
  bool listed = false, use_cells = true;
  ~NGroup() override { free(list); }
  int32_t &L(int i, int m) { return list[(size_t)(m - 1) * ld + (size_t)(i - 1)]; }
  void alloc_tables(int n, bool preserve) {
    int maxcol = 0;
    if (preserve) for (size_t i = 1; i < nn.size(); ++i) maxcol = std::max(maxcol, nn[i]);
    int32_t *nl = (int32_t *)calloc((size_t)n * (size_t)mnb, sizeof(int32_t));
    if (!nl) throw Halt("oracle: cannot allocate list(n,mnb); lower mnb");
    if (preserve && list) {         // t_list(:size,:)=list(:,:) — only the used columns are copied (same semantics)
      size_t oldn = nn.size() - 1;
      for (int m = 1; m <= maxcol; ++m)
        memcpy(nl + (size_t)(m - 1) * n, list + (size_t)(m - 1) * ld, oldn * sizeof(int32_t));
    }
    free(list); list = nl; ld = n;
    if (preserve) nn.resize(n + 1, 0); else nn.assign(n + 1, 0);
  }
  void ngroup_construct(Dana *d);
  void setrc(double rc);
  int attach(Atom *at) override;
  void detach(Atom *at, Node **la_ = nullptr) override;
  void sort_atom(int i);
  void cells_atom(int i);
  void verlet_atom(int i);
  void build_cells();
  void build_verlet();
};

struct TraceRec { int32_t kind; int64_t uid; double val; };

struct Frame { int nat = 0; double zmax = 0; std::vector<int32_t> z; std::vector<double> pos; double scal[6] = {0}; };

// ---------------------------------------------------------------------------------------------
// The program state — dana.F90:1-73 plus module globals
// ---------------------------------------------------------------------------------------------
struct Dana {
  orc_params P;
  std::string err;
  Rng rng;
  // gems_program_types (Program_Types.F90:43-54,81-95)
  double tbox33 = 0, box[3] = {1e6, 1e6, 1e6}, one_box[3] = {1e-6, 1e-6, 1e-6};
  const bool mic = true;
  // gems_neighbor globals (Neighbor.F90:109-112)
  double nb_dcut = 1.0, maxrcut = 0.0; int64_t nupd_vlist = 0;
  std::vector<Group *> gindex;      // 1-based ids
  IGroup sys; NGroup hs; Group gcmc, chunk;
  std::vector<Atom *> chunk_atoms;
  // dana scalars
  int n = 0, nx = 0, nchunk = 0;
  double z0 = 0, z1 = 0, zmax = 0, xi = 0, yi = 0, dif = 0, dif_sei = 0, dif_sc = 0;
  double eps[4][4], r0[4][4];       // 1-based (k,m)
  double prob = 0, max_vel = 0, msd_u = 0, msd_t = 0, msd_max = 0;
  int64_t choques = 0, choques2 = 0, choques3 = 0, try_ = 0, depo = 0;
  double h = 0, t = 0; int nst = 0, nwr = 0;
  bool integrador = false, s_piston = false, s_chunk = false, s_gcmc = false;
  double rho = 0, rho0 = 0, dist = 0, rhomedia = 0, cstdev = 0;
  double cc0 = 0, cc1 = 0, cc2 = 0, sdr = 0, sdv = 0, crv1 = 0, crv2 = 0, skt = 0;
  std::vector<double> ranv; int ranv_n = 0;
  double act = 0; int nadj = 0;
  int step = 0; int64_t next_uid = 0;
  bool tracing = false; std::vector<TraceRec> trace;
  Frame frame;
  // constants dana.F90:12-25
  static constexpr double tau = 0.1, gama = 1.0, Tsist = 300.0;
  static double kB_ui() { const double kB_eVK = 8.617330350e-5, eV_ui = 96.485 * 100.0; return kB_eVK * eV_ui; }
  // Constants.F90:165-167 kB_ui used only inside gcmc_run (dana.F90:594)
  static double kB_ui_module() {
    const double axps_mxs = 1.0e2, uma_kg = 1.6605402e-27, qe_si = 1.60219e-19;   // Constants.F90:100-120
    const double joule_ev = 1.0 / qe_si;
    const double ui_ev = axps_mxs * axps_mxs * uma_kg * joule_ev;
    const double ev_ui = 1.0 / ui_ev;
    return 8.617385e-05 * ev_ui;
  }

  double ran_t(int kind, int64_t uid) { double v = rng.ran(); if (tracing) trace.push_back({kind, uid, v}); return v; }
  double gas_t(int kind, int64_t uid) { double v = rng.gasdev(); if (tracing) trace.push_back({kind, uid, v}); return v; }

  Atom *new_atom() { Atom *a = new Atom; a->uid = next_uid++; return a; }

  void box_setvars() { box[2] = tbox33; box[0] = xi; box[1] = yi; for (int k = 0; k < 3; ++k) one_box[k] = 1.0 / box[k]; }

  // vdistance — Groups.F90:995-1016 (i minus j, idnint minimum image on axes where either atom is periodic)
  void vdistance(double vd[3], const Atom *i, const Atom *j) const {
    for (int l = 0; l < 3; ++l) vd[l] = i->pos[l] - j->pos[l];
    for (int l = 0; l < 3; ++l)
      if (i->pbc[l] || j->pbc[l]) vd[l] = vd[l] - box[l] * idnint(vd[l] * one_box[l]);
  }
  // distance — Program_Types.F90:139-153 (r2 minus r1)
  void distance(double vd[3], const double r1[3], const double r2[3], const bool pbc[3]) const {
    for (int l = 0; l < 3; ++l) vd[l] = r2[l] - r1[l];
    for (int l = 0; l < 3; ++l) if (pbc[l]) vd[l] = vd[l] - box[l] * idnint(vd[l] * one_box[l]);
  }
  static double dot3(const double v[3]) { return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]; }

  void init(const orc_params &p);
  void pos_inic(std::vector<double> &r);
  void config_inic(const std::vector<double> &r, const std::vector<int> &zz);
  void set_ermak();
  void do_pbc();
  void update();
  void test_update();
  void fuerza();
  void ermak_a();
  void ermak_b();
  void cbrownian_hs();
  void atom_pbc(Atom *o1, bool &depos);
  void overlap_moveback();
  void promote();
  void gcmc_run();
  void calc_rho();
  void bloques();
  void maxz();
  void salida();
  void msd_book() { msd_t = msd_t / hs.ref.nat; msd_max = std::max(msd_max, msd_t); }   // dana.F90:201-202
  void step_once();
  void destroy_atom(Atom *a);
  ~Dana();
};

void Group::group_construct(Dana *d) {      // Groups.F90:644-662
  D = d; alist = new Node; alist->next = alist->prev = alist; alist->o = nullptr; nat = 0;
  d->gindex.push_back(this); id = (int)d->gindex.size();
}

// cgroup_tessellate — Cells.F90:180-265 (rcut here is b%rcut = rcut + nb_dcut, Neighbor.F90:332)
void CGroup::tessellate() {
  const double *box = D->box;
  if (tessellated) {
    bool ok1 = true, ok2 = true;
    for (int k = 0; k < 3; ++k) if (!((double)(ncells[k] + 1) >= box[k] / rcut)) ok1 = false;
    if (ok1) {
      for (int k = 0; k < 3; ++k) if (!(rcut < box[k] / (double)ncells[k])) ok2 = false;
      if (ok2) { for (int k = 0; k < 3; ++k) cell[k] = box[k] / (double)ncells[k]; return; }
    }
  }
  if (rcut == 1e10) return;
  for (int k = 0; k < 3; ++k) ncells[k] = (int)(box[k] / rcut);
  if (ncells[0] < 4 && ncells[1] < 4 && ncells[2] < 4) return;
  for (int k = 0; k < 3; ++k) cell[k] = box[k] / (double)ncells[k];
  for (int k = 0; k < 3; ++k) hd[k] = ncells[k] + 2;
  head.assign((size_t)hd[0] * hd[1] * hd[2], 0);
  tessellated = true;
}

void NGroup::ngroup_construct(Dana *d) {    // Neighbor.F90:128-150
  ref.group_construct(d);
  b.cgroup_construct(d);
  igroup_construct(d);
  alloc_tables(pad, false);
}
void NGroup::setrc(double rc) {             // Neighbor.F90:323-334
  rcut = rc; rcut2 = rc * rc;
  D->maxrcut = std::max(D->maxrcut, rc + D->nb_dcut);
  b.rcut = rc + D->nb_dcut;
}
int NGroup::attach(Atom *at) {              // ngroup_attach_atom — Neighbor.F90:173-224
  int l = igroup_attach(at);
  if (l == 0) return 0;
  int n = asize();
  if (listed) {
    if ((int)nn.size() - 1 < n) alloc_tables(n, true);
    int i = at->id[l - 1];
    sort_atom(i);
  } else {
    if ((int)nn.size() - 1 < n) alloc_tables(n, false);
  }
  return l;
}
void NGroup::detach(Atom *at, Node **la_) { // ngroup_detach_atom — Neighbor.F90:226-269
  int i = 0;
  if (listed) { i = at->gid(id); nn[i] = 0; }
  ref.detach(at);
  b.detach(at);
  igroup_detach(at, la_);
  if (listed) { a[i] = limbo; nlimbo++; b_limbo = true; }
}
void NGroup::sort_atom(int i) {             // ngroup_sort_atom — Neighbor.F90:271-318
  nn[i] = 0;
  Atom *at = a[i];
  if (at->gid(ref.id) == 0) { if (use_cells) cells_atom(i); else verlet_atom(i); }
  if (at->gid(b.id) > 0) {
    double rc = rcut + D->nb_dcut; rc = rc * rc;
    Node *la = ref.alist;
    for (int jj = 1; jj <= ref.nat; ++jj) {
      la = la->next; Atom *aj = la->o;
      int j = aj->gid(id);
      if (aj == at) continue;
      double vd[3]; D->vdistance(vd, at, aj);
      double rd = Dana::dot3(vd);
      if (rd > rc) continue;                 // inclusive criterion
      int m = nn[j] + 1;
      if (m > mnb) throw Halt("oracle: neighbour row overflow (mnb)");
      L(j, m) = i; nn[j] = m;
    }
  }
}
void NGroup::cells_atom(int i) {            // ngroup_cells_atom — Neighbor.F90:550-603
  nn[i] = 0;
  Atom *ai = a[i];
  double rc2 = rcut + D->nb_dcut; rc2 = rc2 * rc2;
  int rc[3], nc[3];
  for (int k = 0; k < 3; ++k) rc[k] = (int)(ai->pos[k] / b.cell[k]) + 1;
  for (int nab = 0; nab < 27; ++nab) {
    for (int k = 0; k < 3; ++k) {            // cell_pbc with mic=.true. — Cells.F90:378-404 (wraps z too)
      int r = MAP[nab][k] + rc[k] - 1;
      r = (r + b.ncells[k]) % b.ncells[k];   // Fortran mod: sign of dividend; dividend >= -1+n >= 0 here unless r<-n
      nc[k] = r + 1;
    }
    int j = b.H(nc[0], nc[1], nc[2]);
    while (j > 0) {
      Atom *aj = b.a[j];
      int k = aj->gid(id);
      j = b.next[j];
      if (aj == ai) continue;
      double vd[3]; D->vdistance(vd, aj, ai);
      double rd = Dana::dot3(vd);
      if (rd < rc2) {                        // strict criterion
        nn[i]++;
        if (nn[i] > mnb) throw Halt("oracle: neighbour row overflow (mnb)");
        L(i, nn[i]) = k;
      }
    }
  }
}
void NGroup::build_cells() {                // ngroup_cells — Neighbor.F90:465-548
  std::fill(nn.begin(), nn.end(), 0);
  clean();
  // The reference spawns one OpenMP task per ref atom (Neighbor.F90:492-544); each task owns row i, so a
  // parallel-for over the same atoms is the same computation.  Threads: OMP_NUM_THREADS (1 if built without OpenMP).
  std::vector<int> idx; idx.reserve(ref.nat);
  Node *la = ref.alist;
  for (int ii = 1; ii <= ref.nat; ++ii) { la = la->next; idx.push_back(la->o->gid(id)); }
  int failed = 0;
#pragma omp parallel for schedule(dynamic, 64)
  for (int q = 0; q < (int)idx.size(); ++q) {
    try { cells_atom(idx[q]); } catch (...) { failed = 1; }
  }
  if (failed) throw Halt("oracle: neighbour row overflow (mnb)");
  listed = true;
}
void NGroup::verlet_atom(int i) {           // ngroup_verlet_atom — Neighbor.F90:426-463
  nn[i] = 0;
  Atom *ai = a[i];
  double rc2 = rcut + D->nb_dcut; rc2 = rc2 * rc2;
  int m = 0;
  for (int j = 1; j <= b.amax; ++j) {
    Atom *aj = b.a[j];
    if (!aj) continue;
    if (aj == ai) continue;
    double vd[3]; D->vdistance(vd, ai, aj);
    double rd = Dana::dot3(vd);
    if (rd > rc2) continue;
    m++;
    if (m > mnb) throw Halt("oracle: neighbour row overflow (mnb)");
    L(i, m) = aj->gid(id);
  }
  nn[i] = m;
}
void NGroup::build_verlet() {               // ngroup_verlet — Neighbor.F90:358-424
  std::fill(nn.begin(), nn.end(), 0);
  clean();
  Node *la = ref.alist;
  for (int ii = 1; ii <= ref.nat; ++ii) { la = la->next; verlet_atom(la->o->gid(id)); }
  listed = true;
}

// ---------------------------------------------------------------------------------------------
// Set-up — dana.F90:75-167
// ---------------------------------------------------------------------------------------------
static double quant12(double x) {            // write f25.12 / read back (dana.F90:391,463)
  char buf[64]; snprintf(buf, sizeof buf, "%.12f", x); return strtod(buf, nullptr);
}

struct InitGrid {                            // cell-accelerated variant of the O(N^2) scan in pos_inic
  double cx, cy, cz; int nx, ny, nz; std::vector<std::vector<int>> c;
  InitGrid(double xi, double yi, double alto, double r0) {
    nx = std::max(1, (int)(xi / r0)); ny = std::max(1, (int)(yi / r0)); nz = std::max(1, (int)(alto / r0));
    cx = xi / nx; cy = yi / ny; cz = alto / nz; c.resize((size_t)nx * ny * nz);
  }
  int ix(double x) const { int i = (int)(x / cx); return i < 0 ? 0 : (i >= nx ? nx - 1 : i); }
  int iy(double y) const { int i = (int)(y / cy); return i < 0 ? 0 : (i >= ny ? ny - 1 : i); }
  int iz(double z) const { int i = (int)(z / cz); return i < 0 ? 0 : (i >= nz ? nz - 1 : i); }
};

// pos_inic — dana.F90:330-396.  fast=true replaces the scan over all previous particles with a
// scan over the 27 surrounding grid cells; the accept/reject decision ("any earlier particle
// closer than r0") and hence the RNG stream and the result are identical.
static int gen_pos_inic(Rng &rng, double xi, double yi, double alto, bool fast, std::vector<double> &r,
                        const double box[3], const double one_box[3]) {
  const double r0 = 3.2, o = 0.0;
  // n = Mol*xi*yi*alto*6.022e-4 : the literal is single precision, promoted to double
  int n = (int)(1.0 * xi * yi * alto * (double)6.022e-4f);
  r.assign((size_t)n * 3, 0.0);
  InitGrid G(xi, yi, alto, r0);
  bool usegrid = fast && G.nx >= 3 && G.ny >= 3;
  for (int i = 0; i < n; ++i) {
    int k;
    for (k = 1; k <= 10000; ++k) {
      double p[3];
      p[0] = rng.ran() * xi; p[1] = rng.ran() * yi; p[2] = (rng.ran() * alto) + o;
      bool clash = false;
      auto test = [&](int j) {
        double vd[3];
        for (int l = 0; l < 3; ++l) vd[l] = p[l] - r[(size_t)j * 3 + l];          // distance(v2,v1): v1 - v2
        for (int l = 0; l < 2; ++l) vd[l] = vd[l] - box[l] * idnint(vd[l] * one_box[l]);
        double d2 = (vd[0] * vd[0] + vd[1] * vd[1]) + vd[2] * vd[2];
        return d2 < r0 * r0;
      };
      if (!usegrid) {
        for (int j = 0; j < i && !clash; ++j) clash = test(j);
      } else {
        int cx = G.ix(p[0]), cy = G.iy(p[1]), cz = G.iz(p[2]);
        for (int dz = -1; dz <= 1 && !clash; ++dz) {
          int z = cz + dz; if (z < 0 || z >= G.nz) continue;
          for (int dy = -1; dy <= 1 && !clash; ++dy) {
            int y = (cy + dy + G.ny) % G.ny;
            for (int dx = -1; dx <= 1 && !clash; ++dx) {
              int x = (cx + dx + G.nx) % G.nx;
              for (int j : G.c[(size_t)x + (size_t)G.nx * (y + (size_t)G.ny * z)]) if (test(j)) { clash = true; break; }
            }
          }
        }
      }
      if (clash) continue;
      for (int l = 0; l < 3; ++l) r[(size_t)i * 3 + l] = p[l];
      if (usegrid) G.c[(size_t)G.ix(p[0]) + (size_t)G.nx * (G.iy(p[1]) + (size_t)G.ny * G.iz(p[2]))].push_back(i);
      break;
    }
    if (k == 10001) throw Halt("Maximo numero de intentos alcanzado");
  }
  for (auto &x : r) x = quant12(x);
  return n;
}

void Dana::pos_inic(std::vector<double> &r) {
  tbox33 = zmax; box_setvars();              // dana.F90:341-346
  double alto = zmax - 0.0;
  n = gen_pos_inic(rng, xi, yi, alto, P.fast_init != 0, r, box, one_box);
}

void Dana::set_ermak() {                     // dana.F90:947-971
  cc0 = std::exp(-h * gama);
  cc1 = (1.0 - cc0) / gama;
  cc2 = (1.0 - cc1 / h) / gama;
  sdr = std::sqrt(h / gama * (2.0 - (3.0 - 4.0 * cc0 + cc0 * cc0) / (h * gama)));
  sdv = std::sqrt(1.0 - cc0 * cc0);
  crv1 = (1.0 - cc0) * (1.0 - cc0) / (gama * sdr * sdv);
  crv2 = std::sqrt(1.0 - (crv1 * crv1));
  skt = std::sqrt(kB_ui() * Tsist);
}

void Dana::config_inic(const std::vector<double> &r, const std::vector<int> &zz) {   // dana.F90:430-517
  if (s_chunk) { dist = dist + hs.rcut; z1 = z0 + dist; zmax = z1 + dist; }
  rhomedia = 5.775329e-4; cstdev = 1.0754306e-4;
  for (int i = 0; i < n; ++i) {
    Atom *pa = new_atom();
    for (int k = 0; k < 3; ++k) pa->pos[k] = r[(size_t)i * 3 + k];
    pa->setz(zz.empty() ? 1 : zz[i]);
    for (int k = 0; k < 3; ++k) { pa->force[k] = 0.0; pa->pos_old[k] = pa->pos[k]; }
    sys.attach(pa);
    if (pa->z == 2) hs.b.attach(pa);
    else { hs.ref.attach(pa); hs.b.attach(pa); }
    hs.attach(pa);                            // must be last (dana.F90:471-480)
    pa->pbc[0] = pa->pbc[1] = true; pa->pbc[2] = false;
  }
  ranv_n = n; ranv.assign((size_t)n * 3, 0.0);
  if (integrador) set_ermak();
}

void Dana::init(const orc_params &p) {
  P = p;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { eps[i][j] = 0; r0[i][j] = 0; }
  r0[2][1] = 3.5; r0[1][2] = 3.5;            // dana.F90:87-100
  eps[1][1] = 2313.6; r0[1][1] = 3.2;
  eps[3][3] = 121.0;  r0[3][3] = 3.61;
  eps[1][3] = 529.1;  eps[3][1] = eps[1][3];
  r0[1][3] = 1.564;   r0[3][1] = r0[1][3];
  sys.igroup_construct(this);                // id 1 (dana.F90:103-105)
  rng.idum = p.idum; prob = p.prob; h = p.h; nst = p.nst; nwr = p.nwr; xi = p.xi; yi = p.yi;
  dist = p.dist; z0 = p.z0; zmax = p.zmax; dif_sc = p.dif_sc; dif_sei = p.dif_sei; nb_dcut = p.nb_dcut;
  hs.mnb = p.mnb > 0 ? p.mnb : 10000;
  hs.ngroup_construct(this);                 // ids 2,3,4 (dana.F90:112)
  hs.setrc(3.2);                             // dana.F90:113
  std::vector<double> r; std::vector<int> zz;
  if (p.n_init > 0) {
    tbox33 = zmax; box_setvars();
    n = p.n_init; r.assign(p.init_xyz, p.init_xyz + (size_t)n * 3);
    if (p.init_z) zz.assign(p.init_z, p.init_z + n);
  } else pos_inic(r);
  integrador = p.integrador != 0;            // config_run — dana.F90:399-427
  if (p.reservoir == 1) s_piston = true;
  else if (p.reservoir == 2) { s_chunk = true; chunk.group_construct(this); }
  else if (p.reservoir == 3) { s_gcmc = true; act = p.act; nadj = p.nadj; }
  else throw Halt("Unknown reservoir type");
  config_inic(r, zz);
  if (s_gcmc) {                              // dana.F90:130-137
    gcmc.group_construct(this);
    Node *la = sys.alist;
    for (int j = 1; j <= sys.nat; ++j) { la = la->next; if (la->o->z == 1) gcmc.attach(la->o); }
  }
  test_update();                             // dana.F90:140
  if (integrador) fuerza();                  // dana.F90:142
  calc_rho(); rho0 = rho;                    // dana.F90:145-146
  if (s_chunk) {                             // config_chunk — dana.F90:552-587
    nchunk = p.nchunk;
    for (int j = 0; j < nchunk; ++j) {
      Atom *pb = new Atom;                   // template atoms are not part of sys: no uid
      for (int k = 0; k < 3; ++k) pb->pos[k] = p.chunk_xyz[(size_t)j * 3 + k];
      pb->setz(1);
      for (int k = 0; k < 3; ++k) { pb->force[k] = 0.0; pb->pos_old[k] = pb->pos[k]; }
      pb->pos[2] = pb->pos[2] + zmax;
      pb->pos_old[2] = pb->pos[2] + zmax;    // sic: shifted twice (dana.F90:573-574)
      chunk.attach(pb); chunk_atoms.push_back(pb);
      pb->pbc[0] = pb->pbc[1] = true; pb->pbc[2] = false;
    }
  }
  try_ = 0; depo = 0;
  salida();                                  // dana.F90:165
  try_ = 0; depo = 0;
}

Dana::~Dana() {
  // free atoms reachable from sys and the chunk templates
  std::vector<Atom *> all;
  Node *la = sys.alist ? sys.alist->next : nullptr;
  while (la && la != sys.alist) { all.push_back(la->o); la = la->next; }
  for (Atom *a : all) delete a;
  for (Atom *a : chunk_atoms) delete a;
}

// ---------------------------------------------------------------------------------------------
// Neighbour maintenance — Neighbor.F90:608-713, Groups.F90:1440-1467
// ---------------------------------------------------------------------------------------------
void Dana::do_pbc() {                        // Groups.F90:1440-1467 (over sys)
  Node *la = sys.alist;
  for (int i = 1; i <= sys.nat; ++i) {
    la = la->next; Atom *o = la->o;
    for (int k = 0; k < 3; ++k) if (o->pbc[k]) {
      if (o->pos[k] >= box[k]) { o->pos[k] = o->pos[k] - box[k]; o->pos_old[k] = o->pos_old[k] - box[k]; }
      else if (o->pos[k] < 0.0) { o->pos[k] = o->pos[k] + box[k]; o->pos_old[k] = o->pos_old[k] + box[k]; }
    }
  }
}
void Dana::update() {                        // Neighbor.F90:608-633
  nupd_vlist++;
  Node *la = sys.alist;
  for (int i = 1; i <= sys.nat; ++i) { la = la->next; for (int k = 0; k < 3; ++k) la->o->pos_old[k] = la->o->pos[k]; }
  hs.use_cells = hs.b.tessellated;           // ngroup_setlista — Neighbor.F90:336-353
  if (hs.use_cells) hs.build_cells(); else hs.build_verlet();
}
void Dana::test_update() {                   // Neighbor.F90:668-713
  do_pbc();
  hs.b.tessellate();
  if (hs.b.tessellated) hs.b.sort();
  if (!hs.listed) { update(); return; }
  double d1 = 1e-16, d2 = 1e-16;             // inq_dispmax — Neighbor.F90:635-666 (all atoms of hs)
  Node *la = hs.alist;
  for (int i = 1; i <= hs.nat; ++i) {
    la = la->next; Atom *o = la->o;
    double vd[3]; for (int k = 0; k < 3; ++k) vd[k] = o->pos[k] - o->pos_old[k];
    double rd = dot3(vd);
    if (rd > d1) { d2 = d1; d1 = rd; } else if (rd > d2) d2 = rd;
  }
  if (std::sqrt(d1) + std::sqrt(d2) > nb_dcut) update();
}

// ---------------------------------------------------------------------------------------------
// Physics — dana.F90
// ---------------------------------------------------------------------------------------------
void Dana::fuerza() {                        // dana.F90:1055-1139
  Node *la = hs.ref.alist;
  for (int i = 1; i <= hs.ref.nat; ++i) { la = la->next; Atom *o1 = la->o; o1->force[0] = o1->force[1] = o1->force[2] = 0.0; o1->epot = 0.0; }
  la = hs.ref.alist;
  for (int ii = 1; ii <= hs.ref.nat; ++ii) {
    la = la->next; Atom *o1 = la->o;
    int i = o1->gid(hs.id);
    int k = o1->z;
    for (int jj = 1; jj <= hs.nn[i]; ++jj) {
      int j = hs.L(i, jj);
      Atom *o2 = hs.a[j];
      if (hs.b_limbo && o2 == hs.limbo) continue;
      int m = o2->z;
      double vd[3];
      for (int l = 0; l < 3; ++l) vd[l] = o1->pos[l] - o2->pos[l];
      for (int l = 0; l < 2; ++l) {          // branchy minimum image, x and y only
        if (vd[l] > box[l] * .5) vd[l] = vd[l] - box[l];
        else if (vd[l] < -box[l] * .5) vd[l] = vd[l] + box[l];
      }
      if (o2->z == 2 && o1->z == 2) continue;
      double dr = dot3(vd);
      double r0km = r0[k][m], epskm = eps[k][m];
      if (dr > r0km * r0km) continue;
      dr = std::sqrt(dr);
      double b = pow6(r0km);
      double c = epskm * 12.0 * b;
      c = c / pow7(dr);
      b = b / pow6(dr);
      double aux = c * (b - 1.0);
      for (int l = 0; l < 3; ++l) {
        double f = aux * vd[l] / dr;
        o1->force[l] = o1->force[l] + f;
        o2->force[l] = o2->force[l] - f;
      }
      aux = epskm * b * (b - 2.0);
      aux = aux + epskm;
      o1->epot = o1->epot + aux * .5;
      o2->epot = o2->epot + aux * .5;
    }
  }
}

void Dana::atom_pbc(Atom *o1, bool &depos) { // dana.F90:1187-1250
  depos = false;
  for (int j = 0; j < 3; ++j) {
    if (j < 2) {
      if (o1->pos[j] > box[j]) { o1->pos[j] = o1->pos[j] - box[j]; o1->pos_old[j] = o1->pos_old[j] - box[j]; }
      if (o1->pos[j] < 0.0) { o1->pos[j] = o1->pos[j] + box[j]; o1->pos_old[j] = o1->pos_old[j] + box[j]; }
    } else if (o1->pos[j] > zmax) {
      if (integrador) { o1->pos[2] = o1->pos[2] - 2 * (o1->pos[2] - zmax); o1->vel[2] = -o1->vel[2]; }
      else for (int k = 0; k < 3; ++k) o1->pos[k] = o1->old_cg[k];
    }
  }
  msd_u = o1->vel[0] * o1->vel[0] * h * h;
  msd_t = msd_t + msd_u;
  if (o1->pos[2] <= 0.0) {
    try_ = try_ + 1;
    double ne = ran_t(ORC_TR_UNIF_PBC, o1->uid);
    if (ne < prob) { depo = depo + 1; o1->setz(3); depos = true; }
    for (int k = 0; k < 3; ++k) o1->pos[k] = o1->old_cg[k];
  }
}

void Dana::ermak_a() {                       // dana.F90:974-1028
  Node *la = hs.ref.alist;
  for (int i = 1; i <= hs.ref.nat; ++i) {
    la = la->next; Atom *o1 = la->o;
    for (int k = 0; k < 3; ++k) o1->old_cg[k] = o1->pos[k];
    if (i > ranv_n) throw Halt("oracle: ranv(i,:) out of bounds (ref grew under Ermak)");
    for (int j = 0; j < 3; ++j) {
      double r1 = gas_t(ORC_TR_GAUSS_INTEG, o1->uid);
      double ranr = skt / std::sqrt(o1->mass) * sdr * r1;
      o1->pos[j] = o1->pos[j] + cc1 * o1->vel[j] + cc2 * h * o1->acel[j] + ranr;
      double r2 = gas_t(ORC_TR_GAUSS_INTEG, o1->uid);
      ranv[(size_t)(i - 1) * 3 + j] = skt / std::sqrt(o1->mass) * sdv * (crv1 * r1 + crv2 * r2);
    }
    bool depos; atom_pbc(o1, depos);
    if (depos) continue;
    o1->skip = false;
  }
}
void Dana::ermak_b() {                       // dana.F90:1031-1052
  Node *la = hs.ref.alist;
  for (int i = 1; i <= hs.ref.nat; ++i) {
    la = la->next; Atom *o1 = la->o;
    if (o1->z == 2) continue;
    for (int k = 0; k < 3; ++k) {
      o1->vel[k] = cc0 * o1->vel[k] + (cc1 - cc2) * o1->acel[k] + cc2 * o1->force[k] / o1->mass + ranv[(size_t)(i - 1) * 3 + k];
      o1->acel[k] = o1->force[k] / o1->mass;
    }
  }
}
void Dana::cbrownian_hs() {                  // dana.F90:798-846
  Node *la = hs.ref.alist;
  for (int ii = 1; ii <= hs.ref.nat; ++ii) {
    la = la->next; Atom *o1 = la->o;
    for (int k = 0; k < 3; ++k) o1->old_cg[k] = o1->pos[k];
    if (o1->pos[2] > 80.0) dif = dif_sc; else dif = dif_sei;
    double fac1 = std::sqrt(2.0 * dif * h);
    for (int j = 0; j < 3; ++j) {
      double r1 = gas_t(ORC_TR_GAUSS_INTEG, o1->uid);
      double posold = o1->pos[j];
      o1->pos[j] = posold + r1 * fac1;
      o1->vel[j] = (o1->pos[j] - posold) / h;
    }
    bool depos; atom_pbc(o1, depos);
    if (depos) continue;
    max_vel = std::max(max_vel, dot3(o1->vel));
    o1->skip = false;
  }
}

void Dana::overlap_moveback() {              // dana.F90:849-943 (tail recursion written as a loop)
  std::vector<int64_t> marks;                // choques at entry of each recursion level
  for (int pass = 0;; ++pass) {
    bool again = false;
    // Optional (ov_guard_pass>0, NOT reference behaviour, default off): from that recursion depth on apply the
    // piston-mode guard in every mode, so pairs that overlap at their previous positions — on which the reference
    // recurses forever — are skipped and counted in choques3.  Used only by bench.py on large synthetic boxes.
    bool guard = s_piston || (P.ov_guard_pass > 0 && pass >= P.ov_guard_pass);
    Node *la = hs.ref.alist;
    for (int ii = 1; ii <= hs.ref.nat; ++ii) {
      la = la->next; Atom *o1 = la->o;
      if (o1->skip) continue;
      o1->skip = true;
      int i = o1->gid(hs.id);
      for (int jj = 1; jj <= hs.nn[i]; ++jj) {
        int j = hs.L(i, jj);
        Atom *o2 = hs.a[j];
        if (hs.b_limbo && o2 == hs.limbo) continue;
        double vd[3]; vdistance(vd, o1, o2);
        double dr = dot3(vd);
        if (dr > hs.rcut2) continue;
        if (o2->z == 2) {
          try_ = try_ + 1;
          double ne = ran_t(ORC_TR_UNIF_OVERLAP, o1->uid);
          if (ne < prob) {
            depo = depo + 1; o1->setz(3);
            if (o1->pos[2] > z0) throw Halt("supero z0");
          } else {
            for (int k = 0; k < 3; ++k) o1->pos[k] = o1->old_cg[k];
            o1->skip = false;
          }
          break;
        }
        if (guard) {
          bool same2 = o2->pos[0] == o2->old_cg[0] && o2->pos[1] == o2->old_cg[1] && o2->pos[2] == o2->old_cg[2];
          if (same2) {
            bool same1 = o1->pos[0] == o1->old_cg[0] && o1->pos[1] == o1->old_cg[1] && o1->pos[2] == o1->old_cg[2];
            if (same1) { choques3++; continue; }
          }
        }
        for (int k = 0; k < 3; ++k) { o2->pos[k] = o2->old_cg[k]; o2->acel[k] = 0.0; o2->vel[k] = 0.0; }
        o2->skip = false;
        choques++;
        again = true;
      }
    }
    marks.push_back(choques);
    if (!again) break;
  }
  // choques2=max(choques2,choques-i) evaluated while unwinding: i = choques at the end of that level's pass
  for (size_t lv = 0; lv < marks.size(); ++lv) choques2 = std::max(choques2, choques - marks[lv]);
}

void Dana::promote() {                       // dana.F90:228-236
  Node *la = hs.ref.alist;
  int nloop = hs.ref.nat;
  for (int j = 1; j <= nloop; ++j) {
    la = la->next; Atom *o1 = la->o;
    if (o1->z != 3) continue;
    o1->setz(2);
    hs.ref.detach(o1, &la);
    if (s_gcmc) gcmc.detach(o1);
  }
}

void Dana::destroy_atom(Atom *a) {           // atom_destroy — Groups.F90:433-467 (LIFO by group id)
  while (a->ngr != 0) { Group *g = gindex[a->gr[a->ngr - 1] - 1]; g->detach(a); }
  delete a;
}

void Dana::gcmc_run() {                      // dana.F90:590-713
  Group &g = gcmc;
  double rc = hs.rcut;
  double v = box[0] * box[1] * (zmax - z0);
  int nin = 0;
  Node *la = g.alist;
  for (int j = 1; j <= g.nat; ++j) { la = la->next; if (la->o->pos[2] < z0 || la->o->pos[2] > zmax) continue; nin++; }
  for (int i = 1; i <= nadj; ++i) {
    Atom *ref = g.alist->next->o;
    if (!ref) throw Halt("No more particles");
    double beta = std::sqrt(kB_ui_module() * Tsist / ref->mass);
    if (ran_t(ORC_TR_UNIF_GCMC, -1) < 0.5) {
      if (act * v / (nin + 1) < ran_t(ORC_TR_UNIF_GCMC, -1)) continue;
      double r[3];
      r[0] = ran_t(ORC_TR_UNIF_GCMC, -1) * box[0];
      r[1] = ran_t(ORC_TR_UNIF_GCMC, -1) * box[1];
      r[2] = ran_t(ORC_TR_UNIF_GCMC, -1) * (zmax - z0) + z0;
      bool clash = false;
      la = g.alist;
      for (int j = 1; j <= g.nat; ++j) {
        la = la->next; Atom *o = la->o;
        double vd[3]; distance(vd, o->pos, r, o->pbc);
        double dr = dot3(vd);
        if (dr < rc * rc) { clash = true; break; }
      }
      if (clash) continue;
      nin = nin + 1;
      Atom *o = new_atom();
      // atom_asign(o,ref) — Groups.F90:484-502
      for (int k = 0; k < 3; ++k) { o->pos[k] = ref->pos[k]; o->vel[k] = ref->vel[k]; o->force[k] = ref->force[k]; o->acel[k] = ref->acel[k]; o->pbc[k] = ref->pbc[k]; o->pos_old[k] = ref->pos_old[k]; }
      o->setz(ref->z); o->epot = ref->epot;
      for (int k = 0; k < 3; ++k) { o->pos[k] = r[k]; o->pos_old[k] = r[k]; }
      // la still points at the LAST atom of the gcmc list: the velocity lands there
      for (int j = 0; j < 3; ++j) la->o->vel[j] = beta * gas_t(ORC_TR_GAUSS_GCMC, la->o->uid);
      int ngr = ref->ngr; std::vector<int> grs(ref->gr.begin(), ref->gr.begin() + ngr);
      for (int j = 0; j < ngr; ++j) gindex[grs[j] - 1]->attach(o);
    } else {
      if ((double)nin / (v * act) < ran_t(ORC_TR_UNIF_GCMC, -1)) continue;
      int m = (int)std::floor(ran_t(ORC_TR_UNIF_GCMC, -1) * nin) + 1;
      if (m > nin) m = nin;
      Atom *o = nullptr;
      la = g.alist;
      for (int j = 1; j <= g.nat; ++j) {
        la = la->next; o = la->o;
        if (o->pos[2] < z0) continue;
        if (o->pos[2] > zmax) continue;
        m = m - 1;
        if (m == 0) break;
      }
      if (m > 0) throw Halt("Chosen particle does not exists");
      nin = nin - 1;
      destroy_atom(o);
    }
  }
}

void Dana::calc_rho() {                      // dana.F90:521-549
  int gct = 0;
  double z = s_chunk ? z1 : zmax;
  Node *la = sys.alist;
  for (int i = 1; i <= sys.nat; ++i) { la = la->next; Atom *pa = la->o; if (pa->pos[2] > z0 && pa->pos[2] < z) gct++; }
  double vol = box[0] * box[1] * (z - z0);
  rho = gct / vol;
}

void Dana::bloques() {                       // dana.F90:716-773
  double drho = rho - rhomedia;
  if (std::fabs(drho) < (rhomedia * (double)0.186f)) return;   // 0.186 is a single-precision literal
  z0 = z0 + dist; z1 = z1 + dist; zmax = zmax + dist;
  hs.listed = false;
  Node *la = chunk.alist;
  for (int j = 1; j <= chunk.nat; ++j) {
    la = la->next; Atom *o1 = la->o;
    Atom *o2 = new_atom();
    sys.attach(o2);
    hs.b.attach(o2); hs.ref.attach(o2); hs.attach(o2);
    for (int k = 0; k < 3; ++k) { o2->pos[k] = o1->pos[k]; o2->vel[k] = o1->vel[k]; o2->force[k] = o1->force[k]; o2->acel[k] = o1->acel[k]; o2->pbc[k] = o1->pbc[k]; o2->pos_old[k] = o1->pos_old[k]; }
    o2->setz(o1->z); o2->epot = o1->epot;
    o1->pos[2] = o1->pos[2] + dist;
    o1->pos_old[2] = o1->pos_old[2] + dist;
  }
  n = n + nx;
  tbox33 = zmax; box_setvars();
  test_update();
}

void Dana::maxz() {                          // dana.F90:776-794
  double lohi = ((h / tau) * ((rho0 - rho) / rho));
  for (int i = 1; i <= sys.nat; ++i) { Atom *pa = sys.a[i]; if (pa->pos[2] > z0) pa->pos[2] = pa->pos[2] - lohi * (pa->pos[2] - z0); }
  zmax = zmax - lohi * (zmax - z0);
}

void Dana::salida() {                        // dana.F90:1143-1183 + kion 1342-1376
  frame.nat = sys.nat; frame.zmax = zmax;
  frame.z.resize(sys.nat); frame.pos.resize((size_t)sys.nat * 3);
  double energia = 0.0, vdac = 0.0; int jm = 0;
  Node *la = sys.alist;
  for (int j = 0; j < sys.nat; ++j) {
    la = la->next; Atom *pa = la->o;
    frame.z[j] = pa->z; for (int k = 0; k < 3; ++k) frame.pos[(size_t)j * 3 + k] = pa->pos[k];
    energia = energia + pa->epot;
    if (pa->z != 2) { jm++; double vd = dot3(pa->vel); vd = vd * pa->mass; vdac = vdac + vd; }
  }
  frame.scal[0] = t; frame.scal[1] = energia; frame.scal[2] = vdac / (jm * 3.0 * kB_ui());
  frame.scal[3] = rho; frame.scal[4] = (double)try_; frame.scal[5] = (double)depo;
  try_ = 0; depo = 0;
}

void Dana::step_once() {                     // loop body — dana.F90:173-265
  step++;
  if (integrador) { ermak_a(); fuerza(); ermak_b(); } else cbrownian_hs();
  test_update();
  overlap_moveback();
  test_update();
  msd_book();
  promote();
  if (s_gcmc) gcmc_run();
  calc_rho();
  if (s_chunk) bloques();
  if (step % nwr == 0) salida();
  if (s_piston) maxz();
  t = t + h;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// C interface
// ---------------------------------------------------------------------------------------------
#define GUARD(D, expr) try { expr; } catch (const std::exception &e) { (D)->err = e.what(); return -1; }

extern "C" {

void *orc_create(const orc_params *p) {
  Dana *D = new Dana;
  try { D->init(*p); } catch (const std::exception &e) { D->err = std::string("init: ") + e.what(); }
  return D;
}
void orc_destroy(void *h) { delete (Dana *)h; }
const char *orc_last_error(void *h) { return ((Dana *)h)->err.c_str(); }

int orc_step(void *h, int nsteps) {
  Dana *D = (Dana *)h;
  GUARD(D, for (int i = 0; i < nsteps; ++i) D->step_once());
  return 0;
}
int orc_call(void *h, int op) {
  Dana *D = (Dana *)h;
  GUARD(D, switch (op) {
    case ORC_ERMAK_A: D->step++; D->ermak_a(); break;
    case ORC_CBROWNIAN: D->step++; D->cbrownian_hs(); break;
    case ORC_FUERZA: D->fuerza(); break;
    case ORC_ERMAK_B: D->ermak_b(); break;
    case ORC_TEST_UPDATE: D->test_update(); break;
    case ORC_OVERLAP: D->overlap_moveback(); break;
    case ORC_PROMOTE: D->promote(); break;
    case ORC_GCMC: if (D->s_gcmc) D->gcmc_run(); break;
    case ORC_CALC_RHO: D->calc_rho(); break;
    case ORC_BLOQUES: if (D->s_chunk) D->bloques(); break;
    case ORC_SALIDA: D->salida(); break;
    case ORC_MAXZ: if (D->s_piston) D->maxz(); break;
    case ORC_MSD: D->msd_book(); break;
    case ORC_STEP_END: D->t = D->t + D->h; break;
    default: throw Halt("bad op");
  });
  return 0;
}

void orc_get_scalars(void *h, orc_scalars *s) {
  Dana *D = (Dana *)h;
  memset(s, 0, sizeof *s);
  for (int k = 0; k < 3; ++k) { s->box[k] = D->box[k]; s->cell[k] = D->hs.b.cell[k]; s->ncells[k] = D->hs.b.ncells[k]; }
  s->z0 = D->z0; s->z1 = D->z1; s->zmax = D->zmax; s->rho = D->rho; s->rho0 = D->rho0; s->t = D->t; s->h = D->h;
  s->tessellated = D->hs.b.tessellated; s->listed = D->hs.listed;
  s->nat_sys = D->sys.nat; s->nat_ref = D->hs.ref.nat; s->nat_b = D->hs.b.nat; s->nat_hs = D->hs.nat;
  s->nat_gcmc = D->s_gcmc ? D->gcmc.nat : 0;
  s->hs_amax = D->hs.amax; s->b_amax = D->hs.b.amax;
  s->nupd = D->nupd_vlist; s->choques = D->choques; s->choques2 = D->choques2; s->choques3 = D->choques3;
  s->try_ = D->try_; s->depo = D->depo; s->max_vel = D->max_vel; s->msd_t = D->msd_t; s->msd_max = D->msd_max;
  s->ran_calls = D->rng.calls; s->step = D->step;
  s->cc0 = D->cc0; s->cc1 = D->cc1; s->cc2 = D->cc2; s->sdr = D->sdr; s->sdv = D->sdv; s->crv1 = D->crv1; s->crv2 = D->crv2; s->skt = D->skt;
}

int orc_get_state(void *h, int64_t *uid, int32_t *z, double *pos, double *vel, double *acel, double *force,
                  double *epot, double *pos_old, double *old_cg, int32_t *flags, int32_t *slot_hs, int32_t *slot_b) {
  Dana *D = (Dana *)h;
  Node *la = D->sys.alist;
  int gc = D->s_gcmc ? D->gcmc.id : -1;
  for (int i = 0; i < D->sys.nat; ++i) {
    la = la->next; Atom *a = la->o;
    if (uid) uid[i] = a->uid;
    if (z) z[i] = a->z;
    for (int k = 0; k < 3; ++k) {
      if (pos) pos[3 * i + k] = a->pos[k];
      if (vel) vel[3 * i + k] = a->vel[k];
      if (acel) acel[3 * i + k] = a->acel[k];
      if (force) force[3 * i + k] = a->force[k];
      if (pos_old) pos_old[3 * i + k] = a->pos_old[k];
      if (old_cg) old_cg[3 * i + k] = a->old_cg[k];
    }
    if (epot) epot[i] = a->epot;
    if (flags) flags[i] = (a->gri(D->hs.ref.id) ? 1 : 0) | ((gc > 0 && a->gri(gc)) ? 2 : 0) | (a->skip ? 4 : 0);
    if (slot_hs) slot_hs[i] = a->gid(D->hs.id);
    if (slot_b) slot_b[i] = a->gid(D->hs.b.id);
  }
  return D->sys.nat;
}

int orc_get_rows(void *h, int32_t width, int32_t *nn, int32_t *rows, int64_t *slot_uid) {
  Dana *D = (Dana *)h; NGroup &g = D->hs;
  int over = 0;
  for (int i = 1; i <= g.amax; ++i) {
    Atom *a = g.a[i];
    if (slot_uid) slot_uid[i - 1] = a == nullptr ? -1 : (a == g.limbo ? -2 : a->uid);
    int c = (i < (int)g.nn.size()) ? g.nn[i] : 0;
    if (nn) nn[i - 1] = c;
    if (rows) for (int m = 1; m <= c; ++m) { if (m > width) { over = 1; break; } rows[(size_t)(i - 1) * width + (m - 1)] = g.L(i, m); }
  }
  return over ? -1 : g.amax;
}

int orc_get_cells(void *h, int32_t *cell_of_slot, int32_t *chain_pos) {
  Dana *D = (Dana *)h; CGroup &b = D->hs.b;
  for (int i = 0; i < b.amax; ++i) { cell_of_slot[3 * i] = cell_of_slot[3 * i + 1] = cell_of_slot[3 * i + 2] = -1; chain_pos[i] = -1; }
  if (!b.tessellated) return 0;
  for (int k = 0; k < b.hd[2]; ++k) for (int j = 0; j < b.hd[1]; ++j) for (int i = 0; i < b.hd[0]; ++i) {
    int s = b.H(i, j, k), p = 0;
    while (s > 0) { cell_of_slot[3 * (s - 1)] = i; cell_of_slot[3 * (s - 1) + 1] = j; cell_of_slot[3 * (s - 1) + 2] = k; chain_pos[s - 1] = p++; s = b.next[s]; }
  }
  return b.amax;
}

int orc_get_frame(void *h, int32_t *nat, double *zmax, int32_t *z, double *pos, double *scal) {
  Dana *D = (Dana *)h; Frame &f = D->frame;
  if (nat) *nat = f.nat;
  if (zmax) *zmax = f.zmax;
  if (z) memcpy(z, f.z.data(), f.z.size() * sizeof(int32_t));
  if (pos) memcpy(pos, f.pos.data(), f.pos.size() * sizeof(double));
  if (scal) memcpy(scal, f.scal, sizeof f.scal);
  return f.nat;
}

void orc_trace_enable(void *h, int on) { ((Dana *)h)->tracing = on != 0; }
void orc_trace_clear(void *h) { ((Dana *)h)->trace.clear(); }
int64_t orc_trace_size(void *h) { return (int64_t)((Dana *)h)->trace.size(); }
void orc_trace_get(void *h, int32_t *kind, int64_t *uid, double *val) {
  Dana *D = (Dana *)h;
  for (size_t i = 0; i < D->trace.size(); ++i) { kind[i] = D->trace[i].kind; uid[i] = D->trace[i].uid; val[i] = D->trace[i].val; }
}

int orc_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}
void orc_rng_kat(int32_t idum, int n_ran, double *ran_out, int n_gas, double *gas_out) {
  Rng a; a.idum = idum; for (int i = 0; i < n_ran; ++i) ran_out[i] = a.ran();
  Rng b; b.idum = idum; for (int i = 0; i < n_gas; ++i) gas_out[i] = b.gasdev();
}

int orc_pos_inic(int32_t idum, double xi, double yi, double alto, int fast, double *xyz, int cap, uint64_t *ran_calls) {
  Rng rng; rng.idum = idum;
  double box[3] = {xi, yi, alto}, one_box[3] = {1.0 / xi, 1.0 / yi, 1.0 / alto};
  std::vector<double> r;
  int n;
  try { n = gen_pos_inic(rng, xi, yi, alto, fast != 0, r, box, one_box); } catch (...) { return -1; }
  if (n > cap) return -n;
  memcpy(xyz, r.data(), (size_t)n * 3 * sizeof(double));
  if (ran_calls) *ran_calls = rng.calls;
  return n;
}

} // extern "C"
