// tools/dana_host.cpp — C++ stand-in for the reference's main program (src/dana.F90:1-305) driving libdml.so
// through the same C ABI the Fortran shim binds (fortran/dml_cuda.F90 cannot be compiled in this image).
//
//   dana_b200 [case_dir] [--steps N] [--seed S] [--device D] [--rng philox|reference]
//
// Reads entrada.ini / movedor.ini / chunk.xyz from case_dir (dana.F90:309-327,399-427,552-587), builds the initial
// configuration with the reference's pos_inic rule and RNG (dmlh_pos_inic, stream-identical), runs the loop on the GPU
// and writes pos_inic.xyz, Li.xyz, E.dat, T.dat, rho.dat, try.dat, depo.dat in the reference's list-directed layout
// (dana.F90:1143-1183).  By default the hot path uses counter-based Philox noise, so trajectories are statistically — not
// bitwise — equivalent to the Fortran binary.  With --rng reference the device draws the reference's own ran / gasdev stream
// (src/dana.F90:1379-1428) in the reference's order, continuing from the state pos_inic left, and the last frame of Li.xyz is
// the reference's ref.xyz digit for digit: tools/test_cases.sh is the reference's tests/test.sh run on this binary.
#include "../include/dml.h"
#include "../include/dml_host.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

static std::string first_token(const std::string &line) {
  std::string s = line.substr(0, line.find('!'));
  std::istringstream is(s); std::string t; is >> t; return t;
}
static std::vector<std::string> read_values(const std::string &path) {
  std::ifstream f(path); std::vector<std::string> v; std::string line;
  if (!f) { fprintf(stderr, "#-ERR-> cannot open %s\n", path.c_str()); exit(1); }
  while (std::getline(f, line)) { std::string s = line.substr(0, line.find('!')); if (s.find_first_not_of(" \t\r") != std::string::npos) v.push_back(s); }
  return v;
}

// gfortran list-directed real(8): one blank + G25.17E3 (F form right-justified in 20 columns + 5 blanks, or 1P E form)
static std::string fort_real(double x) {
  char buf[64], out[64];
  if (x == 0.0) { snprintf(out, sizeof out, "   0.0000000000000000     "); return out; }
  snprintf(buf, sizeof buf, "%.16E", x);
  int e = atoi(strchr(buf, 'E') + 1);
  if (e >= -1 && e <= 16) {
    snprintf(buf, sizeof buf, "%.*f", 16 - e, x);
    snprintf(out, sizeof out, " %20s     ", buf);
  } else {
    char mant[40]; strncpy(mant, buf, strchr(buf, 'E') - buf); mant[strchr(buf, 'E') - buf] = 0;
    snprintf(buf, sizeof buf, "%sE%c%03d", mant, e < 0 ? '-' : '+', abs(e));
    snprintf(out, sizeof out, " %25s", buf);
  }
  return out;
}

#define CHECK(rc) do { if ((rc) != 0) { fprintf(stderr, "#-ERR-> %s\n", ctx ? dml_last_error(ctx) : "dml_create failed (no CUDA device? libdml has no CPU fallback)"); return 1; } } while (0)

int main(int argc, char **argv) {
  std::string dir = ".";
  long steps_override = -1; long seed = 20240101; int device = 0; bool rng_reference = false;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps_override = atol(argv[++i]);
    else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = atol(argv[++i]);
    else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--rng") && i + 1 < argc) rng_reference = !strcmp(argv[++i], "reference");
    else dir = argv[i];
  }
  // entrada() — dana.F90:309-327
  auto e = read_values(dir + "/entrada.ini");
  if (e.size() < 13) { fprintf(stderr, "#-ERR-> entrada.ini needs 13 values\n"); return 1; }
  int idum = atoi(first_token(e[0]).c_str());
  double prob = atof(first_token(e[1]).c_str()), h = atof(first_token(e[2]).c_str());
  long nst = atol(first_token(e[3]).c_str()), nwr = atol(first_token(e[4]).c_str());
  double xi = atof(first_token(e[5]).c_str()), yi = atof(first_token(e[6]).c_str()), dist = atof(first_token(e[7]).c_str());
  double z0 = atof(first_token(e[8]).c_str()), zmax = atof(first_token(e[9]).c_str());
  double dif_sc = atof(first_token(e[10]).c_str()), dif_sei = atof(first_token(e[11]).c_str()), nb_dcut = atof(first_token(e[12]).c_str());
  if (steps_override >= 0) nst = steps_override;
  // config_run() — dana.F90:399-427
  auto m = read_values(dir + "/movedor.ini");
  std::string integ = first_token(m[0]), res = first_token(m[1]);
  bool integrador = integ.size() > 1 && (integ[1] == 't' || integ[1] == 'T');
  int reservoir = res == "piston" ? 1 : res == "chunks" ? 2 : res == "gcmc" ? 3 : 0;
  if (!reservoir) { fprintf(stderr, "#-ERR-> Unknown reservoir type\n"); return 1; }
  double act = 0; int nadj = 0;
  if (reservoir == 3) { std::istringstream is(m[2].substr(0, m[2].find('!'))); is >> act >> nadj; }

  // pos_inic() — dana.F90:330-396
  dmlh_rng rng; dmlh_rng_init(&rng, idum);
  int cap0 = (int)(xi * yi * zmax * 6.1e-4) + 16;
  std::vector<double> pos((size_t)cap0 * 3);
  int n = dmlh_pos_inic(&rng, xi, yi, zmax, pos.data(), cap0);
  if (n < 0) { fprintf(stderr, "Maximo numero de intentos alcanzado\n"); return 1; }
  {
    FILE *f = fopen((dir + "/pos_inic.xyz").c_str(), "w");
    fprintf(f, "%12d\n\n", n);
    for (int i = 0; i < n; ++i) fprintf(f, "Li %25.12f %25.12f %25.12f %25.12f\n", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 6.94);
    fclose(f);
  }
  // config_inic() — dana.F90:430-517
  double z1 = 0.0, box3 = zmax;
  const double rhomedia = 5.775329e-4;
  if (reservoir == 2) { dist = dist + 3.2; z1 = z0 + dist; zmax = z1 + dist; }

  dml_config c; memset(&c, 0, sizeof c);
  c.device = device; c.capacity = (int)(n * 2.5) + 65536;
  c.box[0] = xi; c.box[1] = yi; c.box[2] = box3; c.pbc[0] = c.pbc[1] = 1; c.pbc[2] = 0;
  c.rcut = 3.2; c.nb_dcut = nb_dcut;
  c.eps[0] = 2313.6; c.r0[0] = 3.2; c.eps[8] = 121.0; c.r0[8] = 3.61; c.eps[2] = c.eps[6] = 529.1; c.r0[2] = c.r0[6] = 1.564;
  c.r0[1] = c.r0[3] = 3.5;                                  // dana.F90:87-100
  c.mass[0] = c.mass[1] = c.mass[2] = 6.94;
  c.h = h; c.gama = 1.0; c.Tsist = 300.0; c.kB_ui = 8.617330350e-5 * (96.485 * 100.0);
  { const double ui_ev = 1.0e2 * 1.0e2 * 1.6605402e-27 * (1.0 / 1.60219e-19); c.kB_ui_gcmc = 8.617385e-05 * (1.0 / ui_ev); }
  c.dif_sc = dif_sc; c.dif_sei = dif_sei; c.z_sei = 80.0; c.prob = prob; c.z0 = z0; c.z1 = z1; c.zmax = zmax; c.tau = 0.1;
  c.act = act; c.nadj = nadj; c.integrador = integrador; c.reservoir = reservoir; c.rng_mode = rng_reference ? DML_RNG_REFERENCE : DML_RNG_PHILOX; c.seed = (uint64_t)seed;
  c.strict_order = rng_reference ? 1 : 0;                   // digit-for-digit frames need the reference's summation order in fuerza
  dml_ctx *ctx = nullptr;
  CHECK(dml_create(&ctx, &c));
  if (rng_reference) CHECK(dml_set_rng_state(ctx, &rng));      // the stream goes on where pos_inic stopped (one generator for the whole program)

  std::vector<int32_t> z(n, 1), flags(n, DML_F_REF | (reservoir == 3 ? DML_F_GCMC : 0));
  std::vector<double> og((size_t)n * 3, 1e8);
  CHECK(dml_upload(ctx, n, pos.data(), nullptr, nullptr, pos.data(), og.data(), z.data(), flags.data(), nullptr, nullptr));
  CHECK(dml_test_update(ctx));                              // dana.F90:140
  if (integrador) CHECK(dml_fuerza(ctx));                   // dana.F90:142
  double rho = 0; CHECK(dml_calc_rho(ctx, &rho));           // dana.F90:145-146
  dml_scalars sc; CHECK(dml_get_scalars(ctx, &sc)); sc.rho0 = rho; sc.rho = rho; CHECK(dml_set_scalars(ctx, &sc));
  if (reservoir == 2) {                                     // config_chunk — dana.F90:552-587
    std::ifstream f(dir + "/chunk.xyz"); int nch; std::string line; f >> nch; std::getline(f, line); std::getline(f, line);
    std::vector<double> cp((size_t)nch * 3), co((size_t)nch * 3);
    for (int j = 0; j < nch; ++j) { std::string sym; double mm; f >> sym >> cp[3 * j] >> cp[3 * j + 1] >> cp[3 * j + 2]; std::getline(f, line); (void)mm; }
    co = cp;
    for (int j = 0; j < nch; ++j) { cp[3 * j + 2] += zmax; co[3 * j + 2] = cp[3 * j + 2] + zmax; }
    CHECK(dml_set_chunk_template(ctx, nch, cp.data(), co.data(), dist, rhomedia));
  }

  FILE *fxyz = fopen((dir + "/Li.xyz").c_str(), "w"), *fe = fopen((dir + "/E.dat").c_str(), "w"), *ft = fopen((dir + "/T.dat").c_str(), "w"),
       *fr = fopen((dir + "/rho.dat").c_str(), "w"), *ftry = fopen((dir + "/try.dat").c_str(), "w"), *fd = fopen((dir + "/depo.dat").c_str(), "w");
  std::vector<double> P, V, EP; std::vector<int32_t> Z, FL, UID;
  const bool device_sums = getenv("DANA_DEVICE_SUMS") != nullptr;   // E.dat / T.dat from dml_salida_sums instead of the host loop
  auto salida = [&](double t) -> int {                      // dana.F90:1143-1183 (+kion 1342-1376), atoms in sys%alist order
    dml_counters k; if (dml_get_counters(ctx, &k)) return 1;
    int ns = k.n_slots;
    P.resize((size_t)ns * 3); V.resize((size_t)ns * 3); EP.resize(ns); Z.resize(ns); FL.resize(ns); UID.resize(ns);
    if (dml_download(ctx, ns, P.data(), V.data(), nullptr, nullptr, EP.data(), nullptr, nullptr, Z.data(), FL.data(), UID.data(), nullptr)) return 1;
    dml_scalars s; if (dml_get_scalars(ctx, &s)) return 1;
    std::vector<int> order; order.reserve(ns);
    for (int i = 0; i < ns; ++i) if (Z[i] > 0) order.push_back(i);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return UID[a] < UID[b]; });
    fprintf(fxyz, "%12d\n info:%s%12d\n", (int)order.size(), fort_real(s.zmax).c_str(), (int)order.size());
    double energia = 0, vdac = 0; int jm = 0;
    static const char *sym[4] = {"", "Li", "CG", "F"};
    for (int i : order) {
      fprintf(fxyz, " %s%s%s%s%12d\n", sym[Z[i]], fort_real(P[3 * i]).c_str(), fort_real(P[3 * i + 1]).c_str(), fort_real(P[3 * i + 2]).c_str(), Z[i]);
      energia += EP[i];
      if (Z[i] != 2) { jm++; vdac += ((V[3 * i] * V[3 * i] + V[3 * i + 1] * V[3 * i + 1]) + V[3 * i + 2] * V[3 * i + 2]) * 6.94; }
    }
    double temp_out = vdac / (jm * 3.0 * c.kB_ui);
    if (device_sums) {                                        // E.dat / T.dat from the device reductions (dml_salida_sums); the host loop above stays as the check
      double e_dev = 0, e_ref = 0, t_dev = 0; int32_t j_dev = 0;
      if (dml_salida_sums(ctx, &e_dev, &e_ref, &t_dev, &j_dev)) return 1;
      if (j_dev != jm || std::fabs(e_dev - energia) > 1e-9 * std::max(1.0, std::fabs(energia))) { fprintf(stderr, "dml_salida_sums disagrees with the host sums\n"); return 1; }
      energia = e_dev; temp_out = t_dev;
    }
    fprintf(fe, "%s%s\n", fort_real(t).c_str(), fort_real(energia).c_str());
    fprintf(ft, "%s%s\n", fort_real(t).c_str(), fort_real(temp_out).c_str());
    fprintf(fr, "%s%s\n", fort_real(t).c_str(), fort_real(s.rho).c_str());
    fprintf(ftry, "%s%12lld\n", fort_real(t).c_str(), (long long)k.try_);
    fprintf(fd, "%s%12lld\n", fort_real(t).c_str(), (long long)k.depo);
    fflush(fxyz); fflush(fe); fflush(ft); fflush(fr); fflush(ftry); fflush(fd);
    return dml_reset_try_depo(ctx);
  };
  CHECK(dml_reset_try_depo(ctx));
  CHECK(salida(0.0));                                       // dana.F90:165
  double t = 0.0;
  for (long i = 0; i < nst;) {                              // dana.F90:173-265, nwr steps per device call
    long chunk = std::min(nwr - (i % nwr), nst - i);
    bool frame = ((i + chunk) % nwr) == 0;
    if (reservoir == 1 && frame) {
      // salida() runs BEFORE maxz inside a step (dana.F90:250 vs 261): run the last step of the block call site by
      // call site so that the frame holds pre-piston coordinates like the reference's Li.xyz
      if (chunk > 1) CHECK(dml_step(ctx, (int)chunk - 1));
      if (integrador) { CHECK(dml_ermak_a(ctx)); CHECK(dml_fuerza(ctx)); CHECK(dml_ermak_b(ctx)); } else CHECK(dml_cbrownian_hs(ctx));
      CHECK(dml_test_update(ctx)); CHECK(dml_overlap_moveback(ctx)); CHECK(dml_test_update(ctx));
      CHECK(dml_msd_book(ctx)); CHECK(dml_promote(ctx)); CHECK(dml_calc_rho(ctx, &rho));
      i += chunk; t += h * chunk;
      CHECK(salida(t));
      double zm; CHECK(dml_maxz(ctx, &zm));
      continue;
    }
    CHECK(dml_step(ctx, (int)chunk));
    i += chunk; t += h * chunk;
    if (i % nwr == 0) CHECK(salida(t));
  }
  dml_counters k; CHECK(dml_get_counters(ctx, &k));
  printf("#-STD-> vecinos actualizados: %lld veces\n#-STD-> numero total de choques: %lld\n#-STD-> numero total de choques en 2da vuelta: %lld\n"
         "#-STD-> numero total de choques sin solucion: %lld\n#-STD-> MSD maximo en x-y: %10.3E\n#-STD-> particulas: %d (ref %d)\n",
         (long long)k.nupd_vlist, (long long)k.choques, (long long)k.choques2, (long long)k.choques3, k.msd_max, k.nat_sys, k.nat_ref);
  fclose(fxyz); fclose(fe); fclose(ft); fclose(fr); fclose(ftry); fclose(fd);
  dml_destroy(ctx);
  return 0;
}
