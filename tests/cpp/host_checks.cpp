/*
 * host_checks.cpp -- CPU-side check of the product's host-compilable pieces against the oracle:
 *   (1) geometry.h (the fp64 grid_map geometry the kernels run) vs oracle/himm_oracle.cpp, bit for bit;
 *   (2) the closed-form Bresenham + rectangle clipping used by himm_tile_kernel vs the oracle's stepping iterator;
 *   (3) vfh_tables.cpp (the product's VFH::Init) vs the reference VFH class in oracle/_ref.
 * Test infrastructure: links the oracle as the checker.  Prints "OK <n checks>" or the first mismatch.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <random>
#include <vector>

#include "../../oracle/oracle_api.h"
#include "../../ros_navigation_b200/csrc/geometry.h"
#include "../../ros_navigation_b200/csrc/vfh_tables.h"

extern "C" {
void* vfhref_create(const double* p);
void vfhref_destroy(void*);
int vfhref_num_tables(void*);
int vfhref_hist_size(void*);
void vfhref_get_cell_tables(void*, float*, float*, float*);
int vfhref_get_sector_masks(void*, int, unsigned*, int);
void vfhref_get_min_turning_radius(void*, int*, int);
}

using namespace b200nav;

static long long g_checks = 0;
#define FAIL(...)                 \
  do {                            \
    printf("MISMATCH: " __VA_ARGS__); \
    printf("\n");                 \
    return 1;                     \
  } while (0)

static bool same_bits(double a, double b) { return std::memcmp(&a, &b, 8) == 0; }

/* the XU-free fp64 helpers of geometry.h against the plain C++ expressions they stand for */
static int check_f64_helpers(unsigned seed, long long n) {
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  const double res_list[] = {0.05, 0.02, 1.0, 0.1, 0.03, 0.025, 0.3, 1.0 / 3.0, 0.0499999999999, 7.0, 0.001};
  for (double res : res_list) {
    const double y = 1.0 / res;
    for (long long i = 0; i < n; i++) {
      double a;
      switch (i % 4) {
        case 0: a = U(rng) * 200.0; break;
        case 1: a = res * (double)(long long)(U(rng) * 4000.0); break;             /* near exact multiples */
        case 2: a = res * (double)(long long)(U(rng) * 4000.0) * (1.0 + U(rng) * 4e-16); break;
        default: a = U(rng) * 1e-3; break;
      }
      const double want = a / res, got = f64_div_by(a, res, y);
      g_checks++;
      if (!same_bits(want, got)) FAIL("f64_div_by(%.17g, %.17g): %.17g vs %.17g", a, res, got, want);
      if (std::fabs(want) < 2147483000.0) {
        if (f64_trunc_to_int(want) != static_cast<int>(want)) FAIL("f64_trunc_to_int(%.17g)", want);
      }
    }
  }
  for (long long i = 0; i < n; i++) {
    const int k = (int)rng();
    g_checks++;
    if (!same_bits(int_to_f64(k), static_cast<double>(k))) FAIL("int_to_f64(%d)", k);
    const double v = (double)((long long)(rng() % 4000001) - 2000000) + ((i & 1) ? 0.0 : U(rng));
    if (f64_trunc_to_int(v) != static_cast<int>(v)) FAIL("f64_trunc_to_int(%.17g)", v);
    unsigned u = (unsigned)rng();
    if (i % 7 == 0) u &= 0x807fffffu; /* zeros / subnormals */
    float f;
    std::memcpy(&f, &u, 4);
    const double d1 = f32_to_f64(f), d2 = (double)f;
    if (!(same_bits(d1, d2) || (std::isnan(d1) && std::isnan(d2)))) FAIL("f32_to_f64(0x%08x)", u);
  }
  const int ks[] = {0, 1, -1, 2147483647, -2147483647 - 1, 32767, -32768};
  for (int k : ks)
    if (!same_bits(int_to_f64(k), static_cast<double>(k))) FAIL("int_to_f64 edge %d", k);
  const double vs[] = {0.0, -0.0, 0.5, -0.5, 0.9999999999999999, -0.9999999999999999, 1.0, -1.0, 2.5, -2.5, 3.5, -3.5, 1e9, -1e9};
  for (double v : vs)
    if (f64_trunc_to_int(v) != static_cast<int>(v)) FAIL("f64_trunc_to_int edge %.17g", v);
  return 0;
}

static int check_geometry(unsigned seed, int iters) {
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  for (int it = 0; it < iters; it++) {
    const double res_choices[] = {0.05, 0.02, 1.0, 0.1, 0.03};
    const double res = res_choices[rng() % 5];
    const int rows = 3 + (int)(rng() % 300), cols = 3 + (int)(rng() % 300);
    oracle_geom og;
    oracle_geom_init(&og, rows * res, cols * res, res, (U(rng) - 0.5) * 40, (U(rng) - 0.5) * 40);
    if (rng() % 2) {
      og.start0 = (int)(rng() % og.rows);
      og.start1 = (int)(rng() % og.cols);
    }
    GridDims d{og.rows, og.cols, og.res, og.len_x, og.len_y, 1.0 / og.res};
    RobotGeom g{og.pos_x, og.pos_y, og.start0, og.start1};
    for (int k = 0; k < 200; k++) {
      const double x = og.pos_x + (U(rng) - 0.5) * og.len_x * 1.6, y = og.pos_y + (U(rng) - 0.5) * og.len_y * 1.6;
      int r1 = -7, c1 = -7, r2 = -7, c2 = -7;
      const int ok1 = oracle_index_from_position(&og, x, y, &r1, &c1);
      const bool ok2 = grid_index(d, g, x, y, r2, c2);
      g_checks++;
      if ((ok1 != 0) != ok2 || (ok2 && (r1 != r2 || c1 != c2))) FAIL("index_from_position (%g,%g)", x, y);
      if (ok2) {
        double px1, py1, px2, py2;
        oracle_position_from_index(&og, r1, c1, &px1, &py1);
        position_from_index(r1, c1, d.len_x, d.len_y, g.pos_x, g.pos_y, d.res, d.rows, d.cols, g.start0, g.start1, px2,
                            py2);
        g_checks++;
        if (!same_bits(px1, px2) || !same_bits(py1, py2)) FAIL("position_from_index (%d,%d)", r1, c1);
      }
      /* submap */
      {
        const double L = (rng() % 2) ? 1.5 : 2.58;
        int tr, tc, sr, sc;
        double spx, spy, slx, sly;
        const int okA = oracle_submap_info(&og, x, y, L, L, &tr, &tc, &sr, &sc, &spx, &spy, &slx, &sly);
        SubmapInfo si;
        const bool okB = submap_info(d, g, x, y, L, L, si);
        g_checks++;
        /* the oracle's get_submap additionally requires utl+size <= size; fold it in */
        bool okA2 = okA != 0;
        if (okA2) {
          int utr = tr - og.start0, utc = tc - og.start1;
          if (og.start0 || og.start1) {
            wrap_index(utr, og.rows);
            wrap_index(utc, og.cols);
          } else {
            utr = tr;
            utc = tc;
          }
          if (utr + sr > og.rows || utc + sc > og.cols) okA2 = false;
        }
        if (okA2 != okB) FAIL("submap_info ok flag at (%g,%g): %d vs %d", x, y, (int)okA2, (int)okB);
        if (okB && (tr != si.tl_r || tc != si.tl_c || sr != si.size_r || sc != si.size_c || !same_bits(spx, si.pos_x) ||
                    !same_bits(spy, si.pos_y) || !same_bits(slx, si.len_x) || !same_bits(sly, si.len_y)))
          FAIL("submap_info values at (%g,%g)", x, y);
      }
    }
    /* lines: oracle stepping vs make_beam + closed form + rectangle clipping */
    std::vector<int> rc(2 * 70000);
    for (int k = 0; k < 60; k++) {
      const double sx = og.pos_x + (U(rng) - 0.5) * og.len_x * 1.5, sy = og.pos_y + (U(rng) - 0.5) * og.len_y * 1.5;
      const double ex = og.pos_x + (U(rng) - 0.5) * og.len_x * 1.5, ey = og.pos_y + (U(rng) - 0.5) * og.len_y * 1.5;
      const int n = oracle_line_cells(&og, sx, sy, ex, ey, rc.data(), 70000);
      const BeamSeg b = make_beam(d, g, sx, sy, ex, ey, 0);
      g_checks++;
      if ((n == 0) != (b.r0 < 0)) FAIL("line existence");
      int mr, mc;
      const int mok = oracle_index_from_position(&og, ex, ey, &mr, &mc);
      if ((mok != 0) != (b.mr >= 0) || (mok && (mr != b.mr || mc != b.mc))) FAIL("mark cell");
      if (n == 0) continue;
      if (rc[0] != b.r0 || rc[1] != b.c0 || rc[2 * (n - 1)] != b.r1 || rc[2 * (n - 1) + 1] != b.c1) FAIL("line ends");
      const LineForm f = line_form(b);
      if (f.den + 1 != n) FAIL("line length %d vs %d", f.den + 1, n);
      /* closed form at every t */
      const unsigned den = (unsigned)(f.den > 0 ? f.den : 1);
      for (int t = 0; t < n; t++) {
        const int q = (int)(((unsigned)(f.den >> 1) + (unsigned)t * (unsigned)f.add) / den);
        const int mj = f.m0 + f.sm * t, mn = f.n0 + f.sn * q;
        const int r = f.row_major ? mj : mn, c = f.row_major ? mn : mj;
        if (r != rc[2 * t] || c != rc[2 * t + 1]) FAIL("closed form at t=%d", t);
      }
      g_checks += n;
      /* rectangle clipping + the per-lane float-reciprocal start, as the kernel does it */
      for (int rr = 0; rr < 6; rr++) {
        int rlo = (int)(rng() % og.rows), rhi = (int)(rng() % og.rows), clo = (int)(rng() % og.cols),
            chi = (int)(rng() % og.cols);
        if (rlo > rhi) std::swap(rlo, rhi);
        if (clo > chi) std::swap(clo, chi);
        int t0 = 0, t1 = -1;
        const bool any = clip_line_to_rect(f, rlo, rhi, clo, chi, t0, t1);
        int e0 = -1, e1 = -2;
        for (int t = 0; t < n; t++) {
          const bool in = rc[2 * t] >= rlo && rc[2 * t] <= rhi && rc[2 * t + 1] >= clo && rc[2 * t + 1] <= chi;
          if (in) {
            if (e0 < 0) e0 = t;
            e1 = t;
          }
        }
        g_checks++;
        if (any != (e0 >= 0)) FAIL("clip existence");
        if (any && (t0 != e0 || t1 != e1)) FAIL("clip range [%d,%d] vs [%d,%d]", t0, t1, e0, e1);
        if (any) {
          /* every cell in [t0,t1] must be inside (convexity) and the lane formula must reproduce it */
          const unsigned x0 = (unsigned)(f.den >> 1) + (unsigned)t0 * (unsigned)f.add;
          const unsigned q0 = x0 / den;
          const int rem0 = (int)(x0 - q0 * den);
          const float rcp = 1.0f / (float)den;
          const int x32 = 32 * f.add, q32 = small_quotient(x32, rcp), r32 = x32 - q32 * (int)den;
          for (int lane = 0; lane < 32; lane++) {
            const int x = rem0 + lane * f.add;
            int q = small_quotient(x, rcp);
            if (q != x / (int)den) FAIL("small_quotient x=%d den=%u", x, den);
            int rem = x - q * (int)den;
            for (int k2 = lane; k2 < t1 - t0 + 1; k2 += 32) {
              const int t = t0 + k2;
              const int mj = f.m0 + f.sm * t, mn = f.n0 + f.sn * ((int)q0 + q);
              const int r = f.row_major ? mj : mn, c = f.row_major ? mn : mj;
              if (r != rc[2 * t] || c != rc[2 * t + 1]) FAIL("lane walk t=%d", t);
              if (r < rlo || r > rhi || c < clo || c > chi) FAIL("lane walk outside rect");
              rem += r32;
              q += q32;
              if (rem >= (int)den) {
                rem -= (int)den;
                q += 1;
              }
            }
          }
        }
      }
    }
  }
  return 0;
}

static int check_vfh_tables(const b200nav_vfh_params& p) {
  double arr[20] = {p.cell_size, (double)p.window_diameter, (double)p.sector_angle, p.safety_dist_0ms,
                    p.safety_dist_1ms, (double)p.max_speed, (double)p.max_speed_narrow_opening,
                    (double)p.max_speed_wide_opening, (double)p.max_acceleration, (double)p.min_turnrate,
                    (double)p.max_turnrate_0ms, (double)p.max_turnrate_1ms, p.min_turn_radius_safety_factor,
                    p.free_space_cutoff_0ms, p.obs_cutoff_0ms, p.free_space_cutoff_1ms, p.obs_cutoff_1ms,
                    p.weight_desired_dir, p.weight_current_dir, p.robot_radius};
  void* ref = vfhref_create(arr);
  VfhTables t;
  char err[256];
  if (vfh_build_tables(p, t, err, sizeof(err)) != 0) FAIL("vfh_build_tables: %s", err);
  const int W = p.window_diameter;
  if (vfhref_num_tables(ref) != t.c.num_tables || vfhref_hist_size(ref) != t.c.hist_size) FAIL("table counts");
  std::vector<float> d(W * W), s(W * W), b(W * W);
  vfhref_get_cell_tables(ref, d.data(), s.data(), b.data());
  for (int i = 0; i < W * W; i++) {
    g_checks++;
    if (std::memcmp(&d[i], &t.dir_xy[i], 4) || std::memcmp(&s[i], &t.dist_xy[i], 4) || std::memcmp(&b[i], &t.base_xy[i], 4))
      FAIL("cell table W=%d i=%d: dir %g/%g dist %g/%g base %g/%g", W, i, d[i], t.dir_xy[i], s[i], t.dist_xy[i], b[i],
           t.base_xy[i]);
  }
  std::vector<unsigned> m((size_t)W * W * t.c.nwords);
  for (int tab = 0; tab < t.c.num_tables; tab++) {
    if (!vfhref_get_sector_masks(ref, tab, m.data(), t.c.nwords)) FAIL("reference sector list not ascending");
    g_checks++;
    if (std::memcmp(m.data(), t.masks_xy.data() + (size_t)tab * W * W * t.c.nwords, m.size() * 4)) FAIL("sector masks W=%d table %d", W, tab);
  }
  std::vector<int> mtr(p.max_speed + 1);
  vfhref_get_min_turning_radius(ref, mtr.data(), p.max_speed + 1);
  for (int i = 0; i <= p.max_speed; i++)
    if (mtr[i] != t.min_turning_radius[i]) FAIL("min turning radius %d", i);
  vfhref_destroy(ref);
  return 0;
}

/* (4) dda_init / dda_at (the fan walk's fixed-point stepping) against the integer Bresenham recurrence:
 *     q(t) = floor(((den >> 1) + t * add) / den) for every step of every (den, add) tested; also started mid-line
 *     the way the tile kernel does (state taken at t0, then frac += S with carry). */
static int check_dda_pair(unsigned den, unsigned add) {
  unsigned S, B;
  bool diag;
  b200nav::dda_init(add, den, S, B, diag);
  const unsigned h = den >> 1;
  /* closed form at every t */
  for (unsigned t = 0; t <= den; t++) {
    const unsigned want = (unsigned)(((unsigned long long)h + (unsigned long long)t * add) / den);
    const unsigned got = diag ? t : (unsigned)(b200nav::dda_at(S, B, t) >> 32);
    g_checks++;
    if (want != got) {
      printf("dda closed form den=%u add=%u t=%u: want %u got %u\n", den, add, t, want, got);
      return 1;
    }
  }
  /* incremental from a start in the middle */
  const unsigned t0 = den / 3;
  unsigned frac = (unsigned)b200nav::dda_at(S, B, t0);
  unsigned q = diag ? t0 : (unsigned)(b200nav::dda_at(S, B, t0) >> 32);
  for (unsigned t = t0; t < den; t++) {
    const unsigned nf = frac + S;
    q += diag ? 1u : (nf < frac ? 1u : 0u);
    frac = nf;
    const unsigned want = (unsigned)(((unsigned long long)h + (unsigned long long)(t + 1) * add) / den);
    g_checks++;
    if (want != q) {
      printf("dda walk den=%u add=%u t=%u: want %u got %u\n", den, add, t + 1, want, q);
      return 1;
    }
  }
  return 0;
}

static int check_dda(unsigned seed, int iters) {
  const unsigned full = iters > 100 ? 900u : 400u; /* every add for every den up to here */
  for (unsigned den = 1; den <= full; den++)
    for (unsigned add = 0; add <= den; add++)
      if (check_dda_pair(den, add)) return 1;
  /* extremes of every 5th longer line, and random slopes up to the largest grid (32767 cells) */
  for (unsigned den = full + 1; den <= 32767u; den += 5) {
    const unsigned adds[] = {0u, 1u, 2u, den / 2 - 1, den / 2, den / 2 + 1, den - 2, den - 1, den};
    if (den % 97 == 0)
      for (unsigned a : adds)
        if (check_dda_pair(den, a)) return 1;
  }
  std::mt19937 rng(seed);
  for (int i = 0; i < 300 * (iters > 100 ? 5 : 1); i++) {
    const unsigned den = 1u + rng() % 32767u, add = rng() % (den + 1u);
    if (check_dda_pair(den, add)) return 1;
  }
  for (unsigned den : {32767u, 32766u, 65535u, 65534u, 40001u})
    for (unsigned add : {1u, den / 3, den - 1, den})
      if (check_dda_pair(den, add)) return 1;
  return 0;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 40;
  if (check_dda(4242u, iters)) return 1;
  if (check_f64_helpers(777u, 200000LL * (iters > 100 ? 10 : 1))) return 1;
  if (check_geometry(12345u, iters)) return 1;
  b200nav_vfh_params p;
  memset(&p, 0, sizeof(p));
  p.cell_size = 100; p.window_diameter = 30; p.sector_angle = 5; p.safety_dist_0ms = 10; p.safety_dist_1ms = 50;
  p.max_speed = 200; p.max_speed_narrow_opening = 200; p.max_speed_wide_opening = 300; p.max_acceleration = 200;
  p.min_turnrate = 40; p.max_turnrate_0ms = 40; p.max_turnrate_1ms = 40; p.min_turn_radius_safety_factor = 1.0;
  p.free_space_cutoff_0ms = 2e6; p.obs_cutoff_0ms = 4e6; p.free_space_cutoff_1ms = 2e6; p.obs_cutoff_1ms = 4e6;
  p.weight_desired_dir = 10; p.weight_current_dir = 1; p.robot_radius = 178; p.submap_length = 1.5; p.occupied_threshold = 3;
  if (check_vfh_tables(p)) return 1;
  p.window_diameter = 33;
  if (check_vfh_tables(p)) return 1;
  p.window_diameter = 129; p.cell_size = 20;
  if (check_vfh_tables(p)) return 1;
  p.window_diameter = 60; p.cell_size = 100; p.safety_dist_1ms = 10; p.robot_radius = 300; p.max_turnrate_1ms = 20; p.max_speed = 500;
  if (check_vfh_tables(p)) return 1;
  p.sector_angle = 2; p.window_diameter = 41; p.safety_dist_1ms = 200;
  if (check_vfh_tables(p)) return 1;
  printf("OK %lld checks\n", g_checks);
  return 0;
}
