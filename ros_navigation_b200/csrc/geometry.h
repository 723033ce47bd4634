/*
 * geometry.h -- fp64 grid_map geometry used on the device by the HIMM and VFH+ kernels.
 *
 * Behavioural spec (not code) taken from the reference:
 *   grid_map_core/src/GridMapMath.cpp:30-34,55-63,70-100,115-159,202-239,246-296
 *   grid_map_core/src/iterators/LineIterator.cpp:92-150
 * Bit-exactness contract: every expression keeps the reference's association order and is compiled with
 * FMA contraction disabled (nvcc --fmad=false; g++ -ffp-contract=off for the host unit check), because the
 * reference is built by x86-64 g++ which emits no fused multiply-adds.  IEEE add/sub/mul/div/sqrt are
 * correctly rounded on both sides, so results agree to the bit.
 */
#ifndef B200NAV_GEOMETRY_H
#define B200NAV_GEOMETRY_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200nav {

/* Geometry shared by all robots of a grid. */
struct GridDims {
  int rows, cols;
  double res;
  double len_x, len_y;
  double rres; /* RN(1/res), for f64_div_by */
};

/* Per-robot part: map centre and circular-buffer start index. */
struct RobotGeom {
  double pos_x, pos_y;
  int start0, start1;
};

#define B200NAV_DBL_EPSILON 2.2204460492503131e-16

/* ---------------------------------------------------------------------------------------------------------------
 * fp64 helpers that avoid the 64-bit XU operations (MUFU.RCP64H, F2I.F64, I2F.F64, F2F.F64.F32).  On B200 those
 * issue at ~1/64 of the warp rate (ncu: the prep kernel sat at 75 % XU-pipe utilisation with < 3 % of its
 * instructions being such operations).  Every helper returns bit-for-bit what the plain C++ expression returns.
 * ------------------------------------------------------------------------------------------------------------- */
B200_HD unsigned long long f64_bits(double v) {
#if defined(__CUDA_ARCH__)
  return (unsigned long long)__double_as_longlong(v);
#else
  unsigned long long b;
  memcpy(&b, &v, 8);
  return b;
#endif
}
B200_HD double f64_from_bits(unsigned long long b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double v;
  memcpy(&v, &b, 8);
  return v;
#endif
}

/* (double)k for any int32 k: 2^52 + 2^31 + k is exact in the low mantissa bits. */
B200_HD double int_to_f64(int k) {
  return f64_from_bits(0x4330000000000000ull | (unsigned long long)((unsigned)k ^ 0x80000000u)) - 4503601774854144.0;
}

/* (int)v (C truncation toward zero) for |v| < 2^31. */
B200_HD int f64_trunc_to_int(double v) {
  const double t = v + 6755399441055744.0; /* 2^52 + 2^51: round to nearest integer into the low bits */
  int n = (int)(unsigned)(f64_bits(t) & 0xffffffffull);
  const double nd = t - 6755399441055744.0;
  if (v >= 0.0) {
    if (nd > v) n -= 1;
  } else {
    if (nd < v) n += 1;
  }
  return n;
}

/* (double)f for a float f. */
B200_HD double f32_to_f64(float f) {
#if defined(__CUDA_ARCH__)
  const unsigned u = __float_as_uint(f);
#else
  unsigned u;
  memcpy(&u, &f, 4);
#endif
  const unsigned e = (u >> 23) & 0xffu;
  if (e != 0u && e != 255u) /* normal: re-bias the exponent, widen the mantissa */
    return f64_from_bits(((unsigned long long)(u & 0x80000000u) << 32) | ((unsigned long long)(e + 896u) << 52) |
                         ((unsigned long long)(u & 0x7fffffu) << 29));
  return (double)f; /* zero, subnormal, inf, NaN */
}

/* a / b, correctly rounded, given y = RN(1/b) (precomputed once with a true division): Markstein iteration
 * q1 = q0 + (a - q0*b)*y, then an exact-remainder check against the neighbouring double, so the result is the
 * IEEE quotient whatever b is.  Valid for finite a, b with the quotient in the normal range (all uses here:
 * |a| < 1e6 m, b = the grid resolution). */
B200_HD double f64_div_by(double a, double b, double y) {
  const double q0 = a * y;
  const double r0 = fma(-q0, b, a);
  const double q1 = fma(r0, y, q0);
  const double r1 = fma(-q1, b, a);
  if (r1 == 0.0) return q1;
  /* neighbour of q1 on the side of the true quotient */
  const bool up = (r1 > 0.0) == (b > 0.0);
  unsigned long long qb = f64_bits(q1);
  if (q1 == 0.0) return q1;
  const bool away = (q1 > 0.0) == up; /* moving away from zero increments the bit pattern */
  qb = away ? qb + 1ull : qb - 1ull;
  const double qn = f64_from_bits(qb);
  const double r2 = fma(-qn, b, a);
  return (fabs(r2) < fabs(r1)) ? qn : q1;
}

/* One Bresenham line in buffer-index space + the cell to mark + the line's fixed-point DDA constants (dda_init:
 * computed once per beam by the binning kernel instead of once per (beam, tile) by the tile kernel).  32 bytes. */
struct alignas(16) BeamSeg {
  int r0, c0, r1, c1;   /* inclusive endpoints; r0 < 0: no line                    */
  int mr, mc;           /* cell that receives the +30 mark; mr < 0: no mark        */
  unsigned S, B;        /* dda_init(add, den) of the line (0, 0 for 45-degree lines and when there is no line) */
};

B200_HD void wrap_index(int& index, int buffer_size) {
  if (index < 0) index += ((-index / buffer_size) + 1) * buffer_size;
  index = index % buffer_size;
}

/* inside test: 0 <= -((p - c) - L/2) < L on both axes. */
B200_HD bool within_map(double px, double py, double lx, double ly, double mx, double my) {
  const double tx = -((px - mx) - 0.5 * lx);
  const double ty = -((py - my) - 0.5 * ly);
  return tx >= 0.0 && ty >= 0.0 && tx < lx && ty < ly;
}

/* position -> buffer index: -trunc(((p - L/2) - c) / res), then + startIndex mod size. */
B200_HD bool index_from_position(double px, double py, double lx, double ly, double mx, double my, double res,
                                 double rres, int rows, int cols, int s0, int s1, int& r, int& c) {
  if (!within_map(px, py, lx, ly, mx, my)) return false;
  const double vx = f64_div_by((px - 0.5 * lx) - mx, res, rres); /* == ((px - 0.5*lx) - mx) / res */
  const double vy = f64_div_by((py - 0.5 * ly) - my, res, rres);
  int i0 = -f64_trunc_to_int(vx);                                 /* == -static_cast<int>(vx)       */
  int i1 = -f64_trunc_to_int(vy);
  if ((s0 | s1) != 0) {
    i0 += s0;
    i1 += s1;
    wrap_index(i0, rows);
    wrap_index(i1, cols);
  }
  r = i0;
  c = i1;
  return true;
}

B200_HD bool grid_index(const GridDims& d, const RobotGeom& g, double px, double py, int& r, int& c) {
  return index_from_position(px, py, d.len_x, d.len_y, g.pos_x, g.pos_y, d.res, d.rres, d.rows, d.cols, g.start0,
                             g.start1, r, c);
}

/* buffer index -> cell centre: (c + (L/2 - res/2)) + res * (-(unwrapped index)). */
B200_HD void position_from_index(int r, int c, double lx, double ly, double mx, double my, double res, int rows,
                                 int cols, int s0, int s1, double& px, double& py) {
  int u0 = r, u1 = c;
  if ((s0 | s1) != 0) {
    u0 -= s0;
    u1 -= s1;
    wrap_index(u0, rows);
    wrap_index(u1, cols);
  }
  px = (mx + (0.5 * lx - 0.5 * res)) + res * int_to_f64(-u0);
  py = (my + (0.5 * ly - 0.5 * res)) + res * int_to_f64(-u1);
}

B200_HD void limit_position_to_range(double& px, double& py, double lx, double ly, double mx, double my) {
  const double p[2] = {px, py};
  double s[2] = {(px - mx) + 0.5 * lx, (py - my) + 0.5 * ly};
  const double len[2] = {lx, ly};
#pragma unroll
  for (int i = 0; i < 2; i++) {
    double epsilon = 10.0 * B200NAV_DBL_EPSILON;
    if (fabs(p[i]) > 1.0) epsilon *= fabs(p[i]);
    if (s[i] <= 0)
      s[i] = epsilon;
    else if (s[i] >= len[i])
      s[i] = len[i] - epsilon;
  }
  px = (s[0] + mx) - 0.5 * lx;
  py = (s[1] + my) - 0.5 * ly;
}

/* Pull a ray end into the map by stepping (res - eps) along the ray; iterative on purpose (the rounding of
 * the running sum is part of the reference result).  Two guards bound the loop for hostile input, where the
 * reference (LineIterator.cpp:96-103) would spin: a step that no longer moves the point (coordinates beyond 2^53
 * steps: the reference never terminates) and more than B200NAV_CLIP_MAX_STEPS steps (a start hundreds of kilometres
 * from the map: the reference would get there eventually) both give "no line" (defined answers). */
#define B200NAV_CLIP_MAX_STEPS (1 << 22)
B200_HD bool clip_into_map(const GridDims& d, const RobotGeom& g, double sx, double sy, double ex, double ey, int& r,
                           int& c) {
  /* Common case first: the end is inside the map and the direction (sqrt + two divisions) is never needed. */
  if (grid_index(d, g, sx, sy, r, c)) return true;
  double nx = sx, ny = sy;
  double dx = ex - sx, dy = ey - sy;
  const double z = dx * dx + dy * dy;
  if (z > 0.0) {
    const double n = sqrt(z);
    dx = dx / n;
    dy = dy / n;
  }
  const double step = d.res - B200NAV_DBL_EPSILON;
  int steps = 0;
  do {
    const double px = nx, py = ny;
    nx += step * dx;
    ny += step * dy;
    if ((nx == px && ny == py) || ++steps > B200NAV_CLIP_MAX_STEPS) return false; /* no progress / too far away */
    const double qx = ex - nx, qy = ey - ny;
    if (!(sqrt(qx * qx + qy * qy) >= step)) return false; /* also ends the loop on NaN */
  } while (!grid_index(d, g, nx, ny, r, c));
  return true;
}

/* RangeSample -> BeamSeg (LineIterator ctor + the mark cell of lineOnMap). */
B200_HD BeamSeg make_beam(const GridDims& d, const RobotGeom& g, double sx, double sy, double ex, double ey,
                          int clear_end) {
  BeamSeg b;
  b.r0 = b.c0 = b.r1 = b.c1 = -1;
  b.mr = b.mc = -1;
  b.S = b.B = 0u; /* filled in by the caller that needs them (himm_prep_kernel) */
  /* Non-finite coordinates would make the reference's clip loop spin forever; defined as "no line".  The mark
   * only depends on the end point (map_updater.h:44-49). */
  const bool finite = isfinite(sx) && isfinite(sy) && isfinite(ex) && isfinite(ey);
  int r0, c0, r1, c1;
  if (finite && clip_into_map(d, g, sx, sy, ex, ey, r0, c0) && clip_into_map(d, g, ex, ey, sx, sy, r1, c1)) {
    b.r0 = r0;
    b.c0 = c0;
    b.r1 = r1;
    b.c1 = c1;
  }
  if (!clear_end) {
    /* mark cell = index(end).  When the line exists its last cell already is index(end) if the end is inside
     * the map (the clip of the end returns immediately); only re-derive it otherwise. */
    int mr, mc;
    if (b.r0 >= 0 && within_map(ex, ey, d.len_x, d.len_y, g.pos_x, g.pos_y)) {
      b.mr = b.r1;
      b.mc = b.c1;
    } else if (grid_index(d, g, ex, ey, mr, mc)) {
      b.mr = mr;
      b.mc = mc;
    }
  }
  return b;
}

/* Submap window of getSubmapInformation: buffer index of the top-left cell, size, and the submap's own
 * geometry (position = centre, length).  Returns false when the window cannot be formed. */
struct SubmapInfo {
  int tl_r, tl_c;     /* buffer index of the top-left cell            */
  int utl_r, utl_c;   /* the same, unwrapped                          */
  int size_r, size_c;
  double pos_x, pos_y, len_x, len_y;
};

B200_HD bool submap_info(const GridDims& d, const RobotGeom& g, double cx, double cy, double lx, double ly,
                         SubmapInfo& o) {
  double tlx = cx - (-0.5 * lx);
  double tly = cy - (-0.5 * ly);
  limit_position_to_range(tlx, tly, d.len_x, d.len_y, g.pos_x, g.pos_y);
  int tr, tc;
  if (!grid_index(d, g, tlx, tly, tr, tc)) return false;
  int utr = tr, utc = tc;
  if ((g.start0 | g.start1) != 0) {
    utr -= g.start0;
    utc -= g.start1;
    wrap_index(utr, d.rows);
    wrap_index(utc, d.cols);
  }
  double brx = cx + (-0.5 * lx);
  double bry = cy + (-0.5 * ly);
  limit_position_to_range(brx, bry, d.len_x, d.len_y, g.pos_x, g.pos_y);
  int br, bc;
  if (!grid_index(d, g, brx, bry, br, bc)) return false;
  if ((g.start0 | g.start1) != 0) {
    br -= g.start0;
    bc -= g.start1;
    wrap_index(br, d.rows);
    wrap_index(bc, d.cols);
  }
  double cxp, cyp;
  position_from_index(tr, tc, d.len_x, d.len_y, g.pos_x, g.pos_y, d.res, d.rows, d.cols, g.start0, g.start1, cxp,
                      cyp);
  const double half = 0.5 * d.res;
  cxp -= -half;
  cyp -= -half;
  const int sr = br - utr + 1, sc = bc - utc + 1;
  if (sr <= 0 || sc <= 0) return false;
  const double slx = int_to_f64(sr) * d.res, sly = int_to_f64(sc) * d.res;
  const double spx = cxp - 0.5 * slx, spy = cyp - 0.5 * sly;
  if (!within_map(cx, cy, slx, sly, spx, spy)) return false;
  if (utr + sr > d.rows || utc + sc > d.cols) return false;
  o.tl_r = tr;
  o.tl_c = tc;
  o.utl_r = utr;
  o.utl_c = utc;
  o.size_r = sr;
  o.size_c = sc;
  o.pos_x = spx;
  o.pos_y = spy;
  o.len_x = slx;
  o.len_y = sly;
  return true;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Bresenham closed form.  For a line (r0,c0)->(r1,c1) the reference visits n = max(|dr|,|dc|)+1 cells; at step t
 * the driving ("major") coordinate is m0 + sm*t and the other is n0 + sn*floor((den/2 + t*add)/den)
 * (LineIterator.cpp:60-70,106-150; ties |dr| >= |dc| drive along rows).
 * ------------------------------------------------------------------------------------------------------------- */
struct LineForm {
  int den, add;      /* denominator_, numeratorAdd_                 */
  int m0, n0;        /* start major / minor coordinate              */
  int sm, sn;        /* +-1                                         */
  bool row_major;    /* true: rows drive                            */
};

B200_HD LineForm line_form(const BeamSeg& b) {
  LineForm f;
  const int dr = b.r1 - b.r0, dc = b.c1 - b.c0;
  const int adr = (dr < 0 ? -dr : dr), adc = (dc < 0 ? -dc : dc);
  f.row_major = adr >= adc;
  f.den = f.row_major ? adr : adc;
  f.add = f.row_major ? adc : adr;
  f.m0 = f.row_major ? b.r0 : b.c0;
  f.n0 = f.row_major ? b.c0 : b.r0;
  const int sr = (b.r1 >= b.r0) ? 1 : -1, sc = (b.c1 >= b.c0) ? 1 : -1;
  f.sm = f.row_major ? sr : sc;
  f.sn = f.row_major ? sc : sr;
  return f;
}

/* Steps t in [t0,t1] of the line that fall inside the inclusive rectangle; returns false if none. */
B200_HD bool clip_line_to_rect(const LineForm& f, int rlo, int rhi, int clo, int chi, int& t0,
                                                  int& t1) {
  const int mlo = f.row_major ? rlo : clo, mhi = f.row_major ? rhi : chi;
  const int nlo = f.row_major ? clo : rlo, nhi = f.row_major ? chi : rhi;
  int ta, tb;
  if (f.sm > 0) {
    ta = mlo - f.m0;
    tb = mhi - f.m0;
  } else {
    ta = f.m0 - mhi;
    tb = f.m0 - mlo;
  }
  t0 = ta > 0 ? ta : 0;
  t1 = tb < f.den ? tb : f.den;
  int qlo, qhi;
  if (f.sn > 0) {
    qlo = nlo - f.n0;
    qhi = nhi - f.n0;
  } else {
    qlo = f.n0 - nhi;
    qhi = f.n0 - nlo;
  }
  if (qhi < 0 || qlo > f.add || t0 > t1) return false;
  const int num0 = f.den >> 1;
  if (qlo > 0) { /* den/2 + t*add >= qlo*den  (add >= qlo > 0) */
    const unsigned num = (unsigned)(qlo * f.den - num0);
    const int lo = (int)((num + (unsigned)f.add - 1u) / (unsigned)f.add);
    if (lo > t0) t0 = lo;
  }
  if (qhi < f.add) { /* den/2 + t*add <= (qhi+1)*den - 1  (add > qhi >= 0) */
    const unsigned num = (unsigned)((qhi + 1) * f.den - 1 - num0);
    const int hi = (int)(num / (unsigned)f.add);
    if (hi < t1) t1 = hi;
  }
  return t0 <= t1;
}


/* floor(x / den) for 0 <= x < 33*den, den <= 32767, through a correctly rounded float reciprocal: the +0.5 keeps the
 * true quotient at least 0.5/den (>= 1.5e-5) away from an integer while the float error stays below 5e-6. */
B200_HD int small_quotient(int x, float rcp_den) { return (int)(((float)x + 0.5f) * rcp_den); }

/* Exact fixed-point form of the Bresenham recurrence of LineForm (geometry.h): step t of a line visits minor index
 *   q(t) = floor(((den >> 1) + t * add) / den),   0 <= t <= den,  add <= den.
 * With S = ceil(add * 2^32 / den) and B = ceil((den >> 1) * 2^32 / den), X(t) = B + t * S (64 bit) satisfies
 *   X(t) / 2^32 = exact + e,  0 <= e < (t + 1) / 2^32 <= (den + 1) / 2^32,
 * and the exact value's fractional part is a multiple of 1/den, at most 1 - 1/den.  So floor(X(t) / 2^32) = q(t)
 * whenever (den + 1) / 2^32 < 1 / den, i.e. for every den <= 65535 (grids have at most 32767 cells per side).
 * The walk then is: frac += S; carry -> one minor step.  add == den (45 degrees) would need S = 2^32: `diag` folds
 * the minor step into the major one instead.  Verified exhaustively against the integer recurrence in
 * tests/cpp/host_checks.cpp. */
B200_HD void dda_init(unsigned add, unsigned den, unsigned& S, unsigned& B, bool& diag) {
  diag = add >= den;
  /* S = ceil(add * 2^32 / den) by two 16-bit long-division steps (add, den < 2^16) */
  const unsigned n1 = add << 16, hi = n1 / den, r1 = n1 - hi * den;
  const unsigned n2 = r1 << 16, lo = n2 / den, r2 = n2 - lo * den;
  S = diag ? 0u : ((hi << 16) + lo + (r2 != 0u ? 1u : 0u));
  /* B = ceil((den >> 1) * 2^32 / den): 2^31 for even den, 2^31 - floor(2^31 / den) for odd den */
  B = diag ? 0u : 0x80000000u - ((den & 1u) ? 0x80000000u / den : 0u);
}
B200_HD unsigned long long dda_at(unsigned S, unsigned B, unsigned t) {
  return (unsigned long long)B + (unsigned long long)t * (unsigned long long)S;
}

}  // namespace b200nav
#endif
