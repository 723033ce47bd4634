/*
 * himm_oracle.cpp -- CPU ORACLE (test infrastructure, not product; see oracle_api.h).
 *
 * Restates, over a plain column-major float array, the slice of grid_map_core and move_control that
 * the HIMM update and the grid-window -> pseudo-scan stage execute.  Plain C++17, no Eigen / ROS /
 * boost.  Floating point follows the reference expression by expression (association order kept,
 * compiled with -ffp-contract=off, no -ffast-math) so the results are what the reference would
 * produce on this toolchain.  Citations are relative to the reference root.
 */
#include "oracle_api.h"

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct V2 {
  double x, y;
};

/* ---- grid_map_core/src/GridMapMath.cpp -------------------------------------------------------- */

/* GridMapMath.cpp:202-206 (scalar) and :194-200 (vector). */
inline void map_index_within_range(int& index, int buffer_size) {
  if (index < 0) index += ((-index / buffer_size) + 1) * buffer_size;
  index = index % buffer_size;
}

/* GridMapMath.cpp:70-81: index (unwrapped) -> buffer index. */
inline void buffer_index_from_index(const oracle_geom* g, int& i0, int& i1) {
  if (g->start0 == 0 && g->start1 == 0) return;
  i0 += g->start0;
  i1 += g->start1;
  map_index_within_range(i0, g->rows);
  map_index_within_range(i1, g->cols);
}

/* GridMapMath.cpp:484-492: buffer index -> unwrapped index. */
inline void index_from_buffer_index(const oracle_geom* g, int& i0, int& i1) {
  if (g->start0 == 0 && g->start1 == 0) return;
  i0 -= g->start0;
  i1 -= g->start1;
  map_index_within_range(i0, g->rows);
  map_index_within_range(i1, g->cols);
}

/* GridMapMath.cpp:146-159 with getVectorToOrigin (:30-34) and the -I transform (:55-63). */
inline bool within_map(double px, double py, double lx, double ly, double mx, double my) {
  const double ox = 0.5 * lx, oy = 0.5 * ly;
  /* positionTransformed = (-I) * (position - mapPosition - offset); the 2x2 product adds a 0*other term. */
  const double tx = -1.0 * ((px - mx) - ox) + 0.0 * ((py - my) - oy);
  const double ty = 0.0 * ((px - mx) - ox) + -1.0 * ((py - my) - oy);
  return tx >= 0.0 && ty >= 0.0 && tx < lx && ty < ly;
}

/* GridMapMath.cpp:130-144 for an arbitrary geometry (also used for the submap geometry at :293). */
inline bool index_from_position(double px, double py, double lx, double ly, double mx, double my, double res,
                                int rows, int cols, int s0, int s1, int& r, int& c) {
  if (!within_map(px, py, lx, ly, mx, my)) return false;
  const double ox = 0.5 * lx, oy = 0.5 * ly;
  const double vx = ((px - ox) - mx) / res;
  const double vy = ((py - oy) - my) / res;
  /* getIndexFromIndexVector (:93-100): (-I^T) * indexVector.cast<int>() : cast truncates, then negate. */
  int i0 = -static_cast<int>(vx);
  int i1 = -static_cast<int>(vy);
  if (!(s0 == 0 && s1 == 0)) {
    i0 += s0;
    i1 += s1;
    map_index_within_range(i0, rows);
    map_index_within_range(i1, cols);
  }
  r = i0;
  c = i1;
  return true;
}

inline bool geom_index(const oracle_geom* g, double px, double py, int& r, int& c) {
  return index_from_position(px, py, g->len_x, g->len_y, g->pos_x, g->pos_y, g->res, g->rows, g->cols, g->start0,
                             g->start1, r, c);
}

/* GridMapMath.cpp:115-128 (+ :36-53 getVectorToFirstCell, :83-91 getIndexVectorFromIndex). */
inline bool position_from_index(int r, int c, double lx, double ly, double mx, double my, double res, int rows,
                                int cols, int s0, int s1, double& px, double& py) {
  if (!(r >= 0 && c >= 0 && r < rows && c < cols)) return false;
  const double offx = 0.5 * lx - 0.5 * res;
  const double offy = 0.5 * ly - 0.5 * res;
  int u0 = r, u1 = c;
  if (!(s0 == 0 && s1 == 0)) {
    u0 -= s0;
    u1 -= s1;
    map_index_within_range(u0, rows);
    map_index_within_range(u1, cols);
  }
  /* (-I * unwrapped).cast<double>() */
  const double ivx = static_cast<double>(-u0);
  const double ivy = static_cast<double>(-u1);
  px = (mx + offx) + res * ivx;
  py = (my + offy) + res * ivy;
  return true;
}

/* GridMapMath.cpp:208-231 limitPositionToRange. */
inline void limit_position_to_range(double& px, double& py, double lx, double ly, double mx, double my) {
  const double ox = 0.5 * lx, oy = 0.5 * ly;
  double p[2] = {px, py};
  double s[2] = {(px - mx) + ox, (py - my) + oy};
  const double len[2] = {lx, ly};
  for (int i = 0; i < 2; i++) {
    double epsilon = 10.0 * std::numeric_limits<double>::epsilon();
    if (std::fabs(p[i]) > 1.0) epsilon *= std::fabs(p[i]);
    if (s[i] <= 0) {
      s[i] = epsilon;
      continue;
    }
    if (s[i] >= len[i]) {
      s[i] = len[i] - epsilon;
      continue;
    }
  }
  px = (s[0] + mx) - ox;
  py = (s[1] + my) - oy;
}

/* ---- grid_map_core/src/iterators/LineIterator.cpp --------------------------------------------- */

/* LineIterator.cpp:92-104 getIndexLimitedToMapRange. */
inline bool index_limited_to_map_range(const oracle_geom* g, V2 start, V2 end, int& r, int& c) {
  V2 ns = start;
  /* (end - start).normalized(): v / sqrt(x*x + y*y) (Eigen returns v unchanged for the zero vector). */
  V2 d = {end.x - start.x, end.y - start.y};
  const double z = d.x * d.x + d.y * d.y;
  if (z > 0.0) {
    const double n = std::sqrt(z);
    d.x = d.x / n;
    d.y = d.y / n;
  }
  const double step = g->res - std::numeric_limits<double>::epsilon();
  /* Defined answers where the reference's loop would not terminate (or only after millions of steps): a step that
   * no longer moves the point, or more than 2^22 steps, mean "the ray never enters the map". */
  int steps = 0;
  while (!geom_index(g, ns.x, ns.y, r, c)) {
    const V2 before = ns;
    ns.x += step * d.x;
    ns.y += step * d.y;
    if ((ns.x == before.x && ns.y == before.y) || ++steps > (1 << 22)) return false;
    const double ex = end.x - ns.x, ey = end.y - ns.y;
    if (!(std::sqrt(ex * ex + ey * ey) >= step)) return false;
  }
  return true;
}

struct Line {
  int r, c;            /* index_   */
  int inc1r, inc1c;    /* increment1_ */
  int inc2r, inc2c;    /* increment2_ */
  int den, num, add;   /* denominator_, numerator_, numeratorAdd_ */
  int n;               /* nCells_ (0: no line; reference leaves the iterator uninitialised, H4b) */
};

/* LineIterator.cpp:16-23 + :106-150 initializeIterationParameters. */
inline Line make_line(const oracle_geom* g, V2 start, V2 end) {
  Line L;
  std::memset(&L, 0, sizeof(L));
  /* Non-finite coordinates make the reference's clip loop spin forever; defined here as "no line". */
  if (!(std::isfinite(start.x) && std::isfinite(start.y) && std::isfinite(end.x) && std::isfinite(end.y))) return L;
  int sr, sc, er, ec;
  if (!(index_limited_to_map_range(g, start, end, sr, sc) && index_limited_to_map_range(g, end, start, er, ec)))
    return L;
  L.r = sr;
  L.c = sc;
  const int dr = std::abs(er - sr), dc = std::abs(ec - sc);
  L.inc1r = L.inc2r = (er >= sr) ? 1 : -1;
  L.inc1c = L.inc2c = (ec >= sc) ? 1 : -1;
  if (dr >= dc) {
    L.inc1r = 0;
    L.inc2c = 0;
    L.den = dr;
    L.num = dr / 2;
    L.add = dc;
    L.n = dr + 1;
  } else {
    L.inc2r = 0;
    L.inc1c = 0;
    L.den = dc;
    L.num = dc / 2;
    L.add = dr;
    L.n = dc + 1;
  }
  return L;
}

/* LineIterator.cpp:60-70 operator++. */
inline void line_step(Line& L) {
  L.num += L.add;
  if (L.num >= L.den) {
    L.num -= L.den;
    L.r += L.inc1r;
    L.c += L.inc1c;
  }
  L.r += L.inc2r;
  L.c += L.inc2c;
}

/* ---- move_control/include/move_control/map_updater.h ------------------------------------------ */

/* map_updater.h:61-71 clearCell. */
inline void clear_cell(float& v) {
  if (v <= 0 || std::isnan(v))
    v = 0.0;
  else
    v = v - 10.0;
  if (v < 0.0) v = 0.0;
}

/* map_updater.h:52-59 markCell. */
inline void mark_cell(float& v) {
  if (v <= 0 || std::isnan(v))
    v = 30.0;
  else if (v <= 150.0)
    v = v + 30.0;
}

/* map_updater.h:73-78 touch. */
inline void touch(double* bbox, double x, double y) {
  bbox[0] = std::min(bbox[0], x);
  bbox[1] = std::min(bbox[1], y);
  bbox[2] = std::max(bbox[2], x);
  bbox[3] = std::max(bbox[3], y);
}

struct HoistedLayer {
  float* data;
  int rows;
  inline float& at(int r, int c) { return data[static_cast<size_t>(c) * rows + r]; }
};

/* Reference cost model: GridMap::operator[] -> unordered_map<string, Matrix>::at per cell. */
struct KeyedLayers {
  std::unordered_map<std::string, HoistedLayer> data;
  std::string type_name;
  inline float& at(int r, int c) { return data.at(type_name).at(r, c); }
};

template <class LayerT>
long long himm_update_impl(const oracle_geom* g, LayerT& layer, const oracle_sample* s, int n, double* bbox) {
  long long visits = 0;
  /* laser_map_updater.cpp:15-20: samples strictly in buffer order. */
  for (int k = 0; k < n; k++) {
    const V2 a = {s[k].sx, s[k].sy}, b = {s[k].ex, s[k].ey};
    /* map_updater.h:38-42: clear every cell of the line, end cell included. */
    Line L = make_line(g, a, b);
    for (int i = 0; i < L.n; i++) {
      clear_cell(layer.at(L.r, L.c));
      line_step(L);
      visits++;
    }
    /* map_updater.h:44-49: mark the (unclipped) end cell. */
    if (!s[k].clear_end) {
      int er, ec;
      if (geom_index(g, b.x, b.y, er, ec)) mark_cell(layer.at(er, ec));
    }
    if (bbox) {
      touch(bbox, a.x, a.y);
      touch(bbox, b.x, b.y);
    }
  }
  return visits;
}

}  // namespace

extern "C" {

void oracle_geom_init(oracle_geom* g, double len_x, double len_y, double res, double pos_x, double pos_y) {
  /* GridMap.cpp:51-70 */
  g->rows = static_cast<int>(round(len_x / res));
  g->cols = static_cast<int>(round(len_y / res));
  g->res = res;
  g->len_x = static_cast<double>(g->rows) * res;
  g->len_y = static_cast<double>(g->cols) * res;
  g->pos_x = pos_x;
  g->pos_y = pos_y;
  g->start0 = 0;
  g->start1 = 0;
}

int oracle_is_inside(const oracle_geom* g, double x, double y) {
  return within_map(x, y, g->len_x, g->len_y, g->pos_x, g->pos_y) ? 1 : 0;
}

int oracle_index_from_position(const oracle_geom* g, double x, double y, int* r, int* c) {
  int rr = 0, cc = 0;
  const bool ok = geom_index(g, x, y, rr, cc);
  if (ok) {
    *r = rr;
    *c = cc;
  }
  return ok ? 1 : 0;
}

int oracle_position_from_index(const oracle_geom* g, int r, int c, double* x, double* y) {
  return position_from_index(r, c, g->len_x, g->len_y, g->pos_x, g->pos_y, g->res, g->rows, g->cols, g->start0,
                             g->start1, *x, *y)
             ? 1
             : 0;
}

void oracle_index_shift_from_position_shift(double dx, double dy, double res, int* s0, int* s1) {
  /* GridMapMath.cpp:170-184 */
  const double t[2] = {dx / res, dy / res};
  int v[2];
  for (int i = 0; i < 2; i++) v[i] = static_cast<int>(t[i] + 0.5 * (t[i] > 0 ? 1 : -1));
  *s0 = -v[0];
  *s1 = -v[1];
}

int oracle_line_cells(const oracle_geom* g, double sx, double sy, double ex, double ey, int* rc, int cap) {
  Line L = make_line(g, V2{sx, sy}, V2{ex, ey});
  const int n = L.n;
  for (int i = 0; i < n; i++) {
    if (i < cap) {
      rc[2 * i] = L.r;
      rc[2 * i + 1] = L.c;
    }
    line_step(L);
  }
  return n;
}

long long oracle_himm_update(const oracle_geom* g, float* layer, const oracle_sample* s, int n, double* bbox) {
  HoistedLayer h{layer, g->rows};
  return himm_update_impl(g, h, s, n, bbox);
}

long long oracle_himm_update_as_written(const oracle_geom* g, float* layer, const oracle_sample* s, int n,
                                        double* bbox) {
  KeyedLayers k;
  k.type_name = "laser";
  k.data["master"] = HoistedLayer{nullptr, g->rows};
  k.data["range"] = HoistedLayer{nullptr, g->rows};
  k.data["laser"] = HoistedLayer{layer, g->rows};
  return himm_update_impl(g, k, s, n, bbox);
}

void oracle_himm_count(const oracle_geom* g, const oracle_sample* s, int n, long long* visits, long long* marks) {
  long long v = 0, m = 0;
  for (int k = 0; k < n; k++) {
    Line L = make_line(g, V2{s[k].sx, s[k].sy}, V2{s[k].ex, s[k].ey});
    v += L.n;
    int er, ec;
    if (!s[k].clear_end && geom_index(g, s[k].ex, s[k].ey, er, ec)) m++;
  }
  *visits = v;
  *marks = m;
}

int oracle_submap_info(const oracle_geom* g, double cx, double cy, double lx, double ly, int* tl_r, int* tl_c,
                       int* size_r, int* size_c, double* sub_px, double* sub_py, double* sub_lx, double* sub_ly) {
  /* GridMapMath.cpp:246-296.  transform = (-I).cast<double>(). */
  /* topLeftPosition = requested - transform * 0.5 * requestedLength  (:262) */
  double tlx = cx - ((-1.0 * 0.5) * lx + (-0.0 * 0.5) * ly);
  double tly = cy - ((-0.0 * 0.5) * lx + (-1.0 * 0.5) * ly);
  limit_position_to_range(tlx, tly, g->len_x, g->len_y, g->pos_x, g->pos_y);
  int tr, tc;
  if (!geom_index(g, tlx, tly, tr, tc)) return 0;
  int utr = tr, utc = tc;
  index_from_buffer_index(g, utr, utc);

  double brx = cx + ((-1.0 * 0.5) * lx + (-0.0 * 0.5) * ly);
  double bry = cy + ((-0.0 * 0.5) * lx + (-1.0 * 0.5) * ly);
  limit_position_to_range(brx, bry, g->len_x, g->len_y, g->pos_x, g->pos_y);
  int br, bc;
  if (!geom_index(g, brx, bry, br, bc)) return 0;
  index_from_buffer_index(g, br, bc);

  /* top-left corner of the submap (:276-279) */
  double cxp, cyp;
  if (!position_from_index(tr, tc, g->len_x, g->len_y, g->pos_x, g->pos_y, g->res, g->rows, g->cols, g->start0,
                           g->start1, cxp, cyp))
    return 0;
  const double half = 0.5 * g->res;
  cxp -= (-1.0 * half + -0.0 * half);
  cyp -= (-0.0 * half + -1.0 * half);

  const int sr = br - utr + 1, sc = bc - utc + 1; /* :282 */
  const double slx = static_cast<double>(sr) * g->res, sly = static_cast<double>(sc) * g->res; /* :285 */
  const double spx = cxp - 0.5 * slx, spy = cyp - 0.5 * sly;                                     /* :288-290 */
  /* :293-294 requested index in submap must resolve */
  int qr, qc;
  if (!index_from_position(cx, cy, slx, sly, spx, spy, g->res, sr, sc, 0, 0, qr, qc)) return 0;
  *tl_r = tr;
  *tl_c = tc;
  *size_r = sr;
  *size_c = sc;
  *sub_px = spx;
  *sub_py = spy;
  *sub_lx = slx;
  *sub_ly = sly;
  return 1;
}

int oracle_get_submap(const oracle_geom* g, const float* layer, double cx, double cy, double lx, double ly,
                      float* out, int out_cap, int* size_r, int* size_c) {
  int tr, tc, sr, sc;
  double spx, spy, slx, sly;
  if (!oracle_submap_info(g, cx, cy, lx, ly, &tr, &tc, &sr, &sc, &spx, &spy, &slx, &sly)) return 0;
  /* submap.setGeometry(SubmapGeometry) re-derives size = round(length / res) (GridMap.cpp:57-59,316). */
  const int rr = static_cast<int>(round(slx / g->res)), cc = static_cast<int>(round(sly / g->res));
  if (rr != sr || cc != sc) return 0;
  if (sr * sc > out_cap) return 0;
  /* GridMap.cpp:320-336 + getBufferRegionsForSubmap (GridMapMath.cpp:306-412): the <=4 quadrant block copies
   * are equivalent to out(i,j) = layer(wrap(unwrapped_tl + (i,j))). */
  int utr = tr, utc = tc;
  index_from_buffer_index(g, utr, utc);
  if (utr + sr > g->rows || utc + sc > g->cols) return 0; /* :312 */
  for (int j = 0; j < sc; j++)
    for (int i = 0; i < sr; i++) {
      int b0 = utr + i, b1 = utc + j;
      buffer_index_from_index(g, b0, b1);
      out[static_cast<size_t>(j) * sr + i] = layer[static_cast<size_t>(b1) * g->rows + b0];
    }
  *size_r = sr;
  *size_c = sc;
  return 1;
}

void oracle_ranges_from_submap(const oracle_geom* g, const float* master, double rx, double ry, double yaw,
                               double submap_len, double* ranges) {
  /* steerer.cpp:149-150 */
  for (unsigned i = 0; i < 361; i++) ranges[2 * i] = 5000.0;

  /* steerer.cpp:158 -> map_provider.cpp:93-100 -> GridMap.cpp:294-339 */
  int tr, tc, sr, sc;
  double spx, spy, slx, sly;
  if (!oracle_submap_info(g, rx, ry, submap_len, submap_len, &tr, &tc, &sr, &sc, &spx, &spy, &slx, &sly))
    return; /* no submap: reference would iterate an empty map; defined as "no obstacles" */
  std::vector<float> sub(static_cast<size_t>(sr) * sc);
  int qr, qc;
  if (!oracle_get_submap(g, master, rx, ry, submap_len, submap_len, sub.data(), sr * sc, &qr, &qc)) return;
  /* submap.setGeometry: length = size*res (GridMap.cpp:65), startIndex = 0 (:317) */
  const double L0 = static_cast<double>(sr) * g->res, L1 = static_cast<double>(sc) * g->res;

  /* steerer.cpp:161: GridMapIterator = linear index over the column-major buffer (GridMapIterator.cpp:47-50). */
  const size_t lin_size = static_cast<size_t>(sr) * sc;
  for (size_t lin = 0; lin < lin_size; lin++) {
    const int i0 = static_cast<int>(lin) % sr, i1 = static_cast<int>(lin) / sr;
    const float value = sub[static_cast<size_t>(i1) * sr + i0];
    if (std::isnan(value)) continue; /* :164 */
    if (value <= 3) continue;        /* :167 */
    double px, py;
    position_from_index(i0, i1, L0, L1, spx, spy, g->res, sr, sc, 0, 0, px, py); /* :171 */
    const double angle = atan2(py - ry, px - rx);                                 /* :173 */
    /* :175 angles::to_degrees(angles::normalize_angle_positive(angle - yaw + 3.14/2)) */
    const double a = angle - yaw + 3.14 / 2;
    const double np = fmod(fmod(a, 2.0 * M_PI) + 2.0 * M_PI, 2.0 * M_PI);
    const double deg = np * 180.0 / M_PI;
    if (deg > 180) continue; /* :176 */
    const int fl = static_cast<int>(std::floor(deg));
    const int ce = static_cast<int>(std::ceil(deg));
    const double dx = rx - px, dy = ry - py;
    const double distance = std::sqrt(dx * dx + dy * dy) * 1000.0; /* :182 */
    if (ranges[2 * (fl * 2)] > distance) ranges[2 * (fl * 2)] = distance; /* :184 */
    if (ranges[2 * (ce * 2)] > distance) ranges[2 * (ce * 2)] = distance; /* :187 */
  }
}

int oracle_move(oracle_geom* g, float** layers, int nlayers, double x, double y) {
  /* GridMap.cpp:346-412 */
  int shift[2];
  oracle_index_shift_from_position_shift(x - g->pos_x, y - g->pos_y, g->res, &shift[0], &shift[1]);
  /* getPositionShiftFromIndexShift (GridMapMath.cpp:186-192): (-I * shift).cast<double>() * res */
  const double ax = static_cast<double>(-shift[0]) * g->res, ay = static_cast<double>(-shift[1]) * g->res;
  const int size[2] = {g->rows, g->cols};
  int start[2] = {g->start0, g->start1};
  const float nanv = std::numeric_limits<float>::quiet_NaN();
  auto clear_rows = [&](int index, int n) {
    for (int l = 0; l < nlayers; l++)
      for (int c = 0; c < g->cols; c++)
        for (int r = index; r < index + n; r++) layers[l][static_cast<size_t>(c) * g->rows + r] = nanv;
  };
  auto clear_cols = [&](int index, int n) {
    for (int l = 0; l < nlayers; l++)
      for (int c = index; c < index + n; c++)
        for (int r = 0; r < g->rows; r++) layers[l][static_cast<size_t>(c) * g->rows + r] = nanv;
  };
  for (int i = 0; i < 2; i++) {
    if (shift[i] == 0) continue;
    if (std::abs(shift[i]) >= size[i]) {
      clear_rows(0, g->rows); /* clearAll */
    } else {
      const int sign = (shift[i] > 0 ? 1 : -1);
      const int start_index = start[i] - (sign < 0 ? 1 : 0);
      const int end_index = start_index - sign + shift[i];
      const int n_cells = std::abs(shift[i]);
      int index = (sign > 0 ? start_index : end_index);
      map_index_within_range(index, size[i]);
      if (index + n_cells <= size[i]) {
        if (i == 0) clear_rows(index, n_cells);
        else clear_cols(index, n_cells);
      } else {
        const int first_n = size[i] - index;
        if (i == 0) clear_rows(index, first_n);
        else clear_cols(index, first_n);
        const int second_n = n_cells - first_n;
        if (i == 0) clear_rows(0, second_n);
        else clear_cols(0, second_n);
      }
    }
  }
  g->start0 += shift[0];
  g->start1 += shift[1];
  map_index_within_range(g->start0, g->rows);
  map_index_within_range(g->start1, g->cols);
  g->pos_x += ax;
  g->pos_y += ay;
  return (shift[0] != 0 || shift[1] != 0) ? 1 : 0;
}

void oracle_to_occupancy(const oracle_geom* g, const float* layer, float data_min, float data_max,
                         signed char* out) {
  /* GridMapRosConverter.cpp:251-287 */
  const size_t n_cells = static_cast<size_t>(g->rows) * g->cols;
  const float cell_min = 0, cell_max = 100, cell_range = cell_max - cell_min;
  for (size_t lin = 0; lin < n_cells; lin++) {
    const int b0 = static_cast<int>(lin) % g->rows, b1 = static_cast<int>(lin) / g->rows;
    float value = (layer[static_cast<size_t>(b1) * g->rows + b0] - data_min) / (data_max - data_min);
    if (std::isnan(value) || (value < 0))
      value = -1;
    else
      value = cell_min + std::min(std::max(0.0f, value), 1.0f) * cell_range;
    int u0 = b0, u1 = b1;
    index_from_buffer_index(g, u0, u1);
    const size_t index = static_cast<size_t>(u1) * g->rows + u0;
    out[n_cells - index - 1] = static_cast<signed char>(value);
  }
}

int oracle_if_blocked(const oracle_geom* g, const float* master, double x, double y, double radius) {
  /* CircleIterator::findSubmapParameters (CircleIterator.cpp:99-113) */
  const double radius_square = pow(radius, 2);
  double tlx = x + radius, tly = y + radius, brx = x - radius, bry = y - radius;
  limit_position_to_range(tlx, tly, g->len_x, g->len_y, g->pos_x, g->pos_y);
  limit_position_to_range(brx, bry, g->len_x, g->len_y, g->pos_x, g->pos_y);
  int sr, sc, er, ec;
  if (!geom_index(g, tlx, tly, sr, sc) || !geom_index(g, brx, bry, er, ec)) return 0;
  /* getSubmapSizeFromCornerIndeces (GridMapMath.cpp:298-304): unwrapped corner difference + 1 */
  int usr = sr, usc = sc, uer = er, uec = ec;
  index_from_buffer_index(g, usr, usc);
  index_from_buffer_index(g, uer, uec);
  const int nr = uer - usr + 1, nc = uec - usc + 1;
  /* SubmapIterator walks the nr x nc block; ifBlocked returns at the first hit, i.e. "any" */
  for (int i = 0; i < nr; i++)
    for (int j = 0; j < nc; j++) {
      int b0 = usr + i, b1 = usc + j;
      buffer_index_from_index(g, b0, b1);
      double px, py;
      if (!position_from_index(b0, b1, g->len_x, g->len_y, g->pos_x, g->pos_y, g->res, g->rows, g->cols, g->start0,
                               g->start1, px, py))
        continue;
      const double dx = px - x, dy = py - y;
      if (!(dx * dx + dy * dy <= radius_square)) continue; /* CircleIterator::isInside (:90-97) */
      const float v = master[static_cast<size_t>(b1) * g->rows + b0];
      if (std::isnan(v)) continue;
      if (v > 0.0) return 1;
    }
  return 0;
}

void oracle_goal_from_pose(double rx, double ry, double yaw, double tx, double ty, float* desired_angle,
                           float* desired_dist) {
  /* steerer.cpp:228-256: float deltaX/deltaY/desiredDist; RAD2DEG(normalize_angle_positive()).  steerer.cpp includes
   * <math.h>, whose C++ form exposes the float overloads: hypot(float, float) and atan2(float, float) are hypotf and
   * atan2f there (checked against the reference's own Steerer::update, tests/test_reference_pin.py). */
  const float dX = static_cast<float>((tx - rx) * 1000.0);
  const float dY = static_cast<float>((ty - ry) * 1000.0);
  *desired_dist = std::hypot(dX, dY);
  const double a = std::atan2(dY, dX) - yaw + M_PI / 2;
  const double np = fmod(fmod(a, 2.0 * M_PI) + 2.0 * M_PI, 2.0 * M_PI);
  *desired_angle = static_cast<float>(np * 180.0 / M_PI);
}

/* ---------------------------------------------------------------------------------------------------------------
 * LaserScan -> RangeSamples (intake side of LaserMapUpdater::bufferIncomingMsg, laser_map_updater.cpp:38-144).
 * laser_geometry and tf are not in the reference tree, so this restates the SPECIFICATION written down in
 * ros_navigation_b200/csrc/scan_project.h (an independent implementation of the same arithmetic, operation by
 * operation) - parity with the reference is unpinned at this boundary.
 * ------------------------------------------------------------------------------------------------------------- */
void oracle_sincos(double x, double* s_out, double* c_out) {
  /* reduction: n = nearest integer to x * 2/pi, r = x - n * pi/2 with pi/2 split into three 33-bit pieces */
  static const double kInvPio2 = 6.36619772367581382433e-01;
  static const double kPio2[3] = {1.57079632673412561417e+00, 6.07710050630396597660e-11, 2.02226624871116645580e-21};
  static const double kS[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                               2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10};
  static const double kC[6] = {4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                               -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11};
  volatile double t = x * kInvPio2 + 6755399441055744.0; /* round to nearest integer (ties to even) */
  const double fn = t - 6755399441055744.0;
  double r = x - fn * kPio2[0];
  r = r - fn * kPio2[1];
  r = r - fn * kPio2[2];
  const double z = r * r;
  double ps = kS[5];
  for (int i = 4; i >= 0; i--) ps = kS[i] + z * ps; /* Horner, innermost coefficient first */
  double pc = kC[5];
  for (int i = 4; i >= 0; i--) pc = kC[i] + z * pc;
  const double sin_r = r + r * (z * ps);
  const double cos_r = 1.0 - (0.5 * z - (z * z) * pc);
  const long long n = static_cast<long long>(fn);
  switch (static_cast<int>(n & 3)) {
    case 0: *s_out = sin_r; *c_out = cos_r; break;
    case 1: *s_out = cos_r; *c_out = -sin_r; break;
    case 2: *s_out = -sin_r; *c_out = -cos_r; break;
    default: *s_out = -cos_r; *c_out = sin_r; break;
  }
}

int oracle_scan_select(float angle_increment, int n_ranges, int decimate, int* sel, float* increment_used) {
  /* simplifyLaserScan (laser_map_updater.cpp:114-144) is applied when angle_increment < 0.017 (:74) */
  int n = 0;
  *increment_used = angle_increment;
  if (decimate && angle_increment < 0.017 && n_ranges > 0) {
    sel[n++] = 0; /* convertedLaserScan.ranges.push_back(msg->ranges[0]) */
    float increment = 0.0;
    for (int i = 0; i < n_ranges; i++) {
      increment += angle_increment;
      if (increment >= 0.017) {
        *increment_used = increment;
        increment = 0.0;
        sel[n++] = i;
      }
    }
    return n;
  }
  for (int i = 0; i < n_ranges; i++) sel[n++] = i;
  return n;
}

int oracle_project_scan(float angle_min, float increment_used, float range_min, float range_max, const int* sel,
                        int n_used, int decimated, const float* ranges, double x0, double y0, double yaw,
                        oracle_sample* out) {
  double sin_yaw, cos_yaw;
  oracle_sincos(yaw, &sin_yaw, &cos_yaw);
  int n = 0;
  for (int j = 0; j < n_used; j++) {
    const float range = ranges[sel[j]];
    if (!(range < range_max && range >= range_min)) continue; /* laser_geometry keeps range_min <= r < range_max */
    const double angle = static_cast<double>(angle_min) + static_cast<double>(j) * static_cast<double>(increment_used);
    double sa, ca;
    oracle_sincos(angle, &sa, &ca);
    /* point in the sensor frame, stored as float32 in the cloud */
    const float local_x = static_cast<float>(static_cast<double>(range) * ca);
    const float local_y = static_cast<float>(static_cast<double>(range) * sa);
    /* rigid transform into the map frame in double, stored back as float32 */
    const float map_x = static_cast<float>((cos_yaw * local_x - sin_yaw * local_y) + x0);
    const float map_y = static_cast<float>((sin_yaw * local_x + cos_yaw * local_y) + y0);
    out[n].sx = x0; /* getLaserOriginOnGlobal */
    out[n].sy = y0;
    out[n].ex = map_x; /* Position(*itX, *itY) */
    out[n].ey = map_y;
    /* laser_map_updater.cpp:62-66: `index` is the point's position in the projected (possibly thinned) scan, but it is
     * looked up in the ORIGINAL msg->ranges.  Not thinned: that reading passed the filter above, flag false. */
    const float looked_up = decimated ? ranges[j] : range;
    out[n].clear_end = (std::isinf(looked_up) || looked_up == range_max) ? 1 : 0;
    out[n].pad_ = 0;
    n++;
  }
  return n;
}

void oracle_compose_master(const float* range, const float* laser, float* master, long long n) {
  /* The compose MapProvider::composeMasterMapFromLayerdMap carries commented out (map_provider.cpp:218-220):
   *   (range.isNaN() && !laser.isNaN()).select(0, range) + (laser.isNaN() && !range.isNaN()).select(0, laser) */
  for (long long i = 0; i < n; i++) {
    const bool range_nan = std::isnan(range[i]), laser_nan = std::isnan(laser[i]);
    const float first = (range_nan && !laser_nan) ? 0.0f : range[i];
    const float second = (laser_nan && !range_nan) ? 0.0f : laser[i];
    master[i] = first + second;
  }
}

} /* extern "C" */
