/*
 * vfh_tables.cpp -- see vfh_tables.h.  Pure host C++ (compiled by g++, not nvcc, so that <math.h> overloads
 * resolve exactly as in the reference build).
 *
 * Provenance: the tables must be bit-identical to the ones the reference's VFH::Init builds, so the arithmetic
 * expressions below (the quadrant ladder of cell directions, `atanf(..) * (360.0 / 6.28)`, the magnitude polynomial,
 * the sector-overlap test) follow move_control/src/vfh.cpp:237-416 expression by expression; the layout, the bitmask
 * form of the sector lists and everything else are this project's.  The reference file derives from the Player / Orca
 * VFH+ driver (Orca-Components, copyright 2004) and is distributed under the GNU General Public License, version 2
 * or later; whoever ships this file with a product inherits that obligation for it.
 */
#include "vfh_tables.h"

#include <math.h>
#include <stdio.h>

namespace b200nav {

int vfh_get_max_turnrate(int t0, int t1, int speed) {
  /* vfh.cpp:130-138 */
  int val = (t0 - (int)(speed * (t0 - t1) / 1000.0));
  if (val < 0) val = 0;
  return val;
}

static int safety_dist(float s0, float s1, int speed) {
  /* vfh.cpp:195-205 */
  int val = (int)(s0 + (int)(speed * (s1 - s0) / 1000.0));
  if (val < 0) val = 0;
  return val;
}

void vfh_build_min_turning_radius(VfhTables& t, const b200nav_vfh_params& p, int max_speed) {
  /* vfh.cpp:144-166 */
  VfhConst& c = t.c;
  c.current_max_speed = (max_speed < c.max_speed) ? max_speed : c.max_speed;
  t.min_turning_radius.assign((size_t)(c.current_max_speed < 0 ? 0 : c.current_max_speed) + 1, 0);
  for (int x = 0; x <= c.current_max_speed; x++) {
    const double dx = (double)x / 1e6;
    const double dtheta = ((M_PI / 180) * (double)(vfh_get_max_turnrate(c.max_turnrate_0ms, c.max_turnrate_1ms, x))) / 1000.0;
    t.min_turning_radius[x] = (int)(((dx / tan(dtheta)) * 1000.0) * p.min_turn_radius_safety_factor);
  }
}

/* Angular distance helper of the sector test (vfh.cpp:337-375): signed gap from `sector` to `dir`. */
static inline float sector_to_dir(float sector, float dir) {
  if ((sector - dir) > 180) return dir - (sector - 360);
  if ((dir - sector) > 180) return sector - (dir + 360);
  return dir - sector;
}

int vfh_build_tables(const b200nav_vfh_params& p, VfhTables& t, char* err, int errlen) {
  if (!(p.cell_size > 0) || p.window_diameter < 2 || p.window_diameter > 1024 || p.sector_angle < 1 ||
      p.sector_angle > 180 || p.max_speed < 1 || p.max_speed > 100000) {
    snprintf(err, errlen, "vfh params out of range (cell_size>0, 2<=window<=1024, 1<=sector_angle<=180, 1<=max_speed<=100000)");
    return B200NAV_EINVAL;
  }
  VfhConst& c = t.c;
  c.robot_radius = (float)p.robot_radius;
  c.cell_width = (float)p.cell_size;
  c.window = p.window_diameter;
  c.sector_angle = p.sector_angle;
  c.safety_dist_0ms = (float)p.safety_dist_0ms;
  c.safety_dist_1ms = (float)p.safety_dist_1ms;
  c.max_speed = p.max_speed;
  c.current_max_speed = p.max_speed;
  c.max_speed_narrow = p.max_speed_narrow_opening;
  c.max_speed_wide = p.max_speed_wide_opening;
  c.max_acceleration = p.max_acceleration;
  c.max_turnrate_0ms = p.max_turnrate_0ms;
  c.max_turnrate_1ms = p.max_turnrate_1ms;
  c.bin_low_0ms = (float)p.free_space_cutoff_0ms;
  c.bin_high_0ms = (float)p.obs_cutoff_0ms;
  c.bin_low_1ms = (float)p.free_space_cutoff_1ms;
  c.bin_high_1ms = (float)p.obs_cutoff_1ms;
  c.u1 = (float)p.weight_desired_dir;
  c.u2 = (float)p.weight_current_dir;
  c.submap_length = p.submap_length;
  c.occupied_threshold = (float)p.occupied_threshold;
  /* vfh.cpp:96-109 */
  c.num_tables = (c.safety_dist_0ms == c.safety_dist_1ms) ? 1 : 20;
  /* vfh.cpp:247-249 */
  c.center = (int)floor(c.window / 2.0);
  c.hist_size = (int)rint(360.0 / c.sector_angle);
  c.nwords = (c.hist_size + 31) / 32;
  c.front_rows = (int)ceil(c.window / 2.0);
  if (360 % c.sector_angle != 0) {
    snprintf(err, errlen, "sector_angle must divide 360");
    return B200NAV_EINVAL;
  }

  const int W = c.window, C = c.center, T = c.num_tables, NW = c.nwords, FR = c.front_rows;
  t.dir_xy.assign((size_t)W * W, 0.f);
  t.dist_xy.assign((size_t)W * W, 0.f);
  t.base_xy.assign((size_t)W * W, 0.f);
  t.masks_xy.assign((size_t)T * W * W * NW, 0u);

  for (int x = 0; x < W; x++) {
    for (int y = 0; y < W; y++) {
      /* vfh.cpp:271-273 */
      const float dist = sqrt(pow((C - x), 2) + pow((C - y), 2)) * c.cell_width;
      const float base = 15 * pow((3000.0 - dist), 4) / 100000000.0;
      /* vfh.cpp:276-307: direction in degrees, 0 = right, 90 = ahead, with the reference's 360/6.28 scale */
      float dir = 0;
      if (x < C) {
        if (y < C) {
          dir = atanf((float)(C - y) / (float)(C - x));
          dir *= (360.0 / 6.28);
          dir = 180.0 - dir;
        } else if (y == C) {
          dir = 180.0;
        } else {
          dir = atanf((float)(y - C) / (float)(C - x));
          dir *= (360.0 / 6.28);
          dir = 180.0 + dir;
        }
      } else if (x == C) {
        if (y < C) dir = 90.0;
        else if (y == C) dir = -1.0;
        else dir = 270.0;
      } else {
        if (y < C) {
          dir = atanf((float)(C - y) / (float)(x - C));
          dir *= (360.0 / 6.28);
        } else if (y == C) {
          dir = 0.0;
        } else {
          dir = atanf((float)(y - C) / (float)(x - C));
          dir *= (360.0 / 6.28);
          dir = 360.0 - dir;
        }
      }
      t.dir_xy[(size_t)x * W + y] = dir;
      t.dist_xy[(size_t)x * W + y] = dist;
      t.base_xy[(size_t)x * W + y] = base;

      for (int tab = 0; tab < T; tab++) {
        /* vfh.cpp:314-330 */
        const int max_speed_this_table = (int)(((float)(tab + 1) / (float)T) * (float)c.max_speed);
        float enlarge;
        if (dist > 0) {
          const float r = c.robot_radius + safety_dist(c.safety_dist_0ms, c.safety_dist_1ms, max_speed_this_table);
          enlarge = (float)asinf(r / dist) * (180 / M_PI);
        } else {
          enlarge = 0;
        }
        const float plus_dir = dir + enlarge, neg_dir = dir - enlarge;
        uint32_t* m = &t.masks_xy[(((size_t)tab * W + x) * W + y) * NW];
        /* vfh.cpp:337-406: sector i covers [i*a, (i+1)*a]; affected if either edge of the enlarged obstacle
         * falls in it or it lies inside the enlarged obstacle. */
        for (int i = 0; i < (360 / c.sector_angle); i++) {
          const float plus_sector = (i + 1) * (float)c.sector_angle;
          const float neg_sector = i * (float)c.sector_angle;
          const float ns_nd = sector_to_dir(neg_sector, neg_dir);
          const float ps_nd = sector_to_dir(plus_sector, neg_dir);
          const float ps_pd = sector_to_dir(plus_sector, plus_dir);
          const float ns_pd = sector_to_dir(neg_sector, plus_dir);
          const bool neg_dir_bw = (ns_nd >= 0) && (ps_nd <= 0);
          const bool plus_dir_bw = ((ns_pd >= 0) && (ps_pd <= 0)) || ((ps_nd <= 0) && (ps_pd >= 0));
          const bool around = (ns_nd <= 0) && (ns_pd >= 0);
          if (plus_dir_bw || neg_dir_bw || around) m[i >> 5] |= 1u << (i & 31);
        }
      }
    }
  }

  /* device layout: front rows only */
  const size_t nf = (size_t)FR * W;
  t.dir.resize(nf);
  t.dist.resize(nf);
  t.base.resize(nf);
  t.thr.resize(nf);
  t.kidx.resize(nf);
  t.masks.assign((size_t)T * nf * NW, 0u);
  for (int y = 0; y < FR; y++)
    for (int x = 0; x < W; x++) {
      const size_t f = (size_t)y * W + x, xy = (size_t)x * W + y;
      t.dir[f] = t.dir_xy[xy];
      t.dist[f] = t.dist_xy[xy];
      t.base[f] = t.base_xy[xy];
      t.thr[f] = t.dist_xy[xy] + c.cell_width / 2.0;              /* vfh.cpp:1017 */
      t.kidx[f] = (int16_t)(int)rint(t.dir_xy[xy] * 2.0);         /* vfh.cpp:1018 */
      for (int tab = 0; tab < T; tab++)
        for (int k = 0; k < NW; k++)
          t.masks[((size_t)tab * nf + f) * NW + k] = t.masks_xy[(((size_t)tab * W + x) * W + y) * NW + k];
    }
  vfh_build_min_turning_radius(t, p, c.max_speed);
  return B200NAV_OK;
}

}  // namespace b200nav
