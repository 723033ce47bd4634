/*
 * move_control/vfh.h -- drop-in replacement for the reference's VFH+ class header
 * (move_control/include/move_control/vfh.h:182-361; implementation move_control/src/vfh.cpp).  Header-only: remove
 * vfh.cpp from the catkin target and link libb200nav.so.
 *
 * move_control::VFH keeps the reference's public surface exactly - the 19-argument constructor used by
 * Steerer::initVfh (move_control/src/steerer.cpp:122-129), Init(), Update_VFH(double[361][2], int, float, float,
 * float, int&, int&), the getters / setters and the public float* Hist / OriginHist that Steerer::pubHist reads
 * (steerer.cpp:201-220) - and runs every stage of the update (vfh.cpp:480-605 and callees) in one fused sm_100a
 * kernel.  The CUDA context is the process-wide default (b200nav_dropin.hpp), not a constructor argument.
 * Added, non-breaking: Update_VFH_FromGrid(...) fuses Steerer::getRangesFromSubmap (steerer.cpp:147-191) into the
 * same kernel and reads the device copy of the map, which saves the host submap copy and the per-cell atan2.
 */
#ifndef VFH_ALGORITHM_H
#define VFH_ALGORITHM_H

#include <sys/time.h>
#include <stdio.h>
#include <vector>

#include "../b200nav_dropin.hpp"

namespace move_control {

class VFH : public b200nav::VFH
{
public:
    VFH( double cell_size,
         int window_diameter,
         int sector_angle,
         double safety_dist_0ms,
         double safety_dist_1ms,
         int max_speed,
         int max_speed_narrow_opening,
         int max_speed_wide_opening,
         int max_acceleration,
         int min_turnrate,
         int max_turnrate_0ms,
         int max_turnrate_1ms,
         double min_turn_radius_safety_factor,
         double free_space_cutoff_0ms,
         double obs_cutoff_0ms,
         double free_space_cutoff_1ms,
         double obs_cutoff_1ms,
         double weight_desired_dir,
         double weight_current_dir )
        : b200nav::VFH(b200nav::default_context(), cell_size, window_diameter, sector_angle, safety_dist_0ms,
                       safety_dist_1ms, max_speed, max_speed_narrow_opening, max_speed_wide_opening, max_acceleration,
                       min_turnrate, max_turnrate_0ms, max_turnrate_1ms, min_turn_radius_safety_factor,
                       free_space_cutoff_0ms, obs_cutoff_0ms, free_space_cutoff_1ms, obs_cutoff_1ms,
                       weight_desired_dir, weight_current_dir) {}

    // Steerer::getRangesFromSubmap + Update_VFH against the device twin of `map`.  Pass a layer that an updater
    // maintains ("laser"): it is already in HBM.  Host-composed layers ("master") are uploaded on every call.
    int Update_VFH_FromGrid( grid_map::GridMap& map, const std::string& layer, double x, double y, double yaw,
                             int current_speed, float goal_direction, float goal_distance,
                             float goal_distance_tolerance, int &chosen_speed, int &chosen_turnrate ) {
        b200nav::DeviceMap& twin = b200nav::device_map_for(map);
        if (!twin.make_readable(layer))
            return 1;
        return update_from_device_grid(twin.grid(), layer, x, y, yaw, current_speed, goal_direction, goal_distance,
                                       goal_distance_tolerance, chosen_speed, chosen_turnrate);
    }

    // Debug dump of the reference (vfh.cpp:875-978 family); the magnitudes live on the device, nothing is printed.
    void Print_Cells_Mag() {}
};

}
#endif
