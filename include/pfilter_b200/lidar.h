// lidar::Lidar with the reference's interface (/root/reference/include/lidar.h:9-32, src/lidar.cpp).
#pragma once
#include "../pfilter_b200.h"

namespace lidar {
class Lidar {
   public:
    Lidar() {}
    void setScanPeriod(double v) { scan_period = v; }
    void setLines(double v) { num_lines = (int)v; }
    void setVerticalAngle(double v) { vertical_angle = v; }
    void setVerticalResolution(double v) { vertical_angle_resolution = v; }
    void setMaxDistance(double v) { max_distance = v; }
    void setMinDistance(double v) { min_distance = v; }

    double max_distance = 90.0;
    double min_distance = 3.0;
    int num_lines = 64;
    double scan_period = 0.1;
    int points_per_line = 0;
    double horizontal_angle_resolution = 0;
    double horizontal_angle = 0;
    double vertical_angle_resolution = 0;
    double vertical_angle = 0;

    pf_lidar_params c_params() const { return pf_lidar_params{num_lines, min_distance, max_distance, scan_period}; }
};
}  // namespace lidar
