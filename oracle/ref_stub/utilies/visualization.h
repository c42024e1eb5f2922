// utilies/visualization.h — SHADOW of the reference header of the same name (oracle/_ref): the RViz publisher
// (tf2_ros, visualization_msgs, nav_msgs::OccupancyGrid, threads) is outside the front-end solver path and is not
// called by any file compiled into oracle/_ref; utilies/utilies.h includes it unconditionally.
#pragma once
#include "trajectory/camera_type.h"
#include "trajectory/laser_type.h"
#include "utilies/common.h"
#include "utilies/params.h"
