// oracle/shim_ros — TEST INFRASTRUCTURE (see gtsam/shim_gtsam.h).
#pragma once
#include <gtsam/shim_gtsam.h>
