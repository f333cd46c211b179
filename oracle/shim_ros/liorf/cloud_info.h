// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h): the message /root/reference/msg/cloud_info.msg generates.
#pragma once
#include <sensor_msgs/PointCloud2.h>
namespace liorf {
struct cloud_info {
    std_msgs::Header header;
    std::vector<int32_t> startRingIndex, endRingIndex, pointColInd; std::vector<float> pointRange;
    int64_t imuAvailable = 0, odomAvailable = 0;
    float imuRollInit = 0, imuPitchInit = 0, imuYawInit = 0;
    float initialGuessX = 0, initialGuessY = 0, initialGuessZ = 0, initialGuessRoll = 0, initialGuessPitch = 0, initialGuessYaw = 0;
    sensor_msgs::PointCloud2 cloud_deskewed, cloud_corner, cloud_surface, key_frame_cloud, key_frame_color, key_frame_poses, key_frame_map;
    typedef std::shared_ptr<cloud_info const> ConstPtr;
};
typedef std::shared_ptr<cloud_info const> cloud_infoConstPtr;
}
