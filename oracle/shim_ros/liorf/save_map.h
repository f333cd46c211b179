// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h): /root/reference/srv/save_map.srv.
#pragma once
#include <string>
namespace liorf { struct save_mapRequest { float resolution = 0; std::string destination; }; struct save_mapResponse { uint8_t success = 0; }; }
