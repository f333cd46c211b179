// oracle/shim_ros — TEST INFRASTRUCTURE: named by include/utility.h, nothing of it is used on the path.
#pragma once
