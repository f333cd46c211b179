// oracle/shim_ros — TEST INFRASTRUCTURE: named by include/Scancontext.h, nothing of it is used.
#pragma once
