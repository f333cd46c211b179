// oracle/shim — empty stand-in: include/Scancontext.h names this header but Scancontext.cpp uses nothing from it.  TEST INFRASTRUCTURE.
#pragma once
