// stand-in: NVTX ranges are no-ops in the emulated build
#pragma once
inline int nvtxRangePushA(const char*) { return 0; }
inline int nvtxRangePop() { return 0; }
