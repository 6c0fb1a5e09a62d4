// r3d_host.h -- host-side helpers shared by the translation units behind the C ABI:
// argument validation, conversion of the ABI structs into kernel parameter blocks, error plumbing.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>

#include "r3d_b200.h"
#include "r3d_device.cuh"

namespace r3d {

// thread-local message behind r3d_last_error()
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

// validate + convert; return R3D_OK or set the error
int to_device_params(const R3dGrid* grid, GridP& g);
int to_device_params(const R3dRays* rays, RaysP& r);
int to_device_params(const R3dRenderConfig* cfg, const RaysP& r, CfgP& c);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace r3d
