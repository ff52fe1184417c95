// Internal handle layout shared by the host (capi.cpp) and device (device.cu) halves.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/madipm_b200.h"
#include "symbolic.hpp"

namespace madipm {
struct DeviceState;   // device.cu
}

struct madipm_handle {
    madipm_config cfg;
    std::string error;
    madipm::KktPattern pat;
    madipm::Symbolic sym;
    bool has_pattern = false;
    bool analyzed = false;
    int32_t nlb = 0, nub = 0;
    std::vector<int32_t> ind_lb, ind_ub;   // 0-based
    madipm::DeviceState* dev = nullptr;
    int64_t launches = 0;
    const char* last_launch = "";
    std::string describe;                  // buffer of madipm_debug_describe          // name of the kernel launched last (diagnostics of a synchronisation timeout)
};

namespace madipm {
// device.cu entry points used by capi.cpp
int32_t device_create(madipm_handle* h);      // allocate + upload after analyze (lazy)
void device_destroy(madipm_handle* h);
}
