// Shared between the host mini-FEM layer and the device layer.
#pragma once
#include <string>

#define LPF_MAX_ORDER 10

namespace lpf {
void set_error(const std::string &s);
}
