// version.hpp — Digiham::version of the B200 drop-in (reference include/version.hpp:7).
#pragma once
#include "digiham_b200.h"

#include <string>

namespace Digiham {
    inline const std::string version = dh_version();
}
