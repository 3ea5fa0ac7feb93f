// dstar_decoder.hpp — drop-in for the reference header of the same name (reference include/dstar_decoder.hpp); the classes live in
// digiham_b200_modules.hpp and run on the GPU through libdigiham_b200 (C ABI: digiham_b200.h).
#pragma once
#include "digiham_b200_modules.hpp"
