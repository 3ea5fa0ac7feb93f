// zero_heap.hpp — force-included (-include) into every reference translation unit of the oracle build.
// TEST INFRASTRUCTURE ONLY.
//
// The reference reads heap memory it never initialised (src/rrc_filter/rrc_filter.cpp:9,
// src/gfsk_demodulator/gfsk_demodulator.cpp:11, src/dmr_decoder/embedded.cpp:13,
// src/dmr_decoder/talkeralias.cpp:17).  The parity contract of this repository defines that memory as
// zero, so malloc is mapped to calloc.  The system headers are pulled in first so the
// function-like macro cannot touch their declarations.
#pragma once
#include <stdlib.h>
#include <malloc.h>
#include <string.h>
#include <stdint.h>
#ifdef __cplusplus
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <new>
#endif
#define malloc(n) calloc(1, (n))
