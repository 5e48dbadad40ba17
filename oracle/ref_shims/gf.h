// Case shim: the reference includes "gf.h" (NBLDPC.h:6) but ships GF.h.
#pragma once
#include "GF.h"
