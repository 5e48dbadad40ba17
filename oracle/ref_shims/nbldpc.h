// Case shim: the reference includes "nbldpc.h" (Comm.h:9) but ships NBLDPC.h.
#pragma once
#include "NBLDPC.h"
