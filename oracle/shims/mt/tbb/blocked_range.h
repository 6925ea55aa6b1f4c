// Oracle shim: see shims/serial/tbb/blocked_range.h.
#pragma once
#include "../../serial/tbb/blocked_range.h"
