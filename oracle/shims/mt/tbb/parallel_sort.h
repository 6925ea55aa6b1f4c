// Oracle shim: parallel_sort = std::sort (the only call site sorts <= ~1e4 doubles, common.hpp:45-48).
#pragma once
#include "../../serial/tbb/parallel_sort.h"
