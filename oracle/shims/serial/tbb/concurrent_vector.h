// Oracle shim: concurrent_vector = std::vector (arrival order == index order when serial).
#pragma once
#include <vector>
namespace tbb {
template <class T> using concurrent_vector = std::vector<T>;
}  // namespace tbb
