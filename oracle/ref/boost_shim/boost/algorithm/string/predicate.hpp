// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the one Boost string predicate the
// reference's src/DirectoryUtils.cpp:79 uses (see boost_shim/boost/filesystem.hpp).
#pragma once
#include <cstring>
#include <string>
namespace boost {
namespace algorithm {
inline bool ends_with(const std::string &s, const char *suffix) {
    const size_t m = std::strlen(suffix);
    return s.size() >= m && s.compare(s.size() - m, m, suffix) == 0;
}
}  // namespace algorithm
}  // namespace boost
