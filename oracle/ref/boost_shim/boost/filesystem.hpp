// TEST INFRASTRUCTURE ONLY (oracle/).  The reference's consensus stage (src/Consensus.cpp,
// src/DirectoryUtils.cpp, include/DirectoryUtils.h) includes <boost/filesystem.hpp>; Boost is a
// network download of the reference's build (boost-cmake) and is not in this image.  The few
// names those files use (path, directory_iterator, is_regular_file, remove, remove_all,
// create_directory, file_size, system::error_code) exist with the same meaning in C++17
// <filesystem>, so this stand-in header aliases them and lets the UNMODIFIED reference sources
// compile where they lie (oracle/Makefile, target consensus_dropin).
#pragma once
// <boost/filesystem.hpp> pulls these in transitively; DirectoryUtils.cpp relies on it
#include <algorithm>
#include <filesystem>
#include <fstream>
#include <system_error>
namespace boost {
namespace filesystem = std::filesystem;
namespace system {
using error_code = std::error_code;
}
}  // namespace boost
