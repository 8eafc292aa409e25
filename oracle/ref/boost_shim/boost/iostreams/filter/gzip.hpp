// TEST INFRASTRUCTURE ONLY (oracle/): see ../filtering_streambuf.hpp (gzip_decompressor lives there).
#pragma once
#include "../filtering_streambuf.hpp"
