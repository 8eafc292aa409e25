// TEST INFRASTRUCTURE ONLY (oracle/): src/ReadData.cpp:8 includes this header but uses nothing from it.
#pragma once
