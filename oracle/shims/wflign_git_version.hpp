/* TEST INFRASTRUCTURE ONLY: stand-in for the header the reference's CMake generates
 * (src/common/wflign/CMakeLists.txt); only needed to compile the reference sources in place. */
#pragma once
#define WFLIGN_GIT_VERSION "oracle-ref"
