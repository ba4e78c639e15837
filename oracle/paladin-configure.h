/* TEST INFRASTRUCTURE ONLY (oracle/): stands in for the file CMake generates from
 * reference src/paladin-configure.h.in:32-33 when the reference is built by
 * oracle/Makefile. */
#ifndef Paladin_Configure_h
#define Paladin_Configure_h
#define PALADIN_REPO_DATE "oracle-build"
#define PALADIN_REPO_HASH "oracle-build"
#endif
