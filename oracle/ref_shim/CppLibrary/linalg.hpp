// Stand-in header: the reference only #includes <CppLibrary/linalg.hpp> on the hot
// path (source/IC/IntCoord.cpp:1) without using any symbol from it.
#ifndef CL_SHIM_LINALG_HPP
#define CL_SHIM_LINALG_HPP
#include <CppLibrary/utility.hpp>
#endif
