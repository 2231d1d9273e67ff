#ifndef tchem_intcoord_hpp
#define tchem_intcoord_hpp
#include <tchem/IC/IntCoordSet.hpp>
#endif
