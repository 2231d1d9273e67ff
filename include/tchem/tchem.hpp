// umbrella header of the B200 drop-in: the internal-coordinate path of Torch-Chemistry
#ifndef tchem_hpp
#define tchem_hpp
#include <tchem/intcoord.hpp>
#include <tchem/polynomial.hpp>
#include <tchem/b200.hpp>

#endif
