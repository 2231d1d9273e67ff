#ifndef tchem_polynomial_hpp
#define tchem_polynomial_hpp
#include <tchem/polynomial/SAPSet.hpp>
#endif
