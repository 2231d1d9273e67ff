#ifndef tchem_polynomial_hpp
#define tchem_polynomial_hpp
#include <tchem/polynomial/SAPSet.hpp>
#include <tchem/polynomial/PolynomialSet.hpp>
#endif
