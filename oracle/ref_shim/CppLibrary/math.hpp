// Stand-in header, see utility.hpp. Not needed by the hot path; kept so that
// umbrella includes resolve.
#ifndef CL_SHIM_MATH_HPP
#define CL_SHIM_MATH_HPP
#include <cstddef>
namespace CL { namespace math {
inline size_t iCombination(size_t n, size_t k) {
    if (k > n) return 0;
    if (k > n - k) k = n - k;
    size_t c = 1;
    for (size_t i = 1; i <= k; i++) c = c * (n - k + i) / i;
    return c;
}
inline double dFactorial2(int n) {
    double f = 1.0;
    for (int i = n; i > 1; i -= 2) f *= i;
    return f;
}
} }
#endif
